// Lean statistics phase ("v2") for the common statistics masks: the same per-image reductions as stats_tile_t
// (vu_common.cuh), written for the case the round-1 profiles showed to be issue-bound -- few classes, several raters,
// calibration histograms (configs[1]: 279 instructions per voxel, 36 % of the HBM peak).
//
// What is different:
//   * the flags are a compile-time constant, the references are uint8 with word-aligned rows, all three uncertainty types
//     are present and there is no label LUT (anything else takes the general path);
//   * the per-thread partial sums live in registers (StatAcc) for as long as the thread stays on one image, instead of
//     [slot][thread] columns of shared memory that are read, updated and written back on every tile;
//   * calibration: no NaN handling in the unrolled loop (a thread whose 12 values contain a NaN / inf takes a rolled slow
//     path), packed f32x2 arithmetic for the Platt expression of two voxels, the candidate bin / q / histogram address stay in
//     the mantissa domain with the constant parts folded into the table bases (one IMAD each), and the bias this leaves in
//     the q sums is removed when the histogram is flushed;
//   * the references of four voxels are tested as one word per rater with LOP3-friendly byte tests; the rater variance of
//     the NCC is integer arithmetic on byte-transposed words (dp4a) and is scaled once per flush.
//
// Reference semantics (unchanged): evaluation/metrics/ace.py:325-356,431-437; evaluation/metrics/ncc.py:17-27 +
// evaluation/experiment_dataloader.py:283; uncertainty_modeling/test_2D.py:873-899;
// evaluation/uncertainty_aggregation/aggregate_uncertainties.py:37-39,124-125; prediction_shape_stats.py:10-12.
#pragma once
#include "vu_common.cuh"

namespace vu {

constexpr unsigned kMagicBits = 0x4B400000u;  // bit pattern of kRoundMagic (1.5 * 2^23)
// q is accumulated as  qf_bits - bin_bits * kQBinStep  with bin_bits = kMagicBits + bin and qf_bits = kMagicBits + round(conf 2^21):
// that is q + kYBias (mod 2^32); the flush subtracts samples * kYBias again
constexpr unsigned kYBias = kMagicBits - kMagicBits * (unsigned)kQBinStep;
constexpr int kBins2 = kHistBins;  // 20 bins per type and replica (NaN / inf samples never reach the histogram)
// tiles (of four voxels per thread) a thread may go through between two flushes: the Dice counters are byte-sliced (one
// count per voxel position and byte, no popc while streaming), the histogram words hold 16-bit counts of two lanes x 8 raters
constexpr int kMaxTilesPerFlush2 = 255;

template <unsigned FL>
struct StatFlags {
    static constexpr bool sum = (FL & (VU_STAT_IMAGE_SUM | VU_STAT_NCC)) != 0;
    static constexpr bool thr = (FL & VU_STAT_THRESHOLD) != 0;
    static constexpr bool area = (FL & VU_STAT_AREA) != 0;
    static constexpr bool dice = (FL & VU_STAT_DICE) != 0;
    static constexpr bool calib = (FL & VU_STAT_CALIB) != 0;
    static constexpr bool ncc = (FL & VU_STAT_NCC) != 0;
    static constexpr bool refs = (FL & (VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC)) != 0;
};

// shared memory of the v2 statistics threads: [warp][unc][bin][REP] histogram words, then the edge tables
__host__ __device__ inline size_t stats2_smem_bytes(unsigned flags, int threads, int rep) {
    if (!(flags & VU_STAT_CALIB)) return 16;
    return (size_t)(threads / 32) * (VU_N_UNC * kBins2 * rep) * sizeof(uint2) + (size_t)VU_N_UNC * kEdgePad * sizeof(float);
}

// per-thread partial sums of the image the thread is working on
template <unsigned FL, int RMAX>
struct StatAcc {
    using F = StatFlags<FL>;
    double sum[F::sum ? 3 : 1];
    double thr[F::thr ? 3 : 1];
    unsigned thrn[F::thr ? 3 : 1];
    unsigned area, nvox;
    unsigned tp[F::dice ? RMAX : 1], ps[F::dice ? RMAX : 1], gs[F::dice ? RMAX : 1];  // byte j: count of voxel position j
    double bin0[F::calib ? 3 : 1];
    double n1, n2, uu[F::ncc ? 3 : 1], nu[F::ncc ? 3 : 1];  // NCC with n = R sum g^2 - (sum g)^2 = R^2 var(g)
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < (F::sum ? 3 : 1); ++k) sum[k] = 0.0;
#pragma unroll
        for (int k = 0; k < (F::thr ? 3 : 1); ++k) { thr[k] = 0.0; thrn[k] = 0u; }
        area = 0u; nvox = 0u;
#pragma unroll
        for (int r = 0; r < (F::dice ? RMAX : 1); ++r) { tp[r] = 0u; ps[r] = 0u; gs[r] = 0u; }
#pragma unroll
        for (int k = 0; k < (F::calib ? 3 : 1); ++k) bin0[k] = 0.0;
        n1 = 0.0; n2 = 0.0;
#pragma unroll
        for (int k = 0; k < (F::ncc ? 3 : 1); ++k) { uu[k] = 0.0; nu[k] = 0.0; }
    }
};

// ---- type rotation ------------------------------------------------------------------------------------------------------
// With REP == 16 two lanes (l, l + 16) share a histogram replica.  Instead of taking turns, the upper half-warp walks the
// uncertainty types in rotated order -- step s works on type s in the lower half and on type (s + 1) % 3 in the upper half
// -- so the halves never touch the same histogram region in the same step and a __syncwarp() between steps is all the
// ordering there is.  EVERYTHING per type is kept in step order by a lane (its values u, thresholds, Platt constants,
// partial sums); the flush puts the partial sums back into type order.  REP == 32: one replica per lane, no rotation.
template <int REP>
__device__ __forceinline__ bool stats2_rotated() { return REP == 16 && (threadIdx.x & 16) != 0; }
__device__ __forceinline__ int stats2_type_of_step(int s, bool rot) { return rot ? (s + 1) % VU_N_UNC : s; }

// per-thread constants of the statistics loop, in STEP order (loop-invariant registers)
struct Stat2Ctx {
    float a2s[VU_N_UNC], b2[VU_N_UNC], sgn[VU_N_UNC];  // conf = 1 / (1 + 2^(uu a2s + b2)), uu = u sgn
    float thr[VU_N_UNC];
    unsigned ebase[VU_N_UNC];  // shared address of E[type][0] - kMagicBits * 4
    unsigned hbase[VU_N_UNC];  // shared address of this lane's replica of hist[warp][type][0] - kMagicBits * REP * 8
};

__device__ __forceinline__ float2 lds_f2(unsigned addr) {
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint2 lds_u2(unsigned addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts_u2(unsigned addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ unsigned dp4a_u(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ unsigned opaque(unsigned x) { unsigned r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(x)); return r; }
__device__ __forceinline__ float lds_f(unsigned addr) { float r; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr)); return r; }
// all-ones when the comparison holds (false for NaN operands), else 0
__device__ __forceinline__ unsigned mask_ge(float a, float b) { unsigned r; asm("set.ge.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned mask_lt(float a, float b) { unsigned r; asm("set.lt.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }

// zero the histograms of `warps` statistics warps and build the edge tables; called by `nthreads` threads (t = index of the
// caller), which then synchronise among themselves
template <int REP>
__device__ __forceinline__ void stats2_init(const StatParams& sp, void* smem, int t, int nthreads, int warps) {
    if (!(sp.flags & VU_STAT_CALIB)) return;
    uint2* hist = reinterpret_cast<uint2*>(smem);
    const int nh = warps * (VU_N_UNC * kBins2 * REP);
    for (int i = t; i < nh; i += nthreads) hist[i] = make_uint2(0u, 0u);
    // E[k][c] = threshold on uu below which candidate bin c = round(conf * 20) steps down to bin c - 1.  The device's conf is
    // within ~1e-6 of the reference's, so the true bin is c or c - 1 and only this one edge has to be looked at:
    //   c = 0        never steps down (NaN compares false)
    //   c = 1 .. 19  interior edge c - 1 pulled back onto u
    //   c = 20       conf >= 0.975, bin 19 (its upper edge is 1 + 1e-8): always steps down (+inf)
    float* E = reinterpret_cast<float*>(hist + nh);
    for (int i = t; i < VU_N_UNC * kEdgePad; i += nthreads) {
        const int k = i / kEdgePad, e = i % kEdgePad;
        const float nan = __int_as_float(0x7fc00000), inf = __int_as_float(0x7f800000);
        E[i] = (e >= 1 && e <= VU_N_EDGES) ? sp.calib[k].edge[e - 1] : (e == VU_N_EDGES + 1 ? inf : nan);
    }
}

// warp = index of the caller's warp among the `warps` statistics warps
// PIN_ALL = false leaves the Platt constants and thresholds to the compiler (it re-reads them from the constant bank with a
// computed index): 12 registers less for kernels that are short of them.
template <int REP, bool PIN_ALL = true>
__device__ __forceinline__ void stats2_ctx(Stat2Ctx& cx, const StatParams& sp, void* smem, int warp, int warps) {
    uint2* hist = reinterpret_cast<uint2*>(smem);
    const int nh = warps * (VU_N_UNC * kBins2 * REP);
    const float* E = reinterpret_cast<const float*>(hist + nh);
    const int lane = threadIdx.x & 31;
    const bool rot = stats2_rotated<REP>();
#pragma unroll
    for (int s = 0; s < VU_N_UNC; ++s) {
        const int k = stats2_type_of_step(s, rot);
        cx.sgn[s] = sp.calib[k].sgn;
        cx.a2s[s] = sp.calib[k].a2 * sp.calib[k].sgn;
        cx.b2[s] = sp.calib[k].b2;
        cx.thr[s] = sp.thr[k];
        // (biased by the run-time copy of kMagicBits, see StatParams::magic_bits)
        cx.ebase[s] = (unsigned)__cvta_generic_to_shared(E + k * kEdgePad) - sp.magic_bits * 4u;
        cx.hbase[s] = (unsigned)__cvta_generic_to_shared(hist + (warp * VU_N_UNC + k) * (kBins2 * REP) + (lane & (REP - 1))) -
                      sp.magic_bits * (unsigned)(REP * 8);
        // pin the values: without this the compiler re-derives them from the kernel parameters inside the tile loop
        if (PIN_ALL) asm volatile("" : "+f"(cx.sgn[s]), "+f"(cx.a2s[s]), "+f"(cx.b2[s]), "+f"(cx.thr[s]));
        asm volatile("" : "+r"(cx.ebase[s]), "+r"(cx.hbase[s]));
    }
}

// the calibration update of the four voxels of one step: x = the uncertainties (finite), vc = samples | correct << 16 per
// voxel (0 for voxels that must not count), nv = samples, nvf = samples as a float
template <int REP>
__device__ __forceinline__ void calib2_step(const float (&x)[4], float a2s, float b2, float sgn, unsigned ebase, unsigned hbase,
                                            const unsigned (&vc)[4], const unsigned (&nv)[4], const float (&nvf)[4], float& bin0) {
    float uu[4], z[4], conf[4], kf[4], qf[4], e[4], d[4];
    const f32x2 S = pk2(sgn, sgn), A2 = pk2(a2s, a2s), B2 = pk2(b2, b2);
    const f32x2 U01 = mul2(pk2(x[0], x[1]), S), U23 = mul2(pk2(x[2], x[3]), S);
    upk2(U01, uu[0], uu[1]);
    upk2(U23, uu[2], uu[3]);
    upk2(fma2(U01, A2, B2), z[0], z[1]);
    upk2(fma2(U23, A2, B2), z[2], z[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = ex2_approx(z[j]);
    const f32x2 one2 = pk2(1.0f, 1.0f);
    upk2(add2(pk2(e[0], e[1]), one2), d[0], d[1]);
    upk2(add2(pk2(e[2], e[3]), one2), d[2], d[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) conf[j] = rcp_approx(d[j]);  // in [0, 1]
    const f32x2 C01 = pk2(conf[0], conf[1]), C23 = pk2(conf[2], conf[3]);
    const f32x2 twenty2 = pk2(20.0f, 20.0f), magic2 = pk2(kRoundMagic, kRoundMagic), qs2 = pk2((float)(1 << kQBits), (float)(1 << kQBits));
    upk2(fma2(C01, twenty2, magic2), kf[0], kf[1]);
    upk2(fma2(C23, twenty2, magic2), kf[2], kf[3]);
    upk2(fma2(C01, qs2, magic2), qf[0], qf[1]);
    upk2(fma2(C23, qs2, magic2), qf[2], qf[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned kb = __float_as_uint(kf[j]);  // kMagicBits + round(conf * 20)
        const float lower = lds_f(kb * 4u + ebase);      // threshold on uu below which the candidate steps down
        const unsigned bb = kb - (uu[j] < lower ? 1u : 0u);  // kMagicBits + bin
        const unsigned qv = __float_as_uint(qf[j]) - bb * (unsigned)kQBinStep;  // q + kYBias
        const unsigned ha = bb * (unsigned)(REP * 8) + hbase;
        uint2 w = lds_u2(ha);
        w.x += vc[j];
        w.y += qv * nv[j];
        sts_u2(ha, w);
        if (bb == kMagicBits) bin0 = fmaf(conf[j], nvf[j], bin0);
    }
}

// one sample the slow way (NaN / inf among the thread's values): np.digitize sends NaN past the last edge (slot 20);
// k = the uncertainty TYPE of the sample.  Returns what the sample adds to the floating sum of bin 0.
template <int REP>
__device__ __noinline__ double calib2_slow(const StatParams& sp, int k, float x, unsigned vc, unsigned ebase, unsigned hbase, long long b) {
    const unsigned nv = vc & 0xffffu, nc = vc >> 16;
    if (!nv) return 0.0;
    unsigned long long* irow = reinterpret_cast<unsigned long long*>(sp.i64 + b * VU_I64_COLS);
    double* frow = sp.f64 + b * VU_F64_COLS;
    // (an infinite u with a zero slope gives a NaN confidence as well: (-inf) * 0 in ace.py:329)
    const float conf = (x != x) ? x : platt_conf(x, sp.calib[k].a2, sp.calib[k].b2, 0);
    if (conf != conf) {
        atomicAdd(irow + VU_I64_BIN_TOTAL + k * VU_N_BINS + (VU_N_BINS - 1), (unsigned long long)nv);
        if (nc) atomicAdd(irow + VU_I64_BIN_TRUE + k * VU_N_BINS + (VU_N_BINS - 1), (unsigned long long)nc);
        atomicAdd(frow + VU_F64_BIN_SUMS + k * VU_N_BINS + (VU_N_BINS - 1), (double)__int_as_float(0x7fc00000));
        return 0.0;
    }
    const unsigned kb = __float_as_uint(fmaf(conf, 20.0f, kRoundMagic));
    const float uu = x * sp.calib[k].sgn;
    const unsigned bb = kb - (uu < lds_f(kb * 4u + ebase) ? 1u : 0u);
    const unsigned qv = __float_as_uint(fmaf(conf, (float)(1 << kQBits), kRoundMagic)) - bb * (unsigned)kQBinStep;
    const unsigned ha = bb * (unsigned)(REP * 8) + hbase;
    uint2 w = lds_u2(ha);
    w.x += vc;
    w.y += qv * nv;
    sts_u2(ha, w);
    return bb == kMagicBits ? (double)conf * (double)nv : 0.0;
}

// The reference words of the four voxels v .. v + 3 of an image, one per rater (zero past the last rater).  Issued early by
// the kernels that can (the latency of the load then hides behind their streaming phase).
template <unsigned FL, int RMAX>
__device__ __forceinline__ void stats2_load_refs(const StatParams& sp, bool active, const uint8_t* gt_img, long long v, unsigned (&W)[RMAX]) {
    // gt_img: references of the image (gt.data + b * gt.stride_b; the kernels keep it per image instead of forming the
    // 64-bit product for every tile)
#pragma unroll
    for (int r = 0; r < RMAX; ++r) W[r] = 0u;
    if (StatFlags<FL>::refs && active) {
        const int R = sp.gt.R;
        const uint8_t* gp = gt_img + v;
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            if (r < R) W[r] = __ldg(reinterpret_cast<const unsigned*>(gp));
            gp += sp.gt.sr;
        }
    }
}

// The statistics of the four consecutive voxels v .. v + 3 of image b owned by one thread.  U0, U1, U2: the three
// uncertainty values of the voxels in the lane's STEP order (see "type rotation").
// REP == 32: inactive threads may skip the call.  REP == 16: every lane of the warp must make it (active = false past the
// end of the image).
template <unsigned FL, int RMAX, int REP>
__device__ __forceinline__ void stats2_tile(StatAcc<FL, RMAX>& A, const StatParams& sp, const Stat2Ctx& cx, bool active, long long b,
                                            float4 U0, float4 U1, float4 U2, unsigned lab4, const unsigned (&W)[RMAX]) {
    using F = StatFlags<FL>;
    const float u[VU_N_UNC][4] = {{U0.x, U0.y, U0.z, U0.w}, {U1.x, U1.y, U1.z, U1.w}, {U2.x, U2.y, U2.z, U2.w}};
    unsigned nv4 = 0u, nc4 = 0u;  // per voxel (byte j): valid references, references equal to the label
    float chk = 0.f;              // 0 unless one of the twelve values is NaN / inf
    if (active) {
        A.nvox += 4u;
        float s[VU_N_UNC];
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) {
            float lo, hi;
            upk2(add2(pk2(u[k][0], u[k][1]), pk2(u[k][2], u[k][3])), lo, hi);
            s[k] = lo + hi;
            if (F::sum) A.sum[k] += (double)s[k];
        }
        chk = ((s[0] + s[1]) + s[2]) * 0.0f;
        if (F::thr) {
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k) {
                const float t = cx.thr[k];
                float ts = 0.f;
                unsigned n = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool hit = u[k][j] >= t;
                    ts += hit ? u[k][j] : 0.f;
                    n += hit ? 1u : 0u;
                }
                A.thr[k] += (double)ts;
                A.thrn[k] += n;
            }
        }
        if (F::area) A.area += __popc(bytes_nonzero(lab4));
        if (F::refs) {
            const int R = sp.gt.R;
            const unsigned pp_hi = bytes_equal(lab4, kB01);  // predicted foreground (test_2D.py:878)
            if (sp.gt.ign_byte) {
                const unsigned ign4 = sp.gt.ign4;
#pragma unroll
                for (int r = 0; r < RMAX; ++r) {
                    if (r < R) {
                        const unsigned valid_hi = bytes_nonzero(W[r] ^ ign4);           // ace.py:492-499, test_2D.py:880
                        const unsigned eq_hi = ~bytes_nonzero(W[r] ^ lab4) & valid_hi;  // ace.py:488
                        nv4 += valid_hi >> 7;
                        nc4 += eq_hi >> 7;
                        if (F::dice) {
                            const unsigned gp_hi = ~bytes_nonzero(W[r] ^ kB01) & valid_hi;  // test_2D.py:882
                            const unsigned ppv = pp_hi & valid_hi;
                            A.tp[r] += (ppv & gp_hi) >> 7;
                            A.ps[r] += ppv >> 7;
                            A.gs[r] += gp_hi >> 7;
                        }
                    }
                }
            } else {  // no reference can be the ignore value: every rater is valid everywhere
                nv4 = (unsigned)R * kB01;
                const unsigned ps = pp_hi >> 7;
#pragma unroll
                for (int r = 0; r < RMAX; ++r) {
                    if (r < R) {
                        nc4 += (~bytes_nonzero(W[r] ^ lab4) & kB80) >> 7;
                        if (F::dice) {
                            const unsigned gp_hi = ~bytes_nonzero(W[r] ^ kB01) & kB80;
                            A.tp[r] += (pp_hi & gp_hi) >> 7;
                            A.ps[r] += ps;
                            A.gs[r] += gp_hi >> 7;
                        }
                    }
                }
            }
            if (F::ncc) {
                // n_j = R sum_r g_rj^2 - (sum_r g_rj)^2 = R^2 var_r(g_rj), exact in integers (experiment_dataloader.py:283);
                // the bytes of four raters are transposed so that one dp4a sums over the raters of a voxel
                unsigned gs_[4] = {0u, 0u, 0u, 0u}, gq_[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int r0 = 0; r0 < RMAX; r0 += 4) {
                    if (r0 < R) {
                        const unsigned w0 = W[r0], w1 = r0 + 1 < RMAX ? W[r0 + 1] : 0u, w2 = r0 + 2 < RMAX ? W[r0 + 2] : 0u,
                                       w3 = r0 + 3 < RMAX ? W[r0 + 3] : 0u;
                        const unsigned a01 = prmt(w0, w1, 0x5140u), b01 = prmt(w0, w1, 0x7362u);  // (w0.b0 w1.b0 w0.b1 w1.b1), (.. b2 .. b3)
                        const unsigned a23 = prmt(w2, w3, 0x5140u), b23 = prmt(w2, w3, 0x7362u);
                        const unsigned t[4] = {prmt(a01, a23, 0x5410u), prmt(a01, a23, 0x7632u), prmt(b01, b23, 0x5410u),
                                               prmt(b01, b23, 0x7632u)};  // t[j] = (g0j, g1j, g2j, g3j)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            gs_[j] = dp4a_u(t[j], kB01, gs_[j]);
                            gq_[j] = dp4a_u(t[j], t[j], gq_[j]);
                        }
                    }
                }
                float nf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) nf[j] = (float)(int)((unsigned)R * gq_[j] - gs_[j] * gs_[j]);  // < 2^24: exact
                const float n1 = (nf[0] + nf[1]) + (nf[2] + nf[3]);
                const float n2 = fmaf(nf[3], nf[3], fmaf(nf[2], nf[2], fmaf(nf[1], nf[1], nf[0] * nf[0])));
                A.n1 += (double)n1;
                A.n2 += (double)n2;
#pragma unroll
                for (int k = 0; k < VU_N_UNC; ++k) {
                    float uu = 0.f, nu = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uu = fmaf(u[k][j], u[k][j], uu);
                        nu = fmaf(nf[j], u[k][j], nu);
                    }
                    A.uu[k] += (double)uu;
                    A.nu[k] += (double)nu;
                }
            }
        }
    }
    if (F::calib) {
        unsigned vc[4], nv[4];
        float nvf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // samples | correct << 16: byte 0 = byte j of nv4, byte 2 = byte j of nc4; bytes 1 and 3 take the replicated sign bit
            // of byte j of nv4, which is 0 (at most 8 raters)
            vc[j] = prmt(nv4, nc4, (unsigned)j | ((unsigned)(8 + j) << 4) | ((unsigned)(4 + j) << 8) | ((unsigned)(8 + j) << 12));
            nv[j] = prmt(nv4, 0u, 0x4440u + (unsigned)j);
            nvf[j] = (float)nv[j];
        }
        const bool rot = stats2_rotated<REP>();
        const bool slow = chk != 0.0f;  // NaN != 0 as well
        auto slow_samples = [&]() {
#pragma unroll 1
            for (int i = 0; i < VU_N_UNC * 4; ++i) {
                const int s = i >> 2, j = i & 3;
                const float4 Us = s == 0 ? U0 : (s == 1 ? U1 : U2);
                const float x = j == 0 ? Us.x : (j == 1 ? Us.y : (j == 2 ? Us.z : Us.w));
                const unsigned eb = s == 0 ? cx.ebase[0] : (s == 1 ? cx.ebase[1] : cx.ebase[2]);
                const unsigned hb = s == 0 ? cx.hbase[0] : (s == 1 ? cx.hbase[1] : cx.hbase[2]);
                const unsigned vcj = ((nv4 >> (8 * j)) & 0xffu) | (((nc4 >> (8 * j)) & 0xffu) << 16);
                const double b0 = calib2_slow<REP>(sp, stats2_type_of_step(s, rot), x, vcj, eb, hb, b);
                if (s == 0) A.bin0[0] += b0; else if (s == 1) A.bin0[1] += b0; else A.bin0[2] += b0;
            }
        };
        if (REP == 32) {
            if (!slow) {
#pragma unroll
                for (int s = 0; s < VU_N_UNC; ++s) {
                    float b0 = 0.f;
                    calib2_step<REP>(u[s], cx.a2s[s], cx.b2[s], cx.sgn[s], cx.ebase[s], cx.hbase[s], vc, nv, nvf, b0);
                    A.bin0[s] += (double)b0;
                }
            } else {
                slow_samples();
            }
        } else {
            // a slow lane sits the fast steps out (its samples carry zero weight there, its values are replaced by zeros) and
            // catches up afterwards, one half-warp at a time
            const bool any_slow = __any_sync(kFull, slow);
            if (!any_slow) {
#pragma unroll
                for (int s = 0; s < VU_N_UNC; ++s) {
                    float b0 = 0.f;
                    calib2_step<REP>(u[s], cx.a2s[s], cx.b2[s], cx.sgn[s], cx.ebase[s], cx.hbase[s], vc, nv, nvf, b0);
                    A.bin0[s] += (double)b0;
                    __syncwarp();
                }
            } else {
                float xs[VU_N_UNC][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (slow) { vc[j] = 0u; nv[j] = 0u; }
#pragma unroll
                    for (int s = 0; s < VU_N_UNC; ++s) xs[s][j] = slow ? 0.f : u[s][j];
                }
#pragma unroll 1
                for (int s = 0; s < VU_N_UNC; ++s) {
                    float b0 = 0.f;
                    const float x4[4] = {s == 0 ? xs[0][0] : (s == 1 ? xs[1][0] : xs[2][0]), s == 0 ? xs[0][1] : (s == 1 ? xs[1][1] : xs[2][1]),
                                         s == 0 ? xs[0][2] : (s == 1 ? xs[1][2] : xs[2][2]), s == 0 ? xs[0][3] : (s == 1 ? xs[1][3] : xs[2][3])};
                    calib2_step<REP>(x4, s == 0 ? cx.a2s[0] : (s == 1 ? cx.a2s[1] : cx.a2s[2]), s == 0 ? cx.b2[0] : (s == 1 ? cx.b2[1] : cx.b2[2]),
                                     s == 0 ? cx.sgn[0] : (s == 1 ? cx.sgn[1] : cx.sgn[2]), s == 0 ? cx.ebase[0] : (s == 1 ? cx.ebase[1] : cx.ebase[2]),
                                     s == 0 ? cx.hbase[0] : (s == 1 ? cx.hbase[1] : cx.hbase[2]), vc, nv, nvf, b0);
                    if (s == 0) A.bin0[0] += (double)b0; else if (s == 1) A.bin0[1] += (double)b0; else A.bin0[2] += (double)b0;
                    __syncwarp();
                }
#pragma unroll 1
                for (int turn = 0; turn < 2; ++turn) {
                    if (slow && (int)rot == turn) slow_samples();
                    __syncwarp();
                }
            }
        }
    }
}

// Fold a warp's register partials into row b of the global statistics buffers (one atomic per non-zero value and warp)
// and clear them.  Every lane of the warp calls it.
template <unsigned FL, int RMAX, int REP>
__device__ __forceinline__ void stats2_flush_regs(StatAcc<FL, RMAX>& A, const StatParams& sp, long long b) {
    using F = StatFlags<FL>;
    const int lane = threadIdx.x & 31;
    const bool rot = stats2_rotated<REP>();
    double* frow = sp.f64 + b * VU_F64_COLS;
    unsigned long long* irow = reinterpret_cast<unsigned long long*>(sp.i64 + b * VU_I64_COLS);
    auto fadd = [&](int col, double x) {
        x = warp_sum(x);
        if (lane == 0 && x != 0.0) atomicAdd(frow + col, x);
    };
    auto iadd = [&](int col, unsigned x) {
        // per-thread counts stay far below 2^27 between flushes, so the warp total fits 32 bits
        const unsigned t = __reduce_add_sync(kFull, x);
        if (lane == 0 && t) atomicAdd(irow + col, (unsigned long long)t);
    };
    // step order -> type order: type k was step (k + 2) % 3 of a rotated lane
    auto of_type = [&](auto& a, int k) { return rot ? a[(k + 2) % VU_N_UNC] : a[k]; };
    if (F::sum) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) {
            const double s = warp_sum(of_type(A.sum, k));
            if (lane == 0 && s != 0.0) {
                if (FL & VU_STAT_IMAGE_SUM) atomicAdd(frow + VU_F64_SUM + k, s);
                if (FL & VU_STAT_NCC) atomicAdd(frow + VU_F64_NCC_U + k, s);
            }
        }
    }
    if (F::thr) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) { fadd(VU_F64_THR_SUM + k, of_type(A.thr, k)); iadd(VU_I64_THR_COUNT + k, of_type(A.thrn, k)); }
    }
    if (F::area) iadd(VU_I64_AREA, A.area);
    iadd(VU_I64_NVOX, A.nvox);
    if (F::dice) {
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
            if (r < sp.gt.R) {
                iadd(VU_I64_DICE_TP + r, dp4a_u(A.tp[r], kB01, 0u));  // the four byte counters of the thread
                iadd(VU_I64_DICE_PRED + r, dp4a_u(A.ps[r], kB01, 0u));
                iadd(VU_I64_DICE_GT + r, dp4a_u(A.gs[r], kB01, 0u));
            }
    }
    if (F::calib) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) fadd(VU_F64_BIN_SUMS + k * VU_N_BINS, of_type(A.bin0, k));
    }
    if (F::ncc) {
        const double rr = (double)sp.gt.R, r2 = 1.0 / (rr * rr);  // g = n / R^2 (np.var, ddof = 0)
        fadd(VU_F64_NCC_G, A.n1 * r2);
        fadd(VU_F64_NCC_GG, A.n2 * (r2 * r2));
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) { fadd(VU_F64_NCC_UU + k, of_type(A.uu, k)); fadd(VU_F64_NCC_GU + k, of_type(A.nu, k) * r2); }
    }
    A.clear();
}

// Fold the histograms of ONE warp (index `warp` among the statistics warps) into row b and clear them: no barrier with the
// other warps is needed.  Every lane of the warp calls it.
template <int REP>
__device__ __noinline__ void stats2_flush_hist_warp(const StatParams& sp, void* smem, long long b, int warp) {
    uint2* hist = reinterpret_cast<uint2*>(smem) + (size_t)warp * (VU_N_UNC * kBins2 * REP);
    const int lane = threadIdx.x & 31;
    double* frow = sp.f64 + b * VU_F64_COLS;
    unsigned long long* irow = reinterpret_cast<unsigned long long*>(sp.i64 + b * VU_I64_COLS);
    __syncwarp();
    for (int pair = lane; pair < VU_N_UNC * kBins2; pair += 32) {
        unsigned tot = 0u, tru = 0u;
        long long q = 0;
#pragma unroll 4
        for (int i = 0; i < REP; ++i) {
            uint2* h = hist + pair * REP + ((i + lane) & (REP - 1));  // skewed: the lanes read different banks
            const uint2 x = *h;
            *h = make_uint2(0u, 0u);
            tot += x.x & 0xffffu;
            tru += x.x >> 16;
            q += (int)(x.y - (x.x & 0xffffu) * kYBias);
        }
        if (tot) {
            const int k = pair / kBins2, bin = pair % kBins2;
            const int col = k * VU_N_BINS + bin;
            atomicAdd(irow + VU_I64_BIN_TOTAL + col, (unsigned long long)tot);
            if (tru) atomicAdd(irow + VU_I64_BIN_TRUE + col, (unsigned long long)tru);
            if (bin > 0)  // bin 0 is summed in floating point (StatAcc::bin0)
                atomicAdd(frow + VU_F64_BIN_SUMS + col,
                          ((double)tot * (double)(bin * kQBinStep) + (double)q) * (1.0 / (double)(1 << kQBits)));
        }
    }
    __syncwarp();
}

}  // namespace vu
