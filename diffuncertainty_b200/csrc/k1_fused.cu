// K1: the fused streaming pass over the (P, B, C, V) probability slab.
//
// One read of HBM per element produces, per voxel: the member mean (torch's
// cascade order, bit-exact), its argmax label, TU / AU / EU, and -- when
// requested -- the per-image reductions behind image/threshold aggregation,
// area, Dice counts, the 20-bin calibration histograms and the NCC sums.
//
// Reference semantics: uncertainty_modeling/test_2D.py:969-971 (slice + mean),
// :871 (argmax), unc_mod_utils/test_utils.py:833-864 (uncertainty measures),
// evaluation/uncertainty_aggregation/aggregate_uncertainties.py:37-39,124-125,
// evaluation/metrics/ace.py:350-356, evaluation/metrics/ncc.py:17-27.
#include "k1_core.cuh"
#include "vu_host.h"

namespace vu {

extern __shared__ __align__(16) unsigned char vu_dyn_smem[];

struct K1Params {
    const float* x;
    const float* const* mptr;  // optional device array of P member base pointers (then x / sp are unused)
    long long P, B, C, V;
    long long sp, sb, sc, sv;
    long long sd;        // stride between the draws of a member (data form)
    int draws;           // draws per member (>= 1)
    unsigned sflags;     // VU_SLAB_RENORMALIZE | VU_SLAB_DISCRETIZE | VU_SLAB_LOGITS
    float renorm_eps;
    float* tu;
    float* au;
    float* eu;
    uint8_t* lab;
    uint8_t* mlab;  // per-member labels (P, B, V) or NULL
    long long tiles_per_img, total_tiles;
    StatParams st;
};

// ---------------------------------------------------------------------------
// Fast kernel: C compile-time, one thread owns VEC consecutive voxels and walks
// the members in order (that is what makes the mean bit-exact), G members per
// load stage.
//   LEVELS = 1: P <= 17 (plain sequential sum == torch's cascade)
//   LEVELS = 2: P <= 271 (level-0 accumulator folded into level 1 every 16)
// The E = C * VEC values of one member are handled as E/2 packed fp32 pairs
// (FADD2 / FFMA2): pairs of neighbouring voxels for VEC = 2, 4, pairs of classes
// for VEC = 1 (with one scalar leftover when C is odd).  The entropy of a member
// is accumulated class by class with one fma per term in every variant, so all
// variants and the generic kernel produce bit-identical maps.
// Requires stride_v == 1 and VEC-element alignment of every row start.
// ---------------------------------------------------------------------------
template <int VEC>
struct PairLoad;
template <>
struct PairLoad<4> {  // 16 bytes -> two pairs
    __device__ __forceinline__ static void load(const float* p, f32x2* x) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(x[0]), "=l"(x[1]) : "l"(p));
    }
};
template <>
struct PairLoad<2> {
    __device__ __forceinline__ static void load(const float* p, f32x2* x) {
        asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(x[0]) : "l"(p));
    }
};
template <>
struct PairLoad<1> {  // never called (VEC == 1 pairs classes from scalar loads); keeps the dead branch well-formed
    __device__ __forceinline__ static void load(const float*, f32x2*) {}
};

template <int C, int VEC, int LEVELS, int THREADS, int MINB, int G, bool STATS>
__global__ void __launch_bounds__(THREADS, MINB) k1_fast(const __grid_constant__ K1Params prm) {
    constexpr bool do_stats = STATS;
    constexpr long long kTileVox = (long long)THREADS * VEC;
    using Acc = VoxelAcc<C, VEC, LEVELS>;
    constexpr int NP = Acc::NP, NH = Acc::NH;
    constexpr bool ODD = Acc::ODD;
    constexpr int PF = (C * VEC >= 16) ? 1 : (16 / (C * VEC));  // members of the next tile prefetched into L2
    StatsCursor<THREADS> cursor;
    if (do_stats) stats_init<THREADS>(prm.st, vu_dyn_smem);

    const long long P = prm.P, V = prm.V;
    const float Pf = (float)P;
    // consecutive tiles per CTA (total_tiles < 2^31 is checked on the host): (b, vt) advance incrementally
    const int t0 = (int)(prm.total_tiles * (long long)blockIdx.x / gridDim.x);
    const int t1 = (int)(prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x);
    const int tpi = (int)prm.tiles_per_img;
    int b = t0 / tpi, vt = t0 - b * tpi - 1;

    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        const long long v = (long long)vt * kTileVox + (long long)threadIdx.x * VEC;
        if (do_stats) cursor.enter(prm.st, vu_dyn_smem, b, vt, kTileVox);
        const bool active = v < V;
        float u[VU_N_UNC][VEC];
        int label[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { u[0][k] = u[1][k] = u[2][k] = 0.f; label[k] = 0; }

        if (active) {
            if (do_stats) stats_prefetch_gt<VEC>(prm.st, b, v);
            Acc acc;
            acc.init();
            const long long off0 = (long long)b * prm.sb + v;  // offset of this thread's first voxel inside a member

            f32x2 xp[G][NP];
            float xs[G];
            for (long long p0 = 0; p0 < P; p0 += G) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    xs[g] = 0.f;
                    if (G == 1 || p0 + g < P) {
                        const float* r = (prm.mptr ? ld_member_ptr(prm.mptr, p0 + g) : prm.x + (p0 + g) * prm.sp) + off0;
                        if constexpr (VEC >= 2) {
#pragma unroll
                            for (int c = 0; c < C; ++c) PairLoad<VEC>::load(r + c * prm.sc, &xp[g][c * NH]);
                        } else {
#pragma unroll
                            for (int j = 0; j < NP; ++j) xp[g][j] = pk2(ldg_stream(r + (2 * j) * prm.sc), ldg_stream(r + (2 * j + 1) * prm.sc));
                            if constexpr (ODD) xs[g] = ldg_stream(r + (C - 1) * prm.sc);
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < G; ++g)
                    if (G == 1 || p0 + g < P) {
                        acc.add_member(xp[g], xs[g], p0 + g, prm.mlab != nullptr);
                        if (prm.mlab) VecLoad<VEC>::store_u8(prm.mlab + ((p0 + g) * prm.B + b) * V + v, acc.bi);
                    }
            }

            // Pull the first members of the CTA's next tile into L2 while this tile's epilogue and statistics
            // run: no loads of this warp are in flight during those phases otherwise.
            if (PF > 0 && tile + 1 < t1) {
                const bool wrap = (vt + 1 == tpi);
                const long long nv = wrap ? (long long)threadIdx.x * VEC : v + kTileVox;
                if (nv < V) {
                    const long long noff = (long long)(wrap ? b + 1 : b) * prm.sb + nv;
#pragma unroll
                    for (int g = 0; g < PF; ++g)
                        if (g < P) {
                            const float* nrow = (prm.mptr ? ld_member_ptr(prm.mptr, g) : prm.x + g * prm.sp) + noff;
#pragma unroll
                            for (int c = 0; c < C; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + c * prm.sc));
                        }
                }
            }

            acc.finish(Pf, u, label);
            const long long o = (long long)b * V + v;
            if (prm.tu) VecLoad<VEC>::store(prm.tu + o, u[0]);
            if (prm.au) VecLoad<VEC>::store(prm.au + o, u[1]);
            if (prm.eu) VecLoad<VEC>::store(prm.eu + o, u[2]);
            if (prm.lab) VecLoad<VEC>::store_u8(prm.lab + o, label);
        }

        if (do_stats) stats_tile<VEC, THREADS>(prm.st, vu_dyn_smem, active, b, v, u, label);
    }
    if (do_stats && t1 > t0) cursor.finish(prm.st, vu_dyn_smem, vt, kTileVox);
}

// ---------------------------------------------------------------------------
// Generic kernel: any C <= 256, any strides / alignment, any P up to the
// shared-memory limit, all four cascade levels.  Class-outer order: the mean of
// one class is finished in registers before the next class starts; the
// per-member entropies accumulate in shared memory (h[p][thread]).
// Also serves P == 1 (calculate_one_minus_msr, test_utils.py:862-864).
// ---------------------------------------------------------------------------
// dynamic shared memory of the generic kernel before the statistics state: entropies [P][T] (P > 1) and, for member
// labels, best value [P][T] + best index [P][T], rounded up to 16 bytes
// (+ with renormalised / discretised draws: the class sum [P * draws][T] and the argmax [P * draws][T] of every draw)
__host__ __device__ inline size_t generic_smem_bytes(long long P, int threads, bool member_labels, long long draws = 1, unsigned sflags = 0) {
    size_t n = (P > 1 ? (size_t)P * threads * sizeof(float) : 0);
    if (member_labels) n += (size_t)P * threads * (sizeof(float) + 1);
    n = (n + 3) / 4 * 4;
    if (sflags & VU_SLAB_RENORMALIZE) n += (size_t)P * draws * threads * sizeof(float);
    if (sflags & VU_SLAB_LOGITS) n += 2 * (size_t)P * draws * threads * sizeof(float);
    if (sflags & VU_SLAB_DISCRETIZE) n += (size_t)P * draws * threads;
    return (n + 15) / 16 * 16;
}

struct Cascade {
    float a[4];
    __device__ __forceinline__ void reset() { a[0] = a[1] = a[2] = a[3] = 0.f; }
    // i = index of the member being added, n_full = members covered by whole
    // chunks, k = log2(chunk)
    __device__ __forceinline__ void add(float x, long long i, long long n_full, int k) {
        a[0] += x;
        fold(i, n_full, k);
    }
    // a[0] += e * r in one rounding (a member of a slab of logits: probability = exponential x 1 / sum, VoxelAcc::add_member_pre)
    __device__ __forceinline__ void add_product(float e, float r, long long i, long long n_full, int k) {
        a[0] = __fmaf_rn(e, r, a[0]);
        fold(i, n_full, k);
    }
    __device__ __forceinline__ void fold(long long i, long long n_full, int k) {
        const long long done = i + 1;
        const long long mask = (1LL << k) - 1;
        if (i < n_full && (done & mask) == 0) {
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                a[j] += a[j - 1];
                a[j - 1] = 0.f;
                if (done & (mask << (j * k))) break;
            }
        }
    }
    __device__ __forceinline__ float total() const { return ((a[0] + a[1]) + a[2]) + a[3]; }
};

constexpr int kGenAhead = 8;  // members whose loads are in flight per thread

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k1_generic(const __grid_constant__ K1Params prm, int level_k) {
    float* h_smem = reinterpret_cast<float*>(vu_dyn_smem);  // [P][THREADS] (P > 1), then the statistics state
    // per-member argmax state when member labels are wanted: best value [P][THREADS], best index [P][THREADS]
    const bool want_ml = prm.mlab != nullptr;
    float* bv_smem = h_smem + (prm.P > 1 ? (size_t)prm.P * THREADS : 0);
    uint8_t* bi_smem = reinterpret_cast<uint8_t*>(bv_smem + (want_ml ? (size_t)prm.P * THREADS : 0));
    const int D = prm.draws;
    const bool renorm = prm.sflags & VU_SLAB_RENORMALIZE, onehot = prm.sflags & VU_SLAB_DISCRETIZE, logits = prm.sflags & VU_SLAB_LOGITS;
    // logits with nothing else folded in: a member's entropy term comes straight from its softmax (the TMA form does the same)
    const bool plain_logits = logits && !renorm && !onehot && D == 1;
    const size_t pre_bytes = ((prm.P > 1 ? (size_t)prm.P * THREADS * 4 : 0) + (want_ml ? (size_t)prm.P * THREADS * 5 : 0) + 3) / 4 * 4;
    float* norm_smem = reinterpret_cast<float*>(vu_dyn_smem + pre_bytes);  // class sum of every draw (renormalisation)
    float* lmax_smem = norm_smem + (renorm ? (size_t)prm.P * D * THREADS : 0);  // logits: maximum and 1 / sum of every draw's softmax
    float* lrs_smem = lmax_smem + (logits ? (size_t)prm.P * D * THREADS : 0);
    uint8_t* amax_smem = reinterpret_cast<uint8_t*>(lrs_smem + (logits ? (size_t)prm.P * D * THREADS : 0));  // argmax of every draw
    void* st_smem = vu_dyn_smem + generic_smem_bytes(prm.P, THREADS, want_ml, D, prm.sflags);
    const bool do_stats = prm.st.flags != 0;
    StatsCursor<THREADS> cursor;
    if (do_stats) stats_init<THREADS>(prm.st, st_smem);

    const long long P = prm.P, V = prm.V;
    const int C = (int)prm.C;
    const float Pf = (float)P;
    const long long n_full = (P >> level_k) << level_k;
    // the cascade of a sum over the classes of a draw (torch.sum(dim=1)) and over the draws of a member (mean(dim=1))
    auto level_of = [](long long n) { int lg = 0; while ((1LL << lg) < n) ++lg; return lg / 4 > 4 ? lg / 4 : 4; };
    const int kC = level_of(C), kD = level_of(D);
    const long long nfC = ((long long)C >> kC) << kC, nfD = ((long long)D >> kD) << kD;
    const float Df = (float)D;
    // draw d of member p
    auto draw_base = [&](long long p, int d) {
        return prm.mptr ? ld_member_ptr(prm.mptr, p * D + d) : prm.x + p * prm.sp + (long long)d * prm.sd;
    };
    // what the reference's producers make of the raw value x of class c of draw pd at this thread's voxel
    auto softmaxed = [&](float x, long long pd) {  // F.softmax(logits, dim=1), test_2D.py:1181-1256
        return __fmul_rn(ex2_approx(softmax_z(x, lmax_smem[pd * THREADS + threadIdx.x])), lrs_smem[pd * THREADS + threadIdx.x]);
    };
    auto produced = [&](float x, long long pd, int c) {
        if (logits) x = softmaxed(x, pd);
        if (renorm) {
            const float nrm = norm_smem[pd * THREADS + threadIdx.x];
            x = (nrm > prm.renorm_eps) ? __fdiv_rn(x, fmaxf(nrm, prm.renorm_eps)) : x;  // test_2D.py:190-194
        }
        if (onehot) x = (amax_smem[pd * THREADS + threadIdx.x] == (uint8_t)c) ? 1.0f : 0.0f;  // test_2D.py:1274
        return x;
    };
    const int t0 = (int)(prm.total_tiles * (long long)blockIdx.x / gridDim.x);
    const int t1 = (int)(prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x);
    const int tpi = (int)prm.tiles_per_img;
    int b = t0 / tpi, vt = t0 - b * tpi - 1;
    float* h = h_smem + threadIdx.x;

    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        const long long v = (long long)vt * THREADS + threadIdx.x;
        if (do_stats) cursor.enter(prm.st, st_smem, b, vt, THREADS);
        const bool active = v < V;
        float u[VU_N_UNC] = {0.f, 0.f, 0.f};
        int label = 0;
        if (active) {
            const long long off0 = (long long)b * prm.sb + v * prm.sv;
            if (P > 1)
                for (long long p = 0; p < P; ++p) h[p * THREADS] = 0.f;
            if (renorm | onehot | logits) {
                // first pass over the draws: the softmax constants of a draw of logits, the class sum (torch.sum(dim=1):
                // cascade order) and the argmax of the (renormalised) draw
                for (long long pd = 0; pd < P * D; ++pd) {
                    const float* base = draw_base(pd / D, (int)(pd % D)) + off0;
                    if (logits) {
                        float m = ldg_stream(base);
                        for (int c = 1; c < C; ++c) m = max_nan(m, ldg_stream(base + (long long)c * prm.sc));
                        float S = 0.f, EZ = 0.f, e, rS, hh;
                        for (int c = 0; c < C; ++c) softmax_term(ldg_stream(base + (long long)c * prm.sc), m, S, EZ, e);
                        softmax_finish(S, EZ, rS, hh);
                        lmax_smem[pd * THREADS + threadIdx.x] = m;
                        lrs_smem[pd * THREADS + threadIdx.x] = rS;
                        if (plain_logits && P > 1) h[pd * THREADS] = hh;
                    }
                    if (renorm) {
                        Cascade cs;
                        cs.reset();
                        for (int c = 0; c < C; ++c) {
                            float x = ldg_stream(base + (long long)c * prm.sc);
                            if (logits) x = softmaxed(x, pd);
                            cs.add(x, c, nfC, kC);
                        }
                        norm_smem[pd * THREADS + threadIdx.x] = cs.total();
                    }
                    if (onehot) {
                        const float nrm = renorm ? norm_smem[pd * THREADS + threadIdx.x] : 0.f;
                        float bv = 0.f;
                        int bi = 0;
                        for (int c = 0; c < C; ++c) {
                            float x = ldg_stream(base + (long long)c * prm.sc);
                            if (logits) x = softmaxed(x, pd);
                            if (renorm) x = (nrm > prm.renorm_eps) ? __fdiv_rn(x, fmaxf(nrm, prm.renorm_eps)) : x;
                            if (c == 0) { bv = x; bi = 0; } else argmax_step(x, c, bv, bi);
                        }
                        amax_smem[pd * THREADS + threadIdx.x] = (uint8_t)bi;
                    }
                }
            }
            float best = 0.f, tu2 = 0.f;
            for (int c = 0; c < C; ++c) {
                Cascade cas;
                cas.reset();
                const long long offc = off0 + (long long)c * prm.sc;
                // what happens to the value x of member p (the mean sum `cas` took it already when the slab holds plain logits)
                auto consume = [&](long long p, float x) {
                    if (!plain_logits) cas.add(x, p, n_full, level_k);
                    if (P > 1 && !plain_logits) h[p * THREADS] = plog2p_acc(h[p * THREADS], x);
                    if (want_ml) {
                        float& bvp = bv_smem[p * THREADS + threadIdx.x];
                        uint8_t& bip = bi_smem[p * THREADS + threadIdx.x];
                        if (c == 0) { bvp = x; bip = 0; }
                        else {
                            float bb = bvp;
                            int ii = bip;
                            argmax_step(x, c, bb, ii);
                            bvp = bb; bip = (uint8_t)ii;
                        }
                    }
                };
                if (D == 1 && !(renorm | onehot | logits)) {
                    // plain probabilities: the loads of kGenAhead members are issued before the first is consumed (one load in
                    // flight per thread left the kernel latency-bound: 0.5 TB/s; the order of the arithmetic is unchanged)
                    for (long long p0 = 0; p0 < P; p0 += kGenAhead) {
                        float xv[kGenAhead];
#pragma unroll
                        for (int j = 0; j < kGenAhead; ++j)
                            if (p0 + j < P) xv[j] = ldg_stream((prm.mptr ? ld_member_ptr(prm.mptr, p0 + j) : prm.x + (p0 + j) * prm.sp) + offc);
#pragma unroll
                        for (int j = 0; j < kGenAhead; ++j)
                            if (p0 + j < P) consume(p0 + j, xv[j]);
                    }
                } else if (plain_logits) {
                    for (long long p0 = 0; p0 < P; p0 += kGenAhead) {
                        float xv[kGenAhead];
#pragma unroll
                        for (int j = 0; j < kGenAhead; ++j)
                            if (p0 + j < P) xv[j] = ldg_stream(draw_base(p0 + j, 0) + offc);
#pragma unroll
                        for (int j = 0; j < kGenAhead; ++j)
                            if (p0 + j < P) {
                                const long long p = p0 + j;
                                const float e = ex2_approx(softmax_z(xv[j], lmax_smem[p * THREADS + threadIdx.x]));
                                const float r = lrs_smem[p * THREADS + threadIdx.x];
                                cas.add_product(e, r, p, n_full, level_k);
                                consume(p, __fmul_rn(e, r));  // (per-member labels)
                            }
                    }
                } else {
                    for (long long p = 0; p < P; ++p) {
                        // the member is the mean of its draws (test_2D.py:1277): cascade sum, true division
                        Cascade cd;
                        cd.reset();
                        for (int d = 0; d < D; ++d) {
                            const float raw = onehot ? 0.f : ldg_stream(draw_base(p, d) + offc);
                            cd.add(produced(raw, p * D + d, c), d, nfD, kD);
                        }
                        consume(p, D > 1 ? __fdiv_rn(cd.total(), Df) : cd.total());
                    }
                }
                const float mean = __fdiv_rn(cas.total(), Pf);
                if (c == 0) { best = mean; label = 0; } else argmax_step(mean, c, best, label);
                tu2 = plog2p_acc(tu2, mean);
            }
            const long long o = (long long)b * V + v;
            if (P > 1) {
                Cascade cas;
                cas.reset();
                for (long long p = 0; p < P; ++p) cas.add(h[p * THREADS], p, n_full, level_k);
                const float tu = -(tu2 * kLn2);
                const float au = (-(cas.total() * kLn2)) / Pf;
                u[0] = tu; u[1] = au; u[2] = tu - au;
                if (prm.tu) __stcs(prm.tu + o, u[0]);
                if (prm.au) __stcs(prm.au + o, u[1]);
                if (prm.eu) __stcs(prm.eu + o, u[2]);
            } else {
                u[0] = 1.0f - best;  // test_utils.py:863-864
                if (prm.tu) __stcs(prm.tu + o, u[0]);
            }
            if (prm.lab) prm.lab[o] = (uint8_t)label;
            if (want_ml)
                for (long long p = 0; p < P; ++p) prm.mlab[(p * prm.B + b) * V + v] = bi_smem[p * THREADS + threadIdx.x];
        }
        if (do_stats) {
            const float u1[VU_N_UNC][1] = {{u[0]}, {u[1]}, {u[2]}};
            const int l1[1] = {label};
            stats_tile<1, THREADS>(prm.st, st_smem, active, b, v, u1, l1);
        }
    }
    if (do_stats && t1 > t0) cursor.finish(prm.st, st_smem, vt, THREADS);
}

typedef void (*K1Kernel_t)(const K1Params);
// ---------------------------------------------------------------------------
// Class-outer kernel: ANY class count at run time, P <= PMAX members (compile time), unit voxel stride.
// The fast / TMA forms above are compiled for C = 2, 3, 4, 19; every other class count used to fall to the generic kernel
// (one scalar load in flight per thread, shared-memory read-modify-write per element: 0.5 TB/s, r02u).  Here a thread owns VEC
// consecutive voxels and walks the classes in the OUTER loop: the P values of one class are P independent loads (8 at a
// time in flight), their cascade sum is finished in registers (the mean of that class: true division, argmax step, TU term)
// and the per-member entropy sums live in registers too, indexed at compile time because the member loop is fully unrolled
// up to PMAX.  Every per-voxel operation and its order are those of k1_core.cuh, so maps and labels are bit-identical with
// the other forms (tests/test_gpu_parity.py).
// ---------------------------------------------------------------------------
template <int VEC, int PMAX, int THREADS, bool MP, bool STATS>
__global__ void __launch_bounds__(THREADS, (!STATS && PMAX <= 16) ? 4 : 1) k1_classouter(const __grid_constant__ K1Params prm) {
    static_assert(VEC == 1 || VEC == 2, "one or two voxels per thread");
    constexpr int CH = 8;       // members whose loads are in flight together
    constexpr long long kTileVox = (long long)THREADS * VEC;
    StatsCursor<THREADS> cursor;
    if (STATS) stats_init<THREADS>(prm.st, vu_dyn_smem);
    const int P = (int)prm.P, C = (int)prm.C;
    const long long V = prm.V;
    const float Pf = (float)P;
    const bool two = P > 17;  // torch's cascade: chunks of 16 members folded into a second accumulator (k1_core.cuh)
    const int t0 = (int)(prm.total_tiles * (long long)blockIdx.x / gridDim.x);
    const int t1 = (int)(prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x);
    const int tpi = (int)prm.tiles_per_img;
    int b = t0 / tpi, vt = t0 - b * tpi - 1;
    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        const long long v = (long long)vt * kTileVox + (long long)threadIdx.x * VEC;
        if (STATS) cursor.enter(prm.st, vu_dyn_smem, b, vt, kTileVox);
        const bool active = v < V;
        float u[VU_N_UNC][VEC];
        int label[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { u[0][k] = u[1][k] = u[2][k] = 0.f; label[k] = 0; }
        if (active) {
            if (STATS) stats_prefetch_gt<VEC>(prm.st, b, v);
            const long long off0 = (long long)b * prm.sb + v;
            // (member bases are formed per load: 2 x PMAX registers of pointers would halve the occupancy)
            long long sp = prm.sp;
            auto base = [&](int p) { return (MP ? ld_member_ptr(prm.mptr, p) : prm.x + (long long)p * sp) + off0; };
            float h[PMAX][VEC];      // per-member entropy sums (log2 units)
            f32x2 hp[VEC == 2 ? PMAX : 1];  // ... as packed pairs while the classes stream (VEC == 2: FADD2 / FFMA2, as k1_core.cuh)
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
#pragma unroll
                for (int k = 0; k < VEC; ++k) h[p][k] = 0.f;
#pragma unroll
            for (int p = 0; p < (VEC == 2 ? PMAX : 1); ++p) hp[p] = 0ull;
            float best[VEC], tu2[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) { best[k] = 0.f; tu2[k] = 0.f; }
            for (int c = 0; c < C; ++c) {
                const long long oc = (long long)c * prm.sc;
                asm volatile("" : "+l"(sp));  // (keeps the compiler from holding p * stride_p for every p in registers: occupancy)
                float m0[VEC], m1[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) { m0[k] = 0.f; m1[k] = 0.f; }
                // (requesting class c + 1 before class c is consumed was measured: the registers it takes cost more occupancy
                //  than the loads in flight gain -- 1.3 TB/s against 1.6)
                if constexpr (VEC == 2) {
                    f32x2 M0 = 0ull, M1 = 0ull;
#pragma unroll
                    for (int p0 = 0; p0 < PMAX; p0 += CH) {
                        if (p0 < P) {
                            f32x2 X[CH];
#pragma unroll
                            for (int j = 0; j < CH; ++j)
                                if (p0 + j < PMAX && p0 + j < P) PairLoad<2>::load(base(p0 + j) + oc, &X[j]);
#pragma unroll
                            for (int j = 0; j < CH; ++j) {
                                const int p = p0 + j;
                                if (p < PMAX && p < P) {
                                    M0 = add2(M0, X[j]);
                                    f32x2 PC, L;
                                    plog2p_parts2(X[j], PC, L);
                                    hp[p] = fma2(PC, L, hp[p]);
                                    if ((p & 15) == 15 && two) { M1 = add2(M1, M0); M0 = 0ull; }
                                }
                            }
                        }
                    }
                    upk2(M0, m0[0], m0[VEC - 1]);
                    upk2(M1, m1[0], m1[VEC - 1]);
                } else {
#pragma unroll
                    for (int p0 = 0; p0 < PMAX; p0 += CH) {
                        if (p0 < P) {
                            float x[CH][VEC];
#pragma unroll
                            for (int j = 0; j < CH; ++j)
                                if (p0 + j < PMAX && p0 + j < P) VecLoad<VEC>::load(base(p0 + j) + oc, x[j]);
#pragma unroll
                            for (int j = 0; j < CH; ++j) {
                                const int p = p0 + j;
                                if (p < PMAX && p < P) {
#pragma unroll
                                    for (int k = 0; k < VEC; ++k) {
                                        m0[k] = __fadd_rn(m0[k], x[j][k]);
                                        h[p][k] = plog2p_acc(h[p][k], x[j][k]);
                                    }
                                    if ((p & 15) == 15 && two) {
#pragma unroll
                                        for (int k = 0; k < VEC; ++k) { m1[k] = __fadd_rn(m1[k], m0[k]); m0[k] = 0.f; }
                                    }
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float sum = two ? __fadd_rn(m0[k], m1[k]) : m0[k];
                    const float mean = __fdiv_rn(sum, Pf);  // test_2D.py:971
                    if (c == 0) { best[k] = mean; label[k] = 0; } else argmax_step(mean, c, best[k], label[k]);
                    tu2[k] = plog2p_acc(tu2[k], mean);
                }
            }
            if constexpr (VEC == 2) {
#pragma unroll
                for (int p = 0; p < PMAX; ++p) upk2(hp[p], h[p][0], h[p][VEC - 1]);
            }
            // AU: the mean over the members of their entropies, in the same cascade order
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int p = 0; p < PMAX; ++p) {
                    if (p < P) {
                        a0 = __fadd_rn(a0, h[p][k]);
                        if ((p & 15) == 15 && two) { a1 = __fadd_rn(a1, a0); a0 = 0.f; }
                    }
                }
                const float asum = two ? __fadd_rn(a0, a1) : a0;
                const float tu = -(tu2[k] * kLn2);
                const float au = __fdiv_rn(-(asum * kLn2), Pf);
                u[0][k] = tu; u[1][k] = au; u[2][k] = tu - au;
            }
            const long long o = (long long)b * V + v;
            if (prm.tu) VecLoad<VEC>::store(prm.tu + o, u[0]);
            if (prm.au) VecLoad<VEC>::store(prm.au + o, u[1]);
            if (prm.eu) VecLoad<VEC>::store(prm.eu + o, u[2]);
            if (prm.lab) VecLoad<VEC>::store_u8(prm.lab + o, label);
        }
        if (STATS) stats_tile<VEC, THREADS>(prm.st, vu_dyn_smem, active, b, v, u, label);
    }
    if (STATS && t1 > t0) cursor.finish(prm.st, vu_dyn_smem, vt, kTileVox);
}

struct ClassOuterVariant {
    int VEC, PMAX;
    K1Kernel_t fn[2][2];  // [member pointer list][statistics]
};

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef void (*K1Kernel)(const K1Params);
struct FastVariant {
    int C, VEC, LEVELS, THREADS, MINB, G, DB;
    int use;            // automatic selection: 0 = maps-only launches, 1 = launches with statistics, 2 = both, -1 = never
    K1Kernel fn;        // maps + labels only
    K1Kernel fn_stats;  // with the per-image statistics phase
};

#define VU_VARIANT(C, VEC, LEVELS, THREADS, MINB, G, DB, USE)                              \
    { C, VEC, LEVELS, THREADS, MINB, G, DB, USE,                                           \
      (K1Kernel)k1_fast<C, VEC, LEVELS, THREADS, MINB, G, false>,                          \
      (K1Kernel)k1_fast<C, VEC, LEVELS, THREADS, MINB, G, true> }

// First entry matching (C, LEVELS, alignment) wins unless the "k1_variant"
// option selects another by index (used by the tuning sweep, bench/sweep_k1.py).
static const FastVariant kFast[] = {
    // ---- C = 2 (LIDC-like, toy): few loads per member -> group members
    VU_VARIANT(2, 4, 1, 256, 3, 4, 0, 0),    // 0
    VU_VARIANT(2, 4, 2, 256, 3, 4, 0, 0),    // 1
    VU_VARIANT(2, 4, 1, 256, 2, 4, 0, 1),    // 2
    VU_VARIANT(2, 4, 2, 256, 2, 4, 0, 1),    // 3
    VU_VARIANT(2, 4, 1, 256, 2, 8, 0, -1),   // 4
    VU_VARIANT(2, 4, 2, 256, 2, 8, 0, -1),   // 5
    VU_VARIANT(2, 2, 1, 256, 2, 8, 0, 2),    // 6
    VU_VARIANT(2, 2, 2, 256, 2, 8, 0, 2),    // 7
    VU_VARIANT(2, 1, 1, 256, 2, 8, 0, 2),    // 8  (unaligned V)
    VU_VARIANT(2, 1, 2, 256, 2, 8, 0, 2),    // 9
    // ---- C = 19 (Cityscapes / GTA): 19 independent loads per member
    VU_VARIANT(19, 2, 1, 256, 2, 1, 0, 2),   // 10
    VU_VARIANT(19, 2, 2, 256, 1, 1, 0, 2),   // 11
    VU_VARIANT(19, 1, 1, 256, 3, 1, 0, 2),   // 12 (unaligned V)
    VU_VARIANT(19, 1, 2, 256, 2, 1, 0, 2),   // 13
    VU_VARIANT(19, 2, 1, 128, 4, 1, 0, -1),  // 14
    // ---- small C
    VU_VARIANT(3, 4, 1, 256, 2, 2, 0, 2),    // 15
    VU_VARIANT(3, 4, 2, 256, 2, 2, 0, 2),    // 16
    VU_VARIANT(3, 1, 1, 256, 2, 4, 0, 2),    // 17
    VU_VARIANT(3, 1, 2, 256, 2, 4, 0, 2),    // 18
    VU_VARIANT(4, 4, 1, 256, 2, 2, 0, 2),    // 19
    VU_VARIANT(4, 4, 2, 256, 2, 2, 0, 2),    // 20
    VU_VARIANT(4, 1, 1, 256, 2, 4, 0, 2),    // 21
    VU_VARIANT(4, 1, 2, 256, 2, 4, 0, 2),    // 22
};
static const int kNumFast = (int)(sizeof(kFast) / sizeof(kFast[0]));

static bool aligned_for(const vu_fused_args* a, int vec) {
    if (vec == 1) return true;
    const vu_slab& s = a->slab;
    const uintptr_t bytes = (uintptr_t)vec * 4;
    auto ok = [&](const void* p) { return p == nullptr || ((uintptr_t)p % bytes) == 0; };
    if (!ok(s.data) || !ok(a->tu) || !ok(a->au) || !ok(a->eu)) return false;
    if (s.member_ptrs_host)
        for (int64_t p = 0; p < s.P; ++p)
            if (!ok(s.member_ptrs_host[p])) return false;
    if (a->labels && ((uintptr_t)a->labels % vec) != 0) return false;
    if (a->member_labels && ((uintptr_t)a->member_labels % vec) != 0) return false;
    if (s.V % vec || (!s.member_ptrs && s.stride_p % vec) || s.stride_b % vec || s.stride_c % vec) return false;
    return true;
}

static int level_power_for(long long P) {
    // ATen cascade_sum: max(4, ceil(log2 P) / 4)
    int lg = 0;
    while ((1LL << lg) < P) ++lg;
    return lg / 4 > 4 ? lg / 4 : 4;
}

int launch_k1(const vu_fused_args* a, const StatParams& st, cudaStream_t stream) {
    const vu_slab& s = a->slab;
    K1Params prm;
    prm.x = s.data;
    prm.mptr = s.member_ptrs;
    prm.P = s.P; prm.B = s.B; prm.C = s.C; prm.V = s.V;
    prm.sp = s.stride_p; prm.sb = s.stride_b; prm.sc = s.stride_c; prm.sv = s.stride_v;
    prm.sd = s.stride_d; prm.draws = s.draws > 1 ? s.draws : 1; prm.sflags = s.flags; prm.renorm_eps = s.renorm_eps;
    // plain logits or plain one-hot members (--discretize): TMA form or generic kernel; every other combination of the upstream
    // producers (grouped draws, renormalisation): generic kernel
    const bool lg = prm.draws <= 1 && (s.flags == VU_SLAB_LOGITS || s.flags == VU_SLAB_DISCRETIZE);
    const bool produced = !lg && (prm.draws > 1 || s.flags != 0);
    prm.tu = a->tu; prm.au = a->au; prm.eu = a->eu; prm.lab = a->labels;
    prm.mlab = a->member_labels;
    prm.st = st;

    const int sms = device_sm_count();
    const long long forced = get_option("k1_variant", -1);
    // preferred path: the TMA-pipelined kernel (k1_tma.cu); "k1_path" = 1 keeps to the register-streaming
    // kernels, 2 insists on TMA
    if (forced == -1 && !produced) {
        // few-class slabs with reference-based statistics: the unified-warp TMA form (k1_uni.cu)
        const int rcu = launch_k1_uni(a, st, stream);
        if (rcu <= 0) return rcu;
        const int rc = launch_k1_tma(a, st, stream);
        if (rc <= 0) return rc;
        if (get_option("k1_path", 0) == 2) return set_error(VU_ERR_UNSUPPORTED, "slab not eligible for the TMA path");
    }
    const int need_levels = s.P <= 17 ? 1 : (s.P <= 271 ? 2 : 3);

    const FastVariant* pick = nullptr;
    if (s.stride_v == 1 && s.P >= 2 && need_levels <= 2 && forced != -2 && !produced && !lg) {
        if (forced >= 0 && forced < kNumFast) {
            const FastVariant& f = kFast[forced];
            if (f.C == s.C && f.LEVELS >= need_levels && aligned_for(a, f.VEC)) pick = &f;
            else return set_error(VU_ERR_UNSUPPORTED, "k1_variant does not fit this slab");
        } else {
            for (int i = 0; i < kNumFast && !pick; ++i) {
                const FastVariant& f = kFast[i];
                const bool use_ok = f.use == 2 || f.use == (st.flags ? 1 : 0);
                if (use_ok && f.C == s.C && f.LEVELS == need_levels && aligned_for(a, f.VEC)) pick = &f;
            }
        }
    }

    if (pick) {
        const long long tile_vox = (long long)pick->THREADS * pick->VEC;
        prm.tiles_per_img = (s.V + tile_vox - 1) / tile_vox;
        prm.total_tiles = prm.tiles_per_img * s.B;
        if (prm.total_tiles >= (1LL << 31)) return set_error(VU_ERR_UNSUPPORTED, "more than 2^31 tiles in one launch; split the batch");
        K1Kernel fn = st.flags ? pick->fn_stats : pick->fn;
        const size_t dyn = stats_smem_bytes(st.flags, st.gt.R, pick->THREADS) + stats_class_bytes(st.flags, st.gt.R, st.ncls);
        if (dyn > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
            return set_cuda_error("cudaFuncSetAttribute(k1_fast)");
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, pick->THREADS, dyn) != cudaSuccess || occ < 1)
            return set_cuda_error("occupancy query (k1_fast)");
        long long grid = (long long)sms * occ;
        if (grid > prm.total_tiles) grid = prm.total_tiles;
        fn<<<(unsigned)grid, pick->THREADS, dyn, stream>>>(prm);
        count_launch("k1_fast");
        return check_launch("k1_fast");
    }

    // any other class count (and whatever else has no compiled-in form): the class-outer kernels, P <= 32, unit voxel stride --
    // rows through the TMA ring when they are 16-byte aligned and no statistics are asked for (k1_co_tma.cu), else direct loads
    if (forced == -1 && !produced && !lg && s.C != 2 && s.C != 3 && s.C != 4 && s.C != 19) {
        const int rc = launch_k1_co_tma(a, st, stream);
        if (rc <= 0) return rc;
    }
    if (s.stride_v == 1 && s.P >= 2 && s.P <= 32 && !produced && !lg && forced != -2 && !a->member_labels) {
        constexpr int T = 256;
#define VU_CO(VEC, PMAX)                                                                                              \
    { VEC, PMAX, { { (K1Kernel_t)k1_classouter<VEC, PMAX, T, false, false>, (K1Kernel_t)k1_classouter<VEC, PMAX, T, false, true> },  \
                   { (K1Kernel_t)k1_classouter<VEC, PMAX, T, true, false>, (K1Kernel_t)k1_classouter<VEC, PMAX, T, true, true> } } }
        static const ClassOuterVariant kCo[] = {VU_CO(2, 8), VU_CO(2, 16), VU_CO(2, 32), VU_CO(1, 8), VU_CO(1, 16), VU_CO(1, 32)};
#undef VU_CO
        const int vec = aligned_for(a, 2) ? 2 : 1;
        const ClassOuterVariant* co = nullptr;
        for (const ClassOuterVariant& c : kCo)
            if (!co && c.VEC == vec && c.PMAX >= s.P) co = &c;
        const long long tile_vox = (long long)T * vec;
        prm.tiles_per_img = (s.V + tile_vox - 1) / tile_vox;
        prm.total_tiles = prm.tiles_per_img * s.B;
        if (prm.total_tiles >= (1LL << 31)) return set_error(VU_ERR_UNSUPPORTED, "more than 2^31 tiles in one launch; split the batch");
        K1Kernel_t fn = co->fn[s.member_ptrs ? 1 : 0][st.flags ? 1 : 0];
        const size_t dyn = stats_smem_bytes(st.flags, st.gt.R, T) + stats_class_bytes(st.flags, st.gt.R, st.ncls);
        if (dyn <= 200 * 1024) {
            if (dyn > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
                return set_cuda_error("cudaFuncSetAttribute(k1_classouter)");
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, T, dyn) != cudaSuccess || occ < 1)
                return set_cuda_error("occupancy query (k1_classouter)");
            long long grid = (long long)sms * occ;
            if (grid > prm.total_tiles) grid = prm.total_tiles;
            fn<<<(unsigned)grid, T, dyn, stream>>>(prm);
            count_launch("k1_classouter");
            return check_launch("k1_classouter");
        }
    }

    // generic path
    constexpr int T = 64;
    const size_t max_dyn = 160 * 1024;
    size_t dyn = generic_smem_bytes(s.P, T, a->member_labels != nullptr, prm.draws, s.flags) + stats_smem_bytes(st.flags, st.gt.R, T) +
                 stats_class_bytes(st.flags, st.gt.R, st.ncls);
    if (dyn > max_dyn) return set_error(VU_ERR_UNSUPPORTED, "P (x draws) too large for the generic kernel (P*256 B of shared memory, P*576 B with member labels, + P*draws*320 B for renormalised / discretised draws)");
    if (s.P > (1LL << 19)) return set_error(VU_ERR_UNSUPPORTED, "P > 2^19");
    // the attribute is per device (and this may be the first launch on this one): set it whenever it is needed
    if (dyn > 48 * 1024 && cudaFuncSetAttribute(k1_generic<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(k1_generic)");
    prm.tiles_per_img = (s.V + T - 1) / T;
    prm.total_tiles = prm.tiles_per_img * s.B;
    if (prm.total_tiles >= (1LL << 31)) return set_error(VU_ERR_UNSUPPORTED, "more than 2^31 tiles in one launch; split the batch");
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k1_generic<T>, T, dyn) != cudaSuccess || occ < 1)
        return set_cuda_error("occupancy query (k1_generic)");
    long long grid = (long long)sms * occ;
    if (grid > prm.total_tiles) grid = prm.total_tiles;
    k1_generic<T><<<(unsigned)grid, T, dyn, stream>>>(prm, level_power_for(s.P));
    count_launch("k1_generic");
    return check_launch("k1_generic");
}

int num_fast_variants() { return kNumFast; }
int describe_fast_variant(int i, int* out7) {
    if (i < 0 || i >= kNumFast) return -1;
    const FastVariant& f = kFast[i];
    out7[0] = f.C; out7[1] = f.VEC; out7[2] = f.LEVELS; out7[3] = f.THREADS; out7[4] = f.MINB; out7[5] = f.G; out7[6] = f.DB;
    return 0;
}

}  // namespace vu
