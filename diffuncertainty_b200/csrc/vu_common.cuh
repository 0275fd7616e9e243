// Shared device helpers for libvalunc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <type_traits>

#include "valunc.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvalunc is written for sm_100a (B200) only"
#endif

namespace vu {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- streaming loads: read-once data, keep it out of L1 --------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_stream(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

template <int VEC>
struct VecLoad;
template <>
struct VecLoad<4> {
    __device__ __forceinline__ static void load(const float* p, float (&x)[4]) {
        float4 v = ldg_stream(reinterpret_cast<const float4*>(p));
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
    __device__ __forceinline__ static void store(float* p, const float (&x)[4]) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3]));
    }
    __device__ __forceinline__ static void store_u8(uint8_t* p, const int (&l)[4]) {
        uchar4 v = make_uchar4((unsigned char)l[0], (unsigned char)l[1], (unsigned char)l[2], (unsigned char)l[3]);
        __stcs(reinterpret_cast<uchar4*>(p), v);
    }
};
template <>
struct VecLoad<2> {
    __device__ __forceinline__ static void load(const float* p, float (&x)[2]) {
        float2 v = ldg_stream(reinterpret_cast<const float2*>(p));
        x[0] = v.x; x[1] = v.y;
    }
    __device__ __forceinline__ static void store(float* p, const float (&x)[2]) {
        __stcs(reinterpret_cast<float2*>(p), make_float2(x[0], x[1]));
    }
    __device__ __forceinline__ static void store_u8(uint8_t* p, const int (&l)[2]) {
        uchar2 v = make_uchar2((unsigned char)l[0], (unsigned char)l[1]);
        *reinterpret_cast<uchar2*>(p) = v;
    }
};
template <>
struct VecLoad<1> {
    __device__ __forceinline__ static void load(const float* p, float (&x)[1]) { x[0] = ldg_stream(p); }
    __device__ __forceinline__ static void store(float* p, const float (&x)[1]) { __stcs(p, x[0]); }
    __device__ __forceinline__ static void store_u8(uint8_t* p, const int (&l)[1]) { *p = (uint8_t)l[0]; }
};

// ---- p * log2(p) with the reference's skip rule ---------------------------
// test_utils.py:838-840 / 849-851 drop every NaN product, i.e. the term is
// p*log(p) for p > 0 and 0 for p == 0, p < 0 and NaN.  One MUFU.LG2 per
// element; its absolute error (~2^-23.5 on [0.5, 2], measured on B200) is too
// large next to p = 1 (confident pixels: the term is ~ -(1-p)), so within
// |p-1| < 1/16 log2(p) = f * q(f), f = p - 1, with q the degree-4 near-minimax
// fit of log2(1+f)/f on [-1/16, 1/16] (relative error of p*log2(p) < 1.8e-7
// over every float in the window; outside it MUFU.LG2 stays below 1e-6).
// Subnormal p (< FLT_MIN) is treated as 0: its true term is < 1.1e-36.
// Result is in log2 units; callers multiply the class sum by ln 2 once.
__device__ __forceinline__ float plog2p(float p) {
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(p));
    const float f = p - 1.0f;
    float t = fmaf(f, 0.28995341062545776f, -0.36185145378112793f);
    t = fmaf(t, f, 0.4808957874774933f);
    t = fmaf(t, f, -0.721346378326416f);
    t = fmaf(t, f, 1.4426950216293335f);
    const float near_one = t * f;
    const float l = (fabsf(f) < (1.0f / 16.0f)) ? near_one : lg;
    const float term = __fmul_rn(p, l);  // never contracted into the caller's add: all kernels agree bitwise
    return (p >= FLT_MIN) ? term : 0.0f;
}

// torch.argmax tie rule (test_2D.py:817,871): first maximal index, NaN is max.
__device__ __forceinline__ void argmax_step(float v, int c, float& best, int& idx) {
    const bool take = (v > best) || ((v != v) && (best == best));
    best = take ? v : best;
    idx = take ? c : idx;
}

// ---- warp reductions --------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(kFull, v); }

// ---- ground truth view ------------------------------------------------------
struct GtView {
    const void* data;
    int dtype;
    int R;
    long long sb, sr, sv;
    int has_ignore;
    long long ignore;
    int align;  // largest power of two <= 4 such that VEC = align voxels can be fetched with one vector load
};

template <typename T>
__device__ __forceinline__ T ldg_gt(const T* p) { return __ldg(p); }

// references of `VEC` consecutive voxels of rater r (values widened to int for uint8, kept 64-bit for int64)
template <int VEC>
__device__ __forceinline__ void load_gt(const GtView& gt, long long b, int r, long long v, int (&g)[VEC], uint8_t) {
    const uint8_t* base = reinterpret_cast<const uint8_t*>(gt.data) + b * gt.sb + (long long)r * gt.sr;
    if (VEC == 4 && gt.align >= 4) {
        const unsigned w = __ldg(reinterpret_cast<const unsigned*>(base + v));
#pragma unroll
        for (int k = 0; k < VEC; ++k) g[k] = (int)((w >> (8 * k)) & 0xffu);
    } else if (VEC == 2 && gt.align >= 2) {
        const unsigned short w = __ldg(reinterpret_cast<const unsigned short*>(base + v));
#pragma unroll
        for (int k = 0; k < VEC; ++k) g[k] = (int)((w >> (8 * k)) & 0xffu);
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) g[k] = (int)__ldg(base + (v + k) * gt.sv);
    }
}
template <int VEC>
__device__ __forceinline__ void load_gt(const GtView& gt, long long b, int r, long long v, long long (&g)[VEC], long long) {
    const long long* base = reinterpret_cast<const long long*>(gt.data) + b * gt.sb + (long long)r * gt.sr;
    if (VEC >= 2 && gt.align >= 2) {
#pragma unroll
        for (int k = 0; k < VEC; k += 2) {
            const longlong2 w = __ldg(reinterpret_cast<const longlong2*>(base + v + k));
            g[k] = w.x;
            if (k + 1 < VEC) g[k + 1] = w.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) g[k] = __ldg(base + (v + k) * gt.sv);
    }
}

struct CalibDev {
    float a, b;
    float edge[VU_N_EDGES];  // thresholds on u (sign-flipped when conf falls with u); NaN = never reached
    int increasing;
    int identity;  // the map already is the confidence
};

struct StatParams {
    unsigned flags;
    unsigned unc_mask;  // bit k set: uncertainty type k is present (TU/AU/EU; bit 0 only when P == 1)
    long long V;
    GtView gt;
    float thr[VU_N_UNC];
    CalibDev calib[VU_N_UNC];
    const uint8_t* lut;
    const double* ncc_gt_map;
    double* f64;
    long long* i64;
};

// ---- per-CTA statistics state (dynamic shared memory) ---------------------------
// Every thread owns one private column of float64 / packed-integer accumulators
// ("slots", layout [slot][thread]: conflict-free, no atomics, no shuffles while
// streaming).  The calibration histograms are one CTA-wide copy with 32 replicas,
// one per lane ([unc][bin][lane]: a warp's 32 adds never share an address or a
// bank, whatever the bins are), updated with native 32-bit shared-memory atomic
// adds (warps share the replicas).  flush() folds
// everything into the image's row of the global statistics buffers with one
// atomic per non-zero column; it runs when a CTA moves on to another image.
//
// float64 slots                       packed-integer slots (uint64)
//   0..2   sum of u_k                   0      #(u_k >= t_k), 3 x 21 bit
//   3..5   sum of u_k over u_k >= t_k   1      #(label > 0)
//   6..8   sum of conf in bin 0         2+r    rater r: tp | pred << 21 | gt << 42
//   9      sum g        10  sum g*g
//   11..13 sum u_k^2    14..16  sum g*u_k
enum { FS_SUM = 0, FS_THR = 3, FS_BIN0 = 6, FS_G = 9, FS_GG = 10, FS_UU = 11, FS_GU = 14, FS_MAX = 17 };
enum { IS_THRCNT = 0, IS_AREA = 1, IS_DICE = 2, IS_MAX = 2 + VU_MAX_RATERS };
constexpr int kPackBits = 21;
constexpr unsigned long long kPackMask = (1ull << kPackBits) - 1;
constexpr long long kMaxTilesPerFlush = 1LL << 13;  // bounds every packed / 32-bit counter between two flushes
constexpr int kQBits = 24;                           // fixed point of (conf - bin * 0.05): |q| < 2^20, x n_valid <= 8
constexpr int kQSplit = 12;                          // q is accumulated as (q >> 12, q & 0xfff) in two 32-bit words
constexpr int kEdgePad = 24;                         // E[0] = NaN, E[1..19] = edges, E[20..23] = NaN

// histogram words: [4 planes: total, true, q_hi, q_lo][VU_N_UNC][VU_N_BINS][32 lanes] int32
constexpr int kHistPlane = VU_N_UNC * VU_N_BINS * 32;
constexpr int kHistWords = 4 * kHistPlane;

__host__ __device__ inline int stats_num_fslots(unsigned flags) {
    return (flags & VU_STAT_NCC) ? FS_MAX : ((flags & VU_STAT_CALIB) ? FS_G : FS_BIN0);
}
__host__ __device__ inline int stats_num_islots(unsigned flags, int R) { return (flags & VU_STAT_DICE) ? IS_DICE + R : IS_DICE; }
__host__ __device__ inline size_t stats_smem_bytes(unsigned flags, int R, int threads) {
    if (!flags) return 0;
    size_t n = (size_t)(stats_num_fslots(flags) + stats_num_islots(flags, R)) * threads * 8;
    if (flags & VU_STAT_CALIB) n += (size_t)kHistWords * sizeof(int) + (size_t)VU_N_UNC * kEdgePad * sizeof(float);
    return n;
}

template <int THREADS>
struct StatsLayout {
    double* fs;
    unsigned long long* is;
    int* hist;
    float* E;
    int nF, nI;
    __device__ __forceinline__ StatsLayout(const StatParams& sp, void* smem) {
        nF = stats_num_fslots(sp.flags);
        nI = stats_num_islots(sp.flags, sp.gt.R);
        fs = reinterpret_cast<double*>(smem);
        is = reinterpret_cast<unsigned long long*>(fs + (size_t)nF * THREADS);
        hist = reinterpret_cast<int*>(is + (size_t)nI * THREADS);
        E = reinterpret_cast<float*>(hist + kHistWords);
    }
};

template <int THREADS>
__device__ __noinline__ void stats_init(const StatParams& sp, void* smem) {
    StatsLayout<THREADS> L(sp, smem);
    for (int t = threadIdx.x; t < (L.nF + L.nI) * THREADS; t += THREADS) reinterpret_cast<unsigned long long*>(smem)[t] = 0ull;
    if (sp.flags & VU_STAT_CALIB) {
        for (int t = threadIdx.x; t < kHistWords; t += THREADS) L.hist[t] = 0;
        for (int t = threadIdx.x; t < VU_N_UNC * kEdgePad; t += THREADS) {
            const int k = t / kEdgePad, e = t % kEdgePad;
            L.E[t] = (e >= 1 && e <= VU_N_EDGES) ? sp.calib[k].edge[e - 1] : __int_as_float(0x7fc00000);
        }
    }
    __syncthreads();
}

// Add this CTA's partials into image row b and clear them.  nvox = voxels of image b
// the CTA went through since the last flush.  Called by every thread of the CTA.
template <int THREADS>
__device__ __noinline__ void stats_flush(const StatParams& sp, void* smem, long long b, long long nvox) {
    __syncthreads();
    StatsLayout<THREADS> L(sp, smem);
    constexpr int WARPS = THREADS / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned flags = sp.flags;
    double* frow = sp.f64 + b * VU_F64_COLS;
    unsigned long long* irow = reinterpret_cast<unsigned long long*>(sp.i64 + b * VU_I64_COLS);
    for (int q = warp; q < L.nF; q += WARPS) {
        double s = 0.0;
        for (int i = lane; i < THREADS; i += 32) { s += L.fs[q * THREADS + i]; L.fs[q * THREADS + i] = 0.0; }
        s = warp_sum(s);
        if (lane == 0 && s != 0.0) {
            if (q < FS_THR) {
                if (flags & VU_STAT_IMAGE_SUM) atomicAdd(frow + VU_F64_SUM + q, s);
                if (flags & VU_STAT_NCC) atomicAdd(frow + VU_F64_NCC_U + q, s);
            } else if (q < FS_BIN0) atomicAdd(frow + VU_F64_THR_SUM + (q - FS_THR), s);
            else if (q < FS_G) atomicAdd(frow + VU_F64_BIN_SUMS + (q - FS_BIN0) * VU_N_BINS, s);
            else if (q == FS_G) atomicAdd(frow + VU_F64_NCC_G, s);
            else if (q == FS_GG) atomicAdd(frow + VU_F64_NCC_GG, s);
            else if (q < FS_GU) atomicAdd(frow + VU_F64_NCC_UU + (q - FS_UU), s);
            else atomicAdd(frow + VU_F64_NCC_GU + (q - FS_GU), s);
        }
    }
    for (int q = warp; q < L.nI; q += WARPS) {
        unsigned long long f0 = 0, f1 = 0, f2 = 0;
        for (int i = lane; i < THREADS; i += 32) {
            const unsigned long long w = L.is[q * THREADS + i];
            L.is[q * THREADS + i] = 0ull;
            if (q == IS_AREA) f0 += w;
            else { f0 += w & kPackMask; f1 += (w >> kPackBits) & kPackMask; f2 += w >> (2 * kPackBits); }
        }
        f0 = (unsigned long long)warp_sum((double)f0);  // exact: every partial is far below 2^53
        f1 = (unsigned long long)warp_sum((double)f1);
        f2 = (unsigned long long)warp_sum((double)f2);
        if (lane == 0) {
            if (q == IS_THRCNT) {
                if (f0) atomicAdd(irow + VU_I64_THR_COUNT + 0, f0);
                if (f1) atomicAdd(irow + VU_I64_THR_COUNT + 1, f1);
                if (f2) atomicAdd(irow + VU_I64_THR_COUNT + 2, f2);
            } else if (q == IS_AREA) {
                if (f0) atomicAdd(irow + VU_I64_AREA, f0);
            } else {
                const int r = q - IS_DICE;
                if (f0) atomicAdd(irow + VU_I64_DICE_TP + r, f0);
                if (f1) atomicAdd(irow + VU_I64_DICE_PRED + r, f1);
                if (f2) atomicAdd(irow + VU_I64_DICE_GT + r, f2);
            }
        }
    }
    if (flags & VU_STAT_CALIB) {
        for (int t = threadIdx.x; t < VU_N_UNC * VU_N_BINS; t += THREADS) {
            const int k = t / VU_N_BINS, bin = t % VU_N_BINS;
            long long tot = 0, tru = 0, q = 0;
            int* h = L.hist + t * 32;
            for (int i = 0; i < 32; ++i) {
                const int r = (i + threadIdx.x) & 31;  // skewed: the threads of a warp read different banks
                tot += h[r]; tru += h[kHistPlane + r];
                q += (long long)h[2 * kHistPlane + r] * (1 << kQSplit) + h[3 * kHistPlane + r];
                h[r] = 0; h[kHistPlane + r] = 0; h[2 * kHistPlane + r] = 0; h[3 * kHistPlane + r] = 0;
            }
            if (tot) {
                atomicAdd(irow + VU_I64_BIN_TOTAL + t, (unsigned long long)tot);
                if (tru) atomicAdd(irow + VU_I64_BIN_TRUE + t, (unsigned long long)tru);
                // bin 0 is summed in floating point by the threads (slots FS_BIN0); slot 20 only ever holds NaN
                // confidences (np.digitize sends NaN past the last edge), whose sum is NaN
                if (bin == VU_N_BINS - 1) atomicAdd(frow + VU_F64_BIN_SUMS + t, (double)__int_as_float(0x7fc00000));
                else if (bin > 0)
                    atomicAdd(frow + VU_F64_BIN_SUMS + t,
                              (double)tot * (double)((float)bin * 0.05f) + (double)q * (1.0 / (double)(1 << kQBits)));
            }
        }
    }
    if (threadIdx.x == 0 && nvox) atomicAdd(irow + VU_I64_NVOX, (unsigned long long)nvox);
    __syncthreads();
}

// Tracks which image a CTA is working on and flushes when it moves on.  All members are CTA-uniform.
template <int THREADS>
struct StatsCursor {
    int cur_b, vt_begin;
    __device__ __forceinline__ StatsCursor() : cur_b(-1), vt_begin(0) {}
    // top of every tile: tile index vt inside image b, tile_vox voxels per tile
    __device__ __forceinline__ void enter(const StatParams& sp, void* smem, int b, int vt, long long tile_vox) {
        if (b != cur_b || vt - vt_begin >= (int)kMaxTilesPerFlush) {
            if (cur_b >= 0) {
                const long long end = (b != cur_b) ? sp.V : (long long)vt * tile_vox;
                stats_flush<THREADS>(sp, smem, cur_b, end - (long long)vt_begin * tile_vox);
            }
            cur_b = b;
            vt_begin = vt;
        }
    }
    // after the last tile (vt = its index inside the image)
    __device__ __forceinline__ void finish(const StatParams& sp, void* smem, int last_vt, long long tile_vox) {
        if (cur_b >= 0) {
            long long end = (long long)(last_vt + 1) * tile_vox;
            end = end > sp.V ? sp.V : end;
            stats_flush<THREADS>(sp, smem, cur_b, end - (long long)vt_begin * tile_vox);
        }
    }
};

// ace.py:329 in float32: 1 / (1 + exp((-u) * a + b)); ace.py:333 clips to [0, 1].  Bins are decided on u
// itself (calib_bin), so this value only feeds the floating bin_sums: fast exp / reciprocal are enough.
__device__ __forceinline__ float platt_conf(float u, float a, float b, int identity) {
    if (identity) return fminf(fmaxf(u, 0.0f), 1.0f);
    const float z = fmaf(-u, a, b);
    const float c = __fdividef(1.0f, 1.0f + __expf(z));
    return fminf(fmaxf(c, 0.0f), 1.0f);
}

// bin = number of interior edges e_k with conf(u) >= e_k, decided on u itself
// (see vu_calib in valunc.h).  The device confidence is within a few ulp of the
// reference's, so floor(conf * 20) is at most one bin off; the two neighbouring
// thresholds on u settle it exactly.  E is padded with NaN (compares false).
__device__ __forceinline__ int calib_bin(const float* E, int increasing, float u, float conf) {
    const float uu = increasing ? u : -u;
    int k0 = __float2int_rd(conf * 20.0f);
    k0 = k0 > 19 ? 19 : (k0 < 0 ? 0 : k0);
    const float e_lo = E[k0], e_hi = E[k0 + 1];  // E[j] = interior edge j-1, i.e. the lower edge of bin j
    return k0 + (uu >= e_hi ? 1 : 0) - (uu < e_lo ? 1 : 0);
}

// The whole statistics phase for the VEC voxels a thread owns in one tile.  Must
// be called by every thread of the CTA (threads past the end of the image pass
// active = false).  Kept out of line so its registers do not inflate the
// streaming loop of the calling kernel.
template <int VEC> struct FVec;
template <> struct FVec<4> { typedef float4 type; };
template <> struct FVec<2> { typedef float2 type; };
template <> struct FVec<1> { typedef float type; };
__device__ __forceinline__ float4 fvec_pack(const float (&x)[4]) { return make_float4(x[0], x[1], x[2], x[3]); }
__device__ __forceinline__ float2 fvec_pack(const float (&x)[2]) { return make_float2(x[0], x[1]); }
__device__ __forceinline__ float fvec_pack(const float (&x)[1]) { return x[0]; }
__device__ __forceinline__ void fvec_unpack(float4 v, float (&x)[4]) { x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
__device__ __forceinline__ void fvec_unpack(float2 v, float (&x)[2]) { x[0] = v.x; x[1] = v.y; }
__device__ __forceinline__ void fvec_unpack(float v, float (&x)[1]) { x[0] = v; }

// (arguments travel in registers: the three maps as built-in vectors, the VEC labels packed into one word)
template <int VEC, int THREADS, typename GT>
__device__ __noinline__ void stats_tile_t(const StatParams& sp, void* smem, bool active, long long b, long long v,
                                          typename FVec<VEC>::type u0, typename FVec<VEC>::type u1,
                                          typename FVec<VEC>::type u2, unsigned labels_packed) {
    StatsLayout<THREADS> cs(sp, smem);
    float u[VU_N_UNC][VEC];
    int label[VEC];
    fvec_unpack(u0, u[0]);
    fvec_unpack(u1, u[1]);
    fvec_unpack(u2, u[2]);
#pragma unroll
    for (int j = 0; j < VEC; ++j) label[j] = (int)((labels_packed >> (8 * j)) & 0xffu);
    using G = typename std::conditional<sizeof(GT) == 1, int, long long>::type;
    using GAcc = typename std::conditional<sizeof(GT) == 1, int, double>::type;
    const unsigned flags = sp.flags, mask = sp.unc_mask;
    const int tid = threadIdx.x, lane = tid & 31;

    if (active && (flags & (VU_STAT_IMAGE_SUM | VU_STAT_NCC))) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k)
            if ((mask >> k) & 1) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < VEC; ++j) s += u[k][j];
                cs.fs[(FS_SUM + k) * THREADS + tid] += (double)s;
            }
    }
    if (active && (flags & VU_STAT_THRESHOLD)) {
        unsigned long long packed = 0;
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k)
            if ((mask >> k) & 1) {
                float s = 0.f;
                int n = 0;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const bool hit = u[k][j] >= sp.thr[k];
                    s += hit ? u[k][j] : 0.f;
                    n += hit;
                }
                if (n) cs.fs[(FS_THR + k) * THREADS + tid] += (double)s;
                packed |= (unsigned long long)n << (k * kPackBits);
            }
        if (packed) cs.is[IS_THRCNT * THREADS + tid] += packed;
    }
    if (active && (flags & VU_STAT_AREA)) {
        int n = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) n += (label[j] > 0);
        if (n) cs.is[IS_AREA * THREADS + tid] += (unsigned long long)n;
    }
    if (!(flags & (VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC))) return;

    int n_valid[VEC], n_correct[VEC];
    GAcc gsum[VEC], gsq[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { n_valid[j] = 0; n_correct[j] = 0; gsum[j] = 0; gsq[j] = 0; }
    const int R = sp.gt.data ? sp.gt.R : 0;
    if (active && R > 0) {
        G cmp[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) cmp[j] = (G)(sp.lut ? (int)__ldg(sp.lut + label[j]) : label[j]);
        const bool has_ign = sp.gt.has_ignore != 0;
        const G ign = (G)sp.gt.ignore;
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            G g[VEC];
            load_gt<VEC>(sp.gt, b, r, v, g, GT());
            int tp = 0, ps = 0, gs = 0;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const bool valid = !(has_ign && g[j] == ign);     // ace.py:492-499, test_2D.py:880
                const bool pp = (label[j] == 1) && valid, gp = (g[j] == (G)1) && valid;  // test_2D.py:878-886
                tp += (pp && gp); ps += pp; gs += gp;
                n_valid[j] += valid;
                n_correct[j] += (valid && g[j] == cmp[j]);       // ace.py:488
                gsum[j] += (GAcc)g[j];
                gsq[j] += (GAcc)g[j] * (GAcc)g[j];
            }
            if ((flags & VU_STAT_DICE) && (tp | ps | gs))
                cs.is[(IS_DICE + r) * THREADS + tid] +=
                    (unsigned long long)tp | ((unsigned long long)ps << kPackBits) | ((unsigned long long)gs << (2 * kPackBits));
        }
    }
    if (active && (flags & VU_STAT_NCC)) {
        double sg = 0.0, sgg = 0.0, suu[VU_N_UNC] = {0.0, 0.0, 0.0}, sgu[VU_N_UNC] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            double g;
            if (sp.ncc_gt_map) {
                g = sp.ncc_gt_map[b * sp.V + v + j];
            } else {
                // np.var(refs, axis=0), ddof = 0 (experiment_dataloader.py:283): (R*sum g^2 - (sum g)^2) / R^2
                const double rr = (double)R;
                g = ((double)gsq[j] * rr - (double)gsum[j] * (double)gsum[j]) / (rr * rr);
                g = g < 0.0 ? 0.0 : g;
            }
            sg += g;
            sgg += g * g;
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k)
                if ((mask >> k) & 1) {
                    const double x = (double)u[k][j];
                    suu[k] += x * x;
                    sgu[k] += g * x;
                }
        }
        cs.fs[FS_G * THREADS + tid] += sg;
        cs.fs[FS_GG * THREADS + tid] += sgg;
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k)
            if ((mask >> k) & 1) {
                cs.fs[(FS_UU + k) * THREADS + tid] += suu[k];
                cs.fs[(FS_GU + k) * THREADS + tid] += sgu[k];
            }
    }
    if (flags & VU_STAT_CALIB) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) {
            if (!((mask >> k) & 1)) continue;
            float bin0 = 0.f;
            const CalibDev& cal = sp.calib[k];
            const float* E = cs.E + k * kEdgePad;
            int* hk = cs.hist + k * (VU_N_BINS * 32) + lane;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const int nv = n_valid[j];
                if (active && nv > 0) {
                    const float x = u[k][j];
                    int bin = VU_N_BINS - 1, q = 0;  // np.digitize puts NaN past the last edge
                    if (x == x) {
                        const float conf = platt_conf(x, cal.a, cal.b, cal.identity);
                        bin = calib_bin(E, cal.increasing, x, conf);
                        q = __float2int_rn((conf - (float)bin * 0.05f) * (float)(1 << kQBits)) * nv;
                        bin0 += bin == 0 ? conf * (float)nv : 0.f;
                    }
                    int* h = hk + bin * 32;
                    atomicAdd(h, nv);
                    if (n_correct[j]) atomicAdd(h + kHistPlane, n_correct[j]);
                    atomicAdd(h + 2 * kHistPlane, q >> kQSplit);
                    atomicAdd(h + 3 * kHistPlane, q & ((1 << kQSplit) - 1));
                }
            }
            if (bin0 != 0.f) cs.fs[(FS_BIN0 + k) * THREADS + tid] += (double)bin0;
        }
    }
}

template <int VEC, int THREADS>
__device__ __forceinline__ void stats_tile(const StatParams& sp, void* smem, bool active, long long b, long long v,
                                           const float (&u)[VU_N_UNC][VEC], const int (&label)[VEC]) {
    unsigned lp = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) lp |= (unsigned)(label[j] & 0xff) << (8 * j);
    if (sp.gt.dtype == VU_GT_I64)
        stats_tile_t<VEC, THREADS, long long>(sp, smem, active, b, v, fvec_pack(u[0]), fvec_pack(u[1]), fvec_pack(u[2]), lp);
    else
        stats_tile_t<VEC, THREADS, uint8_t>(sp, smem, active, b, v, fvec_pack(u[0]), fvec_pack(u[1]), fvec_pack(u[2]), lp);
}

}  // namespace vu
