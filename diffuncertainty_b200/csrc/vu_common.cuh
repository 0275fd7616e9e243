// Shared device helpers for libvalunc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

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
// element; its absolute error (2^-22 on [0.5, 2]) is too large next to p = 1
// (confident pixels: the term is ~ -(1-p)), so within |p-1| < 1/64 the log is
// taken from a 4-term log1p series instead (relative error < 2e-8).
// Subnormal p (< FLT_MIN) is treated as 0: its true term is < 1.1e-36.
// Result is in log2 units; callers multiply the class sum by ln 2 once.
__device__ __forceinline__ float plog2p(float p) {
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(p));
    const float f = p - 1.0f;
    float t = fmaf(f, -0.25f * kLog2e, kLog2e / 3.0f);
    t = fmaf(t, f, -0.5f * kLog2e);
    t = fmaf(t, f, kLog2e);
    const float near_one = t * f;
    const float l = (fabsf(f) < (1.0f / 64.0f)) ? near_one : lg;
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
    __device__ __forceinline__ long long at(long long b, int r, long long v) const {
        const long long off = b * sb + (long long)r * sr + v * sv;
        return dtype == VU_GT_U8 ? (long long)__ldg(reinterpret_cast<const uint8_t*>(data) + off)
                                 : __ldg(reinterpret_cast<const long long*>(data) + off);
    }
};

struct CalibDev {
    float a, b;
    float edge[32];  // [0..18] real edges (sign-flipped when decreasing), rest NaN
    int increasing;
    int identity;  // the map already is the confidence
};

struct StatParams {
    unsigned flags;
    int n_unc;  // 3 (TU/AU/EU) or 1 (pred_entropy when P == 1)
    long long V;
    GtView gt;
    float thr[VU_N_UNC];
    CalibDev calib[VU_N_UNC];
    const uint8_t* lut;
    const double* ncc_gt_map;
    double* f64;
    long long* i64;
};

// ---- per-CTA statistics state ------------------------------------------------
// One slot per warp: lane 0 of the warp is the only writer between barriers,
// so no shared-memory atomics are needed.  flush() folds the slots and issues
// one global atomic per non-zero column.
struct WarpSlot {
    double f[VU_F64_COLS];
    long long i[VU_I64_COLS];
};

template <int WARPS>
struct CtaStats {
    WarpSlot slot[WARPS];
    float edges[VU_N_UNC][32];

    __device__ void init(const StatParams& sp) {
        for (int t = threadIdx.x; t < WARPS * VU_F64_COLS; t += blockDim.x) slot[t / VU_F64_COLS].f[t % VU_F64_COLS] = 0.0;
        for (int t = threadIdx.x; t < WARPS * VU_I64_COLS; t += blockDim.x) slot[t / VU_I64_COLS].i[t % VU_I64_COLS] = 0;
        for (int t = threadIdx.x; t < VU_N_UNC * 32; t += blockDim.x) edges[t / 32][t % 32] = sp.calib[t / 32].edge[t % 32];
        __syncthreads();
    }
    // add this CTA's partials into image row b and clear them
    __device__ void flush(const StatParams& sp, long long b) {
        __syncthreads();
        for (int c = threadIdx.x; c < VU_F64_COLS; c += blockDim.x) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { s += slot[w].f[c]; slot[w].f[c] = 0.0; }
            if (s != 0.0) atomicAdd(sp.f64 + b * VU_F64_COLS + c, s);
        }
        for (int c = threadIdx.x; c < VU_I64_COLS; c += blockDim.x) {
            long long s = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { s += slot[w].i[c]; slot[w].i[c] = 0; }
            if (s != 0) atomicAdd(reinterpret_cast<unsigned long long*>(sp.i64 + b * VU_I64_COLS + c), (unsigned long long)s);
        }
        __syncthreads();
    }
};

// Per-thread partials for the voxels a thread owns inside one tile.
struct TileAcc {
    double sum[VU_N_UNC], thr_sum[VU_N_UNC];
    double g, gg, u[VU_N_UNC], uu[VU_N_UNC], gu[VU_N_UNC];
    int thr_cnt[VU_N_UNC];
    int area, nvox;
    int tp[VU_MAX_RATERS], ps[VU_MAX_RATERS], gs[VU_MAX_RATERS];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) { sum[k] = thr_sum[k] = u[k] = uu[k] = gu[k] = 0.0; thr_cnt[k] = 0; }
        g = gg = 0.0;
        area = nvox = 0;
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r) tp[r] = ps[r] = gs[r] = 0;
    }
};

// bin = number of interior edges e_k with conf(u) >= e_k, decided on u itself
// (see vu_calib in valunc.h).  edges[] is padded to 32 with NaN (never true).
__device__ __forceinline__ int calib_bin(const float* edges, int increasing, float u) {
    if (u != u) return VU_N_BINS - 1;  // np.digitize puts NaN past the last edge
    const float uu = increasing ? u : -u;
    int pos = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) pos += (uu >= edges[pos + step - 1]) ? step : 0;
    return pos;
}

// ace.py:329 in float32: 1 / (1 + exp((-u) * a + b)); ace.py:333 clips to [0, 1]
__device__ __forceinline__ float platt_conf(float u, float a, float b, int identity) {
    if (identity) return (u != u) ? u : fminf(fmaxf(u, 0.0f), 1.0f);
    const float z = __fadd_rn(__fmul_rn(-u, a), b);
    const float c = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(z)));
    if (c != c) return c;  // np.clip keeps NaN
    return fminf(fmaxf(c, 0.0f), 1.0f);
}

// Statistics of ONE voxel per lane.  Must be called by all 32 lanes of the warp
// (lanes past the end of the image pass active = false).
template <int WARPS>
__device__ __forceinline__ void stats_voxel(const StatParams& sp, CtaStats<WARPS>& cs, TileAcc& acc, bool active,
                                            long long b, long long v, const float (&u)[VU_N_UNC], int label) {
    const unsigned flags = sp.flags;
    const int lane = threadIdx.x & 31;
    WarpSlot& slot = cs.slot[threadIdx.x >> 5];
    if (active) {
        acc.nvox += 1;
        if (flags & VU_STAT_IMAGE_SUM) {
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k)
                if (k < sp.n_unc) acc.sum[k] += (double)u[k];
        }
        if (flags & VU_STAT_THRESHOLD) {
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k)
                if (k < sp.n_unc && u[k] >= sp.thr[k]) { acc.thr_sum[k] += (double)u[k]; acc.thr_cnt[k] += 1; }
        }
        if (flags & VU_STAT_AREA) acc.area += (label > 0);
    }
    if (!(flags & (VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC))) return;

    int n_valid = 0, n_correct = 0;
    double gmean = 0.0, gsq = 0.0;
    const int R = sp.gt.data ? sp.gt.R : 0;
    if (active && R > 0) {
        const int cmp_label = sp.lut ? (int)__ldg(sp.lut + label) : label;
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r) {
            if (r < R) {
                const long long g = sp.gt.at(b, r, v);
                const bool valid = !(sp.gt.has_ignore && g == sp.gt.ignore);
                if (flags & VU_STAT_DICE) {  // test_2D.py:878-886
                    const bool pp = (label == 1) && valid, gp = (g == 1) && valid;
                    acc.tp[r] += (pp && gp);
                    acc.ps[r] += pp;
                    acc.gs[r] += gp;
                }
                n_valid += valid;                             // ace.py:492-499
                n_correct += (valid && g == (long long)cmp_label);  // ace.py:488
                gmean += (double)g;
                gsq += (double)g * (double)g;
            }
        }
    }
    if (flags & VU_STAT_NCC) {
        if (active) {
            double g;
            if (sp.ncc_gt_map) {
                g = sp.ncc_gt_map[b * sp.V + v];
            } else {
                // np.var(refs, axis=0), ddof = 0 (experiment_dataloader.py:283)
                const double mu = gmean / (double)R;
                g = gsq / (double)R - mu * mu;
                g = g < 0.0 ? 0.0 : g;
            }
            acc.g += g;
            acc.gg += g * g;
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k)
                if (k < sp.n_unc) {
                    const double x = (double)u[k];
                    acc.u[k] += x; acc.uu[k] += x * x; acc.gu[k] += g * x;
                }
        }
    }
    if (flags & VU_STAT_CALIB) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) {
            if (k >= sp.n_unc) break;
            int bin = -1;
            double w = 0.0;
            if (active && n_valid > 0) {
                bin = calib_bin(cs.edges[k], sp.calib[k].increasing, u[k]);
                w = (double)platt_conf(u[k], sp.calib[k].a, sp.calib[k].b, sp.calib[k].identity) * (double)n_valid;
            }
            // one round per distinct bin present in the warp (neighbouring
            // voxels mostly share a bin, so this is 1-2 rounds)
            unsigned todo = __ballot_sync(kFull, bin >= 0);
            while (todo) {
                const int leader = __ffs(todo) - 1;
                const int bsel = __shfl_sync(kFull, bin, leader);
                const bool mine = (bin == bsel);
                const int tot = warp_sum(mine ? n_valid : 0);
                const int tru = warp_sum(mine ? n_correct : 0);
                const double sw = warp_sum(mine ? w : 0.0);
                if (lane == 0) {
                    slot.i[VU_I64_BIN_TOTAL + k * VU_N_BINS + bsel] += tot;
                    slot.i[VU_I64_BIN_TRUE + k * VU_N_BINS + bsel] += tru;
                    slot.f[VU_F64_BIN_SUMS + k * VU_N_BINS + bsel] += sw;
                }
                todo &= ~__ballot_sync(kFull, mine);
            }
        }
    }
}

// Fold a thread's tile partials into its warp slot.  All 32 lanes call it.
template <int WARPS>
__device__ __forceinline__ void stats_tile_end(const StatParams& sp, CtaStats<WARPS>& cs, TileAcc& acc) {
    const unsigned flags = sp.flags;
    const int lane = threadIdx.x & 31;
    WarpSlot& slot = cs.slot[threadIdx.x >> 5];
    {
        const int n = warp_sum(acc.nvox);
        if (lane == 0) slot.i[VU_I64_NVOX] += n;
    }
    if (flags & VU_STAT_IMAGE_SUM) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) {
            const double s = warp_sum(acc.sum[k]);
            if (lane == 0) slot.f[VU_F64_SUM + k] += s;
        }
    }
    if (flags & VU_STAT_THRESHOLD) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) {
            const double s = warp_sum(acc.thr_sum[k]);
            const int n = warp_sum(acc.thr_cnt[k]);
            if (lane == 0) { slot.f[VU_F64_THR_SUM + k] += s; slot.i[VU_I64_THR_COUNT + k] += n; }
        }
    }
    if (flags & VU_STAT_AREA) {
        const int n = warp_sum(acc.area);
        if (lane == 0) slot.i[VU_I64_AREA] += n;
    }
    if (flags & VU_STAT_DICE) {
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r) {
            if (r < sp.gt.R) {
                const int a = warp_sum(acc.tp[r]), p = warp_sum(acc.ps[r]), g = warp_sum(acc.gs[r]);
                if (lane == 0) { slot.i[VU_I64_DICE_TP + r] += a; slot.i[VU_I64_DICE_PRED + r] += p; slot.i[VU_I64_DICE_GT + r] += g; }
            }
        }
    }
    if (flags & VU_STAT_NCC) {
        const double g = warp_sum(acc.g), gg = warp_sum(acc.gg);
        if (lane == 0) { slot.f[VU_F64_NCC_G] += g; slot.f[VU_F64_NCC_GG] += gg; }
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) {
            const double a = warp_sum(acc.u[k]), c = warp_sum(acc.uu[k]), d = warp_sum(acc.gu[k]);
            if (lane == 0) { slot.f[VU_F64_NCC_U + k] += a; slot.f[VU_F64_NCC_UU + k] += c; slot.f[VU_F64_NCC_GU + k] += d; }
        }
    }
    acc.clear();
}


// The whole statistics phase for the VEC voxels a thread owns in one tile.
// Kept out of line so its register needs (double accumulators) do not inflate
// the streaming loop of the calling kernel.
template <int VEC, int WARPS>
__device__ __noinline__ void stats_tile(const StatParams& sp, CtaStats<WARPS>& cs, bool active, long long b, long long v,
                                        const float (&u)[VU_N_UNC][VEC], const int (&label)[VEC]) {
    TileAcc acc;
    acc.clear();
#pragma unroll 1
    for (int k = 0; k < VEC; ++k) {
        const float uk[VU_N_UNC] = {u[0][k], u[1][k], u[2][k]};
        stats_voxel<WARPS>(sp, cs, acc, active, b, v + k, uk, label[k]);
    }
    stats_tile_end<WARPS>(sp, cs, acc);
}

}  // namespace vu
