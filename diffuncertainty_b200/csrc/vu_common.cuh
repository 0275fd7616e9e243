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

// base pointer of member p from the device array of the "stack without copying" slab form (vu_slab.member_ptrs)
__device__ __forceinline__ const float* ld_member_ptr(const float* const* a, long long p) {
    return reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(a) + p));
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

// ---- packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE fp32 lanes per instruction) -------------
// The streaming loop is bound by instruction issue as much as by HBM, so every elementwise step that can
// be done on two voxels (or two classes) at once is.  Each lane rounds exactly like the scalar op.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// ---- p * log2(p) with the reference's skip rule ---------------------------
// test_utils.py:838-840 / 849-851 drop every NaN product, i.e. the term is
// p*log(p) for p > 0 and 0 for p == 0, p < 0 and NaN (+inf is kept).  The term is
// formed as  max(p, 0) * L(p)  with  L(p) = log2(max(p, FLT_MIN)):
//   p <= 0, NaN  -> 0 * (-126) = 0          (fmaxf returns the non-NaN operand)
//   subnormal p  -> p * (-126): off by < 2e-37 from the true term
//   +inf         -> inf
// L is one MUFU.LG2, whose absolute error (~2^-23.5 on [0.5, 2], measured on B200)
// is too large next to p = 1 (confident pixels: the term is ~ -(1-p)); within
// |p-1| < 1/16, L = f * q(f), f = p - 1, q = degree-3 near-minimax fit of
// log2(1+f)/f on [-1/16, 1/16].  Relative error of the term: < 6e-7 inside the
// window (checked over every float in it), < 1e-6 outside.
// Result is in log2 units; callers multiply the class sum by ln 2 once.
constexpr float kQ0 = 1.4426945447921753f, kQ1 = -0.7213456630706787f, kQ2 = 0.48202842473983765f, kQ3 = -0.36208757758140564f;
constexpr float kNearOne = 1.0f / 16.0f;

__device__ __forceinline__ float lg2_clamped(float p) {
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(fmaxf(p, FLT_MIN)));
    return lg;
}
// scalar: h + max(p,0) * L(p), one fused multiply-add (all kernels use this same operation order)
__device__ __forceinline__ float plog2p_acc(float h, float p) {
    const float lg = lg2_clamped(p);
    const float f = __fadd_rn(p, -1.0f);
    float t = __fmaf_rn(f, kQ3, kQ2);
    t = __fmaf_rn(t, f, kQ1);
    t = __fmaf_rn(t, f, kQ0);
    const float near_one = __fmul_rn(t, f);
    const float l = (fabsf(f) < kNearOne) ? near_one : lg;
    return __fmaf_rn(fmaxf(p, 0.0f), l, h);
}
// two elements at once: returns (max(p,0), L(p)) for both, the polynomial evaluated with packed ops
__device__ __forceinline__ void plog2p_parts2(f32x2 P, f32x2& PC, f32x2& L) {
    float p0, p1;
    upk2(P, p0, p1);
    const float lg0 = lg2_clamped(p0), lg1 = lg2_clamped(p1);
    const f32x2 F = add2(P, pk2(-1.0f, -1.0f));
    f32x2 T = fma2(F, pk2(kQ3, kQ3), pk2(kQ2, kQ2));
    T = fma2(T, F, pk2(kQ1, kQ1));
    T = fma2(T, F, pk2(kQ0, kQ0));
    const f32x2 N = mul2(T, F);
    float f0, f1, n0, n1;
    upk2(F, f0, f1);
    upk2(N, n0, n1);
    L = pk2((fabsf(f0) < kNearOne) ? n0 : lg0, (fabsf(f1) < kNearOne) ? n1 : lg1);
    PC = pk2(fmaxf(p0, 0.0f), fmaxf(p1, 0.0f));
}

// ---- softmax of a draw that arrives as logits (VU_SLAB_LOGITS; F.softmax(logits, dim=1), test_2D.py:1181-1256) ----------
// p_c = e_c / S,  e_c = exp(x_c - max_c x),  S = sum_c e_c (class order, float32), as torch's softmax kernels compute it.  The
// device evaluates e_c = ex2.approx((x_c - m) * log2 e) (2 ulp; e = 1 exactly at the maximum) and p_c = e_c * RN(1 / S).  A
// member's entropy term sum_c p_c log2 p_c is formed WITHOUT a logarithm per element:
//     sum_c p_c (z_c - log2 S) = (sum_c e_c z_c) / S - log2 S,   z_c = (x_c - m) * log2 e,
// with log2 S from the near-one polynomial above when S - 1 < 1/16 (confident voxels: S = 1 + tiny, and the term is ~ -(S - 1);
// the reference's own value there is quantised by the float32 rounding of S in exactly the same way).
// Special values follow torch: the maximum propagates NaN (max.NaN), so a NaN or +inf logit (or a voxel whose logits are all
// -inf) makes S NaN and with it every probability of the draw; a -inf logit next to finite ones has probability 0.
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float max_nan3(float a, float b, float c) {
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float ex2_approx(float z) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z));
    return e;
}
// log2(S) for S >= 1 (a sum of exponentials whose largest term is 1), NaN for NaN
__device__ __forceinline__ float log2_sum(float S) {
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(S));
    const float f = __fadd_rn(S, -1.0f);
    float t = __fmaf_rn(f, kQ3, kQ2);
    t = __fmaf_rn(t, f, kQ1);
    t = __fmaf_rn(t, f, kQ0);
    return (f < kNearOne) ? __fmul_rn(t, f) : lg;  // (NaN: the comparison is false, lg2(NaN) = NaN)
}
// 1 / S: MUFU.RCP and one Newton step (the correctly rounded reciprocal in all but a few cases per million; the same
// operations in the scalar and the packed form, so every kernel form produces the same bits)
__device__ __forceinline__ float rcp_approx_f(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_newton(float S) {
    const float r = rcp_approx_f(S);
    return __fmaf_rn(__fmaf_rn(S, -r, 1.0f), r, r);
}
// Scalar reference form of one draw's softmax at one voxel, used where a thread walks the classes one by one (generic kernel)
// and as the slow path of the vectorised form: pass 1 = softmax_max over the classes, pass 2 = softmax_term per class
// accumulating S and EZ, then softmax_finish.  -inf logits (and differences that overflow) contribute e = 0 and no e z term.
__device__ __forceinline__ float softmax_z(float x, float m) { return __fmul_rn(__fadd_rn(x, -m), kLog2e); }
__device__ __forceinline__ void softmax_term(float x, float m, float& S, float& EZ, float& e) {
    const float z = softmax_z(x, m);
    e = ex2_approx(z);
    S = __fadd_rn(S, e);
    EZ = (e == 0.0f) ? EZ : __fmaf_rn(e, z, EZ);  // 0 * -inf is not a term; NaN z: e is NaN and stays in
}
// rS = 1 / S; h = the draw's sum_c p_c log2 p_c (never positive), 0 when the draw is all NaN (every term skipped,
// test_utils.py:849-851): fminf returns the operand that is not NaN
__device__ __forceinline__ void softmax_finish(float S, float EZ, float& rS, float& h) {
    rS = rcp_newton(S);
    h = fminf(__fmaf_rn(EZ, rS, -log2_sum(S)), 0.0f);
}
// the same for two voxels at once (packed Newton step, polynomial and final fma)
__device__ __forceinline__ void softmax_finish2(f32x2 S, f32x2 EZ, f32x2& RS, f32x2& H) {
    float s0, s1;
    upk2(S, s0, s1);
    const f32x2 R0 = pk2(rcp_approx_f(s0), rcp_approx_f(s1));
    const f32x2 ERR = fma2(S, mul2(R0, pk2(-1.0f, -1.0f)), pk2(1.0f, 1.0f));
    RS = fma2(ERR, R0, R0);
    float lg0, lg1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg0) : "f"(s0));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg1) : "f"(s1));
    const f32x2 F = add2(S, pk2(-1.0f, -1.0f));
    f32x2 T = fma2(F, pk2(kQ3, kQ3), pk2(kQ2, kQ2));
    T = fma2(T, F, pk2(kQ1, kQ1));
    T = fma2(T, F, pk2(kQ0, kQ0));
    const f32x2 N = mul2(T, F);
    float f0, f1, n0, n1, t0, t1;
    upk2(F, f0, f1);
    upk2(N, n0, n1);
    const f32x2 NL = pk2((f0 < kNearOne) ? -n0 : -lg0, (f1 < kNearOne) ? -n1 : -lg1);
    upk2(fma2(EZ, RS, NL), t0, t1);
    H = pk2(fminf(t0, 0.0f), fminf(t1, 0.0f));
}

// torch.argmax tie rule (test_2D.py:817,871): first maximal index, NaN is max.
__device__ __forceinline__ void argmax_step(float v, int c, float& best, int& idx) {
    const bool take = (v > best) | ((v != v) & (best == best));  // (no short-circuit: three compares, no branches)
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
    int ign_byte;      // the ignore value can occur in a uint8 reference (has_ignore && 0 <= ignore <= 255)
    unsigned ign4;     // ... and its byte replicated four times
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

// uint8 references of `VEC` consecutive voxels of rater r as they lie in memory: byte j = voxel v + j (upper bytes 0).
// The strided fallback is kept out of line: inlined, its 64-bit address arithmetic is hoisted above the branch and executed
// by every tile of the aligned case too (ncu r01g: 7 % of the instructions of K3).
template <int VEC>
__device__ __noinline__ unsigned load_gt_bytes_strided(const uint8_t* base, long long v, long long sv) {
    unsigned w = 0;
#pragma unroll
    for (int k = 0; k < VEC; ++k) w |= (unsigned)__ldg(base + (v + k) * sv) << (8 * k);
    return w;
}
template <int VEC>
__device__ __forceinline__ unsigned load_gt_bytes(const GtView& gt, long long b, int r, long long v) {
    const uint8_t* base = reinterpret_cast<const uint8_t*>(gt.data) + b * gt.sb + (long long)r * gt.sr;
    if (VEC == 4 && gt.align >= 4) return __ldg(reinterpret_cast<const unsigned*>(base + v));
    if (VEC == 2 && gt.align >= 2) return (unsigned)__ldg(reinterpret_cast<const unsigned short*>(base + v));
    if (VEC == 1) return (unsigned)__ldg(base + v * gt.sv);
    return load_gt_bytes_strided<VEC>(base, v, gt.sv);
}

// ---- four voxels per word: byte-wise tests (results: 0x80 in the bytes where the test holds) --------------------------
constexpr unsigned kB01 = 0x01010101u, kB80 = 0x80808080u, kB7f = 0x7f7f7f7fu;
__device__ __forceinline__ unsigned bytes_nonzero(unsigned x) { return (((x & kB7f) + kB7f) | x) & kB80; }
__device__ __forceinline__ unsigned bytes_equal(unsigned a, unsigned b) { return ~bytes_nonzero(a ^ b) & kB80; }

struct CalibDev {
    float a, b;
    float edge[VU_N_EDGES];  // thresholds on u (sign-flipped when conf falls with u); NaN = never reached
    int increasing;
    int identity;  // the map already is the confidence
    float a2, b2;  // -a log2(e), b log2(e): conf = 1 / (1 + 2^(u a2 + b2))
    float sgn;     // +1 when conf rises with u, else -1 (the thresholds are sign-flipped to match)
};
constexpr float kRoundMagic = 12582912.0f;  // 1.5 * 2^23: x + magic rounds x to an integer held in the low mantissa bits

struct StatParams {
    unsigned flags;
    unsigned unc_mask;  // bit k set: uncertainty type k is present (TU/AU/EU; bit 0 only when P == 1)
    unsigned magic_bits;  // bit pattern of kRoundMagic, handed over as a run-time value: table bases biased by it stay one
                          // register (as a compile-time constant the compiler splits the bias off and re-adds it at every use)
    long long V;
    GtView gt;
    float thr[VU_N_UNC];
    CalibDev calib[VU_N_UNC];
    const uint8_t* lut;
    const double* ncc_gt_map;
    double* f64;
    long long* i64;
    // VU_STAT_CLASS_COUNTS: (B, R, ncls, 3) tp / pred / gt per rater and class
    long long* cls;
    int ncls;
    // VU_STAT_PLATT_FIT (dataset-level outputs)
    long long* platt_i64;
    double* platt_f64;
    float platt_edge[VU_N_PLATT_BINS + 1];  // smallest float32 >= edge k of np.logspace(-12, 2, 257)
};

// ---- per-CTA statistics state (dynamic shared memory) ---------------------------
// Every thread owns one private column of float64 / packed-integer accumulators
// ("slots", layout [slot][thread]: conflict-free, no atomics, no shuffles while
// streaming).  The calibration histograms are private to a warp and replicated
// 16 times ([unc][bin][lane & 15], one 64-bit word each): the two half-warps
// update them one after the other with plain load / add / store -- inside a
// half-warp every lane has its own replica, so there are no conflicts whatever
// the bins are, and no atomics (a shared-memory atomic costs ~64 cycles per warp
// on this part).  flush() folds everything into the image's row of the global
// statistics buffers with one atomic per non-zero column; it runs when a CTA
// moves on to another image, and at least every kMaxVoxPerFlush voxels per thread so that
// the packed counters cannot overflow.
//
// float64 slots                       packed-integer slots (uint64)
//   0..2   sum of u_k                   0      #(u_k >= t_k), 3 x 21 bit
//   3..5   sum of u_k over u_k >= t_k   1      #(label > 0)
//   6..8   sum of conf in bin 0         2, 3   samples / correct samples with NaN u_k (bin slot 20), 3 x 21 bit
//   9      sum g        10  sum g*g     4+r    rater r: tp | pred << 21 | gt << 42
//   11..13 sum u_k^2    14..16  sum g*u_k
// histogram word (bins 0..19): x = samples | correct << 16,  y = sum of q * n_valid,
//   q = round(conf * 2^21) - bin * kQBinStep   (bin 0 is also summed in floating point, slots 6..8)
enum { FS_SUM = 0, FS_THR = 3, FS_BIN0 = 6, FS_G = 9, FS_GG = 10, FS_UU = 11, FS_GU = 14, FS_MAX = 17 };
enum { IS_THRCNT = 0, IS_AREA = 1, IS_NANTOT = 2, IS_NANTRU = 3, IS_DICE = 4, IS_MAX = 4 + VU_MAX_RATERS };
constexpr int kPackBits = 21;
constexpr unsigned long long kPackMask = (1ull << kPackBits) - 1;
constexpr int kMaxVoxPerFlush = 1024;   // voxels per thread between flushes: x 2 lanes per replica x n_valid <= 8 keeps the
                                        // 16-bit counts (<= 16384) and the 32-bit q sums (< 2^31) from overflowing
constexpr int kQBits = 21;
constexpr int kQBinStep = 104858;        // round(0.05 * 2^21): what one bin is worth in q units (Sum conf = count * bin * step / 2^21 + Sum q / 2^21)
constexpr int kHistBins = VU_N_BINS - 1;  // 20 real bins; slot 20 (NaN) is counted in the integer slots
// histogram replicas per warp: 16 (two half-warp phases; the register-streaming kernels, 8 warps per CTA) or 32 (one
// replica per lane, no phases; the three statistics warps of the TMA kernel)
constexpr unsigned kRuntimeFlags = 0xffffffffu;  // template value: statistics mask read from StatParams at run time
constexpr int kEdgePad = 24;  // per uncertainty type: float2 EP[k] = (lower, upper) edge of bin k on u; NaN = no such edge

// Platt-fit data: one CTA-wide histogram [unc][bin][samples, correct, q_hi, q_lo] of int32 updated with native shared
// atomics (768 bins are too many to replicate per lane), plus the edge table T[0..256] (padded) and 1 / edge
constexpr int kPlattWords = VU_N_UNC * VU_N_PLATT_BINS * 4;
constexpr int kPlattTab = 264;
constexpr int kPlattQBits = 20;  // q = round((u / edge_bin - 1) * 2^20), |q| <= 2^21 after clamping
constexpr int kPlattQSplit = 12;

__host__ __device__ inline int stats_num_fslots(unsigned flags) {
    return (flags & VU_STAT_NCC) ? FS_MAX : ((flags & VU_STAT_CALIB) ? FS_G : FS_BIN0);
}
__host__ __device__ inline int stats_num_islots(unsigned flags, int R) {
    return (flags & VU_STAT_DICE) ? IS_DICE + R : ((flags & VU_STAT_CALIB) ? IS_DICE : IS_NANTOT);
}
__host__ __device__ inline size_t stats_smem_bytes(unsigned flags, int R, int threads, int rep = 16) {
    if (!flags) return 0;
    size_t n = (size_t)(stats_num_fslots(flags) + stats_num_islots(flags, R)) * threads * 8;
    if (flags & VU_STAT_CALIB)
        n += (size_t)(threads / 32) * (VU_N_UNC * kHistBins * rep) * sizeof(uint2) + (size_t)VU_N_UNC * kEdgePad * sizeof(float2) +
             (size_t)(VU_N_UNC + 1) * sizeof(float4);
    if (flags & VU_STAT_PLATT_FIT) n += (size_t)kPlattWords * sizeof(int) + (size_t)kPlattTab * 2 * sizeof(float);
    return n;
}
// the per-class counters of VU_STAT_CLASS_COUNTS ([rater][class][tp, pred, gt] of uint32, one set per CTA) follow the state above
__host__ __device__ inline size_t stats_class_bytes(unsigned flags, int R, int ncls) {
    return (flags & VU_STAT_CLASS_COUNTS) ? (size_t)R * ncls * 3 * sizeof(unsigned) : 0;
}

template <int THREADS, int REP = 16>
struct StatsLayout {
    static constexpr int kHistWordsPerWarp = VU_N_UNC * kHistBins * REP;  // uint2 words
    double* fs;
    unsigned long long* is;
    uint2* hist;  // [warp][unc][bin][replica]
    float2* E;    // [unc][kEdgePad]
    float4* CC;   // [unc]: (a2, b2, sgn, -) of the Platt expression, so that a lane can fetch its type's constants with one load
    int* phist;   // Platt-fit data [unc][bin][4]
    float* pT;    // [kPlattTab] edges, then [kPlattTab] reciprocals
    unsigned* cc; // class counters [rater][class][3]
    int nF, nI;
    __device__ __forceinline__ StatsLayout(const StatParams& sp, void* smem) {
        nF = stats_num_fslots(sp.flags);
        nI = stats_num_islots(sp.flags, sp.gt.R);
        fs = reinterpret_cast<double*>(smem);
        is = reinterpret_cast<unsigned long long*>(fs + (size_t)nF * THREADS);
        hist = reinterpret_cast<uint2*>(is + (size_t)nI * THREADS);
        E = reinterpret_cast<float2*>(hist + ((sp.flags & VU_STAT_CALIB) ? (THREADS / 32) * kHistWordsPerWarp : 0));
        CC = reinterpret_cast<float4*>(E + ((sp.flags & VU_STAT_CALIB) ? VU_N_UNC * kEdgePad : 0));
        phist = reinterpret_cast<int*>(CC + ((sp.flags & VU_STAT_CALIB) ? VU_N_UNC + 1 : 0));
        pT = reinterpret_cast<float*>(phist + kPlattWords);
        cc = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(smem) + stats_smem_bytes(sp.flags, sp.gt.R, THREADS, REP));
    }
};

// CTA-wide barrier over the THREADS statistics threads: the whole CTA (BAR == 0) or named barrier BAR for
// kernels whose CTA holds other warps as well (the TMA producer warp)
template <int THREADS, int BAR>
__device__ __forceinline__ void stats_sync() {
    if (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(THREADS) : "memory");
}

template <int THREADS, int BAR = 0, int TID0 = 0, int REP = 16>
__device__ __noinline__ void stats_init(const StatParams& sp, void* smem) {
    StatsLayout<THREADS, REP> L(sp, smem);
    constexpr int kHistWordsPerWarp = StatsLayout<THREADS, REP>::kHistWordsPerWarp;
    for (int t = ((int)threadIdx.x - TID0); t < (L.nF + L.nI) * THREADS; t += THREADS) reinterpret_cast<unsigned long long*>(smem)[t] = 0ull;
    if (sp.flags & VU_STAT_CALIB) {
        for (int t = ((int)threadIdx.x - TID0); t < (THREADS / 32) * kHistWordsPerWarp; t += THREADS) L.hist[t] = make_uint2(0u, 0u);
        for (int t = ((int)threadIdx.x - TID0); t < VU_N_UNC * kEdgePad; t += THREADS) {
            const int k = t / kEdgePad, e = t % kEdgePad;  // bin e lies between interior edges e - 1 and e
            const float nan = __int_as_float(0x7fc00000);
            L.E[t] = make_float2((e >= 1 && e <= VU_N_EDGES) ? sp.calib[k].edge[e - 1] : nan,
                                 (e >= 0 && e < VU_N_EDGES) ? sp.calib[k].edge[e] : nan);
        }
        for (int t = ((int)threadIdx.x - TID0); t < VU_N_UNC; t += THREADS)
            L.CC[t] = make_float4(sp.calib[t].a2, sp.calib[t].b2, sp.calib[t].sgn, 0.f);
    }
    if (sp.flags & VU_STAT_CLASS_COUNTS)
        for (int t = ((int)threadIdx.x - TID0); t < sp.gt.R * sp.ncls * 3; t += THREADS) L.cc[t] = 0u;
    if (sp.flags & VU_STAT_PLATT_FIT) {
        for (int t = ((int)threadIdx.x - TID0); t < kPlattWords; t += THREADS) L.phist[t] = 0;
        for (int t = ((int)threadIdx.x - TID0); t < kPlattTab; t += THREADS) {
            const float e = t <= VU_N_PLATT_BINS ? sp.platt_edge[t] : __int_as_float(0x7fc00000);
            L.pT[t] = e;
            L.pT[kPlattTab + t] = t < VU_N_PLATT_BINS ? 1.0f / e : 0.f;
        }
    }
    stats_sync<THREADS, BAR>();
}

// Add this CTA's partials into image row b and clear them.  nvox = voxels of image b
// the CTA went through since the last flush.  Called by every thread of the CTA.
template <int THREADS, int BAR = 0, int TID0 = 0, int REP = 16>
__device__ __noinline__ void stats_flush(const StatParams& sp, void* smem, long long b, long long nvox) {
    stats_sync<THREADS, BAR>();
    StatsLayout<THREADS, REP> L(sp, smem);
    constexpr int kHistRep = REP;
    constexpr int kHistWordsPerWarp = StatsLayout<THREADS, REP>::kHistWordsPerWarp;
    constexpr int WARPS = THREADS / 32;
    const int warp = ((int)threadIdx.x - TID0) >> 5, lane = ((int)threadIdx.x - TID0) & 31;
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
            } else if (q == IS_NANTOT || q == IS_NANTRU) {
                // np.digitize sends NaN confidences past the last edge: slot 20, whose sum of confidences is NaN
                const int col = (q == IS_NANTOT ? VU_I64_BIN_TOTAL : VU_I64_BIN_TRUE) + (VU_N_BINS - 1);
                const unsigned long long f[3] = {f0, f1, f2};
                for (int k = 0; k < VU_N_UNC; ++k)
                    if (f[k]) {
                        atomicAdd(irow + col + k * VU_N_BINS, f[k]);
                        if (q == IS_NANTOT) atomicAdd(frow + VU_F64_BIN_SUMS + k * VU_N_BINS + (VU_N_BINS - 1), (double)__int_as_float(0x7fc00000));
                    }
            } else {
                const int r = q - IS_DICE;
                if (f0) atomicAdd(irow + VU_I64_DICE_TP + r, f0);
                if (f1) atomicAdd(irow + VU_I64_DICE_PRED + r, f1);
                if (f2) atomicAdd(irow + VU_I64_DICE_GT + r, f2);
            }
        }
    }
    if (flags & VU_STAT_CALIB) {
        // 4 adjacent threads fold one (unc, bin): each takes a quarter of its WARPS x 16 replica words
        constexpr int kItems = VU_N_UNC * kHistBins * 4;
        constexpr int kPer = WARPS * kHistRep / 4;
        for (int base = 0; base < ((kItems + THREADS - 1) / THREADS) * THREADS; base += THREADS) {
            const int item = base + ((int)threadIdx.x - TID0);
            const int pair = item >> 2, part = item & 3;
            long long tot = 0, tru = 0, q = 0;
            if (pair < VU_N_UNC * kHistBins) {
                for (int i = 0; i < kPer; ++i) {
                    const int e = part * kPer + i;
                    const int w = e / kHistRep, rep = (e + pair) & (kHistRep - 1);  // skewed: neighbours read different banks
                    uint2* h = L.hist + w * kHistWordsPerWarp + pair * kHistRep + rep;
                    const uint2 v = *h;
                    *h = make_uint2(0u, 0u);
                    tot += v.x & 0xffffu;
                    tru += v.x >> 16;
                    q += (int)v.y;
                }
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                tot += __shfl_xor_sync(kFull, tot, o);
                tru += __shfl_xor_sync(kFull, tru, o);
                q += __shfl_xor_sync(kFull, q, o);
            }
            if (part == 0 && pair < VU_N_UNC * kHistBins && tot) {
                const int k = pair / kHistBins, bin = pair % kHistBins;
                const int col = k * VU_N_BINS + bin;
                atomicAdd(irow + VU_I64_BIN_TOTAL + col, (unsigned long long)tot);
                if (tru) atomicAdd(irow + VU_I64_BIN_TRUE + col, (unsigned long long)tru);
                if (bin > 0)  // bin 0 is summed in floating point by the threads (slots FS_BIN0)
                    atomicAdd(frow + VU_F64_BIN_SUMS + col,
                              ((double)tot * (double)(bin * kQBinStep) + (double)q) * (1.0 / (double)(1 << kQBits)));  // q is relative to bin * kQBinStep
            }
        }
    }
    if (flags & VU_STAT_CLASS_COUNTS) {
        unsigned long long* crow = reinterpret_cast<unsigned long long*>(sp.cls) + b * ((long long)sp.gt.R * sp.ncls * 3);
        for (int t = ((int)threadIdx.x - TID0); t < sp.gt.R * sp.ncls * 3; t += THREADS) {
            const unsigned c = L.cc[t];
            if (c) { atomicAdd(crow + t, (unsigned long long)c); L.cc[t] = 0u; }
        }
    }
    if (flags & VU_STAT_PLATT_FIT) {  // dataset-level: not tied to the image row
        for (int t = ((int)threadIdx.x - TID0); t < VU_N_UNC * VU_N_PLATT_BINS; t += THREADS) {
            int* h = L.phist + t * 4;
            const int tot = h[0], pos = h[1];
            if (tot) {
                const long long q = (long long)h[2] * (1 << kPlattQSplit) + h[3];
                const int bin = t % VU_N_PLATT_BINS;
                atomicAdd(reinterpret_cast<unsigned long long*>(sp.platt_i64) + 2 * t, (unsigned long long)tot);
                if (pos) atomicAdd(reinterpret_cast<unsigned long long*>(sp.platt_i64) + 2 * t + 1, (unsigned long long)pos);
                // sum of u over the bin = edge * (count + sum((u / edge - 1)))
                atomicAdd(sp.platt_f64 + t, (double)L.pT[bin] * ((double)tot + (double)q * (1.0 / (double)(1 << kPlattQBits))));
                h[0] = 0; h[1] = 0; h[2] = 0; h[3] = 0;
            }
        }
    }
    if (((int)threadIdx.x - TID0) == 0 && nvox) atomicAdd(irow + VU_I64_NVOX, (unsigned long long)nvox);
    stats_sync<THREADS, BAR>();
}

// Tracks which image a CTA is working on and flushes when it moves on.  All members are CTA-uniform.
template <int THREADS>
struct StatsCursor {
    int cur_b, vt_begin;
    __device__ __forceinline__ StatsCursor() : cur_b(-1), vt_begin(0) {}
    // top of every tile: tile index vt inside image b, tile_vox voxels per tile
    template <int BAR = 0, int TID0 = 0, int REP = 16>
    __device__ __forceinline__ void enter(const StatParams& sp, void* smem, int b, int vt, long long tile_vox) {
        if (b != cur_b || (long long)(vt - vt_begin) * (tile_vox / THREADS) >= kMaxVoxPerFlush) {
            if (cur_b >= 0) {
                const long long end = (b != cur_b) ? sp.V : (long long)vt * tile_vox;
                stats_flush<THREADS, BAR, TID0, REP>(sp, smem, cur_b, end - (long long)vt_begin * tile_vox);
            }
            cur_b = b;
            vt_begin = vt;
        }
    }
    // after the last tile (vt = its index inside the image)
    template <int BAR = 0, int TID0 = 0, int REP = 16>
    __device__ __forceinline__ void finish(const StatParams& sp, void* smem, int last_vt, long long tile_vox) {
        if (cur_b >= 0) {
            long long end = (long long)(last_vt + 1) * tile_vox;
            end = end > sp.V ? sp.V : end;
            stats_flush<THREADS, BAR, TID0, REP>(sp, smem, cur_b, end - (long long)vt_begin * tile_vox);
        }
    }
};

// Pull the references of a tile towards the SM while the slab is still streaming (L1 is otherwise unused:
// the slab is read with no-allocate loads), so that the statistics phase does not wait on HBM.
template <int VEC>
__device__ __forceinline__ void stats_prefetch_gt(const StatParams& sp, long long b, long long v) {
    if (!(sp.flags & (VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC | VU_STAT_PLATT_FIT | VU_STAT_CLASS_COUNTS)) || !sp.gt.data) return;
    const long long esz = sp.gt.dtype == VU_GT_U8 ? 1 : 8;
    if (sp.gt.sv == 1) {
        // the 32 x VEC voxels of a warp are one run of 128-byte lines per rater: the lanes share them out (one instruction for
        // all raters instead of R per thread -- these launches are issue-bound)
        const int lane = threadIdx.x & 31;
        const int lpr = (int)((32 * VEC * esz + 127) / 128);  // lines per rater
        const long long vb = v - (long long)lane * VEC;  // first voxel of the warp
        const char* base = reinterpret_cast<const char*>(sp.gt.data) + (b * sp.gt.sb + vb) * esz;
        for (int i = lane; i < sp.gt.R * lpr; i += 32) {
            const long long line = i % lpr;
            if (vb + line * (128 / esz) < sp.V)  // never past the end of a rater's row (the last warp of an image is ragged)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (long long)(i / lpr) * sp.gt.sr * esz + line * 128));
        }
        return;
    }
    const char* base = reinterpret_cast<const char*>(sp.gt.data) + (b * sp.gt.sb + v * sp.gt.sv) * esz;
    for (int r = 0; r < sp.gt.R; ++r) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (long long)r * sp.gt.sr * esz));
}

// ace.py:329 in float32: 1 / (1 + exp((-u) * a + b)); ace.py:333 clips to [0, 1].  Bins are decided on u itself (below), so
// this value only feeds the floating bin_sums and the candidate bin: ex2.approx / rcp.approx are enough, and the result
// lies in [0, 1] by construction (NaN for NaN u).  a2 = -a log2(e), b2 = b log2(e).
__device__ __forceinline__ float platt_conf(float u, float a2, float b2, int identity) {
    if (identity) return fminf(fmaxf(u, 0.0f), 1.0f);
    float e, c;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(u, a2, b2)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(1.0f + e));
    return c;
}

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

// The whole statistics phase for the VEC voxels a thread owns in one tile.  Must
// be called by every thread of the CTA (threads past the end of the image pass
// active = false).  Kept out of line so its registers do not inflate the
// streaming loop of the calling kernel; arguments travel in registers (the three
// maps as built-in vectors, the VEC labels packed into one word).
#ifndef VU_STATS_INLINE
#define VU_STATS_INLINE __forceinline__
#endif
// FL: the statistics mask as a compile-time constant (the caller guarantees sp.flags == FL, all three uncertainty
// types present and no label LUT), or kRuntimeFlags.  REP: histogram replicas per warp (16 or 32).
template <int VEC, int THREADS, typename GT, int TID0 = 0, int REP = 16, unsigned FL = kRuntimeFlags>
__device__ VU_STATS_INLINE void stats_tile_t(const StatParams& sp, void* smem, bool active, long long b, long long v,
                                          typename FVec<VEC>::type u0, typename FVec<VEC>::type u1,
                                          typename FVec<VEC>::type u2, unsigned labels_packed) {
    using GAcc = typename std::conditional<sizeof(GT) == 1, int, double>::type;
    constexpr int RB = sizeof(GT) == 1 ? 4 : (VEC == 4 ? 1 : 2);  // raters fetched together
    StatsLayout<THREADS, REP> cs(sp, smem);
    constexpr int kHistRep = REP;
    constexpr int kHistWordsPerWarp = StatsLayout<THREADS, REP>::kHistWordsPerWarp;
    float u[VU_N_UNC][VEC];
    int label[VEC];
    fvec_unpack(u0, u[0]);
    fvec_unpack(u1, u[1]);
    fvec_unpack(u2, u[2]);
#pragma unroll
    for (int j = 0; j < VEC; ++j) label[j] = (int)((labels_packed >> (8 * j)) & 0xffu);
    const unsigned flags = (FL == kRuntimeFlags) ? sp.flags : FL, mask = (FL == kRuntimeFlags) ? sp.unc_mask : 7u;
    const uint8_t* lut = (FL == kRuntimeFlags) ? sp.lut : nullptr;
    const int tid = (int)threadIdx.x - TID0, lane = tid & 31;

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
    if (!(flags & (VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC | VU_STAT_PLATT_FIT | VU_STAT_CLASS_COUNTS))) return;

    // ---- references: per voxel the number of valid / correct raters, as byte j of one word each (<= 8) -------------
    constexpr unsigned kBytes = VEC == 4 ? 0xffffffffu : (VEC == 2 ? 0x0000ffffu : 0x000000ffu);
    unsigned nv4 = 0, nc4 = 0;
    GAcc gsum[VEC], gsq[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { gsum[j] = 0; gsq[j] = 0; }
    const int R = sp.gt.data ? sp.gt.R : 0;
    if (active && R > 0) {
        const unsigned lab4 = labels_packed & kBytes;
        unsigned cmp4 = lab4;  // what a reference has to equal to count as correct (ace.py:488)
        if (lut) {
            cmp4 = 0u;
#pragma unroll
            for (int j = 0; j < VEC; ++j) cmp4 |= (unsigned)__ldg(lut + label[j]) << (8 * j);
        }
        const unsigned pp_hi = bytes_equal(lab4, kB01) & kBytes;  // predicted foreground (test_2D.py:878)
        const bool has_ign = sp.gt.has_ignore != 0;
        const bool ign_byte = sp.gt.ign_byte != 0;  // else no uint8 reference can match the ignore value
        const unsigned ign4 = sp.gt.ign4;
        const long long ign = sp.gt.ignore;
#pragma unroll 1
        for (int r0 = 0; r0 < R; r0 += RB) {
            unsigned w[RB];
            long long g64[sizeof(GT) == 1 ? 1 : RB][VEC];
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                w[i] = 0u;
                if (r0 + i < R) {
                    if (sizeof(GT) == 1) w[i] = load_gt_bytes<VEC>(sp.gt, b, r0 + i, v);
                    else load_gt<VEC>(sp.gt, b, r0 + i, v, g64[sizeof(GT) == 1 ? 0 : i], (long long)0);
                }
            }
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                if (r0 + i >= R) break;
                unsigned valid_hi, eq_hi, gp_hi;  // 0x80 in byte j: reference of voxel j is valid / equals the label / is 1
                if (sizeof(GT) == 1) {
                    valid_hi = (ign_byte ? bytes_nonzero(w[i] ^ ign4) : kB80) & kBytes;  // ace.py:492-499, test_2D.py:880
                    eq_hi = bytes_equal(w[i], cmp4) & valid_hi;
                    gp_hi = bytes_equal(w[i], kB01) & valid_hi;                          // test_2D.py:882
                    if (flags & VU_STAT_NCC) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const int gj = (int)((w[i] >> (8 * j)) & 0xffu);
                            gsum[j] += (GAcc)gj;
                            gsq[j] += (GAcc)(gj * gj);
                        }
                    }
                } else {
                    valid_hi = 0u; eq_hi = 0u; gp_hi = 0u;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        const long long gj = g64[sizeof(GT) == 1 ? 0 : i][j];
                        const bool valid = !(has_ign && gj == ign);
                        valid_hi |= valid ? (0x80u << (8 * j)) : 0u;
                        eq_hi |= (valid && gj == (long long)((cmp4 >> (8 * j)) & 0xffu)) ? (0x80u << (8 * j)) : 0u;
                        gp_hi |= (valid && gj == 1) ? (0x80u << (8 * j)) : 0u;
                        if (flags & VU_STAT_NCC) {
                            gsum[j] += (GAcc)gj;
                            gsq[j] += (GAcc)gj * (GAcc)gj;
                        }
                    }
                }
                nv4 += valid_hi >> 7;
                nc4 += eq_hi >> 7;
                if (flags & VU_STAT_DICE) {
                    const unsigned ppv = pp_hi & valid_hi;
                    const unsigned tp = __popc(ppv & gp_hi), ps = __popc(ppv), gs = __popc(gp_hi);
                    if (tp | ps | gs)
                        cs.is[(IS_DICE + r0 + i) * THREADS + tid] +=
                            (unsigned long long)tp | ((unsigned long long)ps << kPackBits) | ((unsigned long long)gs << (2 * kPackBits));
                }
            }
        }
    }
    if (flags & VU_STAT_CLASS_COUNTS) {
        // Per rater one counter update per (label, reference) pair present in the warp: neighbouring voxels mostly carry the
        // same pair, so the lanes are grouped by it (match.any) and the first lane of a group adds the group's size.  Every
        // lane of the warp walks through here; the references come from L1 (the loop above has just read them).
        const int Rr = sp.gt.data ? sp.gt.R : 0, ncls = sp.ncls;
        for (int r = 0; r < Rr; ++r) {
            long long g[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) g[j] = -1;
            if (active) {
                if (sizeof(GT) == 1) {
                    const unsigned w = load_gt_bytes<VEC>(sp.gt, b, r, v);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) g[j] = (long long)((w >> (8 * j)) & 0xffu);
                } else {
                    load_gt<VEC>(sp.gt, b, r, v, g, (long long)0);
                }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const bool ok = active && !(sp.gt.has_ignore && g[j] == sp.gt.ignore) && g[j] >= 0 && g[j] < ncls && label[j] < ncls;
                const unsigned key = ok ? ((unsigned)label[j] << 8) | (unsigned)g[j] : 0xffffffffu;
                const unsigned grp = __match_any_sync(kFull, key);
                if (ok && lane == __ffs(grp) - 1) {
                    const unsigned n = __popc(grp);
                    unsigned* c = cs.cc + (r * ncls) * 3;
                    atomicAdd(c + label[j] * 3 + 1, n);
                    atomicAdd(c + (int)g[j] * 3 + 2, n);
                    if ((long long)label[j] == g[j]) atomicAdd(c + label[j] * 3, n);
                }
            }
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
    if (flags & VU_STAT_PLATT_FIT) {
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) {
            if (!((mask >> k) & 1)) continue;
            int* hk = cs.phist + k * (VU_N_PLATT_BINS * 4);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const int nv = (int)((nv4 >> (8 * j)) & 0xffu), nc = (int)((nc4 >> (8 * j)) & 0xffu);
                if (active && nv > 0) {
                    const float x = u[k][j];
                    // candidate bin from log10(u), settled exactly by the two neighbouring edges (ace.py:117-126)
                    float lg;
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(x));
                    int b0 = __float2int_rd(fmaf(lg, 0.30102999566f * (256.0f / 14.0f), 12.0f * 256.0f / 14.0f));
                    b0 = b0 < 0 ? 0 : (b0 > VU_N_PLATT_BINS - 1 ? VU_N_PLATT_BINS - 1 : b0);
                    int bin = b0 + (x >= cs.pT[b0 + 1] ? 1 : 0) - (x < cs.pT[b0] ? 1 : 0);
                    bin = bin < 0 ? 0 : (bin > VU_N_PLATT_BINS - 1 ? VU_N_PLATT_BINS - 1 : bin);
                    bin = (x != x) ? VU_N_PLATT_BINS - 1 : bin;  // np.digitize sends NaN past the last edge
                    float rel = fmaf(x, cs.pT[kPlattTab + bin], -1.0f);
                    rel = fminf(fmaxf(rel, -1.0f), 2.0f);  // out-of-range values are clamped into the end bins (NaN -> -1)
                    const int q = __float2int_rn(rel * (float)(1 << kPlattQBits)) * nv;
                    int* h = hk + bin * 4;
                    atomicAdd(h, nv);
                    if (nc) atomicAdd(h + 1, nc);
                    atomicAdd(h + 2, q >> kPlattQSplit);
                    atomicAdd(h + 3, q & ((1 << kPlattQSplit) - 1));
                }
            }
        }
    }
    if (flags & VU_STAT_CALIB) {  // every lane of the warp walks through here (the half-warp phases are warp-synchronous)
        uint2* hw = cs.hist + (tid >> 5) * kHistWordsPerWarp + (lane & (kHistRep - 1));
        const bool upper = lane >= kHistRep;  // REP == 32: never
        // per voxel: samples | correct << 16 (what a hit adds to the count word) and the samples as a float; lanes past the
        // end of the image and voxels without a valid rater carry zeros, so whatever bin they compute receives nothing
        unsigned vc[VEC];
        float nvf[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const unsigned nv = (nv4 >> (8 * j)) & 0xffu, nc = (nc4 >> (8 * j)) & 0xffu;
            vc[j] = nv | (nc << 16);
            nvf[j] = (float)nv;
        }
        unsigned nanbits = 0;  // bit (k * VEC + j): a sample with NaN uncertainty (rare: handled after the loops)
        const bool any_identity = sp.calib[0].identity | sp.calib[1].identity | sp.calib[2].identity;
        if (kHistRep == 16 && mask == 7u && !any_identity) {
            // Two half-warps share a histogram replica.  Instead of taking turns on every update, the lower half works on
            // uncertainty type ks while the upper half works on (ks + 1) % 3: in the same step the halves never address the
            // same histogram region, so ONE read-modify-write serves all 32 lanes and the warp only synchronises when the
            // halves move on to the next pair of regions (3 times per tile instead of 24).
#pragma unroll
            for (int ks = 0; ks < VU_N_UNC; ++ks) {
                const int ku = (ks + 1) % VU_N_UNC;
                const int kk = upper ? ku : ks;
                const float4 cc = cs.CC[kk];
                const float a2 = cc.x, b2 = cc.y, sgn = cc.z;
                const float2* E = cs.E + kk * kEdgePad;
                uint2* hk = hw + kk * (kHistBins * kHistRep);
                float bin0 = 0.f;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float x = upper ? u[ku][j] : u[ks][j];
                    const bool is_nan = x != x;
                    const float conf = is_nan ? 0.0f : platt_conf(x, a2, b2, 0);
                    const float kf = fminf(fmaf(conf, 20.0f, kRoundMagic), kRoundMagic + 19.0f);
                    const float2 e = E[__float_as_int(kf) & 0xff];
                    const float uu = x * sgn;
                    const float t = (kf + ((uu >= e.y) ? 1.0f : 0.0f)) - ((uu < e.x) ? 1.0f : 0.0f);
                    const int bin = __float_as_int(t) & 0xff;
                    // q = round(conf * 2^21) - bin * kQBinStep: the rounding happens in the mantissa of conf * 2^21 + magic
                    const int q = (__float_as_int(fmaf(conf, (float)(1 << kQBits), kRoundMagic)) - __float_as_int(kRoundMagic)) - bin * kQBinStep;
                    const unsigned inc = is_nan ? 0u : vc[j];
                    if (t == kRoundMagic) bin0 = fmaf(conf, nvf[j], bin0);
                    nanbits |= is_nan ? (1u << (kk * VEC + j)) : 0u;
                    uint2* h = hk + bin * kHistRep;
                    uint2 wd = *h;
                    wd.x += inc;
                    wd.y += (unsigned)(q * (int)(inc & 0xffffu));
                    *h = wd;
                }
                if (bin0 != 0.f) cs.fs[(FS_BIN0 + kk) * THREADS + tid] += (double)bin0;
                __syncwarp();
            }
        } else {
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k) {
                if (!((mask >> k) & 1)) continue;
                float bin0 = 0.f;
                const CalibDev& cal = sp.calib[k];
                const float a2 = cal.a2, b2 = cal.b2, sgn = cal.sgn;
                const float2* E = cs.E + k * kEdgePad;
                uint2* hk = hw + k * (kHistBins * kHistRep);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float x = u[k][j];
                    const bool is_nan = x != x;
                    const float conf = is_nan ? 0.0f : platt_conf(x, a2, b2, cal.identity);  // finite from here on
                    // candidate bin round(conf * 20), read off the mantissa (no conversion); the true bin is within one of it and
                    // the two neighbouring thresholds on u settle it exactly (see vu_calib in valunc.h; NaN edges and NaN u
                    // compare false).  t = magic + bin stays in the mantissa domain: bin = low bits, float(bin) = t - magic.
                    const float kf = fminf(fmaf(conf, 20.0f, kRoundMagic), kRoundMagic + 19.0f);
                    const float2 e = E[__float_as_int(kf) & 0xff];
                    const float uu = x * sgn;
                    const float t = (kf + ((uu >= e.y) ? 1.0f : 0.0f)) - ((uu < e.x) ? 1.0f : 0.0f);
                    const int bin = __float_as_int(t) & 0xff;
                    // q = round(conf * 2^21) - bin * kQBinStep: the rounding happens in the mantissa of conf * 2^21 + magic
                    const int q = (__float_as_int(fmaf(conf, (float)(1 << kQBits), kRoundMagic)) - __float_as_int(kRoundMagic)) - bin * kQBinStep;
                    const unsigned inc = is_nan ? 0u : vc[j];
                    const unsigned qq = (unsigned)(q * (int)(inc & 0xffffu));
                    if (t == kRoundMagic) bin0 = fmaf(conf, nvf[j], bin0);  // a NaN sample has conf 0 here
                    nanbits |= is_nan ? (1u << (k * VEC + j)) : 0u;
                    uint2* h = hk + bin * kHistRep;
                    if (kHistRep == 32) {
                        // one replica per lane: plain read-modify-write, lanes that do not hit add zero
                        uint2 wd = *h;
                        wd.x += inc; wd.y += qq;
                        *h = wd;
                    } else {
                        if (!upper) { uint2 wd = *h; wd.x += inc; wd.y += qq; *h = wd; }
                        __syncwarp();
                        if (upper) { uint2 wd = *h; wd.x += inc; wd.y += qq; *h = wd; }
                        __syncwarp();
                    }
                }
                if (bin0 != 0.f) cs.fs[(FS_BIN0 + k) * THREADS + tid] += (double)bin0;
            }
        }
        if (nanbits) {  // np.digitize puts NaN past the last edge: slot 20, counted in the integer slots
            unsigned long long nan_tot = 0, nan_tru = 0;
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k)
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if (((nanbits >> (k * VEC + j)) & 1u) && vc[j] != 0u) {
                        nan_tot += (unsigned long long)(vc[j] & 0xffffu) << (k * kPackBits);
                        nan_tru += (unsigned long long)(vc[j] >> 16) << (k * kPackBits);
                    }
            cs.is[IS_NANTOT * THREADS + tid] += nan_tot;
            cs.is[IS_NANTRU * THREADS + tid] += nan_tru;
        }
    }
}

template <int VEC, int THREADS, int TID0 = 0, int REP = 16, unsigned FL = kRuntimeFlags>
__device__ __forceinline__ void stats_tile(const StatParams& sp, void* smem, bool active, long long b, long long v,
                                           const float (&u)[VU_N_UNC][VEC], const int (&label)[VEC]) {
    unsigned lp = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) lp |= (unsigned)(label[j] & 0xff) << (8 * j);
    if (sp.gt.dtype == VU_GT_I64)
        stats_tile_t<VEC, THREADS, long long, TID0, REP, FL>(sp, smem, active, b, v, fvec_pack(u[0]), fvec_pack(u[1]), fvec_pack(u[2]), lp);
    else
        stats_tile_t<VEC, THREADS, uint8_t, TID0, REP, FL>(sp, smem, active, b, v, fvec_pack(u[0]), fvec_pack(u[1]), fvec_pack(u[2]), lp);
}

}  // namespace vu
