// Per-thread arithmetic of the fused pass, shared by the register-streaming kernel (k1_fused.cu) and the
// TMA-pipelined kernel (k1_tma.cu): member mean in torch's cascade order, per-member entropy, and the
// epilogue (true division, first-max argmax, TU / AU / EU).  Both kernels and the generic kernel produce
// bit-identical maps because every per-voxel operation and its order are defined here.
//
// Reference: uncertainty_modeling/test_2D.py:969-971 (mean), :871 (argmax),
// uncertainty_modeling/unc_mod_utils/test_utils.py:833-859 (TU / AU / EU).
#pragma once
#include "vu_common.cuh"

namespace vu {

// s / P for a member count P (2 <= P <= 271, rP = RN(1 / P)), bit-identical to the IEEE division that
// torch.mean performs (test_2D.py:971): q = RN(s * rP) followed by one residual correction.  Verified exhaustively
// against __fdiv_rn over all 2^24 mantissas of two binades for every P (bench/diag_div.cu: 0 mismatches; the
// uncorrected product differs 1.2e9 times).  Division is exponent-invariant away from under / overflow; outside
// [1e-30, 1e30] (and for NaN / inf) the IEEE divide is taken.
__device__ __forceinline__ float div_by_count(float s, float Pf, float rP) {
    float q = __fmul_rn(s, rP);
    const float a = fabsf(s);
    if (a > 1e-30f && a < 1e30f) q = __fmaf_rn(__fmaf_rn(-q, Pf, s), rP, q);
    else if (s != 0.0f) q = __fdiv_rn(s, Pf);
    return q;
}

// One thread owns VEC consecutive voxels.  The E = C * VEC values of one member are handled as E/2 packed
// fp32 pairs (FADD2 / FFMA2), flattened [class][voxel]: pairs of neighbouring voxels for VEC = 2, 4, pairs
// of classes for VEC = 1 (with one scalar leftover when C is odd).  The entropy of a member is accumulated
// class by class with one fma per term in every configuration.
//   LEVELS = 1: P <= 17 (plain sequential sum == torch's cascade)
//   LEVELS = 2: P <= 271 (level-0 accumulator folded into level 1 every 16 members)
template <int C, int VEC, int LEVELS>
struct VoxelAcc {
    static constexpr int E = C * VEC;
    static constexpr int NP = E / 2;
    static constexpr bool ODD = (E & 1);
    static constexpr int NH = VEC >= 2 ? VEC / 2 : 1;
    static constexpr int NP1 = LEVELS > 1 ? NP : 1;

    f32x2 m0[NP], m1[NP1];  // member sums, cascade level 0 / 1
    float m0s, m1s;         // scalar leftover class (ODD)
    f32x2 a0[NH], a1[NH];   // entropy sums (VEC >= 2: packed over voxels)
    float a0s, a1s;         // entropy sums (VEC == 1)
    float bv[VEC];          // argmax of the member being added (per-member labels, test_2D.py:814-818)
    int bi[VEC];

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < NP; ++j) m0[j] = 0ull;
#pragma unroll
        for (int j = 0; j < NP1; ++j) m1[j] = 0ull;
#pragma unroll
        for (int q = 0; q < NH; ++q) { a0[q] = 0ull; a1[q] = 0ull; }
        m0s = m1s = a0s = a1s = 0.f;
    }

    // ---- a member delivered in class chunks (VEC >= 2 only; used by the TMA kernel when one member's rows do
    // not fit a pipeline stage): begin_member, add_classes<C0, C1> for consecutive class ranges, end_member.
    // Operation order per voxel is the same as add_member's.
    f32x2 hm[NH];
    __device__ __forceinline__ void begin_member() {
#pragma unroll
        for (int q = 0; q < NH; ++q) hm[q] = 0ull;
    }
    // argmax over the classes [C0, C1) of one member, continuing from bv / bi when C0 > 0 (first max, NaN is max)
    template <int C0, int C1>
    __device__ __forceinline__ void member_argmax(const f32x2* xp, float xs) {
        if constexpr (VEC >= 2) {
#pragma unroll
            for (int c = C0; c < C1; ++c) {
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    float x0, x1;
                    upk2(xp[(c - C0) * NH + q], x0, x1);
                    if (c == 0) { bv[2 * q] = x0; bi[2 * q] = 0; bv[2 * q + 1] = x1; bi[2 * q + 1] = 0; }
                    else { argmax_step(x0, c, bv[2 * q], bi[2 * q]); argmax_step(x1, c, bv[2 * q + 1], bi[2 * q + 1]); }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                float x0, x1;
                upk2(xp[j], x0, x1);
                if (j == 0) { bv[0] = x0; bi[0] = 0; } else argmax_step(x0, 2 * j, bv[0], bi[0]);
                argmax_step(x1, 2 * j + 1, bv[0], bi[0]);
            }
            if constexpr (ODD) argmax_step(xs, C - 1, bv[0], bi[0]);
        }
    }
    template <int C0, int C1>
    __device__ __forceinline__ void add_classes(const f32x2 (&xp)[(C1 - C0) * NH], bool want_member_label = false) {
        static_assert(VEC >= 2, "class chunks need VEC >= 2");
        if (want_member_label) member_argmax<C0, C1>(xp, 0.f);
#pragma unroll
        for (int c = C0; c < C1; ++c) {
#pragma unroll
            for (int q = 0; q < NH; ++q) {
                const f32x2 X = xp[(c - C0) * NH + q];
                m0[c * NH + q] = add2(m0[c * NH + q], X);
                f32x2 PC, L;
                plog2p_parts2(X, PC, L);
                hm[q] = fma2(PC, L, hm[q]);
            }
        }
    }
    __device__ __forceinline__ void end_member(long long p) {
#pragma unroll
        for (int q = 0; q < NH; ++q) a0[q] = add2(a0[q], hm[q]);
        cascade_step(p);
    }
    __device__ __forceinline__ void cascade_step(long long p) {
        if (LEVELS > 1 && ((p & 15) == 15)) {
#pragma unroll
            for (int j = 0; j < NP1; ++j) { m1[j] = add2(m1[j], m0[j]); m0[j] = 0ull; }
#pragma unroll
            for (int q = 0; q < NH; ++q) { a1[q] = add2(a1[q], a0[q]); a0[q] = 0ull; }
            m1s = __fadd_rn(m1s, m0s); m0s = 0.f;
            a1s = __fadd_rn(a1s, a0s); a0s = 0.f;
        }
    }

    // member number p (0-based); xp: its NP pairs, xs: the leftover value when ODD
    __device__ __forceinline__ void add_member(const f32x2 (&xp)[NP], float xs, long long p, bool want_member_label = false) {
        if (want_member_label) member_argmax<0, C>(xp, xs);
        if constexpr (VEC >= 2) {
            f32x2 h[NH];
#pragma unroll
            for (int q = 0; q < NH; ++q) h[q] = 0ull;
#pragma unroll
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    const f32x2 X = xp[c * NH + q];
                    m0[c * NH + q] = add2(m0[c * NH + q], X);
                    f32x2 PC, L;
                    plog2p_parts2(X, PC, L);
                    h[q] = fma2(PC, L, h[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < NH; ++q) a0[q] = add2(a0[q], h[q]);
        } else {
            float h = 0.f;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const f32x2 X = xp[j];
                m0[j] = add2(m0[j], X);
                f32x2 PC, L;
                plog2p_parts2(X, PC, L);
                float pc0, pc1, l0, l1;
                upk2(PC, pc0, pc1);
                upk2(L, l0, l1);
                h = __fmaf_rn(pc0, l0, h);
                h = __fmaf_rn(pc1, l1, h);
            }
            if constexpr (ODD) { m0s = __fadd_rn(m0s, xs); h = plog2p_acc(h, xs); }
            a0s = __fadd_rn(a0s, h);
        }
        cascade_step(p);
    }

    // add_member for VEC >= 2 that also hands out L(p) = log2(max(p, FLT_MIN)) (near-one polynomial) of every element, for the
    // member-level likelihood sums folded into the pass (members_fold.cuh).  Same operations, same order.
    __device__ __forceinline__ void add_member_L(const f32x2 (&xp)[NP], long long p, bool want_member_label, f32x2 (&Lout)[NP]) {
        static_assert(VEC >= 2, "add_member_L needs VEC >= 2");
        if (want_member_label) member_argmax<0, C>(xp, 0.f);
        f32x2 h[NH];
#pragma unroll
        for (int q = 0; q < NH; ++q) h[q] = 0ull;
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int q = 0; q < NH; ++q) {
                const f32x2 X = xp[c * NH + q];
                m0[c * NH + q] = add2(m0[c * NH + q], X);
                f32x2 PC, L;
                plog2p_parts2(X, PC, L);
                Lout[c * NH + q] = L;
                h[q] = fma2(PC, L, h[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < NH; ++q) a0[q] = add2(a0[q], h[q]);
        cascade_step(p);
    }

    // ---- a member that arrives as LOGITS (VU_SLAB_LOGITS): softmax over its classes.  xp / xs become the exponentials e_c,
    // rs receives 1 / sum_c e_c of every voxel (packed like the voxels; VEC == 1: both halves the same) -- the probability
    // e_c * rs is not materialised: add_member_pre folds the product into the member sum with one fma, in every kernel form --
    // and hm / hs the member's entropy term sum_c p log2 p (see vu_common.cuh for the arithmetic and the special-value rules).
    // `reload(i)` returns the ORIGINAL pair i (reload_s(): the leftover value): only used on the rare path of a voxel with -inf
    // logits, whose 0 * -inf products have to be taken out of the e z sum.
    template <class Reload, class ReloadS>
    __device__ __forceinline__ void softmax_member(f32x2 (&xp)[NP], float& xs, f32x2 (&rs)[NH], f32x2 (&hm)[NH], float& hs, Reload reload,
                                                   ReloadS reload_s) {
        const f32x2 L2E = pk2(kLog2e, kLog2e);
        if constexpr (VEC >= 2) {
            float mx[VEC];
#pragma unroll
            for (int q = 0; q < NH; ++q) upk2(xp[q], mx[2 * q], mx[2 * q + 1]);
#pragma unroll
            for (int c = 1; c + 1 < C; c += 2) {
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    float a0, a1, b0, b1;
                    upk2(xp[c * NH + q], a0, a1);
                    upk2(xp[(c + 1) * NH + q], b0, b1);
                    mx[2 * q] = max_nan3(mx[2 * q], a0, b0);
                    mx[2 * q + 1] = max_nan3(mx[2 * q + 1], a1, b1);
                }
            }
            if constexpr ((C & 1) == 0) {
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    float a0, a1;
                    upk2(xp[(C - 1) * NH + q], a0, a1);
                    mx[2 * q] = max_nan(mx[2 * q], a0);
                    mx[2 * q + 1] = max_nan(mx[2 * q + 1], a1);
                }
            }
            f32x2 NM[NH], S[NH], EZ[NH];
#pragma unroll
            for (int q = 0; q < NH; ++q) { NM[q] = pk2(-mx[2 * q], -mx[2 * q + 1]); S[q] = 0ull; EZ[q] = 0ull; }
#pragma unroll
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    const f32x2 Z = mul2(add2(xp[c * NH + q], NM[q]), L2E);
                    float z0, z1;
                    upk2(Z, z0, z1);
                    const f32x2 E = pk2(ex2_approx(z0), ex2_approx(z1));
                    S[q] = add2(S[q], E);
                    EZ[q] = fma2(E, Z, EZ[q]);
                    xp[c * NH + q] = E;
                }
            }
            float ez[VEC];
            bool redo = false;
#pragma unroll
            for (int q = 0; q < NH; ++q) {
                upk2(EZ[q], ez[2 * q], ez[2 * q + 1]);
                redo |= (ez[2 * q] != ez[2 * q]) | (ez[2 * q + 1] != ez[2 * q + 1]);
            }
            if (__builtin_expect(redo, 0)) {  // a 0 * -inf product (or a NaN draw, for which this changes nothing): the e z sums term by term
                asm volatile("");  // (keeps this a branch: if-converted, its fmas would run for every member)
#pragma unroll
                for (int k = 0; k < VEC; ++k) ez[k] = 0.f;
#pragma unroll  // (a rolled loop would index xp dynamically and put it into local memory)
                for (int c = 0; c < C; ++c) {
#pragma unroll
                    for (int q = 0; q < NH; ++q) {
                        const f32x2 Z = mul2(add2(reload(c * NH + q), NM[q]), L2E);
                        float z0, z1, e0, e1;
                        upk2(Z, z0, z1);
                        upk2(xp[c * NH + q], e0, e1);
                        ez[2 * q] = (e0 == 0.0f) ? ez[2 * q] : __fmaf_rn(e0, z0, ez[2 * q]);
                        ez[2 * q + 1] = (e1 == 0.0f) ? ez[2 * q + 1] : __fmaf_rn(e1, z1, ez[2 * q + 1]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < NH; ++q) softmax_finish2(S[q], pk2(ez[2 * q], ez[2 * q + 1]), rs[q], hm[q]);
        } else {
            float m;
            {
                float a0, a1;
                upk2(xp[0], a0, a1);
                m = max_nan(a0, a1);
#pragma unroll
                for (int j = 1; j < NP; ++j) {
                    upk2(xp[j], a0, a1);
                    m = max_nan3(m, a0, a1);
                }
                if constexpr (ODD) m = max_nan(m, xs);
            }
            const f32x2 NM = pk2(-m, -m);
            float S = 0.f, EZ = 0.f;
#pragma unroll
            for (int j = 0; j < NP; ++j) {  // packed subtraction / scaling, the sums in class order
                const f32x2 Z = mul2(add2(xp[j], NM), L2E);
                float z0, z1;
                upk2(Z, z0, z1);
                const float e0 = ex2_approx(z0), e1 = ex2_approx(z1);
                S = __fadd_rn(S, e0);
                S = __fadd_rn(S, e1);
                EZ = __fmaf_rn(e0, z0, EZ);
                EZ = __fmaf_rn(e1, z1, EZ);
                xp[j] = pk2(e0, e1);
            }
            if constexpr (ODD) {
                const float z = softmax_z(xs, m);
                const float e = ex2_approx(z);
                S = __fadd_rn(S, e);
                EZ = __fmaf_rn(e, z, EZ);
                xs = e;
            }
            if (__builtin_expect(EZ != EZ, 0)) {
                asm volatile("");
                EZ = 0.f;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const f32x2 Z = mul2(add2(reload(j), NM), L2E);
                    float z0, z1, e0, e1;
                    upk2(Z, z0, z1);
                    upk2(xp[j], e0, e1);
                    EZ = (e0 == 0.0f) ? EZ : __fmaf_rn(e0, z0, EZ);
                    EZ = (e1 == 0.0f) ? EZ : __fmaf_rn(e1, z1, EZ);
                }
                if constexpr (ODD) EZ = (xs == 0.0f) ? EZ : __fmaf_rn(xs, softmax_z(reload_s(), m), EZ);
            }
            float rS;
            softmax_finish(S, EZ, rS, hs);
            rs[0] = pk2(rS, rS);
        }
    }
    // add a member that comes from softmax_member: member sum += e * rs (one fma per element), entropy sum += hm; per-member
    // labels are the argmax of the probabilities RN(e * rs), as the reference takes it (test_2D.py:814-818)
    __device__ __forceinline__ void add_member_pre(const f32x2 (&xp)[NP], float xs, const f32x2 (&rs)[NH], const f32x2 (&hm)[NH], float hs,
                                                   long long p, bool want_member_label = false) {
        if (want_member_label) {
            f32x2 pp[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) pp[j] = mul2(xp[j], rs[VEC >= 2 ? j % NH : 0]);
            float r0, r1;
            upk2(rs[0], r0, r1);
            member_argmax<0, C>(pp, __fmul_rn(xs, r0));
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) m0[j] = fma2(xp[j], rs[VEC >= 2 ? j % NH : 0], m0[j]);
        if constexpr (VEC >= 2) {
#pragma unroll
            for (int q = 0; q < NH; ++q) a0[q] = add2(a0[q], hm[q]);
        } else {
            if constexpr (ODD) {
                float r0, r1;
                upk2(rs[0], r0, r1);
                m0s = __fmaf_rn(xs, r0, m0s);
            }
            a0s = __fadd_rn(a0s, hs);
        }
        cascade_step(p);
    }

    // a member that is replaced by the one-hot vector of its argmax (--discretize, test_2D.py:1272-1275: F.one_hot(argmax)): the
    // member sum takes 1.0 at the label (first maximum, NaN is maximal: torch.argmax) and 0.0 elsewhere; its entropy term is 0
    // (1 log 1, and the skipped 0 log 0 terms), so the entropy sums are not touched.  bi holds the member's label afterwards.
    __device__ __forceinline__ void add_member_onehot(const f32x2 (&xp)[NP], float xs, long long p) {
        member_argmax<0, C>(xp, xs);
        if constexpr (VEC >= 2) {
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int q = 0; q < NH; ++q)
                    m0[c * NH + q] = add2(m0[c * NH + q], pk2(bi[2 * q] == c ? 1.0f : 0.0f, bi[2 * q + 1] == c ? 1.0f : 0.0f));
        } else {
#pragma unroll
            for (int j = 0; j < NP; ++j) m0[j] = add2(m0[j], pk2(bi[0] == 2 * j ? 1.0f : 0.0f, bi[0] == 2 * j + 1 ? 1.0f : 0.0f));
            if constexpr (ODD) m0s = __fadd_rn(m0s, bi[0] == C - 1 ? 1.0f : 0.0f);
        }
        cascade_step(p);
    }

    static constexpr float kExactLo = 2.524354896707238e-29f;  // 2^-95
    static constexpr unsigned kExactLoBits = 0x10000000u;
    // mean (true division, test_2D.py:971), label, TU, AU, EU of the VEC voxels
    __device__ __forceinline__ void finish(float Pf, float (&u)[VU_N_UNC][VEC], int (&label)[VEC]) const {
        const bool fast_div = Pf <= 271.0f;
        const float rP = __frcp_rn(Pf);
        // The sums of all classes first, then ONE test whether every one of them is in the range where the two-FMA division
        // is exact (zero is: it stays zero); the IEEE division is the rare, warp-divergent fallback for the whole thread.
        // (A range test with a branch per value cost more than the division itself: 8 branchy values per thread for C = 2.)
        float sums[C * VEC], asum[VEC];
#pragma unroll
        for (int e = 0; e < NP; ++e) {
            const f32x2 S = (LEVELS > 1) ? add2(m0[e], m1[e < NP1 ? e : 0]) : m0[e];
            upk2(S, sums[2 * e], sums[2 * e + 1]);
        }
        if constexpr (ODD) sums[E - 1] = (LEVELS > 1) ? __fadd_rn(m0s, m1s) : m0s;
        if constexpr (VEC >= 2) {
#pragma unroll
            for (int q = 0; q < NH; ++q) {
                const f32x2 A = (LEVELS > 1) ? add2(a0[q], a1[q]) : a0[q];
                upk2(A, asum[2 * q], asum[2 * q + 1]);
            }
        } else {
            asum[0] = (LEVELS > 1) ? __fadd_rn(a0s, a1s) : a0s;
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) asum[k] = -(asum[k] * kLn2);
        // in range: a == 0 or 2^-95 <= |a| <= 1e30, no NaN.  min.NaN(|a|, 2^-95) is 0 or 2^-95 (one bit: 0x10000000) exactly
        // when the lower bound holds (NaN stays NaN), so an OR over the bit patterns tests all values at once; the upper bound is
        // a test of the largest |a|.
        unsigned low_bits = 0u;
        float largest = 0.f;
        auto in_range = [&](float a) {
            float b;
            asm("min.NaN.f32 %0, %1, %2;" : "=f"(b) : "f"(fabsf(a)), "f"(kExactLo));
            low_bits |= __float_as_uint(b);
            largest = fmaxf(largest, fabsf(a));
        };
#pragma unroll
        for (int e = 0; e < E; ++e) in_range(sums[e]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) in_range(asum[k]);
        const bool exact = fast_div && (low_bits & (kExactLoBits - 1u)) == 0u && largest <= 1e30f;
        float qs[C * VEC], au[VEC];
        if (exact) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const float q = __fmul_rn(sums[e], rP);
                qs[e] = __fmaf_rn(__fmaf_rn(-q, Pf, sums[e]), rP, q);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float q = __fmul_rn(asum[k], rP);
                au[k] = __fmaf_rn(__fmaf_rn(-q, Pf, asum[k]), rP, q);
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) qs[e] = __fdiv_rn(sums[e], Pf);
#pragma unroll
            for (int k = 0; k < VEC; ++k) au[k] = __fdiv_rn(asum[k], Pf);
        }
        float mean[C][VEC], tu2[VEC];
        f32x2 T[NH];  // TU sums (VEC >= 2: packed over voxels, like the member entropies)
#pragma unroll
        for (int q = 0; q < NH; ++q) T[q] = 0ull;
        float ts = 0.f;
#pragma unroll
        for (int e = 0; e < NP; ++e) {
            const float q0 = qs[2 * e], q1 = qs[2 * e + 1];
            mean[(2 * e) / VEC][(2 * e) % VEC] = q0;
            mean[(2 * e + 1) / VEC][(2 * e + 1) % VEC] = q1;
            f32x2 PC, L;
            plog2p_parts2(pk2(q0, q1), PC, L);
            if constexpr (VEC >= 2) {
                T[e % NH] = fma2(PC, L, T[e % NH]);  // pair e = (class e / NH, voxels 2 (e % NH), +1): class order per voxel
            } else {
                float pc0, pc1, l0, l1;
                upk2(PC, pc0, pc1);
                upk2(L, l0, l1);
                ts = __fmaf_rn(pc0, l0, ts);
                ts = __fmaf_rn(pc1, l1, ts);
            }
        }
        if constexpr (ODD) {
            mean[C - 1][0] = qs[E - 1];
            ts = plog2p_acc(ts, mean[C - 1][0]);
        }
        if constexpr (VEC >= 2) {
#pragma unroll
            for (int q = 0; q < NH; ++q) upk2(T[q], tu2[2 * q], tu2[2 * q + 1]);
        } else {
            tu2[0] = ts;
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float best = mean[0][k];
            int idx = 0;
#pragma unroll
            for (int c = 1; c < C; ++c) argmax_step(mean[c][k], c, best, idx);
            const float tu = -(tu2[k] * kLn2);
            u[0][k] = tu; u[1][k] = au[k]; u[2][k] = tu - au[k];
            label[k] = idx;
        }
    }
};

}  // namespace vu
