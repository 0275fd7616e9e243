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
    template <int C0, int C1>
    __device__ __forceinline__ void add_classes(const f32x2 (&xp)[(C1 - C0) * NH]) {
        static_assert(VEC >= 2, "class chunks need VEC >= 2");
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
    __device__ __forceinline__ void add_member(const f32x2 (&xp)[NP], float xs, long long p) {
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

    // mean (true division, test_2D.py:971), label, TU, AU, EU of the VEC voxels
    __device__ __forceinline__ void finish(float Pf, float (&u)[VU_N_UNC][VEC], int (&label)[VEC]) const {
        float mean[C][VEC], asum[VEC];
#pragma unroll
        for (int e = 0; e < NP; ++e) {
            const f32x2 S = (LEVELS > 1) ? add2(m0[e], m1[e < NP1 ? e : 0]) : m0[e];
            float s0, s1;
            upk2(S, s0, s1);
            mean[(2 * e) / VEC][(2 * e) % VEC] = __fdiv_rn(s0, Pf);
            mean[(2 * e + 1) / VEC][(2 * e + 1) % VEC] = __fdiv_rn(s1, Pf);
        }
        if constexpr (ODD) mean[C - 1][0] = __fdiv_rn((LEVELS > 1) ? __fadd_rn(m0s, m1s) : m0s, Pf);
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
        for (int k = 0; k < VEC; ++k) {
            float best = mean[0][k], tu2 = 0.f;
            int idx = 0;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (c > 0) argmax_step(mean[c][k], c, best, idx);
                tu2 = plog2p_acc(tu2, mean[c][k]);
            }
            const float tu = -(tu2 * kLn2);
            const float au = (-(asum[k] * kLn2)) / Pf;
            u[0][k] = tu; u[1][k] = au; u[2][k] = tu - au;
            label[k] = idx;
        }
    }
};

}  // namespace vu
