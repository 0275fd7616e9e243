// Member-level scores folded into the unified-warp fused pass (k1_uni.cu): the GED pair counts of ged_binary_fast
// (evaluation/metrics/ged_fast.py:44-131) and the likelihood sums of _compute_likelihood_stats / _compute_expected_nll
// (uncertainty_modeling/test_2D.py:1043-1120), computed by the warps that stream the slab for the uncertainty maps, while the
// probabilities and their logarithms are in registers -- one read of the slab instead of two (vu_member_scores, k5_members.cu,
// is the stand-alone pass and the A/B reference of this code).
//
// Binary slabs (C == 2), P <= 32 members, R <= 4 uint8 references with word-aligned rows.  A lane owns the four consecutive
// voxels 4 lane + j of its warp's 128-voxel slice of a tile.
//   GED   per member and j one ballot: the label bits of 32 voxels; lane p keeps member p's four words.  After the last
//         member: pair counts popc(word & word) -- lane p accumulates row p of the P x P matrix (a shuffle per partner) and of
//         the P x R matrices, lane i < R row i of the R x R one.
//   NLL   the term of rater r at a voxel is  valid * (l0 + is1 * (l1 - l0)),  l = log2(max(p, eps)) (the pass's own lg2 +
//         near-one polynomial, clamped at log2 eps).  Where every reference of the tile is a class index the first part does
//         not depend on the rater: per member one sum of l0 and one FMA per rater and voxel on 0 / 1 float masks.  The five
//         per-lane values (sum l0, four raters) are folded 8 lanes to 1 by three shuffles and added to lane-private float32
//         columns in shared memory; ln 2 is applied once, in float64, when the warp leaves the image.
// All per-warp state lives in ms_warp_bytes<S>() of shared memory; a warp flushes on its own (global atomics) when it moves to
// another image.
#pragma once
#include "vu_common.cuh"

namespace vu {

struct MsParams {
    unsigned flags;      // VU_MS_NLL | VU_MS_GED, 0 = member scores not requested
    float log2eps;       // log2(eps), eps of test_2D.py:1043
    double* nll_sum;     // (B, R, P)
    unsigned long long* nll_cnt;  // (B, R)
    unsigned long long* nll_bad;  // (B)
    unsigned long long* ged;      // (B, ged_cols)
    int ged_cols;
};

constexpr int kMsP = 32, kMsR = 4;
constexpr int kMsVals = 1 + kMsR;  // per member: sum of l0 over the voxels, then one value per rater
// The five per-lane values of a member have to be summed over the lanes of the warp.  S = shuffle steps taken per member
// before the partial sums go to lane-private columns in shared memory: 32 >> S columns per (member pair, value), the lanes
// whose low S bits are zero own one each.  S = 0: no shuffles at all, 20 KB per warp; S = 3: three steps, 2.5 KB per warp.
// The columns hold the values of two members side by side (f32x2): one packed add serves both.
template <int S> struct MsGeom {
    static constexpr int kCols = 32 >> S;
    static constexpr int kNllBytes = (kMsP / 2) * kMsVals * kCols * 8;
};
// counter words per lane (two 16-bit counters each; a warp flushes at least every 255 tiles of 128 voxels)
enum { MC_PG = 0 /* [r]: pg_tp | pg_pred << 16 */, MC_GG = 4 /* [r]: gg_tp | gg_sum << 16 */, MC_GS = 8 /* [r]: g_sum | valid << 16 */,
       MC_POS = 12 /* pos | bad << 16 */, MC_MAJ = 13 /* tp | pred << 16 */, MC_MAJG = 14, MC_N = 15 };
constexpr int kMsPpBytes = kMsP * 32 * 2, kMsCntBytes = MC_N * 32 * 4;
template <int S> constexpr int ms_warp_bytes() { return MsGeom<S>::kNllBytes + kMsPpBytes + kMsCntBytes; }

__device__ __forceinline__ unsigned lds_u32(unsigned a) { unsigned r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned lds_u16(unsigned a) { unsigned short r; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(a)); return r; }
__device__ __forceinline__ void sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void smem_add_u32(unsigned a, unsigned v) { sts_u32(a, lds_u32(a) + v); }

// add a warp's partials into the rows of image b (layout: valunc.h, vu_member_scores) and clear them.  Out of line and on
// plain addresses, so that the per-tile state of MsWarp stays in registers.
__device__ __noinline__ void ms_flush_warp(unsigned nll_a, int cols, unsigned pp_a, unsigned cnt_a, const MsParams& ms, long long b, int P,
                                           int R, bool want_major) {
    const unsigned lane = threadIdx.x & 31;
    auto cnt = [&](int c, unsigned l) { return lds_u32(cnt_a + ((unsigned)c * 32u + l) * 4u); };
    __syncwarp();
    if (ms.flags & VU_MS_NLL) {
        if (lane < (unsigned)P) {
            // lane p: member p is half (p & 1) of the packed columns of pair p >> 1; the lanes start at different columns so
            // that they do not all hit one bank
            double s[kMsVals];
#pragma unroll
            for (int i = 0; i < kMsVals; ++i) {
                s[i] = 0.0;
                for (int c0 = 0; c0 < cols; ++c0) {
                    const unsigned c = (c0 + lane) & (unsigned)(cols - 1);
                    const unsigned a = nll_a + ((((lane >> 1) * kMsVals + i) * cols + c) * 2u + (lane & 1u)) * 4u;
                    s[i] += (double)__uint_as_float(lds_u32(a));
                    sts_u32(a, 0u);
                }
            }
            for (int r = 0; r < R; ++r) {
                const double t = (s[0] + s[1 + r]) * 0.6931471805599453;  // log2 -> ln
                if (t != 0.0) atomicAdd(ms.nll_sum + (b * R + r) * P + lane, t);
            }
        }
        for (int r = 0; r < R; ++r) {
            const unsigned n = __reduce_add_sync(kFull, cnt(MC_GS + r, lane) >> 16);
            if (lane == 0 && n) atomicAdd(ms.nll_cnt + b * R + r, (unsigned long long)n);
        }
        const unsigned nb = __reduce_add_sync(kFull, cnt(MC_POS, lane) >> 16);
        if (lane == 0 && nb) atomicAdd(ms.nll_bad + b, (unsigned long long)nb);
    }
    if (ms.flags & VU_MS_GED) {
        const int G = R;
        const int o_pg_pred = P * G, o_gs = 2 * P * G, o_pp = o_gs + G, o_pos = o_pp + P * P, o_gg_tp = o_pos + P, o_gg_sum = o_gg_tp + G * G,
                  o_maj = o_gg_sum + G * G;
        unsigned long long* row = ms.ged + b * ms.ged_cols;
        if (lane < (unsigned)P) {
            for (int q = 0; q < P; ++q) {
                const unsigned c = lds_u16(pp_a + ((unsigned)q * 32u + lane) * 2u);
                if (c) atomicAdd(row + o_pp + lane * P + q, (unsigned long long)c);
            }
            const unsigned ps = cnt(MC_POS, lane) & 0xffffu;
            if (ps) atomicAdd(row + o_pos + lane, (unsigned long long)ps);
            for (int r = 0; r < G; ++r) {
                const unsigned a = cnt(MC_PG + r, lane);
                if (a & 0xffffu) atomicAdd(row + lane * G + r, (unsigned long long)(a & 0xffffu));
                if (a >> 16) atomicAdd(row + o_pg_pred + lane * G + r, (unsigned long long)(a >> 16));
            }
        }
        if (lane < (unsigned)G) {
            for (int r = 0; r < G; ++r) {
                const unsigned g = cnt(MC_GG + r, lane);
                if (g & 0xffffu) atomicAdd(row + o_gg_tp + lane * G + r, (unsigned long long)(g & 0xffffu));
                if (g >> 16) atomicAdd(row + o_gg_sum + lane * G + r, (unsigned long long)(g >> 16));
            }
        }
        if (lane == 0) {
            for (int r = 0; r < G; ++r) {
                const unsigned g = cnt(MC_GS + r, 0) & 0xffffu;
                if (g) atomicAdd(row + o_gs + r, (unsigned long long)g);
            }
            if (want_major) {
                const unsigned m = cnt(MC_MAJ, 0), mg = cnt(MC_MAJG, 0);
                if (m & 0xffffu) atomicAdd(row + o_maj, (unsigned long long)(m & 0xffffu));
                if (m >> 16) atomicAdd(row + o_maj + 1, (unsigned long long)(m >> 16));
                if (mg) atomicAdd(row + o_maj + 2, (unsigned long long)mg);
            }
        }
    }
    __syncwarp();
    for (unsigned i = lane; i < (kMsPpBytes + kMsCntBytes) / 4; i += 32) sts_u32(pp_a + 4u * i, 0u);
    __syncwarp();
}

template <int S>
struct MsWarp {
    static constexpr int kCols = MsGeom<S>::kCols, kNllBytes = MsGeom<S>::kNllBytes, kWarpBytes = ms_warp_bytes<S>();
    unsigned nll_a, pp_a, cnt_a;  // shared-memory addresses: nll f32x2 [member pair][value][column], pp u16 [partner q][member =
                                  // lane], cnt u32 [counter][lane]
    f32x2* nll;                   // the same as a pointer (hot loop)
    // per tile
    f32x2 m1[2][4];       // (rater pair, voxel): 1.0 where the reference is class 1 and valid
    float mv;             // 1.0 for a lane inside the image
    unsigned actw;        // ballot of the lanes inside the image
    unsigned okb, oneb, rawb, clsb;  // bit (4 r + j): reference valid / valid and 1 / 1 / valid and a class index (0 or 1)
    unsigned myW[4];      // lane p: label bits of member p, one word per voxel slot j
    bool fast;            // warp-uniform: every reference of the tile is a class index (no ignore value, nothing out of range)

    __device__ __forceinline__ void init(unsigned char* smem_warp) {
        nll = reinterpret_cast<f32x2*>(smem_warp);
        nll_a = (unsigned)__cvta_generic_to_shared(smem_warp);
        pp_a = nll_a + kNllBytes;
        cnt_a = pp_a + kMsPpBytes;
        const int lane = threadIdx.x & 31;
        for (int i = lane; i < kWarpBytes / 4; i += 32) sts_u32(nll_a + 4u * i, 0u);
        __syncwarp();
    }

    // W[r]: the reference bytes of the lane's four voxels (stats2_load_refs); active: the lane is inside the image
    __device__ __forceinline__ void tile_begin(const unsigned (&W)[kMsR], bool active, const GtView& gt, bool want_nll) {
        const unsigned lane = threadIdx.x & 31;
        okb = oneb = rawb = clsb = 0u;
        unsigned bad = 0u, nvalid[kMsR];
        bool all01 = true;
#pragma unroll
        for (int r = 0; r < kMsR; ++r) {
            nvalid[r] = 0u;
            if (r < gt.R && active) {
                const unsigned valid_hi = gt.ign_byte ? bytes_nonzero(W[r] ^ gt.ign4) : kB80;
                const unsigned raw_hi = ~bytes_nonzero(W[r] ^ kB01) & kB80;
                const unsigned zero_hi = ~bytes_nonzero(W[r]) & kB80;
                const unsigned one_hi = raw_hi & valid_hi;
                const unsigned cls_hi = (raw_hi | zero_hi) & valid_hi;  // a class index of a binary slab
                all01 = all01 && cls_hi == kB80;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    okb |= ((valid_hi >> (8 * j + 7)) & 1u) << (4 * r + j);
                    oneb |= ((one_hi >> (8 * j + 7)) & 1u) << (4 * r + j);
                    rawb |= ((raw_hi >> (8 * j + 7)) & 1u) << (4 * r + j);
                    clsb |= ((cls_hi >> (8 * j + 7)) & 1u) << (4 * r + j);
                }
                nvalid[r] = __popc(valid_hi);
                bad += __popc(valid_hi & ~cls_hi);  // torch.gather would raise (test_2D.py:1067)
            }
        }
        mv = active ? 1.0f : 0.0f;
        actw = __ballot_sync(kFull, active);
#pragma unroll
        for (int rp = 0; rp < 2; ++rp)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                m1[rp][j] = pk2(((oneb >> (8 * rp + j)) & 1u) ? 1.0f : 0.0f, ((oneb >> (8 * rp + 4 + j)) & 1u) ? 1.0f : 0.0f);
        fast = __all_sync(kFull, all01);
#pragma unroll
        for (int j = 0; j < 4; ++j) myW[j] = 0u;
        if (want_nll) {
#pragma unroll
            for (int r = 0; r < kMsR; ++r)
                if (nvalid[r]) smem_add_u32(cnt_a + ((MC_GS + r) * 32 + lane) * 4u, nvalid[r] << 16);
            if (bad) smem_add_u32(cnt_a + (MC_POS * 32 + lane) * 4u, bad << 16);
        }
    }

    // the general form of a member's likelihood terms: every term on its own, with torch's NaN rule (torch.clamp keeps NaN)
    // and only where the reference is a class index.  Rare (tiles with ignored voxels, members with NaN values): out of line.
    // (Inlined although it is cold: out of line its array arguments -- and with them the hot path's values -- live in local memory.)
    __device__ __forceinline__ static void member_general(float (&v)[kMsVals], const float (&x0)[4], const float (&x1)[4], const float (&l0)[4],
                                                       const float (&l1)[4], unsigned clsb, unsigned oneb, int R, float log2eps) {
        v[0] = 0.f;
#pragma unroll
        for (int r = 0; r < kMsR; ++r) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // (a valid reference that is neither 0 nor 1 is counted in `bad` and contributes nothing)
                const bool cls = (clsb >> (4 * r + j)) & 1u, one = (oneb >> (4 * r + j)) & 1u;
                const float xs = one ? x1[j] : x0[j];
                const float ls = fmaxf(one ? l1[j] : l0[j], log2eps);
                if (r < R && cls) a += (xs != xs) ? xs : ls;
            }
            v[1 + r] = a;
        }
    }

    // label bits of member p (GED): x = its values (pairs: class 0 voxels 01, 23, class 1 voxels 01, 23)
    __device__ __forceinline__ void member_labels(int p, const f32x2 (&x)[4]) {
        const bool mine = (threadIdx.x & 31) == (unsigned)p;
        float x0[4], x1[4];
        upk2(x[0], x0[0], x0[1]); upk2(x[1], x0[2], x0[3]);
        upk2(x[2], x1[0], x1[1]); upk2(x[3], x1[2], x1[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // torch.argmax over two classes: 1 iff p1 > p0, or p1 is NaN and p0 is not (ged_fast.py:44)
            // (lanes outside the image vote on stale bytes: their bits are masked out once per tile, in tile_end)
            const unsigned w = __ballot_sync(kFull, !(x1[j] <= x0[j]) && (x0[j] == x0[j]));
            myW[j] = mine ? w : myW[j];
        }
    }

    // the five per-lane likelihood values of one member (sum of l0 over the lane's voxels, then one value per rater);
    // x as above, L = log2(max(x, FLT_MIN)) of the same elements
    // 0 unless one of the member's values is NaN (or +inf and -inf meet): such a member takes the general form, because a NaN
    // probability makes its sums NaN (torch.clamp keeps NaN) while the pass's L does not (it clamps first)
    __device__ __forceinline__ static float nan_probe(const f32x2 (&x)[4]) {
        float lo, hi;
        upk2(add2(add2(x[0], x[1]), add2(x[2], x[3])), lo, hi);
        return (lo + hi) * 0.0f;
    }

    // slow: warp-uniform, the tile or the member needs the general form
    __device__ __forceinline__ void member_values(const f32x2 (&x)[4], const f32x2 (&L)[4], int R, float log2eps, bool slow, float (&v)[kMsVals]) {
        float l0[4], l1[4], lo, hi;
        upk2(L[0], l0[0], l0[1]); upk2(L[1], l0[2], l0[3]);
        upk2(L[2], l1[0], l1[1]); upk2(L[3], l1[2], l1[3]);
        if (!slow) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { l0[j] = fmaxf(l0[j], log2eps); l1[j] = fmaxf(l1[j], log2eps); }
            const float d[4] = {l1[0] - l0[0], l1[1] - l0[1], l1[2] - l0[2], l1[3] - l0[3]};
            v[0] = ((l0[0] + l0[1]) + (l0[2] + l0[3])) * mv;
            f32x2 A0 = 0ull, A1 = 0ull;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const f32x2 dj = pk2(d[j], d[j]);
                A0 = fma2(m1[0][j], dj, A0);
                A1 = fma2(m1[1][j], dj, A1);
            }
            upk2(A0, v[1], v[2]);
            upk2(A1, v[3], v[4]);
        } else {
            float x0[4], x1[4];
            upk2(x[0], x0[0], x0[1]); upk2(x[1], x0[2], x0[3]);
            upk2(x[2], x1[0], x1[1]); upk2(x[3], x1[2], x1[3]);
            member_general(v, x0, x1, l0, l1, clsb, oneb, R, log2eps);
        }
    }

    // Add the values of the two members p (even) and p + 1 to the warp's columns: S shuffle steps over the lanes (the two
    // members travel as packed pairs), then the lanes whose low S bits are zero add the partial sums to their column.
    // Without a member p + 1 (odd member count) vb is zero.
    __device__ __forceinline__ void fold_pair(int p, const float (&va)[kMsVals], const float (&vb)[kMsVals]) {
        const unsigned lane = threadIdx.x & 31;
        f32x2 V[kMsVals];
#pragma unroll
        for (int i = 0; i < kMsVals; ++i) V[i] = pk2(va[i], vb[i]);
#pragma unroll
        for (int st = 0; st < S; ++st) {
#pragma unroll
            for (int i = 0; i < kMsVals; ++i) {
                float a, b;
                upk2(V[i], a, b);
                V[i] = add2(V[i], pk2(__shfl_xor_sync(kFull, a, 1 << st), __shfl_xor_sync(kFull, b, 1 << st)));
            }
        }
        if (S == 0 || (lane & ((1u << S) - 1u)) == 0u) {
            f32x2* c = nll + ((unsigned)(p >> 1) * kMsVals) * kCols + (lane >> S);
#pragma unroll
            for (int i = 0; i < kMsVals; ++i) c[i * kCols] = add2(c[i * kCols], V[i]);
        }
    }

    // after the last member of the tile: lab1 = bit j set where the label of the member mean of voxel j is 1 (majority Dice)
    __device__ __forceinline__ void tile_end(int P, int R, unsigned lab1, bool has_ignore, bool ign_is_one, bool want_major) {
        const unsigned lane = threadIdx.x & 31;
        unsigned pos = 0u, pg_tp[kMsR], pg_pred[kMsR], gg_tp[kMsR], gg_sum[kMsR], g_sum[kMsR], maj_tp = 0u, maj_pred = 0u, maj_gt = 0u;
#pragma unroll
        for (int r = 0; r < kMsR; ++r) { pg_tp[r] = 0u; pg_pred[r] = 0u; gg_tp[r] = 0u; gg_sum[r] = 0u; g_sum[r] = 0u; }
        // pair counts member x member: lane p against every partner q
#pragma unroll
        for (int j = 0; j < 4; ++j) { myW[j] &= actw; pos += __popc(myW[j]); }
        for (int q = 0; q < P; ++q) {
            unsigned c = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) c += __popc(myW[j] & __shfl_sync(kFull, myW[j], q));
            const unsigned a = pp_a + ((unsigned)q * 32u + lane) * 2u;
            if (c) sts_u16(a, lds_u16(a) + c);  // lanes >= P hold zero words
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned w = myW[j];
            unsigned myGb = 0u, n1 = 0u, all_valid = 1u;
            unsigned Gw[kMsR], Vw[kMsR];
#pragma unroll
            for (int r = 0; r < kMsR; ++r) {
                Gw[r] = 0u; Vw[r] = 0u;
                if (r < R) {
                    const unsigned ok = (okb >> (4 * r + j)) & 1u, one = (oneb >> (4 * r + j)) & 1u;
                    // ged_fast.py:93 takes (gt == 1) before masking: differs from `one` only when the ignore value itself is 1
                    const unsigned raw1 = ign_is_one ? (rawb >> (4 * r + j)) & 1u : one;
                    const unsigned Gb = ign_is_one ? __ballot_sync(kFull, raw1) : 0u;
                    Gw[r] = __ballot_sync(kFull, one);
                    Vw[r] = has_ignore ? __ballot_sync(kFull, ok) : actw;
                    myGb = (lane == (unsigned)r) ? (ign_is_one ? Gb : Gw[r]) : myGb;
                    n1 += raw1;
                    all_valid &= ok;
                    pg_tp[r] += __popc(w & Gw[r]);
                    pg_pred[r] += __popc(w & Vw[r]);
                    g_sum[r] += __popc(Gw[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < kMsR; ++r) {
                if (r < R) {
                    gg_tp[r] += __popc(myGb & Gw[r]);
                    gg_sum[r] += __popc(myGb & Vw[r]);
                }
            }
            if (want_major) {
                // ged_fast.py:121-131: majority reference (share of raters with label 1 >= 0.5) on voxels no rater ignores
                const bool active = (actw >> lane) & 1u;
                const unsigned Mg = __ballot_sync(kFull, active && 2u * n1 >= (unsigned)R);
                const unsigned Va = has_ignore ? __ballot_sync(kFull, active && all_valid) : actw;
                const unsigned Lw = __ballot_sync(kFull, (lab1 >> j) & 1u);
                maj_tp += __popc(Lw & Mg & Va);
                maj_pred += __popc(Lw & Va);
                maj_gt += __popc(Mg & Va);
            }
        }
        // lane p: rows of member p; lane i < R: row i of the rater x rater matrices; lane 0: the per-rater and majority totals
#pragma unroll
        for (int r = 0; r < kMsR; ++r) {
            if (r < R) {
                const unsigned a = pg_tp[r] | (pg_pred[r] << 16), g = (lane < (unsigned)R) ? (gg_tp[r] | (gg_sum[r] << 16)) : 0u;
                if (a) smem_add_u32(cnt_a + ((MC_PG + r) * 32 + lane) * 4u, a);
                if (g) smem_add_u32(cnt_a + ((MC_GG + r) * 32 + lane) * 4u, g);
                if (lane == 0 && g_sum[r]) smem_add_u32(cnt_a + ((MC_GS + r) * 32) * 4u, g_sum[r]);
            }
        }
        if (pos) smem_add_u32(cnt_a + (MC_POS * 32 + lane) * 4u, pos);
        if (lane == 0 && want_major) {
            smem_add_u32(cnt_a + (MC_MAJ * 32) * 4u, maj_tp | (maj_pred << 16));
            smem_add_u32(cnt_a + (MC_MAJG * 32) * 4u, maj_gt);
        }
    }

    __device__ __forceinline__ void flush(const MsParams& ms, long long b, int P, int R, bool want_major) {
        ms_flush_warp(nll_a, kCols, pp_a, cnt_a, ms, b, P, R, want_major);
    }
};

}  // namespace vu
