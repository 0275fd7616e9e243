// K5: member-level scores on the slab K1 reads:
//   likelihood statistics   uncertainty_modeling/test_2D.py:1043-1120  (_compute_likelihood_stats, _compute_expected_nll)
//   GED counts              evaluation/metrics/ged_fast.py:44-131      (ged_binary_fast)
//
// One pass over the (P, B, C, V) slab and the (B, R, V) references.  A warp owns "passes" of kGroups x 32 voxels of one
// image (lane = voxel inside a group) and walks the members in order, kGroups x C loads in flight per lane:
//   NLL   per member the lane picks log(max(p[gt_r], eps)) for each rater (C == 2: both logs, selected by the reference's
//         bit; C > 2: a gather load per rater), sums its kGroups voxels in float32 and the warp folds the R sums of the
//         member into the image's (R, P) float64 matrix with one atomic each;
//   GED   (C == 2, P <= 32) the member's label bits of a group are one ballot word; lane p keeps member p's words.  After
//         the last member every pair count is popc(word & word): lane p accumulates row p of the P x P intersection
//         matrix (a shuffle per partner), of the P x R prediction / reference counts, lane i < R row i of the R x R
//         reference matrix -- no cross-lane reduction until the CTA is done with the image.
// All counts are integers (bit-exact); the float32 Dice / GED arithmetic of ged_fast.py:60-140 runs on the host on
// P*R + P*P + R*R numbers.  Bound: HBM (the slab is read once more: 4 P C + R g bytes per voxel).
#include "vu_common.cuh"
#include "vu_host.h"

namespace vu {

struct MemberParams {
    const float* data;
    const float* const* member_ptrs;
    long long P, B, C, V;
    long long sp, sb, sc, sv;
    GtView gt;
    const uint8_t* labels;  // (B, V) label of the member mean, or NULL
    unsigned flags;
    float eps;
    double* nll_sum;              // (B, R, P)
    unsigned long long* nll_cnt;  // (B, R)
    unsigned long long* nll_bad;  // (B)
    unsigned long long* ged;      // (B, cols)
    int ged_cols;
    long long chunk;              // voxels per CTA (multiple of 32 * kGroups)
};

constexpr int kMsThreads = 256, kMsWarps = kMsThreads / 32, kGroups = 4;
constexpr int kAhead = 6;             // members whose lines are prefetched into L1 ahead of the one being processed
constexpr unsigned kInvalid = 0xffu;  // class byte of an ignored reference

// ln(max(p, eps)) with torch.clamp's NaN rule (NaN stays NaN).  lg2.approx has an absolute error of ~2^-23.5 on
// [0.5, 2], too much next to p = 1 where the term itself vanishes; there the K1 polynomial takes over (vu_common.cuh).
__device__ __forceinline__ float log_clamped(float p, float eps) {
    const float x = fmaxf(p, eps);
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(x));
    const float f = __fadd_rn(x, -1.0f);
    float t = __fmaf_rn(f, kQ3, kQ2);
    t = __fmaf_rn(t, f, kQ1);
    t = __fmaf_rn(t, f, kQ0);
    const float l2 = (fabsf(f) < kNearOne) ? __fmul_rn(t, f) : lg;
    return (p != p) ? p : l2 * kLn2;
}

__device__ __forceinline__ long long load_ref(const GtView& gt, long long off) {
    return gt.dtype == VU_GT_U8 ? (long long)__ldg(reinterpret_cast<const uint8_t*>(gt.data) + off)
                                : __ldg(reinterpret_cast<const long long*>(gt.data) + off);
}

// Fold one warp's per-lane GED accumulators into the CTA's shared counters (layout: valunc.h, vu_ged_cols).
__device__ __forceinline__ void ged_fold_warp(unsigned* s_ged, int lane, int P, int G, const unsigned (&pp)[32], unsigned pos,
                                              const unsigned (&pg_tp)[VU_MAX_RATERS], const unsigned (&pg_pred)[VU_MAX_RATERS],
                                              const unsigned (&gg_tp)[VU_MAX_RATERS], const unsigned (&gg_sum)[VU_MAX_RATERS],
                                              const unsigned (&g_sum)[VU_MAX_RATERS], unsigned maj_tp, unsigned maj_pred, unsigned maj_gt) {
    const int o_pg_tp = 0, o_pg_pred = P * G, o_gs = 2 * P * G, o_pp = o_gs + G, o_pos = o_pp + P * P, o_gg_tp = o_pos + P,
              o_gg_sum = o_gg_tp + G * G, o_maj = o_gg_sum + G * G;
    (void)o_pg_tp;
    if (lane < P) {
#pragma unroll
        for (int q = 0; q < 32; ++q)
            if (q < P && pp[q]) atomicAdd(&s_ged[o_pp + lane * P + q], pp[q]);
        if (pos) atomicAdd(&s_ged[o_pos + lane], pos);
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r)
            if (r < G) {
                if (pg_tp[r]) atomicAdd(&s_ged[o_pg_tp + lane * G + r], pg_tp[r]);
                if (pg_pred[r]) atomicAdd(&s_ged[o_pg_pred + lane * G + r], pg_pred[r]);
            }
    }
    if (lane < G) {
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r)
            if (r < G) {
                if (gg_tp[r]) atomicAdd(&s_ged[o_gg_tp + lane * G + r], gg_tp[r]);
                if (gg_sum[r]) atomicAdd(&s_ged[o_gg_sum + lane * G + r], gg_sum[r]);
            }
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r)
            if (r < G && g_sum[r]) atomicAdd(&s_ged[o_gs + r], g_sum[r]);
        if (maj_tp) atomicAdd(&s_ged[o_maj], maj_tp);
        if (maj_pred) atomicAdd(&s_ged[o_maj + 1], maj_pred);
        if (maj_gt) atomicAdd(&s_ged[o_maj + 2], maj_gt);
    }
}

template <bool C2, bool GED>
__global__ void __launch_bounds__(kMsThreads) member_scores_kernel(const __grid_constant__ MemberParams prm) {
    extern __shared__ unsigned s_ged[];  // [ged_cols] CTA-wide GED counters (C2 && GED only)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long b = blockIdx.y;
    const long long c0 = (long long)blockIdx.x * prm.chunk;
    const long long c1 = c0 + prm.chunk < prm.V ? c0 + prm.chunk : prm.V;
    const int P = (int)prm.P, R = prm.gt.R;
    const bool want_nll = prm.flags & VU_MS_NLL;
    constexpr bool want_ged = C2 && GED;
    if (want_ged) {
        for (int t = tid; t < prm.ged_cols; t += kMsThreads) s_ged[t] = 0u;
        __syncthreads();
    }
    // per-lane GED accumulators: lane p holds row p of the P x P and P x R matrices, lane i < R row i of the R x R ones
    unsigned pp[32], pg_tp[VU_MAX_RATERS], pg_pred[VU_MAX_RATERS], gg_tp[VU_MAX_RATERS], gg_sum[VU_MAX_RATERS], g_sum[VU_MAX_RATERS];
    unsigned pos = 0, maj_tp = 0, maj_pred = 0, maj_gt = 0;
    if (want_ged) {
#pragma unroll
        for (int q = 0; q < 32; ++q) pp[q] = 0u;
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r) { pg_tp[r] = 0u; pg_pred[r] = 0u; gg_tp[r] = 0u; gg_sum[r] = 0u; g_sum[r] = 0u; }
    }
    unsigned valid_cnt[VU_MAX_RATERS], bad_cnt = 0;  // per CTA: at most chunk <= 2^30 voxels
    double nll_mine[VU_MAX_RATERS];  // P <= 32: lane p keeps member p's sums until the CTA is done (no atomics per pass)
#pragma unroll
    for (int r = 0; r < VU_MAX_RATERS; ++r) { valid_cnt[r] = 0u; nll_mine[r] = 0.0; }
    const bool nll_in_lanes = P <= 32;

    const long long pass_vox = 32LL * kGroups;
    for (long long v0 = c0 + (long long)warp * pass_vox; v0 < c1; v0 += (long long)kMsWarps * pass_vox) {
        // ---- references of the pass: one class byte per (rater, group); kInvalid = ignored or out of range ----------
        unsigned cls[VU_MAX_RATERS][kGroups / 4];  // 4 group bytes per word
        unsigned long long is1 = 0, val = 0;       // bit (r * kGroups + j): reference == 1 (and valid) / valid
        unsigned lab1 = 0;                         // bit j: label of the member mean == 1
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r) {
#pragma unroll
            for (int w = 0; w < kGroups / 4; ++w) cls[r][w] = 0u;
            if (r < R) {
#pragma unroll
                for (int j = 0; j < kGroups; ++j) {
                    const long long v = v0 + 32 * j + lane;
                    unsigned c = kInvalid;
                    if (v < c1) {
                        const long long g = load_ref(prm.gt, b * prm.gt.sb + (long long)r * prm.gt.sr + v * prm.gt.sv);
                        const bool valid = !(prm.gt.has_ignore && g == prm.gt.ignore);
                        if (valid) {
                            val |= 1ull << (r * kGroups + j);
                            if (g == 1) is1 |= 1ull << (r * kGroups + j);
                            valid_cnt[r] += 1;
                            if (g >= 0 && g < prm.C) c = (unsigned)g;
                            else if (want_nll) bad_cnt += 1;  // torch.gather would raise (test_2D.py:1067)
                        }
                    }
                    cls[r][j >> 2] |= c << (8 * (j & 3));
                }
            }
        }
        if (want_ged && prm.labels) {
#pragma unroll
            for (int j = 0; j < kGroups; ++j) {
                const long long v = v0 + 32 * j + lane;
                if (v < c1 && __ldg(prm.labels + b * prm.V + v) == 1) lab1 |= 1u << j;
            }
        }
        // ---- members ---------------------------------------------------------------------------------------------
        unsigned myW[kGroups];  // lane p: label bits of member p, one word per group
#pragma unroll
        for (int j = 0; j < kGroups; ++j) myW[j] = 0u;
        float n0[kGroups], n1[kGroups];  // C == 2: the next member's two class values per group
        if (C2) {
            const float* base0 = (prm.member_ptrs ? ld_member_ptr(prm.member_ptrs, 0) : prm.data) + b * prm.sb;
#pragma unroll
            for (int j = 0; j < kGroups; ++j) {
                const long long v = v0 + 32 * j + lane;
                const bool in = v < c1;
                n0[j] = in ? __ldg(base0 + v * prm.sv) : 1.f;
                n1[j] = in ? __ldg(base0 + prm.sc + v * prm.sv) : 0.f;
            }
            for (int q = 1; q < kAhead && q < P; ++q) {
                const float* fbase = (prm.member_ptrs ? ld_member_ptr(prm.member_ptrs, q) : prm.data + (long long)q * prm.sp) + b * prm.sb;
#pragma unroll
                for (int j = 0; j < kGroups; ++j) {
                    const long long v = v0 + 32 * j + lane;
                    if (v < c1) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fbase + v * prm.sv));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fbase + prm.sc + v * prm.sv));
                    }
                }
            }
        }
        for (int p = 0; p < P; ++p) {
            const float* base = (prm.member_ptrs ? ld_member_ptr(prm.member_ptrs, p) : prm.data + (long long)p * prm.sp) + b * prm.sb;
            float acc[VU_MAX_RATERS];
#pragma unroll
            for (int r = 0; r < VU_MAX_RATERS; ++r) acc[r] = 0.f;
            if (C2) {
                // this member's values were requested while the previous member was processed (n0 / n1 below)
                float p0[kGroups], p1[kGroups];
#pragma unroll
                for (int j = 0; j < kGroups; ++j) { p0[j] = n0[j]; p1[j] = n1[j]; }
                if (p + 1 < P) {
                    const float* nbase = (prm.member_ptrs ? ld_member_ptr(prm.member_ptrs, p + 1) : prm.data + (long long)(p + 1) * prm.sp) + b * prm.sb;
#pragma unroll
                    for (int j = 0; j < kGroups; ++j) {
                        const long long v = v0 + 32 * j + lane;
                        const bool in = v < c1;
                        n0[j] = in ? __ldg(nbase + v * prm.sv) : 1.f;
                        n1[j] = in ? __ldg(nbase + prm.sc + v * prm.sv) : 0.f;
                    }
                }
                if (p + kAhead < P) {
                    // ... and the lines of the member kAhead further on are pulled into L1 now: with ~230 registers per thread
                    // only 8 warps fit on an SM, far too few loads in flight to cover the HBM latency otherwise
                    const float* fbase = (prm.member_ptrs ? ld_member_ptr(prm.member_ptrs, p + kAhead) : prm.data + (long long)(p + kAhead) * prm.sp) + b * prm.sb;
#pragma unroll
                    for (int j = 0; j < kGroups; ++j) {
                        const long long v = v0 + 32 * j + lane;
                        if (v < c1) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(fbase + v * prm.sv));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(fbase + prm.sc + v * prm.sv));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < kGroups; ++j) {
                    if (want_ged) {
                        // torch.argmax over two classes: 1 iff p1 > p0, or p1 is NaN and p0 is not (ged_fast.py:44)
                        const bool one = (p1[j] > p0[j]) || ((p1[j] != p1[j]) && (p0[j] == p0[j]));
                        const unsigned w = __ballot_sync(kFull, one && (v0 + 32 * j + lane < c1));
                        if (lane == p) myW[j] = w;
                    }
                    if (want_nll) {
                        const float l0 = log_clamped(p0[j], prm.eps), l1 = log_clamped(p1[j], prm.eps);
#pragma unroll
                        for (int r = 0; r < VU_MAX_RATERS; ++r) {
                            if (r >= R) break;
                            const unsigned c = (cls[r][j >> 2] >> (8 * (j & 3))) & 0xffu;
                            acc[r] += (c == 0u) ? l0 : ((c == 1u) ? l1 : 0.f);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < VU_MAX_RATERS; ++r) {
                    if (r >= R) break;
                    float x[kGroups];
#pragma unroll
                    for (int j = 0; j < kGroups; ++j) {
                        const unsigned c = (cls[r][j >> 2] >> (8 * (j & 3))) & 0xffu;
                        const long long v = v0 + 32 * j + lane;
                        x[j] = (c != kInvalid) ? __ldg(base + (long long)c * prm.sc + v * prm.sv) : 1.f;  // ln 1 = 0
                    }
#pragma unroll
                    for (int j = 0; j < kGroups; ++j) acc[r] += log_clamped(x[j], prm.eps);
                }
            }
            if (want_nll) {
#pragma unroll
                for (int r = 0; r < VU_MAX_RATERS; ++r) {
                    if (r >= R) break;
                    float s = acc[r];  // 32 x kGroups terms per pass in float32 (butterfly: every lane ends up with the sum) ...
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
                    if (nll_in_lanes) nll_mine[r] += (lane == p) ? (double)s : 0.0;  // ... the passes in float64
                    else if (lane == 0 && s != 0.f) atomicAdd(prm.nll_sum + (b * R + r) * P + p, (double)s);
                }
            }
        }
        // ---- pair counts of the pass -------------------------------------------------------------------------------
        if (want_ged) {
#pragma unroll
            for (int j = 0; j < kGroups; ++j) {
                const unsigned w = myW[j];
                pos += __popc(w);
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    if (q >= P) break;
                    pp[q] += __popc(w & __shfl_sync(kFull, w, q));
                }
                unsigned Gb[VU_MAX_RATERS], Gw[VU_MAX_RATERS], Vw[VU_MAX_RATERS];
                unsigned myGb = 0u, n1 = 0u, all_valid = 1u;
#pragma unroll
                for (int r = 0; r < VU_MAX_RATERS; ++r) {
                    if (r >= R) break;
                    const unsigned one = (unsigned)(is1 >> (r * kGroups + j)) & 1u, ok = (unsigned)(val >> (r * kGroups + j)) & 1u;
                    // is1 is only set on valid voxels; ged_fast.py:93 takes (gt == 1) before masking, which differs only
                    // when the ignore value itself is 1
                    const unsigned raw1 = (prm.gt.has_ignore && prm.gt.ignore == 1) ? (ok ^ 1u) & (unsigned)(v0 + 32 * j + lane < c1) : one;
                    Gb[r] = __ballot_sync(kFull, raw1);
                    Gw[r] = __ballot_sync(kFull, one);
                    Vw[r] = __ballot_sync(kFull, ok);
                    myGb = (lane == r) ? Gb[r] : myGb;
                    n1 += raw1;
                    all_valid &= ok;
                }
#pragma unroll
                for (int r = 0; r < VU_MAX_RATERS; ++r) {
                    if (r >= R) break;
                    pg_tp[r] += __popc(w & Gw[r]);
                    pg_pred[r] += __popc(w & Vw[r]);
                    gg_tp[r] += __popc(myGb & Gw[r]);
                    gg_sum[r] += __popc(myGb & Vw[r]);
                    g_sum[r] += __popc(Gw[r]);
                }
                if (prm.labels) {
                    // ged_fast.py:121-131: majority reference (share of raters with label 1 >= 0.5) on voxels no rater ignores
                    const bool in = v0 + 32 * j + lane < c1;
                    const unsigned Mg = __ballot_sync(kFull, in && 2u * n1 >= (unsigned)R);
                    const unsigned Va = __ballot_sync(kFull, in && (prm.gt.has_ignore ? all_valid : 1u));
                    const unsigned Lw = __ballot_sync(kFull, (lab1 >> j) & 1u);
                    maj_tp += __popc(Lw & Mg & Va);
                    maj_pred += __popc(Lw & Va);
                    maj_gt += __popc(Mg & Va);
                }
            }
        }
    }
    // ---- fold the CTA's partials into the image's rows ---------------------------------------------------------------
    if (want_nll) {
        if (nll_in_lanes && lane < P) {
#pragma unroll
            for (int r = 0; r < VU_MAX_RATERS; ++r)
                if (r < R && nll_mine[r] != 0.0) atomicAdd(prm.nll_sum + (b * R + r) * P + lane, nll_mine[r]);
        }
#pragma unroll
        for (int r = 0; r < VU_MAX_RATERS; ++r) {
            if (r >= R) break;
            const unsigned long long n = (unsigned long long)warp_sum((double)valid_cnt[r]);  // exact: far below 2^53
            if (lane == 0 && n) atomicAdd(prm.nll_cnt + b * R + r, n);
        }
        const unsigned long long nb = (unsigned long long)warp_sum((double)bad_cnt);
        if (lane == 0 && nb) atomicAdd(prm.nll_bad + b, nb);
    }
    if (want_ged) {
        ged_fold_warp(s_ged, lane, P, R, pp, pos, pg_tp, pg_pred, gg_tp, gg_sum, g_sum, maj_tp, maj_pred, maj_gt);
        __syncthreads();
        for (int t = tid; t < prm.ged_cols; t += kMsThreads)
            if (s_ged[t]) atomicAdd(prm.ged + b * prm.ged_cols + t, (unsigned long long)s_ged[t]);
    }
}

// ---- binary fast path: 4 consecutive voxels per lane -----------------------------------------------------------------
// Same algorithm for C == 2 slabs whose voxel rows can be read as float4 (stride_v == 1, 16-byte aligned rows).  ncu of the
// kernel above on configs[3] (r01e): 197 instructions per (member, voxel), issue-bound with 8 warps per SM.  Here
//   * a lane owns voxels v0 + 4 lane + j, j < 4: one 128-bit load per class and member, requested two members ahead, with an
//     L2 prefetch kAheadL2 members ahead;
//   * the likelihood term of rater r is  in01 * l0 + is1 * (l1 - l0)  with 0 / 1 float masks built once per pass: two fused
//     multiply-adds per (member, voxel, rater) instead of byte extraction, compares and selects;
//   * the per-(member, rater) sums stay in lane-private columns of shared memory ([member * R + r][lane], float32, at most
//     chunk / 32 terms each) and are folded into the image's float64 matrix once per CTA.
// RMAX: raters the mask registers are sized for.
constexpr int kAheadL2 = 6;

constexpr int kAccLanes = 8;  // likelihood columns per (member, rater): the 32 lane sums are folded 4 to 1 before they are stored

template <bool GED, int RMAX>
__global__ void __launch_bounds__(kMsThreads, 2) member_scores_c2v4(const __grid_constant__ MemberParams prm) {
    extern __shared__ __align__(16) unsigned char ms_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, warps = blockDim.x >> 5;
    const long long b = blockIdx.y;
    const long long c0 = (long long)blockIdx.x * prm.chunk;
    const long long c1 = c0 + prm.chunk < prm.V ? c0 + prm.chunk : prm.V;
    const int P = (int)prm.P, R = prm.gt.R;
    const bool want_nll = prm.flags & VU_MS_NLL;
    const int n_acc = want_nll ? P * R : 0;
    // shared memory: [warps][n_acc][kAccLanes] float likelihood columns, [warps][P][P] pair counts (GED), [ged_cols] CTA counters
    float* nll_all = reinterpret_cast<float*>(ms_smem);
    float* nll_acc = nll_all + (size_t)warp * n_acc * kAccLanes;
    unsigned* pp_all = reinterpret_cast<unsigned*>(nll_all + (size_t)warps * n_acc * kAccLanes);
    unsigned* pp = pp_all + (size_t)warp * P * P;  // [partner q][member = lane]
    unsigned* s_ged = pp_all + (GED ? (size_t)warps * P * P : 0);
    const int n_words = warps * n_acc * kAccLanes + (GED ? warps * P * P + prm.ged_cols : 0);
    for (int t = tid; t < n_words; t += blockDim.x) reinterpret_cast<unsigned*>(ms_smem)[t] = 0u;
    __syncthreads();

    unsigned pg_tp[RMAX], pg_pred[RMAX], gg_tp[RMAX], gg_sum[RMAX], g_sum[RMAX];
    unsigned pos = 0, maj_tp = 0, maj_pred = 0, maj_gt = 0;
    unsigned valid_cnt[RMAX], bad_cnt = 0;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) { pg_tp[r] = 0u; pg_pred[r] = 0u; gg_tp[r] = 0u; gg_sum[r] = 0u; g_sum[r] = 0u; valid_cnt[r] = 0u; }
    const bool ign_is_one = prm.gt.has_ignore && prm.gt.ignore == 1;

    for (long long v0 = c0 + 128LL * warp; v0 < c1; v0 += 128LL * warps) {
        const long long v = v0 + 4 * lane;
        const bool in = v < c1;  // V % 4 == 0 and the chunk is a multiple of 128: a lane's four voxels are all in or all out
        const long long vs = in ? v : v0;  // lanes past the end read the warp's first voxels (valid memory) and are masked out
        // ---- references: bit (4 r + j) of okb / oneb / rawb = valid / valid & == 1 / == 1, and the likelihood masks --------
        unsigned okb = 0, oneb = 0, rawb = 0, lab1 = 0;
        float m01[RMAX][4], m1[RMAX][4];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { m01[r][j] = 0.f; m1[r][j] = 0.f; }
            if (r < R && in) {
                long long g[4];
                if (prm.gt.dtype == VU_GT_U8) {
                    const unsigned w = load_gt_bytes<4>(prm.gt, b, r, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) g[j] = (long long)((w >> (8 * j)) & 0xffu);
                } else {
                    load_gt<4>(prm.gt, b, r, v, g, (long long)0);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool valid = !(prm.gt.has_ignore && g[j] == prm.gt.ignore);
                    const bool one = g[j] == 1;
                    okb |= valid ? (1u << (4 * r + j)) : 0u;
                    oneb |= (valid && one) ? (1u << (4 * r + j)) : 0u;
                    rawb |= one ? (1u << (4 * r + j)) : 0u;
                    valid_cnt[r] += valid;
                    const bool cls01 = valid && (g[j] == 0 || one);
                    bad_cnt += (valid && !cls01 && want_nll);  // torch.gather would raise (test_2D.py:1067)
                    m01[r][j] = cls01 ? 1.f : 0.f;
                    m1[r][j] = (valid && one) ? 1.f : 0.f;
                }
            }
        }
        if (GED && prm.labels && in) {
            const uint8_t* lp = prm.labels + b * prm.V + v;
#pragma unroll
            for (int j = 0; j < 4; ++j) lab1 |= (__ldg(lp + j) == 1) ? (1u << j) : 0u;
        }
        // ---- members: values requested two members ahead; the row pointer just advances by the member stride ---------------
        const long long row = b * prm.sb + vs;
        auto member_row = [&](int p) {
            return (prm.member_ptrs ? ld_member_ptr(prm.member_ptrs, p) : prm.data + (long long)p * prm.sp) + row;
        };
        const float* q = member_row(0);
        float4 a0 = __ldg(reinterpret_cast<const float4*>(q)), a1 = __ldg(reinterpret_cast<const float4*>(q + prm.sc));
        float4 b0 = a0, b1 = a1;  // member p (a) and p + 1 (b)
        if (P > 1) {
            q = member_row(1);
            b0 = __ldg(reinterpret_cast<const float4*>(q));
            b1 = __ldg(reinterpret_cast<const float4*>(q + prm.sc));
        }
        for (int f = 2; f < kAheadL2 && f < P; ++f) {
            const float* pf = member_row(f);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + prm.sc));
        }
        unsigned myW[4] = {0u, 0u, 0u, 0u};
        for (int p = 0; p < P; ++p) {
            const float x0[4] = {a0.x, a0.y, a0.z, a0.w}, x1[4] = {a1.x, a1.y, a1.z, a1.w};
            a0 = b0; a1 = b1;
            if (p + 2 < P) {
                q = member_row(p + 2);
                b0 = __ldg(reinterpret_cast<const float4*>(q));
                b1 = __ldg(reinterpret_cast<const float4*>(q + prm.sc));
                if (p + kAheadL2 < P) {
                    const float* pf = member_row(p + kAheadL2);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + prm.sc));
                }
            }
            if (GED) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // torch.argmax over two classes: 1 iff p1 > p0, or p1 is NaN and p0 is not (ged_fast.py:44)
                    const bool one = !(x1[j] <= x0[j]) && (x0[j] == x0[j]);
                    const unsigned w = __ballot_sync(kFull, one && in);
                    if (lane == p) myW[j] = w;
                }
            }
            if (want_nll) {
                float l0[4], d[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    l0[j] = log_clamped(x0[j], prm.eps);
                    d[j] = log_clamped(x1[j], prm.eps) - l0[j];
                }
                // A NaN probability must only reach the sums of the raters whose reference selects it (torch.gather picks
                // before the sum, test_2D.py:1064-1068); through the 0 / 1 masks it would reach all of them (0 * NaN = NaN).
                const float chk = (((x0[0] + x0[1]) + (x0[2] + x0[3])) + ((x1[0] + x1[1]) + (x1[2] + x1[3]))) * 0.0f;
                const bool has_nan = __any_sync(kFull, chk != 0.0f);
                float* col = nll_acc + p * R * kAccLanes + (lane >> 2);
#pragma unroll
                for (int r = 0; r < RMAX; ++r) {
                    if (r >= R) break;
                    float acc = 0.f;
                    if (has_nan) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (m01[r][j] != 0.f) acc += (m1[r][j] != 0.f) ? l0[j] + d[j] : l0[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc = fmaf(m1[r][j], d[j], fmaf(m01[r][j], l0[j], acc));
                    }
                    acc += __shfl_xor_sync(kFull, acc, 1);
                    acc += __shfl_xor_sync(kFull, acc, 2);
                    if ((lane & 3) == 0) col[r * kAccLanes] += acc;
                }
            }
        }
        // ---- pair counts of the pass ---------------------------------------------------------------------------------------
        if (GED) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned w = myW[j];
                pos += __popc(w);
                for (int qq = 0; qq < P; ++qq) {
                    const unsigned c = __popc(w & __shfl_sync(kFull, w, qq));
                    if (lane < P && c) pp[qq * P + lane] += c;  // the matrix is symmetric: [partner][member] is conflict-free
                }
                unsigned myGb = 0u, n1 = 0u, all_valid = 1u;
#pragma unroll
                for (int r = 0; r < RMAX; ++r) {
                    if (r >= R) break;
                    const unsigned ok = (okb >> (4 * r + j)) & 1u, one = (oneb >> (4 * r + j)) & 1u;
                    // ged_fast.py:93 takes (gt == 1) before masking: differs from `one` only when the ignore value itself is 1
                    const unsigned raw1 = ign_is_one ? (rawb >> (4 * r + j)) & 1u : one;
                    const unsigned Gb = __ballot_sync(kFull, raw1), Gw = __ballot_sync(kFull, one), Vw = __ballot_sync(kFull, ok);
                    myGb = (lane == r) ? Gb : myGb;
                    n1 += raw1;
                    all_valid &= ok;
                    pg_tp[r] += __popc(w & Gw);
                    pg_pred[r] += __popc(w & Vw);
                    g_sum[r] += __popc(Gw);
                }
                // second sweep: lane i < R needs its own reference word against every rater's
#pragma unroll
                for (int r = 0; r < RMAX; ++r) {
                    if (r >= R) break;
                    const unsigned ok = (okb >> (4 * r + j)) & 1u, one = (oneb >> (4 * r + j)) & 1u;
                    const unsigned Gw = __ballot_sync(kFull, one), Vw = __ballot_sync(kFull, ok);
                    gg_tp[r] += __popc(myGb & Gw);
                    gg_sum[r] += __popc(myGb & Vw);
                }
                if (prm.labels) {
                    // ged_fast.py:121-131: majority reference (share of raters with label 1 >= 0.5) on voxels no rater ignores
                    const unsigned Mg = __ballot_sync(kFull, in && 2u * n1 >= (unsigned)R);
                    const unsigned Va = __ballot_sync(kFull, in && (prm.gt.has_ignore ? all_valid : 1u));
                    const unsigned Lw = __ballot_sync(kFull, (lab1 >> j) & 1u);
                    maj_tp += __popc(Lw & Mg & Va);
                    maj_pred += __popc(Lw & Va);
                    maj_gt += __popc(Mg & Va);
                }
            }
        }
    }
    // ---- fold the CTA's partials into the image's rows ---------------------------------------------------------------------
    if (want_nll) {
        __syncwarp();
        for (int i0 = 0; i0 < n_acc; i0 += 4) {  // 4 (member, rater) pairs x 8 columns per round; i = member * R + r
            const int i = i0 + (lane >> 3);
            double sum = i < n_acc ? (double)nll_acc[i * kAccLanes + (lane & 7)] : 0.0;
            sum += __shfl_xor_sync(kFull, sum, 1);
            sum += __shfl_xor_sync(kFull, sum, 2);
            sum += __shfl_xor_sync(kFull, sum, 4);
            if ((lane & 7) == 0 && i < n_acc && sum != 0.0) atomicAdd(prm.nll_sum + (b * R + i % R) * P + i / R, sum);
        }
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            if (r >= R) break;
            const unsigned long long n = (unsigned long long)warp_sum((double)valid_cnt[r]);  // exact: far below 2^53
            if (lane == 0 && n) atomicAdd(prm.nll_cnt + b * R + r, n);
        }
        const unsigned long long nb = (unsigned long long)warp_sum((double)bad_cnt);
        if (lane == 0 && nb) atomicAdd(prm.nll_bad + b, nb);
    }
    if (GED) {
        // this warp's per-lane counters and pair matrix -> the CTA's counters (layout: valunc.h, vu_ged_cols) -> the image's row
        const int G = R;
        const int o_pg_pred = P * G, o_gs = 2 * P * G, o_pp = o_gs + G, o_pos = o_pp + P * P, o_gg_tp = o_pos + P,
                  o_gg_sum = o_gg_tp + G * G, o_maj = o_gg_sum + G * G;
        __syncwarp();
        for (int t = lane; t < P * P; t += 32)
            if (pp[t]) atomicAdd(&s_ged[o_pp + t], pp[t]);
        if (lane < P) {
            if (pos) atomicAdd(&s_ged[o_pos + lane], pos);
#pragma unroll
            for (int r = 0; r < RMAX; ++r)
                if (r < G) {
                    if (pg_tp[r]) atomicAdd(&s_ged[lane * G + r], pg_tp[r]);
                    if (pg_pred[r]) atomicAdd(&s_ged[o_pg_pred + lane * G + r], pg_pred[r]);
                }
        }
        if (lane < G) {
#pragma unroll
            for (int r = 0; r < RMAX; ++r)
                if (r < G) {
                    if (gg_tp[r]) atomicAdd(&s_ged[o_gg_tp + lane * G + r], gg_tp[r]);
                    if (gg_sum[r]) atomicAdd(&s_ged[o_gg_sum + lane * G + r], gg_sum[r]);
                }
        }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < RMAX; ++r)
                if (r < G && g_sum[r]) atomicAdd(&s_ged[o_gs + r], g_sum[r]);
            if (maj_tp) atomicAdd(&s_ged[o_maj], maj_tp);
            if (maj_pred) atomicAdd(&s_ged[o_maj + 1], maj_pred);
            if (maj_gt) atomicAdd(&s_ged[o_maj + 2], maj_gt);
        }
        __syncthreads();
        for (int t = tid; t < prm.ged_cols; t += blockDim.x)
            if (s_ged[t]) atomicAdd(prm.ged + b * prm.ged_cols + t, (unsigned long long)s_ged[t]);
    }
}

int launch_member_scores(const vu_member_scores_args* a, const GtView& gv, cudaStream_t stream) {
    const vu_slab& s = a->slab;
    MemberParams prm;
    prm.data = s.data; prm.member_ptrs = s.member_ptrs;
    prm.P = s.P; prm.B = s.B; prm.C = s.C; prm.V = s.V;
    prm.sp = s.stride_p; prm.sb = s.stride_b; prm.sc = s.stride_c; prm.sv = s.stride_v;
    prm.gt = gv;
    prm.labels = a->labels;
    prm.flags = a->flags;
    prm.eps = a->eps;
    prm.nll_sum = a->nll_sum;
    prm.nll_cnt = reinterpret_cast<unsigned long long*>(a->nll_count);
    prm.nll_bad = reinterpret_cast<unsigned long long*>(a->nll_bad);
    prm.ged = reinterpret_cast<unsigned long long*>(a->ged_counts);
    prm.ged_cols = (a->flags & VU_MS_GED) ? (int)vu_ged_cols((int)s.P, gv.R) : 0;
    // CTAs per image: enough CTAs to fill the GPU, each with at least one pass per warp, counters below 2^32
    const long long pass = 32LL * kGroups, cta_min = pass * kMsWarps;
    long long per_image = (2LL * device_sm_count() * 4 + s.B - 1) / s.B;
    const long long max_per_image = (s.V + cta_min - 1) / cta_min;
    if (per_image > max_per_image) per_image = max_per_image;
    if (per_image < 1) per_image = 1;
    long long chunk = (s.V + per_image - 1) / per_image;
    chunk = (chunk + pass - 1) / pass * pass;
    if (chunk > (1LL << 30)) chunk = 1LL << 30;
    per_image = (s.V + chunk - 1) / chunk;
    prm.chunk = chunk;
    if (s.B > 65535) return set_error(VU_ERR_UNSUPPORTED, "B > 65535 per vu_member_scores call");
    // binary fast path: rows readable as float4, lane-private likelihood accumulators fit into shared memory
    bool fast = s.C == 2 && s.stride_v == 1 && s.V % 4 == 0 && s.stride_c % 4 == 0 && s.stride_b % 4 == 0 && get_option("k5_path", 0) != 1;
    if (fast) {
        if (s.member_ptrs_host) {
            for (long long p = 0; p < s.P; ++p) fast = fast && ((uintptr_t)s.member_ptrs_host[p] % 16) == 0;
        } else {
            fast = fast && ((uintptr_t)s.data % 16) == 0 && s.stride_p % 4 == 0;
        }
    }
    int warps = kMsWarps;
    const size_t acc_per_warp = ((a->flags & VU_MS_NLL) ? (size_t)s.P * gv.R * kAccLanes * sizeof(float) : 0) +
                                ((a->flags & VU_MS_GED) ? (size_t)s.P * s.P * sizeof(unsigned) : 0);
    const size_t ged_bytes = (size_t)prm.ged_cols * sizeof(unsigned);
    while (fast && warps > 1 && acc_per_warp * warps + ged_bytes > 100 * 1024) warps >>= 1;  // two CTAs per SM
    if (fast && acc_per_warp * warps + ged_bytes > 200 * 1024) fast = false;
    if (fast) {
        // the pass is 128 voxels per warp here; re-derive the chunk for the CTA size chosen above
        const long long cta_vox = 128LL * warps;
        long long pi = (2LL * device_sm_count() * 4 + s.B - 1) / s.B;
        const long long max_pi = (s.V + cta_vox - 1) / cta_vox;
        if (pi > max_pi) pi = max_pi;
        if (pi < 1) pi = 1;
        long long ch = (s.V + pi - 1) / pi;
        ch = (ch + 127) / 128 * 128;
        if (ch > (1LL << 22)) ch = 1LL << 22;  // float32 columns: at most 2^17 terms each
        prm.chunk = ch;
        const dim3 fgrid((unsigned)((s.V + ch - 1) / ch), (unsigned)s.B);
        const size_t fsmem = acc_per_warp * warps + ged_bytes;
        const bool ged = a->flags & VU_MS_GED, r4 = gv.R <= 4;
        void (*fn)(const MemberParams) = ged ? (r4 ? member_scores_c2v4<true, 4> : member_scores_c2v4<true, 8>)
                                             : (r4 ? member_scores_c2v4<false, 4> : member_scores_c2v4<false, 8>);
        if (fsmem > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem) != cudaSuccess)
            return set_cuda_error("cudaFuncSetAttribute(member_scores_c2v4)");
        fn<<<fgrid, warps * 32, fsmem, stream>>>(prm);
        count_launch("member_scores_c2v4");
        return check_launch("member_scores_c2v4");
    }
    dim3 grid((unsigned)per_image, (unsigned)s.B);
    const size_t smem = (size_t)prm.ged_cols * sizeof(unsigned);
    if (s.C == 2 && (a->flags & VU_MS_GED)) member_scores_kernel<true, true><<<grid, kMsThreads, smem, stream>>>(prm);
    else if (s.C == 2) member_scores_kernel<true, false><<<grid, kMsThreads, 0, stream>>>(prm);
    else member_scores_kernel<false, false><<<grid, kMsThreads, 0, stream>>>(prm);
    count_launch("member_scores");
    return check_launch("member_scores");
}

}  // namespace vu
