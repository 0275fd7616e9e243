// K4: exact (weighted) order statistics by radix select, and calibration histograms on arbitrary bin edges.
//
//   vu_radix_hist     one histogram pass of a 3-level radix select (11 + 11 + 10 bits of the order-preserving
//                     float key).  The host walks the cumulative counts between passes (quantile.py), so any
//                     number of maps can be folded into one selection -- find_threshold.py:98-105 takes
//                     np.quantile over the concatenation of all validation maps -- and samples can carry an
//                     integer weight (the number of valid raters of a voxel: ace.py:378-406 ranks one confidence
//                     per (rater, pixel) pair).
//   vu_binned_calib   the three bincounts of calc_eqace (ace.py:392-396) on 19 caller-given thresholds (the
//                     per-image quantile edges pulled back onto the uncertainty axis); float64 sums.
#include "vu_common.cuh"
#include "vu_host.h"
#include <string.h>

namespace vu {

// order-preserving key of a float's bit pattern: key(a) < key(b) <=> a < b; NaN (either sign) sorts last, like np.sort
__device__ __forceinline__ unsigned bits2key(unsigned u) {
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffc00000u;  // the key of the canonical quiet NaN
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// number of raters whose reference is not the ignore value (weight of voxel i); 1 without references
__device__ __forceinline__ int sample_weight(const GtView& gt, long long i) {
    if (!gt.data) return 1;
    int w = 0;
    for (int r = 0; r < gt.R; ++r) {
        const long long off = (long long)r * gt.sr + i * gt.sv;
        const long long g = gt.dtype == VU_GT_U8 ? (long long)__ldg(reinterpret_cast<const uint8_t*>(gt.data) + off)
                                                 : __ldg(reinterpret_cast<const long long*>(gt.data) + off);
        w += !(gt.has_ignore && g == gt.ignore);
    }
    return w;
}

constexpr int kRadixThreads = 256;
constexpr int kMaxPrefixes = 64;

// Batched form (vu_quantile_select_batch / vu_binned_calib_batch): segment s = m * B + b is image b of map m; every segment
// has its own histogram block, state, thresholds and outputs, the references are those of image b.  n_maps == 0: one array.
struct SegBatch {
    int n_maps, B;
    const float* maps[4];
};
__device__ __forceinline__ void seg_gt(GtView& gt, long long b) {
    if (gt.data) gt.data = reinterpret_cast<const char*>(gt.data) + b * gt.sb * (gt.dtype == VU_GT_U8 ? 1 : 8);
}

__global__ void __launch_bounds__(kRadixThreads) radix_hist_kernel(const float* __restrict__ values, long long n, GtView gt, int level,
                                                                  const unsigned* __restrict__ prefixes, int n_prefix,
                                                                  unsigned long long* hist, const vu_radix_state* state, SegBatch seg) {
    __shared__ unsigned h0[2048];
    __shared__ unsigned pre[kMaxPrefixes];
    if (seg.n_maps) {
        const int sgm = blockIdx.y, m = sgm / seg.B, b = sgm - m * seg.B;
        values = seg.maps[m] + (long long)b * n;
        seg_gt(gt, b);
        hist += (size_t)sgm * kMaxPrefixes * 2048;
        state += sgm;
    }
    if (state && level > 0) {  // prefixes left on the device by vu_radix_walk
        n_prefix = state->n_slot;
        prefixes = state->slot_prefix;
    }
    if (level == 0)
        for (int t = threadIdx.x; t < 2048; t += kRadixThreads) h0[t] = 0u;
    else
        for (int t = threadIdx.x; t < n_prefix; t += kRadixThreads) pre[t] = prefixes[t];
    __syncthreads();
    if (level > 0 && n_prefix == 0) return;
    const int lane = threadIdx.x & 31;
    // every lane of a warp runs the same number of iterations (the aggregation below is warp-collective)
    const long long stride = (long long)gridDim.x * kRadixThreads;
    for (long long base = (long long)blockIdx.x * kRadixThreads; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        unsigned target = 0xffffffffu;  // slot << 11 | digit, or "nothing to add"
        int w = 0;
        if (i < n) {
            const unsigned key = bits2key(__ldg(reinterpret_cast<const unsigned*>(values) + i));  // never touched as a float
            w = sample_weight(gt, i);
            if (w > 0) {
                if (level == 0) {
                    target = key >> 21;
                } else {
                    const unsigned p = level == 1 ? key >> 21 : key >> 10;
                    const unsigned digit = level == 1 ? (key >> 10) & 0x7ffu : key & 0x3ffu;
                    for (int s = 0; s < n_prefix; ++s)
                        if (pre[s] == p) { target = ((unsigned)s << 11) | digit; break; }
                }
            }
        }
        if (level == 0) {
            if (target != 0xffffffffu) atomicAdd(&h0[target], (unsigned)w);
        } else {
            // warp-aggregated: equal values are common (background pixels with u == 0), one global atomic per group
            const unsigned grp = __match_any_sync(kFull, target);
            unsigned sum = 0;
#pragma unroll
            for (int bit = 0; bit < 4; ++bit) sum += (unsigned)__popc(grp & __ballot_sync(kFull, (w >> bit) & 1)) << bit;
            if (target != 0xffffffffu && lane == __ffs(grp) - 1) atomicAdd(hist + target, (unsigned long long)sum);
        }
    }
    if (level == 0) {
        __syncthreads();
        for (int t = threadIdx.x; t < 2048; t += kRadixThreads)
            if (h0[t]) atomicAdd(hist + t, (unsigned long long)h0[t]);
    }
}

// The descent of every rank by one level (see vu_radix_walk in valunc.h).  One CTA of 32 warps; a warp takes a rank at a time:
// every lane sums 64 consecutive counters of the rank's slot (independent loads), a warp scan finds the lane in whose run the
// residual rank falls, and that lane walks its 64 counters.  (A single thread walking 2048 counters is a chain of dependent
// global loads: ~0.2 ms per level.)
struct WalkQ { double q[32]; };
constexpr int kWalkThreads = 1024;
__global__ void __launch_bounds__(kWalkThreads) radix_walk_kernel(const unsigned long long* __restrict__ hist, int level, WalkQ wq, int n_q,
                                                                   int q_is_f32, int reverse, vu_radix_state* st, int seg_B) {
    if (seg_B) {  // batched: one CTA per segment; `reverse` is a bit mask over the maps
        st += blockIdx.x;
        hist += (size_t)blockIdx.x * kMaxPrefixes * 2048;
        reverse = (reverse >> (blockIdx.x / seg_B)) & 1;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ long long s_part[32];
    __shared__ long long s_total;
    __shared__ unsigned s_prefix[VU_RADIX_MAX_RANKS];
    __shared__ int s_n_rank;
    if (level == 0) {
        long long part = 0;
        for (int i = tid; i < 2048; i += kWalkThreads) part += (long long)hist[i];
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
        if (lane == 0) s_part[warp] = part;
        __syncthreads();
        if (tid == 0) {
            long long t = 0;
            for (int w = 0; w < 32; ++w) t += s_part[w];
            s_total = t;
            st->total = t;
            st->n_rank = t > 0 ? 2 * n_q + 1 : 0;
            if (t <= 0) st->n_slot = 0;
        }
        __syncthreads();
        const long long total = s_total;
        if (total <= 0) return;
        if (tid < 2 * n_q + 1) {
            long long rank = total - 1;
            if (tid < 2 * n_q) {
                const double q = wq.q[tid >> 1];
                // NumPy's virtual index (n - 1) q in the dtype it uses for q (float64, or float32 for float32 data and a Python
                // scalar q under NumPy 2); a float32 h can round above n - 1: NumPy takes the last element
                const double h = q_is_f32 ? (double)__fmul_rn((float)(total - 1), (float)q) : __dmul_rn((double)(total - 1), q);
                long long lo = (long long)floor(h);
                lo = lo > total - 1 ? total - 1 : (lo < 0 ? 0 : lo);
                const long long hi = lo + 1 > total - 1 ? total - 1 : lo + 1;
                rank = (tid & 1) ? hi : lo;
            }
            if (reverse) rank = total - 1 - rank;
            st->rank[tid] = rank;
            st->residual[tid] = rank;
            st->prefix[tid] = 0u;
            st->slot[tid] = 0;
        }
        if (tid == 0) s_n_rank = 2 * n_q + 1;
    } else if (tid == 0) {
        s_n_rank = st->n_rank;
    }
    __syncthreads();
    const int n_rank = s_n_rank;
    const int bins = level == 2 ? 1024 : 2048, per = bins / 32;
    for (int r = warp; r < n_rank; r += kWalkThreads / 32) {
        const unsigned long long* h = hist + (size_t)st->slot[r] * 2048 + lane * per;
        const long long res = st->residual[r];
        unsigned long long c[64];
        long long mine = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            c[i] = i < per ? h[i] : 0ull;
            mine += (long long)c[i];
        }
        long long incl = mine;  // inclusive scan over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += up;
        }
        const long long before = incl - mine;
        // the first lane whose inclusive count exceeds the residual owns the digit (the last lane if none does: the residual
        // is below the slot's total by construction)
        const unsigned hit = __ballot_sync(kFull, incl > res);
        const int owner = hit ? __ffs(hit) - 1 : 31;
        if (lane == owner) {
            long long cum = before;
            int d = 0;
            for (; d < per - 1; ++d) {
                if (cum + (long long)c[d] > res) break;
                cum += (long long)c[d];
            }
            const unsigned digit = (unsigned)(lane * per + d);
            const unsigned prefix = level == 0 ? digit : (st->prefix[r] << (level == 1 ? 11 : 10)) | digit;
            st->residual[r] = res - cum;
            st->prefix[r] = prefix;
            s_prefix[r] = prefix;
            if (level == 2) st->key[r] = prefix;
        }
    }
    __syncthreads();
    if (level < 2 && tid < n_rank) {
        // ranks with one prefix share a histogram slot: the slot of a rank is the number of distinct prefixes among the ranks
        // before its first occurrence
        const int r = tid;
        int first = r;
        for (int j = 0; j < r; ++j)
            if (s_prefix[j] == s_prefix[r]) { first = j; break; }
        int slot = 0;
        for (int j = 0; j < first; ++j) {
            bool fresh = true;
            for (int k = 0; k < j; ++k)
                if (s_prefix[k] == s_prefix[j]) { fresh = false; break; }
            slot += fresh;
        }
        st->slot[r] = slot;
        if (first == r) st->slot_prefix[slot] = s_prefix[r];
    }
    if (level < 2 && tid == 0) {
        int n_slot = 0;
        for (int j = 0; j < n_rank; ++j) {
            bool fresh = true;
            for (int k = 0; k < j; ++k)
                if (s_prefix[k] == s_prefix[j]) { fresh = false; break; }
            n_slot += fresh;
        }
        st->n_slot = n_slot;
    }
}

int launch_radix_walk(const unsigned long long* hist, int level, const double* q_host, int n_q, int q_is_f32, int reverse,
                      vu_radix_state* state, cudaStream_t stream, int n_seg, int seg_B) {
    WalkQ wq;
    for (int i = 0; i < 32; ++i) wq.q[i] = (q_host && i < n_q) ? q_host[i] : 0.0;
    radix_walk_kernel<<<n_seg > 0 ? n_seg : 1, kWalkThreads, 0, stream>>>(hist, level, wq, n_q, q_is_f32, reverse, state, n_seg > 0 ? seg_B : 0);
    count_launch("radix_walk");
    return check_launch("radix_walk");
}

int launch_radix_hist(const float* values, long long n, const GtView& gt, int level, const unsigned* prefixes, int n_prefix,
                      unsigned long long* hist, cudaStream_t stream, const vu_radix_state* state) {
    long long blocks = (n + kRadixThreads - 1) / kRadixThreads;
    const long long cap = (long long)device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    SegBatch seg;
    memset(&seg, 0, sizeof(seg));
    radix_hist_kernel<<<(unsigned)blocks, kRadixThreads, 0, stream>>>(values, n, gt, level, prefixes, n_prefix, hist, state, seg);
    count_launch("radix_hist");
    return check_launch("radix_hist");
}

// one histogram pass over every segment of a batch: grid.y = segment, grid.x shares 8 CTAs per SM among the segments
int launch_radix_hist_batch(const float* const* maps, int n_maps, long long B, long long V, const GtView& gt, int level,
                            unsigned long long* hist, cudaStream_t stream, const vu_radix_state* states) {
    SegBatch seg;
    memset(&seg, 0, sizeof(seg));
    seg.n_maps = n_maps; seg.B = (int)B;
    for (int m = 0; m < n_maps; ++m) seg.maps[m] = maps[m];
    const long long n_seg = (long long)n_maps * B;
    long long blocks = (V + kRadixThreads - 1) / kRadixThreads;
    long long cap = (long long)device_sm_count() * 8 / n_seg;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    radix_hist_kernel<<<dim3((unsigned)blocks, (unsigned)n_seg), kRadixThreads, 0, stream>>>(nullptr, V, gt, level, nullptr, 0, hist, states, seg);
    count_launch("radix_hist");
    return check_launch("radix_hist");
}

// ---- calibration histogram on caller-given thresholds -----------------------------------------------------------
struct BinnedParams {
    const float* map;
    const uint8_t* labels;
    long long V;
    GtView gt;
    CalibDev cal;  // edge[k]: thresholds on u (sign-adjusted), NaN = never reached
    const uint8_t* lut;
    unsigned long long* counts;  // [2][21]: samples, correct samples
    double* sums;                // [21]
    SegBatch seg;                // batched: per segment its map, image, thresholds (cals[s]) and output rows
    const vu_calib* cals;        // DEVICE array, one per segment
};

constexpr int kBinnedThreads = 256;

__global__ void __launch_bounds__(kBinnedThreads) binned_calib_kernel(const __grid_constant__ BinnedParams prm) {
    constexpr int WARPS = kBinnedThreads / 32;
    __shared__ double s_sum[WARPS][VU_N_BINS];
    __shared__ int s_tot[WARPS][VU_N_BINS], s_tru[WARPS][VU_N_BINS];
    __shared__ float thr[VU_N_EDGES];
    for (int t = threadIdx.x; t < WARPS * VU_N_BINS; t += kBinnedThreads) {
        (&s_sum[0][0])[t] = 0.0; (&s_tot[0][0])[t] = 0; (&s_tru[0][0])[t] = 0;
    }
    const float* map = prm.map;
    const uint8_t* labels = prm.labels;
    GtView gt = prm.gt;
    unsigned long long* counts = prm.counts;
    double* sums = prm.sums;
    bool increasing = prm.cal.increasing, identity = prm.cal.identity;
    float ca = prm.cal.a, cb = prm.cal.b;
    if (prm.seg.n_maps) {
        const int sgm = blockIdx.y, m = sgm / prm.seg.B, b = sgm - m * prm.seg.B;
        map = prm.seg.maps[m] + (long long)b * prm.V;
        if (labels) labels += (long long)b * prm.V;
        seg_gt(gt, b);
        counts += (size_t)sgm * 2 * VU_N_BINS;
        sums += (size_t)sgm * VU_N_BINS;
        const vu_calib& c = prm.cals[sgm];
        increasing = c.mode != VU_CALIB_PLATT_DEC;
        identity = c.mode == VU_CALIB_IDENTITY;
        ca = c.a; cb = c.b;
        if (threadIdx.x < VU_N_EDGES) thr[threadIdx.x] = increasing ? c.edge_u[threadIdx.x] : -c.edge_u[threadIdx.x];
    } else if (threadIdx.x < VU_N_EDGES) {
        thr[threadIdx.x] = prm.cal.edge[threadIdx.x];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * kBinnedThreads;
    for (long long base = (long long)blockIdx.x * kBinnedThreads; base < prm.V; base += stride) {
        const long long i = base + threadIdx.x;
        int bin = -1, nv = 0, nc = 0;
        double w = 0.0;
        if (i < prm.V) {
            const int label = labels ? (int)__ldg(labels + i) : 0;
            const long long cmp = prm.lut ? (long long)__ldg(prm.lut + label) : (long long)label;
            for (int r = 0; r < gt.R; ++r) {
                const long long off = (long long)r * gt.sr + i * gt.sv;
                const long long g = gt.dtype == VU_GT_U8 ? (long long)__ldg(reinterpret_cast<const uint8_t*>(gt.data) + off)
                                                         : __ldg(reinterpret_cast<const long long*>(gt.data) + off);
                const bool valid = !(gt.has_ignore && g == gt.ignore);
                nv += valid;
                nc += valid && g == cmp;
            }
            if (nv > 0) {
                const float u = __ldg(map + i);
                if (u != u) {
                    bin = VU_N_BINS - 1;  // NaN: past the last edge
                    w = (double)u;
                } else {
                    const float uu = increasing ? u : -u;
                    bin = 0;
#pragma unroll
                    for (int k = 0; k < VU_N_EDGES; ++k) bin += (uu >= thr[k]) ? 1 : 0;  // NaN thresholds compare false
                    // ace.py:329 in float32 (the bins do not depend on it, only the float64 sums do)
                    float conf;
                    if (identity) conf = fminf(fmaxf(u, 0.0f), 1.0f);
                    else conf = fminf(fmaxf(__fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fadd_rn(__fmul_rn(-u, ca), cb)))), 0.0f), 1.0f);
                    w = (double)conf * (double)nv;
                }
            }
        }
        // one round per distinct bin present in the warp
        unsigned todo = __ballot_sync(kFull, bin >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int bsel = __shfl_sync(kFull, bin, leader);
            const bool mine = (bin == bsel);
            const int tot = __reduce_add_sync(kFull, mine ? nv : 0);
            const int tru = __reduce_add_sync(kFull, mine ? nc : 0);
            const double sw = warp_sum(mine ? w : 0.0);
            if (lane == 0) { s_tot[warp][bsel] += tot; s_tru[warp][bsel] += tru; s_sum[warp][bsel] += sw; }
            todo &= ~__ballot_sync(kFull, mine);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < VU_N_BINS; t += kBinnedThreads) {
        long long tot = 0, tru = 0;
        double s = 0.0;
        for (int w = 0; w < WARPS; ++w) { tot += s_tot[w][t]; tru += s_tru[w][t]; s += s_sum[w][t]; }
        if (tot) {
            atomicAdd(counts + t, (unsigned long long)tot);
            if (tru) atomicAdd(counts + VU_N_BINS + t, (unsigned long long)tru);
            atomicAdd(sums + t, s);
        }
    }
}

int launch_binned_calib(const float* map, const uint8_t* labels, long long V, const GtView& gt, const CalibDev& cal, const uint8_t* lut,
                        unsigned long long* counts, double* sums, cudaStream_t stream) {
    BinnedParams prm;
    prm.map = map; prm.labels = labels; prm.V = V; prm.gt = gt; prm.cal = cal; prm.lut = lut; prm.counts = counts; prm.sums = sums;
    memset(&prm.seg, 0, sizeof(prm.seg));
    prm.cals = nullptr;
    long long blocks = (V + kBinnedThreads - 1) / kBinnedThreads;
    const long long cap = (long long)device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    binned_calib_kernel<<<(unsigned)blocks, kBinnedThreads, 0, stream>>>(prm);
    count_launch("binned_calib");
    return check_launch("binned_calib");
}

int launch_binned_calib_batch(const float* const* maps, int n_maps, long long B, long long V, const uint8_t* labels, const GtView& gt,
                              const vu_calib* cals_dev, const uint8_t* lut, unsigned long long* counts, double* sums, cudaStream_t stream) {
    BinnedParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.labels = labels; prm.V = V; prm.gt = gt; prm.lut = lut; prm.counts = counts; prm.sums = sums; prm.cals = cals_dev;
    prm.seg.n_maps = n_maps; prm.seg.B = (int)B;
    for (int m = 0; m < n_maps; ++m) prm.seg.maps[m] = maps[m];
    const long long n_seg = (long long)n_maps * B;
    long long blocks = (V + kBinnedThreads - 1) / kBinnedThreads;
    long long cap = (long long)device_sm_count() * 8 / n_seg;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    binned_calib_kernel<<<dim3((unsigned)blocks, (unsigned)n_seg), kBinnedThreads, 0, stream>>>(prm);
    count_launch("binned_calib");
    return check_launch("binned_calib");
}

}  // namespace vu
