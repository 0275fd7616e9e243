// K2: spatial reductions on maps that K1 just wrote (L2-resident):
//   patch_level_aggregation   evaluation/uncertainty_aggregation/aggregate_uncertainties.py:16-34
//   _compute_border           evaluation/uncertainty_aggregation/prediction_shape_stats.py:15-30
#include "vu_common.cuh"
#include "vu_host.h"

namespace vu {

// order-preserving double <-> uint64 so atomicMax works on doubles
__device__ __forceinline__ unsigned long long d2o(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double o2d(unsigned long long o) {
    unsigned long long u = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
    return __longlong_as_double((long long)u);
}

__global__ void patch_init(unsigned long long* max_enc, long long* first, long long B) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < B) { max_enc[i] = 0ull; first[i] = 0x7fffffffffffffffLL; }  // 0 encodes below -inf / NaN
}
__global__ void patch_finish(unsigned long long* max_enc, long long B, double scale) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < B) reinterpret_cast<double*>(max_enc)[i] = o2d(max_enc[i]) * scale;
}

// Fallback for boxes too large for patch_box3 (below).
// "valid" box sum of a (d0, d1, d2) map with a (k0, k1, k2) box, exact in
// float64.  A CTA owns a T1 x T2 window of output (y, x) positions and marches
// along z: per input slice it forms the x-sums, then the y-sums, and keeps the
// last k0 slice values in a shared ring to form the z running sum.
// MODE 0: per-image max (atomicMax on the ordered encoding).
// MODE 1: first row-major index with np.isclose(sum, max) (atomicMin).
constexpr int kT1 = 16, kT2 = 32;

template <int MODE>
__global__ void __launch_bounds__(kT1* kT2) patch_box(const float* __restrict__ maps, long long d0, long long d1, long long d2,
                                                      int k0, int k1, int k2, unsigned long long* max_enc,
                                                      long long* first, double scale, int tiles_y, int tiles_x) {
    extern __shared__ double smem_d[];
    const int in_h = kT1 + k1 - 1, in_w = kT2 + k2 - 1;
    double* xs = smem_d;                                   // [in_h][kT2]  x-sums of the current slice
    double* ring = xs + (size_t)in_h * kT2;                // [k0][kT1*kT2] y/x-sums of the last k0 slices
    float* tile = reinterpret_cast<float*>(ring + (size_t)k0 * kT1 * kT2);  // [in_h][in_w]

    const long long b = blockIdx.y;
    const int ty0 = (blockIdx.x / tiles_x) * kT1, tx0 = (blockIdx.x % tiles_x) * kT2;
    const long long o0 = d0 - k0 + 1, o1 = d1 - k1 + 1, o2 = d2 - k2 + 1;
    const int tid = threadIdx.x, ly = tid / kT2, lx = tid % kT2;
    const long long oy = ty0 + ly, ox = tx0 + lx;
    const bool out_ok = oy < o1 && ox < o2;
    const float* img = maps + b * d0 * d1 * d2;

    double run = 0.0, best = 0.0;
    bool have = false;
    long long best_idx = 0x7fffffffffffffffLL;
    double peak = 0.0, tol = 0.0;
    if (MODE == 1) {
        peak = o2d(max_enc[b]);
        tol = 1e-8 / scale + 1e-5 * fabs(peak);  // np.isclose(v*scale, peak*scale): atol 1e-8, rtol 1e-5
    }
    for (long long z = 0; z < d0; ++z) {
        for (int i = tid; i < in_h * in_w; i += kT1 * kT2) {
            const int yy = i / in_w, xx = i % in_w;
            const long long gy = ty0 + yy, gx = tx0 + xx;
            tile[i] = (gy < d1 && gx < d2) ? __ldg(img + (z * d1 + gy) * d2 + gx) : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < in_h * kT2; i += kT1 * kT2) {
            const int yy = i / kT2, xx = i % kT2;
            double s = 0.0;
            for (int dx = 0; dx < k2; ++dx) s += (double)tile[yy * in_w + xx + dx];
            xs[i] = s;
        }
        __syncthreads();
        double s = 0.0;
        for (int dy = 0; dy < k1; ++dy) s += xs[(ly + dy) * kT2 + lx];
        double* slot = ring + (size_t)(z % k0) * (kT1 * kT2) + tid;
        if (z >= k0) run -= *slot;
        *slot = s;
        run += s;
        if (z >= k0 - 1) {
            // recompute from the ring every k0-th step so add/subtract drift cannot build up
            if (((z - (k0 - 1)) % k0) == 0) {
                double r = 0.0;
                for (int dz = 0; dz < k0; ++dz) r += ring[(size_t)dz * (kT1 * kT2) + tid];
                run = r;
            }
            if (out_ok) {
                const long long oz = z - (k0 - 1);
                if (MODE == 0) {
                    if (!have || run > best) { best = run; have = true; }
                } else if (fabs(run - peak) <= tol) {
                    const long long idx = (oz * o1 + oy) * o2 + ox;
                    if (idx < best_idx) best_idx = idx;
                }
            }
        }
        __syncthreads();
    }
    (void)o0;
    if (MODE == 0) {
        unsigned long long e = have ? d2o(best) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(kFull, e, o);
            e = other > e ? other : e;
        }
        if ((tid & 31) == 0 && e) atomicMax(max_enc + b, e);
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long other = __shfl_xor_sync(kFull, best_idx, o);
            best_idx = other < best_idx ? other : best_idx;
        }
        if ((tid & 31) == 0 && best_idx != 0x7fffffffffffffffLL) atomicMin(first + b, best_idx);
    }
}


// ---- general 3-D / 2-D kernel: sliding sums along all three axes --------------------------------------------------
// A CTA owns a 32 x 32 window of output (y, x) positions and marches along z.
//   step 1  every thread keeps float64 running sums over the last k0 slices for the input columns it owns (add the
//           slice that enters the window, subtract the one that leaves; both were requested one iteration earlier, so
//           their latency hides behind steps 2 and 3) and drops them into shared memory;
//   step 2  x direction: a thread forms kP3Seg consecutive k2-wide sums of one row (first one directly, then slide);
//   step 3  y direction: the same down a column, then max (MODE 0) or first-isclose index (MODE 1).
// 2 barriers per slice.  Inputs are float32 and the accumulators float64, so the add / subtract steps are exact unless
// a window spans more than ~2^29 in magnitude; otherwise the error per step is 2^-53 of the running value (score
// tolerance: 1e-5).
// tile_max (optional workspace, kP3ZChunks + 1 words per CTA): MODE 0 stores the CTA's own maximum and the maxima of
// the chunks of kP3ZSlices output slices; MODE 1 returns at once when the CTA's maximum is not within the isclose
// tolerance of the image's (normally all but one CTA per image), otherwise starts at the first chunk that holds a
// candidate and stops after the first slice with a hit (later slices only have larger row-major indices).
// KT: box edge as a compile-time constant for cubic / square boxes (0: read k0, k1, k2 at run time).
constexpr int kP3T = 32, kP3Threads = 256, kP3Seg = 4;
constexpr int kP3MaxCols = 16;   // input columns per thread: (32 + k - 1)^2 <= 16 * 256  <=>  k <= 33
constexpr int kP3ZSlices = 8;    // output slices per chunk of the workspace
__host__ __device__ inline int p3_zchunks(long long o0) { return (int)((o0 + kP3ZSlices - 1) / kP3ZSlices); }

template <int MODE, int KT>
__global__ void __launch_bounds__(kP3Threads, KT > 0 ? 3 : 2) patch_box3(const float* __restrict__ maps, long long d0, long long d1, long long d2,
                                                           int rk0, int rk1, int rk2, unsigned long long* max_enc, long long* first,
                                                           double scale, int tiles_x, unsigned long long* tile_max) {
    extern __shared__ double smem_d[];
    const int k0 = (KT > 0 && rk0 != 1) ? KT : rk0, k1 = KT > 0 ? KT : rk1, k2 = KT > 0 ? KT : rk2;  // KT with rk0 == 1: a 2-D box
    constexpr int COLS = KT > 0 ? ((kP3T + KT - 1) * (kP3T + KT - 1) + kP3Threads - 1) / kP3Threads : kP3MaxCols;
    const long long b = blockIdx.y;
    const long long o0 = d0 - k0 + 1, o1 = d1 - k1 + 1, o2 = d2 - k2 + 1;
    const int nchunk = p3_zchunks(o0);
    unsigned long long* my_max = tile_max ? tile_max + (b * gridDim.x + blockIdx.x) * (nchunk + 1) : nullptr;
    double peak = 0.0, tol = 0.0;
    long long z_begin = 0;  // first input slice this CTA reads
    if (MODE == 1) {
        peak = o2d(max_enc[b]);
        tol = 1e-8 / scale + 1e-5 * fabs(peak);  // np.isclose(v*scale, peak*scale): atol 1e-8, rtol 1e-5
        if (my_max) {
            const unsigned long long e = my_max[0];
            if (e == 0ull || !(fabs(o2d(e) - peak) <= tol)) return;  // no candidate in this window (CTA-uniform)
            int ch = 0;
            while (ch < nchunk - 1 && !(my_max[1 + ch] != 0ull && fabs(o2d(my_max[1 + ch]) - peak) <= tol)) ++ch;
            z_begin = (long long)ch * kP3ZSlices;  // output slice ch * 8 needs input slices from the same index on
        }
    }
    const int in_h = kP3T + k1 - 1, in_w = kP3T + k2 - 1;
    double* zs = smem_d;                        // [in_h][in_w]  z-window sums of the current slice
    double* xs = zs + (size_t)in_h * in_w;      // [in_h][kP3T]  their x-sums
    const int ty0 = (blockIdx.x / tiles_x) * kP3T, tx0 = (blockIdx.x % tiles_x) * kP3T;
    const int tid = threadIdx.x;
    const long long plane = d1 * d2;
    const float* img = maps + b * d0 * plane;
    const int n_in = in_h * in_w;

    double run[COLS];
    int off[COLS];            // input column tid + i * threads of the window: offset inside a slice (0 when outside the map)
    float nv[COLS], ov[COLS];  // requested one iteration ahead: the column's element of the slice entering / leaving the window
    unsigned inside = 0;      // bit i: column i lies inside the map
#pragma unroll
    for (int i = 0; i < COLS; ++i) {
        const int c = tid + i * kP3Threads;
        run[i] = 0.0;
        off[i] = 0;
        ov[i] = 0.f;
        if (c < n_in) {
            const int yy = c / in_w, xx = c % in_w;
            const long long gy = ty0 + yy, gx = tx0 + xx;
            if (gy < d1 && gx < d2) { off[i] = (int)(gy * d2 + gx); inside |= 1u << i; }
        }
        const float v = __ldg(img + z_begin * plane + off[i]);
        nv[i] = ((inside >> i) & 1) ? v : 0.f;
    }
    double best = 0.0, chunk_best = 0.0;
    bool have = false, chunk_have = false;
    long long best_idx = 0x7fffffffffffffffLL;
    constexpr int kSegs = kP3T / kP3Seg;
    const int x3 = tid % kP3T, y3 = (tid / kP3T) * kP3Seg;  // step-3 item: column x3, outputs y3 .. y3+3 (32 x 8 items)
    const long long ox3 = tx0 + x3;
    __shared__ unsigned long long wmax[kP3Threads / 32];
    __shared__ int s_hit;
    if (MODE == 1 && tid == 0) s_hit = 0;

    for (long long z = z_begin; z < d0; ++z) {
#pragma unroll
        for (int i = 0; i < COLS; ++i) {
            const int c = tid + i * kP3Threads;
            run[i] += (double)nv[i];
            run[i] -= (double)ov[i];
            if (KT > 0 ? (i < COLS - 1 || c < n_in) : (c < n_in)) zs[c] = run[i];
        }
        if (z + 1 < d0) {
            const float* nxt = img + (z + 1) * plane;
            const bool leaving = z + 1 - z_begin >= k0;
            const float* old = leaving ? nxt - (long long)k0 * plane : nxt;
#pragma unroll
            for (int i = 0; i < COLS; ++i) {
                const float a = __ldg(nxt + off[i]), o = __ldg(old + off[i]);
                const bool in = (inside >> i) & 1;
                nv[i] = in ? a : 0.f;
                ov[i] = (in && leaving) ? o : 0.f;
            }
        }
        if (z - z_begin < k0 - 1) continue;  // uniform: the window is not full yet (zs is rewritten next slice by the same threads)
        __syncthreads();
        for (int it = tid; it < in_h * kSegs; it += kP3Threads) {
            const int r = it / kSegs, x0 = (it % kSegs) * kP3Seg;
            const double* row = zs + r * in_w + x0;
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < (KT > 0 ? KT : 1); ++j) s += row[j];
            if (KT == 0)
                for (int j = 1; j < k2; ++j) s += row[j];
            double* o = xs + r * kP3T + x0;
            o[0] = s;
#pragma unroll
            for (int j = 1; j < kP3Seg; ++j) {
                s += row[j + k2 - 1];
                s -= row[j - 1];
                o[j] = s;
            }
        }
        __syncthreads();
        const long long oz = z - (k0 - 1);
        {
            const double* col = xs + y3 * kP3T + x3;
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < (KT > 0 ? KT : 1); ++j) s += col[j * kP3T];
            if (KT == 0)
                for (int j = 1; j < k1; ++j) s += col[j * kP3T];
#pragma unroll
            for (int j = 0; j < kP3Seg; ++j) {
                if (j > 0) {
                    s += col[(j + k1 - 1) * kP3T];
                    s -= col[(j - 1) * kP3T];
                }
                const long long oy = ty0 + y3 + j;
                if (oy < o1 && ox3 < o2) {
                    if (MODE == 0) {
                        if (!chunk_have || s > chunk_best) { chunk_best = s; chunk_have = true; }
                    } else if (fabs(s - peak) <= tol) {
                        const long long idx = (oz * o1 + oy) * o2 + ox3;
                        if (idx < best_idx) { best_idx = idx; s_hit = 1; }
                    }
                }
            }
        }
        if (MODE == 0) {
            // end of a chunk of output slices: fold the chunk's maximum into the CTA's and (with a workspace) store it
            if ((oz + 1) % kP3ZSlices == 0 || oz + 1 == o0) {
                if (chunk_have && (!have || chunk_best > best)) { best = chunk_best; have = true; }
                if (my_max) {
                    unsigned long long e = chunk_have ? d2o(chunk_best) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        unsigned long long other = __shfl_xor_sync(kFull, e, o);
                        e = other > e ? other : e;
                    }
                    if ((tid & 31) == 0) wmax[tid >> 5] = e;
                    __syncthreads();
                    if (tid == 0) {
                        for (int w = 1; w < kP3Threads / 32; ++w) e = wmax[w] > e ? wmax[w] : e;
                        my_max[1 + oz / kP3ZSlices] = e;
                    }
                    // wmax is rewritten at the earliest after two more barriers (the next slice's steps 2 and 3)
                }
                chunk_have = false;
            }
        } else {
            // a hit in this slice: every later slice only has larger indices (the barrier also orders s_hit)
            __syncthreads();
            if (s_hit) break;
        }
        // otherwise no barrier here: the next slice's step 1 only writes zs (last read before the second barrier above),
        // and xs is rewritten only after the next slice's first barrier
    }
    if (MODE == 0) {
        unsigned long long e = have ? d2o(best) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(kFull, e, o);
            e = other > e ? other : e;
        }
        __syncthreads();
        if ((tid & 31) == 0) wmax[tid >> 5] = e;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kP3Threads / 32; ++w) e = wmax[w] > e ? wmax[w] : e;
            if (my_max) my_max[0] = e;
            if (e) atomicMax(max_enc + b, e);
        }
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long other = __shfl_xor_sync(kFull, best_idx, o);
            best_idx = other < best_idx ? other : best_idx;
        }
        if ((tid & 31) == 0 && best_idx != 0x7fffffffffffffffLL) atomicMin(first + b, best_idx);
    }
}

// ---- 3-D maps, cubic box of a compile-time edge K <= 10 (patch_size: 10, aggregation_all.yaml:9): plane kernel -------------
// patch_box3 spends its time in shared memory: per slice a 32 x 32 output window costs (32 + K - 1)^2 float64 stores and two
// sliding passes that read them back (~100 B of shared-memory traffic per output column and slice, 1.64 x redundant global
// loads for the halo; ncu r02s: 36 % issue, short-scoreboard and barrier stalls, and 512 CTAs on 444 slots = two waves).
// Here a CTA of 512 threads owns a 64 x 64 window of input columns (55 x 55 outputs for K = 10: a whole 64^3 LIDC crop per
// plane, 1.35 x halo instead of 1.64 x) and a chunk of output slices:
//   z  every thread keeps float64 running sums over the last K slices for 8 consecutive x columns of one row (two 128-bit
//      loads for the slice entering and the slice leaving, requested one slice ahead);
//   x  the K-wide sums along x slide over the thread's own 8 columns and the 9 that follow, which the next two lanes hold:
//      they come through warp shuffles -- no shared memory, no barrier;
//   y  the x-sums go to shared memory once (XOR-swizzled columns, row stride 65: stores and loads are conflict-free), one
//      barrier (double-buffered), and every thread slides down 8 outputs of one column (17 loads).
// ~40 B of shared-memory traffic per column and slice.  MODE and the workspace as in patch_box3: per CTA its own maximum and the
// maxima of its chunks of kPlChunk output slices; the index pass skips the CTAs that cannot hold a box np.isclose to the
// image's maximum, starts at the first chunk that holds one and stops after the first slice with a hit.
constexpr int kPlW = 64, kPlThreads = 512, kPlStride = 65, kPlRows = 65, kPlMaxZ = 4, kPlAhead = 4;
constexpr int kPlChunk = 4;  // output slices per workspace word (the index pass starts at the first chunk that holds a candidate)
__host__ __device__ inline long long plane_item_words(long long o0, int nz) { return 1 + ((o0 + nz - 1) / nz + kPlChunk - 1) / kPlChunk; }
constexpr size_t kPlSmem = 2 * (size_t)kPlRows * kPlStride * sizeof(double);

template <int MODE, int K>
__global__ void __launch_bounds__(kPlThreads, 2) patch_plane(const float* __restrict__ maps, long long d0, long long d1, long long d2,
                                                             unsigned long long* max_enc, long long* first, double scale, int tiles_y,
                                                             int tiles_x, int nz, unsigned long long* item_max) {
    static_assert(K >= 2 && K <= 10, "the x pass takes its halo from the next two lanes: K - 1 <= 9 columns");
    constexpr int OUT = kPlW - K + 1;
    extern __shared__ double smem_d[];
    const int tid = threadIdx.x;
    const long long b = blockIdx.y;
    int item = blockIdx.x;
    const int txi = item % tiles_x;
    item /= tiles_x;
    const int tyi = item % tiles_y, zc = item / tiles_y;
    const long long o0 = d0 - K + 1, o1 = d1 - K + 1, o2 = d2 - K + 1;
    const long long per = (o0 + nz - 1) / nz, oz_begin = zc * per, oz_end = oz_begin + per < o0 ? oz_begin + per : o0;
    unsigned long long* my_max = item_max ? item_max + (b * gridDim.x + blockIdx.x) * plane_item_words(o0, nz) : nullptr;
    if (oz_begin >= oz_end) {  // (more chunks than output slices)
        if (MODE == 0 && my_max && tid == 0) *my_max = 0ull;
        return;
    }
    double peak = 0.0, tol = 0.0;
    long long z_first = oz_begin;  // first input slice this CTA reads
    if (MODE == 1) {
        peak = o2d(max_enc[b]);
        tol = 1e-8 / scale + 1e-5 * fabs(peak);  // np.isclose(v*scale, peak*scale): atol 1e-8, rtol 1e-5
        if (my_max) {
            const unsigned long long e = *my_max;
            if (e == 0ull || !(fabs(o2d(e) - peak) <= tol)) return;  // no candidate in this window / chunk (CTA-uniform)
            const int nch = (int)((oz_end - oz_begin + kPlChunk - 1) / kPlChunk);
            int ch = 0;
            while (ch < nch - 1 && !(my_max[1 + ch] != 0ull && fabs(o2d(my_max[1 + ch]) - peak) <= tol)) ++ch;
            z_first += (long long)ch * kPlChunk;
        }
    }
    const long long ty0 = (long long)tyi * OUT, tx0 = (long long)txi * OUT;
    const long long plane = d1 * d2;
    const float* img = maps + b * d0 * plane;

    // x role: row ry of the window, columns xseg .. xseg + 7
    const int ry = tid >> 3, l8 = tid & 7, xseg = l8 * 8;
    const long long gy = ty0 + ry, gx0 = tx0 + xseg;
    const bool row_in = gy < d1;
    const float* src = img + (row_in ? gy * d2 : 0) + gx0;  // (only dereferenced where it lies inside the map)
    const bool fast = row_in && gx0 + 8 <= d2 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (plane & 3) == 0;
    auto load8 = [&](long long z, float (&v)[8]) {
        const float* p = src + z * plane;
        if (fast) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p)), c = __ldg(reinterpret_cast<const float4*>(p) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (row_in && gx0 + i < d2) ? __ldg(p + i) : 0.f;
        }
    };
    // y role: column cx of the window, outputs yseg .. yseg + 7
    const int cx = tid & 63, yseg = (tid >> 6) * 8;
    const int cxs = cx ^ ((cx >> 3) & 7);
    const long long ox = tx0 + cx;
    const bool col_ok = cx < OUT && ox < o2 && yseg < OUT;
    int n_valid = 0;  // outputs yseg .. yseg + n_valid - 1 of this thread's column lie inside the window and the map
    if (col_ok) {
        long long lim = OUT - yseg < o1 - (ty0 + yseg) ? OUT - yseg : o1 - (ty0 + yseg);
        n_valid = lim < 0 ? 0 : (lim > 8 ? 8 : (int)lim);
    }

    double run[8];
    float nv[8], ov[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { run[i] = 0.0; ov[i] = 0.f; }
    const long long z_last = oz_end + K - 1;  // input slices [z_first, z_last)
    load8(z_first, nv);
    double best = 0.0, chunk_best = -INFINITY;
    bool have = false, chunk_have = false;
    long long best_idx = 0x7fffffffffffffffLL;
    __shared__ unsigned long long wmax[kPlThreads / 32];

    for (long long z = z_first; z < z_last; ++z) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            run[i] += (double)nv[i];
            run[i] -= (double)ov[i];
        }
        if (z + 1 < z_last) {
            load8(z + 1, nv);
            if (z + 1 - z_first >= K) load8(z + 1 - K, ov);
        }
        // the slice kPlAhead further on goes to L2 now (the leaving slice was read K slices ago and is still there): the register
        // loads above, issued one slice ahead, then only have L2 latency to hide -- every warp of the CTA waits at the same time
        if (row_in && z + kPlAhead < z_last && l8 < 2 && tx0 + 32 * l8 < d2)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (z + kPlAhead) * plane + (l8 ? 32 - xseg : 0)));
        if (z - z_first < K - 1) continue;  // uniform: the window is not full yet
        const int buf = (int)(z & 1);
        {
            // K-wide sums along x: own columns 0..7, the next lane's 8..15, one more from the lane after it (lanes at the end of
            // a row pick up another row's values: they only reach outputs beyond the window, which are masked out below)
            double r[17];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                r[i] = run[i];
                r[8 + i] = __shfl_down_sync(kFull, run[i], 1);
            }
            r[16] = __shfl_down_sync(kFull, run[0], 2);
            double* xw = smem_d + (size_t)buf * (kPlRows * kPlStride) + ry * kPlStride + xseg;
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) s += r[j];
            xw[0 ^ l8] = s;
#pragma unroll
            for (int i = 1; i < 8; ++i) {
                s += r[i + K - 1];
                s -= r[i - 1];
                xw[i ^ l8] = s;
            }
        }
        __syncthreads();
        bool hit = false;
        if (yseg < OUT) {  // (uniform per pair of warps; the last segment holds no valid output)
            const double* col = smem_d + (size_t)buf * (kPlRows * kPlStride) + yseg * kPlStride + cxs;
            const long long oz = z - (K - 1);
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) s += col[j * kPlStride];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i > 0) {
                    s += col[(i + K - 1) * kPlStride];  // (row 64 is never written: it only reaches the masked output 55)
                    s -= col[(i - 1) * kPlStride];
                }
                if (i < n_valid) {
                    if (MODE == 0) {
                        chunk_best = fmax(chunk_best, s);  // (one DMNMX; a NaN sum is passed over, as `s > best` did)
                    } else if (fabs(s - peak) <= tol) {
                        const long long idx = (oz * o1 + ty0 + yseg + i) * o2 + ox;
                        if (idx < best_idx) best_idx = idx;
                        hit = true;
                    }
                }
            }
            if (MODE == 0 && n_valid > 0) chunk_have = true;
        }
        // MODE 1: a hit in this slice ends the search, every later slice only has larger row-major indices.  MODE 0 needs no
        // second barrier: the next slice writes the other buffer, whose readers passed this slice's barrier already.
        if (MODE == 1 && __syncthreads_or(hit)) break;
        if (MODE == 0) {
            const int done = (int)(z - (K - 1) + 1 - oz_begin);  // output slices of this CTA so far
            if (done % kPlChunk == 0 || z + 1 == z_last) {      // end of a chunk: fold it into the CTA's maximum, store it
                if (chunk_have && (!have || chunk_best > best)) { best = chunk_best; have = true; }
                if (my_max) {
                    unsigned long long e = chunk_have ? d2o(chunk_best) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        unsigned long long other = __shfl_xor_sync(kFull, e, o);
                        e = other > e ? other : e;
                    }
                    if ((tid & 31) == 0) wmax[tid >> 5] = e;
                    __syncthreads();
                    if (tid == 0) {
                        for (int w = 1; w < kPlThreads / 32; ++w) e = wmax[w] > e ? wmax[w] : e;
                        my_max[1 + (done - 1) / kPlChunk] = e;
                    }
                    // (wmax is rewritten after the next slice's barrier at the earliest)
                }
                chunk_have = false;
                chunk_best = -INFINITY;
            }
        }
    }
    if (MODE == 0) {
        unsigned long long e = have ? d2o(best) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(kFull, e, o);
            e = other > e ? other : e;
        }
        __syncthreads();
        if ((tid & 31) == 0) wmax[tid >> 5] = e;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kPlThreads / 32; ++w) e = wmax[w] > e ? wmax[w] : e;
            if (my_max) *my_max = e;
            if (e) atomicMax(max_enc + b, e);
        }
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long other = __shfl_xor_sync(kFull, best_idx, o);
            best_idx = other < best_idx ? other : best_idx;
        }
        if ((tid & 31) == 0 && best_idx != 0x7fffffffffffffffLL) atomicMin(first + b, best_idx);
    }
}

// ---- 2-D maps (d0 == 1): one thread per output column, marching down the rows ---------------------------------
// A CTA owns kTX output columns and a chunk of output rows.  Per input row: the row segment goes to shared memory
// (double-buffered, one barrier per row), each thread forms its K-tap x-sum in float64 and keeps the last K of
// them in registers (the row loop is unrolled by K, so the ring is indexed statically); the box sum of an output
// row is re-formed from the ring every time, so no add/subtract drift can build up.  MODE as in patch_box.
constexpr int kTX = 256, kRowsPerCta = 32;  // 32 output rows per CTA: enough CTAs to hide the row loads of one another

template <int K, int MODE>
__global__ void __launch_bounds__(kTX) patch_box2d(const float* __restrict__ maps, long long H, long long W, unsigned long long* max_enc,
                                                  long long* first, double scale, unsigned long long* tile_max) {
    __shared__ float row[2][kTX + K - 1];
    const long long b = blockIdx.z;
    const long long cta = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const long long o1 = H - K + 1, o2 = W - K + 1;
    const long long x0 = (long long)blockIdx.x * kTX, oy0 = (long long)blockIdx.y * kRowsPerCta;
    const long long oy1 = oy0 + kRowsPerCta < o1 ? oy0 + kRowsPerCta : o1;  // output rows [oy0, oy1)
    const int tx = threadIdx.x;
    const long long ox = x0 + tx;
    const bool col_ok = ox < o2;
    const float* img = maps + b * H * W;
    double ring[K];
#pragma unroll
    for (int j = 0; j < K; ++j) ring[j] = 0.0;
    double best = 0.0;
    bool have = false;
    long long best_idx = 0x7fffffffffffffffLL;
    double peak = 0.0, tol = 0.0;
    if (MODE == 1) {
        peak = o2d(max_enc[b]);
        tol = 1e-8 / scale + 1e-5 * fabs(peak);
        if (tile_max) {  // see patch_box3
            const unsigned long long e = tile_max[cta];
            if (e == 0ull || !(fabs(o2d(e) - peak) <= tol)) return;
        }
    }
    // rows travel global -> registers (requested two rows ahead) -> shared (one row ahead) -> the K-tap x-sum
    float pa = 0.f, pb = 0.f;
    auto fetch_row = [&](long long r) {
        const float* src = img + r * W + x0;
        pa = (x0 + tx < W) ? __ldg(src + tx) : 0.f;
        if (tx < K - 1) pb = (x0 + kTX + tx < W) ? __ldg(src + kTX + tx) : 0.f;
    };
    auto stage_row = [&](int buf) {
        row[buf][tx] = pa;
        if (tx < K - 1) row[buf][kTX + tx] = pb;
    };
    const long long r_end = oy1 + K - 1;  // input rows [oy0, r_end)
    fetch_row(oy0);
    stage_row(0);
    if (oy0 + 1 < r_end) fetch_row(oy0 + 1);
    __syncthreads();
    bool done = false;
    for (long long base = oy0; base < r_end && !done; base += K) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const long long r = base + j;
            if (r < r_end && !done) {  // uniform
                const int buf = (int)((r - oy0) & 1);
                if (r + 1 < r_end) stage_row(buf ^ 1);
                if (r + 2 < r_end) fetch_row(r + 2);
                double s = 0.0;
#pragma unroll
                for (int dx = 0; dx < K; ++dx) s += (double)row[buf][tx + dx];
                ring[j] = s;
                const long long oy = r - (K - 1);
                bool hit = false;
                if (oy >= oy0 && col_ok) {
                    double v = 0.0;
#pragma unroll
                    for (int i = 0; i < K; ++i) v += ring[i];
                    if (MODE == 0) {
                        if (!have || v > best) { best = v; have = true; }
                    } else if (fabs(v - peak) <= tol) {
                        const long long idx = oy * o2 + ox;
                        if (idx < best_idx) best_idx = idx;
                        hit = true;
                    }
                }
                // MODE 1: a hit in this row ends the search, every later row only has larger row-major indices
                if (MODE == 1) done = __syncthreads_or(hit) != 0;
                else __syncthreads();
            }
        }
    }
    if (MODE == 0) {
        unsigned long long e = have ? d2o(best) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(kFull, e, o);
            e = other > e ? other : e;
        }
        __shared__ unsigned long long wmax[kTX / 32];
        if ((tx & 31) == 0) wmax[tx >> 5] = e;
        __syncthreads();
        if (tx == 0) {
            for (int w = 1; w < kTX / 32; ++w) e = wmax[w] > e ? wmax[w] : e;
            if (tile_max) tile_max[cta] = e;
            if (e) atomicMax(max_enc + b, e);
        }
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long other = __shfl_xor_sync(kFull, best_idx, o);
            best_idx = other < best_idx ? other : best_idx;
        }
        if ((tx & 31) == 0 && best_idx != 0x7fffffffffffffffLL) atomicMin(first + b, best_idx);
    }
}

// ---- 2-D maps, square box of a compile-time edge K <= 10: strip kernel (r02) -----------------------------------------------
// patch_box2d above spends ~50 instructions per output (K shared-memory loads, K conversions and K float64 additions for the
// x sum of every output, K more for the ring).  Here a WARP is the unit of work: it owns a strip of 128 input columns (116
// outputs) and a chunk of kStripRows output rows, and needs no barrier at all:
//   x  a lane holds 4 consecutive columns of the row (one 128-bit load, requested two rows ahead); the K-wide sums slide over
//      its own values and the 9 that follow, which come from the next three lanes through shuffles (the last three lanes of a
//      warp produce no output: strips overlap by 12 columns);
//   y  the box sum of an output is a running sum over the last K x-sums: add the new row's, subtract the one that leaves -- the
//      x-sums of the last K rows wait in a lane-private ring in shared memory (128-bit stores / loads, conflict-free).
// ~18 instructions per output.  MODE / workspace as in patch_box2d (one word per warp task).
constexpr int kStripLanes = 29, kStripCols = kStripLanes * 4 /* 116 outputs per strip */, kStripRows = 32, kStripWarps = 8;

template <int K, int MODE>
__global__ void __launch_bounds__(kStripWarps * 32) patch_strip2d(const float* __restrict__ maps, long long H, long long W,
                                                                   unsigned long long* max_enc, long long* first, double scale,
                                                                   int strips, int chunks, long long tasks, unsigned long long* task_max) {
    static_assert(K >= 2 && K <= 10, "the halo comes from the next three lanes: K - 1 <= 9 columns after the lane's own four");
    extern __shared__ double smem_d[];  // [warp][K rows][128 columns]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long task = (long long)blockIdx.x * kStripWarps + warp;
    if (task >= tasks) return;
    const long long b = task / ((long long)strips * chunks);
    const int rem = (int)(task - b * strips * chunks), chunk = rem / strips, strip = rem - chunk * strips;
    const long long o1 = H - K + 1, o2 = W - K + 1;
    const long long x0 = (long long)strip * kStripCols, oy0 = (long long)chunk * kStripRows;
    const long long oy1 = oy0 + kStripRows < o1 ? oy0 + kStripRows : o1;
    double peak = 0.0, tol = 0.0;
    if (MODE == 1) {
        peak = o2d(max_enc[b]);
        tol = 1e-8 / scale + 1e-5 * fabs(peak);
        if (task_max) {
            const unsigned long long e = task_max[task];
            if (e == 0ull || !(fabs(o2d(e) - peak) <= tol)) return;  // (warp-uniform)
        }
    }
    const float* img = maps + b * H * W;
    const long long gx = x0 + 4 * lane;
    const float* src = img + gx;
    const bool fast = gx + 4 <= W && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (W & 3) == 0;
    auto load4 = [&](long long r, float (&v)[4]) {
        const float* p = src + r * W;
        if (fast) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = gx + i < W ? __ldg(p + i) : 0.f;
        }
    };
    double* ring = smem_d + ((size_t)warp * K * 128 + 4 * lane);  // row j of the ring: ring + j * 128
    const int n_out = lane < kStripLanes ? (o2 - gx >= 4 ? 4 : (o2 - gx > 0 ? (int)(o2 - gx) : 0)) : 0;  // valid outputs of this lane
    double run[4] = {0.0, 0.0, 0.0, 0.0};
    double best = -INFINITY;
    long long best_idx = 0x7fffffffffffffffLL;
    const long long r_end = oy1 + K - 1;  // input rows [oy0, r_end)
    float va[4], vb[4];
    load4(oy0, va);
    if (oy0 + 1 < r_end) load4(oy0 + 1, vb);
    int slot = 0;  // ring row the current input row goes to (= the row that leaves the window)
    for (long long r = oy0; r < r_end; ++r) {
        double d[13];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = (double)va[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) { va[i] = vb[i]; }
        if (r + 2 < r_end) load4(r + 2, vb);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            d[4 + i] = __shfl_down_sync(kFull, d[i], 1);
            d[8 + i] = __shfl_down_sync(kFull, d[i], 2);
        }
        d[12] = __shfl_down_sync(kFull, d[0], 3);
        double xs[4];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) s += d[j];
        xs[0] = s;
#pragma unroll
        for (int i = 1; i < 4; ++i) {
            s += d[i + K - 1];
            s -= d[i - 1];
            xs[i] = s;
        }
        double* rr = ring + slot * 128;
        if (r - oy0 >= K) {  // (warp-uniform) the x-sums of the row that leaves the window
            const double2 o01 = *reinterpret_cast<const double2*>(rr), o23 = *reinterpret_cast<const double2*>(rr + 2);
            run[0] -= o01.x; run[1] -= o01.y; run[2] -= o23.x; run[3] -= o23.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) run[i] += xs[i];
        *reinterpret_cast<double2*>(rr) = make_double2(xs[0], xs[1]);
        *reinterpret_cast<double2*>(rr + 2) = make_double2(xs[2], xs[3]);
        slot = slot + 1 == K ? 0 : slot + 1;
        const long long oy = r - (K - 1);
        if (oy >= oy0) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < n_out) best = fmax(best, run[i]);
            } else {
                bool hit = false;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < n_out && fabs(run[i] - peak) <= tol) {
                        const long long idx = oy * o2 + gx + i;
                        if (idx < best_idx) best_idx = idx;
                        hit = true;
                    }
                if (__any_sync(kFull, hit)) break;  // every later row only has larger row-major indices
            }
        }
    }
    if (MODE == 0) {
        unsigned long long e = (n_out > 0 && oy1 > oy0) ? d2o(best) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(kFull, e, o);
            e = other > e ? other : e;
        }
        if (lane == 0) {
            if (task_max) task_max[task] = e;
            if (e) atomicMax(max_enc + b, e);
        }
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long other = __shfl_xor_sync(kFull, best_idx, o);
            best_idx = other < best_idx ? other : best_idx;
        }
        if (lane == 0 && best_idx != 0x7fffffffffffffffLL) atomicMin(first + b, best_idx);
    }
}
static long long strip_tasks_per_image(long long H, long long W, int K) {
    const long long o1 = H - K + 1, o2 = W - K + 1;
    return ((o2 + kStripCols - 1) / kStripCols) * ((o1 + kStripRows - 1) / kStripRows);
}
template <int K>
static int launch_strip2d(const float* maps, long long B, long long H, long long W, unsigned long long* enc, long long* first,
                          double scale, unsigned long long* task_max, cudaStream_t stream) {
    const long long o1 = H - K + 1, o2 = W - K + 1;
    const int strips = (int)((o2 + kStripCols - 1) / kStripCols), chunks = (int)((o1 + kStripRows - 1) / kStripRows);
    const long long tasks = (long long)strips * chunks * B;
    const size_t smem = (size_t)kStripWarps * K * 128 * sizeof(double);
    if (cudaFuncSetAttribute(patch_strip2d<K, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(patch_strip2d<K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(patch_strip2d)");
    const unsigned grid = (unsigned)((tasks + kStripWarps - 1) / kStripWarps);
    patch_strip2d<K, 0><<<grid, kStripWarps * 32, smem, stream>>>(maps, H, W, enc, first, scale, strips, chunks, tasks, task_max);
    patch_strip2d<K, 1><<<grid, kStripWarps * 32, smem, stream>>>(maps, H, W, enc, first, scale, strips, chunks, tasks, task_max);
    return VU_OK;
}

template <int K>
static void launch_patch2d(const float* maps, long long B, long long H, long long W, unsigned long long* enc, long long* first,
                           double scale, unsigned long long* tile_max, cudaStream_t stream) {
    const long long o1 = H - K + 1, o2 = W - K + 1;
    dim3 grid((unsigned)((o2 + kTX - 1) / kTX), (unsigned)((o1 + kRowsPerCta - 1) / kRowsPerCta), (unsigned)B);
    patch_box2d<K, 0><<<grid, kTX, 0, stream>>>(maps, H, W, enc, first, scale, tile_max);
    patch_box2d<K, 1><<<grid, kTX, 0, stream>>>(maps, H, W, enc, first, scale, tile_max);
}

static bool patch_uses_2d(long long d0, long long d1, int k0, int k1, int k2) {
    return d0 == 1 && k0 == 1 && k1 == k2 && (k1 == 10 || k1 == 4 || k1 == 16) && (d1 - k1 + 1 + kRowsPerCta - 1) / kRowsPerCta <= 65535;
}
static bool patch_uses_box3(int k1, int k2) { return (kP3T + k1 - 1) * (kP3T + k2 - 1) <= kP3MaxCols * kP3Threads; }
static bool patch_uses_plane(long long d0, long long d1, long long d2, int k0, int k1, int k2) {
    return d0 > 1 && k0 == 10 && k1 == 10 && k2 == 10 && d1 * d2 <= 0x7fffffffLL && get_option("patch_path", 0) != 1;  // option 1: patch_box3 (A/B tests)
}
static long long plane_tiles(long long d, int k) { return (d - k + 1 + (kPlW - k + 1) - 1) / (kPlW - k + 1); }
// chunks of output slices per window: as many as shorten the launch (whole waves of CTAs x slices per CTA, warm-up included)
static int plane_zchunks(long long B, long long d0, long long d1, long long d2, int k) {
    const long long o0 = d0 - k + 1, tiles = plane_tiles(d1, k) * plane_tiles(d2, k), slots = 2LL * device_sm_count();
    int best_nz = 1;
    long long best_cost = -1;
    for (int nz = 1; nz <= kPlMaxZ; ++nz) {
        const long long per = (o0 + nz - 1) / nz;
        const long long cost = ((B * tiles * nz + slots - 1) / slots) * (per + k - 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nz = nz; }
    }
    return best_nz;
}

// CTAs per image of the kernel launch_patch_max picks (one workspace word each)
long long patch_ctas_per_image(long long d0, long long d1, long long d2, int k0, int k1, int k2) {
    const long long o1 = d1 - k1 + 1, o2 = d2 - k2 + 1;
    if (patch_uses_2d(d0, d1, k0, k1, k2)) {  // (sized for either 2-D kernel)
        const long long a = ((o2 + kTX - 1) / kTX) * ((o1 + kRowsPerCta - 1) / kRowsPerCta), c = strip_tasks_per_image(d1, d2, k1);
        return a > c ? a : c;
    }
    if (patch_uses_plane(d0, d1, d2, k0, k1, k2)) {  // (sized for either 3-D kernel: the option can change between the two calls)
        long long a = 0;
        for (int nz = 1; nz <= kPlMaxZ; ++nz) {
            const long long w = plane_tiles(d1, k1) * plane_tiles(d2, k2) * nz * plane_item_words(d0 - k0 + 1, nz);
            a = w > a ? w : a;
        }
        const long long c = ((o1 + kP3T - 1) / kP3T) * ((o2 + kP3T - 1) / kP3T) * (p3_zchunks(d0 - k0 + 1) + 1);
        return a > c ? a : c;
    }
    if (patch_uses_box3(k1, k2)) return ((o1 + kP3T - 1) / kP3T) * ((o2 + kP3T - 1) / kP3T) * (p3_zchunks(d0 - k0 + 1) + 1);
    return 0;  // the fallback kernel does not use the workspace
}

int launch_patch_max(const float* maps, long long B, long long d0, long long d1, long long d2, int k0, int k1, int k2,
                     int mean, double* out_max, long long* out_first, unsigned long long* tile_max, cudaStream_t stream) {
    const double scale = mean ? 1.0 / ((double)k0 * k1 * k2) : 1.0;
    const long long o1 = d1 - k1 + 1, o2 = d2 - k2 + 1;
    if (B > 65535) return set_error(VU_ERR_UNSUPPORTED, "B > 65535 per vu_patch_max call");
    unsigned long long* enc = reinterpret_cast<unsigned long long*>(out_max);
    const unsigned ib = (unsigned)((B + 255) / 256);
    patch_init<<<ib, 256, 0, stream>>>(enc, out_first, B);
    count_launch("patch_init");
    // 2-D maps with a square box of a specialised size: the column-marching kernel
    if (patch_uses_2d(d0, d1, k0, k1, k2)) {
        if ((k1 == 10 || k1 == 4) && get_option("patch_path", 0) != 1 && strip_tasks_per_image(d1, d2, k1) * B < (1LL << 31)) {
            const int rc = k1 == 10 ? launch_strip2d<10>(maps, B, d1, d2, enc, out_first, scale, tile_max, stream)
                                    : launch_strip2d<4>(maps, B, d1, d2, enc, out_first, scale, tile_max, stream);
            if (rc != VU_OK) return rc;
            count_launch("patch_strip2d"); count_launch("patch_strip2d");
            patch_finish<<<ib, 256, 0, stream>>>(enc, B, scale);
            count_launch("patch_finish");
            return check_launch("patch_max");
        }
        if (k1 == 10) launch_patch2d<10>(maps, B, d1, d2, enc, out_first, scale, tile_max, stream);
        else if (k1 == 4) launch_patch2d<4>(maps, B, d1, d2, enc, out_first, scale, tile_max, stream);
        else launch_patch2d<16>(maps, B, d1, d2, enc, out_first, scale, tile_max, stream);
        count_launch("patch_box2d"); count_launch("patch_box2d");
    } else if (patch_uses_plane(d0, d1, d2, k0, k1, k2)) {
        const int tiles_y = (int)plane_tiles(d1, k1), tiles_x = (int)plane_tiles(d2, k2);
        const int nz = plane_zchunks(B, d0, d1, d2, k0);
        if ((long long)tiles_y * tiles_x * nz > 0x7fffffffLL) return set_error(VU_ERR_UNSUPPORTED, "map too large for one vu_patch_max call");
        if (cudaFuncSetAttribute(patch_plane<0, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlSmem) != cudaSuccess ||
            cudaFuncSetAttribute(patch_plane<1, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlSmem) != cudaSuccess)
            return set_cuda_error("cudaFuncSetAttribute(patch_plane)");
        dim3 grid((unsigned)(tiles_y * tiles_x * nz), (unsigned)B);
        patch_plane<0, 10><<<grid, kPlThreads, kPlSmem, stream>>>(maps, d0, d1, d2, enc, out_first, scale, tiles_y, tiles_x, nz, tile_max);
        patch_plane<1, 10><<<grid, kPlThreads, kPlSmem, stream>>>(maps, d0, d1, d2, enc, out_first, scale, tiles_y, tiles_x, nz, tile_max);
        count_launch("patch_plane"); count_launch("patch_plane");
    } else if (patch_uses_box3(k1, k2)) {
        const int tiles_y = (int)((o1 + kP3T - 1) / kP3T), tiles_x = (int)((o2 + kP3T - 1) / kP3T);
        const int in_h = kP3T + k1 - 1, in_w = kP3T + k2 - 1;
        const size_t smem = ((size_t)in_h * in_w + (size_t)in_h * kP3T) * sizeof(double);
        if (smem + 2048 > 200 * 1024) return set_error(VU_ERR_UNSUPPORTED, "patch too large for shared memory");
        if (d1 * d2 > 0x7fffffffLL) return set_error(VU_ERR_UNSUPPORTED, "a slice of the map has more than 2^31 voxels");
        dim3 grid((unsigned)(tiles_y * tiles_x), (unsigned)B);
        const bool k10 = k1 == 10 && k2 == 10 && (k0 == 10 || k0 == 1);  // patch_size: 10 (aggregation_all.yaml:9), 3-D or 2-D
        auto k_max = k10 ? patch_box3<0, 10> : patch_box3<0, 0>;
        auto k_idx = k10 ? patch_box3<1, 10> : patch_box3<1, 0>;
        if (smem + 2048 > 48 * 1024) {  // static shared memory comes on top
            if (cudaFuncSetAttribute(k_max, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
                cudaFuncSetAttribute(k_idx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return set_cuda_error("cudaFuncSetAttribute(patch_box3)");
        }
        k_max<<<grid, kP3Threads, smem, stream>>>(maps, d0, d1, d2, k0, k1, k2, enc, out_first, scale, tiles_x, tile_max);
        k_idx<<<grid, kP3Threads, smem, stream>>>(maps, d0, d1, d2, k0, k1, k2, enc, out_first, scale, tiles_x, tile_max);
        count_launch("patch_box3"); count_launch("patch_box3");
    } else {
        const int tiles_y = (int)((o1 + kT1 - 1) / kT1), tiles_x = (int)((o2 + kT2 - 1) / kT2);
        const size_t smem = ((size_t)(kT1 + k1 - 1) * kT2 + (size_t)k0 * kT1 * kT2) * sizeof(double) +
                            (size_t)(kT1 + k1 - 1) * (kT2 + k2 - 1) * sizeof(float);
        if (smem > 200 * 1024) return set_error(VU_ERR_UNSUPPORTED, "patch too large for shared memory");
        // (the attribute is per device and this may be the first launch on this one: set it whenever it is needed)
        if (smem > 48 * 1024 &&
            (cudaFuncSetAttribute(patch_box<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
             cudaFuncSetAttribute(patch_box<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess))
            return set_cuda_error("cudaFuncSetAttribute(patch_box)");
        dim3 grid((unsigned)(tiles_y * tiles_x), (unsigned)B);
        patch_box<0><<<grid, kT1 * kT2, smem, stream>>>(maps, d0, d1, d2, k0, k1, k2, enc, out_first, scale, tiles_y, tiles_x);
        patch_box<1><<<grid, kT1 * kT2, smem, stream>>>(maps, d0, d1, d2, k0, k1, k2, enc, out_first, scale, tiles_y, tiles_x);
        count_launch("patch_box"); count_launch("patch_box");
    }
    patch_finish<<<ib, 256, 0, stream>>>(enc, B, scale);
    count_launch("patch_finish");
    return check_launch("patch_max");
}

// ---- border -----------------------------------------------------------------
__global__ void __launch_bounds__(256) border_kernel(const uint8_t* __restrict__ lab, long long d0, long long d1, long long d2,
                                                     long long* stats_i64) {
    const long long b = blockIdx.y;
    const long long V = d0 * d1 * d2;
    const uint8_t* L = lab + b * V;
    int cnt = 0;
    for (long long v = blockIdx.x * 256LL + threadIdx.x; v < V; v += (long long)gridDim.x * 256) {
        const long long x = v % d2, y = (v / d2) % d1, z = v / (d1 * d2);
        const uint8_t me = L[v];
        if (x + 1 < d2) cnt += (L[v + 1] != me);
        if (y + 1 < d1) cnt += (L[v + d2] != me);
        if (z + 1 < d0) cnt += (L[v + d1 * d2] != me);
    }
    cnt = warp_sum(cnt);
    __shared__ int part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w = 0; w < 8; ++w) s += part[w];
        if (s) atomicAdd(reinterpret_cast<unsigned long long*>(stats_i64 + b * VU_I64_COLS + VU_I64_BORDER), (unsigned long long)s);
    }
}

// Four voxels along x per thread (rows of a multiple of 4 labels, 4-byte aligned): the labels of a thread are one word, the
// differences to the right / lower / deeper neighbours are byte-wise zero tests of an XOR (see vu_common.cuh).
__global__ void __launch_bounds__(256) border4_kernel(const uint8_t* __restrict__ lab, long long d0, long long d1, long long d2,
                                                      long long* stats_i64) {
    const long long b = blockIdx.y;
    const long long V = d0 * d1 * d2;
    const uint8_t* L = lab + b * V;
    int cnt = 0;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < V / 4; i += (long long)gridDim.x * 256) {
        const long long v = 4 * i, row = v / d2, x = v - row * d2, z = row / d1, y = row - z * d1;
        const unsigned w = __ldg(reinterpret_cast<const unsigned*>(L + v));
        cnt += __popc(bytes_nonzero(w ^ (w >> 8)) & 0x00808080u);                       // (0,1) (1,2) (2,3) inside the word
        if (x + 4 < d2) cnt += ((unsigned)__ldg(L + v + 4) != (w >> 24));               // (3, first label of the next word)
        if (y + 1 < d1) cnt += __popc(bytes_nonzero(w ^ __ldg(reinterpret_cast<const unsigned*>(L + v + d2))));
        if (z + 1 < d0) cnt += __popc(bytes_nonzero(w ^ __ldg(reinterpret_cast<const unsigned*>(L + v + d1 * d2))));
    }
    cnt = warp_sum(cnt);
    __shared__ int part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w = 0; w < 8; ++w) s += part[w];
        if (s) atomicAdd(reinterpret_cast<unsigned long long*>(stats_i64 + b * VU_I64_COLS + VU_I64_BORDER), (unsigned long long)s);
    }
}

int launch_border(const uint8_t* labels, long long B, long long d0, long long d1, long long d2, long long* stats_i64,
                  cudaStream_t stream) {
    if (B > 65535) return set_error(VU_ERR_UNSUPPORTED, "B > 65535 per vu_border_count call");
    const long long V = d0 * d1 * d2;
    long long gx = (V + 256 * 8 - 1) / (256 * 8);
    if (gx < 1) gx = 1;
    if (gx > 4096) gx = 4096;
    if (d2 % 4 == 0 && ((uintptr_t)labels % 4) == 0) border4_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, stream>>>(labels, d0, d1, d2, stats_i64);
    else border_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, stream>>>(labels, d0, d1, d2, stats_i64);
    count_launch("border");
    return check_launch("border");
}

}  // namespace vu
