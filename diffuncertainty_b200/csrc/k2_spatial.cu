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

// ---- 2-D maps (d0 == 1): one thread per output column, marching down the rows ---------------------------------
// A CTA owns kTX output columns and a chunk of output rows.  Per input row: the row segment goes to shared memory
// (double-buffered, one barrier per row), each thread forms its K-tap x-sum in float64 and keeps the last K of
// them in registers (the row loop is unrolled by K, so the ring is indexed statically); the box sum of an output
// row is re-formed from the ring every time, so no add/subtract drift can build up.  MODE as in patch_box.
constexpr int kTX = 256, kRowsPerCta = 128;

template <int K, int MODE>
__global__ void __launch_bounds__(kTX) patch_box2d(const float* __restrict__ maps, long long H, long long W, unsigned long long* max_enc,
                                                  long long* first, double scale) {
    __shared__ float row[2][kTX + K - 1];
    const long long b = blockIdx.z;
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
    }
    auto load_row = [&](long long r, int buf) {
        const float* src = img + r * W + x0;
        row[buf][tx] = (x0 + tx < W) ? __ldg(src + tx) : 0.f;
        if (tx < K - 1) row[buf][kTX + tx] = (x0 + kTX + tx < W) ? __ldg(src + kTX + tx) : 0.f;
    };
    const long long r_end = oy1 + K - 1;  // input rows [oy0, r_end)
    load_row(oy0, 0);
    __syncthreads();
    for (long long base = oy0; base < r_end; base += K) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const long long r = base + j;
            if (r < r_end) {  // uniform
                const int buf = (int)((r - oy0) & 1);
                if (r + 1 < r_end) load_row(r + 1, buf ^ 1);
                double s = 0.0;
#pragma unroll
                for (int dx = 0; dx < K; ++dx) s += (double)row[buf][tx + dx];
                ring[j] = s;
                const long long oy = r - (K - 1);
                if (oy >= oy0 && col_ok) {
                    double v = 0.0;
#pragma unroll
                    for (int i = 0; i < K; ++i) v += ring[i];
                    if (MODE == 0) {
                        if (!have || v > best) { best = v; have = true; }
                    } else if (fabs(v - peak) <= tol) {
                        const long long idx = oy * o2 + ox;
                        if (idx < best_idx) best_idx = idx;
                    }
                }
                __syncthreads();
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
        if ((tx & 31) == 0 && e) atomicMax(max_enc + b, e);
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long other = __shfl_xor_sync(kFull, best_idx, o);
            best_idx = other < best_idx ? other : best_idx;
        }
        if ((tx & 31) == 0 && best_idx != 0x7fffffffffffffffLL) atomicMin(first + b, best_idx);
    }
}

template <int K>
static void launch_patch2d(const float* maps, long long B, long long H, long long W, unsigned long long* enc, long long* first,
                           double scale, cudaStream_t stream) {
    const long long o1 = H - K + 1, o2 = W - K + 1;
    dim3 grid((unsigned)((o2 + kTX - 1) / kTX), (unsigned)((o1 + kRowsPerCta - 1) / kRowsPerCta), (unsigned)B);
    patch_box2d<K, 0><<<grid, kTX, 0, stream>>>(maps, H, W, enc, first, scale);
    patch_box2d<K, 1><<<grid, kTX, 0, stream>>>(maps, H, W, enc, first, scale);
}

int launch_patch_max(const float* maps, long long B, long long d0, long long d1, long long d2, int k0, int k1, int k2,
                     int mean, double* out_max, long long* out_first, cudaStream_t stream) {
    const double scale = mean ? 1.0 / ((double)k0 * k1 * k2) : 1.0;
    const long long o1 = d1 - k1 + 1, o2 = d2 - k2 + 1;
    const int tiles_y = (int)((o1 + kT1 - 1) / kT1), tiles_x = (int)((o2 + kT2 - 1) / kT2);
    const size_t smem = ((size_t)(kT1 + k1 - 1) * kT2 + (size_t)k0 * kT1 * kT2) * sizeof(double) +
                        (size_t)(kT1 + k1 - 1) * (kT2 + k2 - 1) * sizeof(float);
    if (smem > 200 * 1024) return set_error(VU_ERR_UNSUPPORTED, "patch too large for shared memory");
    if (B > 65535) return set_error(VU_ERR_UNSUPPORTED, "B > 65535 per vu_patch_max call");
    static size_t attr = 0;
    if (smem > attr) {
        if (cudaFuncSetAttribute(patch_box<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(patch_box<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return set_cuda_error("cudaFuncSetAttribute(patch_box)");
        attr = smem;
    }
    unsigned long long* enc = reinterpret_cast<unsigned long long*>(out_max);
    const unsigned ib = (unsigned)((B + 255) / 256);
    patch_init<<<ib, 256, 0, stream>>>(enc, out_first, B);
    // 2-D maps with a square box of a specialised size: the column-marching kernel (4x faster than the tiled one)
    if (d0 == 1 && k0 == 1 && k1 == k2 && (k1 == 10 || k1 == 4 || k1 == 16) && (d1 - k1 + 1 + kRowsPerCta - 1) / kRowsPerCta <= 65535) {
        if (k1 == 10) launch_patch2d<10>(maps, B, d1, d2, enc, out_first, scale, stream);
        else if (k1 == 4) launch_patch2d<4>(maps, B, d1, d2, enc, out_first, scale, stream);
        else launch_patch2d<16>(maps, B, d1, d2, enc, out_first, scale, stream);
        patch_finish<<<ib, 256, 0, stream>>>(enc, B, scale);
        count_launch("patch_init"); count_launch("patch_box2d"); count_launch("patch_box2d"); count_launch("patch_finish");
        return check_launch("patch_box2d");
    }
    dim3 grid((unsigned)(tiles_y * tiles_x), (unsigned)B);
    patch_box<0><<<grid, kT1 * kT2, smem, stream>>>(maps, d0, d1, d2, k0, k1, k2, enc, out_first, scale, tiles_y, tiles_x);
    patch_box<1><<<grid, kT1 * kT2, smem, stream>>>(maps, d0, d1, d2, k0, k1, k2, enc, out_first, scale, tiles_y, tiles_x);
    patch_finish<<<ib, 256, 0, stream>>>(enc, B, scale);
    count_launch("patch_init"); count_launch("patch_box"); count_launch("patch_box"); count_launch("patch_finish");
    return check_launch("patch_box");
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

int launch_border(const uint8_t* labels, long long B, long long d0, long long d1, long long d2, long long* stats_i64,
                  cudaStream_t stream) {
    if (B > 65535) return set_error(VU_ERR_UNSUPPORTED, "B > 65535 per vu_border_count call");
    const long long V = d0 * d1 * d2;
    long long gx = (V + 256 * 8 - 1) / (256 * 8);
    if (gx < 1) gx = 1;
    if (gx > 4096) gx = 4096;
    border_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, stream>>>(labels, d0, d1, d2, stats_i64);
    count_launch("border");
    return check_launch("border");
}

}  // namespace vu
