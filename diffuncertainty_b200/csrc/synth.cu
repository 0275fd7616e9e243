// Deterministic on-device synthetic inputs for benchmarks / smoke tests
// (images keyed by (seed, image index) so any GPU count
// sees the same data).  Not part of the reference; plays the role of the
// stochastic forward passes that produce the slab (test_2D.py:1121-1277).
#include "vu_common.cuh"
#include "vu_host.h"

namespace vu {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(256) synth_slab_kernel(float* __restrict__ out, long long P, long long B, long long C,
                                                         long long V, uint64_t seed, long long first_image, float scale) {
    const long long n = P * B * V;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long v = i % V, b = (i / V) % B, p = i / (V * B);
        const uint64_t key = mix64(seed ^ mix64((uint64_t)(first_image + b) * 0x100000001b3ull + (uint64_t)p)) + (uint64_t)v * 0x9e3779b97f4a7c15ull;
        float* o = out + ((p * B + b) * C) * V + v;
        float mx = -1e30f;
        // pass 1: logits -> running max; pass 2 recomputes them (cheaper than storing C values)
        for (int pass = 0; pass < 2; ++pass) {
            float sum = 0.f;
            for (long long c = 0; c < C; c += 2) {
                const uint64_t r = mix64(key + (uint64_t)(c >> 1) * 0xd1342543de82ef95ull);
                const float rad = sqrtf(-2.0f * __logf(u01((uint32_t)r)));
                float s, co;
                __sincosf(6.2831853f * u01((uint32_t)(r >> 32)), &s, &co);
                const float z0 = scale * rad * co, z1 = scale * rad * s;
                if (pass == 0) {
                    mx = fmaxf(mx, z0);
                    if (c + 1 < C) mx = fmaxf(mx, z1);
                } else {
                    const float e0 = __expf(z0 - mx);
                    o[c * V] = e0; sum += e0;
                    if (c + 1 < C) { const float e1 = __expf(z1 - mx); o[(c + 1) * V] = e1; sum += e1; }
                }
            }
            if (pass == 1) {
                const float inv = 1.0f / sum;
                for (long long c = 0; c < C; ++c) o[c * V] *= inv;
            }
        }
    }
}

__global__ void __launch_bounds__(256) synth_gt_kernel(uint8_t* __restrict__ out, const float* __restrict__ slab, long long P,
                                                       long long B, long long C, long long V, int R, uint64_t seed,
                                                       long long first_image, float flip, float ignore_frac, int ignore_value) {
    const long long n = B * V;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long v = i % V, b = i / V;
        const float* x = slab + (b * C) * V + v;  // member 0
        float best = x[0];
        int lab = 0;
        for (long long c = 1; c < C; ++c) { const float t = x[c * V]; if (t > best) { best = t; lab = (int)c; } }
        const uint64_t key = mix64(seed ^ 0x5bd1e995u ^ mix64((uint64_t)(first_image + b))) + (uint64_t)v * 0x9e3779b97f4a7c15ull;
        for (int r = 0; r < R; ++r) {
            const uint64_t h = mix64(key + (uint64_t)r * 0xd1342543de82ef95ull);
            int g = lab;
            if (u01((uint32_t)h) < flip) g = (int)((h >> 32) % (uint64_t)C);
            if (u01((uint32_t)(h >> 20)) < ignore_frac) g = ignore_value;
            out[(b * R + r) * V + v] = (uint8_t)g;
        }
    }
}

int launch_synth_slab(float* out, long long P, long long B, long long C, long long V, uint64_t seed, long long first_image,
                      float scale, cudaStream_t stream) {
    long long blocks = (P * B * V + 255) / 256;
    const long long cap = (long long)device_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    synth_slab_kernel<<<(unsigned)blocks, 256, 0, stream>>>(out, P, B, C, V, seed, first_image, scale);
    count_launch("synth_slab");
    return check_launch("synth_slab");
}

int launch_synth_gt(uint8_t* out, const float* slab, long long P, long long B, long long C, long long V, int R, uint64_t seed,
                    long long first_image, float flip, float ignore_frac, int ignore_value, cudaStream_t stream) {
    long long blocks = (B * V + 255) / 256;
    const long long cap = (long long)device_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    synth_gt_kernel<<<(unsigned)blocks, 256, 0, stream>>>(out, slab, P, B, C, V, R, seed, first_image, flip, ignore_frac, ignore_value);
    count_launch("synth_gt");
    return check_launch("synth_gt");
}

}  // namespace vu
