// K3: the K1 statistics phase on maps / labels that already sit in device memory
// (the reference evaluates stored maps: evaluation/eval_experiments.py:348-355).
#include "vu_common.cuh"
#include "vu_host.h"

namespace vu {

struct K3Params {
    const float* maps[VU_N_UNC];
    const uint8_t* labels;
    long long B, V, tiles_per_img, total_tiles;
    StatParams st;
};

constexpr int kK3Threads = 256, kK3PerThread = 4;

__global__ void __launch_bounds__(kK3Threads) k3_map_stats(const __grid_constant__ K3Params prm) {
    constexpr int WARPS = kK3Threads / 32;
    __shared__ CtaStats<WARPS> cs;
    cs.init(prm.st);
    const long long t0 = prm.total_tiles * (long long)blockIdx.x / gridDim.x;
    const long long t1 = prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x;
    long long cur_b = -1;
    for (long long tile = t0; tile < t1; ++tile) {
        const long long b = tile / prm.tiles_per_img;
        const long long vt = tile - b * prm.tiles_per_img;
        if (b != cur_b) {
            if (cur_b >= 0) cs.flush(prm.st, cur_b);
            cur_b = b;
        }
        TileAcc acc;
        acc.clear();
#pragma unroll 1
        for (int j = 0; j < kK3PerThread; ++j) {
            const long long v = (vt * kK3PerThread + j) * kK3Threads + threadIdx.x;
            const bool active = v < prm.V;
            float u[VU_N_UNC] = {0.f, 0.f, 0.f};
            int label = 0;
            if (active) {
                const long long o = b * prm.V + v;
#pragma unroll
                for (int k = 0; k < VU_N_UNC; ++k)
                    if (prm.maps[k]) u[k] = __ldg(prm.maps[k] + o);
                if (prm.labels) label = __ldg(prm.labels + o);
            }
            stats_voxel<WARPS>(prm.st, cs, acc, active, b, v, u, label);
        }
        stats_tile_end<WARPS>(prm.st, cs, acc);
    }
    if (cur_b >= 0) cs.flush(prm.st, cur_b);
}

int launch_map_stats(const vu_map_stats_args* a, const StatParams& st, cudaStream_t stream) {
    K3Params prm;
    for (int k = 0; k < VU_N_UNC; ++k) prm.maps[k] = a->maps[k];
    prm.labels = a->labels;
    prm.B = a->B; prm.V = a->V;
    const long long tile_vox = (long long)kK3Threads * kK3PerThread;
    prm.tiles_per_img = (a->V + tile_vox - 1) / tile_vox;
    prm.total_tiles = prm.tiles_per_img * a->B;
    prm.st = st;
    long long grid = (long long)device_sm_count() * 4;
    if (grid > prm.total_tiles) grid = prm.total_tiles;
    k3_map_stats<<<(unsigned)grid, kK3Threads, 0, stream>>>(prm);
    count_launch("k3_map_stats");
    return check_launch("k3_map_stats");
}

}  // namespace vu
