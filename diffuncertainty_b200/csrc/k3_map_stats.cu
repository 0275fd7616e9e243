// K3: the K1 statistics phase on maps / labels that already sit in device memory
// (the reference evaluates stored maps: evaluation/eval_experiments.py:348-355).
#include "stats_v2.cuh"
#include "vu_host.h"

namespace vu {

extern __shared__ __align__(16) unsigned char vu_dyn_smem[];

struct K3Params {
    const float* maps[VU_N_UNC];
    const uint8_t* labels;
    long long B, V, tiles_per_img, total_tiles;
    StatParams st;
};

constexpr int kK3Threads = 256;

// VEC = 4: V % 4 == 0 and every pointer 16-byte (labels 4-byte) aligned; VEC = 1 otherwise.
template <int VEC>
__global__ void __launch_bounds__(kK3Threads) k3_map_stats(const __grid_constant__ K3Params prm) {
    constexpr long long kTileVox = (long long)kK3Threads * VEC;
    StatsCursor<kK3Threads> cursor;
    stats_init<kK3Threads>(prm.st, vu_dyn_smem);
    const int t0 = (int)(prm.total_tiles * (long long)blockIdx.x / gridDim.x);
    const int t1 = (int)(prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x);
    const int tpi = (int)prm.tiles_per_img;
    const bool light = !(prm.st.flags & (VU_STAT_CALIB | VU_STAT_PLATT_FIT | VU_STAT_NCC | VU_STAT_CLASS_COUNTS));
    int b = t0 / tpi, vt = t0 - b * tpi - 1;
    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        cursor.enter(prm.st, vu_dyn_smem, b, vt, kTileVox);
        if (light && tile + 1 < t1) {
            // memory-bound masks (sums / thresholds / area: ~20 instructions per voxel): pull the next tile towards the SM
            // while this one is processed (0.077 -> 0.061 ms on 256 x 256^2).  With calibration histograms the pass is
            // issue-bound and the extra instructions only cost (ncu r01f: +7 % instructions, no time gained).
            const int nvt = vt + 1 == tpi ? 0 : vt + 1;
            const long long nb = vt + 1 == tpi ? b + 1 : b;
            const long long nv = (long long)nvt * kTileVox + (long long)threadIdx.x * VEC;
            if (nv < prm.V) {
                const long long o = nb * prm.V + nv;
#pragma unroll
                for (int k = 0; k < VU_N_UNC; ++k)
                    if (prm.maps[k]) asm volatile("prefetch.global.L1 [%0];" ::"l"(prm.maps[k] + o));
                if (prm.labels) asm volatile("prefetch.global.L1 [%0];" ::"l"(prm.labels + o));
            }
        }
        const long long v = (long long)vt * kTileVox + (long long)threadIdx.x * VEC;
        const bool active = v < prm.V;
        float u[VU_N_UNC][VEC];
        int label[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { u[0][j] = u[1][j] = u[2][j] = 0.f; label[j] = 0; }
        if (active) {
            const long long o = (long long)b * prm.V + v;
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k)
                if (prm.maps[k]) VecLoad<VEC>::load(prm.maps[k] + o, u[k]);
            if (prm.labels) {
                if (VEC == 4) {
                    const unsigned w = __ldg(reinterpret_cast<const unsigned*>(prm.labels + o));
#pragma unroll
                    for (int j = 0; j < VEC; ++j) label[j] = (int)((w >> (8 * j)) & 0xffu);
                } else {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) label[j] = (int)__ldg(prm.labels + o + j);
                }
            }
        }
        stats_tile<VEC, kK3Threads>(prm.st, vu_dyn_smem, active, b, v, u, label);
    }
    if (t1 > t0) cursor.finish(prm.st, vu_dyn_smem, vt, kTileVox);
}


// ---- lean form (stats_v2.cuh) for the common masks: uint8 references with word-aligned rows, all three maps present,
// no label LUT, V % 4 == 0.  Register partials per thread, 16 histogram replicas per warp.
constexpr int kK3v2Threads = 256, kK3v2Rep = 16;

template <unsigned FL, int RMAX>
__global__ void __launch_bounds__(kK3v2Threads) k3_map_stats_v2(const __grid_constant__ K3Params prm) {
    constexpr long long kTileVox = (long long)kK3v2Threads * 4;
    constexpr int kWarps = kK3v2Threads / 32;
    const StatParams& sp = prm.st;
    const int tid = threadIdx.x, warp = tid >> 5;
    stats2_init<kK3v2Rep>(sp, vu_dyn_smem, tid, kK3v2Threads, kWarps);
    __syncthreads();
    Stat2Ctx cx;
    stats2_ctx<kK3v2Rep>(cx, sp, vu_dyn_smem, warp, kWarps);
    // the maps in this lane's step order (stats_v2.cuh, "type rotation")
    const bool rot = stats2_rotated<kK3v2Rep>();
    const float* maps[VU_N_UNC];
#pragma unroll
    for (int s = 0; s < VU_N_UNC; ++s) maps[s] = rot ? prm.maps[(s + 1) % VU_N_UNC] : prm.maps[s];
    StatAcc<FL, RMAX> A;
    A.clear();
    const int t0 = (int)(prm.total_tiles * (long long)blockIdx.x / gridDim.x);
    const int t1 = (int)(prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x);
    const int tpi = (int)prm.tiles_per_img;
    int b = t0 / tpi, vt = t0 - b * tpi - 1;
    int cur_b = -1, vt_begin = 0;
    auto flush = [&]() {
        stats2_flush_regs<FL, RMAX, kK3v2Rep>(A, sp, cur_b);
        if (FL & VU_STAT_CALIB) stats2_flush_hist_warp<kK3v2Rep>(sp, vu_dyn_smem, cur_b, warp);
    };
    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        if (b != cur_b || vt - vt_begin >= kMaxTilesPerFlush2) {
            if (cur_b >= 0) flush();
            cur_b = b;
            vt_begin = vt;
        }
        const long long v = (long long)vt * kTileVox + (long long)tid * 4;
        const bool active = v < prm.V;
        float4 U[VU_N_UNC];
        unsigned lab4 = 0u;
#pragma unroll
        for (int k = 0; k < VU_N_UNC; ++k) U[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            const long long o = (long long)b * prm.V + v;
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k) U[k] = ldg_stream(reinterpret_cast<const float4*>(maps[k] + o));
            lab4 = __ldg(reinterpret_cast<const unsigned*>(prm.labels + o));
        }
        unsigned W[RMAX];
        stats2_load_refs<FL, RMAX>(sp, active, reinterpret_cast<const uint8_t*>(sp.gt.data) + (long long)b * sp.gt.sb, v, W);
        stats2_tile<FL, RMAX, kK3v2Rep>(A, sp, cx, active, b, U[0], U[1], U[2], lab4, W);
    }
    if (cur_b >= 0) flush();
}

template <unsigned FL>
static int launch_v2(const K3Params& prm, const StatParams& st, cudaStream_t stream) {
    void (*fn)(const K3Params) = st.gt.R <= 4 ? k3_map_stats_v2<FL, 4> : k3_map_stats_v2<FL, VU_MAX_RATERS>;
    const size_t dyn = stats2_smem_bytes(st.flags, kK3v2Threads, kK3v2Rep);
    if (dyn > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(k3_map_stats_v2)");
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kK3v2Threads, dyn) != cudaSuccess || occ < 1)
        return set_cuda_error("occupancy query (k3_map_stats_v2)");
    long long grid = (long long)device_sm_count() * occ;
    if (grid > prm.total_tiles) grid = prm.total_tiles;
    fn<<<(unsigned)grid, kK3v2Threads, dyn, stream>>>(prm);
    count_launch("k3_map_stats_v2");
    return check_launch("k3_map_stats_v2");
}

// the lean form applies to: compile-time mask, all three maps, uint8 word-aligned references (or none needed), no LUT
bool stats2_eligible(const StatParams& st, long long V) {
    const unsigned needs_gt = VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC;
    if (st.unc_mask != 7u || st.lut || st.ncc_gt_map || (st.flags & VU_STAT_PLATT_FIT) || V % 4) return false;
    if (st.flags & needs_gt) {
        if (!st.gt.data || st.gt.dtype != VU_GT_U8 || st.gt.align < 4) return false;
    }
    for (int k = 0; k < VU_N_UNC; ++k)
        if ((st.flags & VU_STAT_CALIB) && st.calib[k].identity) return false;
    return true;
}

int launch_map_stats(const vu_map_stats_args* a, const StatParams& st, cudaStream_t stream) {
    K3Params prm;
    bool vec4 = (a->V % 4) == 0 && ((uintptr_t)a->labels % 4) == 0;
    for (int k = 0; k < VU_N_UNC; ++k) {
        prm.maps[k] = a->maps[k];
        vec4 = vec4 && ((uintptr_t)a->maps[k] % 16) == 0;
    }
    prm.labels = a->labels;
    prm.B = a->B; prm.V = a->V;
    const long long tile_vox = (long long)kK3Threads * (vec4 ? 4 : 1);
    prm.tiles_per_img = (a->V + tile_vox - 1) / tile_vox;
    prm.total_tiles = prm.tiles_per_img * a->B;
    if (prm.total_tiles >= (1LL << 31)) return set_error(VU_ERR_UNSUPPORTED, "more than 2^31 tiles in one launch; split the batch");
    prm.st = st;
    if (vec4 && a->labels && stats2_eligible(st, a->V) && get_option("stats_path", 0) != 1) {
        switch (st.flags) {
            case 0x07u: return launch_v2<0x07u>(prm, st, stream);
            case 0x0fu: return launch_v2<0x0fu>(prm, st, stream);
            case 0x1du: return launch_v2<0x1du>(prm, st, stream);
            case 0x1fu: return launch_v2<0x1fu>(prm, st, stream);
            case 0x21u: return launch_v2<0x21u>(prm, st, stream);
            case 0x3fu: return launch_v2<0x3fu>(prm, st, stream);
            default: break;
        }
    }
    void (*fn)(const K3Params) = vec4 ? k3_map_stats<4> : k3_map_stats<1>;
    const size_t dyn = stats_smem_bytes(st.flags, st.gt.R, kK3Threads) + stats_class_bytes(st.flags, st.gt.R, st.ncls);
    if (dyn > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(k3_map_stats)");
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kK3Threads, dyn) != cudaSuccess || occ < 1)
        return set_cuda_error("occupancy query (k3_map_stats)");
    long long grid = (long long)device_sm_count() * occ;
    if (grid > prm.total_tiles) grid = prm.total_tiles;
    fn<<<(unsigned)grid, kK3Threads, dyn, stream>>>(prm);
    count_launch("k3_map_stats");
    return check_launch("k3_map_stats");
}

}  // namespace vu
