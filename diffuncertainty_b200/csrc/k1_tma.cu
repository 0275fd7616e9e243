// K1 (TMA-pipelined, warp-specialised form): the fused streaming pass with the slab staged through a
// shared-memory ring by bulk asynchronous copies (cp.async.bulk, the 1-D TMA path) and the per-image
// statistics handed to dedicated warps.
//
// Why: the register-streaming kernel (k1_fused.cu) alternates, per warp, between a load phase and a
// compute / epilogue / statistics phase; while a warp computes it has no loads in flight, and with the
// register budget of C = 19 only 16 warps fit on an SM, so HBM idles in those gaps (ncu r01: with the
// statistics fused in, DRAM throughput drops from 94 % to 70 % of the measured peak although the SM issues
// only 46 % of its slots).  Here one persistent CTA per SM holds three kinds of warps:
//   producer (1 warp)      issues cp.async.bulk copies of [members x classes x TILE voxels] rows into a ring
//                          of NSTAGES stages; completion is counted in bytes on `full` mbarriers;
//   consumers (CT/32 warps) read a stage with conflict-free LDS.64/128, run the shared per-voxel arithmetic
//                          (VoxelAcc), release the stage through `empty` mbarriers, write maps / labels to
//                          HBM and drop (TU, AU, EU, label) of the tile into a double-buffered hand-off area;
//   statistics (4 warps)   pick the tile up from the hand-off area and run the statistics phase
//                          (vu_common.cuh) with their own registers, slots and histograms.
// Loads therefore stay in flight whatever the consumers do, the consumers' code is the same with and
// without statistics, and the latency-bound statistics code overlaps the issue-bound arithmetic.
//
// Reference semantics: see k1_fused.cu / k1_core.cuh.
#include "k1_core.cuh"
#include "tma_common.cuh"
#include "vu_host.h"

namespace vu {

extern __shared__ __align__(128) unsigned char vu_tma_smem[];

struct K1TmaParams {
    const float* x;
    const float* const* mptr;  // optional device array of P member base pointers (then x / sp are unused)
    long long P, B, C, V;
    long long sp, sb, sc;
    float* tu;
    float* au;
    float* eu;
    uint8_t* lab;
    uint8_t* mlab;  // per-member labels (P, B, V) or NULL
    long long tiles_per_img, total_tiles;
    int nstages;
    unsigned bar_offset;    // byte offsets inside dynamic shared memory
    unsigned hand_offset;
    unsigned stats_offset;
    StatParams st;
};

constexpr int kStatVec = 4;        // voxels per statistics thread and pass

// The tile loop of the statistics warps: pick (TU, AU, EU, label) of each tile up from the hand-off buffer the consumers
// filled and run the statistics phase on kStatVec voxels per thread and pass.
// ST: statistics threads (a multiple of 32), REP: histogram replicas per warp (32: one per lane; 16: two half-warp phases).
template <int TV, int TID0, unsigned FL, int ST, int REP>
__device__ __forceinline__ void stat_warps_loop(const StatParams& st, void* st_smem, const unsigned char* hand, unsigned hand_bytes,
                                                unsigned hfull0, unsigned hempty0, int t0, int t1, int tpi, long long V) {
    constexpr int kRep = REP;
    constexpr int kStatThreads = ST;
    constexpr int kPasses = (TV + kStatThreads * kStatVec - 1) / (kStatThreads * kStatVec);
    const int s = (int)threadIdx.x - TID0;
    StatsCursor<kStatThreads> cursor;
    stats_init<kStatThreads, 2, TID0, kRep>(st, st_smem);
    int b = t0 / tpi, vt = t0 - b * tpi - 1;
    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        cursor.template enter<2, TID0, kRep>(st, st_smem, b, vt, TV);
        const int buf = (tile - t0) & 1;
        const unsigned par = ((tile - t0) >> 1) & 1;
        const long long v0 = (long long)vt * TV;
#pragma unroll 1
        for (int pass = 0; pass < kPasses; ++pass) {
            const int q = (pass * kStatThreads + s) * kStatVec;
            if (q < TV && v0 + q < V) stats_prefetch_gt<kStatVec>(st, b, v0 + q);
        }
        mbar_wait_relaxed(hfull0 + 8 * buf, par);  // the consumers have written this tile's maps and labels
        const unsigned char* hb = hand + (size_t)buf * hand_bytes;
#pragma unroll 1
        for (int pass = 0; pass < kPasses; ++pass) {
            const int q = (pass * kStatThreads + s) * kStatVec;  // first voxel of this thread inside the tile
            const bool in_tile = q < TV;                         // (the last pass may overhang the tile)
            float u[VU_N_UNC][kStatVec];
            int label[kStatVec];
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k) {
                const float4 w = in_tile ? *reinterpret_cast<const float4*>(hb + ((size_t)k * TV + q) * sizeof(float))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                u[k][0] = w.x; u[k][1] = w.y; u[k][2] = w.z; u[k][3] = w.w;
            }
            const unsigned lw = in_tile ? *reinterpret_cast<const unsigned*>(hb + (size_t)12 * TV + q) : 0u;
#pragma unroll
            for (int j = 0; j < kStatVec; ++j) label[j] = (int)((lw >> (8 * j)) & 0xffu);
            stats_tile<kStatVec, kStatThreads, TID0, kRep, FL>(st, st_smem, in_tile && v0 + q < V, b, v0 + q, u, label);
        }
        __syncwarp();
        if ((s & 31) == 0) mbar_arrive(hempty0 + 8 * buf);  // this warp is done with the hand-off buffer
    }
    if (t1 > t0) cursor.template finish<2, TID0, kRep>(st, st_smem, vt, TV);
}

// C classes, VEC voxels per consumer thread, CT consumer threads.
// NCH == 1: a stage holds G whole members (G x C rows).  NCH == 2 (G == 1, VEC >= 2): a stage holds half a
// member's classes ((C+1)/2 rows), for C = 19 where a whole member of a 1024-voxel tile would take 76 KB.
// ST statistics threads (0: a launch without statistics), SREP histogram replicas per statistics warp.
// LG: the slab holds logits (VU_SLAB_LOGITS): every member is softmax'ed over its classes as it is consumed (whole members per
// stage only: the maximum over all classes comes first).
// OH: every member is replaced by the one-hot vector of its argmax as it is consumed (VU_SLAB_DISCRETIZE, --discretize).
template <int C, int VEC, int LEVELS, int CT, int G, int NCH, int ST, int SREP, bool LG = false, bool OH = false>
__global__ void __launch_bounds__(CT + 32 + ST, 1) k1_tma(const __grid_constant__ K1TmaParams prm) {
    constexpr bool STATS = ST > 0;
    static_assert(!(LG || OH) || NCH == 1, "logits / one-hot members need whole members per stage");
    static_assert(!(LG && OH), "one producer per kernel");
    constexpr int kStatThreads = ST;
    static_assert(NCH == 1 || (NCH == 2 && G == 1 && VEC >= 2), "class chunks: one member per stage, VEC >= 2");
    constexpr int TV = CT * VEC;  // voxels per tile
    constexpr int CH = (NCH == 1) ? C : (C + 1) / 2;
    constexpr unsigned kRowBytes = TV * sizeof(float);
    constexpr unsigned kStageBytes = (NCH == 1 ? G * C : CH) * kRowBytes;
    constexpr int kStageFloats = kStageBytes / sizeof(float);
    constexpr int kProducer0 = CT, kStat0 = CT + 32;
    constexpr unsigned kHandBytes = 13u * TV;  // per buffer: TU, AU, EU (fp32) + label (u8) of one tile
    using Acc = VoxelAcc<C, VEC, LEVELS>;
    constexpr int NH = Acc::NH;

    const int nstages = prm.nstages;
    float* ring = reinterpret_cast<float*>(vu_tma_smem);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(vu_tma_smem + prm.bar_offset);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + nstages);
    const unsigned hfull0 = smem_u32(bars + 2 * nstages), hempty0 = smem_u32(bars + 2 * nstages + 2);
    unsigned char* hand = vu_tma_smem + prm.hand_offset;
    void* st_smem = vu_tma_smem + prm.stats_offset;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(full0 + 8 * s, 1);         // the producer's arrive.expect_tx
            mbar_init(empty0 + 8 * s, CT / 32);  // one arrive per consumer warp
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(hfull0 + 8 * s, CT / 32);
            mbar_init(hempty0 + 8 * s, STATS ? kStatThreads / 32 : 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long P = prm.P, V = prm.V;
    const int t0 = (int)(prm.total_tiles * (long long)blockIdx.x / gridDim.x);
    const int t1 = (int)(prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x);
    const int tpi = (int)prm.tiles_per_img;
    const int fills = (NCH == 1) ? (int)((P + G - 1) / G) : (int)(2 * P);  // stage fills per tile

    if (tid >= kProducer0 && tid < kStat0) {
        // ------------------------------ producer warp ---------------------------------------------------
        const int lane = tid - kProducer0;
        unsigned long long policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        int b = t0 / tpi, vt = t0 - b * tpi - 1;
        int stage = 0;
        unsigned phase = 0;
        for (int tile = t0; tile < t1; ++tile) {
            if (++vt == tpi) { vt = 0; ++b; }
            const long long v0 = (long long)vt * TV;
            const long long left = V - v0;
            const unsigned row_bytes = left >= TV ? kRowBytes : (unsigned)(left * sizeof(float));
            const long long off0 = (long long)b * prm.sb + v0;  // offset of the tile inside a member
            for (int fi = 0; fi < fills; ++fi) {
                int p0, c0, nrows;
                if (NCH == 1) {
                    p0 = fi * G; c0 = 0;
                    nrows = ((P - p0) < G ? (int)(P - p0) : G) * C;
                } else {
                    p0 = fi >> 1; c0 = (fi & 1) * CH;
                    nrows = (fi & 1) ? (C - CH) : CH;
                }
                mbar_wait(empty0 + 8 * stage, phase ^ 1);  // slot free (the first pass falls through); the producer does not back off:
                                                           // a late refill is a bubble in the HBM stream (measured: -5 % with a 100 ns sleep)
                if (lane == 0) mbar_arrive_expect_tx(full0 + 8 * stage, (unsigned)nrows * row_bytes);
                __syncwarp();
                const unsigned dst0 = smem_u32(ring) + (unsigned)stage * kStageBytes;
                for (int r = lane; r < nrows; r += 32) {
                    const int g = (NCH == 1) ? r / C : 0;
                    const int c = (NCH == 1) ? r - g * C : c0 + r;
                    const float* mem = prm.mptr ? ld_member_ptr(prm.mptr, p0 + g) : prm.x + (long long)(p0 + g) * prm.sp;
                    bulk_g2s(dst0 + (unsigned)r * kRowBytes, mem + off0 + (long long)c * prm.sc, row_bytes, full0 + 8 * stage, policy);
                }
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    if (STATS && tid >= kStat0) {
        // ------------------------------ statistics warps ------------------------------------------------
        // common masks get a compile-time specialisation of the statistics phase (no flag tests, no dead code)
        const StatParams& st = prm.st;
        const bool plain = st.unc_mask == 7u && st.lut == nullptr && st.gt.dtype == VU_GT_U8;
        auto run = [&](auto fl_tag) {
            constexpr unsigned FL = decltype(fl_tag)::value;
            stat_warps_loop<TV, kStat0, FL, (ST > 0 ? ST : 32), SREP>(st, st_smem, hand, kHandBytes, hfull0, hempty0, t0, t1, tpi, V);
        };
        if (plain && st.flags == 0x1fu) run(std::integral_constant<unsigned, 0x1fu>());
        else if (plain && st.flags == 0x3fu) run(std::integral_constant<unsigned, 0x3fu>());
        else if (plain && st.flags == 0x1du) run(std::integral_constant<unsigned, 0x1du>());
        else if (plain && st.flags == 0x21u) run(std::integral_constant<unsigned, 0x21u>());
        else if (st.unc_mask == 7u && st.flags == 0x07u) run(std::integral_constant<unsigned, 0x07u>());
        else run(std::integral_constant<unsigned, kRuntimeFlags>());
        return;
    }

    // ---------------------------------- consumer warps ------------------------------------------------------
    const float Pf = (float)P;
    int b = t0 / tpi, vt = t0 - b * tpi - 1;
    int stage = 0;
    unsigned phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        const long long v = (long long)vt * TV + (long long)tid * VEC;
        const bool active = v < V;
        float u[VU_N_UNC][VEC];
        int label[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { u[0][k] = u[1][k] = u[2][k] = 0.f; label[k] = 0; }

        Acc acc;
        acc.init();
        for (int fi = 0; fi < fills; ++fi) {
            mbar_wait(full0 + 8 * stage, phase);  // the bytes of this stage have landed
            const float* sbase = ring + (size_t)stage * kStageFloats + tid * VEC;
            // (a partial last tile leaves stale bytes behind the image's end: those threads are inactive and
            //  their arithmetic is discarded)
            if constexpr (NCH == 1) {
                const int p0 = fi * G;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (G == 1 || p0 + g < P) {
                        f32x2 xp[Acc::NP];
                        float xs = 0.f;
                        const float* srow = sbase + (size_t)g * C * TV;
                        if constexpr (VEC == 4) {
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                const float4 w = *reinterpret_cast<const float4*>(srow + c * TV);
                                xp[c * 2] = pk2(w.x, w.y);
                                xp[c * 2 + 1] = pk2(w.z, w.w);
                            }
                        } else if constexpr (VEC == 2) {
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                const float2 w = *reinterpret_cast<const float2*>(srow + c * TV);
                                xp[c] = pk2(w.x, w.y);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < Acc::NP; ++j) xp[j] = pk2(srow[(2 * j) * TV], srow[(2 * j + 1) * TV]);
                            if constexpr (Acc::ODD) xs = srow[(C - 1) * TV];
                        }
                        if constexpr (OH) {
                            acc.add_member_onehot(xp, xs, p0 + g);
                        } else if constexpr (LG) {
                            f32x2 hm[Acc::NH], rs[Acc::NH];
                            float hs = 0.f;
                            auto reload = [&](int i) {
                                if constexpr (VEC == 4) {
                                    const float2 w = *reinterpret_cast<const float2*>(srow + (i >> 1) * TV + 2 * (i & 1));
                                    return pk2(w.x, w.y);
                                } else if constexpr (VEC == 2) {
                                    const float2 w = *reinterpret_cast<const float2*>(srow + i * TV);
                                    return pk2(w.x, w.y);
                                } else {
                                    return pk2(srow[(2 * i) * TV], srow[(2 * i + 1) * TV]);
                                }
                            };
                            acc.softmax_member(xp, xs, rs, hm, hs, reload, [&]() { return srow[(C - 1) * TV]; });
                            acc.add_member_pre(xp, xs, rs, hm, hs, p0 + g, prm.mlab != nullptr);
                        } else {
                            acc.add_member(xp, xs, p0 + g, prm.mlab != nullptr);
                        }
                        if (prm.mlab && active) VecLoad<VEC>::store_u8(prm.mlab + ((long long)(p0 + g) * prm.B + b) * V + v, acc.bi);
                    }
                }
            } else {
                auto load_chunk = [&](auto c0_tag, auto c1_tag) {
                    constexpr int C0 = decltype(c0_tag)::value, C1 = decltype(c1_tag)::value;
                    f32x2 xp[(C1 - C0) * NH];
#pragma unroll
                    for (int c = 0; c < C1 - C0; ++c) {
                        if constexpr (VEC == 4) {
                            const float4 w = *reinterpret_cast<const float4*>(sbase + c * TV);
                            xp[c * 2] = pk2(w.x, w.y);
                            xp[c * 2 + 1] = pk2(w.z, w.w);
                        } else {
                            const float2 w = *reinterpret_cast<const float2*>(sbase + c * TV);
                            xp[c] = pk2(w.x, w.y);
                        }
                    }
                    acc.template add_classes<C0, C1>(xp, prm.mlab != nullptr);
                };
                if ((fi & 1) == 0) {
                    acc.begin_member();
                    load_chunk(std::integral_constant<int, 0>(), std::integral_constant<int, CH>());
                } else {
                    load_chunk(std::integral_constant<int, CH>(), std::integral_constant<int, C>());
                    acc.end_member(fi >> 1);
                    if (prm.mlab && active) VecLoad<VEC>::store_u8(prm.mlab + ((long long)(fi >> 1) * prm.B + b) * V + v, acc.bi);
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * stage);  // this warp is done reading the stage
            if (++stage == nstages) { stage = 0; phase ^= 1; }
        }

        if (active) {
            acc.finish(Pf, u, label);
            const long long o = (long long)b * V + v;
            if (prm.tu) VecLoad<VEC>::store(prm.tu + o, u[0]);
            if (prm.au) VecLoad<VEC>::store(prm.au + o, u[1]);
            if (prm.eu) VecLoad<VEC>::store(prm.eu + o, u[2]);
            if (prm.lab) VecLoad<VEC>::store_u8(prm.lab + o, label);
        }
        if (STATS) {
            const int buf = (tile - t0) & 1;
            const unsigned par = ((tile - t0) >> 1) & 1;
            mbar_wait(hempty0 + 8 * buf, par ^ 1);  // the statistics warps are done with this buffer (first use falls through)
            unsigned char* hb = hand + (size_t)buf * kHandBytes;
#pragma unroll
            for (int k = 0; k < VU_N_UNC; ++k) {
                float* dst = reinterpret_cast<float*>(hb + ((size_t)k * TV + tid * VEC) * sizeof(float));
                if constexpr (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(u[k][0], u[k][1], u[k][2], u[k][3]);
                else if constexpr (VEC == 2) *reinterpret_cast<float2*>(dst) = make_float2(u[k][0], u[k][1]);
                else dst[0] = u[k][0];
            }
            unsigned char* ldst = hb + (size_t)12 * TV + tid * VEC;
            if constexpr (VEC == 4) *reinterpret_cast<uchar4*>(ldst) = make_uchar4(label[0], label[1], label[2], label[3]);
            else if constexpr (VEC == 2) *reinterpret_cast<uchar2*>(ldst) = make_uchar2(label[0], label[1]);
            else ldst[0] = (unsigned char)label[0];
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(hfull0 + 8 * buf);
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef void (*K1TmaKernel)(const K1TmaParams);
struct TmaVariant {
    int C, VEC, LEVELS, CT, G, NCH, ST, SREP;
    int use;  // automatic selection: 0 = any launch, 1 = only launches with reference-based statistics on few-class slabs, -1 = never
    int logits;  // 1: built for slabs of logits (VU_SLAB_LOGITS), 2: for members replaced by their one-hot argmax (VU_SLAB_DISCRETIZE)
    K1TmaKernel fn, fn_stats;
};
#define VU_TMA_S(C, VEC, LEVELS, CT, G, NCH, ST, SREP, USE)                                                 \
    { C, VEC, LEVELS, CT, G, NCH, ST, SREP, USE, 0, (K1TmaKernel)k1_tma<C, VEC, LEVELS, CT, G, NCH, 0, 32>, \
      (K1TmaKernel)k1_tma<C, VEC, LEVELS, CT, G, NCH, ST, SREP> }
#define VU_TMA(C, VEC, LEVELS, CT, G, NCH) VU_TMA_S(C, VEC, LEVELS, CT, G, NCH, 96, 32, 0)
#define VU_TMA_OH(C, VEC, LEVELS, CT, G)                                                                            \
    { C, VEC, LEVELS, CT, G, 1, 96, 32, 0, 2, (K1TmaKernel)k1_tma<C, VEC, LEVELS, CT, G, 1, 0, 32, false, true>,    \
      (K1TmaKernel)k1_tma<C, VEC, LEVELS, CT, G, 1, 96, 32, false, true> }
#define VU_TMA_LG(C, VEC, LEVELS, CT, G, USE)                                                                 \
    { C, VEC, LEVELS, CT, G, 1, 96, 32, USE, 1, (K1TmaKernel)k1_tma<C, VEC, LEVELS, CT, G, 1, 0, 32, true>,   \
      (K1TmaKernel)k1_tma<C, VEC, LEVELS, CT, G, 1, 96, 32, true> }

static const TmaVariant kTma[] = {
    VU_TMA(2, 4, 1, 512, 2, 1),  VU_TMA(2, 4, 2, 512, 2, 1),   // 0, 1
    VU_TMA(19, 2, 1, 512, 1, 2), VU_TMA(19, 2, 2, 512, 1, 2),  // 2, 3
    VU_TMA(2, 4, 1, 256, 4, 1),  VU_TMA(2, 4, 2, 256, 4, 1),   // 4, 5
    VU_TMA(19, 2, 1, 256, 1, 1), VU_TMA(19, 2, 2, 256, 1, 1),  // 6, 7
    VU_TMA(3, 4, 1, 256, 2, 1),  VU_TMA(3, 4, 2, 256, 2, 1),   // 8, 9
    VU_TMA(4, 4, 1, 256, 2, 1),  VU_TMA(4, 4, 2, 256, 2, 1),   // 10, 11
    // few classes with reference-based statistics: the statistics phase outweighs the streaming arithmetic, so the warp
    // split follows the work (consumer warps : statistics warps)
    VU_TMA_S(2, 4, 1, 256, 1, 1, 256, 16, -1), VU_TMA_S(2, 4, 2, 512, 2, 1, 192, 32, -1),  // 12, 13   8 : 8 (16 replicas), 16 : 6
    // slabs of logits: whole members per stage; C = 19 with one voxel per thread (19 values + 19 sums in registers)
    VU_TMA_LG(19, 1, 1, 512, 1, 0), VU_TMA_LG(19, 1, 2, 512, 1, 0),  // 14, 15
    VU_TMA_LG(2, 4, 1, 512, 2, 0),  VU_TMA_LG(2, 4, 2, 512, 2, 0),   // 16, 17
    VU_TMA_LG(3, 4, 1, 256, 2, 0),  VU_TMA_LG(3, 4, 2, 256, 2, 0),   // 18, 19
    VU_TMA_LG(4, 4, 1, 256, 2, 0),  VU_TMA_LG(4, 4, 2, 256, 2, 0),   // 20, 21
    VU_TMA_LG(19, 2, 1, 256, 1, -1), VU_TMA_LG(19, 2, 2, 256, 1, -1),  // 22, 23  (two voxels per thread, 8 consumer warps)
    // members replaced by their one-hot argmax (--discretize): same stage shapes
    VU_TMA_OH(19, 1, 1, 512, 1), VU_TMA_OH(19, 1, 2, 512, 1),  // 24, 25
    VU_TMA_OH(2, 4, 1, 512, 2),  VU_TMA_OH(2, 4, 2, 512, 2),   // 26, 27
    VU_TMA_OH(3, 4, 1, 256, 2),  VU_TMA_OH(3, 4, 2, 256, 2),   // 28, 29
    VU_TMA_OH(4, 4, 1, 256, 2),  VU_TMA_OH(4, 4, 2, 256, 2),   // 30, 31
};
static const int kNumTma = (int)(sizeof(kTma) / sizeof(kTma[0]));

// Returns VU_OK after launching, 1 if this slab is not eligible for the TMA path (caller falls through to
// the register-streaming kernels), or a negative vu_status.
int launch_k1_tma(const vu_fused_args* a, const StatParams& st, cudaStream_t stream) {
    const vu_slab& s = a->slab;
    const long long forced = get_option("k1_tma_variant", -1);
    if (get_option("k1_path", 0) == 1) return 1;  // 1 = register-streaming kernels only
    if (s.stride_v != 1 || s.P < 2 || s.P > 271) return 1;
    // bulk copies need 16-byte aligned rows and sizes
    if ((uintptr_t)s.data % 16 || s.V % 4 || (!s.member_ptrs && s.stride_p % 4) || s.stride_b % 4 || s.stride_c % 4) return 1;
    if (s.member_ptrs_host)
        for (int64_t p = 0; p < s.P; ++p)
            if ((uintptr_t)s.member_ptrs_host[p] % 16) return 1;
    const int need_levels = s.P <= 17 ? 1 : 2;
    // 0: probabilities as they are, 1: logits, 2: one-hot members; grouped draws, renormalisation and combinations: generic kernel
    const int lg = s.flags == VU_SLAB_LOGITS ? 1 : (s.flags == VU_SLAB_DISCRETIZE ? 2 : 0);
    if (s.draws > 1 || (s.flags && !lg)) return 1;
    const TmaVariant* pick = nullptr;
    if (forced >= 0 && forced < kNumTma) {
        const TmaVariant& f = kTma[forced];
        if (f.C != s.C || f.LEVELS < need_levels || f.logits != lg) return set_error(VU_ERR_UNSUPPORTED, "k1_tma_variant does not fit this slab");
        pick = &f;
    } else {
        for (int i = 0; i < kNumTma && !pick; ++i)
            if (kTma[i].C == s.C && kTma[i].LEVELS == need_levels && kTma[i].use == 0 && kTma[i].logits == lg) pick = &kTma[i];
    }
    if (!pick) return 1;
    // With few classes the reference-based statistics outweigh the streaming arithmetic; three statistics warps
    // cannot keep up with sixteen consumer warps there, so those launches stay on the register-streaming kernel,
    // where every warp does both (measured r01: cfg2 / cfg4 with Dice + calibration statistics 1.3-2.2x faster
    // that way; without reference-based statistics the TMA form wins everywhere).
    const unsigned heavy = VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC | VU_STAT_PLATT_FIT | VU_STAT_CLASS_COUNTS;
    if (forced < 0 && !lg && get_option("k1_path", 0) != 2 && (st.flags & heavy) && s.C * s.P < 128) return 1;
    const int vec = pick->VEC;
    auto ok = [&](const void* p, uintptr_t al) { return p == nullptr || ((uintptr_t)p % al) == 0; };
    if (!ok(a->tu, 4 * vec) || !ok(a->au, 4 * vec) || !ok(a->eu, 4 * vec) || !ok(a->labels, vec) || !ok(a->member_labels, vec) || s.V % vec) return 1;

    K1TmaParams prm;
    prm.x = s.data;
    prm.mptr = s.member_ptrs;
    prm.P = s.P; prm.B = s.B; prm.C = s.C; prm.V = s.V;
    prm.sp = s.stride_p; prm.sb = s.stride_b; prm.sc = s.stride_c;
    prm.tu = a->tu; prm.au = a->au; prm.eu = a->eu; prm.lab = a->labels;
    prm.mlab = a->member_labels;
    prm.st = st;
    const long long tile_vox = (long long)pick->CT * vec;
    prm.tiles_per_img = (s.V + tile_vox - 1) / tile_vox;
    prm.total_tiles = prm.tiles_per_img * s.B;
    if (prm.total_tiles >= (1LL << 31)) return set_error(VU_ERR_UNSUPPORTED, "more than 2^31 tiles in one launch; split the batch");

    const int rows = pick->NCH == 1 ? pick->G * pick->C : (pick->C + 1) / 2;
    const size_t stage_bytes = (size_t)rows * tile_vox * sizeof(float);
    const size_t stats_bytes = stats_smem_bytes(st.flags, st.gt.R, pick->ST, pick->SREP) + stats_class_bytes(st.flags, st.gt.R, st.ncls);
    const size_t hand_bytes = st.flags ? 2 * 13 * (size_t)tile_vox : 0;
    const size_t budget = 227 * 1024;
    const size_t fixed = 256 /* barriers */ + 256 /* alignment slack */ + stats_bytes + hand_bytes;
    long long nstages = get_option("k1_tma_stages", 0);
    const long long fit = fixed < budget ? (long long)((budget - fixed) / stage_bytes) : 0;
    if (nstages <= 0) nstages = fit < 6 ? fit : 6;
    if (nstages > fit) nstages = fit;
    if (nstages > 8) nstages = 8;
    if (nstages < 2) return 1;  // not enough shared memory for a pipeline: use the register-streaming kernel
    prm.nstages = (int)nstages;
    size_t off = (size_t)nstages * stage_bytes;
    prm.bar_offset = (unsigned)off;
    off = (off + 256 + 127) / 128 * 128;  // 2 x nstages + 4 mbarriers (<= 160 bytes)
    prm.hand_offset = (unsigned)off;
    off = (off + hand_bytes + 127) / 128 * 128;
    prm.stats_offset = (unsigned)off;
    const size_t dyn = off + stats_bytes;

    K1TmaKernel fn = st.flags ? pick->fn_stats : pick->fn;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(k1_tma)");
    long long grid = device_sm_count();
    if (grid > prm.total_tiles) grid = prm.total_tiles;
    const int threads = pick->CT + 32 + (st.flags ? pick->ST : 0);  // no statistics warps without statistics
    fn<<<(unsigned)grid, threads, dyn, stream>>>(prm);
    count_launch("k1_tma");
    return check_launch("k1_tma");
}

int num_tma_variants() { return kNumTma; }

}  // namespace vu
