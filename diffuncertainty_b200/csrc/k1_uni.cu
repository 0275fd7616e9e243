// K1, unified-warp TMA form for binary slabs (C == 2): every consumer warp does the streaming arithmetic AND the statistics of
// its own voxels.
//
// Why a third form: for few-class slabs with reference-based statistics (configs[1]: N = 5, C = 2, 4 raters, calibration
// histograms; configs[3]: N = 32, NCC) the statistics outweigh the streaming arithmetic.  In the warp-specialised form
// (k1_tma.cu) the split between consumer and statistics warps is fixed at compile time, so one of the two groups idles
// (r02a, configs[1]: 16 + 3 warps 1.81 ms, 8 + 8 warps 0.71 ms, 8 + 12 warps 0.80 ms -- never balanced); in the
// register-streaming form (k1_fused.cu) a warp has no loads in flight while it computes (0.81 ms).  Here one producer warp
// keeps the shared-memory ring full with cp.async.bulk copies whatever the other warps do, and each of the CT / 32 consumer
// warps alternates between consuming a tile and running the lean statistics phase (stats_v2.cuh) on the four voxels per
// thread it has just finished, straight from registers: no hand-off buffer, no idle warps, any ratio of the two kinds of
// work.  The ring (>= 96 KB) rides out the statistics phases of the warps.  Partial sums stay in registers, histograms are
// private to a warp, and a warp flushes on its own when it moves to another image: no CTA-wide barrier after start-up.
//
// Reference semantics: see k1_fused.cu / k1_core.cuh (maps, labels) and stats_v2.cuh (statistics).
#include "k1_core.cuh"
#include "members_fold.cuh"
#include "stats_v2.cuh"
#include "tma_common.cuh"
#include <math.h>
#include <string.h>
#include <type_traits>

#include "vu_host.h"

namespace vu {

extern __shared__ __align__(128) unsigned char vu_uni_smem[];

struct K1UniParams {
    const float* x;
    const float* const* mptr;  // optional device array of P member base pointers (then x / sp are unused)
    long long P, B, V;
    long long sp, sb, sc;
    float* tu;
    float* au;
    float* eu;
    uint8_t* lab;
    long long tiles_per_img, total_tiles;
    int nstages;
    unsigned bar_offset;    // byte offsets inside dynamic shared memory
    unsigned stats_offset;
    unsigned ms_offset;     // per-warp state of the member-level scores (MS kernels)
    StatParams st;
    MsParams ms;
};

constexpr int kUniRep = 16;  // histogram replicas per warp (two lanes share one; the half-warps walk the types in rotated order)

// LEVELS: cascade levels of the member sum (1: P <= 17, 2: P <= 271); CT consumer threads; G members per ring stage;
// FL: the statistics mask (compile time); RMAX: raters the reference registers are sized for.
// MINB: CTAs per SM (each with its own producer warp, ring and histograms: independent pipelines that fill each other's gaps).
// MSS >= 0: the member-level scores (GED counts, likelihood sums; members_fold.cuh) are computed in the same pass, with MSS
// shuffle steps per member pair before the partial sums go to shared memory (MsGeom); MSS < 0: no member scores.
// LG: the slab holds logits (VU_SLAB_LOGITS): every member is softmax'ed over its two classes as it is consumed.
// OH: every member is replaced by the one-hot vector of its argmax (VU_SLAB_DISCRETIZE, --discretize).
template <int LEVELS, int CT, int G, unsigned FL, int RMAX, int MINB, int MSS, bool LG = false, bool OH = false>
__global__ void __launch_bounds__(CT + 32, MINB) k1_uni(const __grid_constant__ K1UniParams prm) {
    constexpr bool MS = MSS >= 0;
    static_assert(!(MS && (LG || OH)) && !(LG && OH), "member scores take the members as they are; one producer per kernel");
    constexpr int kMsS = MS ? MSS : 3;
    static_assert(!MS || RMAX == kMsR, "the member-score fold is built for up to four raters");
    constexpr int C = 2, VEC = 4;
    constexpr int TV = CT * VEC;  // voxels per tile
    constexpr unsigned kRowBytes = TV * sizeof(float);
    constexpr unsigned kStageBytes = G * C * kRowBytes;
    constexpr int kStageFloats = kStageBytes / sizeof(float);
    constexpr int kWarps = CT / 32;
    using Acc = VoxelAcc<C, VEC, LEVELS>;

    const int nstages = prm.nstages;
    float* ring = reinterpret_cast<float*>(vu_uni_smem);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(vu_uni_smem + prm.bar_offset);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + nstages);
    void* st_smem = vu_uni_smem + prm.stats_offset;
    const StatParams& sp = prm.st;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(full0 + 8 * s, 1);       // the producer's arrive.expect_tx
            mbar_init(empty0 + 8 * s, kWarps);  // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stats2_init<kUniRep>(sp, st_smem, tid, CT + 32, kWarps);
    __syncthreads();

    const long long V = prm.V;
    const int P = (int)prm.P;
    const int t0 = (int)(prm.total_tiles * (long long)blockIdx.x / gridDim.x);
    const int t1 = (int)(prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x);
    const int tpi = (int)prm.tiles_per_img;
    const int fills = (P + G - 1) / G;  // stage fills per tile
    const int full_fills = P / G, rem = P - full_fills * G;  // ... of which full ones, and the members in the last one otherwise

    if (tid >= CT) {
        // ------------------------------ producer warp ---------------------------------------------------
        const int lane = tid - CT;
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
                const int p0 = fi * G;
                const int nrows = ((P - p0) < G ? (int)(P - p0) : G) * C;
                mbar_wait_hint(empty0 + 8 * stage, phase ^ 1, 4000u);  // slot free (the first pass falls through)
                if (lane == 0) mbar_arrive_expect_tx(full0 + 8 * stage, (unsigned)nrows * row_bytes);
                __syncwarp();
                const unsigned dst0 = smem_u32(ring) + (unsigned)stage * kStageBytes;
                for (int r = lane; r < nrows; r += 32) {
                    const int g = r / C, c = r - g * C;
                    const float* mem = prm.mptr ? ld_member_ptr(prm.mptr, p0 + g) : prm.x + (long long)(p0 + g) * prm.sp;
                    bulk_g2s(dst0 + (unsigned)r * kRowBytes, mem + off0 + (long long)c * prm.sc, row_bytes, full0 + 8 * stage, policy);
                }
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---------------------------------- consumer warps: streaming + statistics --------------------------------------------
    const int warp = tid >> 5;
    const float Pf = (float)P;
    Stat2Ctx cx;
    stats2_ctx<kUniRep, (CT + 32 <= 512)>(cx, sp, st_smem, warp, kWarps);  // more than 16 warps: 96 registers per thread
    const bool rot = stats2_rotated<kUniRep>();
    StatAcc<FL, RMAX> A;
    A.clear();
    int cur_b = -1, vt_begin = 0;
    MsWarp<kMsS> mw;
    const bool ms_nll = MS && (prm.ms.flags & VU_MS_NLL), ms_ged = MS && (prm.ms.flags & VU_MS_GED);
    const bool ign_is_one = sp.gt.has_ignore && sp.gt.ignore == 1;
    if constexpr (MS) mw.init(vu_uni_smem + prm.ms_offset + (unsigned)warp * MsWarp<kMsS>::kWarpBytes);
    auto flush = [&]() {
        stats2_flush_regs<FL, RMAX, kUniRep>(A, sp, cur_b);
        if (FL & VU_STAT_CALIB) stats2_flush_hist_warp<kUniRep>(sp, st_smem, cur_b, warp);
        if constexpr (MS) mw.flush(prm.ms, cur_b, (int)P, sp.gt.R, true);
    };

    int b = t0 / tpi, vt = t0 - b * tpi - 1;
    unsigned stage_off = 0u;  // byte offset of the current stage inside the ring
    unsigned full_bar = full0, empty_bar = empty0;
    unsigned phase = 0;
    const unsigned ring_end = (unsigned)nstages * kStageBytes;
    unsigned my_ring = smem_u32(ring) + (unsigned)tid * (VEC * (unsigned)sizeof(float));
    asm volatile("" : "+r"(my_ring));  // keep it in a register (rematerialised per stage otherwise: S2UR + ULEA + LEA)
    long long img_out = 0;       // b * V: offset of the image in the (B, V) outputs
    const uint8_t* gt_img = nullptr;  // references of the image
    unsigned Wn[RMAX];           // MS: the reference words of the next tile
    if constexpr (MS) if (t0 < t1) {
        const int b0 = t0 / tpi, vt0 = t0 - b0 * tpi;
        const long long v0 = (long long)vt0 * TV + (long long)tid * VEC;
        stats2_load_refs<FL | VU_STAT_DICE, RMAX>(sp, v0 < V, reinterpret_cast<const uint8_t*>(sp.gt.data) + (long long)b0 * sp.gt.sb, v0, Wn);
    }
    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        if (b != cur_b || vt - vt_begin >= kMaxTilesPerFlush2) {  // warp-uniform
            if (cur_b >= 0) flush();
            cur_b = b;
            vt_begin = vt;
            img_out = (long long)b * V;
            gt_img = reinterpret_cast<const uint8_t*>(sp.gt.data) + (long long)b * sp.gt.sb;
        }
        const long long v = (long long)vt * TV + (long long)tid * VEC;
        const bool active = v < V;
        unsigned W[RMAX];
        if constexpr (MS) {
            // the reference words are needed before the first member (likelihood masks): they were requested during the
            // previous tile; request the next tile's now
#pragma unroll
            for (int r = 0; r < RMAX; ++r) W[r] = Wn[r];
            if (tile + 1 < t1) {
                const bool wrap = vt + 1 == tpi;
                const long long nv = (long long)(wrap ? 0 : vt + 1) * TV + (long long)tid * VEC;
                const uint8_t* ngt = wrap ? gt_img + sp.gt.sb : gt_img;
                stats2_load_refs<FL | VU_STAT_DICE, RMAX>(sp, nv < V, ngt, nv, Wn);
            }
            mw.tile_begin(W, active, sp.gt, ms_nll);
        } else {
            stats2_load_refs<FL, RMAX>(sp, active, gt_img, v, W);  // in flight while the members stream
        }

        Acc acc;
        acc.init();
        // One ring stage = G members.  FULL: all G are there (no per-member test); otherwise the first `rem` (the last stage of a
        // tile when G does not divide P).
        auto consume_stage = [&](auto full_tag, const int p0) {
            constexpr bool FULL = decltype(full_tag)::value;
            mbar_wait(full_bar, phase);  // the bytes of this stage have landed
            const unsigned sbase = my_ring + stage_off;
            // (a partial last tile leaves stale bytes behind the image's end: those threads are inactive and
            //  their arithmetic is discarded)
            if constexpr (MS) {
                // two members per stage: their likelihood values are folded over the lanes together
                static_assert(!MS || G == 2, "the member-score fold takes the members in pairs");
                f32x2 xa[Acc::NP], xb[Acc::NP], La[Acc::NP], Lb[Acc::NP];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xa[c * 2]), "=l"(xa[c * 2 + 1]) : "r"(sbase + (unsigned)(c * kRowBytes)));
                    if constexpr (FULL)
                        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xb[c * 2]), "=l"(xb[c * 2 + 1]) : "r"(sbase + (unsigned)((C + c) * kRowBytes)));
                }
                if constexpr (!FULL) {  // odd member count: the second half of the stage holds no member
#pragma unroll
                    for (int i = 0; i < Acc::NP; ++i) xb[i] = 0ull;
                }
                // one member after the other (the registers do not hold both members' logarithms); only the fold over the lanes
                // is shared
                const bool slow = ms_nll && (!mw.fast || __any_sync(kFull, (MsWarp<kMsS>::nan_probe(xa) + MsWarp<kMsS>::nan_probe(xb)) != 0.0f));
                float va[kMsVals], vb[kMsVals];
#pragma unroll
                for (int i = 0; i < kMsVals; ++i) { va[i] = 0.f; vb[i] = 0.f; }
                acc.add_member_L(xa, p0, false, La);
                if (ms_ged) mw.member_labels(p0, xa);
                if (ms_nll) mw.member_values(xa, La, sp.gt.R, prm.ms.log2eps, slow, va);
                if constexpr (FULL) {
                    acc.add_member_L(xb, p0 + 1, false, Lb);
                    if (ms_ged) mw.member_labels(p0 + 1, xb);
                    if (ms_nll) mw.member_values(xb, Lb, sp.gt.R, prm.ms.log2eps, slow, vb);
                }
                if (ms_nll) mw.fold_pair(p0, va, vb);
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (FULL || g < rem) {
                        f32x2 xp[Acc::NP];
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xp[c * 2]), "=l"(xp[c * 2 + 1]) : "r"(sbase + (unsigned)((g * C + c) * kRowBytes)));
                        if constexpr (OH) {
                            acc.add_member_onehot(xp, 0.f, p0 + g);
                        } else if constexpr (LG) {
                            f32x2 hm[Acc::NH], rs[Acc::NH];
                            float xs = 0.f, hs = 0.f;
                            auto reload = [&](int i) {
                                f32x2 r;
                                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(sbase + (unsigned)((g * C + (i >> 1)) * kRowBytes + (i & 1) * 8)));
                                return r;
                            };
                            acc.softmax_member(xp, xs, rs, hm, hs, reload, []() { return 0.f; });
                            acc.add_member_pre(xp, 0.f, rs, hm, hs, p0 + g, false);
                        } else {
                            acc.add_member(xp, 0.f, p0 + g, false);
                        }
                    }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(empty_bar);  // this warp is done reading the stage
            stage_off += kStageBytes; full_bar += 8; empty_bar += 8;
            if (stage_off == ring_end) { stage_off = 0u; full_bar = full0; empty_bar = empty0; phase ^= 1; }
        };
        for (int fi = 0; fi < full_fills; ++fi) consume_stage(std::true_type{}, fi * G);
        if (rem) consume_stage(std::false_type{}, full_fills * G);

        float u[VU_N_UNC][VEC];
        int label[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { u[0][k] = u[1][k] = u[2][k] = 0.f; label[k] = 0; }
        if (active) {
            acc.finish(Pf, u, label);
            const long long o = img_out + v;
            if (prm.tu) VecLoad<VEC>::store(prm.tu + o, u[0]);
            if (prm.au) VecLoad<VEC>::store(prm.au + o, u[1]);
            if (prm.eu) VecLoad<VEC>::store(prm.eu + o, u[2]);
            if (prm.lab) VecLoad<VEC>::store_u8(prm.lab + o, label);
        }
        // statistics of this thread's four voxels, in the lane's step order (stats_v2.cuh, "type rotation")
        const unsigned lab4 = (unsigned)label[0] | ((unsigned)label[1] << 8) | ((unsigned)label[2] << 16) | ((unsigned)label[3] << 24);
        float4 U[VU_N_UNC];
#pragma unroll
        for (int s = 0; s < VU_N_UNC; ++s) {
            const int k1 = (s + 1) % VU_N_UNC;
            U[s] = make_float4(rot ? u[k1][0] : u[s][0], rot ? u[k1][1] : u[s][1], rot ? u[k1][2] : u[s][2], rot ? u[k1][3] : u[s][3]);
        }
        stats2_tile<FL, RMAX, kUniRep>(A, sp, cx, active, b, U[0], U[1], U[2], lab4, W);
        if constexpr (MS) if (ms_ged) {
            const unsigned lab1 = (label[0] == 1 ? 1u : 0u) | (label[1] == 1 ? 2u : 0u) | (label[2] == 1 ? 4u : 0u) | (label[3] == 1 ? 8u : 0u);
            mw.tile_end((int)P, sp.gt.R, lab1, sp.gt.has_ignore != 0, ign_is_one, true);
        }
    }
    if (cur_b >= 0) flush();
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef void (*K1UniKernel)(const K1UniParams);
struct UniVariant {
    int LEVELS, CT, G;
    unsigned FL;
    int RMAX, MINB;
    int use;  // 1: automatic selection, 0: only through the "k1_uni_shape" option (tuning sweep)
    int ms;   // 1: computes the member-level scores as well
    int mss;  // ... with this many shuffle steps per member pair (MsGeom)
    K1UniKernel fn;
    K1UniKernel fn_logits;  // the same launch over a slab of logits (NULL: not built)
    K1UniKernel fn_onehot;  // ... with the members replaced by their one-hot argmax (NULL: not built)
};
#define VU_UNI(LEVELS, CT, G, FL, RMAX, MINB, USE) \
    { LEVELS, CT, G, FL, RMAX, MINB, USE, 0, -1, (K1UniKernel)k1_uni<LEVELS, CT, G, FL, RMAX, MINB, -1>, nullptr, nullptr }
#define VU_UNI_L(LEVELS, CT, G, FL, RMAX, MINB, USE)                                                    \
    { LEVELS, CT, G, FL, RMAX, MINB, USE, 0, -1, (K1UniKernel)k1_uni<LEVELS, CT, G, FL, RMAX, MINB, -1>, \
      (K1UniKernel)k1_uni<LEVELS, CT, G, FL, RMAX, MINB, -1, true>, (K1UniKernel)k1_uni<LEVELS, CT, G, FL, RMAX, MINB, -1, false, true> }
#define VU_UNI_MS(LEVELS, CT, FL, MSS, USE) { LEVELS, CT, 2, FL, 4, 1, USE, 1, MSS, (K1UniKernel)k1_uni<LEVELS, CT, 2, FL, 4, 1, MSS>, nullptr, nullptr }
// Registers are allocated per SM sub-partition: 16 consumer warps + the producer put 5 warps on one of them (96 registers per
// thread), 15 + 1 leave 4 on each (128).  The masks with calibration histograms need the 128 (they spill 260-720 bytes at
// 96); the others do not, and keep the power-of-two tile.
#define VU_UNI_MASKS(LEVELS, G)                                                                                                              \
    VU_UNI_L(LEVELS, 512, G, 0x0du, 4, 1, 1), VU_UNI_L(LEVELS, 512, G, 0x0fu, 4, 1, 1), VU_UNI_L(LEVELS, 480, G, 0x1du, 4, 1, 1),             \
    VU_UNI_L(LEVELS, 512, G, 0x2du, 4, 1, 1),                                                                                                \
    VU_UNI_L(LEVELS, 480, G, 0x1fu, 4, 1, 1), VU_UNI_L(LEVELS, 512, G, 0x21u, 4, 1, 1), VU_UNI_L(LEVELS, 480, G, 0x3fu, 4, 1, 1),             \
    VU_UNI(LEVELS, 480, G, 0x3fu, 8, 1, 1)

// Two members per ring stage: a stage costs ~20 instructions of waiting, releasing and bookkeeping per warp whatever it
// holds (r02e, configs[1]: one member per stage 0.465 ms, two 0.450 ms; two CTAs of 7 + 1 warps per SM 0.53 ms, 11 + 1 warps
// 0.57 ms, 16 + 1 or 19 + 1 warps at 96 registers 0.61-0.64 ms -- they spill).
static const UniVariant kUni[] = {
    VU_UNI_MASKS(1, 2),
    VU_UNI_MASKS(2, 2),
    // tuning candidates ("k1_uni_shape" = CT * 100 + G * 10 + MINB)
    VU_UNI(1, 480, 1, 0x1du, 4, 1, 0), VU_UNI(1, 480, 3, 0x1du, 4, 1, 0),
    // with the member-level scores (GED counts + likelihood sums) in the same pass: 15 + 1 warps (128 registers)
    VU_UNI_MS(1, 480, 0x21u, 3, 1), VU_UNI_MS(2, 480, 0x21u, 3, 1), VU_UNI_MS(1, 480, 0x2du, 3, 1), VU_UNI_MS(2, 480, 0x2du, 3, 1),
    VU_UNI_MS(1, 480, 0x01u, 3, 1), VU_UNI_MS(2, 480, 0x01u, 3, 1),
    // (r02k, configs[3] pipeline: 15 warps with three shuffle steps 2.08 ms; 8 warps without shuffles (166 registers) 2.32 ms,
    //  12 warps with one or two steps 2.28 ms, 10 warps with one step 2.25 ms -- the warp count matters more than the shuffles)
};
static const int kNumUni = (int)(sizeof(kUni) / sizeof(kUni[0]));

bool stats2_eligible(const StatParams& st, long long V);  // k3_map_stats.cu

// Returns VU_OK after launching, 1 if this launch is not one for the unified form (the caller goes on to the
// warp-specialised / register-streaming kernels), or a negative vu_status.
int launch_k1_uni(const vu_fused_args* a, const StatParams& st, cudaStream_t stream, bool dry_run) {
    const vu_slab& s = a->slab;
    const bool want_ms = a->members.flags != 0;
    auto no = [&](const char* why) { return want_ms ? set_error(VU_ERR_UNSUPPORTED, why) : 1; };
    const long long path = get_option("k1_path", 0);
    if (path == 1 || path == 2) return no("k1_path option excludes the unified kernel");  // 1 = register-streaming kernels only, 2 = warp-specialised TMA kernels only
    if (get_option("k1_tma_variant", -1) >= 0 || get_option("stats_path", 0) == 1) return no("tuning options exclude the unified kernel");
    const bool lg = s.flags == VU_SLAB_LOGITS, oh = s.flags == VU_SLAB_DISCRETIZE;
    if ((s.flags && !lg && !oh) || s.draws > 1) return no("the unified kernel takes plain members, logits or one-hot members (no draws / renormalisation)");
    if ((lg || oh) && want_ms) return no("member scores take the members as they are");
    if (s.C != 2 || s.stride_v != 1 || s.P < 2 || s.P > 271 || !st.flags) return no("member scores in the fused pass need C == 2, unit voxel stride, a statistics mask");
    const unsigned heavy = VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC;
    if (!want_ms && !(st.flags & heavy) && path != 3) return 1;  // sums / thresholds / area only: the warp-specialised form is at the HBM roofline
    if (!stats2_eligible(st, s.V)) return no("statistics not eligible for the lean form (uint8 word-aligned references, V % 4 == 0, no LUT)");
    if (want_ms) {
        if (s.P > kMsP) return no("member scores in the fused pass need P <= 32");
        if (!st.gt.data || st.gt.dtype != VU_GT_U8 || st.gt.align < 4 || st.gt.R > kMsR)
            return no("member scores in the fused pass need at most 4 uint8 references with word-aligned rows");
        if ((a->members.flags & ~(VU_MS_NLL | VU_MS_GED))) return set_error(VU_ERR_BAD_ARG, "members.flags");
        if ((a->members.flags & VU_MS_NLL) && (!a->members.nll_sum || !a->members.nll_count || !a->members.nll_bad))
            return set_error(VU_ERR_BAD_ARG, "members: NLL outputs are NULL");
        if ((a->members.flags & VU_MS_GED) && !a->members.ged_counts) return set_error(VU_ERR_BAD_ARG, "members.ged_counts is NULL");
    }
    if (a->member_labels) return no("per-member labels are written by the warp-specialised / register-streaming kernels only");
    // bulk copies need 16-byte aligned rows and sizes
    if ((uintptr_t)s.data % 16 || s.V % 4 || (!s.member_ptrs && s.stride_p % 4) || s.stride_b % 4 || s.stride_c % 4) return no("rows are not 16-byte aligned");
    if (s.member_ptrs_host)
        for (int64_t p = 0; p < s.P; ++p)
            if ((uintptr_t)s.member_ptrs_host[p] % 16) return no("rows are not 16-byte aligned");
    auto ok = [&](const void* p, uintptr_t al) { return p == nullptr || ((uintptr_t)p % al) == 0; };
    if (!ok(a->tu, 16) || !ok(a->au, 16) || !ok(a->eu, 16) || !ok(a->labels, 4)) return no("outputs are not 16-byte aligned");
    const int need_levels = s.P <= 17 ? 1 : 2;
    const int rmax = (st.flags & heavy) && st.gt.R > 4 ? 8 : 4;
    const UniVariant* pick = nullptr;
    const long long shape = get_option("k1_uni_shape", 0);
    for (int i = 0; i < kNumUni && !pick; ++i) {
        const UniVariant& u = kUni[i];
        if (u.LEVELS != need_levels || u.FL != st.flags || u.RMAX < rmax || (u.ms != 0) != want_ms || (lg && !u.fn_logits) || (oh && !u.fn_onehot)) continue;
        const long long ushape = u.ms ? u.CT * 100 + u.G * 10 + u.mss : u.CT * 100 + u.G * 10 + u.MINB;
        if (shape ? ushape == shape : u.use == 1) pick = &u;
    }
    if (!pick) return shape ? set_error(VU_ERR_UNSUPPORTED, "k1_uni_shape: no such kernel for this launch")
                            : no("no unified kernel is built for this statistics mask");
    if (dry_run) return VU_OK;

    K1UniParams prm;
    prm.x = s.data;
    prm.mptr = s.member_ptrs;
    prm.P = s.P; prm.B = s.B; prm.V = s.V;
    prm.sp = s.stride_p; prm.sb = s.stride_b; prm.sc = s.stride_c;
    prm.tu = a->tu; prm.au = a->au; prm.eu = a->eu; prm.lab = a->labels;
    prm.st = st;
    memset(&prm.ms, 0, sizeof(prm.ms));
    if (want_ms) {
        prm.ms.flags = a->members.flags;
        prm.ms.log2eps = (float)log2((double)a->members.eps > 0.0 ? (double)a->members.eps : 1e-45);
        prm.ms.nll_sum = a->members.nll_sum;
        prm.ms.nll_cnt = reinterpret_cast<unsigned long long*>(a->members.nll_count);
        prm.ms.nll_bad = reinterpret_cast<unsigned long long*>(a->members.nll_bad);
        prm.ms.ged = reinterpret_cast<unsigned long long*>(a->members.ged_counts);
        prm.ms.ged_cols = (int)vu_ged_cols((int)s.P, st.gt.R);
    }
    const long long tile_vox = (long long)pick->CT * 4;
    prm.tiles_per_img = (s.V + tile_vox - 1) / tile_vox;
    prm.total_tiles = prm.tiles_per_img * s.B;
    if (prm.total_tiles >= (1LL << 31)) return set_error(VU_ERR_UNSUPPORTED, "more than 2^31 tiles in one launch; split the batch");

    const size_t stage_bytes = (size_t)pick->G * 2 * tile_vox * sizeof(float);
    const size_t ms_warp = !want_ms ? 0 : (pick->mss == 0 ? ms_warp_bytes<0>() : pick->mss == 1 ? ms_warp_bytes<1>() : pick->mss == 2 ? ms_warp_bytes<2>() : ms_warp_bytes<3>());
    const size_t stats_bytes = (stats2_smem_bytes(st.flags, pick->CT, kUniRep) + 15) / 16 * 16 + (size_t)(pick->CT / 32) * ms_warp;
    const size_t budget = (pick->MINB == 1 ? 227 : (pick->MINB == 2 ? 113 : 75)) * 1024 - (pick->MINB > 1 ? 1024 : 0);  // per CTA (1 KB reserved per CTA)
    const size_t fixed = 256 /* barriers */ + 256 /* alignment slack */ + stats_bytes;
    long long nstages = get_option("k1_tma_stages", 0);
    const long long fit = fixed < budget ? (long long)((budget - fixed) / stage_bytes) : 0;
    if (nstages <= 0) nstages = fit < 8 ? fit : 8;
    if (nstages > fit) nstages = fit;
    if (nstages > 8) nstages = 8;
    if (nstages < 2) return 1;
    prm.nstages = (int)nstages;
    size_t off = (size_t)nstages * stage_bytes;
    prm.bar_offset = (unsigned)off;
    off = (off + 256 + 127) / 128 * 128;  // 2 x nstages mbarriers (<= 128 bytes)
    prm.stats_offset = (unsigned)off;
    prm.ms_offset = (unsigned)(off + (stats2_smem_bytes(st.flags, pick->CT, kUniRep) + 15) / 16 * 16);
    const size_t dyn = off + stats_bytes;

    const K1UniKernel fn = lg ? pick->fn_logits : (oh ? pick->fn_onehot : pick->fn);
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(k1_uni)");
    long long grid = (long long)device_sm_count() * pick->MINB;
    if (grid > prm.total_tiles) grid = prm.total_tiles;
    fn<<<(unsigned)grid, pick->CT + 32, dyn, stream>>>(prm);
    count_launch("k1_uni");
    return check_launch("k1_uni");
}

}  // namespace vu
