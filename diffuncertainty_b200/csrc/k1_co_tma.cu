// K1, class-outer TMA form: ANY class count at run time (P <= PMAX members), rows staged through a shared-memory ring by bulk
// asynchronous copies.
//
// The direct-load class-outer kernel (k1_fused.cu) tops out near 60 % of the HBM peak whatever its occupancy or prefetch
// depth (r02x: 512-thread CTAs, a class-ahead register prefetch and a three-class cp.async ring were all measured): every
// warp-level load touches 256 bytes of a different row, and B200's HBM wants longer runs -- the forms that copy 2-8 KB rows
// with the TMA engine (k1_tma.cu, k1_uni.cu) reach 93-98 %.  Same here: a producer warp copies, per (tile, class), the P rows
// of that class (TV voxels each, one cp.async.bulk per row) into a ring stage; the consumer threads own VEC = 2 voxels each and
// do exactly what k1_classouter does -- class mean finished in registers, per-member entropy sums in registers (member loop
// unrolled to PMAX), same operations in the same order as every other form (bit-identical maps and labels).
// The (tile, class) pairs of a CTA form one flat sequence, so the ring never drains between tiles.  With statistics the
// consumer threads run the general statistics phase (vu_common.cuh) on their two voxels after the last class of a tile, on a
// named barrier of their own (the producer warp takes no part), as the register-streaming kernel does.
//
// ET: element type of the slab -- 0 float32, 1 bfloat16, 2 float16 (vu_slab.dtype).  16-bit rows are copied as they are (half
// the bytes) and widened to float32 by the consumers as they read the stage: exact, so the results are those of the upcast slab
// bit for bit.  This form is the one that reads 16-bit slabs, for every class count.  (Two CTAs per SM were measured for
// P <= 16: +5 % for C = 5 / 7, -4 % for C = 21, nothing for 16-bit slabs -- one CTA per SM it stays.)
//
// Reference semantics: see k1_fused.cu / k1_core.cuh.
#include "k1_core.cuh"
#include "tma_common.cuh"
#include "vu_host.h"
#include <type_traits>

namespace vu {

extern __shared__ __align__(128) unsigned char vu_co_smem[];

struct K1CoParams {
    const float* x;  // (16-bit elements when ET != 0; strides are in elements)
    const float* const* mptr;
    long long P, B, C, V;
    long long sp, sb, sc;
    float* tu;
    float* au;
    float* eu;
    uint8_t* lab;
    long long tiles_per_img, total_tiles;
    int nstages;
    unsigned bar_offset, stats_offset;
    StatParams st;
};

// two voxels of a 16-bit row -> packed float32 pair
template <int ET>
__device__ __forceinline__ f32x2 widen2(unsigned w) {
    if (ET == 1) return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));  // bfloat16: the upper half of a float32
    float lo, hi;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(lo), "=f"(hi) : "r"(w));
    return pk2(lo, hi);
}

template <int PMAX, int CT, bool STATS, int ET>
__global__ void __launch_bounds__(CT + 32, 1) k1_co_tma(const __grid_constant__ K1CoParams prm) {
    constexpr int VEC = 2, TV = CT * VEC;
    constexpr unsigned kEs = ET == 0 ? 4u : 2u;  // bytes per element
    constexpr unsigned kRowBytes = TV * kEs;
    const int P = (int)prm.P, C = (int)prm.C;
    const long long V = prm.V;
    const unsigned stage_bytes = (unsigned)P * kRowBytes;
    const int nstages = prm.nstages;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(vu_co_smem + prm.bar_offset);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + nstages);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, CT / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int t0 = (int)(prm.total_tiles * (long long)blockIdx.x / gridDim.x);
    const int t1 = (int)(prm.total_tiles * (long long)(blockIdx.x + 1) / gridDim.x);
    const int tpi = (int)prm.tiles_per_img;

    if (tid >= CT) {
        // ------------------------------ producer warp: P rows per (tile, class) ------------------------------------------
        const int lane = tid - CT;
        unsigned long long policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        int b = t0 / tpi, vt = t0 - b * tpi - 1;
        int stage = 0;
        unsigned phase = 0;
        for (int tile = t0; tile < t1; ++tile) {
            if (++vt == tpi) { vt = 0; ++b; }
            const long long v0 = (long long)vt * TV, left = V - v0;
            const unsigned row_bytes = left >= TV ? kRowBytes : (unsigned)(left * kEs);
            const long long off0 = (long long)b * prm.sb + v0;
            for (int c = 0; c < C; ++c) {
                mbar_wait_hint(empty0 + 8 * stage, phase ^ 1, 2000u);
                if (lane == 0) mbar_arrive_expect_tx(full0 + 8 * stage, (unsigned)P * row_bytes);
                __syncwarp();
                const unsigned dst0 = smem_u32(vu_co_smem) + (unsigned)stage * stage_bytes;
                for (int p = lane; p < P; p += 32) {
                    // (byte arithmetic: the elements are 2 or 4 bytes wide)
                    const char* mem = prm.mptr ? reinterpret_cast<const char*>(ld_member_ptr(prm.mptr, p))
                                               : reinterpret_cast<const char*>(prm.x) + (long long)p * prm.sp * kEs;
                    bulk_g2s(dst0 + (unsigned)p * kRowBytes, mem + (off0 + (long long)c * prm.sc) * kEs, row_bytes, full0 + 8 * stage, policy);
                }
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---------------------------------- consumer warps ---------------------------------------------------------------------
    const float Pf = (float)P;
    const bool two = P > 17;  // torch's cascade: chunks of 16 members folded into a second accumulator (k1_core.cuh)
    int b = t0 / tpi, vt = t0 - b * tpi - 1;
    int stage = 0;
    unsigned phase = 0;
    const unsigned my_ring = smem_u32(vu_co_smem) + (unsigned)tid * (VEC * kEs);
    void* st_smem = vu_co_smem + prm.stats_offset;
    StatsCursor<CT> cursor;
    if (STATS) stats_init<CT, 1, 0, 16>(prm.st, st_smem);
    for (int tile = t0; tile < t1; ++tile) {
        if (++vt == tpi) { vt = 0; ++b; }
        const long long v = (long long)vt * TV + (long long)tid * VEC;
        const bool active = v < V;
        if (STATS) {
            cursor.template enter<1, 0, 16>(prm.st, st_smem, b, vt, TV);
            if (active) stats_prefetch_gt<VEC>(prm.st, b, v);
        }
        f32x2 hp[PMAX];  // per-member entropy sums of the two voxels (log2 units)
#pragma unroll
        for (int p = 0; p < PMAX; ++p) hp[p] = 0ull;
        float best[VEC] = {0.f, 0.f}, tu2[VEC] = {0.f, 0.f};
        int label[VEC] = {0, 0};
        for (int c = 0; c < C; ++c) {
            mbar_wait(full0 + 8 * stage, phase);
            const unsigned src = my_ring + (unsigned)stage * stage_bytes;
            f32x2 M0 = 0ull, M1 = 0ull;
            // (a partial last tile leaves stale bytes behind the image's end: those threads are inactive, their results discarded)
#pragma unroll
            for (int p = 0; p < PMAX; ++p) {
                if (p < P) {
                    f32x2 X;
                    if constexpr (ET == 0) {
                        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(X) : "r"(src + (unsigned)p * kRowBytes));
                    } else {
                        unsigned w;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(src + (unsigned)p * kRowBytes));
                        X = widen2<ET>(w);
                    }
                    M0 = add2(M0, X);
                    f32x2 PC, L;
                    plog2p_parts2(X, PC, L);
                    hp[p] = fma2(PC, L, hp[p]);
                    if ((p & 15) == 15 && two) { M1 = add2(M1, M0); M0 = 0ull; }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * stage);  // this warp is done reading the stage
            if (++stage == nstages) { stage = 0; phase ^= 1; }
            float m0[VEC], m1[VEC];
            upk2(M0, m0[0], m0[1]);
            upk2(M1, m1[0], m1[1]);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float sum = two ? __fadd_rn(m0[k], m1[k]) : m0[k];
                const float mean = __fdiv_rn(sum, Pf);  // test_2D.py:971
                if (c == 0) { best[k] = mean; label[k] = 0; } else argmax_step(mean, c, best[k], label[k]);
                tu2[k] = plog2p_acc(tu2[k], mean);
            }
        }
        float u[VU_N_UNC][VEC] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
        if (active) {
            float h[PMAX][VEC];
#pragma unroll
            for (int p = 0; p < PMAX; ++p) upk2(hp[p], h[p][0], h[p][1]);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                // AU: the mean over the members of their entropies, in the same cascade order
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int p = 0; p < PMAX; ++p) {
                    if (p < P) {
                        a0 = __fadd_rn(a0, h[p][k]);
                        if ((p & 15) == 15 && two) { a1 = __fadd_rn(a1, a0); a0 = 0.f; }
                    }
                }
                const float asum = two ? __fadd_rn(a0, a1) : a0;
                const float tu = -(tu2[k] * kLn2);
                const float au = __fdiv_rn(-(asum * kLn2), Pf);
                u[0][k] = tu; u[1][k] = au; u[2][k] = tu - au;
            }
            const long long o = (long long)b * V + v;
            if (prm.tu) VecLoad<VEC>::store(prm.tu + o, u[0]);
            if (prm.au) VecLoad<VEC>::store(prm.au + o, u[1]);
            if (prm.eu) VecLoad<VEC>::store(prm.eu + o, u[2]);
            if (prm.lab) VecLoad<VEC>::store_u8(prm.lab + o, label);
        } else {
            label[0] = label[1] = 0;
        }
        if (STATS) stats_tile<VEC, CT, 0, 16>(prm.st, st_smem, active, b, v, u, label);
    }
    if (STATS && t1 > t0) cursor.template finish<1, 0, 16>(prm.st, st_smem, vt, TV);
}

// Returns VU_OK after launching, 1 if this launch is not one for this form (caller goes on), or a negative vu_status.
int launch_k1_co_tma(const vu_fused_args* a, const StatParams& st, cudaStream_t stream) {
    const vu_slab& s = a->slab;
    if (get_option("k1_path", 0) == 1 || get_option("k1_variant", -1) != -1) return 1;
    if (a->member_labels || s.draws > 1 || s.flags) return 1;
    if (s.stride_v != 1 || s.P < 2 || s.P > 32) return 1;
    const int et = s.dtype;  // 0 float32, 1 bfloat16, 2 float16
    const int es = et == VU_SLAB_F32 ? 4 : 2, per16 = 16 / es;
    // bulk copies need 16-byte aligned rows and sizes
    if ((uintptr_t)s.data % 16 || s.V % per16 || (!s.member_ptrs && s.stride_p % per16) || s.stride_b % per16 || s.stride_c % per16) return 1;
    if (s.member_ptrs_host)
        for (int64_t p = 0; p < s.P; ++p)
            if ((uintptr_t)s.member_ptrs_host[p] % 16) return 1;
    auto ok = [&](const void* p, uintptr_t al) { return p == nullptr || ((uintptr_t)p % al) == 0; };
    if (!ok(a->tu, 8) || !ok(a->au, 8) || !ok(a->eu, 8) || !ok(a->labels, 2)) return 1;

    typedef void (*Fn)(const K1CoParams);
    // 512 consumer threads (4 KB rows) while the registers allow it, 256 (2 KB rows) for up to 32 members and with statistics
    // (whose per-thread columns and per-warp histograms share the shared memory with the ring)
    const bool stats = st.flags != 0;
    const int ct = (s.P <= 16 && !stats) ? 512 : 256;
    auto pick = [&](auto et_tag) -> Fn {
        constexpr int ET = decltype(et_tag)::value;
        return stats ? (s.P <= 8 ? (Fn)k1_co_tma<8, 256, true, ET> : (s.P <= 16 ? (Fn)k1_co_tma<16, 256, true, ET> : (Fn)k1_co_tma<32, 256, true, ET>))
                     : (s.P <= 8 ? (Fn)k1_co_tma<8, 512, false, ET> : (s.P <= 16 ? (Fn)k1_co_tma<16, 512, false, ET> : (Fn)k1_co_tma<32, 256, false, ET>));
    };
    Fn fn = et == VU_SLAB_F32 ? pick(std::integral_constant<int, 0>()) : (et == VU_SLAB_BF16 ? pick(std::integral_constant<int, 1>()) : pick(std::integral_constant<int, 2>()));
    K1CoParams prm;
    prm.st = st;
    prm.x = s.data; prm.mptr = s.member_ptrs;
    prm.P = s.P; prm.B = s.B; prm.C = s.C; prm.V = s.V;
    prm.sp = s.stride_p; prm.sb = s.stride_b; prm.sc = s.stride_c;
    prm.tu = a->tu; prm.au = a->au; prm.eu = a->eu; prm.lab = a->labels;
    const long long tile_vox = (long long)ct * 2;
    prm.tiles_per_img = (s.V + tile_vox - 1) / tile_vox;
    prm.total_tiles = prm.tiles_per_img * s.B;
    if (prm.total_tiles >= (1LL << 31)) return set_error(VU_ERR_UNSUPPORTED, "more than 2^31 tiles in one launch; split the batch");
    const size_t stage_bytes = (size_t)s.P * tile_vox * es;
    const size_t stats_bytes = stats ? (stats_smem_bytes(st.flags, st.gt.R, ct) + stats_class_bytes(st.flags, st.gt.R, st.ncls) + 127) / 128 * 128 : 0;
    const size_t budget = 225 * 1024 - 512;
    if (stats_bytes + 2 * stage_bytes > budget) return 1;
    long long nstages = (long long)((budget - stats_bytes) / stage_bytes);
    if (nstages > 8) nstages = 8;
    if (nstages < 2) return 1;
    prm.nstages = (int)nstages;
    prm.bar_offset = (unsigned)(nstages * stage_bytes);
    prm.stats_offset = (unsigned)(nstages * stage_bytes + 256);
    const size_t dyn = nstages * stage_bytes + 256 + stats_bytes;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(k1_co_tma)");
    long long grid = device_sm_count();
    if (grid > prm.total_tiles) grid = prm.total_tiles;
    fn<<<(unsigned)grid, ct + 32, dyn, stream>>>(prm);
    count_launch("k1_co_tma");
    return check_launch("k1_co_tma");
}

}  // namespace vu
