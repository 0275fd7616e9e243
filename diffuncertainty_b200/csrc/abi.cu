// extern "C" boundary of libvalunc.so (see include/valunc.h).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#include "vu_common.cuh"
#include "vu_host.h"

namespace vu {

static thread_local char g_err[512] = "";
static std::mutex g_mu;
static std::map<std::string, long long> g_options;
static std::map<std::string, long long> g_counters;

int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int set_cuda_error(const char* where) {
    cudaError_t e = cudaGetLastError();
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return VU_ERR_CUDA;
}
int check_launch(const char* kernel) {
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) return VU_OK;
    return set_cuda_error(kernel);
}
void count_launch(const char* kernel) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_counters["launches"] += 1;
    g_counters[std::string("launches.") + kernel] += 1;
}
long long get_option(const char* key, long long dflt) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_options.find(key);
    return it == g_options.end() ? dflt : it->second;
}
int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    return cached[dev] > 0 ? cached[dev] : 148;
}

// largest power of two <= 4 such that `align` consecutive references can be fetched with one vector load
static int gt_alignment(const vu_gt& gt, long long V) {
    if (gt.stride_v != 1) return 1;
    const long long esz = gt.dtype == VU_GT_U8 ? 1 : 8;
    for (int a = 4; a > 1; a >>= 1) {
        const long long bytes = a * esz > 16 ? 16 : a * esz;
        if ((uintptr_t)gt.data % bytes == 0 && gt.stride_b % a == 0 && gt.stride_r % a == 0 && V % a == 0) return a;
    }
    return 1;
}

static int make_gt_view(GtView& v, const vu_gt* gt, long long V) {
    memset(&v, 0, sizeof(v));
    if (!gt || !gt->data) return VU_OK;
    if (gt->R < 1 || gt->R > VU_MAX_RATERS) return set_error(VU_ERR_UNSUPPORTED, "gt.R must be 1..8");
    if (gt->dtype != VU_GT_U8 && gt->dtype != VU_GT_I64) return set_error(VU_ERR_BAD_ARG, "gt.dtype");
    v.data = gt->data; v.dtype = gt->dtype; v.R = gt->R;
    v.sb = gt->stride_b; v.sr = gt->stride_r; v.sv = gt->stride_v;
    v.has_ignore = gt->has_ignore; v.ignore = gt->ignore_index;
    v.align = gt_alignment(*gt, V);
    v.ign_byte = gt->has_ignore && gt->ignore_index >= 0 && gt->ignore_index <= 255;
    v.ign4 = ((unsigned)gt->ignore_index & 0xffu) * 0x01010101u;
    return VU_OK;
}

static int fill_stat_params(StatParams& st, uint32_t flags, unsigned unc_mask, long long V, const vu_gt& gt,
                            const float* thr, const vu_calib* calib, const uint8_t* lut, const double* ncc_gt_map,
                            double* f64, int64_t* i64, const vu_platt_fit* platt_fit, int64_t* platt_i64, double* platt_f64,
                            int64_t* class_counts = nullptr, int n_classes = 0) {
    memset(&st, 0, sizeof(st));
    st.flags = flags;
    st.unc_mask = unc_mask;
    st.magic_bits = 0x4B400000u;
    st.V = V;
    if (flags == 0) return VU_OK;
    if (!f64 || !i64) return set_error(VU_ERR_BAD_ARG, "stat_flags set but stats_f64 / stats_i64 is NULL");
    const uint32_t known = VU_STAT_IMAGE_SUM | VU_STAT_THRESHOLD | VU_STAT_AREA | VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_NCC |
                           VU_STAT_PLATT_FIT | VU_STAT_CLASS_COUNTS;
    if (flags & ~known) return set_error(VU_ERR_BAD_ARG, "unknown stat flag");
    const bool needs_gt = (flags & (VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_PLATT_FIT | VU_STAT_CLASS_COUNTS)) || ((flags & VU_STAT_NCC) && !ncc_gt_map);
    if (flags & VU_STAT_CLASS_COUNTS) {
        if (!class_counts) return set_error(VU_ERR_BAD_ARG, "VU_STAT_CLASS_COUNTS needs class_counts");
        if (n_classes < 1 || n_classes > 256) return set_error(VU_ERR_BAD_ARG, "VU_STAT_CLASS_COUNTS needs 1..256 classes");
        st.cls = reinterpret_cast<long long*>(class_counts);
        st.ncls = n_classes;
    }
    if (flags & VU_STAT_PLATT_FIT) {
        if (!platt_fit || !platt_i64 || !platt_f64)
            return set_error(VU_ERR_BAD_ARG, "VU_STAT_PLATT_FIT needs platt_fit, platt_i64 and platt_f64");
        for (int k = 0; k <= VU_N_PLATT_BINS; ++k) {
            st.platt_edge[k] = platt_fit->edge_u[k];
            if (!(st.platt_edge[k] > 0.0f) || (k && !(st.platt_edge[k] > st.platt_edge[k - 1])))
                return set_error(VU_ERR_BAD_ARG, "vu_platt_fit.edge_u must be positive and increasing");
        }
        st.platt_i64 = reinterpret_cast<long long*>(platt_i64);
        st.platt_f64 = platt_f64;
    }
    if (needs_gt) {
        if (!gt.data) return set_error(VU_ERR_BAD_ARG, "DICE / CALIB / NCC statistics need ground truth");
        if (gt.R < 1 || gt.R > VU_MAX_RATERS) return set_error(VU_ERR_UNSUPPORTED, "gt.R must be 1..8");
        if (gt.dtype != VU_GT_U8 && gt.dtype != VU_GT_I64) return set_error(VU_ERR_BAD_ARG, "gt.dtype");
    }
    if (gt.data) {
        st.gt.data = gt.data; st.gt.dtype = gt.dtype; st.gt.R = gt.R;
        st.gt.sb = gt.stride_b; st.gt.sr = gt.stride_r; st.gt.sv = gt.stride_v;
        st.gt.has_ignore = gt.has_ignore; st.gt.ignore = gt.ignore_index;
        st.gt.align = gt_alignment(gt, V);
        st.gt.ign_byte = gt.has_ignore && gt.ignore_index >= 0 && gt.ignore_index <= 255;
        st.gt.ign4 = ((unsigned)gt.ignore_index & 0xffu) * 0x01010101u;
    }
    for (int k = 0; k < VU_N_UNC; ++k) {
        st.thr[k] = thr[k];
        st.calib[k].a = calib[k].a;
        st.calib[k].b = calib[k].b;
        const int mode = calib[k].mode;
        if ((flags & VU_STAT_CALIB) && (mode < 0 || mode > 2)) return set_error(VU_ERR_BAD_ARG, "vu_calib.mode");
        st.calib[k].increasing = mode != VU_CALIB_PLATT_DEC;
        st.calib[k].identity = mode == VU_CALIB_IDENTITY;
        for (int e = 0; e < VU_N_EDGES; ++e) st.calib[k].edge[e] = st.calib[k].increasing ? calib[k].edge_u[e] : -calib[k].edge_u[e];
        st.calib[k].a2 = -calib[k].a * kLog2e;
        st.calib[k].b2 = calib[k].b * kLog2e;
        st.calib[k].sgn = st.calib[k].increasing ? 1.0f : -1.0f;
    }
    st.lut = lut;
    st.ncc_gt_map = ncc_gt_map;
    st.f64 = f64;
    st.i64 = reinterpret_cast<long long*>(i64);
    return VU_OK;
}

}  // namespace vu

using namespace vu;

extern "C" {

int vu_abi_version(void) { return VU_ABI_VERSION; }

const char* vu_build_info(void) {
    return "libvalunc sm_100a (compute_100a) nvcc " VU_STR(__CUDACC_VER_MAJOR__) "." VU_STR(__CUDACC_VER_MINOR__)
           " built " __DATE__;
}

const char* vu_last_error(void) { return g_err; }

int vu_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return set_error(VU_ERR_NO_DEVICE, "no CUDA device"); }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return set_error(VU_ERR_NO_DEVICE, "libvalunc needs an sm_100 (B200) device");
    return VU_OK;
}

int vu_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(vu_fused_args);
        case 1: return (int)sizeof(vu_map_stats_args);
        case 2: return (int)sizeof(vu_calib);
        case 3: return (int)sizeof(vu_platt_fit);
        case 4: return (int)sizeof(vu_member_scores_args);
        case 5: return (int)sizeof(vu_member_out);
        case 6: return (int)sizeof(vu_radix_state);
        default: return -1;
    }
}

static int fused_pass_checked(const vu_fused_args* a, void* stream);

int vu_fused_pass(const vu_fused_args* a, void* stream) {
    if (!a) return set_error(VU_ERR_BAD_ARG, "args is NULL");
    if (a->struct_size != sizeof(vu_fused_args)) return set_error(VU_ERR_BAD_ARG, "vu_fused_args.struct_size mismatch");
    if (a->slab.flags & VU_SLAB_LOGITS)
        return set_error(VU_ERR_BAD_ARG, "slab.flags: logits are opt-in (relaxed label contract) -- call vu_fused_pass_logits");
    return fused_pass_checked(a, stream);
}

int vu_fused_pass_logits(const vu_fused_args* a, void* stream) {
    if (!a) return set_error(VU_ERR_BAD_ARG, "args is NULL");
    if (a->struct_size != sizeof(vu_fused_args)) return set_error(VU_ERR_BAD_ARG, "vu_fused_args.struct_size mismatch");
    if (a->members.flags) return set_error(VU_ERR_UNSUPPORTED, "member scores are not available for slabs of logits");
    vu_fused_args b = *a;
    b.slab.flags |= VU_SLAB_LOGITS;
    return fused_pass_checked(&b, stream);
}

static int fused_pass_checked(const vu_fused_args* a, void* stream) {
    const vu_slab& s = a->slab;
    if (s.P < 1 || s.B < 0 || s.C < 1 || s.V < 0) return set_error(VU_ERR_BAD_ARG, "slab sizes must be positive");
    if (s.C > 256) return set_error(VU_ERR_UNSUPPORTED, "C > 256 (labels are uint8)");
    if (s.B == 0 || s.V == 0) return VU_OK;  // empty batch: nothing to do (its data pointer may be NULL)
    if (s.draws < 0 || s.draws > 4096 || (s.flags & ~(VU_SLAB_RENORMALIZE | VU_SLAB_DISCRETIZE | VU_SLAB_LOGITS)))
        return set_error(VU_ERR_BAD_ARG, "slab.draws / slab.flags");
    if (s.dtype != VU_SLAB_F32) {
        if (s.dtype != VU_SLAB_BF16 && s.dtype != VU_SLAB_F16) return set_error(VU_ERR_BAD_ARG, "slab.dtype");
        if (s.draws > 1 || s.flags || a->members.flags || a->member_labels)
            return set_error(VU_ERR_UNSUPPORTED, "16-bit slabs: plain slabs only (no draws / producer flags / member labels / member scores); upcast");
    }
    if ((s.draws > 1 || s.flags) && a->members.flags)
        return set_error(VU_ERR_UNSUPPORTED, "member scores in the fused pass are not available for grouped / renormalised / discretised slabs");
    const int64_t n_ptrs = s.P * (s.draws > 1 ? s.draws : 1);
    if (s.member_ptrs || s.member_ptrs_host) {
        if (!s.member_ptrs || !s.member_ptrs_host)
            return set_error(VU_ERR_BAD_ARG, "slab.member_ptrs needs both the device array and its host copy");
        for (int64_t p = 0; p < n_ptrs; ++p)
            if (!s.member_ptrs_host[p]) return set_error(VU_ERR_BAD_ARG, "slab.member_ptrs_host holds a NULL member");
    } else if (!s.data) {
        return set_error(VU_ERR_BAD_ARG, "slab.data is NULL");
    }
    StatParams st;
    const unsigned unc_mask = s.P > 1 ? 7u : 1u;
    int rc = fill_stat_params(st, a->stat_flags, unc_mask, s.V, a->gt, a->threshold, a->calib, a->calib_label_lut, nullptr,
                              a->stats_f64, a->stats_i64, a->platt_fit, a->platt_i64, a->platt_f64, a->class_counts, (int)s.C);
    if (rc != VU_OK) return rc;
    if (a->members.flags) {
        // member-level scores in the same pass: only the unified-warp kernel computes them
        if (a->stat_flags == 0) return set_error(VU_ERR_UNSUPPORTED, "member scores in the fused pass need a statistics mask (e.g. VU_STAT_IMAGE_SUM)");
        rc = launch_k1_uni(a, st, (cudaStream_t)stream);
        return rc == 1 ? set_error(VU_ERR_UNSUPPORTED, "member scores in the fused pass: launch not eligible") : rc;
    }
    if (s.dtype != VU_SLAB_F32) {  // 16-bit slabs: the class-outer TMA form reads them (k1_co_tma.cu)
        rc = launch_k1_co_tma(a, st, (cudaStream_t)stream);
        return rc == 1 ? set_error(VU_ERR_UNSUPPORTED, "16-bit slab not eligible (2..32 members, unit voxel stride, 16-byte aligned rows); upcast") : rc;
    }
    return launch_k1(a, st, (cudaStream_t)stream);
}

int vu_fused_members_supported(const vu_fused_args* a) {
    if (!a || a->struct_size != sizeof(vu_fused_args) || !a->members.flags || a->stat_flags == 0) return 0;
    const vu_slab& s = a->slab;
    if (s.P < 1 || s.B < 1 || s.C < 1 || s.V < 1) return 0;
    StatParams st;
    if (fill_stat_params(st, a->stat_flags, s.P > 1 ? 7u : 1u, s.V, a->gt, a->threshold, a->calib, a->calib_label_lut, nullptr,
                         a->stats_f64, a->stats_i64, a->platt_fit, a->platt_i64, a->platt_f64) != VU_OK)
        return 0;
    return launch_k1_uni(a, st, nullptr, true) == VU_OK ? 1 : 0;
}

int vu_map_stats(const vu_map_stats_args* a, void* stream) {
    if (!a) return set_error(VU_ERR_BAD_ARG, "args is NULL");
    if (a->struct_size != sizeof(vu_map_stats_args)) return set_error(VU_ERR_BAD_ARG, "vu_map_stats_args.struct_size mismatch");
    if (a->B < 0 || a->V < 0) return set_error(VU_ERR_BAD_ARG, "negative size");
    if (a->B == 0 || a->V == 0 || a->stat_flags == 0) return VU_OK;
    if ((a->stat_flags & (VU_STAT_AREA | VU_STAT_DICE | VU_STAT_CALIB | VU_STAT_CLASS_COUNTS)) && !a->labels)
        return set_error(VU_ERR_BAD_ARG, "AREA / DICE / CALIB / CLASS_COUNTS need labels");
    StatParams st;
    unsigned unc_mask = 0;
    for (int k = 0; k < VU_N_UNC; ++k) unc_mask |= a->maps[k] ? (1u << k) : 0u;
    int rc = fill_stat_params(st, a->stat_flags, unc_mask, a->V, a->gt, a->threshold, a->calib, a->calib_label_lut,
                              a->ncc_gt_map, a->stats_f64, a->stats_i64, a->platt_fit, a->platt_i64, a->platt_f64, a->class_counts,
                              a->n_classes);
    if (rc != VU_OK) return rc;
    return launch_map_stats(a, st, (cudaStream_t)stream);
}

int64_t vu_patch_workspace_bytes(int64_t B, int64_t d0, int64_t d1, int64_t d2, int32_t k0, int32_t k1, int32_t k2) {
    if (B < 1 || d0 < 1 || d1 < 1 || d2 < 1 || k0 < 1 || k1 < 1 || k2 < 1 || k0 > d0 || k1 > d1 || k2 > d2) return 0;
    return B * patch_ctas_per_image(d0, d1, d2, k0, k1, k2) * (int64_t)sizeof(unsigned long long);
}

int vu_patch_max_ws(const float* maps, int64_t B, int64_t d0, int64_t d1, int64_t d2, int32_t k0, int32_t k1, int32_t k2,
                    int32_t mean, double* out_max, int64_t* out_first, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!maps || !out_max || !out_first) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    if (B < 0 || d0 < 1 || d1 < 1 || d2 < 1 || k0 < 1 || k1 < 1 || k2 < 1) return set_error(VU_ERR_BAD_ARG, "bad size");
    if (k0 > d0 || k1 > d1 || k2 > d2) return set_error(VU_ERR_BAD_ARG, "patch larger than the map (\"valid\" output is empty)");
    if (B == 0) return VU_OK;
    if (workspace && workspace_bytes < vu_patch_workspace_bytes(B, d0, d1, d2, k0, k1, k2))
        return set_error(VU_ERR_BAD_ARG, "workspace smaller than vu_patch_workspace_bytes");
    if (workspace && (reinterpret_cast<uintptr_t>(workspace) & 7)) return set_error(VU_ERR_BAD_ARG, "workspace must be 8-byte aligned");
    return launch_patch_max(maps, B, d0, d1, d2, k0, k1, k2, mean, out_max, reinterpret_cast<long long*>(out_first),
                            reinterpret_cast<unsigned long long*>(workspace), (cudaStream_t)stream);
}

int vu_patch_max(const float* maps, int64_t B, int64_t d0, int64_t d1, int64_t d2, int32_t k0, int32_t k1, int32_t k2,
                 int32_t mean, double* out_max, int64_t* out_first, void* stream) {
    return vu_patch_max_ws(maps, B, d0, d1, d2, k0, k1, k2, mean, out_max, out_first, nullptr, 0, stream);
}

int vu_border_count(const uint8_t* labels, int64_t B, int64_t d0, int64_t d1, int64_t d2, int64_t* stats_i64, void* stream) {
    if (!labels || !stats_i64) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    if (B < 0 || d0 < 1 || d1 < 1 || d2 < 1) return set_error(VU_ERR_BAD_ARG, "bad size");
    if (B == 0) return VU_OK;
    return launch_border(labels, B, d0, d1, d2, reinterpret_cast<long long*>(stats_i64), (cudaStream_t)stream);
}

// --- host: invert the Platt map on the bin edges (ace.py:329, 350) ------------
// float32 evaluation of 1 / (1 + exp(x*a + b)), x = -u, exactly as NumPy does it
// for a float32 array and Python-float a, b (NEP 50: a, b are cast to float32).
// expf here is the host libm's; the *counts* the kernels produce are exact with
// respect to whatever conf() is used to build the edges, so the Python layer
// rebuilds the edges with NumPy's own exp (diffuncertainty_b200/calibration.py)
// and only falls back to this helper for C callers.
static float platt_host(float u, float a, float b) {
    volatile float z = (-u) * a;
    z = z + b;
    volatile float e = expf(z);
    volatile float d = 1.0f + e;
    return 1.0f / d;
}
static inline uint32_t f2o(float f) {  // order-preserving float -> uint32
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float o2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int vu_platt_invert_edges_host(double a, double b, vu_calib* calib) {
    if (!calib) return set_error(VU_ERR_BAD_ARG, "calib is NULL");
    const float af = (float)a, bf = (float)b;
    calib->a = af;
    calib->b = bf;
    calib->mode = af >= 0.0f ? VU_CALIB_PLATT_INC : VU_CALIB_PLATT_DEC;
    const bool increasing = calib->mode == VU_CALIB_PLATT_INC;
    if (af == 0.0f) {
        // conf is the constant 1 / (1 + exp(b)) (the reference's fallback a = b = 0 when the fit fails): every u passes the edges at
        // or below it, none passes the others
        const float c0 = platt_host(0.0f, af, bf);
        for (int k = 0; k < VU_N_EDGES; ++k)
            calib->edge_u[k] = (double)c0 >= (1.0 + 1e-8) * (double)(k + 1) / 20.0 ? -INFINITY : INFINITY;
        return VU_OK;
    }
    const uint32_t lo_all = f2o(-INFINITY), hi_all = f2o(INFINITY);
    for (int k = 0; k < VU_N_EDGES; ++k) {
        const double edge = (1.0 + 1e-8) * (double)(k + 1) / 20.0;
        // smallest u (increasing) / largest u (decreasing) with conf(u) >= edge
        uint32_t lo = lo_all, hi = hi_all;
        auto ok = [&](uint32_t o) { float c = platt_host(o2f(o), af, bf); return (double)c >= edge; };
        if (increasing) {
            if (!ok(hi)) { calib->edge_u[k] = NAN; continue; }
            while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if (ok(mid)) hi = mid; else lo = mid + 1; }
            calib->edge_u[k] = o2f(lo);
        } else {
            if (!ok(lo)) { calib->edge_u[k] = NAN; continue; }
            while (lo < hi) { uint32_t mid = lo + (hi - lo + 1) / 2; if (ok(mid)) lo = mid; else hi = mid - 1; }
            calib->edge_u[k] = o2f(lo);
        }
    }
    return VU_OK;
}

int vu_radix_hist(const float* values, int64_t n, const vu_gt* weights_gt, int32_t level, const uint32_t* prefixes, int32_t n_prefix,
                  uint64_t* hist, void* stream) {
    if (n < 0 || level < 0 || level > 2) return set_error(VU_ERR_BAD_ARG, "bad size / level");
    if (level > 0 && (!prefixes || n_prefix < 1 || n_prefix > 64)) return set_error(VU_ERR_BAD_ARG, "levels 1, 2 need 1..64 prefixes");
    if (n == 0) return VU_OK;  // an empty map has no storage: its pointer may be NULL
    if (!values || !hist) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    GtView gv;
    int rc = make_gt_view(gv, weights_gt, n);
    if (rc != VU_OK) return rc;
    return launch_radix_hist(values, n, gv, level, prefixes, n_prefix, reinterpret_cast<unsigned long long*>(hist), (cudaStream_t)stream);
}

int vu_radix_walk(const uint64_t* hist, int32_t level, const double* q_host, int32_t n_q, int32_t q_is_f32, int32_t reverse,
                  vu_radix_state* state, void* stream) {
    if (!hist || !state || level < 0 || level > 2) return set_error(VU_ERR_BAD_ARG, "NULL pointer / bad level");
    if (level == 0 && (n_q < 0 || n_q > 31 || (n_q > 0 && !q_host))) return set_error(VU_ERR_BAD_ARG, "level 0 needs 0..31 quantile fractions");
    if (level == 0)
        for (int i = 0; i < n_q; ++i)
            if (!(q_host[i] >= 0.0 && q_host[i] <= 1.0)) return set_error(VU_ERR_BAD_ARG, "Quantiles must be in the range [0, 1]");
    return launch_radix_walk(reinterpret_cast<const unsigned long long*>(hist), level, q_host, n_q, q_is_f32, reverse, state, (cudaStream_t)stream);
}

int vu_radix_hist_state(const float* values, int64_t n, const vu_gt* weights_gt, int32_t level, const vu_radix_state* state, uint64_t* hist,
                        void* stream) {
    if (n < 0 || level < 0 || level > 2) return set_error(VU_ERR_BAD_ARG, "bad size / level");
    if (level > 0 && !state) return set_error(VU_ERR_BAD_ARG, "levels 1, 2 need the state");
    if (n == 0) return VU_OK;
    if (!values || !hist) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    GtView gv;
    int rc = make_gt_view(gv, weights_gt, n);
    if (rc != VU_OK) return rc;
    return launch_radix_hist(values, n, gv, level, nullptr, 0, reinterpret_cast<unsigned long long*>(hist), (cudaStream_t)stream, state);
}

int vu_quantile_select(const float* values, int64_t n, const vu_gt* weights_gt, const double* q_host, int32_t n_q, int32_t q_is_f32,
                       int32_t reverse, uint64_t* hist, vu_radix_state* state, void* stream) {
    if (!hist || !state || n < 0 || (n > 0 && !values)) return set_error(VU_ERR_BAD_ARG, "NULL pointer / negative size");
    for (int level = 0; level < 3; ++level) {
        if (cudaMemsetAsync(hist, 0, (size_t)VU_RADIX_MAX_RANKS * 2048 * sizeof(uint64_t), (cudaStream_t)stream) != cudaSuccess)
            return set_cuda_error("cudaMemsetAsync(hist)");
        int rc = vu_radix_hist_state(values, n, weights_gt, level, state, hist, stream);
        if (rc != VU_OK) return rc;
        rc = vu_radix_walk(hist, level, q_host, n_q, q_is_f32, reverse, state, stream);
        if (rc != VU_OK) return rc;
    }
    return VU_OK;
}

int vu_binned_calib(const float* map, const uint8_t* labels, int64_t V, const vu_gt* gt, const vu_calib* calib, const uint8_t* label_lut,
                    int64_t* out_counts, double* out_sums, void* stream) {
    if (V < 0) return set_error(VU_ERR_BAD_ARG, "negative size");
    if (V == 0) return VU_OK;
    if (!map || !gt || !gt->data || !calib || !out_counts || !out_sums) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    if (calib->mode < 0 || calib->mode > 2) return set_error(VU_ERR_BAD_ARG, "vu_calib.mode");
    GtView gv;
    int rc = make_gt_view(gv, gt, V);
    if (rc != VU_OK) return rc;
    CalibDev cd;
    cd.a = calib->a; cd.b = calib->b;
    cd.increasing = calib->mode != VU_CALIB_PLATT_DEC;
    cd.identity = calib->mode == VU_CALIB_IDENTITY;
    for (int e = 0; e < VU_N_EDGES; ++e) cd.edge[e] = cd.increasing ? calib->edge_u[e] : -calib->edge_u[e];
    cd.a2 = -calib->a * kLog2e; cd.b2 = calib->b * kLog2e; cd.sgn = cd.increasing ? 1.0f : -1.0f;
    return launch_binned_calib(map, labels, V, gv, cd, label_lut, reinterpret_cast<unsigned long long*>(out_counts), out_sums,
                               (cudaStream_t)stream);
}

static int check_seg_batch(const float* const* maps_host, int32_t n_maps, int64_t B, int64_t V) {
    if (!maps_host || n_maps < 1 || n_maps > 4) return set_error(VU_ERR_BAD_ARG, "1..4 maps per batch");
    if (B < 1 || V < 1 || B * n_maps > 65535) return set_error(VU_ERR_BAD_ARG, "B, V must be positive, n_maps * B <= 65535");
    for (int m = 0; m < n_maps; ++m)
        if (!maps_host[m]) return set_error(VU_ERR_BAD_ARG, "NULL map");
    return VU_OK;
}

int vu_quantile_select_batch(const float* const* maps_host, int32_t n_maps, int64_t B, int64_t V, const vu_gt* weights_gt,
                             const double* q_host, int32_t n_q, int32_t q_is_f32, uint32_t reverse_mask, uint64_t* hist,
                             vu_radix_state* states, void* stream) {
    int rc = check_seg_batch(maps_host, n_maps, B, V);
    if (rc != VU_OK) return rc;
    if (!hist || !states) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    if (n_q < 0 || n_q > 31 || (n_q > 0 && !q_host)) return set_error(VU_ERR_BAD_ARG, "0..31 quantile fractions");
    for (int i = 0; i < n_q; ++i)
        if (!(q_host[i] >= 0.0 && q_host[i] <= 1.0)) return set_error(VU_ERR_BAD_ARG, "Quantiles must be in the range [0, 1]");
    GtView gv;
    rc = make_gt_view(gv, weights_gt, V);
    if (rc != VU_OK) return rc;
    const int n_seg = (int)(n_maps * B);
    for (int level = 0; level < 3; ++level) {
        if (cudaMemsetAsync(hist, 0, (size_t)n_seg * VU_RADIX_MAX_RANKS * 2048 * sizeof(uint64_t), (cudaStream_t)stream) != cudaSuccess)
            return set_cuda_error("cudaMemsetAsync(hist)");
        rc = launch_radix_hist_batch(maps_host, n_maps, B, V, gv, level, reinterpret_cast<unsigned long long*>(hist), (cudaStream_t)stream, states);
        if (rc != VU_OK) return rc;
        rc = launch_radix_walk(reinterpret_cast<const unsigned long long*>(hist), level, q_host, n_q, q_is_f32, (int)reverse_mask, states,
                               (cudaStream_t)stream, n_seg, (int)B);
        if (rc != VU_OK) return rc;
    }
    return VU_OK;
}

int vu_binned_calib_batch(const float* const* maps_host, int32_t n_maps, int64_t B, int64_t V, const uint8_t* labels, const vu_gt* gt,
                          const vu_calib* calibs, const uint8_t* label_lut, int64_t* out_counts, double* out_sums, void* stream) {
    int rc = check_seg_batch(maps_host, n_maps, B, V);
    if (rc != VU_OK) return rc;
    if (!gt || !gt->data || !calibs || !out_counts || !out_sums) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    GtView gv;
    rc = make_gt_view(gv, gt, V);
    if (rc != VU_OK) return rc;
    return launch_binned_calib_batch(maps_host, n_maps, B, V, labels, gv, calibs, label_lut, reinterpret_cast<unsigned long long*>(out_counts),
                                     out_sums, (cudaStream_t)stream);
}

int64_t vu_ged_cols(int32_t P, int32_t R) {
    if (P < 1 || R < 1) return 0;
    return 2LL * P * R + R + (int64_t)P * P + P + 2LL * R * R + 3;
}

int vu_member_scores(const vu_member_scores_args* a, void* stream) {
    if (!a) return set_error(VU_ERR_BAD_ARG, "args is NULL");
    if (a->struct_size != sizeof(vu_member_scores_args)) return set_error(VU_ERR_BAD_ARG, "vu_member_scores_args.struct_size mismatch");
    const vu_slab& s = a->slab;
    if (s.P < 1 || s.B < 0 || s.C < 1 || s.V < 0) return set_error(VU_ERR_BAD_ARG, "slab sizes must be positive");
    if (s.C > VU_MAX_CLASSES) return set_error(VU_ERR_UNSUPPORTED, "C > 255");
    if (!(a->flags & (VU_MS_NLL | VU_MS_GED)) || (a->flags & ~(VU_MS_NLL | VU_MS_GED))) return set_error(VU_ERR_BAD_ARG, "flags");
    if ((a->flags & VU_MS_GED) && s.C != 2) return set_error(VU_ERR_BAD_ARG, "GED counts need C == 2 (ged_fast.py:33)");
    if ((a->flags & VU_MS_GED) && s.P > 32) return set_error(VU_ERR_UNSUPPORTED, "GED counts need P <= 32");
    if (s.draws > 1 || s.flags || s.dtype != VU_SLAB_F32)
        return set_error(VU_ERR_UNSUPPORTED, "member scores take float32 members as they are (no draws / producer flags / 16-bit slabs)");
    if (s.B == 0 || s.V == 0) return VU_OK;
    if (s.member_ptrs || s.member_ptrs_host) {
        if (!s.member_ptrs || !s.member_ptrs_host)
            return set_error(VU_ERR_BAD_ARG, "slab.member_ptrs needs both the device array and its host copy");
    } else if (!s.data) {
        return set_error(VU_ERR_BAD_ARG, "slab.data is NULL");
    }
    if (!a->gt.data) return set_error(VU_ERR_BAD_ARG, "member scores need ground truth");
    if ((a->flags & VU_MS_NLL) && (!a->nll_sum || !a->nll_count || !a->nll_bad)) return set_error(VU_ERR_BAD_ARG, "NLL outputs are NULL");
    if ((a->flags & VU_MS_GED) && !a->ged_counts) return set_error(VU_ERR_BAD_ARG, "ged_counts is NULL");
    GtView gv;
    int rc = make_gt_view(gv, &a->gt, s.V);
    if (rc != VU_OK) return rc;
    return launch_member_scores(a, gv, (cudaStream_t)stream);
}

int vu_platt_fit_edges_host(vu_platt_fit* out) {
    if (!out) return set_error(VU_ERR_BAD_ARG, "out is NULL");
    for (int k = 0; k <= VU_N_PLATT_BINS; ++k) {
        const double e = pow(10.0, -12.0 + 14.0 * (double)k / (double)VU_N_PLATT_BINS);  // np.logspace(-12, 2, 257)
        float f = (float)e;
        if ((double)f < e) f = nextafterf(f, INFINITY);
        out->edge_u[k] = f;
    }
    return VU_OK;
}

int vu_synth_slab(float* out, int64_t P, int64_t B, int64_t C, int64_t V, uint64_t seed, int64_t first_image, float scale,
                  void* stream) {
    if (!out) return set_error(VU_ERR_BAD_ARG, "out is NULL");
    if (P < 1 || B < 0 || C < 1 || V < 0 || C > 256) return set_error(VU_ERR_BAD_ARG, "bad size");
    if (B == 0 || V == 0) return VU_OK;
    return launch_synth_slab(out, P, B, C, V, seed, first_image, scale, (cudaStream_t)stream);
}

int vu_synth_gt(uint8_t* out, const float* slab, int64_t P, int64_t B, int64_t C, int64_t V, int32_t R, uint64_t seed,
                int64_t first_image, float flip, float ignore_frac, int32_t ignore_value, void* stream) {
    if (!out || !slab) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    if (P < 1 || B < 0 || C < 1 || V < 0 || R < 1 || R > VU_MAX_RATERS) return set_error(VU_ERR_BAD_ARG, "bad size");
    if (B == 0 || V == 0) return VU_OK;
    return launch_synth_gt(out, slab, P, B, C, V, R, seed, first_image, flip, ignore_frac, ignore_value, (cudaStream_t)stream);
}

int vu_copy_2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height, int32_t kind,
                     void* stream) {
    if (!dst || !src) return set_error(VU_ERR_BAD_ARG, "NULL pointer");
    if (width < 0 || height < 0 || dst_pitch < width || src_pitch < width) return set_error(VU_ERR_BAD_ARG, "bad width / height / pitch");
    if (kind != 1 && kind != 2) return set_error(VU_ERR_BAD_ARG, "kind must be 1 (host -> device) or 2 (device -> host)");
    if (width == 0 || height == 0) return VU_OK;
    if (cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width, (size_t)height,
                          kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess)
        return set_cuda_error("cudaMemcpy2DAsync");
    return VU_OK;
}

int vu_set_option(const char* key, int64_t value) {
    if (!key) return set_error(VU_ERR_BAD_ARG, "key is NULL");
    std::lock_guard<std::mutex> lk(g_mu);
    g_options[key] = value;
    return VU_OK;
}

int64_t vu_get_counter(const char* key) {
    if (!key) return -1;
    if (strcmp(key, "k1_num_variants") == 0) return num_fast_variants();
    if (strcmp(key, "k1_num_tma_variants") == 0) return num_tma_variants();
    if (strncmp(key, "k1_variant.", 11) == 0) {  // "k1_variant.<i>.<field 0..6>"
        int i = 0, f = 0;
        if (sscanf(key + 11, "%d.%d", &i, &f) != 2 || f < 0 || f > 6) return -1;
        int d[7];
        if (describe_fast_variant(i, d) != 0) return -1;
        return d[f];
    }
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_counters.find(key);
    return it == g_counters.end() ? 0 : it->second;
}

}  // extern "C"
