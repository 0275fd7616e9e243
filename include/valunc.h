/*
 * valunc.h -- C ABI of libvalunc.so: the B200 (sm_100a) implementation of the
 * ValUES per-pixel uncertainty hot path (C2 measures + C3 aggregation +
 * calibration / ambiguity / failure-detection inputs).
 *
 * The reference (JakobLC/DiffUncertainty) is pure Python and has no FFI layer;
 * these entry points are what a binding for its hot path would call.  Each one
 * cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - strides are in ELEMENTS, not bytes; the slab is never copied or made
 *     contiguous (the per-image view softmax_pred[:, i] of test_2D.py:969 is
 *     strided);
 *   - functions only enqueue work on `stream` (a cudaStream_t passed as
 *     void*); they never synchronise, allocate or free caller memory;
 *   - statistics outputs ACCUMULATE (+=) so one buffer can span many calls;
 *     the caller zeroes them;
 *   - return value: VU_OK or a negative vu_status; nothing throws.
 */
#ifndef VALUNC_H
#define VALUNC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VU_ABI_VERSION 4

#if defined(__GNUC__)
#define VU_API __attribute__((visibility("default")))
#else
#define VU_API
#endif

typedef enum vu_status {
    VU_OK = 0,
    VU_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, bad enum     */
    VU_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels handle         */
    VU_ERR_CUDA = -3,         /* a CUDA runtime call failed (see vu_last_error)*/
    VU_ERR_NO_DEVICE = -4     /* no sm_100 device visible                      */
} vu_status;

#define VU_N_UNC 3        /* TU, AU, EU in this order everywhere               */
#define VU_N_BINS 21      /* ace.py:350-356: 20 bins, arrays of length 21      */
#define VU_N_EDGES 19     /* interior edges of np.linspace(0, 1+1e-8, 21)      */
#define VU_MAX_RATERS 8   /* R: LIDC 4, GTA 5         */
#define VU_MAX_CLASSES 255 /* labels are uint8 (test_2D.py:818)                */

/* ground-truth element types handed over by the data loader
 * (batch["seg"].long(), test_2D.py:1122-1124, or uint8 PNG decodes) */
#define VU_GT_U8 0
#define VU_GT_I64 1

/* which per-image statistics vu_fused_pass / vu_map_stats accumulate */
#define VU_STAT_IMAGE_SUM 0x01u /* aggregate_uncertainties.py:37-39           */
#define VU_STAT_THRESHOLD 0x02u /* aggregate_uncertainties.py:124-125         */
#define VU_STAT_AREA 0x04u      /* prediction_shape_stats.py:10-12            */
#define VU_STAT_DICE 0x08u      /* test_2D.py:878-886 (needs gt)              */
#define VU_STAT_CALIB 0x10u     /* ace.py:350-356, 431-437 (needs gt)         */
#define VU_STAT_NCC 0x20u       /* ncc.py:17-27 + experiment_dataloader.py:283*/
#define VU_STAT_PLATT_FIT 0x40u /* ace.py:117-136: the 256-bin data the Platt fit is
                                   run on (validation split; needs gt).  Dataset
                                   level: accumulates into platt_i64 / platt_f64,
                                   not into the per-image rows                   */
#define VU_N_PLATT_BINS 256     /* ace.py:17,31: np.logspace(-12, 2, 257) edges  */
#define VU_STAT_CLASS_COUNTS 0x80u /* test_2D.py:901-918 (multi-class Dice, the risk of the AURC of GTA-like sets): per rater and
                                   class the integers its macro Dice is made of -- tp = #(label == c & gt == c), pred =
                                   #(label == c), gt = #(gt == c) over the voxels the rater does not ignore (needs gt);
                                   accumulates into class_counts, not into the per-image rows                      */

/* ---- layout of one row of the per-image statistics buffers --------------- */
/* double row (VU_F64_COLS doubles per image)                                 */
#define VU_F64_SUM 0       /* [3]  sum of TU, AU, EU                          */
#define VU_F64_THR_SUM 3   /* [3]  sum of map[map >= t]                       */
#define VU_F64_NCC_G 6     /* [1]  sum g      (g = rater variance map)        */
#define VU_F64_NCC_GG 7    /* [1]  sum g*g                                    */
#define VU_F64_NCC_U 8     /* [3]  sum u                                      */
#define VU_F64_NCC_UU 11   /* [3]  sum u*u                                    */
#define VU_F64_NCC_GU 14   /* [3]  sum g*u                                    */
#define VU_F64_BIN_SUMS 17 /* [3][21] sum of Platt confidences per bin        */
#define VU_F64_COLS 80
/* int64 row (VU_I64_COLS int64 per image)                                    */
#define VU_I64_THR_COUNT 0   /* [3]  #(map >= t)                              */
#define VU_I64_AREA 3        /* [1]  #(label > 0)                             */
#define VU_I64_BORDER 4      /* [1]  neighbour-differs count (vu_border_count)*/
#define VU_I64_NVOX 5        /* [1]  voxels seen                              */
#define VU_I64_BIN_TOTAL 6   /* [3][21] samples per bin                       */
#define VU_I64_BIN_TRUE 69   /* [3][21] correct samples per bin               */
#define VU_I64_DICE_TP 132   /* [8]  per rater: pred==1 & gt==1 & valid       */
#define VU_I64_DICE_PRED 140 /* [8]  per rater: pred==1 & valid               */
#define VU_I64_DICE_GT 148   /* [8]  per rater: gt==1 & valid                 */
#define VU_I64_COLS 156

/* The stacked probabilities "softmax_pred" of test_2D.py:1277, shape
 * (P, B, C, V) with V = H*W(*D) flattened; any strides.                      */
typedef struct vu_slab {
    const float* data;
    int64_t P, B, C, V;
    int64_t stride_p, stride_b, stride_c, stride_v;
    /* Optional "stack without copying" form (torch.stack(groups) of test_2D.py:1277 materialises the slab): the P
     * members stay where the forward passes wrote them.  member_ptrs is a DEVICE array of P base pointers, each
     * addressing a (B, C, V) tensor with the strides above (stride_p is ignored, data may be NULL);
     * member_ptrs_host is the same array in HOST memory (used to validate alignment).  Both NULL = `data` form.   */
    const float* const* member_ptrs;
    const float* const* member_ptrs_host;
    /* The upstream elementwise producers of the slab, folded into the read -- all zero = the slab is
     * taken as it is.
     *   draws > 1  every member is a GROUP of `draws` stochastic draws whose mean is the member: the reference's
     *              torch.stack(softmax_pred_groups).mean(dim=1) (test_2D.py:1277; groups are built at :1134-1136, :1160).
     *              `data` form: draw d of member p lies at data + p * stride_p + d * stride_d; member_ptrs form: P * draws
     *              pointers, draw-minor.  The mean is formed in torch's order (cascade sum, true division).
     *   VU_SLAB_RENORMALIZE  each draw is renormalised over its classes first (_renormalize_probabilities,
     *              test_2D.py:188-194: p / max(sum_c p, eps) where sum_c p > eps).
     *   VU_SLAB_DISCRETIZE   each draw is replaced by the one-hot vector of its argmax (--discretize, test_2D.py:1272-1275).
     *   VU_SLAB_LOGITS       the slab holds LOGITS: each draw is softmax'ed over its classes first (F.softmax(output, dim=1),
     *              test_2D.py:1181, 1185, 1225, 1241, 1256), before the two producers above.  The probabilities are never
     *              written anywhere: one full write + read of the slab less than softmax followed by vu_fused_pass.
     *              RELAXED CONTRACT: the device's exponential is not torch's, so the probabilities differ from the
     *              reference's by a few ulp; maps stay within the 1e-5 tolerance (plus 1e-6 absolute, the reference's own
     *              float32 rounding of p next to 1), labels can differ from the reference's where the two largest mean
     *              probabilities of a voxel agree to ~1e-7 relative (measured rates: INTEGRATION.md section 5).
     *              Opt-in: vu_fused_pass rejects the flag, vu_fused_pass_logits sets it.
     * Grouped draws and the two elementwise producers run on the generic kernel (any strides, any class count); plain logits
     * have TMA forms for C = 2, 3, 4, 19.                                                                               */
    int64_t stride_d;
    int32_t draws;
    uint32_t flags;
    float renorm_eps; /* test_2D.py:189: 1e-12 */
    /* Element type of the slab.  VU_SLAB_F32 (0, the default): float32 as declared.  VU_SLAB_BF16 / VU_SLAB_F16: `data` (and
     * every member pointer) addresses 16-bit elements -- what a network under autocast returns; strides stay in ELEMENTS.
     * The kernel widens every value to float32 as it reads it (exact) and computes as for a float32 slab, so the results are
     * those of the upcast slab bit for bit, at half the HBM traffic and without the upcast copy (the reference computes float32
     * maps whatever the input dtype, test_utils.py:836; its own half-precision intermediate arithmetic is not reproduced).
     * Available for plain slabs (no draws / producer flags / per-member labels / member scores) of 2..32 members with unit voxel
     * stride and 16-byte aligned rows (V, strides multiples of 8); vu_fused_pass returns VU_ERR_UNSUPPORTED otherwise and the
     * caller upcasts. */
    int32_t dtype;
} vu_slab;
#define VU_SLAB_F32 0
#define VU_SLAB_BF16 1
#define VU_SLAB_F16 2
#define VU_SLAB_RENORMALIZE 0x1u
#define VU_SLAB_DISCRETIZE 0x2u
#define VU_SLAB_LOGITS 0x4u

/* batch["seg"]: (B, R, V) reference segmentations (test_2D.py:1122-1124).    */
typedef struct vu_gt {
    const void* data; /* NULL: no ground truth                                */
    int32_t dtype;    /* VU_GT_U8 or VU_GT_I64                                */
    int32_t R;        /* raters, 1..VU_MAX_RATERS                             */
    int64_t stride_b, stride_r, stride_v;
    int32_t has_ignore;
    int64_t ignore_index; /* ace.py:492-499 ignore_value / test_2D.py:880     */
} vu_gt;

/* Platt-scaled calibration binning (ace.py:325-329, 350-352) for one
 * uncertainty type.  conf(u) = 1 / (1 + exp(-u*a + b)) in float32.  The bin
 * of a voxel is decided by comparing u with `edge_u` -- the 19 interior bin
 * edges pulled back through conf() on the host with the reference's own
 * float32 expression (vu_platt_invert_edges_host), so counts do not depend on
 * the device's expf.  A NaN edge never matches; NaN u goes to slot 20 as
 * np.digitize does.                                                          */
#define VU_CALIB_PLATT_DEC 0 /* a < 0: conf falls with u, bin = #{k: u <= edge_u[k]} */
#define VU_CALIB_PLATT_INC 1 /* a >= 0: conf rises with u, bin = #{k: u >= edge_u[k]} */
#define VU_CALIB_IDENTITY 2  /* the map already holds confidences (calc_ace(correct,
                                conf) drop-in): conf = clip(u, 0, 1), edge_u[k] = the
                                smallest float32 >= the float64 edge                  */
typedef struct vu_calib {
    float a, b;
    float edge_u[VU_N_EDGES];
    int32_t mode;
} vu_calib;

/* Binning of the Platt-fit data (ace.py:31,117-126): sample u falls into bin
 * #{k : u >= edge_k} - 1, clamped to [0, 255] (NaN -> 255).  edge_u[k] is the
 * smallest float32 >= the float64 edge k, so the float32 comparison has the
 * truth value of NumPy's float64 one.  Fill with vu_platt_fit_edges_host, or
 * from np.logspace itself for bit-exact parity with NumPy's pow.              */
typedef struct vu_platt_fit {
    float edge_u[VU_N_PLATT_BINS + 1];
} vu_platt_fit;

/* Member-level scores computed by vu_fused_pass itself, while it streams the slab for the maps (one read of the slab instead
 * of the two of vu_fused_pass + vu_member_scores): the outputs, their layout and their meaning are those of
 * vu_member_scores_args below (GED: ged_fast.py:44-131 with the majority counts taken from the labels the pass computes;
 * NLL: test_2D.py:1043-1120).  flags == 0: not requested.  Available for binary slabs (C == 2) of at most 32 members with
 * unit voxel stride and 16-byte aligned rows, at most 4 uint8 references with word-aligned rows, and a statistics mask the
 * unified kernel is built for (vu_fused_members_supported); otherwise vu_fused_pass returns VU_ERR_UNSUPPORTED and the caller
 * runs vu_member_scores as a second pass.                                                                                */
typedef struct vu_member_out {
    uint32_t flags;      /* VU_MS_NLL | VU_MS_GED                              */
    float eps;           /* test_2D.py:1043: 1e-12                             */
    double* nll_sum;     /* (B, R, P), accumulated                             */
    int64_t* nll_count;  /* (B, R)                                             */
    int64_t* nll_bad;    /* (B)                                                */
    int64_t* ged_counts; /* (B, vu_ged_cols(P, R))                             */
} vu_member_out;

typedef struct vu_fused_args {
    uint32_t struct_size; /* sizeof(vu_fused_args), ABI check                 */
    uint32_t stat_flags;  /* VU_STAT_* or 0                                   */
    vu_slab slab;
    /* maps, each (B, V) contiguous fp32; NULL = do not store.
     * P >= 2: TU / AU / EU of test_utils.py:833-859.
     * P == 1: `tu` receives 1 - max softmax ("pred_entropy",
     *         test_utils.py:862-864); au / eu are not written.               */
    float* tu;
    float* au;
    float* eu;
    /* argmax of the member mean, (B, V) uint8 (test_2D.py:871, 815-818)      */
    uint8_t* labels;
    vu_gt gt;
    float threshold[VU_N_UNC];   /* VU_STAT_THRESHOLD                         */
    vu_calib calib[VU_N_UNC];    /* VU_STAT_CALIB                             */
    /* optional label lookup applied to the predicted label before it is
     * compared with the references in VU_STAT_CALIB (quirk Q14: LIDC PNGs
     * store foreground as 255).  NULL = compare class ids.  256 entries.     */
    const uint8_t* calib_label_lut;
    double* stats_f64;  /* (B, VU_F64_COLS), accumulated; NULL iff flags == 0 */
    int64_t* stats_i64; /* (B, VU_I64_COLS), accumulated                      */
    /* VU_STAT_PLATT_FIT: dataset-level outputs, accumulated
     *   platt_i64 (3, 256, 2): samples, correct samples per uncertainty type and bin
     *   platt_f64 (3, 256)   : sum of the uncertainties of the samples              */
    const vu_platt_fit* platt_fit; /* HOST pointer; read during the call      */
    int64_t* platt_i64;
    double* platt_f64;
    /* argmax of every member, (P, B, V) uint8 contiguous; NULL = do not store.
     * save_prediction writes one label image per member next to the mean's
     * (test_2D.py:810-818); they are also what prediction_shape_stats
     * (mean_pred=False) and GED consume.                                      */
    uint8_t* member_labels;
    vu_member_out members; /* member-level scores in the same pass (see vu_member_out) */
    /* VU_STAT_CLASS_COUNTS: (B, R, C, 3) int64, [.., c, 0] tp, [.., c, 1] pred, [.., c, 2] gt; accumulated.  References that
     * are not a class index (and not the ignore value) are skipped here; dice() raises on them (dice_wrapped.py:57-58).   */
    int64_t* class_counts;
} vu_fused_args;

/* ABI / build info ---------------------------------------------------------- */
VU_API int vu_abi_version(void);
VU_API const char* vu_build_info(void); /* "sm_100a nvcc 12.9 ..."                   */
VU_API const char* vu_last_error(void); /* text of the last CUDA error on this thread*/
VU_API int vu_device_check(void);       /* VU_OK if the current device is sm_100     */
VU_API int vu_struct_size(int which);   /* 0 vu_fused_args, 1 vu_map_stats_args, 2 vu_calib, 3 vu_platt_fit:
                                    lets a binding verify its struct layout   */

/* The fused streaming pass.  Replaces, for a whole batch in one launch:
 *   mean over members        test_2D.py:971
 *   argmax label             test_2D.py:871, 815-818
 *   calculate_uncertainty    unc_mod_utils/test_utils.py:833-859
 *   calculate_one_minus_msr  unc_mod_utils/test_utils.py:862-864 (P == 1)
 * and, per stat_flags, the reductions of
 *   image_level_aggregation / threshold_aggregation
 *                            aggregate_uncertainties.py:37-39, 124-125
 *   _compute_area            prediction_shape_stats.py:10-12
 *   binary Dice counts       test_2D.py:878-886
 *   calib histograms         ace.py:350-356, 431-437
 *   NCC sums                 ncc.py:17-27
 * reading the slab exactly once.                                             */
VU_API int vu_fused_pass(const vu_fused_args* args, void* stream);
/* The same pass over a slab of LOGITS (args->slab.data / member_ptrs address the network outputs before F.softmax,
 * test_2D.py:1181-1256): softmax over the classes is applied to every draw while it is read (VU_SLAB_LOGITS is set by this
 * entry point; see vu_slab for the relaxed label contract).  Everything else -- outputs, statistics, strides, member lists,
 * grouped draws, renormalise / discretise -- as vu_fused_pass.  Member-level scores (args->members) are not available.  */
VU_API int vu_fused_pass_logits(const vu_fused_args* args, void* stream);
/* 1 if vu_fused_pass can compute args->members in the same pass (see vu_member_out), 0 if not (the reason is left in
 * vu_last_error).  Nothing is launched.                                                                                  */
VU_API int vu_fused_members_supported(const vu_fused_args* args);

/* Same reductions on maps / labels that already exist in device memory (the
 * reference's file-based evaluation, evaluation/eval_experiments.py:348-355,
 * works on stored maps).  maps[k] may be NULL to skip a type.  `labels` is
 * needed for AREA / DICE / CALIB.                                            */
typedef struct vu_map_stats_args {
    uint32_t struct_size;
    uint32_t stat_flags;
    int64_t B, V;
    const float* maps[VU_N_UNC]; /* each (B, V) contiguous                    */
    const uint8_t* labels;       /* (B, V)                                    */
    vu_gt gt;
    float threshold[VU_N_UNC];
    vu_calib calib[VU_N_UNC];
    const uint8_t* calib_label_lut;
    const double* ncc_gt_map; /* optional (B, V) float64 ready-made GT map
                                 (evaluation/utils/gta.py:15-35); NULL = use
                                 the variance of the gt raters               */
    double* stats_f64;
    int64_t* stats_i64;
    const vu_platt_fit* platt_fit; /* HOST pointer (VU_STAT_PLATT_FIT)        */
    int64_t* platt_i64;            /* (3, 256, 2), accumulated                */
    double* platt_f64;             /* (3, 256), accumulated                   */
    int64_t* class_counts;         /* VU_STAT_CLASS_COUNTS: (B, R, n_classes, 3), accumulated */
    int32_t n_classes;             /* classes of `labels` (1..256), VU_STAT_CLASS_COUNTS only */
} vu_map_stats_args;
VU_API int vu_map_stats(const vu_map_stats_args* args, void* stream);

/* patch_level_aggregation (aggregate_uncertainties.py:16-34) for B maps of
 * shape (d0, d1, d2) (2-D maps: d0 = 1) with box (k0, k1, k2) ("valid").
 * out_max[b]   = max box sum (float64)
 * out_first[b] = row-major linear index, in the (d0-k0+1, d1-k1+1, d2-k2+1)
 *                output grid, of the first box with np.isclose(sum, max)
 *                (rtol 1e-5, atol 1e-8).
 * mean != 0 divides the sums by k0*k1*k2 first (patch_level_aggregation's
 * mean=True).  Four small launches on `stream`, no scratch memory.           */
VU_API int vu_patch_max(const float* maps, int64_t B, int64_t d0, int64_t d1, int64_t d2,
                 int32_t k0, int32_t k1, int32_t k2, int32_t mean,
                 double* out_max, int64_t* out_first, void* stream);
/* Same with an optional device workspace of vu_patch_workspace_bytes(...) bytes
 * (8-byte aligned, contents irrelevant): the first pass leaves every CTA's own
 * maximum there and the index pass skips the CTAs that cannot hold a box
 * np.isclose to the image's maximum -- about half the time of vu_patch_max.
 * workspace == NULL behaves like vu_patch_max.                                */
VU_API int64_t vu_patch_workspace_bytes(int64_t B, int64_t d0, int64_t d1, int64_t d2,
                 int32_t k0, int32_t k1, int32_t k2);
VU_API int vu_patch_max_ws(const float* maps, int64_t B, int64_t d0, int64_t d1, int64_t d2,
                 int32_t k0, int32_t k1, int32_t k2, int32_t mean,
                 double* out_max, int64_t* out_first,
                 void* workspace, int64_t workspace_bytes, void* stream);

/* _compute_border (prediction_shape_stats.py:15-30) on (B, d0, d1, d2) uint8
 * label maps; adds into stats_i64[b][VU_I64_BORDER] (row stride VU_I64_COLS) */
VU_API int vu_border_count(const uint8_t* labels, int64_t B, int64_t d0, int64_t d1, int64_t d2,
                    int64_t* stats_i64, void* stream);

/* Member-level scores on the slab vu_fused_pass reads (one more pass over it):
 *   VU_MS_NLL  _compute_likelihood_stats / _compute_expected_nll (test_2D.py:1043-1120):
 *              nll_sum[b][r][p] += sum over the valid voxels of reference r of ln(max(slab[p, b, gt, v], eps))
 *              (float64), nll_count[b][r] += valid voxels (every voxel without gt.has_ignore, test_2D.py:1055-1060),
 *              nll_bad[b] += valid voxels whose reference is not a class index (torch.gather raises there, :1067).
 *   VU_MS_GED  every integer ged_binary_fast (ged_fast.py:44-131) is made of; needs C == 2 and P <= 32.
 *              ged_counts[b][...], vu_ged_cols(P, R) int64 per image, G = R:
 *                [0, PG)            pg_tp[p][g]    member p == 1 & reference g == 1 & g valid      (ged_fast.py:56-60)
 *                [PG, 2PG)          pg_pred[p][g]  member p == 1 & g valid                         (:56,61)
 *                + G                g_sum[g]       reference g == 1 & g valid                      (:57,62)
 *                + P*P              pp_tp[p][q]    member p == 1 & member q == 1                   (:84-86)
 *                + P                pos[p]         member p == 1                                   (:87)
 *                + G*G              gg_tp[i][j]    reference i == 1 & reference j == 1 & j valid   (:98-101)
 *                + G*G              gg_sum[i][j]   reference i == 1 & j valid                      (:100,102)
 *                + 3                majority tp / pred / gt counts (:121-131); needs `labels`, the (B, V) label of the
 *                                   member mean that vu_fused_pass wrote, else they stay untouched
 *              member labels follow torch.argmax (first maximum, NaN is maximal).
 * All outputs accumulate.  The float32 Dice / GED arithmetic on these P*G + P*P + G*G numbers is the caller's.       */
#define VU_MS_NLL 0x1u
#define VU_MS_GED 0x2u
typedef struct vu_member_scores_args {
    uint32_t struct_size;  /* sizeof(vu_member_scores_args) */
    uint32_t flags;
    vu_slab slab;
    vu_gt gt;              /* required */
    const uint8_t* labels; /* (B, V) or NULL */
    float eps;             /* test_2D.py:1043: 1e-12 */
    double* nll_sum;       /* (B, R, P) */
    int64_t* nll_count;    /* (B, R) */
    int64_t* nll_bad;      /* (B) */
    int64_t* ged_counts;   /* (B, vu_ged_cols(P, R)) */
} vu_member_scores_args;
VU_API int64_t vu_ged_cols(int32_t P, int32_t R);
VU_API int vu_member_scores(const vu_member_scores_args* args, void* stream);

/* One histogram pass of an exact radix select over float32 values (order statistics for np.quantile:
 * find_threshold.py:69-77, ace.py:387-388).  key = order-preserving 32-bit image of the float (NaN sorts last).
 *   level 0: hist[key >> 21]                                   += w      (one slot of 2048 counters)
 *   level 1: hist[s * 2048 + ((key >> 10) & 0x7ff)]            += w      for keys with  key >> 21 == prefixes[s]
 *   level 2: hist[s * 2048 + (key & 0x3ff)]                    += w      for keys with  key >> 10 == prefixes[s]
 * w = 1, or with `weights_gt` (R references over the same n voxels, batch stride unused) the number of references
 * that are not the ignore value.  hist is uint64, accumulated: several maps can be folded into one selection.
 * prefixes: DEVICE array of n_prefix <= 64 values (ignored at level 0).                                          */
VU_API int vu_radix_hist(const float* values, int64_t n, const vu_gt* weights_gt, int32_t level, const uint32_t* prefixes,
                         int32_t n_prefix, uint64_t* hist, void* stream);

/* The same selection without host round trips between the passes.  A vu_radix_state lives in DEVICE memory; vu_radix_walk
 * descends one level on the device and leaves there what the next histogram pass needs; vu_radix_hist_state is vu_radix_hist
 * with the prefixes (and their number) taken from the state.  One selection = hist 0, walk 0, hist 1, walk 1, hist 2, walk 2 on
 * one stream and ONE read-back of the state (total, key[]).
 *   walk level 0: total = sum of the level-0 histogram; the ranks to select are derived from `n_q` quantile fractions
 *                 (np.quantile, method "linear": h = (total - 1) q evaluated in float64, or in float32 when q_is_f32;
 *                 lo = min(floor(h), total - 1), hi = min(lo + 1, total - 1)): rank[2i] = lo_i, rank[2i + 1] = hi_i, and
 *                 rank[2 n_q] = total - 1 (a NaN anywhere in the data sorts last).  `reverse`: rank r -> total - 1 - r.
 *                 n_q <= 31.  total == 0 leaves n_rank = 0.
 *   walk level 1, 2: digit of every rank inside its slot, new prefixes; level 2 writes key[] (order-preserving float keys).
 * hist: (64, 2048) uint64, zeroed by the caller before every histogram pass.                                               */
#define VU_RADIX_MAX_RANKS 64
typedef struct vu_radix_state {
    int64_t total;
    int32_t n_rank, n_slot;
    int64_t rank[VU_RADIX_MAX_RANKS];
    int64_t residual[VU_RADIX_MAX_RANKS];
    uint32_t prefix[VU_RADIX_MAX_RANKS];      /* of every rank */
    int32_t slot[VU_RADIX_MAX_RANKS];         /* histogram slot of every rank (ranks with one prefix share a slot) */
    uint32_t slot_prefix[VU_RADIX_MAX_RANKS]; /* prefix of every slot */
    uint32_t key[VU_RADIX_MAX_RANKS];
} vu_radix_state;
VU_API int vu_radix_walk(const uint64_t* hist, int32_t level, const double* q_host, int32_t n_q, int32_t q_is_f32, int32_t reverse,
                         vu_radix_state* state, void* stream);
VU_API int vu_radix_hist_state(const float* values, int64_t n, const vu_gt* weights_gt, int32_t level, const vu_radix_state* state,
                               uint64_t* hist, void* stream);
/* The whole selection over ONE array in one call: clears `hist` (64 x 2048 uint64 of device workspace) before each of the
 * three passes, runs them and the three descents on `stream`.  The caller reads `state` back afterwards.                  */
VU_API int vu_quantile_select(const float* values, int64_t n, const vu_gt* weights_gt, const double* q_host, int32_t n_q,
                              int32_t q_is_f32, int32_t reverse, uint64_t* hist, vu_radix_state* state, void* stream);

/* The three bincounts of calc_eqace (ace.py:392-396) for one map on 19 caller-given thresholds: sample u of voxel v
 * (weight = valid references) falls into bin #{k : u' >= edge_u[k]}, u' = u (mode INC / IDENTITY) or -u with
 * sign-flipped thresholds (mode DEC), NaN u into slot 20.  out_counts (2, 21): samples, correct samples;
 * out_sums (21): float64 sum of the confidences (ace.py:329 in float32, clipped).  Accumulated.                  */
VU_API int vu_binned_calib(const float* map, const uint8_t* labels, int64_t V, const vu_gt* gt, const vu_calib* calib,
                           const uint8_t* label_lut, int64_t* out_counts, double* out_sums, void* stream);

/* Batched forms for per-image eqACE (ace.py:378-406 inside the image loop of calibration_error, ace.py:484-515): segment
 * s = m * B + b is image b of map m (n_maps <= 4 uncertainty types, each a contiguous (B, V) device array whose pointer is
 * given in the HOST array maps_host).  Every segment has its own selection / histogram; the references of image b are
 * gt + b * stride_b.  One call enqueues the work of all segments: 3 x (memset, histogram pass, descent) resp. one binning pass.
 *   vu_quantile_select_batch: hist (n_maps * B, 64, 2048) uint64 workspace, states (n_maps * B); reverse_mask bit m: map m
 *                             selects the mirrored ranks (a confidence that falls with the uncertainty).
 *   vu_binned_calib_batch:    calibs = DEVICE array of n_maps * B vu_calib (edge_u per segment); out_counts (n_maps * B, 2, 21),
 *                             out_sums (n_maps * B, 21), accumulated.                                                        */
VU_API int vu_quantile_select_batch(const float* const* maps_host, int32_t n_maps, int64_t B, int64_t V, const vu_gt* weights_gt,
                                    const double* q_host, int32_t n_q, int32_t q_is_f32, uint32_t reverse_mask, uint64_t* hist,
                                    vu_radix_state* states, void* stream);
VU_API int vu_binned_calib_batch(const float* const* maps_host, int32_t n_maps, int64_t B, int64_t V, const uint8_t* labels,
                                 const vu_gt* gt, const vu_calib* calibs, const uint8_t* label_lut, int64_t* out_counts,
                                 double* out_sums, void* stream);

/* Host helper: pull the 19 interior edges of np.linspace(0, 1+1e-8, 21) back
 * through the reference's float32 Platt expression (ace.py:329) by bisection
 * over float32 bit patterns.  Fills calib->edge_u / increasing from a, b.    */
VU_API int vu_platt_invert_edges_host(double a, double b, vu_calib* calib);
/* Host helper: the 257 edges 10^(-12 + 14 k / 256) rounded up to float32.     */
VU_API int vu_platt_fit_edges_host(vu_platt_fit* out);

/* Deterministic synthetic slab for benchmarks and smoke tests:
 * softmax(scale * N(0,1)) over C, from a counter-based RNG keyed by
 * (seed, first_image + b, p, c, v) so every GPU count sees the same images.
 * Writes a contiguous (P, B, C, V) slab.                                     */
VU_API int vu_synth_slab(float* out, int64_t P, int64_t B, int64_t C, int64_t V,
                  uint64_t seed, int64_t first_image, float scale, void* stream);
/* synthetic (B, R, V) uint8 references: label of member 0 with `flip` of the
 * voxels re-drawn uniformly and `ignore_frac` set to ignore_value.           */
VU_API int vu_synth_gt(uint8_t* out, const float* slab, int64_t P, int64_t B, int64_t C, int64_t V,
                int32_t R, uint64_t seed, int64_t first_image, float flip, float ignore_frac,
                int32_t ignore_value, void* stream);

/* Host <-> device staging helper of the end-to-end path (host_pipeline.py): `height` rows of `width` bytes between pitched
 * buffers in ONE asynchronous copy on `stream` -- the (P, n, C, V) block of n images out of a (P, B, C, V) slab in pinned host
 * memory is P runs of n*C*V floats, `B*C*V*4` bytes apart.  kind: 1 host -> device, 2 device -> host.  The stacked
 * predictions of test_2D.py:1277 live in host memory whenever the forward passes ran on another device or were read back from
 * the files of test_2D.py:857.                                                                                            */
VU_API int vu_copy_2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height,
                            int32_t kind, void* stream);

/* tuning / introspection used by bench.py and the variant sweep */
VU_API int vu_set_option(const char* key, int64_t value); /* e.g. "k1_variant"       */
VU_API int64_t vu_get_counter(const char* key);           /* e.g. "launches"         */

#ifdef __cplusplus
}
#endif
#endif /* VALUNC_H */
