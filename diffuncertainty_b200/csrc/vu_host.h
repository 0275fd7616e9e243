// Host-side plumbing shared by the translation units of libvalunc.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "valunc.h"

#define VU_STR2(x) #x
#define VU_STR(x) VU_STR2(x)

namespace vu {
struct StatParams;
struct GtView;
struct CalibDev;

int set_error(int code, const char* msg);        // records msg, returns code
int set_cuda_error(const char* where);           // records cudaGetLastError text, returns VU_ERR_CUDA
int check_launch(const char* kernel);            // cudaPeekAtLastError after a launch
void count_launch(const char* kernel);           // launch counters for bench.py ("gpu_launches")
long long get_option(const char* key, long long dflt);
int device_sm_count();

int launch_k1(const vu_fused_args* a, const StatParams& st, cudaStream_t stream);
int launch_k1_tma(const vu_fused_args* a, const StatParams& st, cudaStream_t stream);  // 1 = not eligible
int launch_k1_uni(const vu_fused_args* a, const StatParams& st, cudaStream_t stream, bool dry_run = false);  // 1 = not eligible
int launch_k1_co_tma(const vu_fused_args* a, const StatParams& st, cudaStream_t stream);  // class-outer TMA form; 1 = not eligible
int num_tma_variants();
int num_fast_variants();
int describe_fast_variant(int i, int* out7);
int launch_map_stats(const vu_map_stats_args* a, const StatParams& st, cudaStream_t stream);
int launch_patch_max(const float* maps, long long B, long long d0, long long d1, long long d2, int k0, int k1, int k2,
                     int mean, double* out_max, long long* out_first, unsigned long long* tile_max, cudaStream_t stream);
long long patch_ctas_per_image(long long d0, long long d1, long long d2, int k0, int k1, int k2);
int launch_border(const uint8_t* labels, long long B, long long d0, long long d1, long long d2, long long* stats_i64,
                  cudaStream_t stream);
int launch_radix_hist(const float* values, long long n, const GtView& gt, int level, const unsigned* prefixes, int n_prefix,
                      unsigned long long* hist, cudaStream_t stream, const vu_radix_state* state = nullptr);
// n_seg > 0: batched (one CTA per segment, `reverse` is a bit mask over the maps, seg_B images per map)
int launch_radix_walk(const unsigned long long* hist, int level, const double* q_host, int n_q, int q_is_f32, int reverse,
                      vu_radix_state* state, cudaStream_t stream, int n_seg = 0, int seg_B = 0);
int launch_radix_hist_batch(const float* const* maps, int n_maps, long long B, long long V, const GtView& gt, int level,
                            unsigned long long* hist, cudaStream_t stream, const vu_radix_state* states);
int launch_binned_calib_batch(const float* const* maps, int n_maps, long long B, long long V, const uint8_t* labels, const GtView& gt,
                              const vu_calib* cals_dev, const uint8_t* lut, unsigned long long* counts, double* sums, cudaStream_t stream);
int launch_binned_calib(const float* map, const uint8_t* labels, long long V, const GtView& gt, const CalibDev& cal, const uint8_t* lut,
                        unsigned long long* counts, double* sums, cudaStream_t stream);
int launch_member_scores(const vu_member_scores_args* a, const GtView& gt, cudaStream_t stream);
int launch_synth_slab(float* out, long long P, long long B, long long C, long long V, uint64_t seed,
                      long long first_image, float scale, cudaStream_t stream);
int launch_synth_gt(uint8_t* out, const float* slab, long long P, long long B, long long C, long long V, int R,
                    uint64_t seed, long long first_image, float flip, float ignore_frac, int ignore_value,
                    cudaStream_t stream);
}  // namespace vu
