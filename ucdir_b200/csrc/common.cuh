// Shared helpers for the ucdir_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/ucdir_b200.h"

namespace ucdir {

// x * sigmoid(x) with the reference's operation order (model/ucdir.py:48-50): x * (1 / (1 + exp(-x))).
__device__ __forceinline__ float swish_f(float x) { return x * (1.0f / (1.0f + expf(-x))); }
__device__ __forceinline__ float lrelu_f(float x) { return fmaxf(0.2f * x, x); }  // model/ucdir.py:414-416

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// GroupNorm(1, C) scalars of one sample from {sum, sumsq} accumulated in double by producer epilogues.
struct GnScalars { float mean, rstd; };
__device__ __forceinline__ GnScalars gn_scalars(const double* s0, const double* s1, int b, double count, float eps) {
  double sum = s0[2 * b], sq = s0[2 * b + 1];
  if (s1) { sum += s1[2 * b]; sq += s1[2 * b + 1]; }
  double mean = sum / count;
  double var = sq / count - mean * mean;   // biased variance, as F.group_norm
  if (var < 0) var = 0;
  GnScalars g;
  g.mean = (float)mean;
  g.rstd = (float)(1.0 / sqrt(var + (double)eps));
  return g;
}

void set_error(const char* fmt, ...);

// Per-device caches: function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize) and the SM count belong to a DEVICE,
// not to the process -- a process that touches a second GPU must opt in again there.  Launch-side statics are indexed by
// the current device ordinal.
constexpr int UCDIR_MAX_DEV = 64;
inline int cur_dev() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < UCDIR_MAX_DEV) ? d : 0; }
int sm_count();   // multiprocessors of the current device (cached per device, c_abi.cu)
extern long long g_launches;

// launchers implemented per translation unit
int launch_conv_f32(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_sgemm_f32(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_softmax_f32(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_maxpool2(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_guidance(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_time_embed(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_gather_tiles(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_scatter(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_tc_conv(const ucdir_op_t& op, cudaStream_t st, bool dry);
bool tc_mix_halo_applies(const ucdir_op_t& op);
int launch_tc_mix_halo(const ucdir_op_t& op, cudaStream_t st);
bool tc_final_halo_applies(const ucdir_op_t& op);
int launch_tc_final_halo(const ucdir_op_t& op, cudaStream_t st);
bool tc_dense_halo_applies(const ucdir_op_t& op);
int launch_tc_dense_halo(const ucdir_op_t& op, cudaStream_t st);
// setmaxnreg moves registers between warpgroups through a per-CTA pool that holds only what the CTA itself released:
// what `low_threads` threads release going from the launch allocation down to `low` must cover what `high_threads` need.
int check_reg_pool(const void* kernel, const char* name, int low_threads, int low, int high_threads, int high);
int launch_tc_attn(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_gn_apply_bf16(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_cast(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_crop_tiles(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_gn_stats_f32(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_gn_apply_f32(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_layout(const ucdir_op_t& op, cudaStream_t st, bool dry);
int launch_to_image_u8(const ucdir_op_t& op, cudaStream_t st, bool dry);

}  // namespace ucdir
