// Shared helpers for libimm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/imm_b200.h"

namespace immb {

extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launch_count;

int set_error(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return IMMB_OK;
}

#define IMMB_REQUIRE(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return immb::set_error(IMMB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// ---- TF32 split: hi = rna_tf32(v), lo = rna_tf32(v - hi) ------------------------------------------
__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = tf32_rna(v);
  lo = tf32_rna(v - hi);
}
__device__ __forceinline__ void store_split(float* hi_p, float* lo_p, size_t i, float v) {
  if (lo_p) {
    float h, l;
    split_tf32(v, h, l);
    hi_p[i] = h;
    lo_p[i] = l;
  } else {
    hi_p[i] = v;
  }
}
__device__ __forceinline__ float load_split(const float* hi_p, const float* lo_p, size_t i) {
  float v = __ldg(hi_p + i);
  if (lo_p) v += __ldg(lo_p + i);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// BN-backward reduction fused into a dgrad epilogue (conv_tc2_pair_kernel, stats_mode 2): the raw conv output and the
// BN constants of the layer that produced the dgrad's input
struct Tc2BnReduce {
  const float* y;
  int ycs, relu;
  const float *scale, *shift, *mean, *invstd;
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;

}  // namespace immb
