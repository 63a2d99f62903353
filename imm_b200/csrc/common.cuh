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

// ---- scaled fp16 split ("H16" planes): the operands of the kind::f16 tensor-core convolutions --------------------
// A tensor v is stored as two fp16 planes and ONE power-of-two scale 2^e (exact):
//   t = v * 2^e;  hi = rn_f16(t);  lo = rn_f16((t - hi) * 2^11);      v ~= (hi + lo * 2^-11) * 2^-e
// hi carries 11 significant bits, lo the next 11 (the residual t - hi is exact in fp32 and at most half an ulp of hi,
// so lo * 2^-11 never underflows before hi does): 22 bits, the same as the TF32 pair.  A product of two such tensors
// is hi*hi (main accumulator) + (hi*lo + lo*hi) * 2^-11 (cross accumulator), both exact fp16 x fp16 products
// accumulated in fp32 by the tensor core; the dropped lo*lo term is 2^-22 relative.  Since the scale is a power of
// two the result does not depend on e as long as hi neither saturates (|t| <= 65504) nor leaves the normal range
// (|t| >= 2^-14) for the elements that matter: e is chosen so that the tensor's largest magnitude sits near 2^12
// (16x headroom above, 26 binades below).  Scale record in device memory: int32 {e, amax_bits}: producers read e and
// atomicMax the bit pattern of the largest |v| they wrote; immb_scale_update turns amax into the next step's e.
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: launchers cache "already configured" per device slot
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return dev & (kMaxDevices - 1);
}

constexpr int kH16TargetExp = 12;
// Tensors under DELAYED scaling (activations, gradients: this step's exponent comes from the previous step's maximum)
// aim lower: 2^8, i.e. 256x headroom.  Early in training the gradient planes were observed to grow 20-40x from one
// step to the next (tools/soak_step.py: dy maxima at step 31 of config 2), which saturated fp16 under the 16x headroom
// of 2^12.  Weights are rescaled exactly from their current maximum every step and keep 2^12.
constexpr int kH16DelayedTargetExp = 8;
constexpr float kH16LoScale = 2048.f;               // 2^11
constexpr float kH16LoInv = 1.f / 2048.f;

__device__ __forceinline__ uint16_t f16_sat(float t) {      // round-to-nearest-even, saturating to +-65504
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(t));
  return r;
}
__device__ __forceinline__ float f16_to_f32(uint16_t h) {
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
  return f;
}
// t = v * 2^e (already scaled) -> (hi, lo) fp16 bit patterns
__device__ __forceinline__ void split_h16(float t, uint16_t& hi, uint16_t& lo) {
  hi = f16_sat(t);
  lo = f16_sat((t - f16_to_f32(hi)) * kH16LoScale);
}
__device__ __forceinline__ float join_h16(uint16_t hi, uint16_t lo) {      // scaled value t
  return fmaf(f16_to_f32(lo), kH16LoInv, f16_to_f32(hi));
}
__device__ __forceinline__ uint32_t pack2(uint16_t a, uint16_t b) { return (uint32_t)a | ((uint32_t)b << 16); }
__device__ __forceinline__ float exp2i(int e) { return scalbnf(1.f, e); }      // exact 2^e

// warp-wide maximum of non-negative floats -> one atomicMax on the tensor's amax record (bit patterns of non-negative
// floats order like unsigned integers)
__device__ __forceinline__ void h16_track_amax(int32_t* scale_rec, float amax) {
  uint32_t b = __float_as_uint(amax);
  b = __reduce_max_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0 && b != 0) atomicMax(reinterpret_cast<unsigned int*>(scale_rec + 1), b);
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

// scale records ({e, amax bits}, common.cuh "H16 planes") of a conv's activation operand, weight operand and -- when
// the result is written as H16 planes -- output tensor
struct Tc2Scales {
  const int32_t* a;
  const int32_t* b;
  int32_t* o;
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;

}  // namespace immb
