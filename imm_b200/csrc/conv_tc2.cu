// Persistent, halo-reuse tcgen05 convolution for the stride-1 3x3 layers (forward and dgrad).
//
// What limits conv_tc_kernel (conv_tc.cu) on the small-channel / high-resolution layers is L2 -> smem operand
// traffic: every tap re-fetches its own shifted 128-pixel A tile (9 x 2 planes x 16 KB per 32-channel chunk).
// Here ONE (16+2) x 16-pixel halo box per 32-channel chunk is fetched (2 planes x 36 KB: 4x less traffic) and all
// nine taps are read out of it with shifted UMMA descriptors:
//   * tile = 16 rows x 8 columns of one image (M = 128, row m = h*8 + w);
//   * the halo box is 16 pixels wide, so one halo row is exactly 2048 bytes: output row h / tap (r,s) starts at
//     smem row (h+r)*16 + s  ->  every 8-row core-matrix group is SBO = 2048 bytes apart, and all groups of a tap
//     share the same swizzle phase s (the hardware swizzles on absolute smem address bits, so base_offset stays 0);
//   * out-of-image halo pixels are zero-filled by TMA (= SAME padding).
// The CTA is persistent (grid = #SMs, static tile scheduler) with TWO TMEM accumulators, so the epilogue of tile i
// (tcgen05.ld -> bias/ReLU/split -> stores) overlaps the MMAs of tile i+1, and the TMA producer never drains.
// Rings: A (2 halo slots), B (per-tap weight slices, 2-4 slots), TMEM (2 accumulators).
#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace immb {

using namespace ptx;

struct Tc2Tap {
  int b_tap, ro, so;         // weight slice index; halo row / column offset (0..2)
};

struct Tc2Params {
  int tiles_w, tiles_h, n_img, n_tiles_n, total_tiles;
  int kchunks;
  Tc2Tap taps[9];
  float* out_hi;
  float* out_lo;
  const float* bias;
  int relu;
  int H, W, ocs, n_cols, n_store;
  int bo_mode;               // debug (IMMB_TC2_BO): 0 = base_offset 0 (correct), 1 = base_offset s
};

template <int BN, int PASSES>
struct Tc2Cfg {
  static constexpr uint32_t NPL = PASSES == 3 ? 2 : 1;
  static constexpr uint32_t A_PLANE = 18 * 16 * 128;          // 36864 B: (16+2) halo rows x 16 pixels x 32 ch
  static constexpr uint32_t A_SLOT = A_PLANE * NPL;
  static constexpr uint32_t A_SLOTS = 2;
  static constexpr uint32_t B_PLANE = BN * 128;
  static constexpr uint32_t B_SLOT = B_PLANE * NPL;
  static constexpr uint32_t B_SLOTS = BN <= 64 ? 4 : 2;
  static constexpr uint32_t SMEM_BYTES = A_SLOTS * A_SLOT + B_SLOTS * B_SLOT + 1024 + 256;
  static constexpr int ACC_COLS = (BN + 31) / 32 * 32;
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : 256);
};

template <int BN, int PASSES>
__global__ void __launch_bounds__(192, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                const __grid_constant__ Tc2Params p) {
  using Cfg = Tc2Cfg<BN, PASSES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + Cfg::A_SLOTS * Cfg::A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + Cfg::B_SLOTS * Cfg::B_SLOT);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + Cfg::A_SLOTS;
  uint64_t* b_full = a_empty + Cfg::A_SLOTS;
  uint64_t* b_empty = b_full + Cfg::B_SLOTS;
  uint64_t* t_full = b_empty + Cfg::B_SLOTS;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (uint32_t i = 0; i < Cfg::B_SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA_hi);
    prefetch_tmap(&mapB_hi);
    if (PASSES == 3) { prefetch_tmap(&mapA_lo); prefetch_tmap(&mapB_lo); }
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t ai = 0, bi = 0;                       // running slot counters
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int nt = t % p.n_tiles_n;
        int r = t / p.n_tiles_n;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int img = r / p.tiles_h;
        const int n_off = nt * BN;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const uint32_t as = ai % Cfg::A_SLOTS, aph = (ai / Cfg::A_SLOTS) & 1;
          mbar_wait(&a_empty[as], aph ^ 1);
          uint8_t* sa = a_base + as * Cfg::A_SLOT;
          mbar_expect_tx(&a_full[as], Cfg::A_SLOT);
          tma_load_5d(sa, &mapA_hi, &a_full[as], kc * 32, tw * 8 - 1, 0, th * 16 - 1, img);
          if (PASSES == 3) tma_load_5d(sa + Cfg::A_PLANE, &mapA_lo, &a_full[as], kc * 32, tw * 8 - 1, 0, th * 16 - 1, img);
          ++ai;
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t bs = bi % Cfg::B_SLOTS, bph = (bi / Cfg::B_SLOTS) & 1;
            mbar_wait(&b_empty[bs], bph ^ 1);
            uint8_t* sb = b_base + bs * Cfg::B_SLOT;
            mbar_expect_tx(&b_full[bs], Cfg::B_SLOT);
            tma_load_3d(sb, &mapB_hi, &b_full[bs], kc * 32, n_off, p.taps[tap].b_tap);
            if (PASSES == 3) tma_load_3d(sb + Cfg::B_PLANE, &mapB_lo, &b_full[bs], kc * 32, n_off, p.taps[tap].b_tap);
            ++bi;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = idesc_tf32(128, BN, 0, 0);
      uint32_t ai = 0, bi = 0, ti = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
        mbar_wait(&t_empty[acc], tph ^ 1);            // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_COLS;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const uint32_t as = ai % Cfg::A_SLOTS, aph = (ai / Cfg::A_SLOTS) & 1;
          mbar_wait(&a_full[as], aph);
          const uint32_t a_hi = smem_u32(a_base + as * Cfg::A_SLOT);
          const uint32_t a_lo = a_hi + Cfg::A_PLANE;
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t bs = bi % Cfg::B_SLOTS, bph = (bi / Cfg::B_SLOTS) & 1;
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_hi = smem_u32(b_base + bs * Cfg::B_SLOT);
            const uint32_t b_lo = b_hi + Cfg::B_PLANE;
            const Tc2Tap tp = p.taps[tap];
            const uint32_t a_off = (uint32_t)(tp.ro * 16 + tp.so) * 128u;
            // measured on B200: the MMA unit applies the 128B swizzle to the absolute smem address, so a descriptor
            // that starts s rows into the 1024-byte pattern needs base_offset = 0 (setting it to s reads garbage)
            const uint32_t bo = p.bo_mode == 0 ? 0u : (uint32_t)tp.so;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint32_t ko = k4 * 32;
              const uint64_t da_hi = smem_desc_sw128(a_hi + a_off + ko, 16, 2048, 2, bo);
              const uint64_t db_hi = smem_desc_sw128(b_hi + ko, 16, 1024);
              mma_tf32(tmem_d, da_hi, db_hi, idesc, (kc > 0 || tap > 0 || k4 > 0) ? 1u : 0u);
              if (PASSES == 3) {
                const uint64_t da_lo = smem_desc_sw128(a_lo + a_off + ko, 16, 2048, 2, bo);
                const uint64_t db_lo = smem_desc_sw128(b_lo + ko, 16, 1024);
                mma_tf32(tmem_d, da_hi, db_lo, idesc, 1u);
                mma_tf32(tmem_d, da_lo, db_hi, idesc, 1u);
              }
            }
            mma_commit(&b_empty[bs]);
            ++bi;
          }
          mma_commit(&a_empty[as]);
          ++ai;
        }
        mma_commit(&t_full[acc]);
        ++ti;
      }
    }
  } else {
    // ===== epilogue (warps 2..5) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int hl = m >> 3, wl = m & 7;
    uint32_t ti = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int nt = t % p.n_tiles_n;
      int r = t / p.n_tiles_n;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h;
      const int img = r / p.tiles_h;
      const int n_off = nt * BN;
      const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
      const int h = th * 16 + hl, w = tw * 8 + wl;
      const size_t pix = ((size_t)img * p.H + h) * p.W + w;
      mbar_wait(&t_full[acc], tph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + (uint32_t)c0, v);
        const int col0 = n_off + c0;
        if (col0 >= p.n_cols) continue;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.n_cols) v[j] += __ldg(p.bias + col0 + j);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        float* o = p.out_hi + pix * p.ocs + col0;
        if (p.out_lo) {
          float* ol = p.out_lo + pix * p.ocs + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j >= p.n_store) break;
            float4 hi4, lo4;
            split_tf32(v[j], hi4.x, lo4.x);
            split_tf32(v[j + 1], hi4.y, lo4.y);
            split_tf32(v[j + 2], hi4.z, lo4.z);
            split_tf32(v[j + 3], hi4.w, lo4.w);
            *reinterpret_cast<float4*>(o + j) = hi4;
            *reinterpret_cast<float4*>(ol + j) = lo4;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j >= p.n_store) break;
            *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      // all of this warp's TMEM reads are complete (tcgen05.wait::ld inside tmem_ld32): release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
      ++ti;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------
int tc_make_act_map(CUtensorMap* m, const float* base, int N, int H, int W, int C, int cs, bool parity_split,
                    int box_w, int box_h, int box_n, int swizzle_mn);
int tc_make_w_map(CUtensorMap* m, const float* base, int taps, int Nn, int Kd, int bn);
int tc_pick_bn(int ncols);

bool conv_tc2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("IMMB_TC2");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on == 1;
}

// op: 0 forward (x -> y), 1 dgrad (dy -> dx); stride-1 3x3 only
bool conv_tc2_eligible(const immb_conv_desc* d, int op) {
  if (!conv_tc2_enabled()) return false;
  if (d->x_layout != IMMB_XLAYOUT_NHWC || d->kh != 3 || d->kw != 3 || d->stride != 1) return false;
  if (d->H % 16 || d->W % 8) return false;
  if (d->pad_t != 1 || d->pad_l != 1) return false;
  (void)op;
  return true;
}

template <int BN, int PASSES>
static int launch_tc2(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                      const CUtensorMap& b_lo, const Tc2Params& p, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN, PASSES>;
  auto kern = conv_tc2_kernel<BN, PASSES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2 smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(a_hi, a_lo, b_hi, b_lo, p);
  return check_launch("conv_tc2_kernel");
}

// act: the tensor the halo boxes are read from ([N,H,W,act_cs], act_c valid channels); wts: [9][ncols_pad][kd]
int conv_tc2_run(const immb_conv_desc* d, int op, const float* act_hi, const float* act_lo, int act_c, int act_cs,
                 const float* w_hi, const float* w_lo, int w_rows, int kd, const float* bias, int relu,
                 float* out_hi, float* out_lo, int ocs, int ncols, int n_store, cudaStream_t st) {
  const int passes = d->precision == IMMB_PREC_TF32 ? 1 : 3;
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  p.tiles_w = d->W / 8; p.tiles_h = d->H / 16; p.n_img = d->N;
  const int bn = tc_pick_bn(ncols);
  p.n_tiles_n = ceil_div(ncols, bn);
  p.total_tiles = p.tiles_w * p.tiles_h * p.n_img * p.n_tiles_n;
  p.kchunks = ceil_div(kd, 32);
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      Tc2Tap& t = p.taps[r * 3 + s];
      t.b_tap = r * 3 + s;
      t.ro = op == 0 ? r : 2 - r;
      t.so = op == 0 ? s : 2 - s;
    }
  p.out_hi = out_hi; p.out_lo = out_lo; p.bias = bias; p.relu = relu;
  p.H = d->H; p.W = d->W; p.ocs = ocs; p.n_cols = ncols; p.n_store = n_store;
  { const char* e = getenv("IMMB_TC2_BO"); p.bo_mode = e ? atoi(e) : 0; }
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = tc_make_act_map(&a_hi, act_hi, d->N, d->H, d->W, act_c, act_cs, false, 16, 18, 1, 0))) return rc;
  if ((rc = tc_make_w_map(&b_hi, w_hi, 9, w_rows, kd, bn))) return rc;
  a_lo = a_hi; b_lo = b_hi;
  if (passes == 3) {
    if ((rc = tc_make_act_map(&a_lo, act_lo, d->N, d->H, d->W, act_c, act_cs, false, 16, 18, 1, 0))) return rc;
    if ((rc = tc_make_w_map(&b_lo, w_lo, 9, w_rows, kd, bn))) return rc;
  }
#define IMMB_CASE(BN_)                                                                     \
  if (bn == BN_)                                                                           \
    return passes == 3 ? launch_tc2<BN_, 3>(a_hi, a_lo, b_hi, b_lo, p, st)                 \
                       : launch_tc2<BN_, 1>(a_hi, a_lo, b_hi, b_lo, p, st);
  IMMB_CASE(16)
  IMMB_CASE(32)
  IMMB_CASE(64)
  IMMB_CASE(96)
  IMMB_CASE(128)
#undef IMMB_CASE
  return set_error(IMMB_ERR_INVALID, "conv_tc2: unsupported BN %d", bn);
}

}  // namespace immb
