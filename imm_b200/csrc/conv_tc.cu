// tcgen05 + TMA implicit-GEMM convolution engine for sm_100a (forward, dgrad, wgrad).
//
// Forward / dgrad ("shifted-window GEMM", both operands K-major):
//   D[128 pixels, BN channels] = sum over taps (r,s) and 32-channel chunks of  A_tap[128, 32] * B_tap[BN, 32]^T
//   * A_tap is a TH x TW (x TN images) patch of the NHWC activation, shifted by the tap offset, fetched with ONE
//     5-D TMA box (c, w, hpar, h, n): out-of-bounds coordinates are zero-filled by the TMA unit, which is exactly
//     TF 'SAME' padding; stride-2 convs address the input through a parity-split view (c+wpar*C, w/2, hpar, h/2, n).
//   * B_tap is a [BN, 32] slice of the packed weights ([tap][Cout][Cin] for fwd, [tap][Cin][Cout] for dgrad).
//   * both land in shared memory in the canonical K-major SWIZZLE_128B layout (128-byte rows, 1024-byte 8-row
//     atoms), are consumed by tcgen05.mma kind::tf32 (M=128, N=BN, K=8 per instruction; +32 B descriptor advance)
//     and accumulate in TMEM; 3xTF32 issues hi*hi + hi*lo + lo*hi into the same accumulator.
//   * warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue (tcgen05.ld -> bias/ReLU ->
//     vectorised NHWC stores).  mbarrier ring: full[stage] (TMA -> MMA), empty[stage] (tcgen05.commit -> TMA),
//     tmem_full (last commit -> epilogue).
// Wgrad (both operands MN-major): see conv_tc_wgrad_kernel below.
#include <cudaTypedefs.h>
#include <mutex>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace immb {

using namespace ptx;

constexpr int kTileM = 128;
constexpr int kMaxTaps = 9;

struct TapEntry {
  int b_tap, dc, dw, hp, dh;
};

struct TcParams {
  int TW, TH, TN;             // tile box in pixels (TW*TH*TN == 128)
  int tiles_w, tiles_h;
  int n_taps, kchunks;        // class 0 (and the only class of forward / stride-1 dgrad launches)
  TapEntry taps[kMaxTaps];
  // stride-2 dgrad: the 4 output-parity classes run as blockIdx.z of ONE launch; classes 1..3 (<= 2 taps each... <= 4)
  int n_classes;
  int n_taps_c[3];
  TapEntry taps_c[3][4];
  int out_add_c[3][2];
  float* out_hi;
  float* out_lo;
  const float* bias;
  int relu;
  int N, PH, PW;              // tile-space extents (predication)
  int OH, OW, ocs;            // output tensor
  int out_mul, out_add_h, out_add_w;
  int n_cols;                 // valid output channels
  int n_store;                // channels written (n_cols rounded up to 4, <= ocs; the excess is exact zeros)
  // scaled-fp16 operands (F16 kernels): scale records of the A / B operands and of H16 output planes; 32-byte k-steps
  // to issue in the last K chunk
  const int32_t* a_scale;
  const int32_t* b_scale;
  int32_t* o_scale;
  int k_last;
};

// ---------------------------------------------------------------------------------------------------------
template <int BN, int PASSES, int STAGES>
struct FwdCfg {
  static constexpr uint32_t A_BYTES = kTileM * 128;
  static constexpr uint32_t B_BYTES = BN * 128;
  static constexpr uint32_t NPL = PASSES == 3 ? 2 : 1;      // planes per operand
  static constexpr uint32_t STAGE_BYTES = (A_BYTES + B_BYTES) * NPL;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  // 3-pass: A_hi x [w_hi ; w_lo] is ONE MMA of N = 2*BN (hi*lo lands in columns [BN, 2BN), added in the epilogue)
  static constexpr int ACC = PASSES == 3 ? 2 * ((BN + 31) / 32 * 32) : (BN + 31) / 32 * 32;
  static constexpr int TMEM_COLS = ACC <= 32 ? 32 : (ACC <= 64 ? 64 : (ACC <= 128 ? 128 : 256));
};

// F16: scaled fp16 split planes (common.cuh), kind::f16: a K chunk of 128 bytes holds 64 channels; PASSES 3 = hi*hi into
// columns [0,BN), hi*lo (same N-concatenated MMA) and lo*hi into the cross columns [BN,2BN), which carry 2^-11.
template <int BN, int PASSES, int STAGES, bool F16>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
               const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
               const __grid_constant__ TcParams p) {
  using Cfg = FwdCfg<BN, PASSES, STAGES>;
  constexpr int KC = F16 ? 64 : 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tw = blockIdx.x % p.tiles_w;
  const int th = (blockIdx.x / p.tiles_w) % p.tiles_h;
  const int tn = blockIdx.x / (p.tiles_w * p.tiles_h);
  const int n_off = blockIdx.y * BN;
  const int cls = blockIdx.z;
  const int n_taps = cls == 0 ? p.n_taps : p.n_taps_c[cls - 1];
  const TapEntry* taps = cls == 0 ? p.taps : p.taps_c[cls - 1];
  const int out_add_h = cls == 0 ? p.out_add_h : p.out_add_c[cls - 1][0];
  const int out_add_w = cls == 0 ? p.out_add_w : p.out_add_c[cls - 1][1];
  const int num_k = n_taps * p.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA_hi);
    prefetch_tmap(&mapB_hi);
    if (PASSES == 3) {
      prefetch_tmap(&mapA_lo);
      prefetch_tmap(&mapB_lo);
    }
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      // ===== TMA producer (whole warp runs the loop; one elected lane issues) =====
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int tap = kt / p.kchunks, kc = kt - tap * p.kchunks;
        const TapEntry t = taps[tap];
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        if (elect_one()) {
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        const int c = kc * KC + t.dc, w = tw * p.TW + t.dw, h = th * p.TH + t.dh, n = tn * p.TN;
        tma_load_5d(st, &mapA_hi, &full[s], c, w, t.hp, h, n);
        if (PASSES == 3) tma_load_5d(st + Cfg::A_BYTES, &mapA_lo, &full[s], c, w, t.hp, h, n);
        uint8_t* sb = st + Cfg::A_BYTES * Cfg::NPL;
        tma_load_3d(sb, &mapB_hi, &full[s], kc * KC, n_off, t.b_tap);
        if (PASSES == 3) tma_load_3d(sb + Cfg::B_BYTES, &mapB_lo, &full[s], kc * KC, n_off, t.b_tap);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer (whole warp, uniform values; one elected lane issues) =====
      constexpr uint32_t idesc = idesc_kind<F16>(kTileM, BN, 0, 0);
      constexpr uint32_t idesc2 = idesc_kind<F16>(kTileM, 2 * BN, 0, 0);      // [w_hi ; w_lo] (contiguous in the stage)
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t b_hi = a_hi + Cfg::A_BYTES * Cfg::NPL;
        const int kc_ = kt % p.kchunks;
        const int nk = (F16 && kc_ == p.kchunks - 1) ? p.k_last : 4;
        if (elect_one()) {
        // one descriptor per operand and stage; k-steps ADD to its start-address field (32 bytes = +2): rebuilding the
        // fields per MMA cost ~13 instructions per MMA on the single issuing thread (the narrow tiles were bound by it)
        const uint64_t da0_hi = smem_desc_sw128(a_hi, 16, 1024), db0_hi = smem_desc_sw128(b_hi, 16, 1024);
        const uint64_t da0_lo = da0_hi + (uint64_t)(Cfg::A_BYTES >> 4);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          if (k4 < nk) {
          const uint64_t da_hi = da0_hi + (uint64_t)(2 * k4);     // 8 tf32 / 16 fp16 = 32 bytes along K inside the 128-byte swizzle row
          const uint64_t db_hi = db0_hi + (uint64_t)(2 * k4);
          mma_kind<F16>(tmem_base, da_hi, db_hi, PASSES == 3 ? idesc2 : idesc, (kt > 0 || k4 > 0) ? 1u : 0u);
          if (PASSES == 3) {
            const uint64_t da_lo = da0_lo + (uint64_t)(2 * k4);
            // TF32: lo*hi joins the main columns; F16: the lo plane carries 2^11, so it joins hi*lo in the cross columns
            mma_kind<F16>(F16 ? tmem_base + BN : tmem_base, da_lo, db_hi, idesc, 1u);
          }
          }
        }
        mma_commit(&empty[s]);            // frees the smem stage once these MMAs have drained
        if (kt == num_k - 1) mma_commit(tmem_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: warps 2..5; warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int wl = m % p.TW, hl = (m / p.TW) % p.TH, nl = m / (p.TW * p.TH);
    const int n = tn * p.TN + nl, h = th * p.TH + hl, w = tw * p.TW + wl;
    const bool row_ok = (n < p.N) && (h < p.PH) && (w < p.PW);
    const size_t pix = ((size_t)n * p.OH + (size_t)(h * p.out_mul + out_add_h)) * p.OW +
                       (size_t)(w * p.out_mul + out_add_w);
    float s_main = 1.f, s_cross = 1.f, s_out = 1.f, amax = 0.f;
    if (F16) {
      const int e = __ldg(p.a_scale) + __ldg(p.b_scale);
      s_main = exp2i(-e);
      s_cross = exp2i(-e - 11);
      if (p.out_lo) s_out = exp2i(__ldg(p.o_scale));
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      if (PASSES == 3) {
        float v2[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = F16 ? fmaf(v2[j], s_cross, v[j] * s_main) : v[j] + v2[j];
      } else {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (F16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= s_main;
        }
      }
      if (!row_ok) continue;
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
      if (F16 && p.out_lo) {
        // H16 output planes (ocs % 8 == 0, n_store % 8 == 0): 16-byte stores
        uint16_t* oh = reinterpret_cast<uint16_t*>(p.out_hi) + pix * p.ocs + col0;
        uint16_t* ol = reinterpret_cast<uint16_t*>(p.out_lo) + pix * p.ocs + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (col0 + j >= p.n_store) break;
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint16_t h0, l0, h1, l1;
            amax = fmaxf(amax, fmaxf(fabsf(v[j + 2 * u]), fabsf(v[j + 2 * u + 1])));
            split_h16(v[j + 2 * u] * s_out, h0, l0);
            split_h16(v[j + 2 * u + 1] * s_out, h1, l1);
            hw[u] = pack2(h0, h1);
            lw[u] = pack2(l0, l1);
          }
          *reinterpret_cast<uint4*>(oh + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(ol + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        continue;
      }
      float* o = p.out_hi + pix * p.ocs + col0;
      const bool vec8 = (p.ocs % 8 == 0) && (p.n_store % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 31) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out_lo) & 31) == 0);
      if (vec8) {
        // one full 32-byte sector per store (thread = pixel row: 128-bit stores fill half a sector per lane)
        float* ol = p.out_lo ? p.out_lo + pix * p.ocs + col0 : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (col0 + j >= p.n_store) break;
          if (ol) {
            float h[8], l[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) split_tf32(v[j + u], h[u], l[u]);
            st_global_v8(o + j, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
            st_global_v8(ol + j, l[0], l[1], l[2], l[3], l[4], l[5], l[6], l[7]);
          } else {
            st_global_v8(o + j, v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
          }
        }
      } else if (p.out_lo) {
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
    if (F16 && p.out_lo) {
      __syncwarp();
      h16_track_amax(p.o_scale, amax);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// wgrad: dW[tap][ci][co] = sum_pixels X[pix + shift(tap)][ci] * dY[pix][co]
//   GEMM with M = (taps_per_tile x CIT input channels) = 128, N = BN output channels, K = pixels.
//   Both operands are "MN-major" (the channel index is contiguous in memory, pixels are the reduction index):
//   a TMA box (32 channels x 32 pixels) lands as 32 rows of 128 bytes = the canonical MN-major SWIZZLE_128B
//   atom stack ((8 x 16B) along MN) x (8 K-rows, 1024 B apart per K-group); the 128 (A) / BN (B) channels are
//   4 / BN/32 such boxes, LBO = 4096 B apart.  One tcgen05.mma consumes K = 8 pixels (descriptor start + 1024 B).
//   Split-K over pixel tiles across blockIdx.z; partial tiles are reduced with red.global.add.f32.
// ---------------------------------------------------------------------------------------------------------
struct WgParams {
  int PW, PH;                 // pixel box (PW*PH == 32)
  int tiles_w, tiles_h, N;    // pixel tiles per image, images
  int tiles_per_split, total_tiles;
  int cit;                    // input channels per tap inside one M tile (32, 64 or 128)
  int taps_per_tile;          // 128 / cit
  int n_taps;                 // kh*kw
  int m_tiles_per_group;      // for cit==128: Cin_pad/128 M tiles per tap; else 1
  TapEntry taps[kMaxTaps];    // dc, dw, hp, dh of the X box for each tap (b_tap unused)
  float* dw;                  // [taps][Cin][Cout]
  int Cin, Cout;
  int rowwin;                 // 1: rows are (filter row r, j = s*4+c) of the 7x7x3 first layer
};

template <int BN, int PASSES, int STAGES>
struct WgCfg {
  static constexpr uint32_t A_BYTES = 128 * 128;           // 4 boxes x (32 px x 128 B)
  static constexpr uint32_t B_BYTES = BN * 128;            // BN/32 boxes
  static constexpr uint32_t NPL = PASSES == 3 ? 2 : 1;
  static constexpr uint32_t STAGE_BYTES = (A_BYTES + B_BYTES) * NPL;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  // N concatenation (3-pass): the dY lo plane follows the hi plane in the stage, so ONE MMA of N = 2*BN against the hi
  // descriptor gives [x_hi*dy_hi | x_hi*dy_lo] in 2*BN columns (the epilogue adds the halves): x_hi is read from shared
  // memory once per k-step instead of twice -- these launches are bound by those reads, not by the tensor pipe.
#ifndef IMMB_WGT_NCAT
#define IMMB_WGT_NCAT 1
#endif
  static constexpr bool NCAT = IMMB_WGT_NCAT && PASSES == 3 && BN <= 128;
  static constexpr int ACC = NCAT ? 2 * BN : BN;
  static constexpr int TMEM_COLS = ACC <= 32 ? 32 : (ACC <= 64 ? 64 : (ACC <= 128 ? 128 : 256));
};

template <int BN, int PASSES, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap mapX_hi, const __grid_constant__ CUtensorMap mapX_lo,
                     const __grid_constant__ CUtensorMap mapY_hi, const __grid_constant__ CUtensorMap mapY_lo,
                     const __grid_constant__ WgParams p) {
  using Cfg = WgCfg<BN, PASSES, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // M tile -> (first tap, channel offset)
  int tap0, ci0;
  if (p.cit == 128) {
    tap0 = blockIdx.x / p.m_tiles_per_group;
    ci0 = (blockIdx.x % p.m_tiles_per_group) * 128;
  } else {
    tap0 = blockIdx.x * p.taps_per_tile;
    ci0 = 0;
  }
  const int n_off = blockIdx.y * BN;
  const int t_begin = blockIdx.z * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.total_tiles) t_end = p.total_tiles;
  const int num_k = t_end - t_begin;
  const int chunks_per_tap = p.cit / 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (num_k > 0) {
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int tile = t_begin + kt;
        const int twi = tile % p.tiles_w;
        const int thi = (tile / p.tiles_w) % p.tiles_h;
        const int n = tile / (p.tiles_w * p.tiles_h);
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        if (elect_one()) {
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        // A: 4 boxes of 32 channels (taps beyond the filter are loaded with an out-of-range image index -> zeros)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int tl = j / chunks_per_tap, cj = j - tl * chunks_per_tap;
          const int tap = tap0 + tl;
          const bool valid = tap < p.n_taps;
          const TapEntry t = p.taps[valid ? tap : 0];
          const int c = ci0 + cj * 32 + t.dc, w = twi * p.PW + t.dw, h = thi * p.PH + t.dh;
          const int nn = valid ? n : p.N;      // n == N is out of bounds -> TMA zero fill
          tma_load_5d(st + j * 4096, &mapX_hi, &full[s], c, w, t.hp, h, nn);
          if (PASSES == 3) tma_load_5d(st + Cfg::A_BYTES + j * 4096, &mapX_lo, &full[s], c, w, t.hp, h, nn);
        }
        uint8_t* sb = st + Cfg::A_BYTES * Cfg::NPL;
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
          tma_load_5d(sb + j * 4096, &mapY_hi, &full[s], n_off + j * 32, twi * p.PW, 0, thi * p.PH, n);
          if (PASSES == 3)
            tma_load_5d(sb + Cfg::B_BYTES + j * 4096, &mapY_lo, &full[s], n_off + j * 32, twi * p.PW, 0, thi * p.PH, n);
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (num_k > 0) {
      constexpr uint32_t idesc = idesc_tf32(kTileM, BN, 1, 1);
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t b_hi = a_hi + Cfg::A_BYTES * Cfg::NPL;
        if (elect_one()) {
        // MN-major tf32: SWIZZLE_128B_BASE32B; 32-channel groups are LBO = 4096 B apart (one TMA box each),
        // 4-pixel K atoms are SBO = 512 B apart.  One descriptor per operand and stage; k-steps add to the address field.
        const uint64_t da0_hi = smem_desc_sw128(a_hi, 4096, 512, 1), db0_hi = smem_desc_sw128(b_hi, 4096, 512, 1);
        const uint64_t da0_lo = da0_hi + (uint64_t)(Cfg::A_BYTES >> 4), db0_lo = db0_hi + (uint64_t)(Cfg::B_BYTES >> 4);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint64_t ko = (uint64_t)(k4 * (1024 >> 4));     // 8 pixels = 8 rows x 128 B
          const uint64_t da_hi = da0_hi + ko;
          const uint64_t db_hi = db0_hi + ko;
          if (Cfg::NCAT) {
            const uint64_t da_lo = da0_lo + ko;
            mma_tf32(tmem_base, da_hi, db_hi, idesc_tf32(kTileM, 2 * BN, 1, 1), (kt > 0 || k4 > 0) ? 1u : 0u);
            mma_tf32(tmem_base, da_lo, db_hi, idesc, 1u);
          } else {
          mma_tf32(tmem_base, da_hi, db_hi, idesc, (kt > 0 || k4 > 0) ? 1u : 0u);
          if (PASSES == 3) {
            const uint64_t da_lo = da0_lo + ko;
            const uint64_t db_lo = db0_lo + ko;
            mma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
            mma_tf32(tmem_base, da_lo, db_hi, idesc, 1u);
          }
          }
        }
        mma_commit(&empty[s]);
        if (kt == num_k - 1) mma_commit(tmem_full);
        }
        __syncwarp();
      }
    }
  } else if (num_k > 0) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int tl = m / p.cit;
    const int tap = tap0 + tl;
    const int ci = ci0 + (m - tl * p.cit);
    bool row_ok = (tap < p.n_taps) && (ci < p.Cin);
    size_t row_off = ((size_t)tap * p.Cin + ci) * p.Cout;
    if (p.rowwin) {      // dw[r][s][c][co] with j = s*4 + c
      row_ok = (tap < 7) && (ci < 28) && ((ci & 3) < 3);
      row_off = ((size_t)(tap * 7 + (ci >> 2)) * 3 + (ci & 3)) * p.Cout;
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      if (Cfg::NCAT) {
        float v2[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += v2[j];
      } else {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      }
      if (!row_ok) continue;
      const int col0 = n_off + c0;
      if (col0 >= p.Cout) continue;
      float* o = p.dw + row_off + col0;
if ((p.Cout & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (col0 + j < p.Cout) red_add_v4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.Cout) atomicAdd(o + j, v[j]);
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// =========================================================================================================
// host side
// =========================================================================================================
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                    const uint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int esize = 4) {
  auto fn = get_encode_fn();
  if (!fn) return set_error(IMMB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                  const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_b),
                  reinterpret_cast<const cuuint32_t*>(box), reinterpret_cast<const cuuint32_t*>(estr),
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(IMMB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return IMMB_OK;
}

// 5-D activation view (c, w, hpar, h, n) of an NHWC tensor [N,H,W,cs] with `C` valid channels.
// parity_split: (c + wpar*C, w/2, hpar, h/2, n) view used by stride-2 convolutions (needs cs == C, even H, W).
static int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int cs, bool parity_split,
                        int box_w, int box_h, int box_n, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
                        int esize = 4) {
  uint64_t dims[5], str[4];
  const uint64_t es = (uint64_t)esize;
  const uint32_t row_bytes = swz == CU_TENSOR_MAP_SWIZZLE_64B ? 64 : 128;
  uint32_t box[5] = {(uint32_t)(row_bytes / esize), (uint32_t)box_w, 1, (uint32_t)box_h, (uint32_t)box_n};
  if (!parity_split) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
    str[0] = (uint64_t)cs * es; str[1] = (uint64_t)W * cs * es; str[2] = (uint64_t)W * cs * es;
    str[3] = (uint64_t)H * W * cs * es;
  } else {
    dims[0] = 2 * (uint64_t)C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
    str[0] = (uint64_t)2 * C * es; str[1] = (uint64_t)W * C * es; str[2] = (uint64_t)2 * W * C * es;
    str[3] = (uint64_t)H * W * C * es;
  }
  return make_map(m, base, 5, dims, str, box, swz, esize);
}

// Row-window view of the staged first-layer image [N,H,W+8,4]: element (j, w, 0, h, n) = x4[n, h, w + j/4, j%4],
// i.e. the 8 pixels x 4 channels starting at padded column w are ONE 128-byte row (overlapping windows: the w
// stride is 16 bytes).  Filter row r of the 7x7 conv is then a single K = 32 chunk (28 real taps + 4 zero weights).
static int make_rowwin_map(CUtensorMap* m, const float* base, int N, int H, int W, int box_w, int box_h, int box_n,
                           CUtensorMapSwizzle swz) {
  const uint64_t Wp = (uint64_t)W + 8;
  uint64_t dims[5] = {32, (uint64_t)W, 1, (uint64_t)H, (uint64_t)N};
  uint64_t str[4] = {16, Wp * 16, Wp * 16, (uint64_t)H * Wp * 16};
  uint32_t box[5] = {32, (uint32_t)box_w, 1, (uint32_t)box_h, (uint32_t)box_n};
  return make_map(m, base, 5, dims, str, box, swz);
}

// 3-D weight view (k, n, tap) of [taps][Nn][Kd] with box (32, BN, 1)
static int make_w_map(CUtensorMap* m, const void* base, int taps, int Nn, int Kd, int bn, int esize = 4) {
  uint64_t dims[3] = {(uint64_t)Kd, (uint64_t)Nn, (uint64_t)taps};
  uint64_t str[2] = {(uint64_t)Kd * esize, (uint64_t)Nn * Kd * esize};
  uint32_t box[3] = {(uint32_t)(128 / esize), (uint32_t)bn, 1};
  return make_map(m, base, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, esize);
}

// exported to conv_tc2.cu
// swizzle_mn: MN-major operands of the wgrad kernels: the TF32 path needs the 32-byte-atom variant; fp16 planes use
// the standard 128-byte swizzle for both majors
int tc_make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int cs, bool parity_split,
                    int box_w, int box_h, int box_n, int swizzle_mn, int esize) {
  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;
  if (swizzle_mn == 2) swz = CU_TENSOR_MAP_SWIZZLE_64B;                       // 64-byte rows (32 fp16 channels)
  else if (swizzle_mn && esize == 4) swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  return make_act_map(m, base, N, H, W, C, cs, parity_split, box_w, box_h, box_n, swz, esize);
}
int tc_make_w_map(CUtensorMap* m, const void* base, int taps, int Nn, int Kd, int bn, int esize) {
  return make_w_map(m, base, taps, Nn, Kd, bn, esize);
}
int tc_make_rowwin_map(CUtensorMap* m, const float* base, int N, int H, int W, int box_w, int box_h, int box_n) {
  return make_rowwin_map(m, base, N, H, W, box_w, box_h, box_n, CU_TENSOR_MAP_SWIZZLE_128B);
}
bool conv_tc2_rowwin_eligible(const immb_conv_desc* d);
int conv_tc2_rowwin_fwd(const immb_conv_desc* d, const float* x_hi, const float* x_lo, const float* wp_hi,
                        const float* wp_lo, const float* bias, int relu, float* y_hi, float* y_lo, cudaStream_t st,
                        double* stats = nullptr);
bool conv_tc2_eligible(const immb_conv_desc* d, int op);
bool conv_tc2_wgrad_eligible(const immb_conv_desc* d);
int conv_tc2_wgrad_run(const immb_conv_desc* d, const void* x_hi, const void* x_lo, const void* dy_hi,
                       const void* dy_lo, float* dw, cudaStream_t st);
int conv_tc2_run(const immb_conv_desc* d, int op, const void* act_hi, const void* act_lo, int act_c, int act_cs,
                 const void* w_hi, const void* w_lo, int w_rows, int kd, const float* bias, int relu,
                 void* out_hi, void* out_lo, int ocs, int ncols, int n_store, cudaStream_t st,
                 const void* relu_src = nullptr, int relu_cs = 0, double* stats = nullptr,
                 const Tc2BnReduce* bnr = nullptr, const Tc2Scales* sc = nullptr);
int conv_tc2_fwd_stats_rows(const immb_conv_desc* d);
int conv_tc2_dgrad_stats_rows(const immb_conv_desc* d);
int conv_tc2_pair_mode();

static inline bool is_f16(const immb_conv_desc* d) {
  return d->precision == IMMB_PREC_F16X3 || d->precision == IMMB_PREC_F16X2;
}

static bool pick_tile(int PH, int PW, int N, int* TW, int* TH, int* TN) {
  if (PW % 16 == 0 && PH % 8 == 0) { *TW = 16; *TH = 8; *TN = 1; return true; }
  if (PW == 8 && PH == 8) { *TW = 8; *TH = 8; *TN = 2; return true; }
  (void)N;
  return false;
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int pymod(int a, int b) { return a - floordiv(a, b) * b; }

bool conv_tc_eligible(const immb_conv_desc* d, int op) {
  int TW, TH, TN;
  if (is_f16(d)) {
    // fp16 planes: 16-byte pixel strides, no row-window first layer, weight gradients through the halo wgrad kernel only
    if (d->x_layout != IMMB_XLAYOUT_NHWC || d->x_cstride % 8 || d->y_cstride % 8) return false;
    if (op == 2 && !conv_tc2_wgrad_eligible(d)) return false;
  }
  if (d->x_layout == IMMB_XLAYOUT_ROWWIN4) {
    // first encoder layer: 7x7, Cin = 3, stride 1 on the staged [N,H,W+8,4] image (no dgrad: the input is data)
    if (d->kh != 7 || d->kw != 7 || d->Cin != 3 || d->stride != 1 || op == 1) return false;
    if (d->Cout % 32 || d->Cout > 128 || d->y_cstride != d->Cout) return false;
    return d->W % 16 == 0 && d->H % 8 == 0;
  }
  if (d->kh != d->kw || (d->kh != 1 && d->kh != 3)) return false;
  if (d->stride == 2 && (d->kh != 3 || (d->H % 2) || (d->W % 2) || d->x_cstride != d->Cin || d->Cin % 32)) return false;
  if (d->Cin < 1 || d->Cout < 1) return false;
  if (d->x_cstride % 4 || d->y_cstride % 4 || d->y_cstride < d->Cout) return false;
  if (d->Cout > 128 && d->Cout % 32) return false;
  if (d->cin_pad % 32 || d->cin_pad < d->Cin) return false;
  if (op == 0) {
    return pick_tile(d->Ho, d->Wo, d->N, &TW, &TH, &TN);
  } else if (op == 1) {
    if (d->stride == 1) return pick_tile(d->H, d->W, d->N, &TW, &TH, &TN);
    return pick_tile(d->H / 2, d->W / 2, d->N, &TW, &TH, &TN);
  } else {
    if (!(d->cin_pad == 32 || d->cin_pad == 64 || d->cin_pad >= 128)) return false;
    if (d->Wo % 8 || (d->Wo >= 16 && d->Wo % 16)) return false;
    int pw = d->Wo >= 16 ? 16 : 8, phh = 32 / pw;
    if (d->Ho % phh) return false;
    return true;
  }
}

// SHORT: launches whose CTAs run only a handful of K iterations (the stride-2 dgrad parity classes: <= 4 taps x 1-2
// chunks, thousands of CTAs) are bound by per-CTA set-up latency, not by the pipeline depth: two stages leave room for
// two resident CTAs per SM, which overlap each other's prologue / epilogue
template <int BN, int PASSES, bool F16 = false, bool SHORT = false>
static int launch_fwd(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                      const CUtensorMap& b_lo, const TcParams& p, dim3 grid, cudaStream_t st) {
  constexpr int STAGES = SHORT ? 2 : (PASSES == 3 ? (BN <= 64 ? 4 : 3) : (BN <= 64 ? 6 : 4));
  using Cfg = FwdCfg<BN, PASSES, STAGES>;
  auto kern = conv_tc_kernel<BN, PASSES, STAGES, F16>;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device_slot()];      // the shared-memory opt-in is a per-device attribute
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc smem attr: %s", cudaGetErrorString(e));
    if (SHORT) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(a_hi, a_lo, b_hi, b_lo, p);
  return check_launch("conv_tc_kernel");
}

static int dispatch_fwd(int bn, int passes, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                        const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcParams& p, dim3 grid,
                        cudaStream_t st, bool f16 = false) {
  // longest K loop of any CTA of this launch
  int max_k = p.n_taps * p.kchunks;
  for (int c = 1; c < p.n_classes; ++c) max_k = p.n_taps_c[c - 1] * p.kchunks > max_k ? p.n_taps_c[c - 1] * p.kchunks : max_k;
  if (passes == 3 && max_k <= 18 && bn <= 64) {
    if (f16) {
      if (bn == 16) return launch_fwd<16, 3, true, true>(a_hi, a_lo, b_hi, b_lo, p, grid, st);
      if (bn == 32) return launch_fwd<32, 3, true, true>(a_hi, a_lo, b_hi, b_lo, p, grid, st);
      if (bn == 64) return launch_fwd<64, 3, true, true>(a_hi, a_lo, b_hi, b_lo, p, grid, st);
    } else {
      if (bn == 16) return launch_fwd<16, 3, false, true>(a_hi, a_lo, b_hi, b_lo, p, grid, st);
      if (bn == 32) return launch_fwd<32, 3, false, true>(a_hi, a_lo, b_hi, b_lo, p, grid, st);
      if (bn == 64) return launch_fwd<64, 3, false, true>(a_hi, a_lo, b_hi, b_lo, p, grid, st);
    }
  }
  if (f16) {
    if (passes != 3) return set_error(IMMB_ERR_INVALID, "conv_tc: single-pass fp16 is not built");
#define IMMB_CASE16(BN_) \
  if (bn == BN_) return launch_fwd<BN_, 3, true>(a_hi, a_lo, b_hi, b_lo, p, grid, st);
    IMMB_CASE16(16)
    IMMB_CASE16(32)
    IMMB_CASE16(64)
    IMMB_CASE16(96)
    IMMB_CASE16(128)
#undef IMMB_CASE16
    return set_error(IMMB_ERR_INVALID, "conv_tc (fp16): unsupported BN %d", bn);
  }
#define IMMB_CASE(BN_)                                                                         \
  if (bn == BN_)                                                                               \
    return passes == 3 ? launch_fwd<BN_, 3>(a_hi, a_lo, b_hi, b_lo, p, grid, st)               \
                       : launch_fwd<BN_, 1>(a_hi, a_lo, b_hi, b_lo, p, grid, st);
  IMMB_CASE(16)
  IMMB_CASE(32)
  IMMB_CASE(64)
  IMMB_CASE(96)
  IMMB_CASE(128)
#undef IMMB_CASE
  return set_error(IMMB_ERR_INVALID, "conv_tc: unsupported BN %d", bn);
}

static int pick_bn(int ncols);
int tc_pick_bn(int ncols) { return pick_bn(ncols); }
static int pick_bn(int ncols) {
  if (ncols % 128 == 0) return 128;
  if (ncols % 96 == 0) return 96;
  if (ncols % 64 == 0) return 64;
  if (ncols % 32 == 0) return 32;
  if (ncols <= 16) return 16;         // small / ragged channel counts: one tile, TMA zero-fills the missing rows
  if (ncols <= 32) return 32;
  if (ncols <= 64) return 64;
  if (ncols <= 96) return 96;
  return 128;
}

int conv_tc_fwd_stats_rows(const immb_conv_desc* d) { return conv_tc2_fwd_stats_rows(d); }

// number of 32-byte k-steps of the last 128-byte K chunk that hold weights (fp16: 16 channels per k-step)
static inline int k_last_steps(int kd, bool f16) {
  if (!f16) return 4;
  const int rem = kd - (ceil_div(kd, 64) - 1) * 64;
  return ceil_div(rem, 16);
}

int conv_tc_fwd(const immb_conv_desc* d, const void* x_hi, const void* x_lo, const void* wp_hi,
                const void* wp_lo, const float* bias, void* y_hi, void* y_lo, cudaStream_t st, double* stats) {
  const bool f16 = is_f16(d);
  const int esize = f16 ? 2 : 4, kchunk = f16 ? 64 : 32;
  const Tc2Scales sc{d->x_scale, d->w_scale, d->y_scale};
  if (f16 && (!sc.a || !sc.b || (y_lo && !sc.o)))
    return set_error(IMMB_ERR_INVALID, "conv_tc_fwd: fp16 planes need x_scale / w_scale (and y_scale for plane outputs)");
  if (stats && conv_tc2_fwd_stats_rows(d) == 0)
    return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc_fwd: fused BN statistics are not available for this shape");
  if (conv_tc2_eligible(d, 0))
    return conv_tc2_run(d, 0, x_hi, x_lo, d->Cin, d->x_cstride, wp_hi, wp_lo, d->Cout, d->cin_pad, bias,
                        d->epilogue == IMMB_EPI_BIAS_RELU, y_hi, y_lo, d->y_cstride, d->Cout,
                        f16 ? (d->Cout + 7) / 8 * 8 : (d->Cout + 3) / 4 * 4, st, nullptr, 0, stats, nullptr, &sc);
  if (conv_tc2_rowwin_eligible(d))
    return conv_tc2_rowwin_fwd(d, (const float*)x_hi, (const float*)x_lo, (const float*)wp_hi, (const float*)wp_lo, bias,
                               d->epilogue == IMMB_EPI_BIAS_RELU, (float*)y_hi, (float*)y_lo, st, stats);
  const int passes = d->precision == IMMB_PREC_TF32 ? 1 : 3;
  TcParams p;
  memset(&p, 0, sizeof(p));
  if (!pick_tile(d->Ho, d->Wo, d->N, &p.TW, &p.TH, &p.TN)) return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc_fwd: tile");
  p.tiles_w = d->Wo / p.TW; p.tiles_h = d->Ho / p.TH;
  const bool rowwin = d->x_layout == IMMB_XLAYOUT_ROWWIN4;
  if (rowwin) {
    p.n_taps = 7; p.kchunks = 1;
    for (int r = 0; r < 7; ++r) { TapEntry& t = p.taps[r]; t.b_tap = r; t.dc = 0; t.dw = 0; t.hp = 0; t.dh = r - d->pad_t; }
  } else {
    p.n_taps = d->kh * d->kw; p.kchunks = ceil_div(d->cin_pad, kchunk);
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) {
        TapEntry& t = p.taps[r * d->kw + s];
        t.b_tap = r * d->kw + s;
        int th = r - d->pad_t, tw = s - d->pad_l;
        if (d->stride == 1) { t.dc = 0; t.dw = tw; t.hp = 0; t.dh = th; }
        else { t.dh = floordiv(th, 2); t.hp = pymod(th, 2); t.dw = floordiv(tw, 2); t.dc = pymod(tw, 2) * d->Cin; }
      }
  }
  p.out_hi = (float*)y_hi; p.out_lo = (float*)y_lo; p.bias = bias; p.relu = d->epilogue == IMMB_EPI_BIAS_RELU;
  p.N = d->N; p.PH = d->Ho; p.PW = d->Wo; p.OH = d->Ho; p.OW = d->Wo; p.ocs = d->y_cstride;
  p.out_mul = 1; p.out_add_h = 0; p.out_add_w = 0; p.n_cols = d->Cout;
  p.n_store = f16 ? (d->Cout + 7) / 8 * 8 : (d->Cout + 3) / 4 * 4;      // fp16 layers: 8-channel output strides
  p.a_scale = sc.a; p.b_scale = sc.b; p.o_scale = sc.o;
  const int bn = pick_bn(d->Cout);
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  const bool split = d->stride == 2;
  const int kd = rowwin ? 32 : d->cin_pad;
  p.k_last = k_last_steps(kd, f16);
  auto amap = [&](CUtensorMap* m, const void* base) {
    return rowwin ? make_rowwin_map(m, (const float*)base, d->N, d->H, d->W, p.TW, p.TH, p.TN, CU_TENSOR_MAP_SWIZZLE_128B)
                  : make_act_map(m, base, d->N, d->H, d->W, d->Cin, d->x_cstride, split, p.TW, p.TH, p.TN,
                                 CU_TENSOR_MAP_SWIZZLE_128B, esize);
  };
  if ((rc = amap(&a_hi, x_hi))) return rc;
  if ((rc = make_w_map(&b_hi, wp_hi, p.n_taps, d->Cout, kd, bn, esize))) return rc;
  a_lo = a_hi; b_lo = b_hi;
  if (passes == 3) {
    if ((rc = amap(&a_lo, x_lo))) return rc;
    if ((rc = make_w_map(&b_lo, wp_lo, p.n_taps, d->Cout, kd, bn, esize))) return rc;
  }
  dim3 grid(p.tiles_w * p.tiles_h * ceil_div(d->N, p.TN), ceil_div(d->Cout, bn));
  return dispatch_fwd(bn, passes, a_hi, a_lo, b_hi, b_lo, p, grid, st, f16);
}

// dgrad fused with the backward of the ReLU that produced the conv's input and with the operand split of the result:
// out = split(dgrad(dy) * [act_hi > 0]).  Halo pair kernel only (stride-1 3x3, H % 16 == 0, W % 16 == 0).
bool conv_tc_dgrad_relu_eligible(const immb_conv_desc* d) {
  return conv_tc_eligible(d, 1) && conv_tc2_eligible(d, 1) && conv_tc2_pair_mode() == 1 && d->x_cstride % 8 == 0;
}
int conv_tc_dgrad_relu(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo, const void* wh_hi,
                       const void* wh_lo, const void* act_hi, int act_cs, void* out_hi, void* out_lo,
                       int32_t* out_scale, cudaStream_t st) {
  const int ncols = d->cin_pad < d->x_cstride ? d->cin_pad : d->x_cstride;
  const Tc2Scales sc{d->y_scale, d->w_scale, out_scale};
  return conv_tc2_run(d, 1, dy_hi, dy_lo, d->Cout, d->y_cstride, wh_hi, wh_lo, d->cin_pad, d->y_cstride, nullptr, 0,
                      out_hi, out_lo, d->x_cstride, ncols, ncols, st, act_hi, act_cs, nullptr, nullptr, &sc);
}

// dgrad that also accumulates, in its epilogue, the BN-backward sums of the layer that produced its input
int conv_tc_dgrad_stats_rows(const immb_conv_desc* d) { return conv_tc_eligible(d, 1) ? conv_tc2_dgrad_stats_rows(d) : 0; }
int conv_tc_dgrad_bnreduce(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo, const void* wh_hi,
                           const void* wh_lo, float* dx, const float* y_prev, int y_prev_cs, const float* scale,
                           const float* shift, const float* mean, const float* invstd, int relu, double* partials,
                           cudaStream_t st) {
  const int ncols = d->cin_pad < d->x_cstride ? d->cin_pad : d->x_cstride;
  Tc2BnReduce b{y_prev, y_prev_cs, relu, scale, shift, mean, invstd};
  const Tc2Scales sc{d->y_scale, d->w_scale, nullptr};
  return conv_tc2_run(d, 1, dy_hi, dy_lo, d->Cout, d->y_cstride, wh_hi, wh_lo, d->cin_pad, d->y_cstride, nullptr, 0,
                      dx, nullptr, d->x_cstride, ncols, ncols, st, nullptr, 0, partials, &b, &sc);
}

int conv_tc_dgrad(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo, const void* wh_hi,
                  const void* wh_lo, float* dx, cudaStream_t st) {
  const bool f16 = is_f16(d);
  const int esize = f16 ? 2 : 4, kchunk = f16 ? 64 : 32;
  const Tc2Scales sc{d->y_scale, d->w_scale, nullptr};
  if (f16 && (!sc.a || !sc.b)) return set_error(IMMB_ERR_INVALID, "conv_tc_dgrad: fp16 planes need y_scale / w_scale");
  const int passes = d->precision == IMMB_PREC_TF32 ? 1 : 3;
  const int ncols = d->cin_pad < d->x_cstride ? d->cin_pad : d->x_cstride;   // channels of dx that get written
  if (conv_tc2_eligible(d, 1))
    return conv_tc2_run(d, 1, dy_hi, dy_lo, d->Cout, d->y_cstride, wh_hi, wh_lo, d->cin_pad, d->y_cstride, nullptr, 0,
                        dx, nullptr, d->x_cstride, ncols, ncols, st, nullptr, 0, nullptr, nullptr, &sc);
  const int bn = pick_bn(ncols);
  const int classes = d->stride == 1 ? 1 : 4;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.PH = d->H / d->stride; p.PW = d->W / d->stride;
  if (!pick_tile(p.PH, p.PW, d->N, &p.TW, &p.TH, &p.TN)) return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc_dgrad: tile");
  p.tiles_w = p.PW / p.TW; p.tiles_h = p.PH / p.TH;
  p.kchunks = ceil_div(d->y_cstride, kchunk);
  p.k_last = k_last_steps(d->y_cstride, f16);
  p.a_scale = sc.a; p.b_scale = sc.b; p.o_scale = nullptr;
  p.n_classes = classes;
  for (int cls = 0; cls < classes; ++cls) {
    const int phh = cls >> 1, pww = cls & 1;
    int nt = 0;
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) {
        int eh = phh + d->pad_t - r, ew = pww + d->pad_l - s;      // dy index = tile index + e/stride
        if (d->stride == 2 && ((eh & 1) || (ew & 1))) continue;
        if (cls > 0 && nt >= 4) return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc_dgrad: too many taps in a parity class");
        TapEntry& t = cls == 0 ? p.taps[nt] : p.taps_c[cls - 1][nt];
        ++nt;
        t.b_tap = r * d->kw + s; t.dc = 0; t.hp = 0;
        t.dh = d->stride == 1 ? eh : floordiv(eh, 2);
        t.dw = d->stride == 1 ? ew : floordiv(ew, 2);
      }
    if (cls == 0) { p.n_taps = nt; p.out_add_h = phh; p.out_add_w = pww; }
    else { p.n_taps_c[cls - 1] = nt; p.out_add_c[cls - 1][0] = phh; p.out_add_c[cls - 1][1] = pww; }
  }
  p.out_hi = dx; p.out_lo = nullptr; p.bias = nullptr; p.relu = 0;
  p.N = d->N; p.OH = d->H; p.OW = d->W; p.ocs = d->x_cstride;
  p.out_mul = d->stride; p.n_cols = ncols; p.n_store = ncols;
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if ((rc = make_act_map(&a_hi, dy_hi, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, p.TW, p.TH, p.TN, sw, esize))) return rc;
  if ((rc = make_w_map(&b_hi, wh_hi, d->kh * d->kw, d->cin_pad, d->y_cstride, bn, esize))) return rc;
  a_lo = a_hi; b_lo = b_hi;
  if (passes == 3) {
    if ((rc = make_act_map(&a_lo, dy_lo, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, p.TW, p.TH, p.TN, sw, esize))) return rc;
    if ((rc = make_w_map(&b_lo, wh_lo, d->kh * d->kw, d->cin_pad, d->y_cstride, bn, esize))) return rc;
  }
  dim3 grid(p.tiles_w * p.tiles_h * ceil_div(d->N, p.TN), ceil_div(ncols, bn), classes);
  return dispatch_fwd(bn, passes, a_hi, a_lo, b_hi, b_lo, p, grid, st, f16);
}

size_t conv_tc_wgrad_workspace(const immb_conv_desc*) { return 0; }

static constexpr CUtensorMapSwizzle kSwzMN = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;

template <int BN, int PASSES>
static int launch_wg(const CUtensorMap& x_hi, const CUtensorMap& x_lo, const CUtensorMap& y_hi,
                     const CUtensorMap& y_lo, const WgParams& p, dim3 grid, cudaStream_t st) {
  constexpr int STAGES = PASSES == 3 ? (BN <= 64 ? 4 : 3) : (BN <= 64 ? 6 : 4);
  using Cfg = WgCfg<BN, PASSES, STAGES>;
  auto kern = conv_tc_wgrad_kernel<BN, PASSES, STAGES>;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device_slot()];      // the shared-memory opt-in is a per-device attribute
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc_wgrad smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(x_hi, x_lo, y_hi, y_lo, p);
  return check_launch("conv_tc_wgrad_kernel");
}

int conv_tc_wgrad(const immb_conv_desc* d, const void* x_hi_, const void* x_lo_, const void* dy_hi_,
                  const void* dy_lo_, float* dw, void*, size_t, cudaStream_t st) {
  if (conv_tc2_wgrad_eligible(d)) return conv_tc2_wgrad_run(d, x_hi_, x_lo_, dy_hi_, dy_lo_, dw, st);
  if (is_f16(d)) return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc_wgrad: fp16 planes are served by the halo wgrad kernel only");
  const float *x_hi = (const float*)x_hi_, *x_lo = (const float*)x_lo_, *dy_hi = (const float*)dy_hi_, *dy_lo = (const float*)dy_lo_;
  const int passes = d->precision == IMMB_PREC_TF32 ? 1 : 3;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.PW = d->Wo >= 16 ? 16 : 8; p.PH = 32 / p.PW;
  p.tiles_w = d->Wo / p.PW; p.tiles_h = d->Ho / p.PH; p.N = d->N;
  p.total_tiles = p.tiles_w * p.tiles_h * d->N;
  const bool rowwin = d->x_layout == IMMB_XLAYOUT_ROWWIN4;
  const int cin_pad = rowwin ? 32 : d->cin_pad;
  p.rowwin = rowwin ? 1 : 0;
  p.n_taps = rowwin ? 7 : d->kh * d->kw;
  p.cit = cin_pad >= 128 ? 128 : (cin_pad >= 64 ? 64 : 32);
  if (cin_pad < 128 && cin_pad != p.cit) p.cit = 32;       // e.g. cin_pad 96 -> 32-channel groups
  p.taps_per_tile = 128 / p.cit;
  p.m_tiles_per_group = p.cit == 128 ? ceil_div(cin_pad, 128) : ceil_div(cin_pad, p.cit);
  if (rowwin) {
    for (int r = 0; r < 7; ++r) { TapEntry& t = p.taps[r]; t.b_tap = r; t.dc = 0; t.dw = 0; t.hp = 0; t.dh = r - d->pad_t; }
  } else {
    for (int r = 0; r < d->kh; ++r)
      for (int s = 0; s < d->kw; ++s) {
        TapEntry& t = p.taps[r * d->kw + s];
        t.b_tap = r * d->kw + s;
        int th = r - d->pad_t, tw = s - d->pad_l;
        if (d->stride == 1) { t.dc = 0; t.dw = tw; t.hp = 0; t.dh = th; }
        else { t.dh = floordiv(th, 2); t.hp = pymod(th, 2); t.dw = floordiv(tw, 2); t.dc = pymod(tw, 2) * d->Cin; }
      }
  }
  p.dw = dw; p.Cin = d->Cin; p.Cout = d->Cout;
  int m_tiles;
  if (p.cit == 128) m_tiles = p.n_taps * p.m_tiles_per_group;
  else {
    if (cin_pad != p.cit) return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc_wgrad: cin_pad %d", cin_pad);
    m_tiles = ceil_div(p.n_taps, p.taps_per_tile);
  }
  const int bn = d->Cout % 128 == 0 ? 128 : (d->Cout % 64 == 0 ? 64 : 32);
  const int n_tiles = ceil_div(d->Cout, bn);
  int splits = ceil_div(kNumSMs * 2, m_tiles * n_tiles);
  if (splits > p.total_tiles) splits = p.total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = ceil_div(p.total_tiles, splits);
  splits = ceil_div(p.total_tiles, p.tiles_per_split);
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->kh * d->kw * d->Cin * d->Cout, st);
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "wgrad memset: %s", cudaGetErrorString(e));
  CUtensorMap mx_hi, mx_lo, my_hi, my_lo;
  int rc;
  const bool split = d->stride == 2;
  auto xmap = [&](CUtensorMap* m, const float* base) {
    return rowwin ? make_rowwin_map(m, base, d->N, d->H, d->W, p.PW, p.PH, 1, kSwzMN)
                  : make_act_map(m, base, d->N, d->H, d->W, d->Cin, d->x_cstride, split, p.PW, p.PH, 1, kSwzMN);
  };
  if ((rc = xmap(&mx_hi, x_hi))) return rc;
  if ((rc = make_act_map(&my_hi, dy_hi, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, p.PW, p.PH, 1, kSwzMN))) return rc;
  mx_lo = mx_hi; my_lo = my_hi;
  if (passes == 3) {
    if ((rc = xmap(&mx_lo, x_lo))) return rc;
    if ((rc = make_act_map(&my_lo, dy_lo, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, p.PW, p.PH, 1, kSwzMN))) return rc;
  }
  dim3 grid(m_tiles, n_tiles, splits);
#define IMMB_WG(BN_)                                                                            \
  if (bn == BN_)                                                                                \
    return passes == 3 ? launch_wg<BN_, 3>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st)             \
                       : launch_wg<BN_, 1>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
  IMMB_WG(32)
  IMMB_WG(64)
  IMMB_WG(128)
#undef IMMB_WG
  return set_error(IMMB_ERR_INVALID, "conv_tc_wgrad: bn");
}

}  // namespace immb
