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
#include <type_traits>
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
  int m_tiles, total_pairs;      // cluster mode: ceil(m_tiles / 2) * n_tiles_n pair iterations
  int tw_sh, th_sh, nn_sh;       // log2 + 1 of tiles_w / tiles_h / n_tiles_n when a power of two, else 0 (pair kernel: the
                                 // tile decode is on the epilogue's per-tile critical path; three divisions cost ~100 instructions)
  int kchunks;
  Tc2Tap taps[9];
  float* out_hi;
  float* out_lo;
  const float* bias;
  int relu;
  int H, W, ocs, n_cols, n_store;
  int bo_mode;               // debug (IMMB_TC2_BO): 0 = base_offset 0 (correct), 1 = base_offset s
  // pair kernel only: geometry of the activation box (defaults = the 3x3 halo box)
  int n_taps;                // 9 (3x3 halo) or 7 (filter rows of the 7x7 first layer on the row-window view)
  int a_sbo;                 // bytes between the 8-row core-matrix groups of the A operand = box width * 128
  int a_plane_bytes;         // bytes one TMA box deposits per plane (expect_tx)
  int box_dw, box_dh;        // box origin relative to the tile origin (-1,-1 for 3x3 SAME; 0,-3 for the row-window view)
  uint32_t a_off[9];         // byte offset of each tap's first row inside the box
  const float* relu_src;     // optional [.., relu_cs] tensor laid out like the output: out = (relu_src > 0) ? out : 0
  int relu_cs;               // (backward of the ReLU of the layer that produced the dgrad's input, fused)
  double* stats;             // optional per-channel reduction partials: row (cta*4 + epilogue warp) of [2][n_cols] doubles
  int stats_mode;            // 1: forward BN moments (sum y, sum y^2);  2: BN-backward sums of the layer that produced
                             // this dgrad's input (sum dz, sum dz*xhat), dz = g * relu'(z), xhat = (y - mean) * invstd
  const float* bnr_y;        // mode 2: that layer's raw conv output y [.., bnr_ycs] and its BN constants [n_cols]
  int bnr_ycs, bnr_relu;
  const float *bnr_scale, *bnr_shift, *bnr_mean, *bnr_invstd;
  // scaled-fp16 operands (F16 kernels): scale records {e, amax bits} of the activation / weight operands and, when the
  // output is written as H16 planes, of the output tensor; k-steps (of 32 bytes) to issue in the LAST K chunk
  const int32_t* a_scale;
  const int32_t* b_scale;
  int32_t* o_scale;
  int k_last;
  // pair kernel: the whole weight operand of the layer (n_taps x kchunks slices of this CTA's rows) fits the weight
  // ring and there is a single N tile: it is loaded ONCE per CTA and stays resident for all of the CTA's tiles
  int b_resident;
};

template <int BN, int PASSES>
struct Tc2Cfg {
  // PASSES: 1 = hi*hi; 2 = hi*hi + lo*hi (weight operand exactly TF32); 3 = hi*hi + hi*lo + lo*hi
  static constexpr uint32_t NPLA = PASSES >= 2 ? 2 : 1;       // activation planes
  static constexpr uint32_t NPLB = PASSES == 3 ? 2 : 1;       // weight planes
  static constexpr uint32_t A_PLANE = 18 * 16 * 128;          // 36864 B: (16+2) halo rows x 16 pixels x 32 ch
  static constexpr uint32_t A_SLOT = A_PLANE * NPLA;
  static constexpr uint32_t A_SLOTS = 2;
  static constexpr uint32_t B_PLANE = BN * 128;
  static constexpr uint32_t B_SLOT = B_PLANE * NPLB;
  static constexpr uint32_t B_SLOTS = (B_SLOT <= 16384) ? 4 : ((A_SLOTS * A_SLOT + 3 * B_SLOT <= 225000) ? 3 : 2);
  static constexpr uint32_t SMEM_BYTES = A_SLOTS * A_SLOT + B_SLOTS * B_SLOT + 1024 + 256;
  // 3-pass: ONE MMA of N = 2*BN against the contiguous [w_hi ; w_lo] rows produces hi*hi in columns [0,BN) and
  // hi*lo in [BN,2BN) (the A_hi tile is read from shared memory once instead of twice); lo*hi (N = BN) accumulates
  // into [0,BN); the epilogue adds the two column blocks.
  static constexpr int BNP = (BN + 31) / 32 * 32;
  static constexpr int ACC_COLS = PASSES == 3 ? 2 * BNP : BNP;
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512));
  static_assert(PASSES != 3 || BN % 32 == 0 || BN == 16, "3-pass N-concatenation needs w_lo to start right after w_hi");
};

// CL = 2: the CTA pair of a 2-CTA cluster works on two M tiles of the SAME N tile in lockstep; each CTA fetches half
// of every weight slice and TMA-multicasts it into both CTAs' shared memory, which halves the L2 -> smem traffic of
// the weight operand (the limiter of the big layers).  b_empty then counts 2 arrivals (own + peer MMA commits).
template <int BN, int PASSES, int CL>
__global__ void __launch_bounds__(192, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                const __grid_constant__ Tc2Params p) {
  using Cfg = Tc2Cfg<BN, PASSES>;
  const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u;
  const int tile0 = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // first pair-tile / tile of this CTA
  const int tstep = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_iter_total = CL == 2 ? p.total_pairs : p.total_tiles;
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

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (uint32_t i = 0; i < Cfg::B_SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA_hi);
    prefetch_tmap(&mapB_hi);
    if (PASSES >= 2) prefetch_tmap(&mapA_lo);
    if (PASSES == 3) prefetch_tmap(&mapB_lo);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();                   // peer barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // iteration -> (m tile, n tile); with CL == 2 the pair shares the n tile and CTA `crank` takes m tile 2*jm + crank
  // (clamped to the last m tile when the count is odd: that CTA recomputes it with stores disabled)
  auto decode = [&](int it, int& img, int& th, int& tw, int& n_off, bool& live) {
    const int nt = it % p.n_tiles_n;
    int mt = it / p.n_tiles_n;
    live = true;
    if (CL == 2) {
      mt = mt * 2 + (int)crank;
      if (mt >= p.m_tiles) { mt = p.m_tiles - 1; live = false; }
    }
    tw = mt % p.tiles_w;
    int r = mt / p.tiles_w;
    th = r % p.tiles_h;
    img = r / p.tiles_h;
    n_off = nt * BN;
  };

  if (warp == 0) {
    {
      // ===== TMA producer (whole warp runs the loop; one elected lane issues) =====
      uint32_t ai = 0, bi = 0;                       // running slot counters
      for (int t = tile0; t < n_iter_total; t += tstep) {
        int img, th, tw, n_off;
        bool live;
        decode(t, img, th, tw, n_off, live);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const uint32_t as = ai % Cfg::A_SLOTS, aph = (ai / Cfg::A_SLOTS) & 1;
          mbar_wait(&a_empty[as], aph ^ 1);
          uint8_t* sa = a_base + as * Cfg::A_SLOT;
          if (elect_one()) {
          mbar_expect_tx(&a_full[as], Cfg::A_SLOT);
          tma_load_5d(sa, &mapA_hi, &a_full[as], kc * 32, tw * 8 - 1, 0, th * 16 - 1, img);
          if (PASSES >= 2) tma_load_5d(sa + Cfg::A_PLANE, &mapA_lo, &a_full[as], kc * 32, tw * 8 - 1, 0, th * 16 - 1, img);
          }
          __syncwarp();
          ++ai;
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t bs = bi % Cfg::B_SLOTS, bph = (bi / Cfg::B_SLOTS) & 1;
            mbar_wait(&b_empty[bs], bph ^ 1);
            uint8_t* sb = b_base + bs * Cfg::B_SLOT;
            if (elect_one()) {
            mbar_expect_tx(&b_full[bs], Cfg::B_SLOT);
            if (CL == 2) {
              // my half of the rows, into both CTAs
              const uint32_t ho = crank * (Cfg::B_PLANE / 2);
              const int row0 = n_off + (int)crank * (BN / 2);
              tma_load_3d_mc(sb + ho, &mapB_hi, &b_full[bs], kc * 32, row0, p.taps[tap].b_tap, (uint16_t)3);
              if (PASSES == 3)
                tma_load_3d_mc(sb + Cfg::B_PLANE + ho, &mapB_lo, &b_full[bs], kc * 32, row0, p.taps[tap].b_tap, (uint16_t)3);
            } else {
              tma_load_3d(sb, &mapB_hi, &b_full[bs], kc * 32, n_off, p.taps[tap].b_tap);
              if (PASSES == 3) tma_load_3d(sb + Cfg::B_PLANE, &mapB_lo, &b_full[bs], kc * 32, n_off, p.taps[tap].b_tap);
            }
            }
            __syncwarp();
            ++bi;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer (whole warp runs the loop with uniform values; one elected lane issues) =====
      constexpr uint32_t idesc = idesc_tf32(128, BN, 0, 0);
      constexpr uint32_t idesc2 = idesc_tf32(128, 2 * BN, 0, 0);      // [w_hi ; w_lo]
      uint32_t ai = 0, bi = 0, ti = 0;
      for (int t = tile0; t < n_iter_total; t += tstep) {
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
            if (elect_one()) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint32_t ko = k4 * 32;
              const uint64_t da_hi = smem_desc_sw128(a_hi + a_off + ko, 16, 2048, 2, bo);
              const uint64_t db_hi = smem_desc_sw128(b_hi + ko, 16, 1024);
              mma_tf32(tmem_d, da_hi, db_hi, PASSES == 3 ? idesc2 : idesc, (kc > 0 || tap > 0 || k4 > 0) ? 1u : 0u);
              if (PASSES >= 2) {
                const uint64_t da_lo = smem_desc_sw128(a_lo + a_off + ko, 16, 2048, 2, bo);
                mma_tf32(tmem_d, da_lo, db_hi, idesc, 1u);
              }
            }
            if (CL == 2) mma_commit_mc(&b_empty[bs], (uint16_t)3);
            else mma_commit(&b_empty[bs]);
            if (tap == 8) mma_commit(&a_empty[as]);
            if (tap == 8 && kc == p.kchunks - 1) mma_commit(&t_full[acc]);
            }
            __syncwarp();
            ++bi;
          }
          ++ai;
        }
        ++ti;
      }
    }
  } else {
    // ===== epilogue (warps 2..5) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int hl = m >> 3, wl = m & 7;
    uint32_t ti = 0;
    for (int t = tile0; t < n_iter_total; t += tstep) {
      int img, th, tw, n_off;
      bool live;
      decode(t, img, th, tw, n_off, live);
      const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
      const int h = th * 16 + hl, w = tw * 8 + wl;
      const size_t pix = ((size_t)img * p.H + h) * p.W + w;
      mbar_wait(&t_full[acc], tph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + (uint32_t)c0, v);
        if (PASSES == 3) {
          float v2[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + (uint32_t)(BN + c0), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        }
        const int col0 = n_off + c0;
        if (col0 >= p.n_cols || !live) continue;
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
  if (CL == 2) cluster_sync_all();                   // nobody exits while the peer may still signal / write into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// =============================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a 2-CTA cluster (one TPC) execute ONE MMA of
// M = 256 pixels (two M tiles, CTA r owns tile 2*jm + r: its own halo box, its own 128 TMEM lanes) x N = BN2 output
// channels.  The weight operand is SPLIT across the pair: CTA r fetches and keeps only rows [r*BN2/2, (r+1)*BN2/2) of
// every weight slice, so per output the L2 -> SM weight traffic AND the per-SM shared-memory operand reads are half of
// what two independent CTAs need, and N up to 256 fits (one activation halo serves 256 output channels).
// Why: at M = 128 / N = 128 every SM has to ingest ~48-52 B/clk of operands to keep the TF32 pipe busy -- above the
// ~42 B/clk/SM the L2 delivers chip-wide -- so the single-CTA kernel is L2-fed-bound on every wide layer.
// Protocol: tc_ptx.cuh "CTA pair".  full barriers live in the leader (expect 2x the bytes: own + peer loads);
// empty / accumulator-full barriers are per CTA and receive the leader's multicast commits; the accumulator-empty
// barrier lives in the leader and counts the 4 epilogue warps of BOTH CTAs.
// =============================================================================================================
// (value, compensation) fp32 pair accumulation (Knuth TwoSum; -fmad / reassociation cannot touch plain adds and subs)
__device__ __forceinline__ float2 two_sum_acc(float2 a, float x) {
  const float t = __fadd_rn(a.x, x);
  const float bp = __fadd_rn(t, -a.x);
  const float e = __fadd_rn(__fadd_rn(a.x, -__fadd_rn(t, -bp)), __fadd_rn(x, -bp));
  return make_float2(t, __fadd_rn(a.y, e));
}
__device__ __forceinline__ double pair_value(double slot) {       // the double slot holds a float2 pair
  const float2 v = *reinterpret_cast<const float2*>(&slot);
  return (double)v.x + (double)v.y;
}

template <int BN2, int PASSES, bool F16 = false>
struct Tc2PairCfg {
  static constexpr uint32_t NPLA = PASSES >= 2 ? 2 : 1;
  static constexpr uint32_t NPLB = PASSES == 3 ? 2 : 1;
  // one activation plane of a slot: the 18 x 10-pixel halo box (23040 B) or the first layer's 22 x 8 row-window box
  // (22528 B), rounded up to the 1024-byte swizzle pattern
  static constexpr uint32_t A_PLANE = 23 * 1024;
  static constexpr uint32_t A_SLOT = A_PLANE * NPLA;
#ifndef IMMB_PAIR_A_SLOTS
#define IMMB_PAIR_A_SLOTS 3
#endif
  static constexpr uint32_t A_SLOTS = IMMB_PAIR_A_SLOTS;
  static constexpr uint32_t BH = BN2 / 2;                     // weight rows held by each CTA
  static constexpr uint32_t B_PLANE = BH * 128;
  static constexpr uint32_t B_SLOT = B_PLANE * NPLB;
  // fused BN statistics (3-pass layers only: the frozen 2-pass tower has no BN): per epilogue warp, per N tile (<= 2),
  // sum and sum of squares of BN2 channels in double
  // Narrow N tiles are bound by the LATENCY of the epilogue (TMEM read -> bias -> statistics shuffles -> stores: one warp
  // per 32 accumulator lanes has nothing to overlap them with; ncu: the MMA warp waits on t_empty 57 % of the time at
  // BN2 = 32), so they run TWO sets of four epilogue warps, set s draining accumulator stage s (tiles alternate).
#ifndef IMMB_PAIR_EPI_SETS
#define IMMB_PAIR_EPI_SETS 2
#endif
#ifndef IMMB_PAIR_WIDE_SETS
#define IMMB_PAIR_WIDE_SETS 1
#endif
  static constexpr int EPI_SETS = (BN2 <= 64 || IMMB_PAIR_WIDE_SETS) ? IMMB_PAIR_EPI_SETS : 1;
  // how the two sets share the work: BN2 = 64 -> both sets drain EVERY tile, set s its 32-column chunk s (halves the
  // per-tile drain latency; works with two accumulator stages); BN2 = 32 (one chunk) -> tiles alternate between the sets,
  // and the accumulator ring is 4 deep so that the MMA warp can run ahead of the longer per-tile drain
  static constexpr bool EPI_SPLIT_COLS = EPI_SETS == 2 && BN2 >= 64;      // set s: 32-column chunks s, s + 2, ...
  static constexpr bool EPI_ALT_TILES = EPI_SETS == 2 && !EPI_SPLIT_COLS;
  static constexpr int THREADS = 64 + 128 * EPI_SETS;
  static constexpr uint32_t STATS_ROWS = 4 * EPI_SETS;
  // statistics columns a warp holds per N tile: all BN2 (tiles alternate) or only its own chunks (columns split)
  static constexpr uint32_t STATS_W = EPI_SPLIT_COLS ? ((BN2 / 32 + 1) / 2) * 32 : BN2;
  static constexpr uint32_t STATS_BYTES = PASSES == 3 ? STATS_ROWS * 2 * 2 * STATS_W * 8 + 4 * 256 * 4 : 0;     // + BN constants [4][256]
  static constexpr uint32_t BIAS_BYTES = 2048;            // the layer's bias vector (<= 512 channels), staged once
  static constexpr uint32_t ROOM = 232448 - 1024 - 512 - A_SLOTS * A_SLOT - STATS_BYTES - BIAS_BYTES;
  static constexpr uint32_t B_FIT = ROOM / B_SLOT;
  // ring depth 6 for streaming; layers whose whole weight operand fits (narrow layers: 9 taps x 1-2 chunks of <= 8 KB)
  // use the region as resident storage (Tc2Params::b_resident), so take what the budget gives up to 18 slots
#ifndef IMMB_PAIR_B_CAP
#define IMMB_PAIR_B_CAP 18
#endif
  static constexpr uint32_t B_SLOTS = B_FIT > IMMB_PAIR_B_CAP ? IMMB_PAIR_B_CAP : B_FIT;
  static_assert(B_SLOTS >= 2, "weight ring needs two slots");
  static constexpr uint32_t SMEM_BYTES = A_SLOTS * A_SLOT + B_SLOTS * B_SLOT + STATS_BYTES + BIAS_BYTES + 1024 + 512;
  // 3-pass: the two cross terms (hi*lo, lo*hi; ~2^-11 of the main term) accumulate in their OWN TMEM columns
  // [BN2, 2*BN2) and are added to the main accumulator once, in the epilogue.  The tensor core truncates the fp32
  // accumulator at every MMA, a biased error that grows linearly with the number of accumulating MMAs; keeping the
  // small terms out of the main accumulator leaves it with ONE truncating add per K-step, like a plain TF32 GEMM
  // (measured: all three products into one accumulator pushed the BN-beta gradients of config 1 from <1e-2 to 1.01e-2).
  // F16: the lo planes carry an extra factor 2^11, so every product with a lo operand (also the single one of the
  // 2-pass frozen-weight product) goes to the cross accumulator and is scaled by 2^-11 when the epilogue adds it.
  static constexpr bool CROSS = PASSES == 3 || (F16 && PASSES == 2);
  // Narrow N tiles of the 3-pass product are bound by the shared-memory reads of the A operand (4 KB per MMA that lasts
  // 16-32 cycles), not by the tensor pipe: there A_hi is read ONCE per k-step by an MMA of N = 2*BN2 against each CTA's
  // contiguous [w_hi half ; w_lo half] rows (columns per CTA half h: [h*2*BH, h*2*BH + BH) = main, the next BH = hi*lo),
  // and A_lo * w_hi accumulates in a third block of BN2 columns: 2 A reads and 2 MMAs per k-step instead of 3.
  static constexpr bool CONCAT = PASSES == 3 && BN2 <= 64;
  static constexpr int ACC_COLS = CONCAT ? 3 * BN2 : (CROSS ? 2 * BN2 : BN2);
  static_assert(2 * ACC_COLS <= 512, "two accumulator sets must fit the 512 TMEM columns (cross accumulator: BN2 <= 128)");
#ifndef IMMB_PAIR_ACC_STAGES
#define IMMB_PAIR_ACC_STAGES 4
#endif
  static constexpr int ACC_STAGES = (BN2 <= 64 && IMMB_PAIR_ACC_STAGES * ACC_COLS <= 512) ? IMMB_PAIR_ACC_STAGES : 2;
  static constexpr int ACC_TOTAL = ACC_STAGES * ACC_COLS;
  static constexpr int TMEM_COLS = ACC_TOTAL <= 64 ? 64 : (ACC_TOTAL <= 128 ? 128 : (ACC_TOTAL <= 256 ? 256 : 512));
  static_assert(BN2 % 32 == 0 && BN2 <= 256, "pair N tile: multiple of 32 up to 256");
};

// BNR: the BN-backward-sums epilogue (stats_mode 2) is a separate instantiation so that its registers do not weigh on
// the forward / plain-dgrad variants (a high register count squeezes the glue kernels that co-run on the side streams)
template <int BN2, int PASSES, bool BNR, bool F16>
__global__ void __launch_bounds__((Tc2PairCfg<BN2, PASSES, F16>::THREADS), 1)
conv_tc2_pair_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                     const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                     const __grid_constant__ Tc2Params p) {
  using Cfg = Tc2PairCfg<BN2, PASSES, F16>;
  constexpr int KC = F16 ? 64 : 32;                 // channels per 128-byte K chunk
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int tile0 = (int)(blockIdx.x >> 1);
  const int tstep = (int)(gridDim.x >> 1);
  const int n_iter_total = p.total_pairs;
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
  uint64_t* t_empty = t_full + Cfg::ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + Cfg::ACC_STAGES);
  static_assert((2 * Cfg::A_SLOTS + 2 * Cfg::B_SLOTS + 2 * Cfg::ACC_STAGES + 1) * 8 <= 512, "barrier block");
  double* stats_sm = reinterpret_cast<double*>(b_base + Cfg::B_SLOTS * Cfg::B_SLOT + 512);     // [epilogue warps][2 n tiles][2][BN2]
  float* bias_sm = reinterpret_cast<float*>(b_base + Cfg::B_SLOTS * Cfg::B_SLOT + 512 + Cfg::STATS_BYTES);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (uint32_t i = 0; i < Cfg::B_SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < Cfg::ACC_STAGES; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], Cfg::EPI_SPLIT_COLS ? 16 : 8);      // one arrival per draining warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA_hi);
    prefetch_tmap(&mapB_hi);
    if (PASSES >= 2) prefetch_tmap(&mapA_lo);
    if (PASSES == 3) prefetch_tmap(&mapB_lo);
  }
  if (warp == 1) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs: barriers initialised and TMEM allocated before any load / MMA / arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pair iteration -> (n tile, the two m tiles); CTA `crank` owns m tile 2*jm + crank (clamped to the last m tile when
  // the count is odd: that CTA recomputes it with stores disabled)
  auto decode = [&](int it, int& img, int& th, int& tw, int& n_off, bool& live) {
    int nt, jm;
    if (p.nn_sh) { nt = it & (p.n_tiles_n - 1); jm = it >> (p.nn_sh - 1); }
    else { nt = it % p.n_tiles_n; jm = it / p.n_tiles_n; }
    int mt = jm * 2 + (int)crank;
    live = true;
    if (mt >= p.m_tiles) { mt = p.m_tiles - 1; live = false; }
    int r;
    if (p.tw_sh) { tw = mt & (p.tiles_w - 1); r = mt >> (p.tw_sh - 1); }
    else { tw = mt % p.tiles_w; r = mt / p.tiles_w; }
    if (p.th_sh) { th = r & (p.tiles_h - 1); img = r >> (p.th_sh - 1); }
    else { th = r % p.tiles_h; img = r / p.tiles_h; }
    n_off = nt * BN2;
  };

  if (warp == 0) {
    // ===== TMA producer (both CTAs; every load signals the leader's full barrier) =====
    uint32_t ai = 0, bi = 0;
    if (p.b_resident && tile0 < n_iter_total) {
      // the layer's whole weight operand (this CTA's rows of every slice), once: slot = kc * n_taps + tap
      const int row0r = (int)crank * (int)Cfg::BH;
      if (elect_one()) {
        if (leader) mbar_expect_tx(&b_full[0], 2u * (uint32_t)(p.kchunks * p.n_taps) * Cfg::B_SLOT);
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int tap = 0; tap < p.n_taps; ++tap) {
            uint8_t* sb = b_base + (uint32_t)(kc * p.n_taps + tap) * Cfg::B_SLOT;
            tma_load_3d_pair(sb, &mapB_hi, &b_full[0], kc * KC, row0r, p.taps[tap].b_tap);
            if (PASSES == 3) tma_load_3d_pair(sb + Cfg::B_PLANE, &mapB_lo, &b_full[0], kc * KC, row0r, p.taps[tap].b_tap);
          }
      }
      __syncwarp();
    }
    for (int t = tile0; t < n_iter_total; t += tstep) {
      int img, th, tw, n_off;
      bool live;
      decode(t, img, th, tw, n_off, live);
      const int row0 = n_off + (int)crank * (int)Cfg::BH;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        const uint32_t as = ai % Cfg::A_SLOTS, aph = (ai / Cfg::A_SLOTS) & 1;
        mbar_wait(&a_empty[as], aph ^ 1);
        uint8_t* sa = a_base + as * Cfg::A_SLOT;
        if (elect_one()) {
          if (leader) mbar_expect_tx(&a_full[as], 2 * Cfg::NPLA * (uint32_t)p.a_plane_bytes);
          tma_load_5d_pair(sa, &mapA_hi, &a_full[as], kc * KC, tw * 8 + p.box_dw, 0, th * 16 + p.box_dh, img);
          if (PASSES >= 2)
            tma_load_5d_pair(sa + Cfg::A_PLANE, &mapA_lo, &a_full[as], kc * KC, tw * 8 + p.box_dw, 0, th * 16 + p.box_dh, img);
        }
        __syncwarp();
        ++ai;
        if (p.b_resident) continue;
        for (int tap = 0; tap < p.n_taps; ++tap) {
          const uint32_t bs = bi % Cfg::B_SLOTS, bph = (bi / Cfg::B_SLOTS) & 1;
          mbar_wait(&b_empty[bs], bph ^ 1);
          uint8_t* sb = b_base + bs * Cfg::B_SLOT;
          if (elect_one()) {
            if (leader) mbar_expect_tx(&b_full[bs], 2 * Cfg::B_SLOT);
            tma_load_3d_pair(sb, &mapB_hi, &b_full[bs], kc * KC, row0, p.taps[tap].b_tap);
            if (PASSES == 3) tma_load_3d_pair(sb + Cfg::B_PLANE, &mapB_lo, &b_full[bs], kc * KC, row0, p.taps[tap].b_tap);
          }
          __syncwarp();
          ++bi;
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer (leader CTA only): M = 256 across the pair, N = BN2 =====
      constexpr uint32_t idesc = idesc_kind<F16>(256, BN2, 0, 0);
      constexpr uint32_t idesc2 = idesc_kind<F16>(256, 2 * BN2 <= 256 ? 2 * BN2 : BN2, 0, 0);      // [w_hi ; w_lo] rows of both CTAs
      uint32_t ai = 0, bi = 0, ti = 0;
      const bool resident = p.b_resident != 0;
      if (resident && tile0 < n_iter_total) {
        mbar_wait(&b_full[0], 0);                     // the whole weight operand of both CTAs has landed
        tc_fence_after();
      }
      for (int t = tile0; t < n_iter_total; t += tstep) {
        const uint32_t acc = ti % Cfg::ACC_STAGES, tph = (ti / Cfg::ACC_STAGES) & 1;
        mbar_wait(&t_empty[acc], tph ^ 1);            // the epilogues of BOTH CTAs have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_COLS;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const uint32_t as = ai % Cfg::A_SLOTS, aph = (ai / Cfg::A_SLOTS) & 1;
          mbar_wait(&a_full[as], aph);
          const uint32_t a_hi = smem_u32(a_base + as * Cfg::A_SLOT);
          const int last_tap = p.n_taps - 1;
          const uint32_t sbo = (uint32_t)p.a_sbo;
          const int nk = (kc == p.kchunks - 1) ? p.k_last : 4;       // 32-byte k-steps holding real channels
          // Descriptors advance by ADDING to a 64-bit value (start-address field = addr >> 4; the k-step is 32 bytes = +2):
          // rebuilding them from their fields per MMA cost ~13 SASS instructions per MMA on the single issuing thread, and
          // the narrow layers (18-36 short MMAs per tile) were bound by exactly that instruction stream (ncu source view).
          const uint64_t da0_hi = smem_desc_sw128(a_hi, 16, sbo, 2, 0);
          const uint64_t db0_hi = smem_desc_sw128(smem_u32(b_base), 16, 1024);
          const uint32_t accf0 = kc > 0 ? 1u : 0u;
          // nk_c: compile-time count of k-steps (straight-line MMA stream: 4 = full 128-byte rows, 2 = the 32-channel fp16
          // layers) or 0 = use the runtime value
          auto issue_tap = [&](int tap, uint32_t bs, auto nk_c) {
            constexpr int NKC = decltype(nk_c)::value;
            const uint64_t a_off16 = (uint64_t)(p.a_off[tap] >> 4);
            const uint64_t da_hi0 = da0_hi + a_off16;
            const uint64_t db_hi0 = db0_hi + (uint64_t)bs * (uint64_t)(Cfg::B_SLOT >> 4);
            const uint32_t accf = (accf0 | (tap > 0 ? 1u : 0u));
#pragma unroll
            for (int k4 = 0; k4 < (NKC ? NKC : 4); ++k4) {
              if (NKC || k4 < nk) {
                const uint64_t da_hi = da_hi0 + (uint64_t)(2 * k4), da_lo = da_hi0 + (uint64_t)((Cfg::A_PLANE >> 4) + 2 * k4);
                const uint64_t db_hi = db_hi0 + (uint64_t)(2 * k4), db_lo = db_hi0 + (uint64_t)((Cfg::B_PLANE >> 4) + 2 * k4);
                (void)db_lo;
                const uint32_t first = k4 > 0 ? 1u : accf;
                if (Cfg::CONCAT) {
                  mma_kind_pair<F16>(tmem_d, da_hi, db_hi, idesc2, first);               // hi*hi and hi*lo in one pass over A_hi
                  mma_kind_pair<F16>(tmem_d + 2 * BN2, da_lo, db_hi, idesc, first);      // lo*hi: third column block
                } else if (PASSES == 3) {
                  mma_kind_pair<F16>(tmem_d, da_hi, db_hi, idesc, first);
                  mma_kind_pair<F16>(tmem_d + BN2, da_hi, db_lo, idesc, first);      // cross terms: own accumulator
                  mma_kind_pair<F16>(tmem_d + BN2, da_lo, db_hi, idesc, 1u);
                } else if (PASSES == 2) {
                  mma_kind_pair<F16>(tmem_d, da_hi, db_hi, idesc, first);
                  if (F16) mma_kind_pair<F16>(tmem_d + BN2, da_lo, db_hi, idesc, first);   // lo carries 2^11: cross accumulator
                  else mma_kind_pair<F16>(tmem_d, da_lo, db_hi, idesc, 1u);
                } else {
                  mma_kind_pair<F16>(tmem_d, da_hi, db_hi, idesc, first);
                }
              }
            }
          };
          if (resident) {
            // the whole weight operand is in place: one election per chunk, no per-tap barrier traffic
            tc_fence_after();                         // (the a_full wait above ordered this chunk's activation box)
            if (elect_one()) {
              const uint32_t bs0 = (uint32_t)(kc * p.n_taps);
              if (nk == 4) { for (int tap = 0; tap <= last_tap; ++tap) issue_tap(tap, bs0 + tap, std::integral_constant<int, 4>{}); }
              else if (nk == 2) { for (int tap = 0; tap <= last_tap; ++tap) issue_tap(tap, bs0 + tap, std::integral_constant<int, 2>{}); }
              else { for (int tap = 0; tap <= last_tap; ++tap) issue_tap(tap, bs0 + tap, std::integral_constant<int, 0>{}); }
              mma_commit_pair(&a_empty[as], (uint16_t)3);
              if (kc == p.kchunks - 1) mma_commit_pair(&t_full[acc], (uint16_t)3);
            }
            __syncwarp();
          } else {
            for (int tap = 0; tap <= last_tap; ++tap) {
              const uint32_t bs = bi % Cfg::B_SLOTS, bph = (bi / Cfg::B_SLOTS) & 1;
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              if (elect_one()) {
                if (nk == 4) issue_tap(tap, bs, std::integral_constant<int, 4>{});
                else issue_tap(tap, bs, std::integral_constant<int, 0>{});
                mma_commit_pair(&b_empty[bs], (uint16_t)3);
                if (tap == last_tap) mma_commit_pair(&a_empty[as], (uint16_t)3);
                if (tap == last_tap && kc == p.kchunks - 1) mma_commit_pair(&t_full[acc], (uint16_t)3);
              }
              __syncwarp();
              ++bi;
            }
          }
          ++ai;
        }
        ++ti;
      }
    }
  } else {
    // ===== epilogue (warps 2.. of both CTAs; each CTA drains its own 128 TMEM lanes; with two sets of four warps, set s
    // takes the tiles of accumulator stage s) =====
    const int q = warp & 3;
    const int eset = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int hl = m >> 3, wl = m & 7;
    const bool vec8 = (p.ocs % 8 == 0) && (p.n_store % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 31) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.out_lo) & 31) == 0);
    const bool do_stats = PASSES == 3 && p.stats != nullptr;
    // F16: result = (main + cross * 2^-11) * 2^-(e_a + e_b); H16 output planes are written with the output tensor's 2^e_o
    float s_main = 1.f, s_cross = 1.f, s_out = 1.f, amax = 0.f;
    if (F16) {
      const int e = __ldg(p.a_scale) + __ldg(p.b_scale);
      s_main = exp2i(-e);
      s_cross = exp2i(-e - 11);
      if (p.out_lo) s_out = exp2i(__ldg(p.o_scale));
    }
    const bool vec16h = F16 && p.out_lo && (p.ocs % 16 == 0) && (p.n_store % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out_hi) & 31) == 0) && ((reinterpret_cast<uintptr_t>(p.out_lo) & 31) == 0);
    constexpr int SW = (int)Cfg::STATS_W;
    double* my_stats = stats_sm + (size_t)(eset * 4 + q) * (2 * 2 * SW);      // this warp's private rows: no cross-warp races
    float* bn_const = reinterpret_cast<float*>(stats_sm + Cfg::STATS_ROWS * 2 * 2 * SW);     // [scale | shift | mean | invstd][256]
    const bool bias_staged = p.bias != nullptr && p.n_cols <= 512;
    if (bias_staged) {
      // every epilogue warp writes the same values (benign); it only reads them after its own __syncwarp
      for (int c = lane; c < p.n_cols; c += 32) bias_sm[c] = __ldg(p.bias + c);
      __syncwarp();
    }
    if (do_stats) {
      for (int i = lane; i < 2 * 2 * SW; i += 32) my_stats[i] = 0.0;
      if (BNR) {
        // every epilogue warp stages the full table itself (identical values: benign), so no cross-warp barrier is needed
        for (int c = lane; c < p.n_cols && c < 256; c += 32) {
          bn_const[c] = __ldg(p.bnr_scale + c);
          bn_const[256 + c] = __ldg(p.bnr_shift + c);
          bn_const[512 + c] = __ldg(p.bnr_mean + c);
          bn_const[768 + c] = __ldg(p.bnr_invstd + c);
        }
      }
      __syncwarp();
    }
    uint32_t ti = 0;
    for (int t = tile0; t < n_iter_total; t += tstep) {
      const uint32_t acc = ti % Cfg::ACC_STAGES, tph = (ti / Cfg::ACC_STAGES) & 1;
      if (Cfg::EPI_ALT_TILES && (int)(ti & 1) != eset) { ++ti; continue; }
      int img, th, tw, n_off;
      bool live;
      decode(t, img, th, tw, n_off, live);
      const int h = th * 16 + hl, w = tw * 8 + wl;
      const size_t pix = ((size_t)img * p.H + h) * p.W + w;
#ifndef IMMB_PAIR_NO_PREFETCH
      if ((BNR && do_stats) || p.relu_src) {
        // the epilogue's own global reads (the BN layer's y row / the ReLU mask row of this thread's pixel) would otherwise
        // pay a full DRAM latency per tile with nothing to overlap: pull the NEXT tile's rows into L2 now
        const int tn = t + tstep * (Cfg::EPI_ALT_TILES ? 2 : 1);
        if (tn < n_iter_total) {
          int img2, th2, tw2, n_off2;
          bool live2;
          decode(tn, img2, th2, tw2, n_off2, live2);
          const size_t pix2 = ((size_t)img2 * p.H + th2 * 16 + hl) * p.W + tw2 * 8 + wl;
          for (int c0 = Cfg::EPI_SPLIT_COLS ? 32 * eset : 0; c0 < BN2; c0 += (Cfg::EPI_SPLIT_COLS ? 64 : 32)) {
            if (n_off2 + c0 >= p.n_cols) break;
            const void* ptr = (BNR && do_stats) ? (const void*)(p.bnr_y + pix2 * p.bnr_ycs + n_off2 + c0)
                            : (F16 ? (const void*)(reinterpret_cast<const uint16_t*>(p.relu_src) + pix2 * p.relu_cs + n_off2 + c0)
                                   : (const void*)(p.relu_src + pix2 * p.relu_cs + n_off2 + c0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
          }
        }
      }
#endif
      mbar_wait(&t_full[acc], tph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = Cfg::EPI_SPLIT_COLS ? 32 * eset : 0; c0 < BN2; c0 += (Cfg::EPI_SPLIT_COLS ? 64 : 32)) {
        const int col0 = n_off + c0;
        if (col0 >= p.n_cols) break;             // warp-uniform
        float v[32];
        if (Cfg::CONCAT) {
          // columns: CTA-half h of the N tile holds [main | hi*lo] for its BH channels; lo*hi follows at 2*BN2 in channel order
          const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS;
          float va[32], vb[32], vc[32];
          tmem_ld32_nowait(tb + (uint32_t)(2 * BN2 + c0), vc);
          if (Cfg::BH == 16) {                  // BN2 = 32: 16 main + 16 cross columns per half
            tmem_ld32_nowait(tb, va);
            tmem_ld32_nowait(tb + 32u, vb);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float c_lo = va[16 + j] + vc[j], c_hi = vb[16 + j] + vc[16 + j];
              v[j] = F16 ? fmaf(c_lo, s_cross, va[j] * s_main) : va[j] + c_lo;
              v[16 + j] = F16 ? fmaf(c_hi, s_cross, vb[j] * s_main) : vb[j] + c_hi;
            }
          } else {                              // BN2 = 64: a 32-channel chunk is one CTA half
            tmem_ld32_nowait(tb + (uint32_t)(2 * c0), va);
            tmem_ld32_nowait(tb + (uint32_t)(2 * c0 + 32), vb);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float c = vb[j] + vc[j];
              v[j] = F16 ? fmaf(c, s_cross, va[j] * s_main) : va[j] + c;
            }
          }
        } else if (Cfg::CROSS) {
          float v2[32];
          tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + (uint32_t)c0, v);
          tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + (uint32_t)(BN2 + c0), v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = F16 ? fmaf(v2[j], s_cross, v[j] * s_main) : v[j] + v2[j];
        } else {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + (uint32_t)c0, v);
        }
        if (Cfg::CONCAT || Cfg::CROSS) {
        } else if (F16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= s_main;
        }
        if (!live) continue;
        if (p.bias) {
          if (bias_staged && col0 + 32 <= p.n_cols) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_sm + col0 + j);
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.n_cols) v[j] += __ldg(p.bias + col0 + j);
          }
        }
        if (do_stats) {
          // per-channel sum / sum of squares of this warp's 32 pixel rows: butterfly transpose-reduce (31 shuffles per
          // quantity; lane j ends up with channel col0 + j), accumulated in double in the warp's own smem rows
          float s1[32], s2[32];
          if (BNR) {
            const float* yp = p.bnr_y + pix * p.bnr_ycs + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 yy = __ldg(reinterpret_cast<const float4*>(yp + j));       // n_cols % 32 == 0 in this mode
              const float ya[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int c = col0 + j + u;
                const float z = fmaf(ya[u], bn_const[c], bn_const[256 + c]);
                const float dz = (p.bnr_relu && !(z > 0.f)) ? 0.f : v[j + u];
                s1[j + u] = dz;
                s2[j + u] = dz * ((ya[u] - bn_const[512 + c]) * bn_const[768 + c]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float t = (col0 + j < p.n_cols) ? v[j] : 0.f;
              s1[j] = t;
              s2[j] = t * t;
            }
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float send1 = up ? s1[i] : s1[i + off], keep1 = up ? s1[i + off] : s1[i];
              const float send2 = up ? s2[i] : s2[i + off], keep2 = up ? s2[i + off] : s2[i];
              s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
              s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
            }
          }
          const int nt = (n_off / BN2) & 1;
          // running sums as an unevaluated fp32 pair (value, compensation) in the 8 bytes of the double slot: TwoSum keeps
          // the rounding error of every add (~2^-46 relative overall), and costs 7 FADDs where the DADD it replaces showed
          // as 30 % of the epilogue's samples (math-pipe stall; ncu source view of the 128-wide statistics layers)
          float2* row = reinterpret_cast<float2*>(my_stats + (size_t)nt * (2 * SW));
          const int sc0 = Cfg::EPI_SPLIT_COLS ? (c0 >> 6) * 32 : c0;       // column split: this warp's chunks are packed
          row[sc0 + lane] = two_sum_acc(row[sc0 + lane], s1[0]);
          row[SW + sc0 + lane] = two_sum_acc(row[SW + sc0 + lane], s2[0]);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.relu_src && F16) {
          // hi plane (fp16) of the post-ReLU activation: positive <=> sign clear and magnitude non-zero
          const uint16_t* a = reinterpret_cast<const uint16_t*>(p.relu_src) + pix * p.relu_cs + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j >= p.n_store) break;
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(a + j));
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t h0 = w[u] & 0xFFFFu, h1 = w[u] >> 16;
              v[j + 2 * u] = (h0 != 0u && h0 < 0x8000u) ? v[j + 2 * u] : 0.f;
              v[j + 2 * u + 1] = (h1 != 0u && h1 < 0x8000u) ? v[j + 2 * u + 1] : 0.f;
            }
          }
        } else if (p.relu_src) {
          const float* a = p.relu_src + pix * p.relu_cs + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j >= p.n_store) break;
            const float4 t = __ldg(reinterpret_cast<const float4*>(a + j));
            v[j] = t.x > 0.f ? v[j] : 0.f;
            v[j + 1] = t.y > 0.f ? v[j + 1] : 0.f;
            v[j + 2] = t.z > 0.f ? v[j + 2] : 0.f;
            v[j + 3] = t.w > 0.f ? v[j + 3] : 0.f;
          }
        }
        if (F16 && p.out_lo) {
          // H16 output planes: 16 channels = one 32-byte sector per plane per store
          uint16_t* oh = reinterpret_cast<uint16_t*>(p.out_hi) + pix * p.ocs + col0;
          uint16_t* ol = reinterpret_cast<uint16_t*>(p.out_lo) + pix * p.ocs + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 16) {
            if (col0 + j >= p.n_store) break;
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              uint16_t h0, l0, h1, l1;
              amax = fmaxf(amax, fmaxf(fabsf(v[j + 2 * u]), fabsf(v[j + 2 * u + 1])));
              split_h16(v[j + 2 * u] * s_out, h0, l0);
              split_h16(v[j + 2 * u + 1] * s_out, h1, l1);
              hw[u] = pack2(h0, h1);
              lw[u] = pack2(l0, l1);
            }
            if (vec16h) {
              st_global_v8u(oh + j, hw[0], hw[1], hw[2], hw[3], hw[4], hw[5], hw[6], hw[7]);
              st_global_v8u(ol + j, lw[0], lw[1], lw[2], lw[3], lw[4], lw[5], lw[6], lw[7]);
            } else {
              *reinterpret_cast<uint4*>(oh + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(ol + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              if (col0 + j + 8 < p.n_store) {
                *reinterpret_cast<uint4*>(oh + j + 8) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                *reinterpret_cast<uint4*>(ol + j + 8) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
              }
            }
          }
          continue;
        }
        float* o = p.out_hi + pix * p.ocs + col0;
        if (vec8) {
          // one full 32-byte sector per store
          if (p.out_lo) {
            float* ol = p.out_lo + pix * p.ocs + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j >= p.n_store) break;
              float h[8], l[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) split_tf32(v[j + u], h[u], l[u]);
              st_global_v8(o + j, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
              st_global_v8(ol + j, l[0], l[1], l[2], l[3], l[4], l[5], l[6], l[7]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j >= p.n_store) break;
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&t_empty[acc]);
      ++ti;
    }
    if (F16 && p.out_lo) {
      __syncwarp();
      h16_track_amax(p.o_scale, amax);
    }
    if (do_stats) {
      // one partial row per (CTA, epilogue warp): [sum over n_cols | sum of squares over n_cols]; summed in a fixed
      // order by the second level (immb_bn_stats_from_partials): deterministic, no atomics
      __syncwarp();
      double* out = p.stats + ((size_t)blockIdx.x * 4 + q) * (size_t)(2 * p.n_cols);
      if (Cfg::EPI_SPLIT_COLS) {
        // the two sets own disjoint columns of the same row: each warp writes its own chunks, nothing to merge
        for (int nt = 0; nt < p.n_tiles_n && nt < 2; ++nt)
          for (int c0 = 32 * eset; c0 < BN2; c0 += 64) {
            const int col = nt * BN2 + c0 + lane, sc = (c0 >> 6) * 32 + lane;
            if (col < p.n_cols) {
              out[col] = pair_value(my_stats[(size_t)nt * (2 * SW) + sc]);
              out[p.n_cols + col] = pair_value(my_stats[(size_t)nt * (2 * SW) + SW + sc]);
            }
          }
      } else {
        if (Cfg::EPI_SETS == 2) asm volatile("bar.sync 1, 256;" ::: "memory");      // set 1's rows are final: set 0 folds them in
        if (eset == 0) {
          const double* other = my_stats + (size_t)4 * (2 * 2 * SW);                // same lane quarter, second set
          for (int nt = 0; nt < p.n_tiles_n && nt < 2; ++nt)
            for (int c = lane; c < BN2; c += 32) {
              const int col = nt * BN2 + c;
              if (col < p.n_cols) {
                double a = pair_value(my_stats[(size_t)nt * (2 * SW) + c]), b = pair_value(my_stats[(size_t)nt * (2 * SW) + SW + c]);
                if (Cfg::EPI_SETS == 2) {
                  a += pair_value(other[(size_t)nt * (2 * SW) + c]);
                  b += pair_value(other[(size_t)nt * (2 * SW) + SW + c]);
                }
                out[col] = a;
                out[p.n_cols + col] = b;
              }
            }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // nobody exits (or frees TMEM) while the peer may still signal / compute into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------
// esize: bytes per element of the planes (4 = TF32 pairs in fp32 storage, 2 = scaled fp16); the box is always 128 bytes
// of channels wide (32 fp32 / 64 fp16 elements)
int tc_make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int cs, bool parity_split,
                    int box_w, int box_h, int box_n, int swizzle_mn, int esize = 4);
int tc_make_w_map(CUtensorMap* m, const void* base, int taps, int Nn, int Kd, int bn, int esize = 4);
int tc_pick_bn(int ncols);

static inline int pow2_sh(int v) {          // log2(v) + 1 for a power of two, else 0
  if (v <= 0 || (v & (v - 1))) return 0;
  int s = 0;
  while ((1 << s) < v) ++s;
  return s + 1;
}

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

// IMMB_TC2_CLUSTER: 0 (default) = never, 1 = for wide N tiles on large problems, 2 = whenever the tile shape allows
int conv_tc2_cluster_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("IMMB_TC2_CLUSTER");
    mode = e ? atoi(e) : 0;      // measured on B200: no gain at batch 64 (the weight operand is not the limiter) -> off by default
  }
  return mode;
}

template <int BN, int PASSES>
static int launch_tc2(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                      const CUtensorMap& b_lo, const Tc2Params& p, bool cluster, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN, PASSES>;
  if (!cluster) {
    auto kern = conv_tc2_kernel<BN, PASSES, 1>;
    static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device_slot()];      // the shared-memory opt-in is a per-device attribute
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
      if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2 smem attr: %s", cudaGetErrorString(e));
      configured = true;
    }
    int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(a_hi, a_lo, b_hi, b_lo, p);
    return check_launch("conv_tc2_kernel");
  }
  auto kern = conv_tc2_kernel<BN, PASSES, 2>;
  static bool configured2_dev[kMaxDevices] = {};
  bool& configured2 = configured2_dev[current_device_slot()];
  if (!configured2) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2 smem attr: %s", cudaGetErrorString(e));
    configured2 = true;
  }
  int pairs = p.total_pairs < kNumSMs / 2 ? p.total_pairs : kNumSMs / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a_hi, a_lo, b_hi, b_lo, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2 cluster launch: %s", cudaGetErrorString(e));
  return IMMB_OK;
}

// IMMB_TC2_PAIR: 1 (default) = cta_group::2 pair kernel for every halo-conv, 0 = single-CTA kernel only,
// N > 1 = pair kernel only when the layer has at least N output columns (development aid)
int conv_tc2_pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("IMMB_TC2_PAIR");
    mode = e ? atoi(e) : 1;
  }
  return mode;
}
// N tile of the pair kernel: as few N tiles as possible (one activation halo then serves up to 256 output channels;
// 128 for the 3-pass product, whose cross-term accumulator doubles the TMEM columns), each the smallest instantiated
// width that covers its share (288 columns -> 3 x 96 for 3 passes, 2 x 160 otherwise)
static int pair_bn(int ncols, int passes, bool f16 = false) {
  const int cap = (passes == 3 || (f16 && passes == 2)) ? 128 : 256;      // a cross accumulator doubles the TMEM columns
  const int nt = ceil_div(ncols, cap);
  const int need = ceil_div(ncols, nt);
  static const int kWidths[] = {32, 64, 96, 128, 160, 192, 256};
  for (int w : kWidths)
    if (w >= need) return w;
  return cap;
}

template <int BN2, int PASSES, bool BNR = false, bool F16 = false>
static int launch_tc2_pair(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                           const CUtensorMap& b_lo, const Tc2Params& p, cudaStream_t st) {
  if constexpr ((PASSES == 3 || (F16 && PASSES == 2)) && BN2 > 128) {
    return set_error(IMMB_ERR_INVALID, "conv_tc2 pair: N tile %d > 128 with a cross accumulator", BN2);
  } else if constexpr (BNR && PASSES != 3) {
    return set_error(IMMB_ERR_INVALID, "conv_tc2 pair: the BN-backward epilogue exists for the 3-pass product only");
  } else if constexpr (F16 && PASSES == 1) {
    return set_error(IMMB_ERR_INVALID, "conv_tc2 pair: single-pass fp16 is not built");
  } else {
  if (!BNR && p.stats_mode == 2) return launch_tc2_pair<BN2, PASSES, true, F16>(a_hi, a_lo, b_hi, b_lo, p, st);
  using Cfg = Tc2PairCfg<BN2, PASSES, F16>;
  auto kern = conv_tc2_pair_kernel<BN2, PASSES, BNR, F16>;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device_slot()];      // the shared-memory opt-in is a per-device attribute
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2_pair smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  int pairs = p.total_pairs < kNumSMs / 2 ? p.total_pairs : kNumSMs / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // resident weight operand: single N tile and every slice of the layer fits the weight region (IMMB_B_RESIDENT=0: off)
  static int resident_on = -1;
  if (resident_on < 0) { const char* ev = getenv("IMMB_B_RESIDENT"); resident_on = (ev && atoi(ev) == 0) ? 0 : 1; }
  Tc2Params pp = p;
  pp.b_resident = (resident_on && p.n_tiles_n == 1 && p.n_taps * p.kchunks <= (int)Cfg::B_SLOTS) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a_hi, a_lo, b_hi, b_lo, pp);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2_pair launch: %s", cudaGetErrorString(e));
  return IMMB_OK;
  }
}

static inline bool prec_is_f16(int precision) { return precision == IMMB_PREC_F16X3 || precision == IMMB_PREC_F16X2; }
static inline int prec_passes(int precision) {
  return precision == IMMB_PREC_TF32 ? 1 : ((precision == IMMB_PREC_TF32X2 || precision == IMMB_PREC_F16X2) ? 2 : 3);
}

// act: the tensor the halo boxes are read from ([N,H,W,act_cs], act_c valid channels); wts: [9][ncols_pad][kd]
// F16 precisions: planes are scaled fp16 (common.cuh); `sc` = scale records of act / weights / output planes
int conv_tc2_run(const immb_conv_desc* d, int op, const void* act_hi, const void* act_lo, int act_c, int act_cs,
                 const void* w_hi, const void* w_lo, int w_rows, int kd, const float* bias, int relu,
                 void* out_hi, void* out_lo, int ocs, int ncols, int n_store, cudaStream_t st,
                 const void* relu_src, int relu_cs, double* stats, const Tc2BnReduce* bnr, const Tc2Scales* sc) {
  const int passes = prec_passes(d->precision);
  const bool f16 = prec_is_f16(d->precision);
  const int esize = f16 ? 2 : 4, kchunk = f16 ? 64 : 32;
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  p.tiles_w = d->W / 8; p.tiles_h = d->H / 16; p.n_img = d->N;
  const int pmode = conv_tc2_pair_mode();
  const bool pair = pmode > 0 && ncols >= (pmode > 1 ? pmode : 1);
  if (f16 && !pair) return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc2: the fp16 product needs the pair kernel");
  if (f16 && (!sc || !sc->a || !sc->b || (out_lo && !sc->o)))
    return set_error(IMMB_ERR_INVALID, "conv_tc2: fp16 planes need their scale records");
  if (f16 && (act_cs % 8 || (out_lo && (ocs % 8 || n_store % 8))))
    return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc2: fp16 planes need channel strides that are multiples of 8");
  const int bn = pair ? pair_bn(ncols, passes, f16) : tc_pick_bn(ncols);
  p.n_tiles_n = ceil_div(ncols, bn);
  p.total_tiles = p.tiles_w * p.tiles_h * p.n_img * p.n_tiles_n;
  p.m_tiles = p.tiles_w * p.tiles_h * p.n_img;
  p.total_pairs = ceil_div(p.m_tiles, 2) * p.n_tiles_n;
  p.tw_sh = pow2_sh(p.tiles_w); p.th_sh = pow2_sh(p.tiles_h); p.nn_sh = pow2_sh(p.n_tiles_n);
  // cluster mode pays off when the weight slice dominates the operand traffic (wide N tiles) and there is enough work
  const int cmode = conv_tc2_cluster_mode();
  const bool cluster = !pair && cmode > 0 && bn >= 32 && (bn % 32 == 0 || bn == 96) && (cmode == 2 ? bn >= 32 : bn >= 64) &&
                       (cmode == 2 || p.m_tiles >= 2 * kNumSMs);
  p.kchunks = ceil_div(kd, kchunk);
  // k-steps (32 bytes: 8 tf32 / 16 fp16 channels) of the last chunk that hold real (or zero-padded) weights
  p.k_last = f16 ? ceil_div(kd - (p.kchunks - 1) * kchunk, 16) : 4;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      Tc2Tap& t = p.taps[r * 3 + s];
      t.b_tap = r * 3 + s;
      t.ro = op == 0 ? r : 2 - r;
      t.so = op == 0 ? s : 2 - s;
    }
  // halo box width: the pair kernel fetches exactly the 8 + 2 columns a tile needs (the descriptors' group stride is
  // then 10 * 128 B; groups may start anywhere on a 128-byte boundary because both TMA and the MMA unit swizzle on
  // absolute shared-memory address bits); the single-CTA kernel keeps its 16-pixel box (SBO = 2048 B).
  const int box_w = pair ? 10 : 16;        // (a 16-pixel box for the pair kernel measured 0.6 % slower and needs 60 % more smem)
  p.n_taps = 9; p.a_sbo = box_w * 128; p.a_plane_bytes = 18 * box_w * 128; p.box_dw = -1; p.box_dh = -1;
  for (int i = 0; i < 9; ++i) p.a_off[i] = (uint32_t)(p.taps[i].ro * box_w + p.taps[i].so) * 128u;
  p.out_hi = reinterpret_cast<float*>(out_hi); p.out_lo = reinterpret_cast<float*>(out_lo); p.bias = bias; p.relu = relu;
  p.H = d->H; p.W = d->W; p.ocs = ocs; p.n_cols = ncols; p.n_store = n_store;
  p.relu_src = reinterpret_cast<const float*>(relu_src); p.relu_cs = relu_cs; p.stats = stats; p.stats_mode = stats ? 1 : 0;
  if (sc) { p.a_scale = sc->a; p.b_scale = sc->b; p.o_scale = sc->o; }
  if (bnr) {
    if (!stats || ncols % 32 || ncols > 256)
      return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc2: fused BN-backward sums need a partials buffer and 32 | channels <= 256");
    p.stats_mode = 2;
    p.bnr_y = bnr->y; p.bnr_ycs = bnr->ycs; p.bnr_relu = bnr->relu;
    p.bnr_scale = bnr->scale; p.bnr_shift = bnr->shift; p.bnr_mean = bnr->mean; p.bnr_invstd = bnr->invstd;
  }
  if (stats && (!pair || passes != 3 || p.n_tiles_n > 2))
    return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc2: fused BN statistics need the 3-pass pair kernel and <= 2 N tiles");
  if (relu_src && !pair) return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc2: the fused ReLU-backward epilogue needs the pair kernel");
  { const char* e = getenv("IMMB_TC2_BO"); p.bo_mode = e ? atoi(e) : 0; }
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = tc_make_act_map(&a_hi, act_hi, d->N, d->H, d->W, act_c, act_cs, false, box_w, 18, 1, 0, esize))) return rc;
  const int b_box = (cluster || pair) ? bn / 2 : bn;  // cluster / pair mode: each CTA loads half of the rows
  if ((rc = tc_make_w_map(&b_hi, w_hi, 9, w_rows, kd, b_box, esize))) return rc;
  a_lo = a_hi; b_lo = b_hi;
  if (passes >= 2) {
    if ((rc = tc_make_act_map(&a_lo, act_lo, d->N, d->H, d->W, act_c, act_cs, false, box_w, 18, 1, 0, esize))) return rc;
  }
  if (passes == 3) {
    if ((rc = tc_make_w_map(&b_lo, w_lo, 9, w_rows, kd, b_box, esize))) return rc;
  }
  if (pair && f16) {
#define IMMB_PCASE(BN_)                                                                        \
  if (bn == BN_)                                                                               \
    return passes == 3 ? launch_tc2_pair<BN_, 3, false, true>(a_hi, a_lo, b_hi, b_lo, p, st)   \
                       : launch_tc2_pair<BN_, 2, false, true>(a_hi, a_lo, b_hi, b_lo, p, st);
    IMMB_PCASE(32)
    IMMB_PCASE(64)
    IMMB_PCASE(96)
    IMMB_PCASE(128)
#undef IMMB_PCASE
    return set_error(IMMB_ERR_INVALID, "conv_tc2 pair (fp16): unsupported BN %d", bn);
  }
  if (pair) {
#define IMMB_PCASE(BN_)                                                                  \
  if (bn == BN_)                                                                         \
    return passes == 3 ? launch_tc2_pair<BN_, 3>(a_hi, a_lo, b_hi, b_lo, p, st)          \
         : passes == 2 ? launch_tc2_pair<BN_, 2>(a_hi, a_lo, b_hi, b_lo, p, st)          \
                       : launch_tc2_pair<BN_, 1>(a_hi, a_lo, b_hi, b_lo, p, st);
    IMMB_PCASE(32)
    IMMB_PCASE(64)
    IMMB_PCASE(96)
    IMMB_PCASE(128)
    IMMB_PCASE(160)
    IMMB_PCASE(192)
    IMMB_PCASE(256)
#undef IMMB_PCASE
    return set_error(IMMB_ERR_INVALID, "conv_tc2 pair: unsupported BN %d", bn);
  }
#define IMMB_CASE(BN_)                                                                     \
  if (bn == BN_)                                                                           \
    return passes == 3 ? launch_tc2<BN_, 3>(a_hi, a_lo, b_hi, b_lo, p, cluster, st)        \
         : passes == 2 ? launch_tc2<BN_, 2>(a_hi, a_lo, b_hi, b_lo, p, cluster, st)        \
                       : launch_tc2<BN_, 1>(a_hi, a_lo, b_hi, b_lo, p, cluster, st);
  IMMB_CASE(16)
  IMMB_CASE(32)
  IMMB_CASE(64)
  IMMB_CASE(96)
  IMMB_CASE(128)
#undef IMMB_CASE
  return set_error(IMMB_ERR_INVALID, "conv_tc2: unsupported BN %d", bn);
}


// ---- 7x7 / Cin = 3 first encoder layer on the CTA-pair kernel --------------------------------------------------
// The staged image [N,H,W+8,4] is read through the overlapping row-window view (conv_tc.cu make_rowwin_map): one
// 128-byte smem row = the 8 pixels x 4 channels starting at a pixel, so filter row r is ONE K = 32 chunk and needs no
// column halo.  Tile = 16 rows x 8 columns; ONE box of (16 + 6) window rows x 8 pixels per plane (22.5 KB) serves all
// seven filter rows through descriptors that start r * 1024 bytes into it (conv_tc_kernel fetched one 16 KB tile per
// filter row: 5x the L2 -> smem traffic, which was this layer's limiter).
int tc_make_rowwin_map(CUtensorMap* m, const float* base, int N, int H, int W, int box_w, int box_h, int box_n);

bool conv_tc2_rowwin_eligible(const immb_conv_desc* d) {
  if (!conv_tc2_enabled() || conv_tc2_pair_mode() <= 0) return false;
  if (d->x_layout != IMMB_XLAYOUT_ROWWIN4 || d->kh != 7 || d->kw != 7 || d->Cin != 3 || d->stride != 1) return false;
  if (d->pad_t != 3 || d->pad_l != 3) return false;
  return d->H % 16 == 0 && d->W % 8 == 0 && d->Cout % 32 == 0 && d->Cout <= 256 && d->y_cstride == d->Cout;
}

// rows of BN-statistics partials the pair kernel writes for this forward conv (4 per CTA), 0 = not served by it
int conv_tc2_fwd_stats_rows(const immb_conv_desc* d) {
  const bool rowwin = conv_tc2_rowwin_eligible(d);
  if (!rowwin && !(conv_tc2_eligible(d, 0) && conv_tc2_pair_mode() == 1)) return 0;
  if (d->precision != IMMB_PREC_TF32X3 && d->precision != IMMB_PREC_F16X3) return 0;
  if (rowwin && d->precision != IMMB_PREC_TF32X3) return 0;
  const int m_tiles = (d->W / 8) * (d->H / 16) * d->N;
  const int bn = pair_bn(d->Cout, 3);
  const int n_tiles_n = ceil_div(d->Cout, bn);
  if (n_tiles_n > 2) return 0;
  const int total_pairs = ceil_div(m_tiles, 2) * n_tiles_n;
  const int pairs = total_pairs < kNumSMs / 2 ? total_pairs : kNumSMs / 2;
  return 2 * pairs * 4;
}

// rows of BN-backward partials the pair kernel writes for this dgrad (4 per CTA), 0 = not served
int conv_tc2_dgrad_stats_rows(const immb_conv_desc* d) {
  if (!(conv_tc2_eligible(d, 1) && conv_tc2_pair_mode() == 1) ||
      (d->precision != IMMB_PREC_TF32X3 && d->precision != IMMB_PREC_F16X3)) return 0;
  const int ncols = d->cin_pad < d->x_cstride ? d->cin_pad : d->x_cstride;
  if (ncols != d->Cin || ncols % 32 || ncols > 256) return 0;
  const int m_tiles = (d->W / 8) * (d->H / 16) * d->N;
  const int bn = pair_bn(ncols, 3);
  const int n_tiles_n = ceil_div(ncols, bn);
  if (n_tiles_n > 2) return 0;
  const int total_pairs = ceil_div(m_tiles, 2) * n_tiles_n;
  const int pairs = total_pairs < kNumSMs / 2 ? total_pairs : kNumSMs / 2;
  return 2 * pairs * 4;
}

int conv_tc2_rowwin_fwd(const immb_conv_desc* d, const float* x_hi, const float* x_lo, const float* wp_hi,
                        const float* wp_lo, const float* bias, int relu, float* y_hi, float* y_lo, cudaStream_t st,
                        double* stats) {
  const int passes = d->precision == IMMB_PREC_TF32 ? 1 : (d->precision == IMMB_PREC_TF32X2 ? 2 : 3);
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  p.tiles_w = d->W / 8; p.tiles_h = d->H / 16; p.n_img = d->N;
  const int bn = pair_bn(d->Cout, passes);
  p.n_tiles_n = ceil_div(d->Cout, bn);
  p.m_tiles = p.tiles_w * p.tiles_h * p.n_img;
  p.total_tiles = p.m_tiles * p.n_tiles_n;
  p.total_pairs = ceil_div(p.m_tiles, 2) * p.n_tiles_n;
  p.tw_sh = pow2_sh(p.tiles_w); p.th_sh = pow2_sh(p.tiles_h); p.nn_sh = pow2_sh(p.n_tiles_n);
  p.kchunks = 1; p.k_last = 4;
  p.n_taps = 7; p.a_sbo = 1024; p.a_plane_bytes = 22 * 8 * 128; p.box_dw = 0; p.box_dh = -3;
  for (int r = 0; r < 7; ++r) { p.taps[r].b_tap = r; p.a_off[r] = (uint32_t)r * 1024u; }
  p.out_hi = y_hi; p.out_lo = y_lo; p.bias = bias; p.relu = relu;
  p.H = d->H; p.W = d->W; p.ocs = d->y_cstride; p.n_cols = d->Cout; p.n_store = d->Cout;
  p.stats = stats; p.stats_mode = stats ? 1 : 0;
  if (stats && (passes != 3 || p.n_tiles_n > 2))
    return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc2 rowwin: fused BN statistics need the 3-pass kernel and <= 2 N tiles");
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = tc_make_rowwin_map(&a_hi, x_hi, d->N, d->H, d->W, 8, 22, 1))) return rc;
  if ((rc = tc_make_w_map(&b_hi, wp_hi, 7, d->Cout, 32, bn / 2))) return rc;
  a_lo = a_hi; b_lo = b_hi;
  if (passes >= 2 && (rc = tc_make_rowwin_map(&a_lo, x_lo, d->N, d->H, d->W, 8, 22, 1))) return rc;
  if (passes == 3 && (rc = tc_make_w_map(&b_lo, wp_lo, 7, d->Cout, 32, bn / 2))) return rc;
#define IMMB_PCASE(BN_)                                                                  \
  if (bn == BN_)                                                                         \
    return passes == 3 ? launch_tc2_pair<BN_, 3>(a_hi, a_lo, b_hi, b_lo, p, st)          \
         : passes == 2 ? launch_tc2_pair<BN_, 2>(a_hi, a_lo, b_hi, b_lo, p, st)          \
                       : launch_tc2_pair<BN_, 1>(a_hi, a_lo, b_hi, b_lo, p, st);
  IMMB_PCASE(32)
  IMMB_PCASE(64)
  IMMB_PCASE(96)
  IMMB_PCASE(128)
  IMMB_PCASE(160)
  IMMB_PCASE(192)
  IMMB_PCASE(256)
#undef IMMB_PCASE
  return set_error(IMMB_ERR_INVALID, "conv_tc2 rowwin: unsupported BN %d", bn);
}

// =============================================================================================================
// Halo-reuse wgrad for the stride-1 3x3 layers.
//   dW[r][s][ci][co] = sum_pixels X[h+r-1, w+s-1][ci] * dY[h, w][co]
// One CTA owns a 32-input-channel chunk x BN output channels x a range of pixel tiles (split-K).  Per stage it
// fetches ONE (4+2) x 16-pixel halo box of X (12 KB / plane) and the matching 4 x 8-pixel dY boxes, and issues, for
// each filter row r and each image row of the tile, ONE MMA with M = 128 = 4 column taps (s = 0..3, the 4th is
// ignored) x 32 channels: the four "MN groups" of the MN-major A descriptor are the SAME halo rows shifted by one
// pixel each (LBO = 128 B), K = 8 pixels of one image row (two 4-row BASE32B atoms, SBO = 512 B).
// Three TMEM accumulators (one per r).  Operand traffic per 32 pixels: 24 KB (X) + BN/32 * 8 KB (dY), versus
// 9 x 8 KB + taps-tiles x dY in conv_tc_wgrad_kernel.
// =============================================================================================================
struct Wg2Params {
  int tiles_w, tiles_h, n_img, total_tiles, tiles_per_split;
  float* dw;
  int Cin, Cout;
};

template <int BN, int PASSES, int STAGES>
struct Wg2Cfg {
  static constexpr uint32_t NPL = PASSES == 3 ? 2 : 1;
  static constexpr uint32_t A_PLANE = 6 * 16 * 128;           // 12288 B halo box
  static constexpr uint32_t B_PLANE = BN * 128;               // BN/32 boxes of 32 px x 128 B
  static constexpr uint32_t STAGE_BYTES = (A_PLANE + B_PLANE) * NPL;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  // 3-pass, BN <= 64: X_hi x [dY_hi | dY_lo] as ONE MMA of N = 2*BN (the lo plane follows the hi plane in the stage)
  static constexpr bool CONCAT = PASSES == 3 && BN <= 64;
  static constexpr int ACC = CONCAT ? 2 * BN : BN;            // columns per filter-row accumulator
  static constexpr int TMEM_COLS = 3 * ACC <= 128 ? 128 : (3 * ACC <= 256 ? 256 : 512);
};

template <int BN, int PASSES, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_tc2_wgrad_kernel(const __grid_constant__ CUtensorMap mapX_hi, const __grid_constant__ CUtensorMap mapX_lo,
                      const __grid_constant__ CUtensorMap mapY_hi, const __grid_constant__ CUtensorMap mapY_lo,
                      const __grid_constant__ Wg2Params p) {
  using Cfg = Wg2Cfg<BN, PASSES, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int ci0 = blockIdx.x * 32;
  const int n_off = blockIdx.y * BN;
  const int t_begin = blockIdx.z * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.total_tiles) t_end = p.total_tiles;
  const int num_k = t_end - t_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
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
        tma_load_5d(st, &mapX_hi, &full[s], ci0, twi * 8 - 1, 0, thi * 4 - 1, n);
        if (PASSES == 3) tma_load_5d(st + Cfg::A_PLANE, &mapX_lo, &full[s], ci0, twi * 8 - 1, 0, thi * 4 - 1, n);
        uint8_t* sb = st + Cfg::A_PLANE * Cfg::NPL;
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
          tma_load_5d(sb + j * 4096, &mapY_hi, &full[s], n_off + j * 32, twi * 8, 0, thi * 4, n);
          if (PASSES == 3) tma_load_5d(sb + Cfg::B_PLANE + j * 4096, &mapY_lo, &full[s], n_off + j * 32, twi * 8, 0, thi * 4, n);
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (num_k > 0) {
      constexpr uint32_t idesc = idesc_tf32(128, BN, 1, 1);
      constexpr uint32_t idesc2 = idesc_tf32(128, 2 * BN <= 256 ? 2 * BN : BN, 1, 1);
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t a_lo = a_hi + Cfg::A_PLANE;
        const uint32_t b_hi = a_hi + Cfg::A_PLANE * Cfg::NPL;
        const uint32_t b_lo = b_hi + Cfg::B_PLANE;
        if (elect_one()) {
        // one descriptor per operand and stage; the MMAs add constants to its start-address field
        const uint64_t da0_hi = smem_desc_sw128(a_hi, 128, 512, 1), db0_hi = smem_desc_sw128(b_hi, 4096, 512, 1);
        const uint64_t da0_lo = da0_hi + (uint64_t)(Cfg::A_PLANE >> 4), db0_lo = db0_hi + (uint64_t)(Cfg::B_PLANE >> 4);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const uint32_t tmem_d = tmem_base + (uint32_t)(r * Cfg::ACC);
#pragma unroll
          for (int hl = 0; hl < 4; ++hl) {
            const uint64_t ao = (uint64_t)(((hl + r) * 16 * 128) >> 4);   // halo row hl+r, column tap s via LBO
            const uint64_t bo_ = (uint64_t)((hl * 1024) >> 4);            // dY rows hl*8 .. hl*8+7
            const uint64_t da_hi = da0_hi + ao;
            const uint64_t db_hi = db0_hi + bo_;
            mma_tf32(tmem_d, da_hi, db_hi, Cfg::CONCAT ? idesc2 : idesc, (kt > 0 || hl > 0) ? 1u : 0u);
            if (PASSES == 3) {
              const uint64_t da_lo = da0_lo + ao;
              if (!Cfg::CONCAT) {
                const uint64_t db_lo = db0_lo + bo_;
                mma_tf32(tmem_d, da_hi, db_lo, idesc, 1u);
              }
              mma_tf32(tmem_d, da_lo, db_hi, idesc, 1u);
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
    const int s = m >> 5;                 // column tap (0..3; 3 is the overhang group)
    const int ci = ci0 + (m & 31);
    const bool row_ok = (s < 3) && (ci < p.Cin);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * Cfg::ACC + c0), v);
        if (Cfg::CONCAT) {
          float v2[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * Cfg::ACC + BN + c0), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        }
        if (!row_ok) continue;
        const int col0 = n_off + c0;
        if (col0 >= p.Cout) continue;
        float* o = p.dw + ((size_t)(r * 3 + s) * p.Cin + ci) * p.Cout + col0;
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// =============================================================================================================
// Halo wgrad on scaled fp16 planes (kind::f16, both operands MN-major, standard 128-byte swizzle).
//   dW[r][s][ci][co] = sum_pixels X[h+r-1, w+s-1][ci] * dY[h, w][co]
// One CTA owns ONE filter row r, a 64-input-channel chunk (one 128-byte smem row = 64 fp16 channels), BN output
// channels and a range of 4x8-pixel tiles (split-K).  Per stage it fetches the 4 x 16-pixel window of X rows
// h + r - 1 (8 KB / plane) and the 4 x 8-pixel dY boxes (4 KB per 64 channels / plane).  One MMA has M = 128 = two
// column taps x 64 channels -- the two MN atoms of the A descriptor are the SAME window rows one pixel apart
// (LBO = 128 B) -- and K = 16 = the 8 pixels of two consecutive image rows (two K atoms, SBO = the 2048-byte row
// pitch of the window / the 1024-byte row pitch of a dY box).  Taps s = 0,1 share one MMA, tap s = 2 takes a second
// one whose upper MN atom is ignored.  The lo planes carry 2^11 (common.cuh), so hi*hi accumulates in "main" columns
// and hi*lo + lo*hi in "cross" columns: four accumulators of BN columns ([s01 | s2] x [main | cross]).
// Grid: (64-channel chunks, BN tiles, 3 * splits).  Epilogue: (main + cross * 2^-11) * 2^-(e_x + e_dy), vector red.
// =============================================================================================================
struct Wg16Params {
  int tiles_w, tiles_h, n_img, total_tiles, tiles_per_split, splits;
  float* dw;
  int Cin, Cout;
  const int32_t* x_scale;
  const int32_t* dy_scale;
};

// RPC = filter rows per CTA: 1 (the CTA's row comes from blockIdx.z) or 3 (narrow N tiles: the twelve accumulators of all
// three rows fit TMEM, X and dY are streamed once instead of three times; the window then carries the 2 halo rows)
// C32 (layers with <= 32 input channels): the X window is fetched as 64-byte rows (32 channels, SWIZZLE_64B), so that an
// MMA's M = 128 is FOUR column-tap atoms of 32 channels one pixel apart (s = 0..3, the 4th ignored) instead of two
// half-empty 64-channel atoms: one MMA group per filter row instead of two.
template <int BN, int STAGES, int RPC = 1, bool C32 = false>
struct Wg16Cfg {
  static constexpr uint32_t A_ROWS = 4 + RPC - 1;
  static constexpr uint32_t A_ROW_BYTES = C32 ? 64 : 128;
  static constexpr uint32_t A_PITCH = 16 * A_ROW_BYTES;                    // bytes between image rows of the window
  static constexpr uint32_t A_PLANE = (A_ROWS * A_PITCH + 1023) / 1024 * 1024;
  static constexpr int GROUPS = C32 ? 1 : 2;                              // MMA groups per filter row
  static constexpr uint32_t NB = (BN + 63) / 64;                          // dY boxes of 64 channels
  // N concatenation (IMMB_WG_NCAT): the dY lo plane directly follows the hi plane in the stage, so ONE MMA of N = 2*BN
  // against the hi descriptor yields [x_hi*dy_hi | x_hi*dy_lo] = the [main | cross] accumulator pair -- x_hi is read
  // from shared memory once instead of twice (these kernels are bound by exactly those reads at N <= 64).  BN = 32 needs
  // 32-channel N atoms for that: dY boxes of 32 channels with the 64-byte swizzle (B32), like the C32 x operand.
#ifndef IMMB_WG_NCAT
#define IMMB_WG_NCAT 1
#endif
  static constexpr bool NCAT = IMMB_WG_NCAT != 0;
  static constexpr bool B32 = NCAT && BN == 32;
  static constexpr uint32_t B_PLANE = B32 ? 2048 : NB * 4096;
  static constexpr uint32_t STAGE_BYTES = (A_PLANE + B_PLANE) * 2;        // hi + lo planes of both operands
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int ACC_COLS = RPC * GROUPS * 2 * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512));
  static_assert(ACC_COLS <= 512, "[main | cross] accumulators of BN columns per MMA group and filter row");
};

template <int BN, int STAGES, int RPC, bool C32>
__global__ void __launch_bounds__(192, 1)
conv_tc2_wgrad16_kernel(const __grid_constant__ CUtensorMap mapX_hi, const __grid_constant__ CUtensorMap mapX_lo,
                        const __grid_constant__ CUtensorMap mapY_hi, const __grid_constant__ CUtensorMap mapY_lo,
                        const __grid_constant__ Wg16Params p) {
  using Cfg = Wg16Cfg<BN, STAGES, RPC, C32>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int ci0 = blockIdx.x * (C32 ? 32 : 64);
  const int n_off = blockIdx.y * BN;
  const int r = RPC == 3 ? 0 : blockIdx.z / p.splits;       // first filter row of this CTA
  const int split = RPC == 3 ? blockIdx.z : blockIdx.z - r * p.splits;
  const int t_begin = split * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.total_tiles) t_end = p.total_tiles;
  const int num_k = t_end - t_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
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
        mbar_expect_tx(&full[s], 2 * (Cfg::A_ROWS * Cfg::A_PITCH + Cfg::B_PLANE));
        tma_load_5d(st, &mapX_hi, &full[s], ci0, twi * 8 - 1, 0, thi * 4 + r - 1, n);
        tma_load_5d(st + Cfg::A_PLANE, &mapX_lo, &full[s], ci0, twi * 8 - 1, 0, thi * 4 + r - 1, n);
        uint8_t* sb = st + Cfg::A_PLANE * 2;
#pragma unroll
        for (uint32_t j = 0; j < Cfg::NB; ++j) {
          tma_load_5d(sb + j * 4096, &mapY_hi, &full[s], n_off + (int)j * 64, twi * 8, 0, thi * 4, n);
          tma_load_5d(sb + Cfg::B_PLANE + j * 4096, &mapY_lo, &full[s], n_off + (int)j * 64, twi * 8, 0, thi * 4, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (num_k > 0) {
      constexpr uint32_t idesc = idesc_f16(128, BN, 1, 1);
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t a_lo = a_hi + Cfg::A_PLANE;
        const uint32_t b_hi = a_hi + Cfg::A_PLANE * 2;
        const uint32_t b_lo = b_hi + Cfg::B_PLANE;
        if (elect_one()) {
          // one descriptor per operand and stage, advanced by ADDING to the start-address field (rebuilding the fields
          // per MMA cost ~13 instructions per MMA on the single issuing thread: more than a 16-cycle N = 32 MMA lasts).
          // A: LBO = one pixel (the next column tap), SBO = one image row (the next 8 K rows); C32: 64-byte swizzle
          const uint64_t da0_hi = smem_desc_sw128(a_hi, Cfg::A_ROW_BYTES, Cfg::A_PITCH, C32 ? 4 : 2);
          const uint64_t da0_lo = da0_hi + (uint64_t)(Cfg::A_PLANE >> 4);
          // B: N atoms (64 channels x 8 pixels, or 32 x 8 with the 64-byte swizzle) are LBO apart, 8-pixel K groups SBO apart
          const uint64_t db0_hi = Cfg::B32 ? smem_desc_sw128(b_hi, Cfg::B_PLANE, 512, 4) : smem_desc_sw128(b_hi, 4096, 1024, 2);
          const uint64_t db0_lo = db0_hi + (uint64_t)(Cfg::B_PLANE >> 4);
          constexpr uint32_t idesc2 = idesc_f16(128, 2 * BN, 1, 1);
          constexpr uint32_t kBStep = Cfg::B32 ? 1024 : 2048;              // two image rows of the dY tile
#pragma unroll
          for (int j = 0; j < 2; ++j) {                                   // image rows 2j, 2j+1 of the tile = K 16
            const uint32_t acc = (kt > 0 || j > 0) ? 1u : 0u;
            const uint64_t db_hi = db0_hi + (uint64_t)(j * (kBStep >> 4));
            const uint64_t db_lo = db0_lo + (uint64_t)(j * (kBStep >> 4));
            (void)db_lo;
#pragma unroll
            for (int rr = 0; rr < RPC; ++rr) {
#pragma unroll
              for (int g = 0; g < Cfg::GROUPS; ++g) {                     // g = 0: taps s = 0,1 (C32: s = 0..3);  g = 1: tap s = 2 (+ ignored)
                const uint64_t ao = (uint64_t)(((uint32_t)(2 * j + rr) * Cfg::A_PITCH + (uint32_t)(2 * g) * Cfg::A_ROW_BYTES) >> 4);
                const uint64_t da_hi = da0_hi + ao;
                const uint64_t da_lo = da0_lo + ao;
                const uint32_t d_main = tmem_base + (uint32_t)((rr * Cfg::GROUPS + g) * 2 * BN);
                const uint32_t d_cross = d_main + (uint32_t)BN;
                if (Cfg::NCAT) {
                  mma_f16(d_main, da_hi, db_hi, idesc2, acc);              // [main | x_hi*dy_lo]
                } else {
                  mma_f16(d_main, da_hi, db_hi, idesc, acc);
                  mma_f16(d_cross, da_hi, db_lo, idesc, acc);
                }
                mma_f16(d_cross, da_lo, db_hi, idesc, 1u);
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
    const int g_row = C32 ? (m >> 5) : (m >> 6);      // MN atom of this TMEM lane: column tap s = 2*g + g_row (C32: s = g_row)
    const int ci = ci0 + (C32 ? (m & 31) : (m & 63));
    const int e = __ldg(p.x_scale) + __ldg(p.dy_scale);
    const float s_main = exp2i(-e), s_cross = exp2i(-e - 11);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int rg = 0; rg < Cfg::GROUPS * RPC; ++rg) {
      const int rr = rg / Cfg::GROUPS, g = rg % Cfg::GROUPS;
      const int s_tap = 2 * g + g_row;
      const bool row_ok = (s_tap < 3) && (ci < p.Cin);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32], v2[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(rg * 2 * BN + c0), v);
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(rg * 2 * BN + BN + c0), v2);
        tmem_ld_wait();
        if (!row_ok) continue;
        const int col0 = n_off + c0;
        if (col0 >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(v2[j], s_cross, v[j] * s_main);
        float* o = p.dw + ((size_t)((r + rr) * 3 + s_tap) * p.Cin + ci) * p.Cout + col0;
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int RPC, bool C32 = false>
static int launch_wg16(const CUtensorMap& x_hi, const CUtensorMap& x_lo, const CUtensorMap& y_hi,
                       const CUtensorMap& y_lo, const Wg16Params& p, dim3 grid, cudaStream_t st) {
  constexpr int STAGES = BN <= 64 ? 6 : 5;
  using Cfg = Wg16Cfg<BN, STAGES, RPC, C32>;
  auto kern = conv_tc2_wgrad16_kernel<BN, STAGES, RPC, C32>;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device_slot()];      // the shared-memory opt-in is a per-device attribute
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2_wgrad16 smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(x_hi, x_lo, y_hi, y_lo, p);
  return check_launch("conv_tc2_wgrad16_kernel");
}

static int conv_tc2_wgrad16_run(const immb_conv_desc* d, const void* x_hi, const void* x_lo, const void* dy_hi,
                                const void* dy_lo, float* dw, cudaStream_t st) {
  if (!d->x_scale || !d->y_scale) return set_error(IMMB_ERR_INVALID, "conv_tc2_wgrad: fp16 planes need x_scale / y_scale");
  if (d->x_cstride % 8 || d->y_cstride % 8)
    return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc2_wgrad: fp16 planes need channel strides that are multiples of 8");
  Wg16Params p;
  memset(&p, 0, sizeof(p));
  p.tiles_w = d->W / 8; p.tiles_h = d->H / 4; p.n_img = d->N;
  p.total_tiles = p.tiles_w * p.tiles_h * d->N;
  p.dw = dw; p.Cin = d->Cin; p.Cout = d->Cout;
  p.x_scale = d->x_scale; p.dy_scale = d->y_scale;
  const int bn = d->Cout > 64 ? 128 : (d->Cout > 32 ? 64 : 32);      // (narrower tiles would read past their TMEM columns)
  const int n_tiles = ceil_div(d->Cout, bn);
  static int c32_on = -1;
  if (c32_on < 0) { const char* ev = getenv("IMMB_WG_C32"); c32_on = (ev && atoi(ev) == 0) ? 0 : 1; }
  const bool c32 = c32_on && bn == 32 && d->Cin <= 32;                 // 32-channel window atoms (SWIZZLE_64B)
  const int c_tiles = ceil_div(d->Cin, c32 ? 32 : 64);
  const int rpc = bn == 32 ? 3 : 1;                                  // all three filter rows in one CTA when TMEM allows
  int splits = kNumSMs / (c_tiles * n_tiles * (rpc == 3 ? 1 : 3));
  if (splits > p.total_tiles) splits = p.total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = ceil_div(p.total_tiles, splits);
  splits = ceil_div(p.total_tiles, p.tiles_per_split);
  p.splits = splits;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)9 * d->Cin * d->Cout, st);
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "wgrad memset: %s", cudaGetErrorString(e));
  CUtensorMap mx_hi, mx_lo, my_hi, my_lo;
  int rc;
  const int box_h = 4 + rpc - 1;
  const int xsw = c32 ? 2 : 1;        // tc_make_act_map: 2 = 32-channel rows with the 64-byte swizzle
  if ((rc = tc_make_act_map(&mx_hi, x_hi, d->N, d->H, d->W, d->Cin, d->x_cstride, false, 16, box_h, 1, xsw, 2))) return rc;
  if ((rc = tc_make_act_map(&mx_lo, x_lo, d->N, d->H, d->W, d->Cin, d->x_cstride, false, 16, box_h, 1, xsw, 2))) return rc;
  const int ysw = (IMMB_WG_NCAT && bn == 32) ? 2 : 1;       // N concatenation at BN = 32: 32-channel boxes, 64-byte swizzle
  if ((rc = tc_make_act_map(&my_hi, dy_hi, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, 8, 4, 1, ysw, 2))) return rc;
  if ((rc = tc_make_act_map(&my_lo, dy_lo, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, 8, 4, 1, ysw, 2))) return rc;
  dim3 grid(c_tiles, n_tiles, (rpc == 3 ? 1 : 3) * splits);
  if (bn == 128) return launch_wg16<128, 1>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
  if (bn == 64) return launch_wg16<64, 1>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
  if (c32) return launch_wg16<32, 3, true>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
  return launch_wg16<32, 3>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
}

// =============================================================================================================
// fp16 wgrad of the STRIDE-2 3x3 layers (TF SAME on even sizes: no padding before, one row / column after):
//   dW[r][s][ci][co] = sum_{n,i,j} X[n, 2i+r, 2j+s, ci] * dY[n, i, j, co]
// X is read through its parity-split view (c + wpar*C, w/2, hpar, h/2, n): tap (r, s) is the view's channel block
// wpar = s & 1 at row parity hpar = r & 1, shifted by (r >> 1, s >> 1) view pixels -- so each tap is again a plain
// shifted window and the kernel is the stride-1 one with a different operand map.  One CTA = one filter row r, one
// chunk of input channels, BN output channels, a split-K range of 4x8-pixel dY tiles.
//   C32 (C = 32): one 128-byte view row holds BOTH column parities of a pixel pair = taps s = 0 | 1 (32 rows each);
//       the same window one view pixel further = s = 2 | (unused): ONE MMA of M = 128 covers the filter row.
//   else (64 | C): the wpar = 0 and wpar = 1 chunks are separate windows: MMA group 0 = the wpar-0 window at shifts
//       0 and 1 (s = 0, 2), group 1 = the wpar-1 window (s = 1, upper atom unused).
// =============================================================================================================
template <int BN, int STAGES, bool C32>
struct Wg16S2Cfg {
  static constexpr uint32_t NWIN = C32 ? 1 : 2;                           // X windows per plane per stage
  static constexpr uint32_t A_WIN = 4 * 16 * 128;                         // 4 view rows x 16 view pixels x 128 B
  static constexpr uint32_t A_PLANE = NWIN * A_WIN;
  static constexpr uint32_t NB = (BN + 63) / 64;
  static constexpr uint32_t B_PLANE = NB * 4096;
  static constexpr uint32_t STAGE_BYTES = (A_PLANE + B_PLANE) * 2;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int GROUPS = C32 ? 1 : 2;
  static constexpr int ACC_COLS = GROUPS * 2 * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512));
  static_assert(ACC_COLS <= 512, "accumulators");
};

template <int BN, int STAGES, bool C32>
__global__ void __launch_bounds__(192, 1)
conv_tc2_wgrad16_s2_kernel(const __grid_constant__ CUtensorMap mapX_hi, const __grid_constant__ CUtensorMap mapX_lo,
                           const __grid_constant__ CUtensorMap mapY_hi, const __grid_constant__ CUtensorMap mapY_lo,
                           const __grid_constant__ Wg16Params p) {
  using Cfg = Wg16S2Cfg<BN, STAGES, C32>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int ci0 = C32 ? 0 : blockIdx.x * 64;
  const int n_off = blockIdx.y * BN;
  const int r = blockIdx.z / p.splits;
  const int split = blockIdx.z - r * p.splits;
  const int hpar = r & 1, dh = r >> 1;
  const int t_begin = split * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.total_tiles) t_end = p.total_tiles;
  const int num_k = t_end - t_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
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
#pragma unroll
        for (uint32_t wv = 0; wv < Cfg::NWIN; ++wv) {
          const int c = ci0 + (int)wv * p.Cin;                            // view channel of the wpar = wv chunk
          tma_load_5d(st + wv * Cfg::A_WIN, &mapX_hi, &full[s], c, twi * 8, hpar, thi * 4 + dh, n);
          tma_load_5d(st + Cfg::A_PLANE + wv * Cfg::A_WIN, &mapX_lo, &full[s], c, twi * 8, hpar, thi * 4 + dh, n);
        }
        uint8_t* sb = st + Cfg::A_PLANE * 2;
#pragma unroll
        for (uint32_t j = 0; j < Cfg::NB; ++j) {
          tma_load_5d(sb + j * 4096, &mapY_hi, &full[s], n_off + (int)j * 64, twi * 8, 0, thi * 4, n);
          tma_load_5d(sb + Cfg::B_PLANE + j * 4096, &mapY_lo, &full[s], n_off + (int)j * 64, twi * 8, 0, thi * 4, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (num_k > 0) {
      constexpr uint32_t idesc = idesc_f16(128, BN, 1, 1);
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t a_lo = a_hi + Cfg::A_PLANE;
        const uint32_t b_hi = a_hi + Cfg::A_PLANE * 2;
        const uint32_t b_lo = b_hi + Cfg::B_PLANE;
        if (elect_one()) {
          const uint64_t da0_hi = smem_desc_sw128(a_hi, 128, 2048, 2), db0_hi = smem_desc_sw128(b_hi, 4096, 1024, 2);
          const uint64_t da0_lo = da0_hi + (uint64_t)(Cfg::A_PLANE >> 4), db0_lo = db0_hi + (uint64_t)(Cfg::B_PLANE >> 4);
#pragma unroll
          for (int j = 0; j < 2; ++j) {                                   // dY rows 2j, 2j+1 of the tile = K 16
            const uint32_t acc = (kt > 0 || j > 0) ? 1u : 0u;
            const uint64_t db_hi = db0_hi + (uint64_t)(j * (2048 >> 4));
            const uint64_t db_lo = db0_lo + (uint64_t)(j * (2048 >> 4));
#pragma unroll
            for (int g = 0; g < Cfg::GROUPS; ++g) {
              // window g, view rows 2j / 2j+1; MN atoms = the window at view-pixel shifts 0 and 1 (LBO = 128 B)
              const uint64_t ao = (uint64_t)(((uint32_t)g * Cfg::A_WIN + (uint32_t)(2 * j) * 2048u) >> 4);
              const uint64_t da_hi = da0_hi + ao;
              const uint64_t da_lo = da0_lo + ao;
              const uint32_t d_main = tmem_base + (uint32_t)(g * 2 * BN);
              const uint32_t d_cross = d_main + (uint32_t)BN;
              if (IMMB_WG_NCAT && BN >= 64) {                               // the lo plane follows the hi plane: N = 2*BN
                mma_f16(d_main, da_hi, db_hi, idesc_f16(128, 2 * BN, 1, 1), acc);
              } else {
                mma_f16(d_main, da_hi, db_hi, idesc, acc);
                mma_f16(d_cross, da_hi, db_lo, idesc, acc);
              }
              mma_f16(d_cross, da_lo, db_hi, idesc, 1u);
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
    const int atom = m >> 6;                   // view-pixel shift of this TMEM lane's MN atom
    const int e = __ldg(p.x_scale) + __ldg(p.dy_scale);
    const float s_main = exp2i(-e), s_cross = exp2i(-e - 11);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int g = 0; g < Cfg::GROUPS; ++g) {
      // column tap and input channel of this lane
      int s_tap, ci;
      if (C32) { s_tap = 2 * atom + ((m >> 5) & 1); ci = m & 31; }
      else { s_tap = g == 0 ? 2 * atom : (atom == 0 ? 1 : 3); ci = ci0 + (m & 63); }
      const bool row_ok = (s_tap < 3) && (ci < p.Cin);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32], v2[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 2 * BN + c0), v);
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 2 * BN + BN + c0), v2);
        tmem_ld_wait();
        if (!row_ok) continue;
        const int col0 = n_off + c0;
        if (col0 >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(v2[j], s_cross, v[j] * s_main);
        float* o = p.dw + ((size_t)(r * 3 + s_tap) * p.Cin + ci) * p.Cout + col0;
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool C32>
static int launch_wg16_s2(const CUtensorMap& x_hi, const CUtensorMap& x_lo, const CUtensorMap& y_hi,
                          const CUtensorMap& y_lo, const Wg16Params& p, dim3 grid, cudaStream_t st) {
  constexpr int STAGES = BN <= 64 ? 6 : 4;
  using Cfg = Wg16S2Cfg<BN, STAGES, C32>;
  auto kern = conv_tc2_wgrad16_s2_kernel<BN, STAGES, C32>;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device_slot()];      // the shared-memory opt-in is a per-device attribute
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2_wgrad16_s2 smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(x_hi, x_lo, y_hi, y_lo, p);
  return check_launch("conv_tc2_wgrad16_s2_kernel");
}

// stride-2 3x3, even sizes (TF SAME: pads before = 0), contiguous input channels, C = 32 or a multiple of 64
bool conv_tc2_wgrad_s2_eligible(const immb_conv_desc* d) {
  if (!conv_tc2_enabled()) return false;
  if (d->x_layout != IMMB_XLAYOUT_NHWC || d->kh != 3 || d->kw != 3 || d->stride != 2) return false;
  if (d->H % 2 || d->W % 2 || d->pad_t != 0 || d->pad_l != 0) return false;
  if (d->Ho % 4 || d->Wo % 8 || d->x_cstride != d->Cin) return false;
  return d->Cin == 32 || d->Cin % 64 == 0;
}

static int conv_tc2_wgrad16_s2_run(const immb_conv_desc* d, const void* x_hi, const void* x_lo, const void* dy_hi,
                                   const void* dy_lo, float* dw, cudaStream_t st) {
  if (!d->x_scale || !d->y_scale) return set_error(IMMB_ERR_INVALID, "conv_tc2_wgrad: fp16 planes need x_scale / y_scale");
  if (d->y_cstride % 8) return set_error(IMMB_ERR_UNSUPPORTED, "conv_tc2_wgrad: fp16 planes need 8-channel strides");
  Wg16Params p;
  memset(&p, 0, sizeof(p));
  p.tiles_w = d->Wo / 8; p.tiles_h = d->Ho / 4; p.n_img = d->N;
  p.total_tiles = p.tiles_w * p.tiles_h * d->N;
  p.dw = dw; p.Cin = d->Cin; p.Cout = d->Cout;
  p.x_scale = d->x_scale; p.dy_scale = d->y_scale;
  const bool c32 = d->Cin == 32;
  const int bn = d->Cout > 64 ? 128 : (d->Cout > 32 ? 64 : 32);
  const int n_tiles = ceil_div(d->Cout, bn);
  const int c_tiles = c32 ? 1 : d->Cin / 64;
  int splits = kNumSMs / (c_tiles * n_tiles * 3);
  if (splits > p.total_tiles) splits = p.total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = ceil_div(p.total_tiles, splits);
  splits = ceil_div(p.total_tiles, p.tiles_per_split);
  p.splits = splits;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)9 * d->Cin * d->Cout, st);
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "wgrad memset: %s", cudaGetErrorString(e));
  CUtensorMap mx_hi, mx_lo, my_hi, my_lo;
  int rc;
  // parity-split view of X, 16 view pixels x 4 view rows per window
  if ((rc = tc_make_act_map(&mx_hi, x_hi, d->N, d->H, d->W, d->Cin, d->x_cstride, true, 16, 4, 1, 1, 2))) return rc;
  if ((rc = tc_make_act_map(&mx_lo, x_lo, d->N, d->H, d->W, d->Cin, d->x_cstride, true, 16, 4, 1, 1, 2))) return rc;
  if ((rc = tc_make_act_map(&my_hi, dy_hi, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, 8, 4, 1, 1, 2))) return rc;
  if ((rc = tc_make_act_map(&my_lo, dy_lo, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, 8, 4, 1, 1, 2))) return rc;
  dim3 grid(c_tiles, n_tiles, 3 * splits);
  if (c32) {
    if (bn == 128) return launch_wg16_s2<128, true>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
    if (bn == 64) return launch_wg16_s2<64, true>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
    return launch_wg16_s2<32, true>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
  }
  if (bn == 128) return launch_wg16_s2<128, false>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
  if (bn == 64) return launch_wg16_s2<64, false>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
  return launch_wg16_s2<32, false>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
}

bool conv_tc2_wgrad_eligible(const immb_conv_desc* d) {
  if (!conv_tc2_enabled()) return false;
  if (prec_is_f16(d->precision) && conv_tc2_wgrad_s2_eligible(d)) return true;
  if (d->x_layout != IMMB_XLAYOUT_NHWC || d->kh != 3 || d->kw != 3 || d->stride != 1) return false;
  if (d->H % 4 || d->W % 8 || d->pad_t != 1 || d->pad_l != 1) return false;
  return true;
}

template <int BN, int PASSES>
static int launch_wg2(const CUtensorMap& x_hi, const CUtensorMap& x_lo, const CUtensorMap& y_hi,
                      const CUtensorMap& y_lo, const Wg2Params& p, dim3 grid, cudaStream_t st) {
  constexpr int STAGES = PASSES == 3 ? (BN <= 64 ? 4 : 3) : 4;
  using Cfg = Wg2Cfg<BN, PASSES, STAGES>;
  auto kern = conv_tc2_wgrad_kernel<BN, PASSES, STAGES>;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device_slot()];      // the shared-memory opt-in is a per-device attribute
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "conv_tc2_wgrad smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(x_hi, x_lo, y_hi, y_lo, p);
  return check_launch("conv_tc2_wgrad_kernel");
}

int conv_tc2_wgrad_run(const immb_conv_desc* d, const void* x_hi_, const void* x_lo_, const void* dy_hi_,
                       const void* dy_lo_, float* dw, cudaStream_t st) {
  if (prec_is_f16(d->precision) && d->stride == 2) return conv_tc2_wgrad16_s2_run(d, x_hi_, x_lo_, dy_hi_, dy_lo_, dw, st);
  if (prec_is_f16(d->precision)) return conv_tc2_wgrad16_run(d, x_hi_, x_lo_, dy_hi_, dy_lo_, dw, st);
  const float *x_hi = (const float*)x_hi_, *x_lo = (const float*)x_lo_, *dy_hi = (const float*)dy_hi_, *dy_lo = (const float*)dy_lo_;
  const int passes = d->precision == IMMB_PREC_TF32 ? 1 : 3;
  Wg2Params p;
  memset(&p, 0, sizeof(p));
  p.tiles_w = d->W / 8; p.tiles_h = d->H / 4; p.n_img = d->N;
  p.total_tiles = p.tiles_w * p.tiles_h * d->N;
  p.dw = dw; p.Cin = d->Cin; p.Cout = d->Cout;
  const int bn = d->Cout % 128 == 0 ? 128 : (d->Cout % 64 == 0 ? 64 : 32);
  const int n_tiles = ceil_div(d->Cout, bn);
  const int c_tiles = d->cin_pad / 32;
  // split-K factor: the grid must not exceed a whole number of waves (1 CTA per SM: 19 splits x 16 tiles = 304 CTAs
  // used to run as 148 + 148 + 8, a third round for 3 % of the work)
  static int wg_waves = -1;
  if (wg_waves < 0) { const char* e = getenv("IMMB_WG_WAVES"); wg_waves = e ? atoi(e) : 1; if (wg_waves < 1) wg_waves = 1; }
  int splits = (kNumSMs * wg_waves) / (c_tiles * n_tiles);
  if (splits > p.total_tiles) splits = p.total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = ceil_div(p.total_tiles, splits);
  splits = ceil_div(p.total_tiles, p.tiles_per_split);
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)9 * d->Cin * d->Cout, st);
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "wgrad memset: %s", cudaGetErrorString(e));
  CUtensorMap mx_hi, mx_lo, my_hi, my_lo;
  int rc;
  if ((rc = tc_make_act_map(&mx_hi, x_hi, d->N, d->H, d->W, d->Cin, d->x_cstride, false, 16, 6, 1, 1))) return rc;
  if ((rc = tc_make_act_map(&my_hi, dy_hi, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, 8, 4, 1, 1))) return rc;
  mx_lo = mx_hi; my_lo = my_hi;
  if (passes == 3) {
    if ((rc = tc_make_act_map(&mx_lo, x_lo, d->N, d->H, d->W, d->Cin, d->x_cstride, false, 16, 6, 1, 1))) return rc;
    if ((rc = tc_make_act_map(&my_lo, dy_lo, d->N, d->Ho, d->Wo, d->Cout, d->y_cstride, false, 8, 4, 1, 1))) return rc;
  }
  dim3 grid(c_tiles, n_tiles, splits);
#define IMMB_WG2(BN_)                                                                          \
  if (bn == BN_)                                                                               \
    return passes == 3 ? launch_wg2<BN_, 3>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st)           \
                       : launch_wg2<BN_, 1>(mx_hi, mx_lo, my_hi, my_lo, p, grid, st);
  IMMB_WG2(32)
  IMMB_WG2(64)
  IMMB_WG2(128)
#undef IMMB_WG2
  return set_error(IMMB_ERR_INVALID, "conv_tc2_wgrad: bn");
}

}  // namespace immb
