// fp32 CUDA-core implicit-GEMM convolution (forward / dgrad / wgrad).
// Used for the shapes the tcgen05 engine does not take (Cin = 3 / 1, Cout = 9 / n_maps: <2% of the step
// FLOPs) and as the on-device cross-check of the tensor-core engine.  C[M,N] = sum_k A(m,k) B(k,n) with a
// 64x64x16 block tile, 256 threads, 4x4 register tile per thread; A/B elements are gathered on the fly.
#include "common.cuh"

namespace immb {

constexpr int BM = 64, BN = 64, BK = 16;

struct ConvP {
  int N, H, W, Cin, Cout, kh, kw, stride, Ho, Wo, pad_t, pad_l, xcs, ycs;
  int M, Nn, K;          // GEMM extents
  int relu;              // fwd epilogue
  long long k_per_split; // wgrad split-K
};

enum { OP_FWD = 0, OP_DGRAD = 1, OP_WGRAD = 2 };

// ---- element fetchers --------------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ float fetch_a(const ConvP& p, const float* __restrict__ a_hi,
                                         const float* __restrict__ a_lo, int m, long long k) {
  if (OP == OP_FWD) {
    // m = (n, ho, wo); k = (tap, ci)
    if (m >= p.M || k >= p.K) return 0.f;
    int ci = (int)(k % p.Cin);
    int tap = (int)(k / p.Cin);
    int r = tap / p.kw, s = tap - r * p.kw;
    int wo = m % p.Wo;
    int t = m / p.Wo;
    int ho = t % p.Ho;
    int n = t / p.Ho;
    int h = ho * p.stride + r - p.pad_t, w = wo * p.stride + s - p.pad_l;
    if (h < 0 || h >= p.H || w < 0 || w >= p.W) return 0.f;
    return load_split(a_hi, a_lo, (((size_t)n * p.H + h) * p.W + w) * p.xcs + ci);
  } else if (OP == OP_DGRAD) {
    // m = (n, h, w) input pixel; k = (tap, co); A = dy[n, (h+pt-r)/st, (w+pl-s)/st, co]
    if (m >= p.M || k >= p.K) return 0.f;
    int co = (int)(k % p.Cout);
    int tap = (int)(k / p.Cout);
    int r = tap / p.kw, s = tap - r * p.kw;
    int w = m % p.W;
    int t = m / p.W;
    int h = t % p.H;
    int n = t / p.H;
    int hn = h + p.pad_t - r, wn = w + p.pad_l - s;
    if (hn < 0 || wn < 0) return 0.f;
    if (p.stride > 1 && ((hn % p.stride) || (wn % p.stride))) return 0.f;
    int ho = hn / p.stride, wo = wn / p.stride;
    if (ho >= p.Ho || wo >= p.Wo) return 0.f;
    return load_split(a_hi, a_lo, (((size_t)n * p.Ho + ho) * p.Wo + wo) * p.ycs + co);
  } else {
    // wgrad: m = (tap, ci); k = output pixel (n, ho, wo); A = x[n, ho*st+r-pt, wo*st+s-pl, ci]
    if (m >= p.M || k >= p.K) return 0.f;
    int ci = m % p.Cin;
    int tap = m / p.Cin;
    int r = tap / p.kw, s = tap - r * p.kw;
    int wo = (int)(k % p.Wo);
    long long t = k / p.Wo;
    int ho = (int)(t % p.Ho);
    int n = (int)(t / p.Ho);
    int h = ho * p.stride + r - p.pad_t, w = wo * p.stride + s - p.pad_l;
    if (h < 0 || h >= p.H || w < 0 || w >= p.W) return 0.f;
    return load_split(a_hi, a_lo, (((size_t)n * p.H + h) * p.W + w) * p.xcs + ci);
  }
}

template <int OP>
__device__ __forceinline__ float fetch_b(const ConvP& p, const float* __restrict__ b_hi,
                                         const float* __restrict__ b_lo, long long k, int n) {
  if (k >= p.K || n >= p.Nn) return 0.f;
  if (OP == OP_FWD) {
    // w[tap][ci][co] : k = tap*Cin+ci
    return __ldg(b_hi + (size_t)k * p.Cout + n);
  } else if (OP == OP_DGRAD) {
    // k = (tap, co), n = ci -> w[tap][ci][co]
    int co = (int)(k % p.Cout);
    int tap = (int)(k / p.Cout);
    return __ldg(b_hi + ((size_t)tap * p.Cin + n) * p.Cout + co);
  } else {
    // dy[pixel k][co n]
    return load_split(b_hi, b_lo, (size_t)k * p.ycs + n);
  }
}

template <int OP>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvP p, const float* __restrict__ a_hi,
                                                        const float* __restrict__ a_lo,
                                                        const float* __restrict__ b_hi,
                                                        const float* __restrict__ b_lo,
                                                        const float* __restrict__ bias, float* out_hi,
                                                        float* out_lo) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  long long k_begin = 0, k_end = p.K;
  if (OP == OP_WGRAD) {
    k_begin = (long long)blockIdx.z * p.k_per_split;
    k_end = k_begin + p.k_per_split;
    if (k_end > p.K) k_end = p.K;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
    // A tile: BM x BK.  Contiguity: fwd/dgrad along k (channels); wgrad along m (channels).
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm, kk;
      if (OP == OP_WGRAD) {
        mm = tid & 63;
        kk = (tid >> 6) + 4 * i;
      } else {
        kk = tid & 15;
        mm = (tid >> 4) + 16 * i;
      }
      long long kg = k0 + kk;
      As[kk][mm] = (kg < k_end) ? fetch_a<OP>(p, a_hi, a_lo, m0 + mm, kg) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int nn, kk;
      if (OP == OP_DGRAD) {
        kk = tid & 15;
        nn = (tid >> 4) + 16 * i;
      } else {
        nn = tid & 63;
        kk = (tid >> 6) + 4 * i;
      }
      long long kg = k0 + kk;
      Bs[kk][nn] = (kg < k_end) ? fetch_b<OP>(p, b_hi, b_lo, kg, n0 + nn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.Nn) continue;
      float v = acc[i][j];
      if (OP == OP_FWD) {
        if (bias) v += __ldg(bias + n);
        if (p.relu) v = fmaxf(v, 0.f);
        store_split(out_hi, out_lo, (size_t)m * p.ycs + n, v);
      } else if (OP == OP_DGRAD) {
        out_hi[(size_t)m * p.xcs + n] = v;
      } else {
        atomicAdd(out_hi + (size_t)m * p.Cout + n, v);   // dw[(tap,ci)][co], split-K partials
      }
    }
  }
}

static ConvP make_p(const immb_conv_desc* d) {
  ConvP p;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.kh = d->kh; p.kw = d->kw;
  p.stride = d->stride; p.Ho = d->Ho; p.Wo = d->Wo; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  p.xcs = d->x_cstride; p.ycs = d->y_cstride; p.relu = (d->epilogue == IMMB_EPI_BIAS_RELU);
  p.M = p.Nn = p.K = 0; p.k_per_split = 0;
  return p;
}

int conv_simt_fwd(const immb_conv_desc* d, const float* x_hi, const float* x_lo, const float* w,
                  const float* bias, float* y_hi, float* y_lo, cudaStream_t st) {
  ConvP p = make_p(d);
  p.M = d->N * d->Ho * d->Wo; p.Nn = d->Cout; p.K = d->kh * d->kw * d->Cin;
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.Nn, BN));
  conv_simt_kernel<OP_FWD><<<grid, 256, 0, st>>>(p, x_hi, x_lo, w, nullptr, bias, y_hi, y_lo);
  return check_launch("conv_simt_fwd");
}

int conv_simt_dgrad(const immb_conv_desc* d, const float* dy_hi, const float* dy_lo, const float* w,
                    float* dx, cudaStream_t st) {
  ConvP p = make_p(d);
  p.M = d->N * d->H * d->W; p.Nn = d->Cin; p.K = d->kh * d->kw * d->Cout;
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.Nn, BN));
  conv_simt_kernel<OP_DGRAD><<<grid, 256, 0, st>>>(p, dy_hi, dy_lo, w, nullptr, nullptr, dx, nullptr);
  return check_launch("conv_simt_dgrad");
}

int conv_simt_wgrad(const immb_conv_desc* d, const float* x_hi, const float* x_lo, const float* dy_hi,
                    const float* dy_lo, float* dw, cudaStream_t st) {
  ConvP p = make_p(d);
  p.M = d->kh * d->kw * d->Cin; p.Nn = d->Cout; p.K = d->N * d->Ho * d->Wo;
  int tiles = ceil_div(p.M, BM) * ceil_div(p.Nn, BN);
  int splits = (kNumSMs * 4 + tiles - 1) / tiles;
  long long max_splits = (p.K + 4 * BK - 1) / (4 * BK);
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  p.k_per_split = (((long long)p.K + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (int)((p.K + p.k_per_split - 1) / p.k_per_split);
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p.M * p.Nn, st);
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "wgrad memset: %s", cudaGetErrorString(e));
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.Nn, BN), splits);
  conv_simt_kernel<OP_WGRAD><<<grid, 256, 0, st>>>(p, x_hi, x_lo, dy_hi, dy_lo, nullptr, dw, nullptr);
  return check_launch("conv_simt_wgrad");
}

}  // namespace immb
