// Bandwidth-bound kernels of the IMM hot path: batch norm (stats / apply / backward), legacy bilinear
// upsample, landmark bottleneck (softargmax + Gaussian maps), perceptual-loss glue, max-pool, and the
// fused clip+Adam optimiser.  All NHWC fp32; channel index is the fastest-moving thread index so every
// warp-level access is a contiguous 128-byte line.
#include "common.cuh"

namespace immb {

constexpr float kBnEps = 1e-3f;       // tf.layers.batch_normalization default epsilon
constexpr float kBnMomentum = 0.99f;  // default momentum

// =====================================================================================================
// per-channel reductions over pixels.  block = (32 channels, 8 pixel rows); grid = (pixel chunks, C/32)
// =====================================================================================================
constexpr int kRedRows = 8;
constexpr int kRedPixPerBlock = 512;

template <int NV, class F>
__device__ __forceinline__ void channel_reduce(int64_t npix, int C, double* out, int out_stride, F f) {
  const int c = blockIdx.y * 32 + threadIdx.x;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  if (c < C) {
    int64_t p0 = (int64_t)blockIdx.x * kRedPixPerBlock;
    int64_t p1 = p0 + kRedPixPerBlock;
    if (p1 > npix) p1 = npix;
    for (int64_t p = p0 + threadIdx.y; p < p1; p += kRedRows) {
      float v[NV];
      f(p, c, v);
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] += (double)v[i];
    }
  }
  __shared__ double sm[NV][kRedRows][33];
#pragma unroll
  for (int i = 0; i < NV; ++i) sm[i][threadIdx.y][threadIdx.x] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < kRedRows; ++r) s += sm[i][r][threadIdx.x];
      atomicAdd(out + (size_t)i * out_stride + c, s);
    }
  }
}

__global__ void bn_stats_kernel(const float* __restrict__ y, int64_t npix, int C, int ycs, double* sums) {
  channel_reduce<2>(npix, C, sums, C, [&](int64_t p, int c, float* v) {
    float t = __ldg(y + p * ycs + c);
    v[0] = t;
    v[1] = t * t;
  });
}

// sumsq in fp32 of a single value is exact enough; accumulation is in double.  (t*t rounds once.)

__device__ __forceinline__ void bn_finalize_channel(int c, int training, double sum, double sumsq, double count,
                                                    const float* gamma, const float* beta, float* mm, float* mv,
                                                    float* scale, float* shift, float* mean_o, float* invstd_o);

__global__ void bn_finalize_kernel(const double* sums, double count, int C, const float* gamma,
                                   const float* beta, float* mm, float* mv, int training, float* scale,
                                   float* shift, float* mean_o, float* invstd_o) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  bn_finalize_channel(c, training, training ? sums[c] : 0.0, training ? sums[C + c] : 0.0, count, gamma, beta, mm, mv,
                      scale, shift, mean_o, invstd_o);
}

__device__ __forceinline__ void bn_finalize_channel(int c, int training, double sum, double sumsq, double count,
                                                    const float* gamma, const float* beta, float* mm, float* mv,
                                                    float* scale, float* shift, float* mean_o, float* invstd_o) {
  float mean, var;
  if (training) {
    double m = sum / count;
    double v = sumsq / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    double vu = v * (count / (count > 1.0 ? count - 1.0 : 1.0));
    mm[c] = mm[c] - (mm[c] - mean) * (1.0f - kBnMomentum);
    mv[c] = mv[c] - (mv[c] - (float)vu) * (1.0f - kBnMomentum);
  } else {
    mean = mm[c];
    var = mv[c];
  }
  float inv = rsqrtf(var + kBnEps);
  // one Newton step: rsqrtf is ~2 ulp; make it correctly rounded to match the CPU rsqrt path
  inv = inv * (1.5f - 0.5f * (var + kBnEps) * inv * inv);
  float sc = gamma[c] * inv;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_o[c] = mean;
  invstd_o[c] = inv;
}

__device__ __forceinline__ float bn_act(float y, float sc, float sh, int relu) {
  float z = fmaf(y, sc, sh);
  return (relu && z < 0.f) ? 0.f : z;
}

__global__ void bn_apply_kernel(const float* __restrict__ y, int64_t npix, int C, int ycs,
                                const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                float* out_hi, float* out_lo, int ocs) {
  int64_t total = npix * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C;
    int c = (int)(i - p * C);
    float v = bn_act(__ldg(y + p * ycs + c), __ldg(scale + c), __ldg(shift + c), relu);
    store_split(out_hi, out_lo, (size_t)(p * ocs + c), v);
  }
}

// fused [relu](bn(y)) + TF1 legacy bilinear x2 (src = dst/2, no half-pixel): out[2i]=a[i], out[2i+1]=(a[i]+a[min(i+1,n-1)])/2
__global__ void bn_apply_up2x_kernel(const float* __restrict__ y, int N, int H, int W, int C, int ycs,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     int relu, float* out_hi, float* out_lo, int ocs) {
  const int Ho = 2 * H, Wo = 2 * W;
  int64_t total = (int64_t)N * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int wo = (int)(p % Wo);
    int64_t q = p / Wo;
    int ho = (int)(q % Ho);
    int n = (int)(q / Ho);
    int h0 = ho >> 1, w0 = wo >> 1;
    int h1 = min(h0 + 1, H - 1), w1 = min(w0 + 1, W - 1);
    float fh = (ho & 1) ? 0.5f : 0.f, fw = (wo & 1) ? 0.5f : 0.f;
    float sc = __ldg(scale + c), sh = __ldg(shift + c);
    const float* base = y + (int64_t)n * H * W * ycs + c;
    float a00 = bn_act(__ldg(base + ((int64_t)h0 * W + w0) * ycs), sc, sh, relu);
    float a01 = bn_act(__ldg(base + ((int64_t)h0 * W + w1) * ycs), sc, sh, relu);
    float a10 = bn_act(__ldg(base + ((int64_t)h1 * W + w0) * ycs), sc, sh, relu);
    float a11 = bn_act(__ldg(base + ((int64_t)h1 * W + w1) * ycs), sc, sh, relu);
    float top = a00 + (a01 - a00) * fw;
    float bot = a10 + (a11 - a10) * fw;
    float v = top + (bot - top) * fh;
    store_split(out_hi, out_lo, (size_t)(p * ocs + c), v);
  }
}

__global__ void upsample2x_bwd_kernel(const float* __restrict__ gup, int N, int H, int W, int C, int gcs,
                                      float* g) {
  const int Ho = 2 * H, Wo = 2 * W;
  int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int w = (int)(p % W);
    int64_t q = p / W;
    int h = (int)(q % H);
    int n = (int)(q / H);
    const float* base = gup + (int64_t)n * Ho * Wo * gcs + c;
    float acc = 0.f;
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      int hh = 2 * h + dh;
      if (hh < 0) continue;
      float wh = dh == 0 ? 1.f : (dh < 0 ? 0.5f : (h == H - 1 ? 1.f : 0.5f));
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        int ww = 2 * w + dw;
        if (ww < 0) continue;
        float wwt = dw == 0 ? 1.f : (dw < 0 ? 0.5f : (w == W - 1 ? 1.f : 0.5f));
        acc += wh * wwt * __ldg(base + ((int64_t)hh * Wo + ww) * gcs);
      }
    }
    g[i] = acc;
  }
}

__global__ void bn_bwd_reduce_kernel(const float* __restrict__ g, int gcs, const float* __restrict__ y,
                                     int ycs, int64_t npix, int C, const float* __restrict__ scale,
                                     const float* __restrict__ shift, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, int relu, double* sums) {
  channel_reduce<2>(npix, C, sums, C, [&](int64_t p, int c, float* v) {
    float yy = __ldg(y + p * ycs + c);
    float gg = __ldg(g + p * gcs + c);
    float z = fmaf(yy, __ldg(scale + c), __ldg(shift + c));
    float dz = (relu && !(z > 0.f)) ? 0.f : gg;
    float xh = (yy - __ldg(mean + c)) * __ldg(invstd + c);
    v[0] = dz;
    v[1] = dz * xh;
  });
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ g, int gcs, const float* __restrict__ y,
                                    int ycs, int64_t npix, int C, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, int relu, const double* __restrict__ sums,
                                    float* dy_hi, float* dy_lo, float* dgamma, float* dbeta,
                                    double* dbias_acc) {
  const double inv_n = 1.0 / (double)npix;
  if (blockIdx.x == 0 && threadIdx.y == 0) {
    int c = blockIdx.y * 32 + threadIdx.x;
    if (c < C) {
      dbeta[c] = (float)sums[c];
      dgamma[c] = (float)sums[C + c];
    }
  }
  channel_reduce<1>(npix, C, dbias_acc, C, [&](int64_t p, int c, float* v) {
    float yy = __ldg(y + p * ycs + c);
    float gg = __ldg(g + p * gcs + c);
    float sc = __ldg(scale + c);
    float z = fmaf(yy, sc, __ldg(shift + c));
    float dz = (relu && !(z > 0.f)) ? 0.f : gg;
    float xh = (yy - __ldg(mean + c)) * __ldg(invstd + c);
    float mdz = (float)(sums[c] * inv_n);
    float mdzx = (float)(sums[C + c] * inv_n);
    float dy = sc * (dz - mdz - xh * mdzx);
    store_split(dy_hi, dy_lo, (size_t)(p * C + c), dy);
    v[0] = dy;
  });
}

__global__ void cast_d2f_kernel(const double* s, float* d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (float)s[i];
}

__global__ void split_planes_kernel(const float* __restrict__ v, float* hi, float* lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    store_split(hi, lo, (size_t)i, __ldg(v + i));
}


// =====================================================================================================
// float4-vectorised variants (C % 4 == 0, 16-byte aligned rows): one thread = 4 consecutive channels.
// =====================================================================================================
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store_split4(float* hi_p, float* lo_p, size_t i, float4 v) {
  if (lo_p) {
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    st4(hi_p + i, h);
    st4(lo_p + i, l);
  } else {
    st4(hi_p + i, v);
  }
}
__device__ __forceinline__ float4 load_split4(const float* hi_p, const float* lo_p, size_t i) {
  float4 v = ld4(hi_p + i);
  if (lo_p) {
    float4 l = ld4(lo_p + i);
    v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
  }
  return v;
}

// -----------------------------------------------------------------------------------------------------
// Split-plane accessors.  Every kernel that reads or writes conv operands is a template on the plane format:
//   Pl<false>: fp32 storage, TF32 pair (hi, lo); lo == nullptr means a plain fp32 tensor in `hi`;
//   Pl<true> : scaled fp16 pair ("H16 planes", common.cuh) with the tensor's scale record {e, amax bits}.
// load*: the real value (planes joined, scale removed).  store*: split + write, tracking max |v| in `amax`, which the
// kernel hands to finish() once, at its end, with whole warps converged.
// -----------------------------------------------------------------------------------------------------
template <bool H16>
struct Pl;
template <>
struct Pl<false> {
  float* hi;
  float* lo;
  __device__ __forceinline__ void init() {}
  __device__ __forceinline__ float load(size_t i) const { return load_split(hi, lo, i); }
  __device__ __forceinline__ float4 load4(size_t i) const { return load_split4(hi, lo, i); }
  __device__ __forceinline__ void store(size_t i, float v, float&) const { store_split(hi, lo, i, v); }
  __device__ __forceinline__ void store4(size_t i, float4 v, float&) const { store_split4(hi, lo, i, v); }
  __device__ __forceinline__ void finish(float) const {}
};
template <>
struct Pl<true> {
  uint16_t* hi;
  uint16_t* lo;
  int32_t* rec;
  float mul, inv;
  __device__ __forceinline__ void init() {
    const int e = __ldg(rec);
    mul = exp2i(e);
    inv = exp2i(-e);
  }
  __device__ __forceinline__ float load(size_t i) const { return join_h16(__ldg(hi + i), __ldg(lo + i)) * inv; }
  __device__ __forceinline__ float4 load4(size_t i) const {
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + i)), l = __ldg(reinterpret_cast<const uint2*>(lo + i));
    return make_float4(join_h16((uint16_t)(h.x & 0xFFFFu), (uint16_t)(l.x & 0xFFFFu)) * inv,
                       join_h16((uint16_t)(h.x >> 16), (uint16_t)(l.x >> 16)) * inv,
                       join_h16((uint16_t)(h.y & 0xFFFFu), (uint16_t)(l.y & 0xFFFFu)) * inv,
                       join_h16((uint16_t)(h.y >> 16), (uint16_t)(l.y >> 16)) * inv);
  }
  __device__ __forceinline__ void store(size_t i, float v, float& amax) const {
    uint16_t h, l;
    amax = fmaxf(amax, fabsf(v));
    split_h16(v * mul, h, l);
    hi[i] = h;
    lo[i] = l;
  }
  __device__ __forceinline__ void store4(size_t i, float4 v, float& amax) const {
    uint16_t h0, l0, h1, l1, h2, l2, h3, l3;
    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    split_h16(v.x * mul, h0, l0);
    split_h16(v.y * mul, h1, l1);
    split_h16(v.z * mul, h2, l2);
    split_h16(v.w * mul, h3, l3);
    *reinterpret_cast<uint2*>(hi + i) = make_uint2(pack2(h0, h1), pack2(h2, h3));
    *reinterpret_cast<uint2*>(lo + i) = make_uint2(pack2(l0, l1), pack2(l2, l3));
  }
  __device__ __forceinline__ void finish(float amax) const { h16_track_amax(rec, amax); }
};
template <bool H16>
static inline Pl<H16> make_pl(const void* hi, const void* lo, const int32_t* rec);
template <>
inline Pl<false> make_pl<false>(const void* hi, const void* lo, const int32_t*) {
  return Pl<false>{const_cast<float*>(static_cast<const float*>(hi)), const_cast<float*>(static_cast<const float*>(lo))};
}
template <>
inline Pl<true> make_pl<true>(const void* hi, const void* lo, const int32_t* rec) {
  return Pl<true>{const_cast<uint16_t*>(static_cast<const uint16_t*>(hi)), const_cast<uint16_t*>(static_cast<const uint16_t*>(lo)),
                  const_cast<int32_t*>(rec), 1.f, 1.f};
}

// pixels per block: adaptive (vred_pix): enough blocks to fill the GPU, enough work per block to amortise the
// cross-row reduction + atomics

// per-channel reduction of NV quantities; f(p, c4, out[NV]) yields float4 per quantity for pixel p, channels c4*4..+3
// partials != nullptr: block b writes its NV*C partial sums to partials[b*NV*C ...] (combined by reduce_partials_kernel
// in block order: deterministic, no same-address atomics); else atomicAdd into out.
template <int NV, class F>
__device__ __forceinline__ void channel_reduce4(int64_t npix, int C, double* out, int out_stride, int pix_per_block,
                                                double* partials, F f) {
  const int q = C >> 2;                       // channel quads
  const int rows = blockDim.x / q;            // pixel rows handled concurrently by the block
  const int cq = threadIdx.x % q, pr = threadIdx.x / q;
  double acc[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0;
  int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  int64_t p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  if (pr < rows) {
    // 4 pixels in flight per thread (memory-level parallelism), fp32 partial over the 4, double across iterations
    int64_t p = p0 + pr;
    for (; p + 3 * (int64_t)rows < p1; p += 4 * (int64_t)rows) {
      float4 v0[NV], v1[NV], v2[NV], v3[NV];
      f(p, cq, v0);
      f(p + rows, cq, v1);
      f(p + 2 * (int64_t)rows, cq, v2);
      f(p + 3 * (int64_t)rows, cq, v3);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        acc[i][0] += (double)((v0[i].x + v1[i].x) + (v2[i].x + v3[i].x));
        acc[i][1] += (double)((v0[i].y + v1[i].y) + (v2[i].y + v3[i].y));
        acc[i][2] += (double)((v0[i].z + v1[i].z) + (v2[i].z + v3[i].z));
        acc[i][3] += (double)((v0[i].w + v1[i].w) + (v2[i].w + v3[i].w));
      }
    }
    for (; p < p1; p += rows) {
      float4 v[NV];
      f(p, cq, v);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        acc[i][0] += (double)v[i].x; acc[i][1] += (double)v[i].y;
        acc[i][2] += (double)v[i].z; acc[i][3] += (double)v[i].w;
      }
    }
  }
  extern __shared__ double vsm[];             // [NV*4][blockDim.x]
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) vsm[(i * 4 + j) * blockDim.x + threadIdx.x] = acc[i][j];
  __syncthreads();
  if (pr == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double sacc = 0.0;
        for (int r = 0; r < rows; ++r) sacc += vsm[(i * 4 + j) * blockDim.x + r * q + cq];
        if (partials) partials[((size_t)blockIdx.x * NV + i) * C + cq * 4 + j] = sacc;
        else atomicAdd(out + (size_t)i * out_stride + cq * 4 + j, sacc);
      }
  }
}

__global__ void bn_stats4_kernel(const float* __restrict__ y, int64_t npix, int C, int ycs, double* sums, int ppb,
                                 double* partials) {
  channel_reduce4<2>(npix, C, sums, C, ppb, partials, [&](int64_t p, int c4, float4* v) {
    float4 t = ld4(y + p * ycs + c4 * 4);
    v[0] = t;
    v[1] = make_float4(t.x * t.x, t.y * t.y, t.z * t.z, t.w * t.w);
  });
}

__device__ __forceinline__ float4 bn_act4(float4 y, float4 sc, float4 sh, int relu) {
  return make_float4(bn_act(y.x, sc.x, sh.x, relu), bn_act(y.y, sc.y, sh.y, relu), bn_act(y.z, sc.z, sh.z, relu),
                     bn_act(y.w, sc.w, sh.w, relu));
}

// qs = log2(C/4) when C/4 is a power of two (every layer of the shipped configs), else -1: the index split costs a shift
// instead of a 64-bit division per 16 bytes.  Four independent elements per iteration keep 64 bytes of loads in flight
// per thread.
template <bool H16>
__global__ void bn_apply4_kernel(const float* __restrict__ y, int64_t npix, int C, int ycs,
                                 const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                 Pl<H16> out, int ocs, int qs) {
  const int q = C >> 2;
  const int64_t total = npix * q, stride = (int64_t)gridDim.x * blockDim.x;
  out.init();
  float amax = 0.f;
  auto split = [&](int64_t i, int64_t& p, int& c) {
    if (qs >= 0) { p = i >> qs; c = (int)(i & (q - 1)) * 4; }
    else { p = i / q; c = (int)(i - p * q) * 4; }
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < total; i += 4 * stride) {
    int64_t p[4];
    int c[4];
    float4 yy[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      split(i + u * stride, p[u], c[u]);
      yy[u] = ld4(y + p[u] * ycs + c[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      out.store4((size_t)(p[u] * ocs + c[u]), bn_act4(yy[u], ld4(scale + c[u]), ld4(shift + c[u]), relu), amax);
  }
  for (; i < total; i += stride) {
    int64_t p;
    int c;
    split(i, p, c);
    out.store4((size_t)(p * ocs + c), bn_act4(ld4(y + p * ycs + c), ld4(scale + c), ld4(shift + c), relu), amax);
  }
  out.finish(amax);
}

// One thread = one INPUT pixel x 4 channels -> its 2x2 block of the x2 output (TF1 legacy bilinear: out[2i] = in[i],
// out[2i+1] = (in[i] + in[i+1]) / 2, last odd sample clamps): the four activated neighbours are computed once and
// serve four outputs (the per-output version evaluated 16 BN+ReLU and 4 loads per output quad).  Same expressions,
// same rounding as before.
template <bool H16>
__global__ void bn_apply_up2x4_kernel(const float* __restrict__ y, int N, int H, int W, int C, int ycs,
                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                      int relu, Pl<H16> out, int ocs) {
  const int Wo = 2 * W, q = C >> 2;
  const int64_t total = (int64_t)N * H * W * q;
  out.init();
  float amax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % q) * 4;
    int64_t p = i / q;
    const int w0 = (int)(p % W);
    int64_t t = p / W;
    const int h0 = (int)(t % H);
    const int n = (int)(t / H);
    const int h1 = min(h0 + 1, H - 1), w1 = min(w0 + 1, W - 1);
    const float4 sc = ld4(scale + c), sh = ld4(shift + c);
    const float* base = y + (int64_t)n * H * W * ycs + c;
    const float4 a00 = bn_act4(ld4(base + ((int64_t)h0 * W + w0) * ycs), sc, sh, relu);
    const float4 a01 = bn_act4(ld4(base + ((int64_t)h0 * W + w1) * ycs), sc, sh, relu);
    const float4 a10 = bn_act4(ld4(base + ((int64_t)h1 * W + w0) * ycs), sc, sh, relu);
    const float4 a11 = bn_act4(ld4(base + ((int64_t)h1 * W + w1) * ycs), sc, sh, relu);
    const int64_t orow = ((int64_t)n * 2 * H + 2 * h0) * Wo + 2 * w0;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const float fh = dh ? 0.5f : 0.f, fw = dw ? 0.5f : 0.f;
        float4 v;
#define IMMB_LERP(f)                                          \
        {                                                     \
          float top = a00.f + (a01.f - a00.f) * fw;           \
          float bot = a10.f + (a11.f - a10.f) * fw;           \
          v.f = top + (bot - top) * fh;                       \
        }
        IMMB_LERP(x) IMMB_LERP(y) IMMB_LERP(z) IMMB_LERP(w)
#undef IMMB_LERP
        out.store4((size_t)((orow + (int64_t)dh * Wo + dw) * ocs + c), v, amax);
      }
  }
  out.finish(amax);
}

__global__ void upsample2x_bwd4_kernel(const float* __restrict__ gup, int N, int H, int W, int C, int gcs,
                                       float* g) {
  const int Ho = 2 * H, Wo = 2 * W, q = C >> 2;
  int64_t total = (int64_t)N * H * W * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) * 4;
    int64_t p = i / q;
    int w = (int)(p % W);
    int64_t t = p / W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    const float* base = gup + (int64_t)n * Ho * Wo * gcs + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      int hh = 2 * h + dh;
      if (hh < 0) continue;
      float wh = dh == 0 ? 1.f : (dh < 0 ? 0.5f : (h == H - 1 ? 1.f : 0.5f));
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        int ww = 2 * w + dw;
        if (ww < 0) continue;
        float wt = wh * (dw == 0 ? 1.f : (dw < 0 ? 0.5f : (w == W - 1 ? 1.f : 0.5f)));
        float4 gv = ld4(base + ((int64_t)hh * Wo + ww) * gcs);
        acc.x += wt * gv.x; acc.y += wt * gv.y; acc.z += wt * gv.z; acc.w += wt * gv.w;
      }
    }
    st4(g + p * C + c, acc);
  }
}

struct BnBwdArgs {
  const float *g, *y, *scale, *shift, *mean, *invstd;
  int gcs, ycs, relu;
};

// per-thread constants of the BN backward: the thread's channel quad is fixed for its whole pixel loop
struct BnBwdRegs {
  float4 sc, sh, mu, is;
};

__device__ __forceinline__ void bn_bwd_elem4(const BnBwdArgs& a, const BnBwdRegs& r, int64_t p, int c, float4& dz,
                                             float4& xh) {
  float4 yy = ld4(a.y + p * a.ycs + c), gg = ld4(a.g + p * a.gcs + c);
#define IMMB_E(f)                                                        \
  {                                                                      \
    float z = fmaf(yy.f, r.sc.f, r.sh.f);                                \
    dz.f = (a.relu && !(z > 0.f)) ? 0.f : gg.f;                          \
    xh.f = (yy.f - r.mu.f) * r.is.f;                                     \
  }
  IMMB_E(x) IMMB_E(y) IMMB_E(z) IMMB_E(w)
#undef IMMB_E
}

__global__ void bn_bwd_reduce4_kernel(BnBwdArgs a, int64_t npix, int C, double* sums, int ppb, double* partials) {
  const int c = (threadIdx.x % (C >> 2)) * 4;
  const BnBwdRegs r{ld4(a.scale + c), ld4(a.shift + c), ld4(a.mean + c), ld4(a.invstd + c)};
  channel_reduce4<2>(npix, C, sums, C, ppb, partials, [&](int64_t p, int, float4* v) {
    float4 dz, xh;
    bn_bwd_elem4(a, r, p, c, dz, xh);
    v[0] = dz;
    v[1] = make_float4(dz.x * xh.x, dz.y * xh.y, dz.z * xh.z, dz.w * xh.w);
  });
}

template <bool H16>
__global__ void bn_bwd_apply4_kernel(BnBwdArgs a, int64_t npix, int C, const double* __restrict__ sums,
                                     Pl<H16> dyp, float* dgamma, float* dbeta, double* dbias_acc,
                                     int ppb, double* partials) {
  const double inv_n = 1.0 / (double)npix;
  dyp.init();
  float amax = 0.f;
  if (blockIdx.x == 0 && threadIdx.x < C) {
    dbeta[threadIdx.x] = (float)sums[threadIdx.x];
    dgamma[threadIdx.x] = (float)sums[C + threadIdx.x];
  }
  const int c = (threadIdx.x % (C >> 2)) * 4;
  const BnBwdRegs r{ld4(a.scale + c), ld4(a.shift + c), ld4(a.mean + c), ld4(a.invstd + c)};
  const float4 mdz = make_float4((float)(sums[c] * inv_n), (float)(sums[c + 1] * inv_n), (float)(sums[c + 2] * inv_n),
                                 (float)(sums[c + 3] * inv_n));
  const float4 mdzx = make_float4((float)(sums[C + c] * inv_n), (float)(sums[C + c + 1] * inv_n),
                                  (float)(sums[C + c + 2] * inv_n), (float)(sums[C + c + 3] * inv_n));
  channel_reduce4<1>(npix, C, dbias_acc, C, ppb, partials, [&](int64_t p, int, float4* v) {
    float4 dz, xh;
    bn_bwd_elem4(a, r, p, c, dz, xh);
    float4 dy;
    dy.x = r.sc.x * (dz.x - mdz.x - xh.x * mdzx.x);
    dy.y = r.sc.y * (dz.y - mdz.y - xh.y * mdzx.y);
    dy.z = r.sc.z * (dz.z - mdz.z - xh.z * mdzx.z);
    dy.w = r.sc.w * (dz.w - mdz.w - xh.w * mdzx.w);
    dyp.store4((size_t)(p * C + c), dy, amax);
    v[0] = dy;
  });
  __syncwarp();
  dyp.finish(amax);
}

// out[v] += sum_b partials[b*nvals + v] in a fixed (deterministic) order: 32 partial-block lanes per value (strided
// over b) with EIGHT independent loads in flight per thread (the kernel is a dependent-L2-load chain otherwise: 69
// launches per step used to cost 1.1 ms), then a fixed-order combine through shared memory.
// block = (32 values, 32 lanes).
// (also: out_f != nullptr writes the float value of the sum there instead of accumulating into `out`: the bias gradient
// needs no double accumulator and no cast kernel)
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double* __restrict__ partials, int nblocks, int nvals,
                                                                double* out, float* out_f) {
  __shared__ double sm[32][33];
  const int v = blockIdx.x * 32 + threadIdx.x;
  double s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.0;
  if (v < nvals) {
    const double* pv = partials + v;
    int b = threadIdx.y;
    for (; b + 7 * 32 < nblocks; b += 8 * 32) {
      double t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = __ldcg(pv + (size_t)(b + 32 * j) * nvals);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (b + 32 * j < nblocks) s[j] += __ldcg(pv + (size_t)(b + 32 * j) * nvals);
  }
  sm[threadIdx.y][threadIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (threadIdx.y == 0 && v < nvals) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
    for (int r = 0; r < 32; r += 4) {
      t0 += sm[r][threadIdx.x]; t1 += sm[r + 1][threadIdx.x]; t2 += sm[r + 2][threadIdx.x]; t3 += sm[r + 3][threadIdx.x];
    }
    if (out_f) out_f[v] = (float)((t0 + t1) + (t2 + t3));
    else out[v] += (t0 + t1) + (t2 + t3);
  }
}

// Many second-level reductions in one launch (the bias gradients of all BN layers, deferred to the end of the backward
// pass): item i owns blocks [block0, block0 + ceil(nvals / 32)); same summation order as reduce_partials_kernel; the
// float value of each sum is written to out_f.
__global__ void __launch_bounds__(1024) reduce_partials_multi_kernel(const immb_reduce_item* __restrict__ items, int n_items) {
  __shared__ double sm[32][33];
  int it = 0;
  while (it + 1 < n_items && (int)blockIdx.x >= items[it + 1].block0) ++it;
  const immb_reduce_item m = items[it];
  const int v = ((int)blockIdx.x - m.block0) * 32 + threadIdx.x;
  const int nblocks = m.nblocks, nvals = m.nvals;
  double s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.0;
  if (v < nvals) {
    const double* pv = m.partials + v;
    int b = threadIdx.y;
    for (; b + 7 * 32 < nblocks; b += 8 * 32) {
      double t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = __ldcg(pv + (size_t)(b + 32 * j) * nvals);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (b + 32 * j < nblocks) s[j] += __ldcg(pv + (size_t)(b + 32 * j) * nvals);
  }
  sm[threadIdx.y][threadIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (threadIdx.y == 0 && v < nvals) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
    for (int r = 0; r < 32; r += 4) {
      t0 += sm[r][threadIdx.x]; t1 += sm[r + 1][threadIdx.x]; t2 += sm[r + 2][threadIdx.x]; t3 += sm[r + 3][threadIdx.x];
    }
    m.out_f[v] = (float)((t0 + t1) + (t2 + t3));
  }
}

// Second level of the conv-epilogue BN statistics fused with bn_finalize: block = 32 channels x 32 row lanes sums the
// partial rows of sum(y) and sum(y^2) in the same fixed order as reduce_partials_kernel, then lane row 0 finalises its
// channel (one launch instead of two per BN layer; deterministic, no atomics).
__global__ void __launch_bounds__(1024) bn_finalize_partials_kernel(const double* __restrict__ partials, int nblocks, int C,
                                                                     double count, const float* gamma, const float* beta,
                                                                     float* mm, float* mv, float* scale, float* shift,
                                                                     float* mean_o, float* invstd_o) {
  __shared__ double sm[2][32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int nvals = 2 * C;
#pragma unroll
  for (int qn = 0; qn < 2; ++qn) {
    double s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.0;
    if (c < C) {
      const double* pv = partials + qn * C + c;
      int b = threadIdx.y;
      for (; b + 7 * 32 < nblocks; b += 8 * 32) {
        double t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = __ldcg(pv + (size_t)(b + 32 * j) * nvals);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += t[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (b + 32 * j < nblocks) s[j] += __ldcg(pv + (size_t)(b + 32 * j) * nvals);
    }
    sm[qn][threadIdx.y][threadIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  }
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double tot[2];
#pragma unroll
    for (int qn = 0; qn < 2; ++qn) {
      double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
      for (int r = 0; r < 32; r += 4) {
        t0 += sm[qn][r][threadIdx.x]; t1 += sm[qn][r + 1][threadIdx.x]; t2 += sm[qn][r + 2][threadIdx.x]; t3 += sm[qn][r + 3][threadIdx.x];
      }
      tot[qn] = (t0 + t1) + (t2 + t3);
    }
    bn_finalize_channel(c, 1, tot[0], tot[1], count, gamma, beta, mm, mv, scale, shift, mean_o, invstd_o);
  }
}

template <bool H16>
__global__ void bias_grad_kernel(Pl<H16> g, int gcs, int64_t npix, int C, double* acc) {
  g.init();
  channel_reduce<1>(npix, C, acc, C, [&](int64_t p, int c, float* v) {
    v[0] = g.load((size_t)(p * gcs + c));
  });
}

template <bool H16>
__global__ void split_planes4_kernel(const float* __restrict__ v, Pl<H16> out, int64_t n4) {
  out.init();
  float amax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    out.store4((size_t)i * 4, ld4(v + i * 4), amax);
  out.finish(amax);
}

template <bool HI, bool HO>
__global__ void maxpool2x2_fwd4_kernel(Pl<HI> x, int N, int H, int W, int C, Pl<HO> o) {
  int Ho = H / 2, Wo = W / 2, q = C >> 2;
  int64_t total = (int64_t)N * Ho * Wo * q;
  x.init();
  o.init();
  float amax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) * 4;
    int64_t p = i / q;
    int wo = (int)(p % Wo);
    int64_t t = p / Wo;
    int ho = (int)(t % Ho);
    int n = (int)(t / Ho);
    size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    float4 a = x.load4(base), b = x.load4(base + C);
    float4 d = x.load4(base + (size_t)W * C), e = x.load4(base + (size_t)W * C + C);
    float4 v = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                           fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
    o.store4((size_t)(p * C + c), v, amax);
  }
  o.finish(amax);
}

__global__ void maxpool2x2_bwd4_kernel(const float* __restrict__ g_out, const float* __restrict__ x_hi,
                                       const float* __restrict__ x_lo, int N, int H, int W, int C, float* g_in) {
  int Ho = H / 2, Wo = W / 2, q = C >> 2;
  int64_t total = (int64_t)N * Ho * Wo * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) * 4;
    int64_t p = i / q;
    int wo = (int)(p % Wo);
    int64_t t = p / Wo;
    int ho = (int)(t % Ho);
    int n = (int)(t / Ho);
    size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    size_t idx[4] = {base, base + C, base + (size_t)W * C, base + (size_t)W * C + C};
    float4 xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xv[k] = load_split4(x_hi, x_lo, idx[k]);
    float4 gg = ld4(g_out + p * C + c);
    float4 o[4];
#define IMMB_E(f)                                                  \
    {                                                              \
      int best = 0;                                                \
      float bv = xv[0].f;                                          \
      _Pragma("unroll") for (int k = 1; k < 4; ++k) if (xv[k].f > bv) { bv = xv[k].f; best = k; } \
      _Pragma("unroll") for (int k = 0; k < 4; ++k) o[k].f = (k == best) ? gg.f : 0.f;            \
    }
    IMMB_E(x) IMMB_E(y) IMMB_E(z) IMMB_E(w)
#undef IMMB_E
#pragma unroll
    for (int k = 0; k < 4; ++k) st4(g_in + idx[k], o[k]);
  }
}

__device__ __forceinline__ float mask_at(const float* __restrict__ mask, int64_t p, int h, int w, int R) {
  if (!mask) return 1.f;
  int s = R / h;
  int ww = (int)(p % w);
  int64_t t = p / w;
  int hh = (int)(t % h);
  int64_t b = t / h;
  return __ldg(mask + (b * R + (int64_t)hh * s) * R + (int64_t)ww * s);
}

template <bool H16>
__global__ void perceptual_level_sum4_kernel(Pl<H16> fg, int gcs, Pl<H16> fp, int pcs, int B, int h, int w, int C,
                                             const float* __restrict__ mask, int R, double* acc) {
  const int q = C >> 2;
  int64_t total = (int64_t)B * h * w * q;
  double local = 0.0;
  fg.init();
  fp.init();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) * 4;
    int64_t p = i / q;
    float4 a = fg.load4((size_t)(p * gcs + c)), b = fp.load4((size_t)(p * pcs + c));
    float m = mask_at(mask, p, h, w, R);
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    local += (double)(m * (dx * dx)) + (double)(m * (dy * dy)) + (double)(m * (dz * dz)) + (double)(m * (dw * dw));
  }
  local = warp_sum(local);
  __shared__ double sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = local;
  __syncthreads();
  if (wid == 0) {
    double v = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(acc, v);
  }
}

// 2x2/2 max pool of BOTH halves of a perceptual level ([gt ; pred] stacked on the batch axis) fused with that level's
// masked squared-difference sum (imm_model.py:143-147): the pool reads every element of the level anyway, so the
// separate perceptual_level_sum pass (4 more planes) disappears.  acc[0] += sum mask * (f_gt - f_pred)^2.
template <bool HI, bool HO>
__global__ void maxpool2x2_fwd_levelsum4_kernel(Pl<HI> x, int B, int H, int W, int C, Pl<HO> o,
                                                const float* __restrict__ mask, int R, double* acc) {
  x.init();
  o.init();
  float amax = 0.f;
  const int Ho = H / 2, Wo = W / 2, q = C >> 2;
  const int64_t total = (int64_t)B * Ho * Wo * q;
  const size_t half_in = (size_t)B * H * W * C, half_out = (size_t)B * Ho * Wo * C;
  const int s = R / H;
  double local = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) * 4;
    int64_t p = i / q;
    int wo = (int)(p % Wo);
    int64_t t = p / Wo;
    int ho = (int)(t % Ho);
    int n = (int)(t / Ho);
    size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    const size_t off[4] = {0, (size_t)C, (size_t)W * C, (size_t)W * C + C};
    float4 vg = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), vp = vg;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 g = x.load4(base + off[k]);
      const float4 pr = x.load4(half_in + base + off[k]);
      vg = make_float4(fmaxf(vg.x, g.x), fmaxf(vg.y, g.y), fmaxf(vg.z, g.z), fmaxf(vg.w, g.w));
      vp = make_float4(fmaxf(vp.x, pr.x), fmaxf(vp.y, pr.y), fmaxf(vp.z, pr.z), fmaxf(vp.w, pr.w));
      const int hh = 2 * ho + (k >> 1), ww = 2 * wo + (k & 1);
      const float m = mask ? __ldg(mask + ((int64_t)n * R + (int64_t)hh * s) * R + (int64_t)ww * s) : 1.f;
      const float dx = g.x - pr.x, dy = g.y - pr.y, dz = g.z - pr.z, dw = g.w - pr.w;
      local += (double)(m * (dx * dx)) + (double)(m * (dy * dy)) + (double)(m * (dz * dz)) + (double)(m * (dw * dw));
    }
    o.store4((size_t)(p * C + c), vg, amax);
    o.store4(half_out + (size_t)(p * C + c), vp, amax);
  }
  o.finish(amax);
  local = warp_sum(local);
  __shared__ double sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = local;
  __syncthreads();
  if (wid == 0) {
    double v = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(acc, v);
  }
}

template <bool HI, bool HO>
__global__ void vgg_bwd_combine4_kernel(const float* __restrict__ g_next, Pl<HI> fgp, Pl<HI> fpp, int B, int h, int w, int C,
                                        const float* __restrict__ mask, int R, const float* __restrict__ coef,
                                        Pl<HO> dyp) {
  const int q = C >> 2;
  int64_t total = (int64_t)B * h * w * q;
  float cf = coef ? __ldg(coef) : 0.f;
  fgp.init();
  fpp.init();
  dyp.init();
  float amax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / q;
    size_t e = (size_t)i * 4;
    float4 fp = fpp.load4(e);
    float4 g = g_next ? ld4(g_next + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (coef) {
      float4 fg = fgp.load4(e);
      float m = cf * mask_at(mask, p, h, w, R);
      g.x += m * (fg.x - fp.x); g.y += m * (fg.y - fp.y); g.z += m * (fg.z - fp.z); g.w += m * (fg.w - fp.w);
    }
    g.x = fp.x > 0.f ? g.x : 0.f; g.y = fp.y > 0.f ? g.y : 0.f; g.z = fp.z > 0.f ? g.z : 0.f; g.w = fp.w > 0.f ? g.w : 0.f;
    dyp.store4(e, g, amax);
  }
  dyp.finish(amax);
}

// max-pool backward fused with the loss-term / ReLU-backward combine of the layer that fed the pool (the pred half of
// a VGG activation): dy = split( [x > 0] * ( [x is the window's first max] * g_out  +  coef * mask * (f_gt - x) ) ),
// i.e. maxpool2x2_bwd4_kernel followed by vgg_bwd_combine4_kernel without the fp32 gradient round trip and without
// re-reading x.  coef == nullptr: no loss term at this layer (fg_* unused).
template <bool HI, bool HO>
__global__ void maxpool2x2_bwd_combine4_kernel(const float* __restrict__ g_out, Pl<HI> fgp, Pl<HI> xp, int N, int H, int W,
                                               int C, const float* __restrict__ mask, int R,
                                               const float* __restrict__ coef, Pl<HO> dyp) {
  fgp.init();
  xp.init();
  dyp.init();
  float amax = 0.f;
  const int Ho = H / 2, Wo = W / 2, q = C >> 2;
  const int64_t total = (int64_t)N * Ho * Wo * q;
  const float cf = coef ? __ldg(coef) : 0.f;
  const int s = R / H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) * 4;
    int64_t p = i / q;
    int wo = (int)(p % Wo);
    int64_t t = p / Wo;
    int ho = (int)(t % Ho);
    int n = (int)(t / Ho);
    size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    size_t idx[4] = {base, base + C, base + (size_t)W * C, base + (size_t)W * C + C};
    float4 xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xv[k] = xp.load4(idx[k]);
    float4 gg = ld4(g_out + p * C + c);
    float4 o[4];
#define IMMB_E(f)                                                  \
    {                                                              \
      int best = 0;                                                \
      float bv = xv[0].f;                                          \
      _Pragma("unroll") for (int k = 1; k < 4; ++k) if (xv[k].f > bv) { bv = xv[k].f; best = k; } \
      _Pragma("unroll") for (int k = 0; k < 4; ++k) o[k].f = (k == best) ? gg.f : 0.f;            \
    }
    IMMB_E(x) IMMB_E(y) IMMB_E(z) IMMB_E(w)
#undef IMMB_E
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 g = o[k];
      if (coef) {
        const float4 fg = fgp.load4(idx[k]);
        const int hh = 2 * ho + (k >> 1), ww = 2 * wo + (k & 1);
        const float m = cf * (mask ? __ldg(mask + ((int64_t)n * R + (int64_t)hh * s) * R + (int64_t)ww * s) : 1.f);
        g.x += m * (fg.x - xv[k].x); g.y += m * (fg.y - xv[k].y); g.z += m * (fg.z - xv[k].z); g.w += m * (fg.w - xv[k].w);
      }
      g.x = xv[k].x > 0.f ? g.x : 0.f; g.y = xv[k].y > 0.f ? g.y : 0.f;
      g.z = xv[k].z > 0.f ? g.z : 0.f; g.w = xv[k].w > 0.f ? g.w : 0.f;
      dyp.store4(idx[k], g, amax);
    }
  }
  dyp.finish(amax);
}

// =====================================================================================================
// landmark bottleneck: one warp per (b, k)
// =====================================================================================================
__device__ __forceinline__ float lin_coord(int i, int S) {
  // tf.linspace(-1, 1, S): start + i * (stop - start) / (S - 1)
  return S > 1 ? -1.0f + (float)i * (2.0f / (float)(S - 1)) : -1.0f;
}

template <bool H16>
__global__ void softargmax_gauss_fwd_kernel(const float* __restrict__ heat, int B, int S, int K, int hcs,
                                            float inv_std, float* mu, float* py, float* px, int Sg,
                                            Pl<H16> maps, int ocs, int c_off) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B * K) return;                 // whole warps leave together
  maps.init();
  float amax = 0.f;
  int b = warp / K, k = warp - b * K;
  const float* hb = heat + (int64_t)b * S * S * hcs + k;
  float ly = 0.f, lx = 0.f;
  if (lane < S) {
    for (int j = 0; j < S; ++j) {
      ly += __ldg(hb + ((int64_t)lane * S + j) * hcs);   // mean over w for row `lane`
      lx += __ldg(hb + ((int64_t)j * S + lane) * hcs);   // mean over h for col `lane`
    }
    ly /= (float)S;
    lx /= (float)S;
  } else {
    ly = -INFINITY;
    lx = -INFINITY;
  }
  float my = ly, mx = lx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  float ey = lane < S ? expf(ly - my) : 0.f;
  float ex = lane < S ? expf(lx - mx) : 0.f;
  float sy = warp_sum(ey), sx = warp_sum(ex);
  float pyv = ey / sy, pxv = ex / sx;
  float c = lin_coord(lane, S);
  float muy = warp_sum(lane < S ? pyv * c : 0.f);
  float mux = warp_sum(lane < S ? pxv * c : 0.f);
  if (lane < S) {
    py[((int64_t)b * S + lane) * K + k] = pyv;
    px[((int64_t)b * S + lane) * K + k] = pxv;
  }
  if (lane == 0) {
    mu[((int64_t)b * K + k) * 2 + 0] = muy;
    mu[((int64_t)b * K + k) * 2 + 1] = mux;
  }
  if (maps.hi) {
    float inv2 = inv_std * inv_std;
    for (int t = lane; t < Sg * Sg; t += 32) {
      int i = t / Sg, j = t - i * Sg;
      float dy = lin_coord(i, Sg) - muy, dx = lin_coord(j, Sg) - mux;
      float gval = expf(-(dy * dy + dx * dx) * inv2);
      maps.store((size_t)(((int64_t)b * Sg * Sg + t) * ocs + c_off + k), gval, amax);
    }
    __syncwarp();
    maps.finish(amax);
  }
}

__global__ void softargmax_gauss_bwd_kernel(const float* __restrict__ gmaps, int gcs, int c_off,
                                            const float* __restrict__ mu, const float* __restrict__ py,
                                            const float* __restrict__ px, int B, int S, int K, int Sg,
                                            float inv_std, float* gheat, int ghcs) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B * K) return;
  int b = warp / K, k = warp - b * K;
  float muy = mu[((int64_t)b * K + k) * 2 + 0], mux = mu[((int64_t)b * K + k) * 2 + 1];
  float inv2 = inv_std * inv_std;
  float gy = 0.f, gx = 0.f;
  for (int t = lane; t < Sg * Sg; t += 32) {
    int i = t / Sg, j = t - i * Sg;
    float dy = lin_coord(i, Sg) - muy, dx = lin_coord(j, Sg) - mux;
    float gval = expf(-(dy * dy + dx * dx) * inv2);
    float gg = __ldg(gmaps + ((int64_t)b * Sg * Sg + t) * gcs + c_off + k) * gval * 2.0f * inv2;
    gy += gg * dy;
    gx += gg * dx;
  }
  gy = warp_sum(gy);
  gx = warp_sum(gx);
  // d mu_y / d ly[h] = py[h] (c[h] - mu_y);  ly[h] = mean_w x[h, w]
  float ay = 0.f, ax = 0.f;
  if (lane < S) {
    float c = lin_coord(lane, S);
    ay = py[((int64_t)b * S + lane) * K + k] * (c - muy) * gy / (float)S;
    ax = px[((int64_t)b * S + lane) * K + k] * (c - mux) * gx / (float)S;
  }
  for (int h = 0; h < S; ++h) {
    float ayh = __shfl_sync(0xffffffffu, ay, h);
    if (lane < S) gheat[((int64_t)b * S * S + (int64_t)h * S + lane) * ghcs + k] = ayh + ax;
  }
}

__global__ void gaussian_maps_kernel(const float* __restrict__ mu, int B, int K, int S, float inv_std,
                                     float* maps) {
  int64_t total = (int64_t)B * S * S * K;
  float inv2 = inv_std * inv_std;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    int64_t p = i / K;
    int j = (int)(p % S);
    int64_t q = p / S;
    int ii = (int)(q % S);
    int b = (int)(q / S);
    float dy = lin_coord(ii, S) - mu[((int64_t)b * K + k) * 2 + 0];
    float dx = lin_coord(j, S) - mu[((int64_t)b * K + k) * 2 + 1];
    maps[i] = expf(-(dy * dy + dx * dx) * inv2);
  }
}

// =====================================================================================================
// perceptual-loss glue
// =====================================================================================================
__device__ __forceinline__ float gray_norm(const float* __restrict__ p) {
  float v = (__ldg(p) + __ldg(p + 1) + __ldg(p + 2)) / 3.0f;
  v = v / 255.0f;
  return v - (114.451f / 255.0f);
}

__global__ void vgg_prologue_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int pcs,
                                    int B, int R, float* out_hi, float* out_lo) {
  int64_t per = (int64_t)B * R * R;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = i < per ? gray_norm(gt + i * 3) : gray_norm(pred + (i - per) * pcs);
    store_split(out_hi, out_lo, (size_t)i, v);
  }
}

// 3x3 SAME patches of the grayscale image: out[n,h,w,r*3+s] = gray[n,h+r-1,w+s-1] (0 outside), channels 9..11 = 0
__global__ void vgg_prologue_patch_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int pcs,
                                          int B, int R, float* out_hi, float* out_lo) {
  int64_t per = (int64_t)B * R * R;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per * 12;
       i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % 12);
    int64_t p = i / 12;
    float v = 0.f;
    if (t < 9) {
      int w = (int)(p % R);
      int64_t q = p / R;
      int h = (int)(q % R);
      int64_t n = q / R;
      int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
      if (hh >= 0 && hh < R && ww >= 0 && ww < R) {
        int64_t src = (n * R + hh) * R + ww;
        v = src < per ? gray_norm(gt + src * 3) : gray_norm(pred + (src - per) * pcs);
      }
    }
    store_split(out_hi, out_lo, (size_t)i, v);
  }
}

// VGG conv1_1 (Cin = 1) fused with the gray/normalise prologue.  block = 256 threads = 64 pixels x 4 channel groups;
// each thread produces Cout/4 (= 16) channels of one pixel from its 9 gray taps; weights live in shared memory.
template <int COUT, bool H16>
__global__ void __launch_bounds__(256) vgg_conv1_1_fused_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                                int pcs, int B, int R, const float* __restrict__ w,
                                                                const float* __restrict__ bias, Pl<H16> out,
                                                                int64_t p_begin, int64_t p_end) {
  constexpr int CG = COUT / 4;
  out.init();
  float amax = 0.f;
  __shared__ float ws[9][COUT];
  __shared__ float bs[COUT];
  for (int i = threadIdx.x; i < 9 * COUT; i += 256) ws[i / COUT][i % COUT] = __ldg(w + i);
  for (int i = threadIdx.x; i < COUT; i += 256) bs[i] = __ldg(bias + i);
  __syncthreads();
  const int64_t per = (int64_t)B * R * R;
  const int cg = threadIdx.x & 3;                  // channel group
  for (int64_t p = p_begin + (int64_t)blockIdx.x * 64 + (threadIdx.x >> 2); p < p_end; p += (int64_t)gridDim.x * 64) {
    int wq = (int)(p % R);
    int64_t q = p / R;
    int h = (int)(q % R);
    int64_t n = q / R;
    float g[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      int hh = h + t / 3 - 1, ww = wq + t % 3 - 1;
      float v = 0.f;
      if (hh >= 0 && hh < R && ww >= 0 && ww < R) {
        int64_t src = (n * R + hh) * R + ww;
        v = src < per ? gray_norm(gt + src * 3) : gray_norm(pred + (src - per) * pcs);
      }
      g[t] = v;
    }
    float acc[CG];
#pragma unroll
    for (int j = 0; j < CG; ++j) acc[j] = bs[cg * CG + j];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < CG; ++j) acc[j] = fmaf(g[t], ws[t][cg * CG + j], acc[j]);
    size_t o = (size_t)p * COUT + cg * CG;
    // 256-bit stores: a thread owns 16 consecutive channels (64 B per plane); 128-bit stores would fill only half of
    // each 32-byte sector per instruction (this layer is pure HBM write traffic: 2 planes x 64 ch x 4 B per pixel)
    if constexpr (H16) {
      // 16 consecutive channels = 32 bytes per plane: one full sector per store
      static_assert(CG == 16, "fp16 planes: a thread owns 16 channels");
      uint32_t hw[8], lw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float v0 = fmaxf(acc[2 * u], 0.f), v1 = fmaxf(acc[2 * u + 1], 0.f);
        uint16_t h0, l0, h1, l1;
        amax = fmaxf(amax, fmaxf(v0, v1));
        split_h16(v0 * out.mul, h0, l0);
        split_h16(v1 * out.mul, h1, l1);
        hw[u] = pack2(h0, h1);
        lw[u] = pack2(l0, l1);
      }
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out.hi + o), "r"(hw[0]), "r"(hw[1]),
                   "r"(hw[2]), "r"(hw[3]), "r"(hw[4]), "r"(hw[5]), "r"(hw[6]), "r"(hw[7]) : "memory");
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out.lo + o), "r"(lw[0]), "r"(lw[1]),
                   "r"(lw[2]), "r"(lw[3]), "r"(lw[4]), "r"(lw[5]), "r"(lw[6]), "r"(lw[7]) : "memory");
    } else {
#pragma unroll
    for (int j = 0; j < CG; j += 8) {
      float h[8], l[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) split_tf32(fmaxf(acc[j + u], 0.f), h[u], l[u]);
      asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out.hi + o + j), "f"(h[0]), "f"(h[1]),
                   "f"(h[2]), "f"(h[3]), "f"(h[4]), "f"(h[5]), "f"(h[6]), "f"(h[7]) : "memory");
      asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out.lo + o + j), "f"(l[0]), "f"(l[1]),
                   "f"(l[2]), "f"(l[3]), "f"(l[4]), "f"(l[5]), "f"(l[6]), "f"(l[7]) : "memory");
    }
    }
  }
  out.finish(amax);
}

// Tiled version (R % 32 == 0): one block = an 8 x 32-pixel tile of one image.  Phase 1 puts the normalised gray values
// of the 10 x 34 halo into shared memory (the per-thread version recomputed 9 taps x 3 loads for every 16 channels:
// ncu showed 935 instructions per thread and the L1 pipe at 93 %).  Phase 2: thread = 16 channels x 4 consecutive pixels;
// the 3 x 6 gray window lives in registers, each filter tap's 16 weights are read once per 4 pixels from a bank-conflict-
// free layout (channel groups 20 floats apart).  ~300 instructions per 16 outputs.
template <int COUT, bool H16>
__global__ void __launch_bounds__(256) vgg_conv1_1_tiled_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                                int pcs, int B, int R, const float* __restrict__ w,
                                                                const float* __restrict__ bias, Pl<H16> out, int img_begin) {
  static_assert(COUT == 64, "4 channel groups of 16");
  constexpr int TH = 8, TW = 32, GW = TW + 2 + 2;             // gray row pitch 36 floats
  __shared__ float gs[TH + 2][GW];
  __shared__ __align__(16) float ws[9][4][20];
  __shared__ float bs[COUT];
  out.init();
  float amax = 0.f;
  for (int i = threadIdx.x; i < 9 * COUT; i += 256) ws[i / COUT][(i % COUT) >> 4][i & 15] = __ldg(w + i);
  for (int i = threadIdx.x; i < COUT; i += 256) bs[i] = __ldg(bias + i);
  const int tiles_w = R / TW, tiles_h = R / TH;
  const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h;
  const int64_t n = img_begin + blockIdx.x / (tiles_w * tiles_h);      // image index in [gt ; pred]
  const float* src = n < B ? gt + (size_t)n * R * R * 3 : pred + (size_t)(n - B) * R * R * pcs;
  const int sp = n < B ? 3 : pcs;
  for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += 256) {
    const int gy = i / (TW + 2), gx = i - gy * (TW + 2);
    const int hh = th * TH + gy - 1, ww = tw * TW + gx - 1;
    float v = 0.f;
    if (hh >= 0 && hh < R && ww >= 0 && ww < R) v = gray_norm(src + ((size_t)hh * R + ww) * sp);
    gs[gy][gx] = v;
  }
  __syncthreads();
  const int cg = threadIdx.x & 3, slot = threadIdx.x >> 2;
  const int row = slot >> 3, col0 = (slot & 7) * 4;
  float win[3][6];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) win[a][b] = gs[row + a][col0 + b];
  float acc[4][16];
#pragma unroll
  for (int px = 0; px < 4; ++px)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[px][j] = bs[cg * 16 + j];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float wv[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 q = *reinterpret_cast<const float4*>(&ws[t][cg][j]);
      wv[j] = q.x; wv[j + 1] = q.y; wv[j + 2] = q.z; wv[j + 3] = q.w;
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const float g = win[t / 3][px + t % 3];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[px][j] = fmaf(g, wv[j], acc[px][j]);
    }
  }
  const size_t pix0 = ((size_t)n * R + (th * TH + row)) * R + tw * TW + col0;
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const size_t o = (pix0 + px) * COUT + cg * 16;
    if constexpr (H16) {
      uint32_t hw[8], lw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float v0 = fmaxf(acc[px][2 * u], 0.f), v1 = fmaxf(acc[px][2 * u + 1], 0.f);
        uint16_t h0, l0, h1, l1;
        amax = fmaxf(amax, fmaxf(v0, v1));
        split_h16(v0 * out.mul, h0, l0);
        split_h16(v1 * out.mul, h1, l1);
        hw[u] = pack2(h0, h1);
        lw[u] = pack2(l0, l1);
      }
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out.hi + o), "r"(hw[0]), "r"(hw[1]),
                   "r"(hw[2]), "r"(hw[3]), "r"(hw[4]), "r"(hw[5]), "r"(hw[6]), "r"(hw[7]) : "memory");
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out.lo + o), "r"(lw[0]), "r"(lw[1]),
                   "r"(lw[2]), "r"(lw[3]), "r"(lw[4]), "r"(lw[5]), "r"(lw[6]), "r"(lw[7]) : "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        float h[8], l[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) split_tf32(fmaxf(acc[px][j + u], 0.f), h[u], l[u]);
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out.hi + o + j), "f"(h[0]), "f"(h[1]),
                     "f"(h[2]), "f"(h[3]), "f"(h[4]), "f"(h[5]), "f"(h[6]), "f"(h[7]) : "memory");
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out.lo + o + j), "f"(l[0]), "f"(l[1]),
                     "f"(l[2]), "f"(l[3]), "f"(l[4]), "f"(l[5]), "f"(l[6]), "f"(l[7]) : "memory");
      }
    }
  }
  out.finish(amax);
}

// first-layer staging: image [N,H,W,3] -> [N,H,W+8,4] (3 zero columns left, 5 right, 4th channel zero)
__global__ void stage_image_rowwin_kernel(const float* __restrict__ img, int N, int H, int W, float* x_hi,
                                          float* x_lo) {
  const int Wp = W + 8;
  int64_t total = (int64_t)N * H * Wp * 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i & 3);
    int64_t p = i >> 2;
    int wp = (int)(p % Wp);
    int64_t row = p / Wp;
    int w = wp - 3;
    float v = (c < 3 && w >= 0 && w < W) ? __ldg(img + (row * W + w) * 3 + c) : 0.f;
    store_split(x_hi, x_lo, (size_t)i, v);
  }
}

// [7,7,3,Cout] -> [7][Cout][32], k = s*4 + c
__global__ void pack_weights_rowwin_kernel(const float* __restrict__ w, int Cout, float* wp_hi, float* wp_lo) {
  int total = 7 * Cout * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int k = i & 31;
    int co = (i >> 5) % Cout;
    int r = (i >> 5) / Cout;
    int s = k >> 2, c = k & 3;
    float v = (s < 7 && c < 3) ? __ldg(w + ((size_t)(r * 7 + s) * 3 + c) * Cout + co) : 0.f;
    store_split(wp_hi, wp_lo, (size_t)i, v);
  }
}

__global__ void maxpool2x2_fwd_kernel(const float* __restrict__ x_hi, const float* __restrict__ x_lo, int N,
                                      int H, int W, int C, float* o_hi, float* o_lo) {
  int Ho = H / 2, Wo = W / 2;
  int64_t total = (int64_t)N * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int wo = (int)(p % Wo);
    int64_t q = p / Wo;
    int ho = (int)(q % Ho);
    int n = (int)(q / Ho);
    size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    float v = load_split(x_hi, x_lo, base);
    v = fmaxf(v, load_split(x_hi, x_lo, base + C));
    v = fmaxf(v, load_split(x_hi, x_lo, base + (size_t)W * C));
    v = fmaxf(v, load_split(x_hi, x_lo, base + (size_t)W * C + C));
    store_split(o_hi, o_lo, (size_t)i, v);
  }
}

__global__ void maxpool2x2_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ x_hi,
                                      const float* __restrict__ x_lo, int N, int H, int W, int C,
                                      float* g_in) {
  int Ho = H / 2, Wo = W / 2;
  int64_t total = (int64_t)N * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int wo = (int)(p % Wo);
    int64_t q = p / Wo;
    int ho = (int)(q % Ho);
    int n = (int)(q / Ho);
    size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    size_t idx[4] = {base, base + C, base + (size_t)W * C, base + (size_t)W * C + C};
    int best = 0;
    float bv = load_split(x_hi, x_lo, idx[0]);
#pragma unroll
    for (int t = 1; t < 4; ++t) {
      float v = load_split(x_hi, x_lo, idx[t]);
      if (v > bv) {
        bv = v;
        best = t;
      }
    }
    float gg = __ldg(g_out + i);
#pragma unroll
    for (int t = 0; t < 4; ++t) g_in[idx[t]] = (t == best) ? gg : 0.f;
  }
}

__global__ void perceptual_level_sum_kernel(const float* __restrict__ fg_hi, const float* __restrict__ fg_lo,
                                            int gcs, const float* __restrict__ fp_hi,
                                            const float* __restrict__ fp_lo, int pcs, int B, int h, int w, int C,
                                            const float* __restrict__ mask, int R, double* acc) {
  int64_t per = (int64_t)B * h * w;
  int64_t total = per * C;
  int s = R / h;
  double local = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    float d = load_split(fg_hi, fg_lo, (size_t)(p * gcs + c)) - load_split(fp_hi, fp_lo, (size_t)(p * pcs + c));
    float m = 1.f;
    if (mask) {
      int ww = (int)(p % w);
      int64_t q = p / w;
      int hh = (int)(q % h);
      int b = (int)(q / h);
      m = __ldg(mask + ((int64_t)b * R + (int64_t)hh * s) * R + (int64_t)ww * s);
    }
    local += (double)(m * (d * d));
  }
  local = warp_sum(local);
  __shared__ double sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = local;
  __syncthreads();
  if (wid == 0) {
    double v = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(acc, v);
  }
}

__global__ void perceptual_finalize_kernel(const double* acc, const double* counts, int n_levels, float* agg,
                                           int training, float* levels, float* rec_loss, float* coef) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  float total = 0.f;
  for (int k = 0; k < n_levels; ++k) {
    float s = (float)(acc[k] / counts[k]);
    float a = agg[k];
    float wl = a + (1.0f - 0.99f) * (s - a);
    float L = s / wl;
    levels[k] = L;
    total += L;
    coef[k] = (float)(1000.0 * 0.99 * (double)a / ((double)wl * (double)wl) * (-2.0 / counts[k]));
    if (training) agg[k] = wl;
  }
  rec_loss[0] = 1000.0f * total;
}

__global__ void vgg_bwd_combine_kernel(const float* __restrict__ g_next, const float* __restrict__ fg_hi,
                                       const float* __restrict__ fg_lo, const float* __restrict__ fp_hi,
                                       const float* __restrict__ fp_lo, int B, int h, int w, int C,
                                       const float* __restrict__ mask, int R, const float* __restrict__ coef,
                                       float* dy_hi, float* dy_lo) {
  int64_t per = (int64_t)B * h * w;
  int64_t total = per * C;
  int s = R / h;
  float cf = coef ? __ldg(coef) : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C;
    float fp = load_split(fp_hi, fp_lo, (size_t)i);
    float g = g_next ? __ldg(g_next + i) : 0.f;
    if (coef) {
      float fg = load_split(fg_hi, fg_lo, (size_t)i);
      float m = 1.f;
      if (mask) {
        int ww = (int)(p % w);
        int64_t q = p / w;
        int hh = (int)(q % h);
        int b = (int)(q / h);
        m = __ldg(mask + ((int64_t)b * R + (int64_t)hh * s) * R + (int64_t)ww * s);
      }
      g += cf * m * (fg - fp);
    }
    store_split(dy_hi, dy_lo, (size_t)i, fp > 0.f ? g : 0.f);
  }
}

__global__ void pred_grad_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int pcs,
                                 const float* __restrict__ mask, const float* __restrict__ coef_in,
                                 const float* __restrict__ g_vggin, int g_is_patch, int B, int R, float* g_hi,
                                 float* g_lo) {
  int64_t per = (int64_t)B * R * R;
  float cf = __ldg(coef_in);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per * pcs;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % pcs);
    int64_t p = i / pcs;
    float g = 0.f;
    if (c < 3) {
      float m = mask ? __ldg(mask + p) : 1.f;
      g = cf * m * (__ldg(gt + p * 3 + c) - __ldg(pred + i));
      if (g_vggin) {
        float gg;
        if (!g_is_patch) {
          gg = __ldg(g_vggin + p);
        } else {      // adjoint of the 3x3 patch extraction
          int w = (int)(p % R);
          int64_t q = p / R;
          int h = (int)(q % R);
          int64_t n = q / R;
          gg = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            int hh = h - (t / 3 - 1), ww = w - (t % 3 - 1);
            if (hh >= 0 && hh < R && ww >= 0 && ww < R) gg += __ldg(g_vggin + ((n * R + hh) * R + ww) * 12 + t);
          }
        }
        g += gg * (1.0f / 3.0f) * (1.0f / 255.0f);
      }
    }
    store_split(g_hi, g_lo, (size_t)i, g);
  }
}

// Backward of the whole VGG input stage in one kernel: dgrad of conv1_1 (3x3, Cin = 1) + adjoint of the gray /
// normalise prologue + the 'input' level's loss term = the gradient wrt the renderer output (split planes):
//   g_pred[c<3] = coef_in * mask * (gt_c - pred_c) + (1/(3*255)) * sum_t sum_co dy[q - off_t][co] * w[t][co]
// One block = one 16x16-pixel tile: phase 1 contracts the 64 channels of dy for the 18x18 halo pixels and the nine taps
// (4 threads per pixel, 16 channels each, exact fp32) into shared memory, phase 2 gathers the nine shifted values.
// Replaces conv2d_dgrad (1x1 over the [.,12] patch tensor on the tensor cores) + pred_grad and their 48 B/pixel tensor.
template <bool HI, bool HO>
__global__ void __launch_bounds__(256) vgg_conv1_1_bwd_fused_kernel(
    Pl<HI> dyp, const float* __restrict__ w,
    const float* __restrict__ gt, const float* __restrict__ pred, int pcs, const float* __restrict__ mask,
    const float* __restrict__ coef_in, int R, Pl<HO> gp) {
  constexpr int COUT = 64, T = 16, HW = T + 2, NP = HW * HW;
  // weights: channel groups 20 floats apart (the four 16-channel quarters of a pixel read different banks);
  // per-tap planes of the contracted halo (phase 2 reads consecutive addresses across a warp)
  __shared__ __align__(16) float ws[9][4][20];
  __shared__ float patch[9][NP + 4];
  dyp.init();
  gp.init();
  float amax = 0.f;
  for (int i = threadIdx.x; i < 9 * COUT; i += 256) ws[i / COUT][(i % COUT) >> 4][i & 15] = __ldg(w + i);
  const int tiles = R / T;
  const int tw = blockIdx.x % tiles, th = (blockIdx.x / tiles) % tiles;
  const int64_t n = blockIdx.x / (tiles * tiles);
  __syncthreads();
  const int part = threadIdx.x & 3;                       // 16-channel quarter of a pixel
  for (int it = 0; it < (NP + 63) / 64; ++it) {            // uniform trip count: the shuffles below need whole warps
    const int q = it * 64 + (threadIdx.x >> 2);
    const bool valid = q < NP;
    const int hh = th * T + q / HW - 1, ww = tw * T + q % HW - 1;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    if (valid && hh >= 0 && hh < R && ww >= 0 && ww < R) {
      const size_t base = (((size_t)n * R + hh) * R + ww) * COUT + part * 16;
      float v[16];
      if constexpr (HI) {
        // 16 channels = 32 bytes per plane: two 128-bit loads each
        const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(dyp.hi + base)), h1 = __ldg(reinterpret_cast<const uint4*>(dyp.hi + base + 8));
        const uint4 l0 = __ldg(reinterpret_cast<const uint4*>(dyp.lo + base)), l1 = __ldg(reinterpret_cast<const uint4*>(dyp.lo + base + 8));
        const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[2 * u] = join_h16((uint16_t)(hw[u] & 0xFFFFu), (uint16_t)(lw[u] & 0xFFFFu)) * dyp.inv;
          v[2 * u + 1] = join_h16((uint16_t)(hw[u] >> 16), (uint16_t)(lw[u] >> 16)) * dyp.inv;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 t4 = dyp.load4(base + j);
          v[j] = t4.x; v[j + 1] = t4.y; v[j + 2] = t4.z; v[j + 3] = t4.w;
        }
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(&ws[t][part][j]);
          acc[t] += v[j] * q4.x + v[j + 1] * q4.y + v[j + 2] * q4.z + v[j + 3] * q4.w;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 1);
      acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 2);
    }
    if (valid && part == 0) {
#pragma unroll
      for (int t = 0; t < 9; ++t) patch[t][q] = acc[t];
    }
  }
  __syncthreads();
  const int hl = threadIdx.x / T, wl = threadIdx.x % T;
  const int h = th * T + hl, wq = tw * T + wl;
  float gg = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) gg += patch[t][(hl + 1 - (t / 3 - 1)) * HW + (wl + 1 - (t % 3 - 1))];
  gg *= (1.0f / 3.0f) * (1.0f / 255.0f);
  const size_t p = ((size_t)n * R + h) * R + wq;
  const float cm = __ldg(coef_in) * (mask ? __ldg(mask + p) : 1.f);
  for (int c = 0; c < pcs; ++c) {
    float g = 0.f;
    if (c < 3) g = cm * (__ldg(gt + p * 3 + c) - __ldg(pred + p * pcs + c)) + gg;
    gp.store(p * pcs + c, g, amax);
  }
  gp.finish(amax);
}

__device__ __forceinline__ void ac_src(int o, int n_in, int n_out, int& lo, int& hi, float& f) {
  float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
  float s = (float)o * scale;
  lo = min((int)floorf(s), n_in - 1);
  hi = min(lo + 1, n_in - 1);
  f = s - (float)lo;
}

template <bool HI, bool HO>
__global__ void resize_ac_fwd_kernel(Pl<HI> x, int xcs, int N, int H, int W, int C, int Ho, int Wo, Pl<HO> o, int ocs) {
  int64_t total = (int64_t)N * Ho * Wo * C;
  x.init();
  o.init();
  float amax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int wo = (int)(p % Wo);
    int64_t q = p / Wo;
    int ho = (int)(q % Ho);
    int n = (int)(q / Ho);
    int h0, h1, w0, w1;
    float fh, fw;
    ac_src(ho, H, Ho, h0, h1, fh);
    ac_src(wo, W, Wo, w0, w1, fw);
    size_t nb = (size_t)n * H * W;
    float a00 = x.load((nb + (size_t)h0 * W + w0) * xcs + c);
    float a01 = x.load((nb + (size_t)h0 * W + w1) * xcs + c);
    float a10 = x.load((nb + (size_t)h1 * W + w0) * xcs + c);
    float a11 = x.load((nb + (size_t)h1 * W + w1) * xcs + c);
    float top = a00 + (a01 - a00) * fw, bot = a10 + (a11 - a10) * fw;
    o.store((size_t)(p * ocs + c), top + (bot - top) * fh, amax);
  }
  o.finish(amax);
}

// adjoint by scatter; g_in must be zeroed by the caller
__global__ void resize_ac_bwd_kernel(const float* __restrict__ g_out, int gcs, int N, int H, int W, int C,
                                     int Ho, int Wo, float* g_in) {
  int64_t total = (int64_t)N * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int wo = (int)(p % Wo);
    int64_t q = p / Wo;
    int ho = (int)(q % Ho);
    int n = (int)(q / Ho);
    int h0, h1, w0, w1;
    float fh, fw;
    ac_src(ho, H, Ho, h0, h1, fh);
    ac_src(wo, W, Wo, w0, w1, fw);
    float g = __ldg(g_out + p * gcs + c);
    size_t nb = (size_t)n * H * W;
    atomicAdd(g_in + (nb + (size_t)h0 * W + w0) * C + c, g * (1.f - fh) * (1.f - fw));
    atomicAdd(g_in + (nb + (size_t)h0 * W + w1) * C + c, g * (1.f - fh) * fw);
    atomicAdd(g_in + (nb + (size_t)h1 * W + w0) * C + c, g * fh * (1.f - fw));
    atomicAdd(g_in + (nb + (size_t)h1 * W + w1) * C + c, g * fh * fw);
  }
}

// =====================================================================================================
// optimiser: per-tensor clip_by_norm + TF Adam over flat buffers; one block per chunk
// =====================================================================================================
__global__ void adam_norms_kernel(const float* __restrict__ p, const float* __restrict__ g,
                                  const int32_t* __restrict__ chunk_tensor, const int64_t* __restrict__ chunk_off,
                                  const int32_t* __restrict__ chunk_len, const float* __restrict__ tensor_wd,
                                  float gscale, double* sq, double* wsq) {
  int ch = blockIdx.x;
  int t = chunk_tensor[ch];
  int64_t off = chunk_off[ch];
  int len = chunk_len[ch];
  float wd = tensor_wd[t];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    float pv = __ldg(p + off + i);
    float gv = fmaf(wd, pv, __ldg(g + off + i) * gscale);
    a += (double)gv * (double)gv;
    b += (double)pv * (double)pv;
  }
  a = warp_sum(a);
  b = warp_sum(b);
  __shared__ double sa[32], sb[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    sa[wid] = a;
    sb[wid] = b;
  }
  __syncthreads();
  if (wid == 0) {
    a = lane < (blockDim.x >> 5) ? sa[lane] : 0.0;
    b = lane < (blockDim.x >> 5) ? sb[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      atomicAdd(sq + t, a);
      atomicAdd(wsq + t, b);
    }
  }
}

__global__ void adam_apply_kernel(float* p, const float* __restrict__ g, float* m, float* v,
                                  const int32_t* __restrict__ chunk_tensor, const int64_t* __restrict__ chunk_off,
                                  const int32_t* __restrict__ chunk_len, const float* __restrict__ tensor_wd,
                                  float gscale, const double* __restrict__ sq, float clip, float lr_t,
                                  float beta1, float beta2, float eps, const float* __restrict__ lr_t_dev,
                                  float* amax) {
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);          // CUDA-graph replays: the step-dependent scalar lives in device memory
  float pmax = 0.f;
  int ch = blockIdx.x;
  int t = chunk_tensor[ch];
  int64_t off = chunk_off[ch];
  int len = chunk_len[ch];
  float wd = tensor_wd[t];
  // tf.clip_by_norm: g * clip / max(||g||, clip)
  float nrm = (float)sqrt(sq[t]);
  float cscale = clip > 0.f ? clip / fmaxf(nrm, clip) : 1.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    int64_t j = off + i;
    float pv = p[j];
    float gv = fmaf(wd, pv, __ldg(g + j) * gscale) * cscale;
    float mv = beta1 * m[j] + (1.f - beta1) * gv;
    float vv = beta2 * v[j] + (1.f - beta2) * gv * gv;
    m[j] = mv;
    v[j] = vv;
    const float pn = pv - lr_t * mv / (sqrtf(vv) + eps);
    p[j] = pn;
    pmax = fmaxf(pmax, fabsf(pn));
  }
  if (amax) {       // per-tensor max |p| of the UPDATED parameters (zeroed by the caller): scale of the fp16 weight planes
    __syncwarp();
    const uint32_t b = __reduce_max_sync(0xffffffffu, __float_as_uint(pmax));
    if ((threadIdx.x & 31) == 0 && b != 0) atomicMax(reinterpret_cast<unsigned int*>(amax + t), b);
  }
}

__global__ void total_loss_kernel(const float* rec_loss, const double* wsq, const float* tensor_wd,
                                  int n_tensors, float* weights_loss, float* total, const int32_t* overflow) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double s = 0.0;
  for (int t = 0; t < n_tensors; ++t) s += 0.5 * (double)tensor_wd[t] * wsq[t];
  weights_loss[0] = (float)s;
  total[0] = rec_loss[0] + (float)s;
  // an fp16 operand plane saturated (immb_scale_update counted it): the step is not trustworthy -> fail loudly through
  // the caller's NaN guard (cnn_train_multi.py:463) instead of training on clipped values
  if (overflow && overflow[0] > 0) total[0] = __int_as_float(0x7fc00000);
}

// HWIO master -> packed [tap][Cout][cin_pad] and split [tap][cin_pad][cout_pad], both as (hi, lo) planes
__global__ void pack_weights_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, int cin_pad,
                                    int cout_pad, float* wp_hi, float* wp_lo, float* wh_hi, float* wh_lo) {
  int64_t total = (int64_t)taps * cin_pad * cout_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int co = (int)(i % cout_pad);
    int64_t q = i / cout_pad;
    int ci = (int)(q % cin_pad);
    int tap = (int)(q / cin_pad);
    float val = (ci < Cin && co < Cout) ? __ldg(w + ((int64_t)tap * Cin + ci) * Cout + co) : 0.f;
    float hi, lo;
    split_tf32(val, hi, lo);
    if (wh_hi) {
      wh_hi[i] = hi;
      if (wh_lo) wh_lo[i] = lo;
    }
    if (wp_hi && co < Cout) {
      size_t j = ((size_t)tap * Cout + co) * cin_pad + ci;
      wp_hi[j] = hi;
      if (wp_lo) wp_lo[j] = lo;
    }
  }
}

// H16 variant: the same two layouts as scaled fp16 pairs.  The tensor's scale exponent is derived here from its largest
// magnitude (amax[0], produced by the optimiser / immb_multi_amax in the same step: weights need no delayed scaling)
// and published in rec[0] for the convolutions that consume the planes.
__device__ __forceinline__ int h16_exp_for(float amax, int target = kH16TargetExp) {
  if (!(amax > 0.f) || !isfinite(amax)) return 0;
  int ex;
  frexpf(amax, &ex);                    // amax = m * 2^ex, m in [0.5, 1)  ->  amax * 2^(target - ex) < 2^target
  int e = target - ex;
  return e > 100 ? 100 : (e < -100 ? -100 : e);
}
__global__ void pack_weights_h16_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, int cin_pad,
                                        int cout_pad, uint16_t* wp_hi, uint16_t* wp_lo, uint16_t* wh_hi, uint16_t* wh_lo,
                                        const float* __restrict__ amax, int32_t* rec) {
  const int e = h16_exp_for(__ldg(amax));
  const float mul = exp2i(e);
  if (blockIdx.x == 0 && threadIdx.x == 0) rec[0] = e;
  int64_t total = (int64_t)taps * cin_pad * cout_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int co = (int)(i % cout_pad);
    int64_t q = i / cout_pad;
    int ci = (int)(q % cin_pad);
    int tap = (int)(q / cin_pad);
    float val = (ci < Cin && co < Cout) ? __ldg(w + ((int64_t)tap * Cin + ci) * Cout + co) : 0.f;
    uint16_t hi, lo;
    split_h16(val * mul, hi, lo);
    if (wh_hi) {
      wh_hi[i] = hi;
      if (wh_lo) wh_lo[i] = lo;
    }
    if (wp_hi && co < Cout) {
      size_t j = ((size_t)tap * Cout + co) * cin_pad + ci;
      wp_hi[j] = hi;
      if (wp_lo) wp_lo[j] = lo;
    }
  }
}

// All weight tensors of the model in ONE launch (the 25 per-layer pack launches sat back to back on the critical path at
// the end of every step): `items` is a device table (immb_pack_item), block b works on the item whose [block0, block0 +
// nblocks) range contains it, 2048 elements per block.  kind 0 = fp32 TF32 planes, 1 = scaled fp16 planes, 2 = the 7x7
// first layer's row-window layout (TF32 planes).
__global__ void pack_weights_multi_kernel(const immb_pack_item* __restrict__ items, int n_items) {
  int it = 0;
  while (it + 1 < n_items && (int)blockIdx.x >= items[it + 1].block0) ++it;
  const immb_pack_item m = items[it];
  const int64_t base = (int64_t)((int)blockIdx.x - m.block0) * 2048;
  const float* w = m.w;
  if (m.kind == 2) {
    const int total = 7 * m.Cout * 32;
    float* wp_hi = (float*)m.wp_hi;
    float* wp_lo = (float*)m.wp_lo;
    for (int64_t i = base + threadIdx.x; i < base + 2048 && i < total; i += blockDim.x) {
      int k = (int)i & 31;
      int co = ((int)i >> 5) % m.Cout;
      int r = ((int)i >> 5) / m.Cout;
      int sx = k >> 2, c = k & 3;
      float v = (sx < 7 && c < 3) ? __ldg(w + ((size_t)(r * 7 + sx) * 3 + c) * m.Cout + co) : 0.f;
      store_split(wp_hi, wp_lo, (size_t)i, v);
    }
    return;
  }
  const int64_t total = (int64_t)m.taps * m.cin_pad * m.cout_pad;
  int e = 0;
  float mul = 1.f;
  if (m.kind == 1) {
    e = h16_exp_for(__ldg(m.amax));
    mul = exp2i(e);
    if ((int)blockIdx.x == m.block0 && threadIdx.x == 0) m.rec[0] = e;
  }
  for (int64_t i = base + threadIdx.x; i < base + 2048 && i < total; i += blockDim.x) {
    int co = (int)(i % m.cout_pad);
    int64_t q = i / m.cout_pad;
    int ci = (int)(q % m.cin_pad);
    int tap = (int)(q / m.cin_pad);
    float val = (ci < m.Cin && co < m.Cout) ? __ldg(w + ((int64_t)tap * m.Cin + ci) * m.Cout + co) : 0.f;
    const size_t j = ((size_t)tap * m.Cout + co) * m.cin_pad + ci;
    if (m.kind == 1) {
      uint16_t hi, lo;
      split_h16(val * mul, hi, lo);
      if (m.wh_hi) { ((uint16_t*)m.wh_hi)[i] = hi; ((uint16_t*)m.wh_lo)[i] = lo; }
      if (m.wp_hi && co < m.Cout) { ((uint16_t*)m.wp_hi)[j] = hi; ((uint16_t*)m.wp_lo)[j] = lo; }
    } else {
      float hi, lo;
      split_tf32(val, hi, lo);
      if (m.wh_hi) { ((float*)m.wh_hi)[i] = hi; ((float*)m.wh_lo)[i] = lo; }
      if (m.wp_hi && co < m.Cout) { ((float*)m.wp_hi)[j] = hi; ((float*)m.wp_lo)[j] = lo; }
    }
  }
}

// per-tensor max |p| over the chunk table of the flat parameter buffer (amax[t] zeroed by the caller; bit patterns of
// non-negative floats order like unsigned integers)
__global__ void multi_amax_kernel(const float* __restrict__ p, const int32_t* __restrict__ chunk_tensor,
                                  const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ chunk_len,
                                  float* amax) {
  const int ch = blockIdx.x;
  const int64_t off = chunk_off[ch];
  const int len = chunk_len[ch];
  float m = 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(p + off + i)));
  uint32_t b = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
  if ((threadIdx.x & 31) == 0 && b != 0) atomicMax(reinterpret_cast<unsigned int*>(amax + chunk_tensor[ch]), b);
}

// Delayed scaling of the activation / gradient planes: rec[i] = {e, amax bits}.  For every tensor that was written
// since the last update: count a saturation if its largest magnitude did not fit fp16 under the exponent it was
// written with, choose the next exponent from the observed maximum, clear the maximum.
__global__ void scale_update_kernel(int32_t* recs, int n, int32_t* overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t bits = (uint32_t)recs[2 * i + 1];
  if (bits == 0) return;
  const float a = __uint_as_float(bits);
  const int e_old = recs[2 * i];
  if (!isfinite(a) || a * exp2i(e_old) > 65504.f) atomicAdd(overflow, 1);
  if (isfinite(a)) recs[2 * i] = h16_exp_for(a, kH16DelayedTargetExp);
  recs[2 * i + 1] = 0;
}

// Thin-plate-spline warp: one thread per output pixel; the (Hc*Wc+3) x 2 parameters of the image are staged in
// shared memory; U(|g-c|^2) is evaluated on the fly (no [H*W, M+3] matrix in HBM).
__global__ void __launch_bounds__(256) tps_warp_kernel(const float* __restrict__ src, int H, int W, int C,
                                                       const float* __restrict__ w_tps, int Hc, int Wc, float* dst) {
  extern __shared__ float wsm[];                 // [(M+3)*2]
  const int b = blockIdx.y;
  const int M = Hc * Wc;
  for (int i = threadIdx.x; i < (M + 3) * 2; i += blockDim.x) wsm[i] = __ldg(w_tps + (size_t)b * (M + 3) * 2 + i);
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int ho = p / W, wo = p - ho * W;
  const float gx = W > 1 ? -1.f + 2.f * (float)wo / (float)(W - 1) : -1.f;     // np.linspace(-1, 1, W)
  const float gy = H > 1 ? -1.f + 2.f * (float)ho / (float)(H - 1) : -1.f;
  float ox = wsm[M * 2] + gx * wsm[(M + 1) * 2] + gy * wsm[(M + 2) * 2];
  float oy = wsm[M * 2 + 1] + gx * wsm[(M + 1) * 2 + 1] + gy * wsm[(M + 2) * 2 + 1];
  for (int j = 0; j < M; ++j) {
    const int cj = j % Wc, ci = j / Wc;
    const float cx = Wc > 1 ? -1.f + 2.f * (float)cj / (float)(Wc - 1) : -1.f;
    const float cy = Hc > 1 ? -1.f + 2.f * (float)ci / (float)(Hc - 1) : -1.f;
    float d2 = (gx - cx) * (gx - cx) + (gy - cy) * (gy - cy);
    d2 = fmaxf(d2, 1e-8f);
    const float k = logf(d2) * d2;
    ox = fmaf(k, wsm[j * 2], ox);
    oy = fmaf(k, wsm[j * 2 + 1], oy);
  }
  // grid_sample, bilinear, zero padding, corner-aligned: pixel = (coord + 1) / 2 * (size - 1)
  const float ix = (ox + 1.f) * 0.5f * (float)(W - 1), iy = (oy + 1.f) * 0.5f * (float)(H - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = ix - fx, ay = iy - fy;
  const float wts[4] = {(1.f - ax) * (1.f - ay), ax * (1.f - ay), (1.f - ax) * ay, ax * ay};
  const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
  const float* sb = src + (size_t)b * H * W * C;
  float* o = dst + ((size_t)b * H * W + p) * C;
  for (int c = 0; c < C; ++c) {
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (xs[t] >= 0 && xs[t] < W && ys[t] >= 0 && ys[t] < H) v += wts[t] * __ldg(sb + ((size_t)ys[t] * W + xs[t]) * C + c);
    o[c] = v;
  }
}

static inline int ew_grid(int64_t total, int block = 256) {
  int64_t g = (total + block - 1) / block;
  // 148 * 12 blocks = a whole number of waves at 2, 3, 4 or 6 resident blocks per SM (the grid-stride kernels hold
  // 45-112 registers; ncu showed 5.33 waves at the old cap of 148 * 16 with 3 blocks per SM: a third-full last wave)
  int64_t cap = (int64_t)kNumSMs * 12;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
static inline int log2_or_neg(int v) {       // log2(v) for a power of two, else -1
  if (v <= 0 || (v & (v - 1))) return -1;
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }
// vectorised per-channel kernels: C a multiple of 4 that divides 256*4 so that a 256-thread block covers whole rows
static inline bool vec_ok(int C, int cs) { return C % 4 == 0 && cs % 4 == 0 && C <= 1024 && (1024 % C) == 0; }
// pixels per block of the vectorised per-channel reductions.  With a scratch buffer (two-level reduction, no atomics)
// small tensors get ~4 blocks per SM; without it blocks stay large because every block ends in 2*C same-address atomics.
static inline int vred_pix(int64_t npix, int C, bool two_level) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("IMMB_VRED_PIX"); forced = e ? atoi(e) : 0; }
  if (forced > 0) return forced;
  if (!two_level) {
    int64_t ppb = npix / ((int64_t)kNumSMs * 8);
    ppb = (ppb + 255) / 256 * 256;
    if (ppb < 256) ppb = 256;
    if (ppb > 2048) ppb = 2048;
    return (int)ppb;
  }
  // ONE wave: these kernels hold 76-80 registers (3 blocks of 256 threads per SM), and a grid of 2.3 waves (ncu: 1024
  // blocks on the 128x128x32 tensors) leaves the last wave a third full -> at most 148 * 3 blocks, equal shares
  const int64_t unit = (int64_t)(1024 / C) * 4;          // pixel rows per block (256 threads / (C/4) quads) x 4 in flight
  int64_t ppb = (npix + (int64_t)kNumSMs * 3 - 1) / ((int64_t)kNumSMs * 3);
  ppb = (ppb + unit - 1) / unit * unit;
  if (ppb < unit) ppb = unit;
  return (int)ppb;
}
static inline int vred_grid(int64_t npix, int ppb) { return (int)((npix + ppb - 1) / ppb); }
static inline size_t vred_smem(int nv) { return sizeof(double) * nv * 4 * 256; }
static inline dim3 red_grid(int64_t npix, int C) {
  return dim3((unsigned)((npix + kRedPixPerBlock - 1) / kRedPixPerBlock), (unsigned)((C + 31) / 32));
}

}  // namespace immb

using namespace immb;
#define ST(s) ((cudaStream_t)(s))

extern "C" int immb_split_planes(const float* v, void* hi, void* lo, int64_t n, int32_t* scale, void* stream) {
  IMMB_REQUIRE(v && hi && n >= 0, "split_planes: bad args");
  if (n == 0) return IMMB_OK;
  if (scale) {
    IMMB_REQUIRE(lo && (n & 3) == 0 && aligned16(v) && aligned16(hi) && aligned16(lo), "split_planes: fp16 planes need n % 4 == 0");
    split_planes4_kernel<true><<<ew_grid(n / 4), 256, 0, ST(stream)>>>(v, make_pl<true>(hi, lo, scale), n / 4);
  } else if ((n & 3) == 0 && aligned16(v) && aligned16(hi) && aligned16(lo))
    split_planes4_kernel<false><<<ew_grid(n / 4), 256, 0, ST(stream)>>>(v, make_pl<false>(hi, lo, nullptr), n / 4);
  else
    split_planes_kernel<<<ew_grid(n), 256, 0, ST(stream)>>>(v, (float*)hi, (float*)lo, n);
  return check_launch("split_planes");
}

extern "C" size_t immb_bn_scratch_elems(int64_t npix, int C) {
  if (npix <= 0 || C <= 0 || !vec_ok(C, C)) return 0;
  int ppb = vred_pix(npix, C, true);
  return (size_t)vred_grid(npix, ppb) * 2 * (size_t)C;
}

// launches the second level of a two-level reduction
static int launch_reduce_partials(const double* partials, int nblocks, int nvals, double* out, cudaStream_t st,
                                  float* out_f = nullptr) {
  reduce_partials_kernel<<<ceil_div(nvals, 32), dim3(32, 32), 0, st>>>(partials, nblocks, nvals, out, out_f);
  return check_launch("reduce_partials");
}

extern "C" int immb_bn_stats(const float* y, int64_t npix, int C, int ycs, double* sums, double* scratch,
                             size_t scratch_elems, void* stream) {
  IMMB_REQUIRE(y && sums && npix > 0 && C > 0 && ycs >= C, "bn_stats: bad args");
  if (vec_ok(C, ycs) && aligned16(y)) {
    const bool two = scratch && scratch_elems >= immb_bn_scratch_elems(npix, C);
    const int ppb = vred_pix(npix, C, two), grid = vred_grid(npix, ppb);
    bn_stats4_kernel<<<grid, 256, vred_smem(2), ST(stream)>>>(y, npix, C, ycs, sums, ppb, two ? scratch : nullptr);
    int rc = check_launch("bn_stats");
    if (rc || !two) return rc;
    return launch_reduce_partials(scratch, grid, 2 * C, sums, ST(stream));
  }
  bn_stats_kernel<<<red_grid(npix, C), dim3(32, kRedRows), 0, ST(stream)>>>(y, npix, C, ycs, sums);
  return check_launch("bn_stats");
}

extern "C" int immb_bn_stats_from_partials(const double* partials, int rows, int C, double* sums, void* stream) {
  IMMB_REQUIRE(partials && sums && rows > 0 && C > 0, "bn_stats_from_partials: bad args");
  return launch_reduce_partials(partials, rows, 2 * C, sums, ST(stream));
}

extern "C" int immb_bn_finalize_partials(const double* partials, int rows, int64_t count, int C, const float* gamma,
                                         const float* beta, float* mm, float* mv, float* scale, float* shift,
                                         float* mean, float* invstd, void* stream) {
  IMMB_REQUIRE(partials && rows > 0 && gamma && beta && mm && mv && scale && shift && mean && invstd && C > 0,
               "bn_finalize_partials: bad args");
  bn_finalize_partials_kernel<<<ceil_div(C, 32), dim3(32, 32), 0, ST(stream)>>>(partials, rows, C, (double)count, gamma, beta,
                                                                                mm, mv, scale, shift, mean, invstd);
  return check_launch("bn_finalize_partials");
}

extern "C" int immb_bn_finalize(const double* sums, int64_t count, int C, const float* gamma,
                                const float* beta, float* mm, float* mv, int training, float* scale,
                                float* shift, float* mean, float* invstd, void* stream) {
  IMMB_REQUIRE(gamma && beta && mm && mv && scale && shift && mean && invstd && C > 0, "bn_finalize: bad args");
  IMMB_REQUIRE(!training || sums, "bn_finalize: training needs sums");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, ST(stream)>>>(sums, (double)count, C, gamma, beta, mm, mv,
                                                               training, scale, shift, mean, invstd);
  return check_launch("bn_finalize");
}

extern "C" int immb_bn_apply(const float* y, int N, int H, int W, int C, int ycs, const float* scale,
                             const float* shift, int relu, int up2x, void* out_hi, void* out_lo, int ocs,
                             int32_t* out_scale, void* stream) {
  IMMB_REQUIRE(y && scale && shift && out_hi && ycs >= C && ocs >= C, "bn_apply: bad args");
  int64_t npix = (int64_t)N * H * W;
  const bool v4 = vec_ok(C, ycs) && (ocs % 4 == 0) && aligned16(y) && aligned16(out_hi) && aligned16(out_lo);
  if (out_scale) {
    IMMB_REQUIRE(v4 && out_lo, "bn_apply: fp16 planes need the vectorised path (C % 4 == 0, aligned rows)");
    const Pl<true> o = make_pl<true>(out_hi, out_lo, out_scale);
    if (up2x)
      bn_apply_up2x4_kernel<true><<<ew_grid(npix * C / 4), 256, 0, ST(stream)>>>(y, N, H, W, C, ycs, scale, shift, relu, o, ocs);
    else
      bn_apply4_kernel<true><<<ew_grid(npix * C / 4 / 4), 256, 0, ST(stream)>>>(y, npix, C, ycs, scale, shift, relu, o, ocs, log2_or_neg(C / 4));
    return check_launch("bn_apply");
  }
  float *ohi = (float*)out_hi, *olo = (float*)out_lo;
  const Pl<false> o = make_pl<false>(out_hi, out_lo, nullptr);
  if (up2x) {
    if (v4)
      bn_apply_up2x4_kernel<false><<<ew_grid(npix * C / 4), 256, 0, ST(stream)>>>(y, N, H, W, C, ycs, scale, shift, relu, o, ocs);
    else
      bn_apply_up2x_kernel<<<ew_grid(npix * 4 * C), 256, 0, ST(stream)>>>(y, N, H, W, C, ycs, scale, shift,
                                                                           relu, ohi, olo, ocs);
  } else {
    if (v4)
      bn_apply4_kernel<false><<<ew_grid(npix * C / 4 / 4), 256, 0, ST(stream)>>>(y, npix, C, ycs, scale, shift, relu, o, ocs, log2_or_neg(C / 4));
    else
      bn_apply_kernel<<<ew_grid(npix * C), 256, 0, ST(stream)>>>(y, npix, C, ycs, scale, shift, relu, ohi, olo, ocs);
  }
  return check_launch("bn_apply");
}

extern "C" int immb_upsample2x_bwd(const float* g_up, int N, int H, int W, int C, int gcs, float* g,
                                   void* stream) {
  IMMB_REQUIRE(g_up && g && gcs >= C, "upsample2x_bwd: bad args");
  if (vec_ok(C, gcs) && aligned16(g_up) && aligned16(g))
    upsample2x_bwd4_kernel<<<ew_grid((int64_t)N * H * W * C / 4), 256, 0, ST(stream)>>>(g_up, N, H, W, C, gcs, g);
  else
    upsample2x_bwd_kernel<<<ew_grid((int64_t)N * H * W * C), 256, 0, ST(stream)>>>(g_up, N, H, W, C, gcs, g);
  return check_launch("upsample2x_bwd");
}

extern "C" int immb_bn_bwd_reduce(const float* g, int gcs, const float* y, int ycs, int64_t npix, int C,
                                  const float* scale, const float* shift, const float* mean,
                                  const float* invstd, int relu, double* sums, double* scratch,
                                  size_t scratch_elems, void* stream) {
  IMMB_REQUIRE(g && y && sums && gcs >= C && ycs >= C, "bn_bwd_reduce: bad args");
  if (vec_ok(C, gcs) && ycs % 4 == 0 && aligned16(g) && aligned16(y)) {
    BnBwdArgs a{g, y, scale, shift, mean, invstd, gcs, ycs, relu};
    const bool two = scratch && scratch_elems >= immb_bn_scratch_elems(npix, C);
    const int ppb = vred_pix(npix, C, two), grid = vred_grid(npix, ppb);
    bn_bwd_reduce4_kernel<<<grid, 256, vred_smem(2), ST(stream)>>>(a, npix, C, sums, ppb, two ? scratch : nullptr);
    int rc = check_launch("bn_bwd_reduce");
    if (rc || !two) return rc;
    return launch_reduce_partials(scratch, grid, 2 * C, sums, ST(stream));
  } else {
    bn_bwd_reduce_kernel<<<red_grid(npix, C), dim3(32, kRedRows), 0, ST(stream)>>>(
        g, gcs, y, ycs, npix, C, scale, shift, mean, invstd, relu, sums);
  }
  return check_launch("bn_bwd_reduce");
}

extern "C" int immb_bn_bwd_apply(const float* g, int gcs, const float* y, int ycs, int64_t npix, int C,
                                 const float* scale, const float* shift, const float* mean,
                                 const float* invstd, int relu, const double* sums, void* dy_hi,
                                 void* dy_lo, float* dgamma, float* dbeta, double* dbias_acc, double* scratch,
                                 size_t scratch_elems, int32_t* dy_scale, float* dbias_out, int defer_dbias,
                                 void* stream) {
  IMMB_REQUIRE(g && y && sums && dy_hi && dgamma && dbeta && dbias_acc, "bn_bwd_apply: bad args");
  const bool v4 = vec_ok(C, gcs) && ycs % 4 == 0 && C <= 256 && aligned16(g) && aligned16(y) && aligned16(dy_hi) &&
                  aligned16(dy_lo);
  IMMB_REQUIRE(!dy_scale || (v4 && dy_lo), "bn_bwd_apply: fp16 planes need the vectorised path");
  if (v4) {
    BnBwdArgs a{g, y, scale, shift, mean, invstd, gcs, ycs, relu};
    // two-level path: one row of C partial sums per block (immb_bn_bwd_apply_blocks rows)
    const bool two = scratch && scratch_elems >= (size_t)vred_grid(npix, vred_pix(npix, C, true)) * (size_t)C;
    const int ppb = vred_pix(npix, C, two), grid = vred_grid(npix, ppb);
    if (dy_scale)
      bn_bwd_apply4_kernel<true><<<grid, 256, vred_smem(1), ST(stream)>>>(a, npix, C, sums, make_pl<true>(dy_hi, dy_lo, dy_scale),
                                                                          dgamma, dbeta, dbias_acc, ppb, two ? scratch : nullptr);
    else
      bn_bwd_apply4_kernel<false><<<grid, 256, vred_smem(1), ST(stream)>>>(a, npix, C, sums, make_pl<false>(dy_hi, dy_lo, nullptr),
                                                                           dgamma, dbeta, dbias_acc, ppb, two ? scratch : nullptr);
    int rc = check_launch("bn_bwd_apply");
    if (rc) return rc;
    if (!two) {
      IMMB_REQUIRE(!dbias_out && !defer_dbias, "bn_bwd_apply: the float bias-gradient output needs the two-level (scratch) path");
      return rc;
    }
    if (defer_dbias) return rc;       // the caller reduces scratch[immb_bn_bwd_apply_blocks][C] later (immb_reduce_partials_multi)
    return launch_reduce_partials(scratch, grid, C, dbias_acc, ST(stream), dbias_out);
  } else {
    bn_bwd_apply_kernel<<<red_grid(npix, C), dim3(32, kRedRows), 0, ST(stream)>>>(
        g, gcs, y, ycs, npix, C, scale, shift, mean, invstd, relu, sums, (float*)dy_hi, (float*)dy_lo, dgamma, dbeta, dbias_acc);
  }
  return check_launch("bn_bwd_apply");
}

extern "C" int immb_bn_bwd_apply_blocks(int64_t npix, int C) {
  if (npix <= 0 || C <= 0 || !vec_ok(C, C) || C > 256) return 0;
  return vred_grid(npix, vred_pix(npix, C, true));
}

extern "C" int immb_reduce_partials_multi(const immb_reduce_item* items, int n_items, int total_blocks, void* stream) {
  IMMB_REQUIRE(items && n_items > 0 && total_blocks > 0, "reduce_partials_multi: bad args");
  reduce_partials_multi_kernel<<<total_blocks, dim3(32, 32), 0, ST(stream)>>>(items, n_items);
  return check_launch("reduce_partials_multi");
}

extern "C" int immb_bias_grad(const void* g_hi, const void* g_lo, int gcs, int64_t npix, int C, double* acc,
                              const int32_t* g_scale, void* stream) {
  IMMB_REQUIRE(g_hi && acc && gcs >= C, "bias_grad: bad args");
  if (g_scale)
    bias_grad_kernel<true><<<red_grid(npix, C), dim3(32, kRedRows), 0, ST(stream)>>>(make_pl<true>(g_hi, g_lo, g_scale), gcs, npix, C, acc);
  else
    bias_grad_kernel<false><<<red_grid(npix, C), dim3(32, kRedRows), 0, ST(stream)>>>(make_pl<false>(g_hi, g_lo, nullptr), gcs, npix, C, acc);
  return check_launch("bias_grad");
}

extern "C" int immb_cast_d2f(const double* src, float* dst, int64_t n, void* stream) {
  IMMB_REQUIRE(src && dst && n > 0, "cast_d2f: bad args");
  cast_d2f_kernel<<<ceil_div(n, 256), 256, 0, ST(stream)>>>(src, dst, n);
  return check_launch("cast_d2f");
}

extern "C" int immb_softargmax_gauss_fwd(const float* heat, int B, int S, int K, int hcs, float inv_std,
                                         float* mu, float* py, float* px, int Sg, void* maps_hi,
                                         void* maps_lo, int ocs, int c_off, int32_t* maps_scale, void* stream) {
  IMMB_REQUIRE(heat && mu && py && px && S >= 1 && S <= 32 && K >= 1 && hcs >= K, "softargmax_fwd: bad args (S<=32)");
  IMMB_REQUIRE(!maps_hi || (Sg >= 1 && ocs >= c_off + K), "softargmax_fwd: bad map args");
  IMMB_REQUIRE(!maps_scale || (maps_hi && maps_lo), "softargmax_fwd: fp16 map planes need both planes");
  int warps = B * K;
  if (maps_scale)
    softargmax_gauss_fwd_kernel<true><<<ceil_div((int64_t)warps * 32, 128), 128, 0, ST(stream)>>>(
        heat, B, S, K, hcs, inv_std, mu, py, px, Sg, make_pl<true>(maps_hi, maps_lo, maps_scale), ocs, c_off);
  else
    softargmax_gauss_fwd_kernel<false><<<ceil_div((int64_t)warps * 32, 128), 128, 0, ST(stream)>>>(
        heat, B, S, K, hcs, inv_std, mu, py, px, Sg, make_pl<false>(maps_hi, maps_lo, nullptr), ocs, c_off);
  return check_launch("softargmax_gauss_fwd");
}

extern "C" int immb_softargmax_gauss_bwd(const float* g_maps, int gcs, int c_off, const float* mu,
                                         const float* py, const float* px, int B, int S, int K, int Sg,
                                         float inv_std, float* g_heat, int ghcs, void* stream) {
  IMMB_REQUIRE(g_maps && mu && py && px && g_heat && S <= 32 && ghcs >= K, "softargmax_bwd: bad args");
  int warps = B * K;
  softargmax_gauss_bwd_kernel<<<ceil_div((int64_t)warps * 32, 128), 128, 0, ST(stream)>>>(
      g_maps, gcs, c_off, mu, py, px, B, S, K, Sg, inv_std, g_heat, ghcs);
  return check_launch("softargmax_gauss_bwd");
}

extern "C" int immb_gaussian_maps(const float* mu, int B, int K, int S, float inv_std, float* maps,
                                  void* stream) {
  IMMB_REQUIRE(mu && maps && B > 0 && K > 0 && S > 0, "gaussian_maps: bad args");
  gaussian_maps_kernel<<<ew_grid((int64_t)B * S * S * K), 256, 0, ST(stream)>>>(mu, B, K, S, inv_std, maps);
  return check_launch("gaussian_maps");
}

extern "C" int immb_vgg_prologue(const float* gt, const float* pred, int pcs, int B, int R, int patches,
                                 float* out_hi, float* out_lo, void* stream) {
  IMMB_REQUIRE(gt && pred && out_hi && pcs >= 3, "vgg_prologue: bad args");
  if (patches)
    vgg_prologue_patch_kernel<<<ew_grid((int64_t)2 * B * R * R * 12), 256, 0, ST(stream)>>>(gt, pred, pcs, B, R,
                                                                                         out_hi, out_lo);
  else
    vgg_prologue_kernel<<<ew_grid((int64_t)2 * B * R * R), 256, 0, ST(stream)>>>(gt, pred, pcs, B, R, out_hi,
                                                                                out_lo);
  return check_launch("vgg_prologue");
}

extern "C" int immb_vgg_conv1_1_fused(const float* gt, const float* pred, int pcs, int B, int R, const float* w,
                                      const float* bias, int Cout, void* out_hi, void* out_lo, int which,
                                      int32_t* out_scale, void* stream) {
  IMMB_REQUIRE(w && bias && out_hi && pcs >= 3 && which >= 0 && which <= 2, "vgg_conv1_1_fused: bad args");
  IMMB_REQUIRE((which == 2 || gt) && (which == 1 || pred), "vgg_conv1_1_fused: missing input for the requested half");
  IMMB_REQUIRE(Cout == 64, "vgg_conv1_1_fused: Cout must be 64 (VGG16 conv1_1)");
  IMMB_REQUIRE(out_lo && aligned32(out_hi) && aligned32(out_lo), "vgg_conv1_1_fused: split output planes must be 32-byte aligned");
  const int64_t per = (int64_t)B * R * R;
  const int64_t p0 = which == 2 ? per : 0, p1 = which == 1 ? per : 2 * per;
  static int tiled_on = -1;
  if (tiled_on < 0) { const char* ev = getenv("IMMB_CONV11_TILED"); tiled_on = (ev && atoi(ev) == 0) ? 0 : 1; }
  if (tiled_on && R % 32 == 0) {
    const int img0 = which == 2 ? B : 0, n_img = which == 0 ? 2 * B : B;
    const int grid = n_img * (R / 32) * (R / 8);
    if (out_scale)
      vgg_conv1_1_tiled_kernel<64, true><<<grid, 256, 0, ST(stream)>>>(gt, pred, pcs, B, R, w, bias,
                                                                       make_pl<true>(out_hi, out_lo, out_scale), img0);
    else
      vgg_conv1_1_tiled_kernel<64, false><<<grid, 256, 0, ST(stream)>>>(gt, pred, pcs, B, R, w, bias,
                                                                        make_pl<false>(out_hi, out_lo, nullptr), img0);
    return check_launch("vgg_conv1_1_fused");
  }
  int64_t blocks = (p1 - p0 + 63) / 64;
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  if (out_scale)
    vgg_conv1_1_fused_kernel<64, true><<<(int)blocks, 256, 0, ST(stream)>>>(gt, pred, pcs, B, R, w, bias,
                                                                            make_pl<true>(out_hi, out_lo, out_scale), p0, p1);
  else
    vgg_conv1_1_fused_kernel<64, false><<<(int)blocks, 256, 0, ST(stream)>>>(gt, pred, pcs, B, R, w, bias,
                                                                             make_pl<false>(out_hi, out_lo, nullptr), p0, p1);
  return check_launch("vgg_conv1_1_fused");
}

extern "C" int immb_stage_image_rowwin(const float* image, int N, int H, int W, float* x4_hi, float* x4_lo,
                                       void* stream) {
  IMMB_REQUIRE(image && x4_hi && N > 0 && H > 0 && W > 0, "stage_image_rowwin: bad args");
  stage_image_rowwin_kernel<<<ew_grid((int64_t)N * H * (W + 8) * 4), 256, 0, ST(stream)>>>(image, N, H, W, x4_hi,
                                                                                          x4_lo);
  return check_launch("stage_image_rowwin");
}

extern "C" int immb_pack_weights_rowwin(const float* w, int Cout, float* wp_hi, float* wp_lo, void* stream) {
  IMMB_REQUIRE(w && wp_hi && Cout > 0, "pack_weights_rowwin: bad args");
  pack_weights_rowwin_kernel<<<ew_grid((int64_t)7 * Cout * 32), 256, 0, ST(stream)>>>(w, Cout, wp_hi, wp_lo);
  return check_launch("pack_weights_rowwin");
}

template <bool H>
static inline Pl<H> pl_of(const void* hi, const void* lo, const int32_t* rec) { return make_pl<H>(hi, lo, rec); }

extern "C" int immb_maxpool2x2_fwd(const void* x_hi, const void* x_lo, int N, int H, int W, int C,
                                   void* o_hi, void* o_lo, const int32_t* x_scale, int32_t* o_scale, void* stream) {
  IMMB_REQUIRE(x_hi && o_hi && (H % 2 == 0) && (W % 2 == 0), "maxpool_fwd: bad args (even sizes only)");
  const bool v4 = C % 4 == 0 && aligned16(x_hi) && aligned16(x_lo) && aligned16(o_hi) && aligned16(o_lo);
  IMMB_REQUIRE((!x_scale && !o_scale) || (v4 && x_lo && o_lo), "maxpool_fwd: fp16 planes need 4 | C and both planes");
  const int grid = ew_grid((int64_t)N * (H / 2) * (W / 2) * C / 4);
  if (x_scale && o_scale)
    maxpool2x2_fwd4_kernel<true, true><<<grid, 256, 0, ST(stream)>>>(pl_of<true>(x_hi, x_lo, x_scale), N, H, W, C, pl_of<true>(o_hi, o_lo, o_scale));
  else if (x_scale)
    maxpool2x2_fwd4_kernel<true, false><<<grid, 256, 0, ST(stream)>>>(pl_of<true>(x_hi, x_lo, x_scale), N, H, W, C, pl_of<false>(o_hi, o_lo, nullptr));
  else if (o_scale)
    return set_error(IMMB_ERR_UNSUPPORTED, "maxpool_fwd: fp32 -> fp16 planes is not built");
  else if (v4)
    maxpool2x2_fwd4_kernel<false, false><<<grid, 256, 0, ST(stream)>>>(pl_of<false>(x_hi, x_lo, nullptr), N, H, W, C, pl_of<false>(o_hi, o_lo, nullptr));
  else
    maxpool2x2_fwd_kernel<<<ew_grid((int64_t)N * (H / 2) * (W / 2) * C), 256, 0, ST(stream)>>>(
        (const float*)x_hi, (const float*)x_lo, N, H, W, C, (float*)o_hi, (float*)o_lo);
  return check_launch("maxpool2x2_fwd");
}

extern "C" int immb_maxpool2x2_fwd_levelsum(const void* x_hi, const void* x_lo, int B, int H, int W, int C,
                                            void* o_hi, void* o_lo, const float* mask, int R, double* acc,
                                            const int32_t* x_scale, int32_t* o_scale, void* stream) {
  IMMB_REQUIRE(x_hi && x_lo && o_hi && o_lo && acc && B > 0 && (H % 2 == 0) && (W % 2 == 0) && C % 4 == 0,
               "maxpool_fwd_levelsum: bad args (even sizes, split planes, 4 | C)");
  IMMB_REQUIRE(!mask || (R >= H && R % H == 0), "maxpool_fwd_levelsum: mask resolution must be a multiple of the level's");
  IMMB_REQUIRE(aligned16(x_hi) && aligned16(x_lo) && aligned16(o_hi) && aligned16(o_lo), "maxpool_fwd_levelsum: alignment");
  const int grid = ew_grid((int64_t)B * (H / 2) * (W / 2) * C / 4);
  if (x_scale && o_scale)
    maxpool2x2_fwd_levelsum4_kernel<true, true><<<grid, 256, 0, ST(stream)>>>(pl_of<true>(x_hi, x_lo, x_scale), B, H, W, C,
                                                                              pl_of<true>(o_hi, o_lo, o_scale), mask, R, acc);
  else if (x_scale)
    maxpool2x2_fwd_levelsum4_kernel<true, false><<<grid, 256, 0, ST(stream)>>>(pl_of<true>(x_hi, x_lo, x_scale), B, H, W, C,
                                                                               pl_of<false>(o_hi, o_lo, nullptr), mask, R, acc);
  else if (o_scale)
    return set_error(IMMB_ERR_UNSUPPORTED, "maxpool_fwd_levelsum: fp32 -> fp16 planes is not built");
  else
    maxpool2x2_fwd_levelsum4_kernel<false, false><<<grid, 256, 0, ST(stream)>>>(pl_of<false>(x_hi, x_lo, nullptr), B, H, W, C,
                                                                                pl_of<false>(o_hi, o_lo, nullptr), mask, R, acc);
  return check_launch("maxpool2x2_fwd_levelsum");
}

extern "C" int immb_maxpool2x2_bwd(const float* g_out, const float* x_hi, const float* x_lo, int N, int H,
                                   int W, int C, float* g_in, void* stream) {
  IMMB_REQUIRE(g_out && x_hi && g_in && (H % 2 == 0) && (W % 2 == 0), "maxpool_bwd: bad args");
  if (C % 4 == 0 && aligned16(g_out) && aligned16(x_hi) && aligned16(x_lo) && aligned16(g_in))
    maxpool2x2_bwd4_kernel<<<ew_grid((int64_t)N * (H / 2) * (W / 2) * C / 4), 256, 0, ST(stream)>>>(g_out, x_hi, x_lo,
                                                                                                  N, H, W, C, g_in);
  else
    maxpool2x2_bwd_kernel<<<ew_grid((int64_t)N * (H / 2) * (W / 2) * C), 256, 0, ST(stream)>>>(g_out, x_hi, x_lo,
                                                                                              N, H, W, C, g_in);
  return check_launch("maxpool2x2_bwd");
}

extern "C" int immb_maxpool2x2_bwd_combine(const float* g_out, const void* fg_hi, const void* fg_lo,
                                           const void* fp_hi, const void* fp_lo, int B, int H, int W, int C,
                                           const float* mask, int R, const float* coef, void* dy_hi, void* dy_lo,
                                           const int32_t* f_scale, int32_t* dy_scale, void* stream) {
  IMMB_REQUIRE(g_out && fp_hi && fp_lo && dy_hi && dy_lo && (H % 2 == 0) && (W % 2 == 0) && C % 4 == 0 &&
               (!coef || (fg_hi && fg_lo)), "maxpool_bwd_combine: bad args");
  IMMB_REQUIRE(!mask || (R >= H && R % H == 0), "maxpool_bwd_combine: mask resolution must be a multiple of the level's");
  IMMB_REQUIRE(aligned16(g_out) && aligned16(fg_hi) && aligned16(fg_lo) && aligned16(fp_hi) && aligned16(fp_lo) &&
               aligned16(dy_hi) && aligned16(dy_lo), "maxpool_bwd_combine: alignment");
  IMMB_REQUIRE((f_scale != nullptr) == (dy_scale != nullptr), "maxpool_bwd_combine: activation and dy planes share a format");
  const int grid = ew_grid((int64_t)B * (H / 2) * (W / 2) * C / 4);
  if (f_scale)
    maxpool2x2_bwd_combine4_kernel<true, true><<<grid, 256, 0, ST(stream)>>>(
        g_out, pl_of<true>(fg_hi, fg_lo, f_scale), pl_of<true>(fp_hi, fp_lo, f_scale), B, H, W, C, mask, R, coef,
        pl_of<true>(dy_hi, dy_lo, dy_scale));
  else
    maxpool2x2_bwd_combine4_kernel<false, false><<<grid, 256, 0, ST(stream)>>>(
        g_out, pl_of<false>(fg_hi, fg_lo, nullptr), pl_of<false>(fp_hi, fp_lo, nullptr), B, H, W, C, mask, R, coef,
        pl_of<false>(dy_hi, dy_lo, nullptr));
  return check_launch("maxpool2x2_bwd_combine");
}

extern "C" int immb_perceptual_level_sum(const void* fg_hi, const void* fg_lo, int gcs, const void* fp_hi,
                                         const void* fp_lo, int pcs, int B, int h, int w, int C,
                                         const float* mask, int R, double* acc, const int32_t* f_scale, void* stream) {
  IMMB_REQUIRE(fg_hi && fp_hi && acc && gcs >= C && pcs >= C && h > 0 && (!mask || R % h == 0),
               "perceptual_level_sum: bad args");
  const bool v4 = C % 4 == 0 && gcs % 4 == 0 && pcs % 4 == 0 && aligned16(fg_hi) && aligned16(fg_lo) && aligned16(fp_hi) &&
                  aligned16(fp_lo);
  IMMB_REQUIRE(!f_scale || (v4 && fg_lo && fp_lo), "perceptual_level_sum: fp16 planes need 4 | C and both planes");
  if (f_scale)
    perceptual_level_sum4_kernel<true><<<ew_grid((int64_t)B * h * w * C / 4), 256, 0, ST(stream)>>>(
        pl_of<true>(fg_hi, fg_lo, f_scale), gcs, pl_of<true>(fp_hi, fp_lo, f_scale), pcs, B, h, w, C, mask, R, acc);
  else if (v4)
    perceptual_level_sum4_kernel<false><<<ew_grid((int64_t)B * h * w * C / 4), 256, 0, ST(stream)>>>(
        pl_of<false>(fg_hi, fg_lo, nullptr), gcs, pl_of<false>(fp_hi, fp_lo, nullptr), pcs, B, h, w, C, mask, R, acc);
  else
    perceptual_level_sum_kernel<<<ew_grid((int64_t)B * h * w * C), 256, 0, ST(stream)>>>(
        (const float*)fg_hi, (const float*)fg_lo, gcs, (const float*)fp_hi, (const float*)fp_lo, pcs, B, h, w, C, mask, R, acc);
  return check_launch("perceptual_level_sum");
}

extern "C" int immb_perceptual_finalize(const double* acc, const double* counts, int n_levels, float* agg,
                                        int training, float* levels, float* rec_loss, float* coef,
                                        void* stream) {
  IMMB_REQUIRE(acc && counts && agg && levels && rec_loss && coef && n_levels > 0, "perceptual_finalize: bad args");
  perceptual_finalize_kernel<<<1, 32, 0, ST(stream)>>>(acc, counts, n_levels, agg, training, levels, rec_loss,
                                                       coef);
  return check_launch("perceptual_finalize");
}

extern "C" int immb_vgg_bwd_combine(const float* g_next, const void* fg_hi, const void* fg_lo,
                                    const void* fp_hi, const void* fp_lo, int B, int h, int w, int C,
                                    const float* mask, int R, const float* coef, void* dy_hi, void* dy_lo,
                                    const int32_t* f_scale, int32_t* dy_scale, void* stream) {
  IMMB_REQUIRE(fp_hi && dy_hi && (g_next || coef) && (!coef || fg_hi), "vgg_bwd_combine: bad args");
  const bool v4 = C % 4 == 0 && aligned16(g_next) && aligned16(fg_hi) && aligned16(fg_lo) && aligned16(fp_hi) &&
                  aligned16(fp_lo) && aligned16(dy_hi) && aligned16(dy_lo);
  IMMB_REQUIRE((f_scale != nullptr) == (dy_scale != nullptr), "vgg_bwd_combine: activation and dy planes share a format");
  IMMB_REQUIRE(!f_scale || (v4 && fp_lo && dy_lo), "vgg_bwd_combine: fp16 planes need 4 | C and both planes");
  const int grid = ew_grid((int64_t)B * h * w * C / 4);
  if (f_scale)
    vgg_bwd_combine4_kernel<true, true><<<grid, 256, 0, ST(stream)>>>(g_next, pl_of<true>(fg_hi, fg_lo, f_scale),
                                                                      pl_of<true>(fp_hi, fp_lo, f_scale), B, h, w, C, mask, R,
                                                                      coef, pl_of<true>(dy_hi, dy_lo, dy_scale));
  else if (v4)
    vgg_bwd_combine4_kernel<false, false><<<grid, 256, 0, ST(stream)>>>(g_next, pl_of<false>(fg_hi, fg_lo, nullptr),
                                                                        pl_of<false>(fp_hi, fp_lo, nullptr), B, h, w, C, mask,
                                                                        R, coef, pl_of<false>(dy_hi, dy_lo, nullptr));
  else
    vgg_bwd_combine_kernel<<<ew_grid((int64_t)B * h * w * C), 256, 0, ST(stream)>>>(
        g_next, (const float*)fg_hi, (const float*)fg_lo, (const float*)fp_hi, (const float*)fp_lo, B, h, w, C, mask, R, coef,
        (float*)dy_hi, (float*)dy_lo);
  return check_launch("vgg_bwd_combine");
}

extern "C" int immb_pred_grad(const float* gt, const float* pred, int pcs, const float* mask,
                              const float* coef_input, const float* g_vggin, int g_is_patch, int B, int R,
                              float* g_hi, float* g_lo, void* stream) {
  IMMB_REQUIRE(gt && pred && coef_input && g_hi && pcs >= 3, "pred_grad: bad args");
  pred_grad_kernel<<<ew_grid((int64_t)B * R * R * pcs), 256, 0, ST(stream)>>>(
      gt, pred, pcs, mask, coef_input, g_vggin, g_is_patch, B, R, g_hi, g_lo);
  return check_launch("pred_grad");
}

extern "C" int immb_vgg_conv1_1_bwd_fused(const void* dy_hi, const void* dy_lo, const float* w, int Cout,
                                          const float* gt, const float* pred, int pcs, const float* mask,
                                          const float* coef_input, int B, int R, void* g_hi, void* g_lo,
                                          const int32_t* dy_scale, int32_t* g_scale, void* stream) {
  IMMB_REQUIRE(dy_hi && dy_lo && w && gt && pred && coef_input && g_hi && g_lo && pcs >= 3 && B > 0,
               "vgg_conv1_1_bwd_fused: bad args");
  IMMB_REQUIRE(Cout == 64 && R % 16 == 0, "vgg_conv1_1_bwd_fused: Cout must be 64 and 16 | R");
  IMMB_REQUIRE(aligned16(dy_hi) && aligned16(dy_lo), "vgg_conv1_1_bwd_fused: alignment");
  const int grid = B * (R / 16) * (R / 16);
  if (dy_scale && g_scale)
    vgg_conv1_1_bwd_fused_kernel<true, true><<<grid, 256, 0, ST(stream)>>>(pl_of<true>(dy_hi, dy_lo, dy_scale), w, gt, pred, pcs,
                                                                           mask, coef_input, R, pl_of<true>(g_hi, g_lo, g_scale));
  else if (dy_scale)
    vgg_conv1_1_bwd_fused_kernel<true, false><<<grid, 256, 0, ST(stream)>>>(pl_of<true>(dy_hi, dy_lo, dy_scale), w, gt, pred, pcs,
                                                                            mask, coef_input, R, pl_of<false>(g_hi, g_lo, nullptr));
  else if (g_scale)
    vgg_conv1_1_bwd_fused_kernel<false, true><<<grid, 256, 0, ST(stream)>>>(pl_of<false>(dy_hi, dy_lo, nullptr), w, gt, pred, pcs,
                                                                            mask, coef_input, R, pl_of<true>(g_hi, g_lo, g_scale));
  else
    vgg_conv1_1_bwd_fused_kernel<false, false><<<grid, 256, 0, ST(stream)>>>(pl_of<false>(dy_hi, dy_lo, nullptr), w, gt, pred, pcs,
                                                                             mask, coef_input, R, pl_of<false>(g_hi, g_lo, nullptr));
  return check_launch("vgg_conv1_1_bwd_fused");
}

extern "C" int immb_resize_ac_fwd(const void* x_hi, const void* x_lo, int xcs, int N, int H, int W, int C,
                                  int Ho, int Wo, void* o_hi, void* o_lo, int ocs, const int32_t* x_scale,
                                  int32_t* o_scale, void* stream) {
  IMMB_REQUIRE(x_hi && o_hi && xcs >= C && ocs >= C, "resize_ac_fwd: bad args");
  IMMB_REQUIRE((!x_scale || x_lo) && (!o_scale || o_lo), "resize_ac_fwd: fp16 planes need both planes");
  const int grid = ew_grid((int64_t)N * Ho * Wo * C);
  if (x_scale && o_scale)
    resize_ac_fwd_kernel<true, true><<<grid, 256, 0, ST(stream)>>>(pl_of<true>(x_hi, x_lo, x_scale), xcs, N, H, W, C, Ho, Wo,
                                                                   pl_of<true>(o_hi, o_lo, o_scale), ocs);
  else if (x_scale)
    resize_ac_fwd_kernel<true, false><<<grid, 256, 0, ST(stream)>>>(pl_of<true>(x_hi, x_lo, x_scale), xcs, N, H, W, C, Ho, Wo,
                                                                    pl_of<false>(o_hi, o_lo, nullptr), ocs);
  else if (o_scale)
    resize_ac_fwd_kernel<false, true><<<grid, 256, 0, ST(stream)>>>(pl_of<false>(x_hi, x_lo, nullptr), xcs, N, H, W, C, Ho, Wo,
                                                                    pl_of<true>(o_hi, o_lo, o_scale), ocs);
  else
    resize_ac_fwd_kernel<false, false><<<grid, 256, 0, ST(stream)>>>(pl_of<false>(x_hi, x_lo, nullptr), xcs, N, H, W, C, Ho, Wo,
                                                                     pl_of<false>(o_hi, o_lo, nullptr), ocs);
  return check_launch("resize_ac_fwd");
}

extern "C" int immb_resize_ac_bwd(const float* g_out, int gcs, int N, int H, int W, int C, int Ho, int Wo,
                                  float* g_in, void* stream) {
  IMMB_REQUIRE(g_out && g_in && gcs >= C, "resize_ac_bwd: bad args");
  cudaError_t e = cudaMemsetAsync(g_in, 0, sizeof(float) * (size_t)N * H * W * C, ST(stream));
  if (e != cudaSuccess) return set_error(IMMB_ERR_CUDA, "resize_ac_bwd memset: %s", cudaGetErrorString(e));
  resize_ac_bwd_kernel<<<ew_grid((int64_t)N * Ho * Wo * C), 256, 0, ST(stream)>>>(g_out, gcs, N, H, W, C, Ho,
                                                                                 Wo, g_in);
  return check_launch("resize_ac_bwd");
}

extern "C" int immb_tps_warp(const float* src, int B, int H, int W, int C, const float* w_tps, int Hc, int Wc,
                             float* dst, void* stream) {
  IMMB_REQUIRE(src && w_tps && dst && B > 0 && H > 0 && W > 0 && C > 0 && C <= 8 && Hc > 0 && Wc > 0,
               "tps_warp: bad args (C <= 8)");
  IMMB_REQUIRE(src != dst, "tps_warp: in-place warp is not supported");
  size_t smem = sizeof(float) * (size_t)(Hc * Wc + 3) * 2;
  IMMB_REQUIRE(smem <= 40000, "tps_warp: too many control points");
  tps_warp_kernel<<<dim3(ceil_div((int64_t)H * W, 256), B), 256, smem, ST(stream)>>>(src, H, W, C, w_tps, Hc, Wc, dst);
  return check_launch("tps_warp");
}

extern "C" int immb_adam_norms(const float* p, const float* g, int64_t n, const int32_t* chunk_tensor,
                               const int64_t* chunk_off, const int32_t* chunk_len, int n_chunks,
                               const float* tensor_wd, float gscale, double* sq, double* wsq, void* stream) {
  IMMB_REQUIRE(p && g && chunk_tensor && chunk_off && chunk_len && tensor_wd && sq && wsq && n_chunks > 0,
               "adam_norms: bad args");
  adam_norms_kernel<<<n_chunks, 256, 0, ST(stream)>>>(p, g, chunk_tensor, chunk_off, chunk_len, tensor_wd,
                                                      gscale, sq, wsq);
  return check_launch("adam_norms");
}

extern "C" int immb_adam_apply(float* p, const float* g, float* m, float* v, int64_t n,
                               const int32_t* chunk_tensor, const int64_t* chunk_off, const int32_t* chunk_len,
                               int n_chunks, const float* tensor_wd, float gscale, const double* sq, float clip,
                               float lr_t, float beta1, float beta2, float eps, float* amax, void* stream) {
  IMMB_REQUIRE(p && g && m && v && chunk_tensor && chunk_off && chunk_len && tensor_wd && sq && n_chunks > 0,
               "adam_apply: bad args");
  adam_apply_kernel<<<n_chunks, 256, 0, ST(stream)>>>(p, g, m, v, chunk_tensor, chunk_off, chunk_len,
                                                      tensor_wd, gscale, sq, clip, lr_t, beta1, beta2, eps, nullptr, amax);
  return check_launch("adam_apply");
}

extern "C" int immb_adam_apply_dev(float* p, const float* g, float* m, float* v, int64_t n,
                                   const int32_t* chunk_tensor, const int64_t* chunk_off, const int32_t* chunk_len,
                                   int n_chunks, const float* tensor_wd, float gscale, const double* sq, float clip,
                                   const float* lr_t_dev, float beta1, float beta2, float eps, float* amax,
                                   void* stream) {
  IMMB_REQUIRE(p && g && m && v && chunk_tensor && chunk_off && chunk_len && tensor_wd && sq && n_chunks > 0 && lr_t_dev,
               "adam_apply_dev: bad args");
  adam_apply_kernel<<<n_chunks, 256, 0, ST(stream)>>>(p, g, m, v, chunk_tensor, chunk_off, chunk_len,
                                                      tensor_wd, gscale, sq, clip, 0.f, beta1, beta2, eps, lr_t_dev, amax);
  return check_launch("adam_apply_dev");
}

extern "C" int immb_total_loss(const float* rec_loss, const double* wsq, const float* tensor_wd, int n_tensors,
                               float* weights_loss, float* total, const int32_t* overflow, void* stream) {
  IMMB_REQUIRE(rec_loss && wsq && tensor_wd && weights_loss && total, "total_loss: bad args");
  total_loss_kernel<<<1, 32, 0, ST(stream)>>>(rec_loss, wsq, tensor_wd, n_tensors, weights_loss, total, overflow);
  return check_launch("total_loss");
}

extern "C" int immb_pack_weights_multi(const immb_pack_item* items, int n_items, int total_blocks, void* stream) {
  IMMB_REQUIRE(items && n_items > 0 && total_blocks > 0, "pack_weights_multi: bad args");
  pack_weights_multi_kernel<<<total_blocks, 256, 0, ST(stream)>>>(items, n_items);
  return check_launch("pack_weights_multi");
}

extern "C" int immb_multi_amax(const float* p, const int32_t* chunk_tensor, const int64_t* chunk_off,
                               const int32_t* chunk_len, int n_chunks, float* amax, void* stream) {
  IMMB_REQUIRE(p && chunk_tensor && chunk_off && chunk_len && amax && n_chunks > 0, "multi_amax: bad args");
  multi_amax_kernel<<<n_chunks, 256, 0, ST(stream)>>>(p, chunk_tensor, chunk_off, chunk_len, amax);
  return check_launch("multi_amax");
}

extern "C" int immb_scale_update(int32_t* recs, int n, int32_t* overflow, void* stream) {
  IMMB_REQUIRE(recs && overflow && n > 0, "scale_update: bad args");
  scale_update_kernel<<<ceil_div(n, 128), 128, 0, ST(stream)>>>(recs, n, overflow);
  return check_launch("scale_update");
}

extern "C" int immb_pack_weights(const float* w, int kh, int kw, int Cin, int Cout, int cin_pad, int cout_pad,
                                 void* wp_hi, void* wp_lo, void* wh_hi, void* wh_lo, const float* amax,
                                 int32_t* w_scale, void* stream) {
  IMMB_REQUIRE(w && (wp_hi || wh_hi) && cin_pad >= Cin && cout_pad >= Cout, "pack_weights: bad args");
  IMMB_REQUIRE((amax != nullptr) == (w_scale != nullptr), "pack_weights: fp16 planes need both the tensor's amax and its scale record");
  if (w_scale)
    pack_weights_h16_kernel<<<ew_grid((int64_t)kh * kw * cin_pad * cout_pad), 256, 0, ST(stream)>>>(
        w, kh * kw, Cin, Cout, cin_pad, cout_pad, (uint16_t*)wp_hi, (uint16_t*)wp_lo, (uint16_t*)wh_hi, (uint16_t*)wh_lo,
        amax, w_scale);
  else
    pack_weights_kernel<<<ew_grid((int64_t)kh * kw * cin_pad * cout_pad), 256, 0, ST(stream)>>>(
        w, kh * kw, Cin, Cout, cin_pad, cout_pad, (float*)wp_hi, (float*)wp_lo, (float*)wh_hi, (float*)wh_lo);
  return check_launch("pack_weights");
}
