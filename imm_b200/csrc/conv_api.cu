// C-ABI entry points for the convolutions: validates the descriptor and dispatches to the tcgen05 engine
// (conv_tc.cu) or the SIMT engine (conv_simt.cu).
#include "common.cuh"

namespace immb {
int conv_simt_fwd(const immb_conv_desc*, const float*, const float*, const float*, const float*, float*, float*,
                  cudaStream_t);
int conv_simt_dgrad(const immb_conv_desc*, const float*, const float*, const float*, float*, cudaStream_t);
int conv_simt_wgrad(const immb_conv_desc*, const float*, const float*, const float*, const float*, float*,
                    cudaStream_t);
// tcgen05 engine (conv_tc.cu)
bool conv_tc_eligible(const immb_conv_desc* d, int op);
int conv_tc_fwd(const immb_conv_desc*, const void* x_hi, const void* x_lo, const void* wp_hi,
                const void* wp_lo, const float* bias, void* y_hi, void* y_lo, cudaStream_t, double* stats = nullptr);
int conv_tc_fwd_stats_rows(const immb_conv_desc* d);
int conv_tc_dgrad(const immb_conv_desc*, const void* dy_hi, const void* dy_lo, const void* wh_hi,
                  const void* wh_lo, float* dx, cudaStream_t);
int conv_tc_dgrad_stats_rows(const immb_conv_desc* d);
int conv_tc_dgrad_bnreduce(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo, const void* wh_hi,
                           const void* wh_lo, float* dx, const float* y_prev, int y_prev_cs, const float* scale,
                           const float* shift, const float* mean, const float* invstd, int relu, double* partials,
                           cudaStream_t st);
bool conv_tc_dgrad_relu_eligible(const immb_conv_desc* d);
int conv_tc_dgrad_relu(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo, const void* wh_hi,
                       const void* wh_lo, const void* act_hi, int act_cs, void* out_hi, void* out_lo, int32_t* out_scale,
                       cudaStream_t);
size_t conv_tc_wgrad_workspace(const immb_conv_desc*);
int conv_tc_wgrad(const immb_conv_desc*, const void* x_hi, const void* x_lo, const void* dy_hi,
                  const void* dy_lo, float* dw, void* ws, size_t ws_bytes, cudaStream_t);

static inline bool is_f16_prec(int p) { return p == IMMB_PREC_F16X3 || p == IMMB_PREC_F16X2; }

static int validate(const immb_conv_desc* d) {
  IMMB_REQUIRE(d, "conv: null descriptor");
  IMMB_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv: bad extents");
  IMMB_REQUIRE(d->kh > 0 && d->kw > 0 && d->kh <= 7 && d->kw <= 7, "conv: kernel size must be 1..7");
  IMMB_REQUIRE(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  IMMB_REQUIRE(d->Ho == (d->H + d->stride - 1) / d->stride && d->Wo == (d->W + d->stride - 1) / d->stride,
               "conv: Ho/Wo must be ceil(H/stride), ceil(W/stride) (TF SAME)");
  IMMB_REQUIRE(d->x_cstride >= d->Cin && d->y_cstride >= d->Cout, "conv: channel strides too small");
  IMMB_REQUIRE(d->pad_t >= 0 && d->pad_l >= 0 && d->pad_t < d->kh && d->pad_l < d->kw, "conv: bad padding");
  IMMB_REQUIRE(d->x_layout == IMMB_XLAYOUT_NHWC || d->engine != IMMB_ENGINE_SIMT,
               "conv: the SIMT engine reads plain NHWC only");
  IMMB_REQUIRE(d->precision >= IMMB_PREC_TF32X3 && d->precision <= IMMB_PREC_F16X2, "conv: bad precision");
  IMMB_REQUIRE(!is_f16_prec(d->precision) || d->engine != IMMB_ENGINE_SIMT,
               "conv: fp16 planes are read by the tcgen05 engine only");
  return IMMB_OK;
}

static int pick_engine(const immb_conv_desc* d, int op, int* engine) {
  bool ok = conv_tc_eligible(d, op);
  if (d->engine == IMMB_ENGINE_TC) {
    if (!ok) return set_error(IMMB_ERR_UNSUPPORTED, "conv: shape not eligible for the tcgen05 engine");
    *engine = IMMB_ENGINE_TC;
  } else if (d->engine == IMMB_ENGINE_SIMT) {
    *engine = IMMB_ENGINE_SIMT;
  } else {
    *engine = ok ? IMMB_ENGINE_TC : IMMB_ENGINE_SIMT;
    if (!ok && d->x_layout != IMMB_XLAYOUT_NHWC)
      return set_error(IMMB_ERR_UNSUPPORTED, "conv: x_layout ROWWIN4 needs the tcgen05 engine");
    if (!ok && is_f16_prec(d->precision))
      return set_error(IMMB_ERR_UNSUPPORTED, "conv: this shape / op has no fp16-plane kernel (use a TF32 precision)");
  }
  return IMMB_OK;
}
}  // namespace immb

using namespace immb;

extern "C" int immb_conv_engine_for(const immb_conv_desc* d, int op) {
  if (validate(d) != IMMB_OK) return IMMB_ERR_INVALID;
  immb_conv_desc t = *d;
  t.engine = IMMB_ENGINE_AUTO;
  int e = IMMB_ENGINE_SIMT;
  pick_engine(&t, op, &e);
  return e;
}

extern "C" int immb_conv2d_fwd(const immb_conv_desc* d, const void* x_hi, const void* x_lo, const float* w,
                               const void* wp_hi, const void* wp_lo, const float* bias, void* y_hi,
                               void* y_lo, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  IMMB_REQUIRE(x_hi && y_hi, "conv2d_fwd: null tensors");
  int engine = IMMB_ENGINE_SIMT;
  if ((rc = pick_engine(d, 0, &engine))) return rc;
  if (engine == IMMB_ENGINE_TC) {
    IMMB_REQUIRE(wp_hi && (d->precision == IMMB_PREC_TF32 || (wp_lo && x_lo)),
                 "conv2d_fwd: tcgen05 engine needs packed weights and (for TF32x3 / TF32x2) lo planes");
    return conv_tc_fwd(d, x_hi, x_lo, wp_hi, wp_lo, bias, y_hi, y_lo, (cudaStream_t)stream);
  }
  IMMB_REQUIRE(w, "conv2d_fwd: SIMT engine needs the master weights");
  return conv_simt_fwd(d, (const float*)x_hi, (const float*)x_lo, w, bias, (float*)y_hi, (float*)y_lo, (cudaStream_t)stream);
}

extern "C" int immb_conv2d_fwd_stats_rows(const immb_conv_desc* d) {
  if (validate(d) != IMMB_OK) return 0;
  int engine = IMMB_ENGINE_SIMT;
  if (pick_engine(d, 0, &engine) || engine != IMMB_ENGINE_TC) return 0;
  return conv_tc_fwd_stats_rows(d);
}

extern "C" int immb_conv2d_fwd_bnstats(const immb_conv_desc* d, const void* x_hi, const void* x_lo,
                                       const void* wp_hi, const void* wp_lo, const float* bias, float* y,
                                       double* partials, size_t partial_elems, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  IMMB_REQUIRE(x_hi && x_lo && wp_hi && wp_lo && y && partials, "conv2d_fwd_bnstats: null tensors");
  const int rows = immb_conv2d_fwd_stats_rows(d);
  if (rows <= 0) return immb::set_error(IMMB_ERR_UNSUPPORTED, "conv2d_fwd_bnstats: shape not served by the 3-pass pair kernel");
  IMMB_REQUIRE(partial_elems >= (size_t)rows * 2 * (size_t)d->Cout, "conv2d_fwd_bnstats: partials buffer too small");
  return conv_tc_fwd(d, x_hi, x_lo, wp_hi, wp_lo, bias, y, nullptr, (cudaStream_t)stream, partials);
}

extern "C" int immb_conv2d_dgrad(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo,
                                 const float* w, const void* wh_hi, const void* wh_lo, float* dx,
                                 void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  IMMB_REQUIRE(dy_hi && dx, "conv2d_dgrad: null tensors");
  int engine = IMMB_ENGINE_SIMT;
  if ((rc = pick_engine(d, 1, &engine))) return rc;
  if (engine == IMMB_ENGINE_TC) {
    IMMB_REQUIRE(wh_hi && (d->precision == IMMB_PREC_TF32 || (wh_lo && dy_lo)),
                 "conv2d_dgrad: tcgen05 engine needs split weights and (for TF32x3) lo planes");
    return conv_tc_dgrad(d, dy_hi, dy_lo, wh_hi, wh_lo, dx, (cudaStream_t)stream);
  }
  IMMB_REQUIRE(w, "conv2d_dgrad: SIMT engine needs the master weights");
  return conv_simt_dgrad(d, (const float*)dy_hi, (const float*)dy_lo, w, dx, (cudaStream_t)stream);
}

extern "C" int immb_conv2d_dgrad_stats_rows(const immb_conv_desc* d) {
  if (validate(d) != IMMB_OK) return 0;
  int engine = IMMB_ENGINE_SIMT;
  if (pick_engine(d, 1, &engine) || engine != IMMB_ENGINE_TC) return 0;
  return conv_tc_dgrad_stats_rows(d);
}

extern "C" int immb_conv2d_dgrad_bnreduce(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo,
                                          const void* wh_hi, const void* wh_lo, float* dx, const float* y_prev,
                                          int y_prev_cstride, const float* scale, const float* shift, const float* mean,
                                          const float* invstd, int relu, double* partials, size_t partial_elems,
                                          void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  IMMB_REQUIRE(dy_hi && dy_lo && wh_hi && wh_lo && dx && y_prev && scale && shift && mean && invstd && partials,
               "conv2d_dgrad_bnreduce: null tensors");
  IMMB_REQUIRE(y_prev_cstride >= d->Cin && y_prev_cstride % 4 == 0, "conv2d_dgrad_bnreduce: bad y stride");
  const int rows = immb_conv2d_dgrad_stats_rows(d);
  if (rows <= 0) return immb::set_error(IMMB_ERR_UNSUPPORTED, "conv2d_dgrad_bnreduce: shape not served by the 3-pass pair kernel");
  IMMB_REQUIRE(partial_elems >= (size_t)rows * 2 * (size_t)d->Cin, "conv2d_dgrad_bnreduce: partials buffer too small");
  return conv_tc_dgrad_bnreduce(d, dy_hi, dy_lo, wh_hi, wh_lo, dx, y_prev, y_prev_cstride, scale, shift, mean, invstd, relu,
                                partials, (cudaStream_t)stream);
}

extern "C" int immb_conv2d_dgrad_relu(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo,
                                      const void* wh_hi, const void* wh_lo, const void* act_hi, int act_cstride,
                                      void* out_hi, void* out_lo, int32_t* out_scale, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  IMMB_REQUIRE(dy_hi && wh_hi && act_hi && out_hi && out_lo && act_cstride >= d->Cin, "conv2d_dgrad_relu: null tensors");
  IMMB_REQUIRE(d->precision == IMMB_PREC_TF32 ||
                   (dy_lo && (d->precision == IMMB_PREC_TF32X2 || d->precision == IMMB_PREC_F16X2 || wh_lo)),
               "conv2d_dgrad_relu: lo planes missing");
  IMMB_REQUIRE(!is_f16_prec(d->precision) || out_scale, "conv2d_dgrad_relu: fp16 output planes need their scale record");
  if (!conv_tc_dgrad_relu_eligible(d))
    return immb::set_error(IMMB_ERR_UNSUPPORTED, "conv2d_dgrad_relu: shape not covered by the halo pair kernel");
  return conv_tc_dgrad_relu(d, dy_hi, dy_lo, wh_hi, wh_lo, act_hi, act_cstride, out_hi, out_lo, out_scale,
                            (cudaStream_t)stream);
}

extern "C" int immb_conv2d_dgrad_relu_supported(const immb_conv_desc* d) {
  return (validate(d) == IMMB_OK && conv_tc_dgrad_relu_eligible(d)) ? 1 : 0;
}

extern "C" size_t immb_conv2d_wgrad_workspace(const immb_conv_desc* d) {
  if (validate(d) != IMMB_OK) return 0;
  int engine = IMMB_ENGINE_SIMT;
  if (pick_engine(d, 2, &engine)) return 0;
  return engine == IMMB_ENGINE_TC ? conv_tc_wgrad_workspace(d) : 0;
}

extern "C" int immb_conv2d_wgrad(const immb_conv_desc* d, const void* x_hi, const void* x_lo,
                                 const void* dy_hi, const void* dy_lo, float* dw, void* workspace,
                                 size_t ws_bytes, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  IMMB_REQUIRE(x_hi && dy_hi && dw, "conv2d_wgrad: null tensors");
  int engine = IMMB_ENGINE_SIMT;
  if ((rc = pick_engine(d, 2, &engine))) return rc;
  if (engine == IMMB_ENGINE_TC) {
    IMMB_REQUIRE(d->precision == IMMB_PREC_TF32 || (x_lo && dy_lo), "conv2d_wgrad: TF32x3 needs lo planes");
    return conv_tc_wgrad(d, x_hi, x_lo, dy_hi, dy_lo, dw, workspace, ws_bytes, (cudaStream_t)stream);
  }
  return conv_simt_wgrad(d, (const float*)x_hi, (const float*)x_lo, (const float*)dy_hi, (const float*)dy_lo, dw,
                         (cudaStream_t)stream);
}
