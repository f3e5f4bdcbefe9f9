// fc_linspline.cu — piecewise-linear spline layer (forward / inverse / backward) for sm_100a  (SURVEY.md §8(f) n3).
//
// Replaces linear_spline / unconstrained_linear_spline (flowcon/transforms/splines/linear.py:9-105), searchsorted
// (utils/torchutils.py:147-149), sum_except_batch (:25-30) and the coupling column split / scatter
// (transforms/coupling.py:82-83,96-98) for PiecewiseLinearCouplingTransform (coupling.py:299-352),
// MaskedPiecewiseLinearAutoregressiveTransform (autoregressive.py:321-372) and PiecewiseLinearCDF
// (nonlinearities.py:250-283).  Same kernel skeletons as the rational-quadratic layer (fc_pipeline.cuh: per-warp TMA
// ring; fc_staged.cuh: general strides); element math: fc_math.cuh.  HBM-bound: 4 (K + 2) bytes per transformed element.
#include "fc_pipeline.cuh"
#include "fc_made_inverse.cuh"

namespace fc {

template <int KC>
struct LinSplineOp {
  static constexpr int kTileWarps = 12;  // little arithmetic per row: more consumer warps in flight (fc_pipeline.cuh)
  LinSplineParams c;
  __device__ __forceinline__ int P() const { return c.K; }
  __device__ __forceinline__ void eval(float x, const float* p, float& y, float& lad, unsigned& status) const {
    linspline_eval<KC>(c, x, p, y, lad, status);
  }
  __device__ __forceinline__ void backward(float x, const float* p, float gy, float gl, float& gx, float* gp) const {
    linspline_backward_elem<KC>(c, x, p, gy, gl, gx, gp);
  }
};

#define FC_DISPATCH_LIN_K(K, CALL) \
  switch (K) {                     \
    case 4: CALL(4); break;        \
    case 8: CALL(8); break;        \
    case 10: CALL(10); break;      \
    case 16: CALL(16); break;      \
    default: CALL(0); break;       \
  }

static int make_linspline_params(int32_t num_bins, int32_t tails, float left, float right, float bottom, float top,
                                 int32_t inverse, LinSplineParams& c) {
  if (num_bins < 1) return FC_ERR_INVALID_ARGUMENT;
  if (num_bins > FC_MAX_BINS_GENERIC) return FC_ERR_UNSUPPORTED;
  if (tails != FC_TAILS_NONE && tails != FC_TAILS_LINEAR) return FC_ERR_INVALID_ARGUMENT;
  if (!(right > left) || !(top > bottom)) return FC_ERR_INVALID_ARGUMENT;
  c.K = num_bins;
  c.tails = tails;
  c.inverse = inverse != 0;
  c.left = left; c.right = right; c.bottom = bottom; c.top = top;
  c.log_k = (float)log((double)num_bins);
  c.inv_w = (float)(1.0 / ((double)right - (double)left));
  c.inv_h = (float)(1.0 / ((double)top - (double)bottom));
  return FC_OK;
}

template <int KC>
struct MadeLinSplineOp {  // incremental autoregressive inverse (fc_made_inverse.cuh)
  LinSplineParams c;
  __device__ __forceinline__ void eval(float z, const float* pc, float& x, float& lad, unsigned& status) const {
    float p[KC ? KC : FC_MAX_BINS_GENERIC];
    made_load_params(pc, c.K, p);
    linspline_eval<KC>(c, z, p, x, lad, status);
  }
};

}  // namespace fc

using namespace fc;

extern "C" int fc_made_inverse_linspline(const fc_made_program* prog, const float* z, int64_t z_row_stride, float* x,
                                         int64_t x_row_stride, float* logabsdet, int32_t accumulate_logabsdet, int64_t B,
                                         int32_t num_bins, int32_t tails, float left, float right, float bottom, float top,
                                         int32_t* status, void* stream) {
  LinSplineParams c;
  int rc = make_linspline_params(num_bins, tails, left, right, bottom, top, 1, c);
  if (rc != FC_OK) return rc;
  MadeArgs a{};
  rc = made_check(prog, z, z_row_stride, x, x_row_stride, logabsdet, B, c.K, a);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  a.accumulate = accumulate_logabsdet;
  a.status = status;
#define CALL(KC)                                     \
  {                                                  \
    MadeLinSplineOp<KC> op;                          \
    op.c = c;                                        \
    return launch_made(a, op, (cudaStream_t)stream); \
  }
  FC_DISPATCH_LIN_K(c.K, CALL)
#undef CALL
  return FC_OK;
}

extern "C" int fc_linspline_apply(const float* x, int64_t x_row_stride, const float* params, int64_t params_row_stride,
                                  float* y, int64_t y_row_stride, float* logabsdet, int32_t accumulate_logabsdet,
                                  int64_t B, int32_t D_t, fc_cols tcols, fc_cols ccols, int32_t num_bins, int32_t tails,
                                  float left, float right, float bottom, float top, int32_t inverse, int32_t* status,
                                  void* stream) {
  LinSplineParams c;
  int rc = make_linspline_params(num_bins, tails, left, right, bottom, top, inverse, c);
  if (rc != FC_OK) return rc;
  rc = check_layer_args(x, params, y, B, D_t, tcols, ccols);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  if (!logabsdet) return FC_ERR_INVALID_ARGUMENT;
  LayerArgs a;
  a.x = x; a.params = params; a.y = y; a.lad = logabsdet; a.status = status;
  a.x_stride = x_row_stride; a.p_stride = params_row_stride; a.y_stride = y_row_stride;
  a.B = B; a.D_t = D_t; a.n_copy = ccols.n; a.tcols = tcols.idx; a.ccols = ccols.idx;
  a.accumulate = accumulate_logabsdet;
  const size_t smem = plan_tiles(a, c.K);
#define CALL(KC)                                                                                 \
  {                                                                                              \
    LinSplineOp<KC> op;                                                                          \
    op.c = c;                                                                                    \
    const int piped = try_launch_pipelined(a, op, c.K, (int)x_row_stride, (cudaStream_t)stream); \
    if (piped != 0) return piped < 0 ? piped : FC_OK;                                            \
    return launch_apply(a, op, smem, (cudaStream_t)stream);                                      \
  }
  FC_DISPATCH_LIN_K(c.K, CALL)
#undef CALL
  return FC_OK;
}

extern "C" int fc_linspline_backward(const float* x, int64_t x_row_stride, const float* params, int64_t params_row_stride,
                                     const float* grad_y, int64_t gy_row_stride, const float* grad_logabsdet,
                                     float* grad_x, int64_t gx_row_stride, float* grad_params, int64_t gp_row_stride,
                                     int64_t B, int32_t D_t, fc_cols tcols, fc_cols ccols, int32_t num_bins,
                                     int32_t tails, float left, float right, float bottom, float top, int32_t inverse,
                                     void* stream) {
  LinSplineParams c;
  int rc = make_linspline_params(num_bins, tails, left, right, bottom, top, inverse, c);
  if (rc != FC_OK) return rc;
  rc = check_layer_args(x, params, grad_x, B, D_t, tcols, ccols);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  if (!grad_y || !grad_params) return FC_ERR_INVALID_ARGUMENT;
  LayerBwdArgs a;
  a.x = x; a.params = params; a.gy = grad_y; a.gl = grad_logabsdet; a.gx = grad_x; a.gp = grad_params;
  a.x_stride = x_row_stride; a.p_stride = params_row_stride; a.gy_stride = gy_row_stride;
  a.gx_stride = gx_row_stride; a.gp_stride = gp_row_stride;
  a.B = B; a.D_t = D_t; a.n_copy = ccols.n; a.tcols = tcols.idx; a.ccols = ccols.idx;
  const size_t smem = plan_tiles(a, c.K);
#define CALL(KC)                                                                                          \
  {                                                                                                       \
    LinSplineOp<KC> op;                                                                                   \
    op.c = c;                                                                                             \
    const int piped = try_launch_pipelined_backward(a, op, c.K, (int)x_row_stride, (cudaStream_t)stream); \
    if (piped != 0) return piped < 0 ? piped : FC_OK;                                                     \
    return launch_backward(a, op, smem, (cudaStream_t)stream);                                            \
  }
  FC_DISPATCH_LIN_K(c.K, CALL)
#undef CALL
  return FC_OK;
}
