// fc_quadspline.cu — piecewise-quadratic spline layer (forward / inverse / backward) for sm_100a  (SURVEY.md §8(f) n3).
//
// Replaces quadratic_spline / unconstrained_quadratic_spline (flowcon/transforms/splines/quadratic.py:11-159),
// searchsorted (utils/torchutils.py:147-149), sum_except_batch (:25-30) and the coupling column split / scatter for
// PiecewiseQuadraticCouplingTransform (coupling.py:355-427), MaskedPiecewiseQuadraticAutoregressiveTransform
// (autoregressive.py:375-450) and PiecewiseQuadraticCDF (nonlinearities.py:286-340).  Kernel skeletons: fc_pipeline.cuh
// (per-warp TMA ring) / fc_staged.cuh (general strides); element math: fc_math.cuh.  HBM-bound: 4 (P + 2) bytes per
// transformed element, P = 2K+1 (no tails) or 2K-1 (linear tails).
#include "fc_pipeline.cuh"
#include "fc_made_inverse.cuh"

namespace fc {

template <int KC>
struct QuadSplineOp {
  static constexpr int kTileWarps = 12;  // tile ring: 12 consumer warps x 2 CTAs (68 registers)
  static constexpr int kMinBlocks = 1;  // arithmetic-heavy: the full 128 registers instead of spills (fc_pipeline.cuh)
  QuadSplineParams c;
  __device__ __forceinline__ int P() const { return c.tails == FC_TAILS_LINEAR ? 2 * c.K - 1 : 2 * c.K + 1; }
  __device__ __forceinline__ void eval(float x, const float* p, float& y, float& lad, unsigned& status) const {
    quadspline_eval<KC>(c, x, p, y, lad, status);
  }
  __device__ __forceinline__ void backward(float x, const float* p, float gy, float gl, float& gx, float* gp) const {
    quadspline_backward_elem<KC>(c, x, p, gy, gl, gx, gp);
  }
};

#define FC_DISPATCH_QUAD_K(K, CALL) \
  switch (K) {                      \
    case 4: CALL(4); break;         \
    case 8: CALL(8); break;         \
    case 10: CALL(10); break;       \
    default: CALL(0); break;        \
  }

static int make_quadspline_params(const fc_quadspline_config* cfg, QuadSplineParams& c) {
  if (!cfg) return FC_ERR_INVALID_ARGUMENT;
  if (cfg->num_bins < 1 || (cfg->tails == FC_TAILS_LINEAR && cfg->num_bins < 2)) return FC_ERR_INVALID_ARGUMENT;
  if (cfg->num_bins > FC_MAX_BINS_GENERIC) return FC_ERR_UNSUPPORTED;
  if (cfg->tails != FC_TAILS_NONE && cfg->tails != FC_TAILS_LINEAR) return FC_ERR_INVALID_ARGUMENT;
  if (!(cfg->right > cfg->left) || !(cfg->top > cfg->bottom)) return FC_ERR_INVALID_ARGUMENT;
  // quadratic.py:77-80
  if (cfg->min_bin_width * cfg->num_bins > 1.f || cfg->min_bin_height * cfg->num_bins > 1.f) return FC_ERR_INVALID_ARGUMENT;
  c.K = cfg->num_bins;
  c.tails = cfg->tails;
  c.inverse = cfg->inverse != 0;
  c.left = cfg->left; c.right = cfg->right; c.bottom = cfg->bottom; c.top = cfg->top;
  c.inv_w = (float)(1.0 / ((double)cfg->right - (double)cfg->left));
  c.inv_h = (float)(1.0 / ((double)cfg->top - (double)cfg->bottom));
  c.min_w = cfg->min_bin_width;
  c.min_h = cfg->min_bin_height;
  c.wh_scale = cfg->wh_scale;
  return FC_OK;
}

template <int KC>
struct MadeQuadSplineOp {  // incremental autoregressive inverse (fc_made_inverse.cuh)
  QuadSplineParams c;
  int P;
  __device__ __forceinline__ void eval(float z, const float* pc, float& x, float& lad, unsigned& status) const {
    float p[2 * (KC ? KC : FC_MAX_BINS_GENERIC) + 1];
    made_load_params(pc, P, p);
    quadspline_eval<KC>(c, z, p, x, lad, status);
  }
};

}  // namespace fc

using namespace fc;

extern "C" int fc_made_inverse_quadspline(const fc_made_program* prog, const float* z, int64_t z_row_stride, float* x,
                                          int64_t x_row_stride, float* logabsdet, int32_t accumulate_logabsdet, int64_t B,
                                          const fc_quadspline_config* cfg, int32_t* status, void* stream) {
  QuadSplineParams c;
  int rc = make_quadspline_params(cfg, c);
  if (rc != FC_OK) return rc;
  if (!c.inverse) return FC_ERR_INVALID_ARGUMENT;
  const int P = c.tails == FC_TAILS_LINEAR ? 2 * c.K - 1 : 2 * c.K + 1;
  MadeArgs a{};
  rc = made_check(prog, z, z_row_stride, x, x_row_stride, logabsdet, B, P, a);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  a.accumulate = accumulate_logabsdet;
  a.status = status;
#define CALL(KC)                                     \
  {                                                  \
    MadeQuadSplineOp<KC> op;                         \
    op.c = c;                                        \
    op.P = P;                                        \
    return launch_made(a, op, (cudaStream_t)stream); \
  }
  FC_DISPATCH_QUAD_K(c.K, CALL)
#undef CALL
  return FC_OK;
}

extern "C" int fc_quadspline_apply(const float* x, int64_t x_row_stride, const float* params, int64_t params_row_stride,
                                   float* y, int64_t y_row_stride, float* logabsdet, int32_t accumulate_logabsdet,
                                   int64_t B, int32_t D_t, fc_cols tcols, fc_cols ccols,
                                   const fc_quadspline_config* cfg, int32_t* status, void* stream) {
  QuadSplineParams c;
  int rc = make_quadspline_params(cfg, c);
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
  const int P = c.tails == FC_TAILS_LINEAR ? 2 * c.K - 1 : 2 * c.K + 1;
  const size_t smem = plan_tiles(a, P);
#define CALL(KC)                                                                               \
  {                                                                                            \
    QuadSplineOp<KC> op;                                                                       \
    op.c = c;                                                                                  \
    const int piped = try_launch_pipelined(a, op, P, (int)x_row_stride, (cudaStream_t)stream); \
    if (piped != 0) return piped < 0 ? piped : FC_OK;                                          \
    return launch_apply(a, op, smem, (cudaStream_t)stream);                                    \
  }
  FC_DISPATCH_QUAD_K(c.K, CALL)
#undef CALL
  return FC_OK;
}

extern "C" int fc_quadspline_backward(const float* x, int64_t x_row_stride, const float* params,
                                      int64_t params_row_stride, const float* grad_y, int64_t gy_row_stride,
                                      const float* grad_logabsdet, float* grad_x, int64_t gx_row_stride,
                                      float* grad_params, int64_t gp_row_stride, int64_t B, int32_t D_t, fc_cols tcols,
                                      fc_cols ccols, const fc_quadspline_config* cfg, void* stream) {
  QuadSplineParams c;
  int rc = make_quadspline_params(cfg, c);
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
  const int P = c.tails == FC_TAILS_LINEAR ? 2 * c.K - 1 : 2 * c.K + 1;
  const size_t smem = plan_tiles(a, P);
#define CALL(KC)                                                                                        \
  {                                                                                                     \
    QuadSplineOp<KC> op;                                                                                \
    op.c = c;                                                                                           \
    const int piped = try_launch_pipelined_backward(a, op, P, (int)x_row_stride, (cudaStream_t)stream); \
    if (piped != 0) return piped < 0 ? piped : FC_OK;                                                   \
    return launch_backward(a, op, smem, (cudaStream_t)stream);                                          \
  }
  FC_DISPATCH_QUAD_K(c.K, CALL)
#undef CALL
  return FC_OK;
}
