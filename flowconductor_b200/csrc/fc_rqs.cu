// fc_rqs.cu — rational-quadratic spline layer (forward / inverse / backward) for sm_100a.
//
// Replaces the eager op chain of unconstrained_rational_quadratic_spline / rational_quadratic_spline
// (flowcon/transforms/splines/rational_quadratic.py:13-181), searchsorted (utils/torchutils.py:147-149),
// sum_except_batch (utils/torchutils.py:25-30) and the coupling column split / scatter
// (transforms/coupling.py:82-83,96-98): ~151 ATen calls, 3 host syncs and ~158 KB of tensor traffic per
// sample per layer become one launch that touches each byte once.  Kernel skeleton: fc_staged.cuh;
// element math: fc_math.cuh.
#include "fc_pipeline.cuh"

namespace fc {

template <int KC>
struct RqsOp {
  RqsParams c;
  __device__ __forceinline__ int P() const { return c.P; }
  __device__ __forceinline__ void eval(float x, const float* p, float& y, float& lad, unsigned& status) const {
    rqs_eval<KC>(c, x, p, y, lad, status);
  }
  __device__ __forceinline__ void backward(float x, const float* p, float gy, float gl, float& gx, float* gp) const {
    rqs_backward_elem<KC>(c, x, p, gy, gl, gx, gp);
  }
};

// compile-time bin counts with register-resident knot arrays; anything else takes the runtime-K path
#define FC_DISPATCH_K(K, CALL) \
  switch (K) {                 \
    case 4: CALL(4); break;    \
    case 5: CALL(5); break;    \
    case 8: CALL(8); break;    \
    case 10: CALL(10); break;  \
    case 16: CALL(16); break;  \
    default: CALL(0); break;   \
  }

// Debug / parity aid: the bin every element falls into, computed by exactly the device arithmetic of the layer kernels
// (rqs_domain + rqs_locate of fc_math.cuh: base-2 softmax on the SFU, running-sum knots, "last true wins" search).
template <int KC>
__global__ void rqs_bins_kernel(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ params,
                                int64_t p_stride, const int32_t* __restrict__ tcols, int64_t B, int D_t, RqsParams c,
                                int32_t* __restrict__ bins, float* __restrict__ knot_dist) {
  const int64_t total = B * D_t;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D_t;
    const int j = (int)(i % D_t);
    const float xv = x[r * x_stride + (tcols ? tcols[j] : j)];
    const float* p = params + r * p_stride + (int64_t)j * c.P;
    unsigned status = 0;
    float xs;
    const bool inside = rqs_domain(c, xv, xs, status);
    const int K = KC ? KC : c.K;
    float ew[KC ? KC : FC_MAX_BINS_GENERIC], eh[KC ? KC : FC_MAX_BINS_GENERIC];
    float inv_w, inv_h;
    RqsBin b;
    rqs_locate<KC>(c, K, xs, p, b, ew, eh, inv_w, inv_h);
    bins[i] = inside ? b.k : -1;
    if (knot_dist) {
      // distance to the nearest knot of the searched axis, in units of the interval length
      const float lo = c.inverse ? b.ch : b.cw, sz = c.inverse ? b.h : b.w;
      const float span = c.inverse ? (c.top - c.bottom) : (c.right - c.left);
      knot_dist[i] = inside ? fminf(fabsf(xs - lo), fabsf(lo + sz - xs)) / span : INFINITY;
    }
  }
}

}  // namespace fc

using namespace fc;

extern "C" int fc_elementwise_last_path(void) { return last_path(); }

extern "C" int fc_rqs_bins(const float* x, int64_t x_row_stride, const float* params, int64_t params_row_stride, int64_t B,
                           int32_t D_t, fc_cols tcols, const fc_rqs_config* cfg, int32_t* bins, float* knot_dist,
                           void* stream) {
  RqsParams c;
  int rc = make_rqs_params(cfg, c);
  if (rc != FC_OK) return rc;
  if (B < 0 || D_t <= 0 || (tcols.idx && tcols.n != D_t)) return FC_ERR_INVALID_ARGUMENT;
  if (B == 0) return FC_OK;
  if (!x || !params || !bins) return FC_ERR_INVALID_ARGUMENT;
  const int64_t total = B * D_t;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
#define CALL(KC)                                                                                                        \
  rqs_bins_kernel<KC><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, x_row_stride, params, params_row_stride, tcols.idx, B, \
                                                                D_t, c, bins, knot_dist);
  FC_DISPATCH_K(c.K, CALL)
#undef CALL
  FC_CHECK_LAUNCH();
  return FC_OK;
}

extern "C" int fc_rqs_apply(const float* x, int64_t x_row_stride, const float* params, int64_t params_row_stride,
                            float* y, int64_t y_row_stride, float* logabsdet, int32_t accumulate_logabsdet, int64_t B,
                            int32_t D_t, fc_cols tcols, fc_cols ccols, const fc_rqs_config* cfg, int32_t* status,
                            void* stream) {
  RqsParams c;
  int rc = make_rqs_params(cfg, c);
  if (rc != FC_OK) return rc;
  rc = check_layer_args(x, params, y, B, D_t, tcols, ccols);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  if (!logabsdet) return FC_ERR_INVALID_ARGUMENT;
  if (B == 0) return FC_OK;
  LayerArgs a;
  a.x = x; a.params = params; a.y = y; a.lad = logabsdet; a.status = status;
  a.x_stride = x_row_stride; a.p_stride = params_row_stride; a.y_stride = y_row_stride;
  a.B = B; a.D_t = D_t; a.n_copy = ccols.n; a.tcols = tcols.idx; a.ccols = ccols.idx;
  a.accumulate = accumulate_logabsdet;
  const size_t smem = plan_tiles(a, c.P);
#define CALL(KC)                                                                                   \
  {                                                                                                \
    RqsOp<KC> op;                                                                                  \
    op.c = c;                                                                                      \
    const int piped = try_launch_pipelined(a, op, c.P, (int)x_row_stride, (cudaStream_t)stream);   \
    if (piped != 0) return piped < 0 ? piped : FC_OK;                                              \
    return launch_apply(a, op, smem, (cudaStream_t)stream);                                        \
  }
  FC_DISPATCH_K(c.K, CALL)
#undef CALL
  return FC_OK;
}

extern "C" int fc_rqs_backward(const float* x, int64_t x_row_stride, const float* params, int64_t params_row_stride,
                               const float* grad_y, int64_t gy_row_stride, const float* grad_logabsdet, float* grad_x,
                               int64_t gx_row_stride, float* grad_params, int64_t gp_row_stride, int64_t B,
                               int32_t D_t, fc_cols tcols, fc_cols ccols, const fc_rqs_config* cfg, void* stream) {
  RqsParams c;
  int rc = make_rqs_params(cfg, c);
  if (rc != FC_OK) return rc;
  rc = check_layer_args(x, params, grad_x, B, D_t, tcols, ccols);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  if (!grad_y || !grad_params) return FC_ERR_INVALID_ARGUMENT;
  if (B == 0) return FC_OK;
  LayerBwdArgs a;
  a.x = x; a.params = params; a.gy = grad_y; a.gl = grad_logabsdet; a.gx = grad_x; a.gp = grad_params;
  a.x_stride = x_row_stride; a.p_stride = params_row_stride; a.gy_stride = gy_row_stride;
  a.gx_stride = gx_row_stride; a.gp_stride = gp_row_stride;
  a.B = B; a.D_t = D_t; a.n_copy = ccols.n; a.tcols = tcols.idx; a.ccols = ccols.idx;
  const size_t smem = plan_tiles(a, c.P);
#define CALL(KC)                                                        \
  {                                                                     \
    RqsOp<KC> op;                                                       \
    op.c = c;                                                           \
    const int piped = try_launch_pipelined_backward(a, op, c.P, (int)x_row_stride, (cudaStream_t)stream); \
    if (piped != 0) return piped < 0 ? piped : FC_OK;                   \
    return launch_backward(a, op, smem, (cudaStream_t)stream);          \
  }
  FC_DISPATCH_K(c.K, CALL)
#undef CALL
  return FC_OK;
}
