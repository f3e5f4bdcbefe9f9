// fc_affine.cu — affine element-wise transforms and the standard-normal tail for sm_100a.
//
// Replaces AffineCouplingTransform._coupling_transform_forward/_inverse (flowcon/transforms/coupling.py
// :234-252, blocked params [shift | raw_scale]), MaskedAffineAutoregressiveTransform._elementwise_forward/
// _inverse (transforms/autoregressive/autoregressive.py:97-129, interleaved [raw_scale, shift] pairs) and
// StandardNormal._log_prob + the final add of Flow._log_prob (distributions/normal.py:23-33,
// flows/base.py:48).  16 B/element of traffic: no staging needed, one lane per feature, shuffle log-det.
#include "fc_pipeline.cuh"

namespace fc {

struct AffineArgs {
  const float* x;
  const float* params;
  const float* gy;
  const float* gl;
  float* y;   // forward: outputs; backward: grad_x
  float* lad; // forward only
  float* gp;  // backward only
  int64_t x_stride, p_stride, y_stride, gy_stride, gp_stride;
  int64_t B;
  int D_t, n_copy;
  const int32_t* tcols;
  const int32_t* ccols;
  int accumulate, layout, activation, inverse, seg;
};

__device__ __forceinline__ void affine_fetch(const AffineArgs& a, const float* prow, int j, float& raw, float& shift) {
  if (a.layout == FC_AFFINE_BLOCKED) {
    shift = prow[j];
    raw = prow[a.D_t + j];
  } else {
    raw = prow[2 * j];
    shift = prow[2 * j + 1];
  }
}

// Op for the per-warp TMA ring of fc_pipeline.cuh (contiguous rows whose column lists cover the row): parameters
// per feature are (raw scale, shift) pairs (interleaved) or prow[0] = shift, prow[D_t] = raw scale (blocked).
struct AffineOp {
  static constexpr int kTileBwdWarps = 12;
  static constexpr int kTileWarps = 12;  // little arithmetic per row: more consumer warps in flight (fc_pipeline.cuh)
  int D_t, layout, activation, inverse;
  __device__ __forceinline__ int P() const { return 2; }
  __device__ __forceinline__ int feature_stride() const { return layout == FC_AFFINE_BLOCKED ? 1 : 2; }
  __device__ __forceinline__ void eval(float x, const float* p, float& y, float& lad, unsigned&) const {
    const float raw = layout == FC_AFFINE_BLOCKED ? p[D_t] : p[0];
    const float shift = layout == FC_AFFINE_BLOCKED ? p[0] : p[1];
    affine_eval(x, raw, shift, activation, inverse, y, lad);
  }
  // gp may alias p (the slot is rewritten in place): both parameters are read before the first write
  __device__ __forceinline__ void backward(float x, const float* p, float gy, float gl, float& gx, float* gp) const {
    const bool blocked = layout == FC_AFFINE_BLOCKED;
    const float raw = blocked ? p[D_t] : p[0];
    const float shift = blocked ? p[0] : p[1];
    float graw, gshift;
    affine_backward_elem(x, raw, shift, activation, inverse, gy, gl, gx, graw, gshift);
    gp[blocked ? D_t : 0] = graw;
    gp[blocked ? 0 : 1] = gshift;
  }
};

// Forward / inverse, general path.  16 B per element and almost no arithmetic: the kernel lives on loads in flight, so every warp
// handles kU consecutive row groups per step and issues all of their loads (x, raw scale, shift, identity columns)
// before the first dependent instruction.
constexpr int kAffineUnroll = 4;

__global__ void __launch_bounds__(kThreads) affine_forward_kernel(const AffineArgs a) {
  constexpr int kU = kAffineUnroll;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = a.seg, rpw = 32 / seg;
  const int sub = lane / seg, j0 = lane % seg;
  const int64_t groups = (a.B + rpw - 1) / rpw;
  const int64_t steps = (groups + kU - 1) / kU;
  for (int64_t st = (int64_t)blockIdx.x * kWarps + warp; st < steps; st += (int64_t)gridDim.x * kWarps) {
    int64_t row[kU];
    bool ok[kU];
    float lad_acc[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      row[u] = (st * kU + u) * rpw + sub;
      ok[u] = row[u] < a.B;
      lad_acc[u] = 0.f;
    }
    for (int j = j0; j < a.D_t; j += seg) {
      const int col = a.tcols ? __ldg(a.tcols + j) : j;
      float xv[kU], raw[kU], shift[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (ok[u]) {
          xv[u] = __ldcs(a.x + row[u] * a.x_stride + col);
          const float* prow = a.params + row[u] * a.p_stride;
          if (a.layout == FC_AFFINE_BLOCKED) {
            shift[u] = __ldcs(prow + j);
            raw[u] = __ldcs(prow + a.D_t + j);
          } else {
            const float2 rs = __ldcs(reinterpret_cast<const float2*>(prow) + j);
            raw[u] = rs.x;
            shift[u] = rs.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (ok[u]) {
          float yv, lv;
          affine_eval(xv[u], raw[u], shift[u], a.activation, a.inverse, yv, lv);
          __stcs(a.y + row[u] * a.y_stride + col, yv);
          lad_acc[u] += lv;
        }
      }
    }
    for (int i = j0; i < a.n_copy; i += seg) {  // identity columns
      const int col = __ldg(a.ccols + i);
      float v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (ok[u]) v[u] = __ldcs(a.x + row[u] * a.x_stride + col);
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (ok[u]) __stcs(a.y + row[u] * a.y_stride + col, v[u]);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const float tot = seg_reduce_sum(lad_acc[u], seg);
      if (ok[u] && j0 == 0) a.lad[row[u]] = a.accumulate ? a.lad[row[u]] + tot : tot;
    }
  }
}

template <bool kBackward>
__global__ void __launch_bounds__(kThreads) affine_kernel(const AffineArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = a.seg, rpw = 32 / seg;
  const int sub = lane / seg, j0 = lane % seg;
  const int64_t groups = (a.B + rpw - 1) / rpw;
  for (int64_t g = (int64_t)blockIdx.x * kWarps + warp; g < groups; g += (int64_t)gridDim.x * kWarps) {
    const int64_t row = g * rpw + sub;
    const bool row_ok = row < a.B;
    float lad_acc = 0.f;
    if (row_ok) {
      const float* xrow = a.x + row * a.x_stride;
      const float* prow = a.params + row * a.p_stride;
      float* yrow = a.y + row * a.y_stride;
      if (!kBackward) {
        for (int j = j0; j < a.D_t; j += seg) {
          const int col = a.tcols ? a.tcols[j] : j;
          float raw, shift, yv, lv;
          affine_fetch(a, prow, j, raw, shift);
          affine_eval(xrow[col], raw, shift, a.activation, a.inverse, yv, lv);
          yrow[col] = yv;
          lad_acc += lv;
        }
        for (int i = j0; i < a.n_copy; i += seg) yrow[a.ccols[i]] = xrow[a.ccols[i]];
      } else {
        const float* gyrow = a.gy + row * a.gy_stride;
        float* gprow = a.gp + row * a.gp_stride;
        const float gl = a.gl ? a.gl[row] : 0.f;
        for (int j = j0; j < a.D_t; j += seg) {
          const int col = a.tcols ? a.tcols[j] : j;
          float raw, shift, gx, graw, gshift;
          affine_fetch(a, prow, j, raw, shift);
          affine_backward_elem(xrow[col], raw, shift, a.activation, a.inverse, gyrow[col], gl, gx, graw, gshift);
          yrow[col] = gx;
          if (a.layout == FC_AFFINE_BLOCKED) {
            gprow[j] = gshift;
            gprow[a.D_t + j] = graw;
          } else {
            gprow[2 * j] = graw;
            gprow[2 * j + 1] = gshift;
          }
        }
        for (int i = j0; i < a.n_copy; i += seg) yrow[a.ccols[i]] = gyrow[a.ccols[i]];
      }
    }
    if (!kBackward) {
      lad_acc = seg_reduce_sum(lad_acc, seg);
      if (row_ok && j0 == 0) a.lad[row] = a.accumulate ? a.lad[row] + lad_acc : lad_acc;
    }
  }
}

__global__ void __launch_bounds__(kThreads) stdnormal_kernel(const float* __restrict__ z, int64_t z_stride,
                                                             const float* __restrict__ lad, float* __restrict__ out,
                                                             int64_t B, int D, int seg, float log_z) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / seg, sub = lane / seg, j0 = lane % seg;
  const int64_t groups = (B + rpw - 1) / rpw;
  for (int64_t g = (int64_t)blockIdx.x * kWarps + warp; g < groups; g += (int64_t)gridDim.x * kWarps) {
    const int64_t row = g * rpw + sub;
    const bool row_ok = row < B;
    float acc = 0.f;
    if (row_ok) {
      const float* zr = z + row * z_stride;
      for (int j = j0; j < D; j += seg) acc += zr[j] * zr[j];
    }
    acc = seg_reduce_sum(acc, seg);
    if (row_ok && j0 == 0) out[row] = (-0.5f * acc - log_z) + (lad ? lad[row] : 0.f);
  }
}

static int affine_grid(int64_t B, int seg) {
  const int64_t groups = (B + (32 / seg) - 1) / (32 / seg);
  int64_t g = (groups + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)device_info().sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_affine_apply(const float* x, int64_t x_row_stride, const float* params, int64_t params_row_stride,
                               float* y, int64_t y_row_stride, float* logabsdet, int32_t accumulate_logabsdet,
                               int64_t B, int32_t D_t, fc_cols tcols, fc_cols ccols, int32_t layout,
                               int32_t activation, int32_t inverse, void* stream) {
  int rc = check_layer_args(x, params, y, B, D_t, tcols, ccols);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  if (!logabsdet) return FC_ERR_INVALID_ARGUMENT;
  if (layout != FC_AFFINE_BLOCKED && layout != FC_AFFINE_INTERLEAVED) return FC_ERR_INVALID_ARGUMENT;
  if (activation < FC_SCALE_SIGMOID2 || activation > FC_SCALE_SOFTPLUS_EPS) return FC_ERR_INVALID_ARGUMENT;
  if (B == 0) return FC_OK;
  AffineArgs a = {};
  a.x = x; a.params = params; a.y = y; a.lad = logabsdet;
  a.x_stride = x_row_stride; a.p_stride = params_row_stride; a.y_stride = y_row_stride;
  a.B = B; a.D_t = D_t; a.n_copy = ccols.n; a.tcols = tcols.idx; a.ccols = ccols.idx;
  a.accumulate = accumulate_logabsdet; a.layout = layout; a.activation = activation; a.inverse = inverse;
  a.seg = lane_map(D_t).seg;
  {  // fast path: whole contiguous rows through the per-warp TMA ring
    LayerArgs la = {};
    la.x = x; la.params = params; la.y = y; la.lad = logabsdet; la.status = nullptr;
    la.x_stride = x_row_stride; la.p_stride = params_row_stride; la.y_stride = y_row_stride;
    la.B = B; la.D_t = D_t; la.n_copy = ccols.n; la.tcols = tcols.idx; la.ccols = ccols.idx;
    la.accumulate = accumulate_logabsdet;
    AffineOp op = {D_t, layout, activation, inverse};
    const int piped = try_launch_pipelined(la, op, 2, (int)x_row_stride, (cudaStream_t)stream);
    if (piped != 0) return piped < 0 ? piped : FC_OK;
  }
  // interleaved (raw, shift) pairs are read as float2: the parameter rows must keep 8-byte alignment
  const bool pairs_ok = layout == FC_AFFINE_BLOCKED ||
                        ((reinterpret_cast<uintptr_t>(params) & 7) == 0 && (params_row_stride & 1) == 0);
  last_path() = kPathStaged;
  if (pairs_ok) {
    affine_forward_kernel<<<affine_grid((B + kAffineUnroll - 1) / kAffineUnroll, a.seg), kThreads, 0, (cudaStream_t)stream>>>(a);
  } else {
    affine_kernel<false><<<affine_grid(B, a.seg), kThreads, 0, (cudaStream_t)stream>>>(a);
  }
  FC_CHECK_LAUNCH();
  return FC_OK;
}

extern "C" int fc_affine_backward(const float* x, int64_t x_row_stride, const float* params, int64_t params_row_stride,
                                  const float* grad_y, int64_t gy_row_stride, const float* grad_logabsdet,
                                  float* grad_x, int64_t gx_row_stride, float* grad_params, int64_t gp_row_stride,
                                  int64_t B, int32_t D_t, fc_cols tcols, fc_cols ccols, int32_t layout,
                                  int32_t activation, int32_t inverse, void* stream) {
  int rc = check_layer_args(x, params, grad_x, B, D_t, tcols, ccols);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  if (!grad_y || !grad_params) return FC_ERR_INVALID_ARGUMENT;
  if (layout != FC_AFFINE_BLOCKED && layout != FC_AFFINE_INTERLEAVED) return FC_ERR_INVALID_ARGUMENT;
  if (activation < FC_SCALE_SIGMOID2 || activation > FC_SCALE_SOFTPLUS_EPS) return FC_ERR_INVALID_ARGUMENT;
  if (B == 0) return FC_OK;
  AffineArgs a = {};
  a.x = x; a.params = params; a.gy = grad_y; a.gl = grad_logabsdet; a.y = grad_x; a.gp = grad_params;
  a.x_stride = x_row_stride; a.p_stride = params_row_stride; a.y_stride = gx_row_stride;
  a.gy_stride = gy_row_stride; a.gp_stride = gp_row_stride;
  a.B = B; a.D_t = D_t; a.n_copy = ccols.n; a.tcols = tcols.idx; a.ccols = ccols.idx;
  a.layout = layout; a.activation = activation; a.inverse = inverse;
  a.seg = lane_map(D_t).seg;
  {  // fast path: whole contiguous rows through the per-warp TMA ring (bulk loads in, bulk stores out)
    LayerBwdArgs lb = {};
    lb.x = x; lb.params = params; lb.gy = grad_y; lb.gl = grad_logabsdet; lb.gx = grad_x; lb.gp = grad_params;
    lb.x_stride = x_row_stride; lb.p_stride = params_row_stride; lb.gy_stride = gy_row_stride;
    lb.gx_stride = gx_row_stride; lb.gp_stride = gp_row_stride;
    lb.B = B; lb.D_t = D_t; lb.n_copy = ccols.n; lb.tcols = tcols.idx; lb.ccols = ccols.idx;
    AffineOp op = {D_t, layout, activation, inverse};
    const int piped = try_launch_pipelined_backward(lb, op, 2, (int)x_row_stride, (cudaStream_t)stream);
    if (piped != 0) return piped < 0 ? piped : FC_OK;
  }
  last_path() = kPathStaged;
  affine_kernel<true><<<affine_grid(B, a.seg), kThreads, 0, (cudaStream_t)stream>>>(a);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

extern "C" int fc_stdnormal_log_prob(const float* z, int64_t z_row_stride, const float* logabsdet, float* out,
                                     int64_t B, int32_t D, void* stream) {
  if (B < 0 || D < 1) return FC_ERR_INVALID_ARGUMENT;
  if (B == 0) return FC_OK;
  if (!z || !out) return FC_ERR_INVALID_ARGUMENT;
  const int seg = lane_map(D).seg;
  // distributions/normal.py:18-21: log_z = 0.5 * D * log(2 pi) (fp64 buffer, applied to an fp32 tensor)
  const float log_z = (float)(0.5 * (double)D * log(2.0 * 3.14159265358979323846));
  stdnormal_kernel<<<affine_grid(B, seg), kThreads, 0, (cudaStream_t)stream>>>(z, z_row_stride, logabsdet, out, B, D,
                                                                               seg, log_z);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

extern "C" const char* fc_version(void) { return "flowcon_b200 0.1.0 (sm_100a)"; }
extern "C" int fc_built_for_sm(void) { return 100; }
