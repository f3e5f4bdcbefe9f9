// fc_actnorm.cu — activation normalisation as a flow layer.
//
// Replaces ActNorm.forward / inverse (flowcon/transforms/normalization.py:144-218, 2-D inputs): per-feature
//   forward  y = exp(log_scale) * x + shift,      logabsdet = +sum(log_scale)
//   inverse  y = (x - shift) * exp(-log_scale),   logabsdet = -sum(log_scale)
// and their gradients.  8 bytes per element forward: HBM-bound; one grid-stride pass with 128-bit accesses, the D scales /
// shifts in shared memory.  The parameter gradients are reductions over the batch: every CTA reduces its row block per
// column in a fixed order into a workspace, a second kernel sums the blocks in order (deterministic: replays of a
// captured training step stay bitwise equal).
#include "fc_common.cuh"

namespace fc {

constexpr int kActThreads = 256;

struct ActArgs {
  const float* x;
  const float* log_scale;
  const float* shift;
  float* y;
  float* lad;
  int64_t xs, ys, B;
  int D, inverse, accumulate;
};

__global__ void __launch_bounds__(kActThreads) actnorm_apply_kernel(const ActArgs a) {
  extern __shared__ float sm[];  // [D] scale, [D] shift
  float* sc = sm;
  float* sh = sm + a.D;
  __shared__ float total_s;
  float part = 0.f;
  for (int d = threadIdx.x; d < a.D; d += kActThreads) {
    const float ls = __ldg(a.log_scale + d);
    sc[d] = expf(a.inverse ? -ls : ls);
    sh[d] = __ldg(a.shift + d);
    part += ls;
  }
  // sum(log_scale) in a fixed order: per-thread partial sums, then thread 0 adds the 256 partials
  __shared__ float parts[kActThreads];
  parts[threadIdx.x] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kActThreads; ++i) t += parts[i];
    total_s = a.inverse ? -t : t;
  }
  __syncthreads();
  const float total = total_s;
  const int D = a.D;
  const bool vec = (D & 3) == 0 && (a.xs & 3) == 0 && (a.ys & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.y)) & 15) == 0;
  const int64_t tid0 = (int64_t)blockIdx.x * kActThreads + threadIdx.x, nthr = (int64_t)gridDim.x * kActThreads;
  if (vec) {
    const int D4 = D >> 2;
    const int64_t n4 = a.B * D4;
    for (int64_t i = tid0; i < n4; i += nthr) {
      const int64_t r = i / D4;
      const int d = (int)(i - r * D4) << 2;
      const float4 v = __ldcs(reinterpret_cast<const float4*>(a.x + r * a.xs + d));
      float4 o;
      if (a.inverse) {
        o.x = (v.x - sh[d]) * sc[d], o.y = (v.y - sh[d + 1]) * sc[d + 1];
        o.z = (v.z - sh[d + 2]) * sc[d + 2], o.w = (v.w - sh[d + 3]) * sc[d + 3];
      } else {
        o.x = fmaf(sc[d], v.x, sh[d]), o.y = fmaf(sc[d + 1], v.y, sh[d + 1]);
        o.z = fmaf(sc[d + 2], v.z, sh[d + 2]), o.w = fmaf(sc[d + 3], v.w, sh[d + 3]);
      }
      __stcs(reinterpret_cast<float4*>(a.y + r * a.ys + d), o);
    }
  } else {
    const int64_t n = a.B * D;
    for (int64_t i = tid0; i < n; i += nthr) {
      const int64_t r = i / D;
      const int d = (int)(i - r * D);
      const float v = a.x[r * a.xs + d];
      a.y[r * a.ys + d] = a.inverse ? (v - sh[d]) * sc[d] : fmaf(sc[d], v, sh[d]);
    }
  }
  for (int64_t r = tid0; r < a.B; r += nthr) a.lad[r] = a.accumulate ? a.lad[r] + total : total;
}

struct ActBwdArgs {
  const float* x;
  const float* log_scale;
  const float* shift;
  const float* gy;
  const float* gl;  // may be null
  float* gx;
  float* ws;  // [blocks][2][D] column sums, then [blocks] sums of gl
  int64_t xs, gys, gxs, B;
  int D, inverse, blocks;
  int64_t rows_per_block;
};

// CTA g: rows [g R, (g + 1) R).  Threads (tx = column lane, ty = row lane); a thread owns the V-wide column groups
// tx, tx + CW, ... (V = 4: 128-bit accesses when D and the strides allow) and walks the CTA's rows two at a time, so that
// four 16-byte loads are in flight per thread.
template <int V>
__global__ void __launch_bounds__(kActThreads) actnorm_backward_kernel(const ActBwdArgs a, int cw) {
  extern __shared__ float sm[];  // [D] scale, [D] shift, then [ry][2][D] partial sums
  const int D = a.D;
  float* sc = sm;
  float* sh = sm + D;
  float* red = sm + 2 * D;
  for (int d = threadIdx.x; d < D; d += kActThreads) {
    const float ls = __ldg(a.log_scale + d);
    sc[d] = expf(a.inverse ? -ls : ls);
    sh[d] = __ldg(a.shift + d);
  }
  __syncthreads();
  const int ry = kActThreads / cw, tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const int64_t r_lo = (int64_t)blockIdx.x * a.rows_per_block;
  int64_t r_hi = r_lo + a.rows_per_block;
  r_hi = r_hi < a.B ? r_hi : a.B;
  constexpr int kMaxGroups = 16 / V;  // column groups per thread (D <= 16 * 256 scalar, 4 * 4 * 256 vector)
  float s_ls[kMaxGroups][V], s_sh[kMaxGroups][V];
#pragma unroll
  for (int c = 0; c < kMaxGroups; ++c)
#pragma unroll
    for (int v = 0; v < V; ++v) s_ls[c][v] = s_sh[c][v] = 0.f;
  auto one = [&](int c, int d, const float* g, const float* xv, float* gx) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float s = sc[d + v];
      gx[v] = g[v] * s;
      if (a.inverse) {
        s_ls[c][v] -= g[v] * ((xv[v] - sh[d + v]) * s);  // d y / d log_scale = -y
        s_sh[c][v] -= g[v] * s;
      } else {
        s_ls[c][v] += g[v] * (xv[v] * s);
        s_sh[c][v] += g[v];
      }
    }
  };
  for (int64_t r = r_lo + ty; r < r_hi; r += 2 * ry) {
    const bool two = r + ry < r_hi;
#pragma unroll
    for (int c = 0; c < kMaxGroups; ++c) {
      const int d = (tx + c * cw) * V;
      if (d < D) {
        float g0[V], x0[V], g1[V], x1[V], o0[V], o1[V];
        if constexpr (V == 4) {
          *reinterpret_cast<float4*>(g0) = __ldcs(reinterpret_cast<const float4*>(a.gy + r * a.gys + d));
          *reinterpret_cast<float4*>(x0) = __ldcs(reinterpret_cast<const float4*>(a.x + r * a.xs + d));
          if (two) {
            *reinterpret_cast<float4*>(g1) = __ldcs(reinterpret_cast<const float4*>(a.gy + (r + ry) * a.gys + d));
            *reinterpret_cast<float4*>(x1) = __ldcs(reinterpret_cast<const float4*>(a.x + (r + ry) * a.xs + d));
          }
        } else {
          g0[0] = a.gy[r * a.gys + d], x0[0] = a.x[r * a.xs + d];
          if (two) g1[0] = a.gy[(r + ry) * a.gys + d], x1[0] = a.x[(r + ry) * a.xs + d];
        }
        one(c, d, g0, x0, o0);
        if constexpr (V == 4) {
          __stcs(reinterpret_cast<float4*>(a.gx + r * a.gxs + d), *reinterpret_cast<float4*>(o0));
        } else {
          a.gx[r * a.gxs + d] = o0[0];
        }
        if (two) {
          one(c, d, g1, x1, o1);
          if constexpr (V == 4) {
            __stcs(reinterpret_cast<float4*>(a.gx + (r + ry) * a.gxs + d), *reinterpret_cast<float4*>(o1));
          } else {
            a.gx[(r + ry) * a.gxs + d] = o1[0];
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kMaxGroups; ++c) {
    const int d = (tx + c * cw) * V;
    if (d < D) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        red[(ty * 2 + 0) * D + d + v] = s_ls[c][v];
        red[(ty * 2 + 1) * D + d + v] = s_sh[c][v];
      }
    }
  }
  __syncthreads();
  float* out = a.ws + (size_t)blockIdx.x * 2 * D;
  for (int d = threadIdx.x; d < D; d += kActThreads) {
    float t0 = 0.f, t1 = 0.f;
    for (int y = 0; y < ry; ++y) {
      t0 += red[(y * 2 + 0) * D + d];
      t1 += red[(y * 2 + 1) * D + d];
    }
    out[d] = t0;
    out[D + d] = t1;
  }
  // sum of grad_logabsdet over the CTA's rows (each log_scale[d] receives +-sum_b gl[b])
  if (a.gl) {
    __shared__ float gparts[kActThreads];
    float t = 0.f;
    for (int64_t r = r_lo + threadIdx.x; r < r_hi; r += kActThreads) t += a.gl[r];
    gparts[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < kActThreads; ++i) s += gparts[i];
      a.ws[(size_t)a.blocks * 2 * D + blockIdx.x] = s;
    }
  }
}

// Second stage: one CTA per column sums the per-CTA partials of the first stage in a fixed order (strided partial sums per
// thread, then a shuffle / shared-memory tree), so the result does not depend on the launch.  (One thread per column
// walking all partials serially cost 50-90 us of the 0.79 ms of a 4 M-row call.)
constexpr int kFinishThreads = 128;
__global__ void __launch_bounds__(kFinishThreads) actnorm_finish_kernel(const float* ws, int blocks, int D, int has_gl,
                                                                        int inverse, float* g_ls, float* g_sh) {
  const int d = blockIdx.x;
  float t0 = 0.f, t1 = 0.f, g = 0.f;
  for (int b = threadIdx.x; b < blocks; b += kFinishThreads) {
    t0 += ws[(size_t)b * 2 * D + d];
    t1 += ws[(size_t)b * 2 * D + D + d];
    if (has_gl) g += ws[(size_t)blocks * 2 * D + b];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    g += __shfl_xor_sync(0xffffffffu, g, o);
  }
  __shared__ float parts[3][kFinishThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    parts[0][threadIdx.x >> 5] = t0;
    parts[1][threadIdx.x >> 5] = t1;
    parts[2][threadIdx.x >> 5] = g;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a0 = 0.f, a1 = 0.f, ag = 0.f;
    for (int w = 0; w < kFinishThreads / 32; ++w) {
      a0 += parts[0][w];
      a1 += parts[1][w];
      ag += parts[2][w];
    }
    g_ls[d] = a0 + (inverse ? -ag : ag);
    g_sh[d] = a1;
  }
}

static int act_blocks(int64_t B) {
  const int64_t want = (B + 127) / 128;  // at least 128 rows per CTA
  const int cap = 8 * device_info().sm_count;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace fc

using namespace fc;

extern "C" int fc_actnorm_apply(const float* x, int64_t x_row_stride, const float* log_scale, const float* shift, float* y,
                                int64_t y_row_stride, float* logabsdet, int32_t accumulate_logabsdet, int64_t B, int32_t D,
                                int32_t inverse, void* stream) {
  if (B < 0 || D < 1 || !log_scale || !shift) return FC_ERR_INVALID_ARGUMENT;
  if (B == 0) return FC_OK;
  if (!x || !y || !logabsdet || x_row_stride < D || y_row_stride < D) return FC_ERR_INVALID_ARGUMENT;
  if (D > 8192) return FC_ERR_UNSUPPORTED;
  ActArgs a;
  a.x = x; a.log_scale = log_scale; a.shift = shift; a.y = y; a.lad = logabsdet;
  a.xs = x_row_stride; a.ys = y_row_stride; a.B = B; a.D = D; a.inverse = inverse; a.accumulate = accumulate_logabsdet;
  const int64_t work = (B * D + 4 * kActThreads - 1) / (4 * kActThreads);
  const int cap = 8 * device_info().sm_count;
  const int grid = (int)(work < 1 ? 1 : (work > cap ? cap : work));
  actnorm_apply_kernel<<<grid, kActThreads, 2 * D * sizeof(float), (cudaStream_t)stream>>>(a);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

extern "C" int64_t fc_actnorm_workspace_floats(int64_t B, int32_t D) {
  if (B < 0 || D < 1) return FC_ERR_INVALID_ARGUMENT;
  const int blocks = act_blocks(B);
  return (int64_t)blocks * 2 * D + blocks;
}

extern "C" int fc_actnorm_backward(const float* x, int64_t x_row_stride, const float* log_scale, const float* shift,
                                   const float* grad_y, int64_t gy_row_stride, const float* grad_logabsdet, float* grad_x,
                                   int64_t gx_row_stride, float* grad_log_scale, float* grad_shift, float* workspace,
                                   int64_t B, int32_t D, int32_t inverse, void* stream) {
  if (B < 0 || D < 1 || !log_scale || !shift || !grad_log_scale || !grad_shift) return FC_ERR_INVALID_ARGUMENT;
  if (B > 0 && (!x || !grad_y || !grad_x || !workspace)) return FC_ERR_INVALID_ARGUMENT;
  if (D > 4096) return FC_ERR_UNSUPPORTED;  // 16 columns per thread
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    if (cudaMemsetAsync(grad_log_scale, 0, sizeof(float) * D, st) != cudaSuccess) return FC_ERR_CUDA;
    if (cudaMemsetAsync(grad_shift, 0, sizeof(float) * D, st) != cudaSuccess) return FC_ERR_CUDA;
    return FC_OK;
  }
  ActBwdArgs a;
  a.x = x; a.log_scale = log_scale; a.shift = shift; a.gy = grad_y; a.gl = grad_logabsdet; a.gx = grad_x; a.ws = workspace;
  a.xs = x_row_stride; a.gys = gy_row_stride; a.gxs = gx_row_stride; a.B = B; a.D = D; a.inverse = inverse;
  a.blocks = act_blocks(B);
  a.rows_per_block = (B + a.blocks - 1) / a.blocks;
  const bool vec = (D & 3) == 0 && (x_row_stride & 3) == 0 && (gy_row_stride & 3) == 0 && (gx_row_stride & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_y) | reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0;
  int cw = next_pow2(vec ? D / 4 : D);
  cw = cw > kActThreads ? kActThreads : cw;
  const int ry = kActThreads / cw;
  const size_t smem = sizeof(float) * ((size_t)2 * D + (size_t)ry * 2 * D);
  auto kern = vec ? actnorm_backward_kernel<4> : actnorm_backward_kernel<1>;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return FC_ERR_CUDA;
  kern<<<a.blocks, kActThreads, smem, st>>>(a, cw);
  actnorm_finish_kernel<<<D, kFinishThreads, 0, st>>>(workspace, a.blocks, D, grad_logabsdet != nullptr, inverse,
                                                      grad_log_scale, grad_shift);
  FC_CHECK_LAUNCH();
  return FC_OK;
}
