// fc_staged.cuh — the "staged tile" kernel skeleton shared by the spline and sum-of-sigmoids layers.
//
// Work decomposition:
//   * a CTA owns a tile of R consecutive rows; the tile's parameters are one contiguous block of
//     R * D_t * P floats which is staged global -> shared with 128-bit coalesced loads;
//   * a warp owns row groups of the tile; lane j of a `seg`-lane segment owns feature j of a row
//     (seg = min(32, next_pow2(D_t))), so the P parameters of a lane's element sit at shared-memory word
//     (row*D_t + j) * P: consecutive lanes are P words apart and P (3K-1, 3n+1) is odd for even K / n,
//     i.e. the per-parameter reads are bank-conflict free;
//   * the element math runs in registers (fc_math.cuh), the per-sample log|det J| is a shuffle
//     reduction over the segment (sum_except_batch, flowcon/utils/torchutils.py:25-30), identity
//     columns of a coupling layer are copied by the same warp (coupling.py:96-98);
//   * the backward writes the P parameter gradients over the staged parameters in place and streams
//     the tile back with 128-bit stores.
//
// An `Op` provides:  int P() ; void eval(x, p, y&, lad&, status&) ; void backward(x, p, gy, gl, gx&, gp)
#pragma once
#include "fc_common.cuh"

namespace fc {

struct LayerArgs {
  const float* x;
  const float* params;
  float* y;
  float* lad;
  int32_t* status;
  int64_t x_stride, p_stride, y_stride;
  int64_t B;
  int D_t, n_copy;
  const int32_t* tcols;
  const int32_t* ccols;
  int accumulate;
  int tile_rows, seg;
  int64_t num_tiles;
};

struct LayerBwdArgs {
  const float* x;
  const float* params;
  const float* gy;
  const float* gl;
  float* gx;
  float* gp;
  int64_t x_stride, p_stride, gy_stride, gx_stride, gp_stride;
  int64_t B;
  int D_t, n_copy;
  const int32_t* tcols;
  const int32_t* ccols;
  int tile_rows, seg;
  int64_t num_tiles;
};

__device__ __forceinline__ void stage_tile_in(float* smem, const float* params, int64_t p_stride, int64_t row0,
                                              int rows, int row_floats) {
  if (p_stride == row_floats) {
    stage_in(smem, params + row0 * p_stride, (int64_t)rows * row_floats);
  } else {
    for (int r = 0; r < rows; ++r) stage_in(smem + (int64_t)r * row_floats, params + (row0 + r) * p_stride, row_floats);
  }
}

// kStage: parameters staged through shared memory (true) or read straight from global (rows too long).
template <class Op, bool kStage>
__global__ void __launch_bounds__(kThreads) staged_apply_kernel(const LayerArgs a, const Op op) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = a.seg, rpw = 32 / seg;
  const int sub = lane / seg, j0 = lane % seg;
  const int P = op.P();
  const int row_floats = a.D_t * P;
  unsigned status = 0;

  for (int64_t tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * a.tile_rows;
    const int rows = (int)min((int64_t)a.tile_rows, a.B - row0);
    if (kStage) {
      stage_tile_in(smem, a.params, a.p_stride, row0, rows, row_floats);
      __syncthreads();
    }
    const int groups = (rows + rpw - 1) / rpw;
    for (int g = warp; g < groups; g += kWarps) {
      const int r = g * rpw + sub;
      const bool row_ok = r < rows;
      const int64_t row = row0 + r;
      float lad_acc = 0.f;
      if (row_ok) {
        const float* xrow = a.x + row * a.x_stride;
        float* yrow = a.y + row * a.y_stride;
        const float* prow = kStage ? smem + (int64_t)r * row_floats : a.params + row * a.p_stride;
        for (int j = j0; j < a.D_t; j += seg) {
          const int col = a.tcols ? a.tcols[j] : j;
          float yv, lv;
          op.eval(xrow[col], prow + (int64_t)j * P, yv, lv, status);
          yrow[col] = yv;
          lad_acc += lv;
        }
        for (int i = j0; i < a.n_copy; i += seg) {
          const int col = a.ccols[i];
          yrow[col] = xrow[col];
        }
      }
      lad_acc = seg_reduce_sum(lad_acc, seg);
      if (row_ok && j0 == 0) a.lad[row] = a.accumulate ? a.lad[row] + lad_acc : lad_acc;
    }
    if (kStage) __syncthreads();
  }
  if (status && a.status) atomicOr(a.status, (int)status);
}

template <class Op, bool kStage>
__global__ void __launch_bounds__(kThreads) staged_backward_kernel(const LayerBwdArgs a, const Op op) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = a.seg, rpw = 32 / seg;
  const int sub = lane / seg, j0 = lane % seg;
  const int P = op.P();
  const int row_floats = a.D_t * P;

  for (int64_t tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * a.tile_rows;
    const int rows = (int)min((int64_t)a.tile_rows, a.B - row0);
    if (kStage) {
      stage_tile_in(smem, a.params, a.p_stride, row0, rows, row_floats);
      __syncthreads();
    }
    const int groups = (rows + rpw - 1) / rpw;
    for (int g = warp; g < groups; g += kWarps) {
      const int r = g * rpw + sub;
      if (r >= rows) continue;
      const int64_t row = row0 + r;
      const float* xrow = a.x + row * a.x_stride;
      const float* gyrow = a.gy + row * a.gy_stride;
      float* gxrow = a.gx + row * a.gx_stride;
      const float gl = a.gl ? a.gl[row] : 0.f;
      for (int j = j0; j < a.D_t; j += seg) {
        const int col = a.tcols ? a.tcols[j] : j;
        float gxv;
        if (kStage) {
          float* slot = smem + (int64_t)r * row_floats + (int64_t)j * P;  // gradients overwrite the staged params
          op.backward(xrow[col], slot, gyrow[col], gl, gxv, slot);
        } else {
          op.backward(xrow[col], a.params + row * a.p_stride + (int64_t)j * P, gyrow[col], gl, gxv,
                      a.gp + row * a.gp_stride + (int64_t)j * P);
        }
        gxrow[col] = gxv;
      }
      for (int i = j0; i < a.n_copy; i += seg) {
        const int col = a.ccols[i];
        gxrow[col] = gyrow[col];
      }
    }
    if (kStage) {
      __syncthreads();
      if (a.gp_stride == row_floats) {
        stage_out(a.gp + row0 * a.gp_stride, smem, (int64_t)rows * row_floats);
      } else {
        for (int r = 0; r < rows; ++r)
          stage_out(a.gp + (row0 + r) * a.gp_stride, smem + (int64_t)r * row_floats, row_floats);
      }
      __syncthreads();
    }
  }
}

template <typename Kern>
inline int prepare_kernel(Kern kern, size_t smem) {
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return FC_ERR_CUDA;
  }
  return FC_OK;
}

inline int grid_for(int64_t tiles, size_t smem) {
  const DeviceInfo& d = device_info();
  int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;  // 8 CTAs x 256 threads = 2048 resident threads
  int64_t g = (int64_t)d.sm_count * per_sm;
  if (g > tiles) g = tiles;
  if (g < 1) g = 1;
  return (int)g;
}

// Fill tiling fields; returns the dynamic shared-memory size (0 => unstaged kernel).
template <class Args>
inline size_t plan_tiles(Args& a, int P) {
  const int row_floats = a.D_t * P;
  const LaneMap lm = lane_map(a.D_t);
  a.seg = lm.seg;
  const bool stage = (int64_t)row_floats * 4 <= kSmemMaxBytes;
  a.tile_rows = stage ? tile_rows(row_floats, lm, a.B) : kWarps * lm.rows_per_warp;
  a.num_tiles = (a.B + a.tile_rows - 1) / a.tile_rows;
  return stage ? (size_t)a.tile_rows * row_floats * 4 : 0;
}

// Which skeleton the last element-wise layer call of this thread launched (fc_elementwise_last_path: tests, profiles)
enum { kPathStaged = 0, kPathWarpRing = 1, kPathTileRing = 2 };
inline int& last_path() {
  static thread_local int p = -1;
  return p;
}

template <class Op>
inline int launch_apply(const LayerArgs& a, const Op& op, size_t smem, cudaStream_t st) {
  last_path() = kPathStaged;
  const int grid = grid_for(a.num_tiles, smem);
  if (smem) {
    if (prepare_kernel(staged_apply_kernel<Op, true>, smem) != FC_OK) return FC_ERR_CUDA;
    staged_apply_kernel<Op, true><<<grid, kThreads, smem, st>>>(a, op);
  } else {
    staged_apply_kernel<Op, false><<<grid, kThreads, 0, st>>>(a, op);
  }
  FC_CHECK_LAUNCH();
  return FC_OK;
}

template <class Op>
inline int launch_backward(const LayerBwdArgs& a, const Op& op, size_t smem, cudaStream_t st) {
  last_path() = kPathStaged;
  const int grid = grid_for(a.num_tiles, smem);
  if (smem) {
    if (prepare_kernel(staged_backward_kernel<Op, true>, smem) != FC_OK) return FC_ERR_CUDA;
    staged_backward_kernel<Op, true><<<grid, kThreads, smem, st>>>(a, op);
  } else {
    staged_backward_kernel<Op, false><<<grid, kThreads, 0, st>>>(a, op);
  }
  FC_CHECK_LAUNCH();
  return FC_OK;
}

inline int check_layer_args(const void* x, const void* params, const void* out, int64_t B, int32_t D_t, fc_cols tcols,
                            fc_cols ccols) {
  if (B < 0 || D_t < 1) return FC_ERR_INVALID_ARGUMENT;
  if (B > 0 && (!x || !params || !out)) return FC_ERR_INVALID_ARGUMENT;  // empty batches carry null pointers
  if (tcols.idx && tcols.n != D_t) return FC_ERR_INVALID_ARGUMENT;
  if (ccols.n < 0 || (ccols.n > 0 && !ccols.idx)) return FC_ERR_INVALID_ARGUMENT;
  return FC_OK;
}

}  // namespace fc
