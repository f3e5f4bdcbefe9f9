// fc_pipeline.cuh — TMA-pipelined persistent forward/inverse layer kernel (the fast path of fc_*_apply).
//
// Every warp runs its OWN software pipeline, so there is no CTA-wide synchronisation at all:
//   * a warp owns a ring of S slots in shared memory; a slot holds `slot_rows` consecutive rows:
//     their parameters (slot_rows * D_t * P floats, one contiguous block of global memory) and their inputs
//     (slot_rows * D floats, contiguous too);
//   * lane 0 fills a slot with two 1-D TMA bulk copies (cp.async.bulk.shared.global, SASS UBLKCP) that
//     complete on the slot's mbarrier (expect_tx = bytes); all lanes wait on the barrier's phase parity;
//   * the warp evaluates the bijection from shared memory (lane <-> feature mapping and conflict-free
//     parameter reads as in fc_staged.cuh), writes the transformed values over the inputs IN the slot, reduces
//     the per-sample log|det J| with shuffles, then streams the finished rows (identity columns included:
//     they were staged with the inputs) to global memory with 128-bit coalesced stores;
//   * lane 0 immediately re-arms the slot with the rows S groups ahead (fence.proxy.async orders the warp's
//     generic-proxy accesses to the slot before the async-proxy refill).
// Row groups are dealt round-robin over all warps of the grid, so neighbouring warps stream neighbouring
// DRAM pages.  HBM traffic = algorithmic bytes: params + x read once, y + logabsdet written once.
//
// Requirements (checked by the host, else the staged kernel of fc_staged.cuh runs): contiguous params, x and
// y rows (stride == width), 16-byte aligned bases, row sizes that are multiples of 16 bytes.
#pragma once
#include "fc_staged.cuh"

namespace fc {

struct PipeArgs {
  LayerArgs a;
  int D;           // full row width of x / y
  int slot_rows;   // rows per slot (multiple of 32/seg)
  int stages;      // slots per warp
  int warps;       // warps per CTA
  int64_t num_groups;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Words between the parameter blocks of two neighbouring features: P for the spline / sum-of-sigmoids layouts; an
// Op may override it (blocked affine parameters [shift(D_t) | raw_scale(D_t)]: 1, the Op reads prow[0] and prow[D_t]).
template <class Op>
__device__ __forceinline__ auto feature_stride(const Op& op, int) -> decltype(op.feature_stride()) {
  return op.feature_stride();
}
template <class Op>
__device__ __forceinline__ int feature_stride(const Op& op, long) {
  return op.P();
}

// Second __launch_bounds__ argument of the forward ring.  0 (unspecified) lets ptxas aim for two resident CTAs, i.e. at
// most 64 registers per thread, which the HBM-bound spline / affine rings want; an Op that is bound by arithmetic sets
// kMinBlocks = 1 to get the full 128 registers instead of spills.
template <class Op, class = void>
struct PipeMinBlocks {
  static constexpr int value = 0;
};
template <class Op>
struct PipeMinBlocks<Op, decltype((void)Op::kMinBlocks)> {
  static constexpr int value = Op::kMinBlocks;
};

// kSimple: D_t == 32 (one feature per lane, one row per warp pass): the feature loop and the segment logic fold away.
template <class Op, bool kSimple>
__global__ void __launch_bounds__(512, PipeMinBlocks<Op>::value) pipelined_apply_kernel(const PipeArgs pa, const Op op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const LayerArgs& a = pa.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = kSimple ? 32 : a.seg, rpw = 32 / seg;
  const int sub = lane / seg, j0 = lane % seg;
  const int P = op.P();
  const int D = pa.D, D_t = kSimple ? 32 : a.D_t, S = pa.stages, R = pa.slot_rows;
  const int row_floats = D_t * P;
  const int slot_floats = R * (row_floats + D);
  const int32_t* __restrict__ tcols = a.tcols;
  const int accumulate = a.accumulate;
  const int col0 = (j0 < D_t) ? (tcols ? __ldg(tcols + j0) : j0) : 0;  // this lane's first column
  unsigned status = 0;

  // per-warp carve-up: [S slots][S mbarriers]; everything below is addressed with 32-bit shared offsets
  float* const wslots = reinterpret_cast<float*>(smem_raw) + warp * S * slot_floats;
  const uint32_t wslots_s = smem_u32(wslots);
  const uint32_t bars_s =
      smem_u32(smem_raw + (size_t)pa.warps * S * slot_floats * sizeof(float)) + (uint32_t)warp * S * 8u;
  const uint32_t slot_bytes = (uint32_t)slot_floats * 4u;
  const uint32_t xoff_bytes = (uint32_t)R * row_floats * 4u;  // inputs sit behind the parameters in a slot

  const int64_t gstride = (int64_t)gridDim.x * pa.warps;
  const int64_t g0 = (int64_t)blockIdx.x * pa.warps + warp;
  const int64_t B = a.B;

  // running state of the producer side (lane 0): next group to fetch and its global addresses
  int64_t fetch_row = g0 * R;
  const int64_t row_step = gstride * R;
  const float* fetch_p = a.params + fetch_row * row_floats;
  const float* fetch_x = a.x + fetch_row * D;
  const int64_t step_p = row_step * row_floats, step_x = row_step * D;

  auto issue = [&](int slot) {  // lane 0 only; fetches the group at fetch_row into `slot`, then advances
    const int rows = (int)min((int64_t)R, B - fetch_row);
    const uint32_t dst = wslots_s + (uint32_t)slot * slot_bytes;
    const uint32_t bar = bars_s + (uint32_t)slot * 8u;
    const uint32_t pbytes = (uint32_t)(rows * row_floats) * 4u, xbytes = (uint32_t)(rows * D) * 4u;
    mbar_expect_tx(bar, pbytes + xbytes);
    bulk_g2s(dst, fetch_p, pbytes, bar);
    bulk_g2s(dst + xoff_bytes, fetch_x, xbytes, bar);
    fetch_row += row_step;
    fetch_p += step_p;
    fetch_x += step_x;
  };

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars_s + (uint32_t)s * 8u, 1);
    fence_mbar_init();
    fence_proxy_async();
    for (int s = 0; s < S; ++s)
      if (fetch_row < B) issue(s);
  }
  __syncwarp();

  int slot = 0;
  uint32_t parity = 0;
  int64_t row0 = g0 * R;
  float* yout = a.y + row0 * D;
  float* ladout = a.lad + row0;
  const int64_t step_y = row_step * D;
  for (; row0 < B; row0 += row_step, yout += step_y, ladout += row_step) {
    mbar_wait(bars_s + (uint32_t)slot * 8u, parity);
    const int rows = (int)min((int64_t)R, B - row0);
    const float* sp = wslots + slot * slot_floats;
    float* sx = wslots + slot * slot_floats + R * row_floats;
    float* xrow = sx + sub * D;                       // strength-reduced row pointers
    const int fstride = feature_stride(op, 0);
    const float* prow = sp + sub * row_floats + j0 * fstride;
    const int xstep = rpw * D, pstep = rpw * row_floats, jstep = seg * fstride;
    for (int r = sub; r < rows + sub; r += rpw, xrow += xstep, prow += pstep) {  // uniform trip count
      const bool row_ok = r < rows;
      float lad_acc = 0.f;
      if (kSimple) {
        float yv;
        op.eval(xrow[col0], prow, yv, lad_acc, status);
        xrow[col0] = yv;
      } else if (row_ok) {
        const float* pj = prow;
        for (int j = j0; j < D_t; j += seg, pj += jstep) {
          const int col = (j == j0) ? col0 : (tcols ? __ldg(tcols + j) : j);
          float yv, lv;
          op.eval(xrow[col], pj, yv, lv, status);
          xrow[col] = yv;  // compose the output row in place; identity columns are already there
          lad_acc += lv;
        }
      }
      lad_acc = seg_reduce_sum(lad_acc, seg);
      if (row_ok && j0 == 0) ladout[r] = accumulate ? ladout[r] + lad_acc : lad_acc;
    }
    __syncwarp();
    {  // finished rows -> global, 128-bit coalesced
      const int n4 = (rows * D) >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(sx);
      float4* d4 = reinterpret_cast<float4*>(yout);
      if (n4 == 32) {
        d4[lane] = s4[lane];
      } else {
        for (int i = lane; i < n4; i += 32) d4[i] = s4[i];
      }
    }
    __syncwarp();
    if (lane == 0 && fetch_row < B) {
      fence_proxy_async();
      issue(slot);
    }
    if (++slot == S) {
      slot = 0;
      parity ^= 1u;
    }
  }
  if (status && a.status) atomicOr(a.status, (int)status);
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Try the pipelined kernel; returns 1 if it was launched, 0 if the call does not qualify, <0 on error.
template <class Op>
inline int try_launch_pipelined(const LayerArgs& a, const Op& op, int P, int D, cudaStream_t st) {
  if (env_int("FC_PIPE", 1) == 0) return 0;
  const int64_t row_floats = (int64_t)a.D_t * P;
  const bool aligned = ((reinterpret_cast<uintptr_t>(a.params) | reinterpret_cast<uintptr_t>(a.x) |
                         reinterpret_cast<uintptr_t>(a.y)) & 15) == 0;
  if (!aligned || a.p_stride != row_floats || a.x_stride != D || a.y_stride != D) return 0;
  if (a.D_t + a.n_copy != D) return 0;  // whole rows are streamed: the column lists must cover the row
  if ((row_floats * 4) % 16 != 0 || (D * 4) % 16 != 0) return 0;
  const DeviceInfo& dev = device_info();
  const LaneMap lm = lane_map(a.D_t);
  const int64_t row_bytes = (row_floats + D) * 4;
  // slot: at least one pass of the warp; short rows (affine: 512 B) are grouped up to ~2 KB — small slots keep the
  // shared-memory footprint low enough for 2-3 CTAs of 16 warps per SM, which is what hides the latency (measured on
  // affine coupling rows: 2 KB slots 85 % of the copy peak, 4 KB 68 %, 8 KB 63 %)
  int slot_rows = lm.rows_per_warp;
  const int target = env_int("FC_PIPE_SLOT_BYTES", 2048);
  while ((int64_t)(slot_rows * 2) * row_bytes <= target) slot_rows *= 2;
  slot_rows = env_int("FC_PIPE_SLOT_ROWS", slot_rows);
  int stages = env_int("FC_PIPE_STAGES", 2);
  int warps = env_int("FC_PIPE_WARPS", 16);
  int ctas_per_sm = env_int("FC_PIPE_CTAS", slot_rows * row_bytes <= 2048 ? 3 : 2);
  const int64_t slot_bytes = slot_rows * row_bytes;
  if (slot_bytes > 96 * 1024) return 0;  // rows too long for a per-warp ring
  auto smem_need = [&](int w, int s) { return (int64_t)w * s * (slot_bytes + 8) + 128; };
  const int64_t sm_budget = 224 * 1024;  // 228 KB per SM minus the per-CTA reservation
  while ((int64_t)ctas_per_sm * (smem_need(warps, stages) + 1024) > sm_budget) {
    if (stages > 2) --stages;
    else if (ctas_per_sm > 1) --ctas_per_sm;  // keep 16 warps per CTA before giving up the second CTA
    else if (warps > 1) --warps;
    else return 0;
  }
  if (warps > 16) warps = 16;
  PipeArgs pa;
  pa.a = a;
  pa.a.seg = lm.seg;
  pa.D = D;
  pa.slot_rows = slot_rows;
  pa.stages = stages;
  pa.warps = warps;
  pa.num_groups = (a.B + slot_rows - 1) / slot_rows;
  const size_t smem = (size_t)smem_need(warps, stages);
  int64_t grid = (int64_t)dev.sm_count * ctas_per_sm;
  const int64_t need = (pa.num_groups + warps - 1) / warps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (a.D_t == 32) {
    if (prepare_kernel(pipelined_apply_kernel<Op, true>, smem) != FC_OK) return FC_ERR_CUDA;
    pipelined_apply_kernel<Op, true><<<(int)grid, warps * 32, smem, st>>>(pa, op);
  } else {
    if (prepare_kernel(pipelined_apply_kernel<Op, false>, smem) != FC_OK) return FC_ERR_CUDA;
    pipelined_apply_kernel<Op, false><<<(int)grid, warps * 32, smem, st>>>(pa, op);
  }
  if (cudaGetLastError() != cudaSuccess) return FC_ERR_CUDA;
  return 1;
}

// ------------------------------------------------------------------------------------------------------------
// Backward on the same per-warp ring.  A slot holds the parameters, the inputs and the upstream gradients of a few
// consecutive rows (three 1-D TMA bulk loads).  The element adjoint writes the P parameter gradients OVER the
// parameters in the slot and the input gradient over the inputs; the two finished blocks leave with two TMA bulk
// stores (cp.async.bulk.global.shared::cta), so the warp issues no global load / store of its own except the per-row
// grad_logabsdet scalar.  HBM traffic = algorithmic: x, params, grad_y read once; grad_x, grad_params written once.
// ------------------------------------------------------------------------------------------------------------
struct PipeBwdArgs {
  LayerBwdArgs a;
  int D;
  int slot_rows, stages, warps;
};

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <class Op>
__global__ void __launch_bounds__(512) pipelined_backward_kernel(const PipeBwdArgs pa, const Op op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const LayerBwdArgs& a = pa.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = a.seg, rpw = 32 / seg;
  const int sub = lane / seg, j0 = lane % seg;
  const int P = op.P();
  const int D = pa.D, D_t = a.D_t, S = pa.stages, R = pa.slot_rows;
  const int row_floats = D_t * P;
  const int slot_floats = R * (row_floats + 2 * D);
  const int32_t* __restrict__ tcols = a.tcols;
  const int32_t* __restrict__ ccols = a.ccols;

  float* const wslots = reinterpret_cast<float*>(smem_raw) + warp * S * slot_floats;
  const uint32_t wslots_s = smem_u32(wslots);
  const uint32_t bars_s =
      smem_u32(smem_raw + (size_t)pa.warps * S * slot_floats * sizeof(float)) + (uint32_t)warp * S * 8u;
  const uint32_t slot_bytes = (uint32_t)slot_floats * 4u;
  const uint32_t xoff_bytes = (uint32_t)R * row_floats * 4u;       // inputs behind the parameters,
  const uint32_t goff_bytes = xoff_bytes + (uint32_t)R * D * 4u;   // upstream gradients behind the inputs

  const int64_t gstride = (int64_t)gridDim.x * pa.warps;
  const int64_t g0 = (int64_t)blockIdx.x * pa.warps + warp;
  const int64_t B = a.B;
  int64_t fetch_row = g0 * R;
  const int64_t row_step = gstride * R;

  auto issue = [&](int slot) {  // lane 0 only
    const int rows = (int)min((int64_t)R, B - fetch_row);
    const uint32_t dst = wslots_s + (uint32_t)slot * slot_bytes;
    const uint32_t bar = bars_s + (uint32_t)slot * 8u;
    const uint32_t pbytes = (uint32_t)(rows * row_floats) * 4u, xbytes = (uint32_t)(rows * D) * 4u;
    mbar_expect_tx(bar, pbytes + 2 * xbytes);
    bulk_g2s(dst, a.params + fetch_row * row_floats, pbytes, bar);
    bulk_g2s(dst + xoff_bytes, a.x + fetch_row * D, xbytes, bar);
    bulk_g2s(dst + goff_bytes, a.gy + fetch_row * D, xbytes, bar);
    fetch_row += row_step;
  };

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars_s + (uint32_t)s * 8u, 1);
    fence_mbar_init();
    fence_proxy_async();
    for (int s = 0; s < S; ++s)
      if (fetch_row < B) issue(s);
  }
  __syncwarp();

  int slot = 0;
  uint32_t parity = 0;
  const int fstride = feature_stride(op, 0);
  for (int64_t row0 = g0 * R; row0 < B; row0 += row_step) {
    mbar_wait(bars_s + (uint32_t)slot * 8u, parity);
    const int rows = (int)min((int64_t)R, B - row0);
    float* sp = wslots + slot * slot_floats;
    float* sx = sp + R * row_floats;
    const float* sg = sx + R * D;
    for (int r = sub; r < rows + sub; r += rpw) {  // uniform trip count
      if (r < rows) {
        float* xrow = sx + r * D;
        const float* grow = sg + r * D;
        const float gl = a.gl ? __ldg(a.gl + row0 + r) : 0.f;
        for (int j = j0; j < D_t; j += seg) {
          const int col = tcols ? __ldg(tcols + j) : j;
          float* pj = sp + r * row_floats + j * fstride;
          float gxv;
          op.backward(xrow[col], pj, grow[col], gl, gxv, pj);  // parameter gradients overwrite the parameters
          xrow[col] = gxv;                                      // input gradient overwrites the input
        }
        for (int i = j0; i < a.n_copy; i += seg) {  // identity columns: grad_x = grad_y
          const int col = __ldg(ccols + i);
          xrow[col] = grow[col];
        }
      }
    }
    fence_proxy_async();  // the bulk stores read the slot through the async proxy
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(a.gp + row0 * row_floats, wslots_s + (uint32_t)slot * slot_bytes, (uint32_t)(rows * row_floats) * 4u);
      bulk_s2g(a.gx + row0 * D, wslots_s + (uint32_t)slot * slot_bytes + xoff_bytes, (uint32_t)(rows * D) * 4u);
      bulk_commit();
      if (fetch_row < B) {
        bulk_wait_read0();  // the stores have read the slot: it may be refilled
        issue(slot);
      }
    }
    __syncwarp();
    if (++slot == S) {
      slot = 0;
      parity ^= 1u;
    }
  }
  if (lane == 0) bulk_wait_all0();  // global writes complete before the kernel ends
}

// Try the pipelined backward kernel; returns 1 if it was launched, 0 if the call does not qualify, <0 on error.
template <class Op>
inline int try_launch_pipelined_backward(const LayerBwdArgs& a, const Op& op, int P, int D, cudaStream_t st) {
  if (env_int("FC_PIPE", 1) == 0 || env_int("FC_PIPE_BWD", 1) == 0) return 0;
  const int64_t row_floats = (int64_t)a.D_t * P;
  const bool aligned = ((reinterpret_cast<uintptr_t>(a.params) | reinterpret_cast<uintptr_t>(a.x) |
                         reinterpret_cast<uintptr_t>(a.gy) | reinterpret_cast<uintptr_t>(a.gx) |
                         reinterpret_cast<uintptr_t>(a.gp)) & 15) == 0;
  if (!aligned || a.p_stride != row_floats || a.gp_stride != row_floats || a.x_stride != D || a.gy_stride != D ||
      a.gx_stride != D)
    return 0;
  if (a.D_t + a.n_copy != D) return 0;  // whole rows are streamed: the column lists must cover the row
  if ((row_floats * 4) % 16 != 0 || (D * 4) % 16 != 0) return 0;
  const DeviceInfo& dev = device_info();
  const LaneMap lm = lane_map(a.D_t);
  const int64_t row_bytes = (row_floats + 2 * D) * 4;
  int slot_rows = lm.rows_per_warp;
  const int target = env_int("FC_PIPE_BWD_SLOT_BYTES", 2048);
  while ((int64_t)(slot_rows * 2) * row_bytes <= target) slot_rows *= 2;
  int stages = env_int("FC_PIPE_BWD_STAGES", 2);
  int warps = env_int("FC_PIPE_BWD_WARPS", 16);
  int ctas_per_sm = env_int("FC_PIPE_BWD_CTAS", slot_rows * row_bytes <= 2048 ? 3 : 2);
  const int64_t slot_bytes = slot_rows * row_bytes;
  if (slot_bytes > 96 * 1024) return 0;
  auto smem_need = [&](int w, int s) { return (int64_t)w * s * (slot_bytes + 8) + 128; };
  const int64_t sm_budget = 224 * 1024;
  while ((int64_t)ctas_per_sm * (smem_need(warps, stages) + 1024) > sm_budget) {
    if (stages > 2) --stages;
    else if (warps > 8) --warps;
    else if (ctas_per_sm > 1) --ctas_per_sm;
    else if (warps > 1) --warps;
    else return 0;
  }
  PipeBwdArgs pa;
  pa.a = a;
  pa.a.seg = lm.seg;
  pa.D = D;
  pa.slot_rows = slot_rows;
  pa.stages = stages;
  pa.warps = warps;
  const size_t smem = (size_t)smem_need(warps, stages);
  const int64_t groups = (a.B + slot_rows - 1) / slot_rows;
  int64_t grid = (int64_t)dev.sm_count * ctas_per_sm;
  const int64_t need = (groups + warps - 1) / warps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (prepare_kernel(pipelined_backward_kernel<Op>, smem) != FC_OK) return FC_ERR_CUDA;
  pipelined_backward_kernel<Op><<<(int)grid, warps * 32, smem, st>>>(pa, op);
  if (cudaGetLastError() != cudaSuccess) return FC_ERR_CUDA;
  return 1;
}

}  // namespace fc
