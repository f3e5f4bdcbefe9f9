// fc_pipeline.cuh — TMA-pipelined persistent element-wise layer kernels (the fast paths of fc_*_apply / fc_*_backward).
//
// Two skeletons share the bijections' Op::eval / Op::backward:
//   * the TILE RING (tiled_apply_kernel / tiled_backward_kernel, further down; round 2, the default): one producer warp per
//     CTA moves ~50 KB tiles of rows with bulk loads and stores, 8-16 consumer warps evaluate them;
//   * the PER-WARP RING described next (round 1; FC_TILE=0, or rows too long for a CTA-level tile).
//
// Per-warp ring: every warp runs its OWN software pipeline, so there is no CTA-wide synchronisation at all:
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
// y rows (stride == width), 16-byte aligned bases; the per-warp ring also needs row sizes that are multiples of 16
// bytes, the tile ring further down (the default) takes any row length.
#pragma once
#include <map>
#include <mutex>
#include <tuple>

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

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------
// Tile ring: the same streaming scheme with the ring shared by the CTA.  One producer warp moves TILES of `tile_rows`
// consecutive rows (parameters + inputs, two bulk loads; finished rows leave with ONE bulk store, the per-row
// log|det J| with one coalesced store of the producer warp), `consumers` warps evaluate the rows of a tile side by side.
// The per-slot bookkeeping of the per-warp ring above (barrier wait, re-arm, address arithmetic, output copy: ≈ 270 of
// the ≈ 400 warp instructions per row of a K = 8 linear-spline row, where the ring is issue-bound at 54 % of the copy
// peak) is paid once per tile by a warp that does nothing else, and once per tile — not per row — by a consumer.
//   full[s]: armed by the producer's bulk loads (expect_tx);   done[s]: one arrival per consumer warp.
// ------------------------------------------------------------------------------------------------------------
struct TileArgs {
  LayerArgs a;
  int D;            // full row width of x / y
  int tile_rows;    // rows per tile (multiple of 4)
  int stages;       // tiles in flight per CTA
  int consumers;    // consumer warps per CTA (the producer is warp `consumers`)
  int64_t num_tiles;
};

__device__ __forceinline__ void ring_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int kTileMaxWarps = 17;

// Consumer warps per CTA and CTAs per SM of the tile ring: 8 x 2 unless the Op says otherwise (kTileWarps / kTileCtas) —
// rows with little arithmetic (affine, linear spline) want more consumers in flight, the register-heavy cubic spline one
// larger CTA (scripts/bench_tile_ring.py --sweep).
template <class Op, class = void>
struct TileWarps {
  static constexpr int value = 8;
};
template <class Op>
struct TileWarps<Op, decltype((void)Op::kTileWarps)> {
  static constexpr int value = Op::kTileWarps;
};
template <class Op, class = void>
struct TileBwdWarps {  // the adjoints hold 80-100 registers: one CTA of 16 consumer warps per SM (0.98-1.00 of the copy peak on
  static constexpr int value = 16;  // the rational-quadratic rows) unless the Op says otherwise (kTileBwdWarps)
};
template <class Op>
struct TileBwdWarps<Op, decltype((void)Op::kTileBwdWarps)> {
  static constexpr int value = Op::kTileBwdWarps;
};
// Block-size bound of the backward kernel in warps.  ptxas rounds the bound up to a multiple of 128 threads when it sizes the
// register budget: 17 warps (16 consumers) -> 640 threads -> 96 registers, 16 warps -> 512 threads -> 128 registers.  The
// sum-of-sigmoids adjoint spills at 96 (0.68 of the copy peak, 0.91 with 128); the spline adjoints prefer the 16th consumer.
template <class Op, class = void>
struct TileBwdMaxWarps {
  static constexpr int value = kTileMaxWarps;
};
template <class Op>
struct TileBwdMaxWarps<Op, decltype((void)Op::kTileBwdMaxWarps)> {
  static constexpr int value = Op::kTileBwdMaxWarps;
};
template <class Op, class = void>
struct TileCtas {
  static constexpr int value = 2;
};
template <class Op>
struct TileCtas<Op, decltype((void)Op::kTileCtas)> {
  static constexpr int value = Op::kTileCtas;
};

template <class Op, bool kSimple>
__global__ void __launch_bounds__(kTileMaxWarps * 32, PipeMinBlocks<Op>::value) tiled_apply_kernel(const TileArgs ta, const Op op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const LayerArgs& a = ta.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int NW = ta.consumers, S = ta.stages, TR = ta.tile_rows;
  const int P = op.P();
  const int D = ta.D, D_t = kSimple ? 32 : a.D_t;
  const int row_floats = D_t * P;
  const int tile_floats = TR * (row_floats + D + 1);  // parameters, inputs, one log|det J| per row
  float* const tiles = reinterpret_cast<float*>(smem_raw);
  const uint32_t tiles_s = smem_u32(tiles);
  const uint32_t full_s = tiles_s + (uint32_t)S * tile_floats * 4u, done_s = full_s + (uint32_t)S * 8u;
  const int64_t B = a.B, nt = ta.num_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_s + (uint32_t)s * 8u, 1);
      mbar_init(done_s + (uint32_t)s * 8u, (uint32_t)NW);
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();

  if (warp == NW) {  // ---- producer warp
    int64_t t_fetch = blockIdx.x;
    // Rows need not be multiples of 16 bytes: a tile starts at a multiple of 4 rows, so a FULL tile is an aligned block of
    // a 16-byte-multiple size whatever the row length.  Only a ragged last tile of such rows is moved with plain loads /
    // stores of the producer's lanes (released to the consumers by an ordinary arrival on the same barrier).
    auto bulk_ok = [&](int rows) { return ((rows * row_floats) & 3) == 0 && ((rows * D) & 3) == 0; };
    auto issue = [&](int slot) {  // whole warp
      const int64_t row0 = t_fetch * TR;
      const int rows = (int)min((int64_t)TR, B - row0);
      const uint32_t bar = full_s + (uint32_t)slot * 8u;
      if (bulk_ok(rows)) {
        if (lane == 0) {
          const uint32_t dst = tiles_s + (uint32_t)slot * tile_floats * 4u;
          const uint32_t pbytes = (uint32_t)(rows * row_floats) * 4u, xbytes = (uint32_t)(rows * D) * 4u;
          mbar_expect_tx(bar, pbytes + xbytes);
          bulk_g2s(dst, a.params + row0 * row_floats, pbytes, bar);
          bulk_g2s(dst + (uint32_t)TR * row_floats * 4u, a.x + row0 * D, xbytes, bar);
        }
      } else {
        float* d = tiles + slot * tile_floats;
        const float* gp = a.params + row0 * row_floats;
        const float* gx = a.x + row0 * D;
        for (int i = lane; i < rows * row_floats; i += 32) d[i] = __ldg(gp + i);
        for (int i = lane; i < rows * D; i += 32) d[TR * row_floats + i] = __ldg(gx + i);
        __syncwarp();
        if (lane == 0) ring_arrive(bar);
      }
      t_fetch += gridDim.x;
    };
    for (int s = 0; s < S; ++s)
      if (t_fetch < nt) issue(s);
    int slot = 0;
    uint32_t parity = 0;
    const int accumulate = a.accumulate;
    for (int64_t t = blockIdx.x; t < nt; t += gridDim.x) {
      mbar_wait(done_s + (uint32_t)slot * 8u, parity);
      const int64_t row0 = t * TR;
      const int rows = (int)min((int64_t)TR, B - row0);
      const float* sx = tiles + slot * tile_floats + TR * row_floats;
      const float* slad = sx + TR * D;
      for (int r = lane; r < rows; r += 32) {
        float* g = a.lad + row0 + r;
        *g = accumulate ? *g + slad[r] : slad[r];
      }
      if (bulk_ok(rows)) {
        __syncwarp();
        if (lane == 0) {
          bulk_s2g(a.y + row0 * D, tiles_s + (uint32_t)(slot * tile_floats + TR * row_floats) * 4u, (uint32_t)(rows * D) * 4u);
          bulk_commit();
          if (t_fetch < nt) bulk_wait_read0();  // the store has read the tile: it may be refilled
        }
      } else {
        float* gy = a.y + row0 * D;
        for (int i = lane; i < rows * D; i += 32) gy[i] = sx[i];
      }
      __syncwarp();
      if (t_fetch < nt) issue(slot);
      if (++slot == S) {
        slot = 0;
        parity ^= 1u;
      }
    }
    if (lane == 0) bulk_wait_all0();  // global writes complete before the kernel ends
    return;
  }

  // ---- consumer warps
  const int seg = kSimple ? 32 : a.seg, rpw = 32 / seg;
  const int sub = lane / seg, j0 = lane % seg;
  const int32_t* __restrict__ tcols = a.tcols;
  const int col0 = (j0 < D_t) ? (tcols ? __ldg(tcols + j0) : j0) : 0;  // this lane's first column
  const int fstride = feature_stride(op, 0);
  const int jstep = seg * fstride;
  const int rstep = NW * rpw;
  unsigned status = 0;
  int slot = 0;
  uint32_t parity = 0;
  for (int64_t t = blockIdx.x; t < nt; t += gridDim.x) {
    mbar_wait(full_s + (uint32_t)slot * 8u, parity);
    const int rows = (int)min((int64_t)TR, B - t * TR);
    float* const sp = tiles + slot * tile_floats;
    float* const sx = sp + TR * row_floats;
    float* const slad = sx + TR * D;
    for (int rb = warp * rpw; rb < rows; rb += rstep) {  // uniform trip count within the warp
      const int r = rb + sub;
      const bool row_ok = r < rows;
      float* xrow = sx + r * D;
      const float* prow = sp + r * row_floats + j0 * fstride;
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
      if (row_ok && j0 == 0) slad[r] = lad_acc;
    }
    fence_proxy_async();  // the bulk store / the refill touch the tile through the async proxy
    __syncwarp();
    if (lane == 0) ring_arrive(done_s + (uint32_t)slot * 8u);
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

// CTAs of `kern` that fit one SM with this block size / dynamic shared memory (registers included); cached per
// (kernel, threads, smem) — the query costs microseconds and the launchers below run once per layer call.
template <class Kern>
inline int resident_ctas(Kern kern, int threads, size_t smem) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, size_t>, int> cache;
  const auto key = std::make_tuple(reinterpret_cast<const void*>(kern), threads, smem);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess) nb = 0;
  cache.emplace(key, nb);
  return nb;
}

// Launch the tile ring if the shape fits; 1 launched, 0 not applicable, <0 error.  Same preconditions as the per-warp ring
// (checked by the caller).
template <class Op>
inline int try_launch_tiled(const LayerArgs& a, const Op& op, int P, int D, cudaStream_t st) {
  if (env_int("FC_TILE", 1) == 0) return 0;
  const DeviceInfo& dev = device_info();
  const LaneMap lm = lane_map(a.D_t);
  const int64_t row_bytes = ((int64_t)a.D_t * P + D + 1) * 4;
  int consumers = env_int("FC_TILE_WARPS", TileWarps<Op>::value);
  if (consumers < 1) consumers = 1;
  if (consumers > kTileMaxWarps - 1) consumers = kTileMaxWarps - 1;
  // two CTAs per SM, two tiles in flight each, tiles as large as the shared memory allows (≈ 50 KB): measured best on the
  // rational-quadratic (0.98 of the copy peak), affine (0.97) and linear-spline rows (scripts/bench_tile_ring.py --sweep)
  int ctas_per_sm = env_int("FC_TILE_CTAS", TileCtas<Op>::value);
  int stages = env_int("FC_TILE_STAGES", 2);
  int passes = env_int("FC_TILE_PASSES", 32);  // rows (groups of rows_per_warp) a consumer warp evaluates per tile
  const int64_t sm_budget = 224 * 1024;
  const int passes_max = passes;
  auto tile_rows = [&](int ps) { return (consumers * lm.rows_per_warp * ps + 3) / 4 * 4; };
  auto smem_need = [&](int ps, int s) { return (int64_t)s * (tile_rows(ps) * row_bytes + 16) + 128; };
  auto resident = [&](size_t smem) {  // CTAs per SM the registers of this instantiation allow
    return a.D_t == 32 ? resident_ctas(tiled_apply_kernel<Op, true>, (consumers + 1) * 32, smem)
                       : resident_ctas(tiled_apply_kernel<Op, false>, (consumers + 1) * 32, smem);
  };
  for (;;) {
    passes = passes_max;
    while ((int64_t)ctas_per_sm * (smem_need(passes, stages) + 1024) > sm_budget) {
      if (passes > 1) --passes;
      else if (stages > 2) --stages;
      else if (ctas_per_sm > 1) --ctas_per_sm;
      else if (consumers > 1) consumers = (consumers + 1) / 2;  // very long rows: fewer rows per tile
      else return 0;
    }
    if (a.D_t == 32 ? prepare_kernel(tiled_apply_kernel<Op, true>, (size_t)smem_need(passes, stages)) != FC_OK
                    : prepare_kernel(tiled_apply_kernel<Op, false>, (size_t)smem_need(passes, stages)) != FC_OK)
      return FC_ERR_CUDA;
    const int nb = resident((size_t)smem_need(passes, stages));
    if (nb <= 0) return 0;
    if (nb >= ctas_per_sm) break;
    ctas_per_sm = nb;  // fewer, larger CTAs
  }
  // a short batch: smaller tiles so that every SM gets one
  while (passes > 1 && (a.B + tile_rows(passes) - 1) / tile_rows(passes) < (int64_t)dev.sm_count * ctas_per_sm) --passes;
  TileArgs ta;
  ta.a = a;
  ta.a.seg = lm.seg;
  ta.D = D;
  ta.tile_rows = tile_rows(passes);
  ta.stages = stages;
  ta.consumers = consumers;
  ta.num_tiles = (a.B + ta.tile_rows - 1) / ta.tile_rows;
  const size_t smem = (size_t)smem_need(passes, stages);
  int64_t grid = (int64_t)dev.sm_count * ctas_per_sm;
  if (grid > ta.num_tiles) grid = ta.num_tiles;
  if (grid < 1) grid = 1;
  if (a.D_t == 32) {
    if (prepare_kernel(tiled_apply_kernel<Op, true>, smem) != FC_OK) return FC_ERR_CUDA;
    tiled_apply_kernel<Op, true><<<(int)grid, (consumers + 1) * 32, smem, st>>>(ta, op);
  } else {
    if (prepare_kernel(tiled_apply_kernel<Op, false>, smem) != FC_OK) return FC_ERR_CUDA;
    tiled_apply_kernel<Op, false><<<(int)grid, (consumers + 1) * 32, smem, st>>>(ta, op);
  }
  if (cudaGetLastError() != cudaSuccess) return FC_ERR_CUDA;
  last_path() = kPathTileRing;
  return 1;
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
  {
    const int tiled = try_launch_tiled(a, op, P, D, st);  // any row length
    if (tiled != 0) return tiled;
  }
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
  last_path() = kPathWarpRing;
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

// ------------------------------------------------------------------------------------------------------------
// Backward on the tile ring: a tile holds the parameters, the inputs, the upstream gradients and the grad_logabsdet
// scalars of `tile_rows` rows (four bulk loads by the producer warp); the consumers write the parameter gradients over
// the parameters and the input gradients over the inputs; the producer sends the two finished blocks off with two bulk
// stores.  Same traffic as the per-warp backward ring, a fraction of its bookkeeping instructions.
// ------------------------------------------------------------------------------------------------------------
struct TileBwdArgs {
  LayerBwdArgs a;
  int D;
  int tile_rows, stages, consumers;
  int gl_bulk;  // grad_logabsdet rows travel with the tile (base 16-byte aligned); the ragged last tile reads them directly
  int64_t num_tiles;
};

template <class Op>
__global__ void __launch_bounds__(TileBwdMaxWarps<Op>::value * 32) tiled_backward_kernel(const TileBwdArgs ta, const Op op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const LayerBwdArgs& a = ta.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int NW = ta.consumers, S = ta.stages, TR = ta.tile_rows;
  const int P = op.P();
  const int D = ta.D, D_t = a.D_t;
  const int row_floats = D_t * P;
  const int tile_floats = TR * (row_floats + 2 * D + 1);  // parameters, inputs, upstream gradients, grad_logabsdet
  float* const tiles = reinterpret_cast<float*>(smem_raw);
  const uint32_t tiles_s = smem_u32(tiles);
  const uint32_t full_s = tiles_s + (uint32_t)S * tile_floats * 4u, done_s = full_s + (uint32_t)S * 8u;
  const int64_t B = a.B, nt = ta.num_tiles;
  const uint32_t xoff = (uint32_t)TR * row_floats * 4u, goff = xoff + (uint32_t)TR * D * 4u, loff = goff + (uint32_t)TR * D * 4u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_s + (uint32_t)s * 8u, 1);
      mbar_init(done_s + (uint32_t)s * 8u, (uint32_t)NW);
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();

  if (warp == NW) {  // ---- producer warp (lane 0 moves the tiles; the other lanes only help with a ragged last tile)
    int64_t t_fetch = blockIdx.x;
    auto bulk_ok = [&](int rows) { return ((rows * row_floats) & 3) == 0 && ((rows * D) & 3) == 0; };  // see the forward ring
    auto issue = [&](int slot) {  // whole warp
      const int64_t row0 = t_fetch * TR;
      const int rows = (int)min((int64_t)TR, B - row0);
      const uint32_t bar = full_s + (uint32_t)slot * 8u;
      if (bulk_ok(rows)) {
        if (lane == 0) {
          const uint32_t dst = tiles_s + (uint32_t)slot * tile_floats * 4u;
          const uint32_t pbytes = (uint32_t)(rows * row_floats) * 4u, xbytes = (uint32_t)(rows * D) * 4u;
          const uint32_t lbytes = (ta.gl_bulk && rows == TR) ? (uint32_t)TR * 4u : 0u;
          mbar_expect_tx(bar, pbytes + 2 * xbytes + lbytes);
          bulk_g2s(dst, a.params + row0 * row_floats, pbytes, bar);
          bulk_g2s(dst + xoff, a.x + row0 * D, xbytes, bar);
          bulk_g2s(dst + goff, a.gy + row0 * D, xbytes, bar);
          if (lbytes) bulk_g2s(dst + loff, a.gl + row0, lbytes, bar);
        }
      } else {
        float* d = tiles + slot * tile_floats;
        const float* gp = a.params + row0 * row_floats;
        const float* gx = a.x + row0 * D;
        const float* gg = a.gy + row0 * D;
        for (int i = lane; i < rows * row_floats; i += 32) d[i] = __ldg(gp + i);
        for (int i = lane; i < rows * D; i += 32) {
          d[TR * row_floats + i] = __ldg(gx + i);
          d[TR * (row_floats + D) + i] = __ldg(gg + i);
        }
        __syncwarp();
        if (lane == 0) ring_arrive(bar);
      }
      t_fetch += gridDim.x;
    };
    for (int s = 0; s < S; ++s)
      if (t_fetch < nt) issue(s);
    int slot = 0;
    uint32_t parity = 0;
    for (int64_t t = blockIdx.x; t < nt; t += gridDim.x) {
      mbar_wait(done_s + (uint32_t)slot * 8u, parity);
      const int64_t row0 = t * TR;
      const int rows = (int)min((int64_t)TR, B - row0);
      if (bulk_ok(rows)) {
        if (lane == 0) {
          const uint32_t src = tiles_s + (uint32_t)slot * tile_floats * 4u;
          bulk_s2g(a.gp + row0 * row_floats, src, (uint32_t)(rows * row_floats) * 4u);
          bulk_s2g(a.gx + row0 * D, src + xoff, (uint32_t)(rows * D) * 4u);
          bulk_commit();
          if (t_fetch < nt) bulk_wait_read0();  // the stores have read the tile: it may be refilled
        }
      } else {
        const float* d = tiles + slot * tile_floats;
        float* ogp = a.gp + row0 * row_floats;
        float* ogx = a.gx + row0 * D;
        for (int i = lane; i < rows * row_floats; i += 32) ogp[i] = d[i];
        for (int i = lane; i < rows * D; i += 32) ogx[i] = d[TR * row_floats + i];
      }
      __syncwarp();
      if (t_fetch < nt) issue(slot);
      if (++slot == S) {
        slot = 0;
        parity ^= 1u;
      }
    }
    if (lane == 0) bulk_wait_all0();  // global writes complete before the kernel ends
    return;
  }

  // ---- consumer warps
  const int seg = a.seg, rpw = 32 / seg;
  const int sub = lane / seg, j0 = lane % seg;
  const int32_t* __restrict__ tcols = a.tcols;
  const int32_t* __restrict__ ccols = a.ccols;
  const int fstride = feature_stride(op, 0);
  const int rstep = NW * rpw;
  int slot = 0;
  uint32_t parity = 0;
  for (int64_t t = blockIdx.x; t < nt; t += gridDim.x) {
    mbar_wait(full_s + (uint32_t)slot * 8u, parity);
    const int64_t row0 = t * TR;
    const int rows = (int)min((int64_t)TR, B - row0);
    float* const sp = tiles + slot * tile_floats;
    float* const sx = sp + TR * row_floats;
    const float* const sg = sx + TR * D;
    const float* const sl = sg + TR * D;
    const bool gl_tile = ta.gl_bulk && rows == TR;
    for (int r = warp * rpw + sub; r < rows; r += rstep) {
      float* xrow = sx + r * D;
      const float* grow = sg + r * D;
      const float gl = a.gl ? (gl_tile ? sl[r] : __ldg(a.gl + row0 + r)) : 0.f;
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
    fence_proxy_async();  // the bulk stores / the refill touch the tile through the async proxy
    __syncwarp();
    if (lane == 0) ring_arrive(done_s + (uint32_t)slot * 8u);
    if (++slot == S) {
      slot = 0;
      parity ^= 1u;
    }
  }
}

template <class Op>
inline int try_launch_tiled_backward(const LayerBwdArgs& a, const Op& op, int P, int D, cudaStream_t st) {
  if (env_int("FC_TILE", 1) == 0 || env_int("FC_TILE_BWD", 1) == 0) return 0;
  const DeviceInfo& dev = device_info();
  const LaneMap lm = lane_map(a.D_t);
  const int64_t row_bytes = ((int64_t)a.D_t * P + 2 * D + 1) * 4;
  int consumers = env_int("FC_TILE_BWD_WARPS", TileBwdWarps<Op>::value);
  if (consumers < 1) consumers = 1;
  if (consumers > TileBwdMaxWarps<Op>::value - 1) consumers = TileBwdMaxWarps<Op>::value - 1;
  int ctas_per_sm = env_int("FC_TILE_BWD_CTAS", TileCtas<Op>::value);
  int stages = env_int("FC_TILE_BWD_STAGES", 2);
  int passes = env_int("FC_TILE_BWD_PASSES", 32);
  const int64_t sm_budget = 224 * 1024;
  const int passes_max = passes;
  auto tile_rows = [&](int ps) { return (consumers * lm.rows_per_warp * ps + 3) / 4 * 4; };
  auto smem_need = [&](int ps, int s) { return (int64_t)s * (tile_rows(ps) * row_bytes + 16) + 128; };
  for (;;) {
    passes = passes_max;
    while ((int64_t)ctas_per_sm * (smem_need(passes, stages) + 1024) > sm_budget) {
      if (passes > 1) --passes;
      else if (stages > 2) --stages;
      else if (ctas_per_sm > 1) --ctas_per_sm;
      else if (consumers > 1) consumers = (consumers + 1) / 2;  // very long rows: fewer rows per tile
      else return 0;
    }
    const size_t smem = (size_t)smem_need(passes, stages);
    if (prepare_kernel(tiled_backward_kernel<Op>, smem) != FC_OK) return FC_ERR_CUDA;
    const int nb = resident_ctas(tiled_backward_kernel<Op>, (consumers + 1) * 32, smem);  // registers included
    if (nb <= 0) return 0;
    if (nb >= ctas_per_sm) break;
    ctas_per_sm = nb;  // fewer, larger CTAs
  }
  while (passes > 1 && (a.B + tile_rows(passes) - 1) / tile_rows(passes) < (int64_t)dev.sm_count * ctas_per_sm) --passes;
  TileBwdArgs ta;
  ta.a = a;
  ta.a.seg = lm.seg;
  ta.D = D;
  ta.tile_rows = tile_rows(passes);
  ta.stages = stages;
  ta.consumers = consumers;
  ta.gl_bulk = a.gl != nullptr && (reinterpret_cast<uintptr_t>(a.gl) & 15) == 0;
  ta.num_tiles = (a.B + ta.tile_rows - 1) / ta.tile_rows;
  const size_t smem = (size_t)smem_need(passes, stages);
  int64_t grid = (int64_t)dev.sm_count * ctas_per_sm;
  if (grid > ta.num_tiles) grid = ta.num_tiles;
  if (grid < 1) grid = 1;
  if (prepare_kernel(tiled_backward_kernel<Op>, smem) != FC_OK) return FC_ERR_CUDA;
  tiled_backward_kernel<Op><<<(int)grid, (consumers + 1) * 32, smem, st>>>(ta, op);
  if (cudaGetLastError() != cudaSuccess) return FC_ERR_CUDA;
  last_path() = kPathTileRing;
  return 1;
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
  {
    const int tiled = try_launch_tiled_backward(a, op, P, D, st);  // any row length
    if (tiled != 0) return tiled;
  }
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
  last_path() = kPathWarpRing;
  return 1;
}

}  // namespace fc
