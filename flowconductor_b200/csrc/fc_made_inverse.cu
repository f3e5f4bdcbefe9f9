// fc_made_inverse.cu — the INVERSE of a masked autoregressive layer (sampling direction) in ONE kernel, with the MADE
// conditioner evaluated incrementally.
//
// Replaces AutoregressiveTransform.inverse (flowcon/transforms/autoregressive/autoregressive.py:44-53): the reference
// runs the WHOLE conditioner D times, each pass on the partially inverted outputs, and keeps one more correct feature
// per pass.  Feature f's parameters depend only on features < f (the masks of flowcon/transforms/made.py:28-51), and in
// a MADE whose hidden units are ordered by degree every masked weight row is a PREFIX of its input: a hidden unit of
// degree m reads the units of degree <= m of the layer below.  So pass f only has to compute the hidden units that
// become valid with feature f - 1 (about H / (D - 1) per layer) from the prefix computed so far, and the P outputs of
// feature f.  Summed over the D passes that is HALF of one conditioner evaluation instead of D of them (cfg 3: 32 x
// fewer multiply-adds), and nothing but the inputs and outputs touches HBM.
//
// The host (flowconductor_b200/made_inverse.py) turns a residual MADE into a straight-line PROGRAM of steps — "units
// [j0, j0 + nj) of array `out` = bias + W[:, :k_count] . act(array `in`[:k_count]) (+ array `res`)" — and lays the masked,
// degree-sorted weights out in step order; after the step that completes a feature's parameters the bijection's
// inverse is evaluated for that feature.  The kernel is an interpreter for that program:
//
//   * one CTA = 32 rows (lane = row).  The row tile's state lives in shared memory as [unit][32 rows] fp32 arrays:
//     the features inverted so far, one array per hidden layer output (1 + 2 x blocks arrays of H units), the
//     parameter tile of the feature being inverted.  That state (5 KB per row at cfg 3) is what limits a CTA to 32
//     rows, and is why this is an fp32 CUDA-core kernel: the products are 32 x <=24 x k slivers on a serial chain of
//     ~100 steps, far below a tcgen05 tile.
//   * the weights are one linear stream, identical for every row tile: a producer warp walks the program and feeds a
//     ring of 6 KB slots (<= 64 k-values x <= 24 outputs) with 1-D TMA bulk copies on mbarriers; it runs ahead across
//     steps and row tiles.
//   * the 8 compute warps split the reduction: warp w multiplies k = w, w + 8, ... of a slot (one conflict-free
//     activation load per k, the slot's weight row broadcast as 128-bit loads, packed fp32x2 FMAs into 24
//     accumulators), the partial sums meet in shared memory, bias / skip connection are added and the units stored.
//   * warp 0 then inverts the feature (same element arithmetic as the layer kernels: fc_math.cuh) and the next pass
//     starts.  log|det J| is the sum of the per-feature terms (what the reference's last pass returns).
#include "fc_common.cuh"
#include "fc_tc.cuh"

namespace fc {

using namespace tc;

constexpr int kMR = 32;                 // rows per CTA
constexpr int kMW = 8;                  // compute warps
constexpr int kMThreads = (kMW + 1) * 32;
constexpr int kMSlotK = 64;             // k-values per ring slot
constexpr int kMJT = FC_MADE_MAX_NJ;    // outputs per step (24)
constexpr int kMStages = 4;
constexpr int kMSlotBytes = kMSlotK * kMJT * 4;
constexpr int kMRingBytes = kMStages * kMSlotBytes;
constexpr int kMScratchFloats = kMW * kMJT * kMR;

struct MadeArgs {
  const int4* steps;  // fc_made_step[n_steps] as 3 x int4 each
  int n_steps;
  const float* weights;
  const float* bias;
  int D, PS, n_arrays, hidden;
  const float* z;
  long long ldz;
  float* x;
  long long ldx;
  float* lad;
  int accumulate;
  long long M;
  int num_tiles;
  int32_t* status;
};

// Bounded mbarrier wait: a protocol error ends the kernel with a trap (the launch fails) instead of hanging the GPU.
__device__ __forceinline__ void made_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

template <int NJ4>
__device__ __forceinline__ void made_kloop(float2 (&acc)[kMJT / 2], const float* __restrict__ in_k, const float* __restrict__ slot,
                                           int nk, int warp, bool relu) {
#pragma unroll 4
  for (int kk = warp; kk < nk; kk += kMW) {
    float av = in_k[kk * kMR];
    if (relu) av = fmaxf(av, 0.f);
    const float2 a2 = make_float2(av, av);
    const float4* wr = reinterpret_cast<const float4*>(slot + kk * (NJ4 * 4));
#pragma unroll
    for (int q = 0; q < NJ4; ++q) {
      const float4 w = wr[q];
      acc[2 * q] = __ffma2_rn(a2, make_float2(w.x, w.y), acc[2 * q]);
      acc[2 * q + 1] = __ffma2_rn(a2, make_float2(w.z, w.w), acc[2 * q + 1]);
    }
  }
}

template <class Op>
__global__ void __launch_bounds__(kMThreads, 1) made_inverse_kernel(const MadeArgs a, const Op op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t raw_s = s32(smem_raw);
  const uint32_t base = (raw_s + 127u) & ~127u;
  unsigned char* const gbase = smem_raw + (base - raw_s);
  float* const ring_g = reinterpret_cast<float*>(gbase);
  float* const X = reinterpret_cast<float*>(gbase + kMRingBytes);    // [D][32]: z until a feature is inverted, then x
  float* const H = X + a.D * kMR;                                    // [n_arrays][hidden][32]
  float* const PT = H + (size_t)a.n_arrays * a.hidden * kMR;         // [32 rows][PS]
  float* const SC = PT + kMR * a.PS;                                 // [warp][24][32] partial sums
  const uint32_t bars = base + kMRingBytes + 4u * (uint32_t)(a.D * kMR + a.n_arrays * a.hidden * kMR + kMR * a.PS + kMScratchFloats);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMStages + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kMW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kMW) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        for (int i = 0; i < a.n_steps; ++i) {
          const int4 s0 = __ldg(a.steps + 3 * i), s1 = __ldg(a.steps + 3 * i + 1), s2 = __ldg(a.steps + 3 * i + 2);
          const int k_count = s0.z, nj4 = s1.y;
          const float* src = a.weights + (size_t)(unsigned)s2.y * 4;
          for (int k0 = 0; k0 < k_count; k0 += kMSlotK) {
            const int nk = k_count - k0 < kMSlotK ? k_count - k0 : kMSlotK;
            const uint32_t bytes = (uint32_t)(nk * nj4 * 16);
            made_wait(empty_bar(s), ph ^ 1u);
            mbar_expect_tx(full_bar(s), bytes);
            bulk_load_1d(base + (uint32_t)(s * kMSlotBytes), src + (size_t)k0 * nj4 * 4, bytes, full_bar(s));
            if (++s == kMStages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps
  const int tid = threadIdx.x;  // 0..255
  int s = 0;
  uint32_t ph = 0;
  unsigned status = 0;
  const int D = a.D;
  const size_t arr_floats = (size_t)a.hidden * kMR;
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * kMR;
    for (int i = tid; i < kMR * D; i += kMW * 32) {
      const int r = i / D, f = i - r * D;
      X[f * kMR + r] = (row0 + r < a.M) ? __ldg(a.z + (row0 + r) * a.ldz + f) : 0.f;
    }
    float lad_acc = 0.f;
    named_barrier_sync(1, kMW * 32);
    for (int i = 0; i < a.n_steps; ++i) {
      const int4 s0 = __ldg(a.steps + 3 * i), s1 = __ldg(a.steps + 3 * i + 1), s2 = __ldg(a.steps + 3 * i + 2);
      const int in_array = s0.x, out_array = s0.y, k_count = s0.z, j0 = s0.w;
      const int nj = s1.x, nj4 = s1.y, relu_in = s1.z, res_array = s1.w;
      const int feature = s2.x, b_off = s2.z;
      float2 acc[kMJT / 2];
#pragma unroll
      for (int j = 0; j < kMJT / 2; ++j) acc[j] = make_float2(0.f, 0.f);
      const float* in = (in_array == 0 ? X : H + (size_t)(in_array - 1) * arr_floats) + lane;
      for (int k0 = 0; k0 < k_count; k0 += kMSlotK) {
        const int nk = k_count - k0 < kMSlotK ? k_count - k0 : kMSlotK;
        made_wait(full_bar(s), ph);
        const float* slot = ring_g + s * (kMSlotBytes / 4);
        const float* in_k = in + (size_t)k0 * kMR;
        switch (nj4) {
          case 1: made_kloop<1>(acc, in_k, slot, nk, warp, relu_in != 0); break;
          case 2: made_kloop<2>(acc, in_k, slot, nk, warp, relu_in != 0); break;
          case 3: made_kloop<3>(acc, in_k, slot, nk, warp, relu_in != 0); break;
          case 4: made_kloop<4>(acc, in_k, slot, nk, warp, relu_in != 0); break;
          case 5: made_kloop<5>(acc, in_k, slot, nk, warp, relu_in != 0); break;
          default: made_kloop<6>(acc, in_k, slot, nk, warp, relu_in != 0); break;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
        if (++s == kMStages) {
          s = 0;
          ph ^= 1u;
        }
      }
      // partial sums of this warp's share of the reduction
      {
        float* sc = SC + (warp * kMJT) * kMR + lane;
#pragma unroll
        for (int j = 0; j < kMJT / 2; ++j) {
          if (2 * j < nj) sc[(2 * j) * kMR] = acc[j].x;
          if (2 * j + 1 < nj) sc[(2 * j + 1) * kMR] = acc[j].y;
        }
      }
      named_barrier_sync(1, kMW * 32);
      for (int jj = warp; jj < nj; jj += kMW) {
        float v = __ldg(a.bias + b_off + jj);
        const float* sc = SC + jj * kMR + lane;
#pragma unroll
        for (int w = 0; w < kMW; ++w) v += sc[w * (kMJT * kMR)];
        if (res_array > 0) v += H[(size_t)(res_array - 1) * arr_floats + (size_t)(j0 + jj) * kMR + lane];
        if (out_array > 0) {
          H[(size_t)(out_array - 1) * arr_floats + (size_t)(j0 + jj) * kMR + lane] = v;
        } else {
          PT[lane * a.PS + j0 + jj] = v;
        }
      }
      named_barrier_sync(1, kMW * 32);
      if (feature >= 0) {
        // the feature's parameters are complete: invert it (autoregressive.py:50-52 for the one column that becomes final)
        if (warp == 0) {
          const float zf = X[feature * kMR + lane];
          float xf, lf;
          op.eval(zf, PT + lane * a.PS, xf, lf, status);
          X[feature * kMR + lane] = xf;
          lad_acc += lf;
        }
        named_barrier_sync(1, kMW * 32);
      }
    }
    for (int i = tid; i < kMR * D; i += kMW * 32) {
      const int r = i / D, f = i - r * D;
      if (row0 + r < a.M) a.x[(row0 + r) * a.ldx + f] = X[f * kMR + r];
    }
    if (warp == 0 && row0 + lane < a.M) {
      const long long row = row0 + lane;
      a.lad[row] = a.accumulate ? a.lad[row] + lad_acc : lad_acc;
    }
    named_barrier_sync(1, kMW * 32);
  }
  if (warp == 0 && status != 0 && a.status) atomicOr(a.status, (int)status);
}

template <int KC>
struct MadeRqsOp {
  RqsParams c;
  __device__ __forceinline__ void eval(float z, const float* p, float& x, float& lad, unsigned& status) const {
    rqs_eval<KC>(c, z, p, x, lad, status);
  }
};

struct MadeAffineOp {  // interleaved (raw scale, shift) pairs: autoregressive.py:124-129
  int activation;
  __device__ __forceinline__ void eval(float z, const float* p, float& x, float& lad, unsigned&) const {
    affine_eval(z, p[0], p[1], activation, 1, x, lad);
  }
};

static size_t made_smem_bytes(int D, int PS, int n_arrays, int hidden) {
  return (size_t)kMRingBytes + 4ull * ((size_t)D * kMR + (size_t)n_arrays * hidden * kMR + (size_t)kMR * PS + kMScratchFloats) +
         8 * 2 * kMStages + 128;
}

static int made_check(const fc_made_program* prog, const float* z, int64_t ldz, float* x, int64_t ldx, float* lad, int64_t B,
                      int P, MadeArgs& a) {
  if (!prog || !prog->steps || !prog->weights || !prog->bias || prog->n_steps <= 0) return FC_ERR_INVALID_ARGUMENT;
  if (prog->features <= 0 || prog->hidden <= 0 || prog->n_arrays <= 0 || prog->params_per_feature != P)
    return FC_ERR_INVALID_ARGUMENT;
  if (B < 0) return FC_ERR_INVALID_ARGUMENT;
  if (B > 0 && (!z || !x || !lad)) return FC_ERR_INVALID_ARGUMENT;
  if (ldz < prog->features || ldx < prog->features) return FC_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(prog->weights) & 15) || (reinterpret_cast<uintptr_t>(prog->steps) & 15)) return FC_ERR_UNSUPPORTED;
  if (B >= ((int64_t)1 << 31) * kMR) return FC_ERR_UNSUPPORTED;
  a.steps = reinterpret_cast<const int4*>(prog->steps);
  a.n_steps = prog->n_steps;
  a.weights = prog->weights;
  a.bias = prog->bias;
  a.D = prog->features;
  a.PS = (P & 1) ? P : P + 1;  // odd row stride of the parameter tile: conflict-free for lane = row
  a.n_arrays = prog->n_arrays;
  a.hidden = prog->hidden;
  a.z = z;
  a.ldz = ldz;
  a.x = x;
  a.ldx = ldx;
  a.lad = lad;
  a.M = B;
  a.num_tiles = (int)((B + kMR - 1) / kMR);
  if (made_smem_bytes(a.D, a.PS, a.n_arrays, a.hidden) > (size_t)device_info().max_smem_optin) return FC_ERR_UNSUPPORTED;
  return FC_OK;
}

template <class Op>
static int launch_made(const MadeArgs& a, const Op& op, cudaStream_t stream) {
  auto kern = made_inverse_kernel<Op>;
  const size_t smem = made_smem_bytes(a.D, a.PS, a.n_arrays, a.hidden);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return FC_ERR_CUDA;
  const int sms = device_info().sm_count;
  const int grid = a.num_tiles < sms ? a.num_tiles : sms;
  kern<<<grid, kMThreads, smem, stream>>>(a, op);
  FC_CHECK_LAUNCH();
  return FC_OK;
}

}  // namespace fc

using namespace fc;

extern "C" int64_t fc_made_inverse_smem_bytes(int32_t features, int32_t params_per_feature, int32_t n_arrays, int32_t hidden) {
  if (features <= 0 || params_per_feature <= 0 || n_arrays <= 0 || hidden <= 0) return FC_ERR_INVALID_ARGUMENT;
  const int PS = (params_per_feature & 1) ? params_per_feature : params_per_feature + 1;
  return (int64_t)made_smem_bytes(features, PS, n_arrays, hidden);
}

extern "C" int fc_made_inverse_rqs(const fc_made_program* prog, const float* z, int64_t z_row_stride, float* x,
                                   int64_t x_row_stride, float* logabsdet, int32_t accumulate_logabsdet, int64_t B,
                                   const fc_rqs_config* cfg, int32_t* status, void* stream) {
  RqsParams c;
  int rc = make_rqs_params(cfg, c);
  if (rc != FC_OK) return rc;
  if (!c.inverse) return FC_ERR_INVALID_ARGUMENT;
  MadeArgs a{};
  rc = made_check(prog, z, z_row_stride, x, x_row_stride, logabsdet, B, c.P, a);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  a.accumulate = accumulate_logabsdet;
  a.status = status;
#define CALL(KC)                                        \
  {                                                     \
    MadeRqsOp<KC> op;                                   \
    op.c = c;                                           \
    return launch_made(a, op, (cudaStream_t)stream);    \
  }
  switch (c.K) {
    case 8: CALL(8);
    case 10: CALL(10);
    case 16: CALL(16);
    default: CALL(0);
  }
#undef CALL
}

extern "C" int fc_made_inverse_affine(const fc_made_program* prog, const float* z, int64_t z_row_stride, float* x,
                                      int64_t x_row_stride, float* logabsdet, int32_t accumulate_logabsdet, int64_t B,
                                      int32_t activation, void* stream) {
  if (activation < FC_SCALE_SIGMOID2 || activation > FC_SCALE_SOFTPLUS_EPS) return FC_ERR_INVALID_ARGUMENT;
  MadeArgs a{};
  int rc = made_check(prog, z, z_row_stride, x, x_row_stride, logabsdet, B, 2, a);
  if (rc != FC_OK) return rc;
  if (B == 0) return FC_OK;
  a.accumulate = accumulate_logabsdet;
  a.status = nullptr;
  MadeAffineOp op;
  op.activation = activation;
  return launch_made(a, op, (cudaStream_t)stream);
}
