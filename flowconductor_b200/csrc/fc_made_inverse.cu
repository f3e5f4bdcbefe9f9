// fc_made_inverse.cu — entry points of the incremental autoregressive inverse for the rational-quadratic spline and
// affine layers (kernel: fc_made_inverse.cuh; the other families instantiate it in their own files).
#include "fc_made_inverse.cuh"

namespace fc {

// Bijections: `pc` points at this row's column of the parameter tile ([P][32 rows]: parameter i at pc[32 i]).
template <int KC>
struct MadeRqsOp {
  RqsParams c;
  __device__ __forceinline__ void eval(float z, const float* pc, float& x, float& lad, unsigned& status) const {
    constexpr int PM = 3 * (KC ? KC : FC_MAX_BINS_GENERIC) + 1;
    float p[PM];
    if constexpr (KC != 0) {
#pragma unroll
      for (int i = 0; i < PM; ++i) p[i] = i < c.P ? pc[i * kMR] : 0.f;
    } else {
      for (int i = 0; i < c.P; ++i) p[i] = pc[i * kMR];
    }
    rqs_eval<KC, (KC != 0), (KC > 1 ? KC + 1 : 1)>(c, z, p, x, lad, status);
  }
};

struct MadeAffineOp {  // interleaved (raw scale, shift) pairs: autoregressive.py:124-129
  int activation;
  __device__ __forceinline__ void eval(float z, const float* pc, float& x, float& lad, unsigned&) const {
    affine_eval(z, pc[0], pc[kMR], activation, 1, x, lad);
  }
};

}  // namespace fc

using namespace fc;

extern "C" int fc_made_inverse_profile(unsigned long long* out32) {
  if (!out32) return FC_ERR_INVALID_ARGUMENT;
  if (cudaMemcpyFromSymbol(out32, g_made_prof, sizeof(unsigned long long) * 32) != cudaSuccess) return FC_ERR_CUDA;
  return FC_OK;
}

extern "C" int64_t fc_made_inverse_smem_bytes(int32_t features, int32_t params_per_feature, int32_t n_arrays, int32_t hidden,
                                              int32_t n_bias) {
  if (features <= 0 || params_per_feature <= 0 || n_arrays <= 0 || hidden <= 0 || n_bias <= 0) return FC_ERR_INVALID_ARGUMENT;
  return (int64_t)MadeSmem::total(2, features, params_per_feature, n_arrays, hidden, n_bias);  // with the smallest weight ring
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
