// Micro-benchmark: per-SMSP throughput of the instruction sequences the row threads of fc_conditioner.cu are made of
// (operand conversion, packed fp32x2 fma, the rational-quadratic spline evaluation) as a function of warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I flowconductor_b200/csrc \
//        -o scripts/microbench/alu_rate scripts/microbench/alu_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include "../../flowconductor_b200/csrc/fc_common.cuh"
using namespace fc;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters, RqsParams c) {
  float v[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) v[j] = (float)(threadIdx.x + j) * 0.01f - 1.f;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // convert 64 values: scale, cvt hi, unpack, sub, cvt lo (as convert_chunk)
      const float2 s2 = make_float2(1.0009765625f, 1.0009765625f), neg1 = make_float2(-1.f, -1.f);
      uint32_t x = 0;
#pragma unroll
      for (int p = 0; p < 32; ++p) {
        const float2 a2 = __fmul2_rn(make_float2(v[2 * p], v[2 * p + 1]), s2);
        const __half2 h2 = __floats2half2_rn(a2.x, a2.y);
        const float2 r2 = __ffma2_rn(__half22float2(h2), neg1, a2);
        const __half2 l2 = __floats2half2_rn(r2.x, r2.y);
        x ^= *reinterpret_cast<const uint32_t*>(&h2) + *reinterpret_cast<const uint32_t*>(&l2);
        v[2 * p] = r2.x + a2.x;
        v[2 * p + 1] = r2.y + a2.y;
      }
      acc += __uint_as_float(x & 0x3fffffffu);
    } else if (MODE == 1) {  // 64 scalar FFMA (3 register operands)
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = fmaf(v[j], v[(j + 1) & 63], v[(j + 7) & 63]);
    } else if (MODE == 2) {  // 32 FFMA2
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        const float2 r = __ffma2_rn(make_float2(v[j], v[j + 1]), make_float2(v[(j + 2) & 63], v[(j + 3) & 63]),
                                    make_float2(v[(j + 8) & 63], v[(j + 9) & 63]));
        v[j] = r.x;
        v[j + 1] = r.y;
      }
    } else if (MODE == 3) {  // one spline evaluation (8 bins)
      unsigned st = 0;
      float y, l;
      rqs_eval<8, true>(c, v[0], v + 8, y, l, st);
      v[0] = y * 0.5f;
      v[9] += l * 1e-3f;
    } else if (MODE == 4) {  // two independent spline evaluations
      unsigned st = 0;
      float y0, l0, y1, l1;
      rqs_eval<8, true>(c, v[0], v + 8, y0, l0, st);
      rqs_eval<8, true>(c, v[1], v + 36, y1, l1, st);
      v[0] = y0 * 0.5f;
      v[1] = y1 * 0.5f;
      v[9] += l0 * 1e-3f;
      v[37] += l1 * 1e-3f;
    } else if (MODE == 5) {  // 32 F2FP packs only
      uint32_t x = 0;
#pragma unroll
      for (int p = 0; p < 32; ++p) {
        const __half2 h2 = __floats2half2_rn(v[2 * p], v[2 * p + 1]);
        x ^= *reinterpret_cast<const uint32_t*>(&h2);
      }
      v[it & 63] += __uint_as_float(x & 0x3fffffffu);
    } else if (MODE == 7) {  // 64 x relu (FMNMX with zero)
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f) - 0.25f;
    } else if (MODE == 8) {  // 64 x relu with integer ops: v & ~(v >> 31)
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int b = __float_as_int(v[j]);
        v[j] = __int_as_float(b & ~(b >> 31)) - 0.25f;
      }
    } else if (MODE == 9) {  // 64-value max |v| as unsigned integers
      unsigned m = 0;
#pragma unroll
      for (int j = 0; j < 64; ++j) m = max(m, __float_as_uint(v[j]) & 0x7fffffffu);
      v[it & 63] = __uint_as_float(m) * 0.999f;
    } else if (MODE == 10) {  // 64-value max of non-negative values as unsigned integers (no mask)
      unsigned m = 0;
#pragma unroll
      for (int j = 0; j < 64; ++j) m = max(m, __float_as_uint(v[j]));
      v[it & 63] = __uint_as_float(m & 0x3fffffffu) * 0.999f;
    } else if (MODE == 11) {  // 64 x (FADD only: subtract)
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = v[j] - 0.25f;
    } else if (MODE == 6) {  // 64 max-abs (FMNMX3)
      float m = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) m = fmaxf(m, fabsf(v[j]));
      v[it & 63] = m * 0.999f;
    }
  }
  const long long t1 = clock64();
#pragma unroll
  for (int j = 0; j < 64; ++j) acc += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc, const RqsParams& c) {
  for (int threads : {128, 256, 512}) {
    const int iters = 2000;
    k<MODE><<<1, threads>>>(out, cyc, iters, c);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-34s warps/SMSP=%d : %8.1f cycles / iteration / warp (%s)\n", name, threads / 128, (double)h / iters,
           cudaGetErrorString(e));
  }
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4 * 512 * 4);
  cudaMalloc(&cyc, 8);
  fc_rqs_config cfg = {8, FC_TAILS_LINEAR, 0, 0, -3.f, 3.f, -3.f, 3.f, 1e-3f, 1e-3f, 1e-3f, 0.0625f};
  RqsParams c;
  make_rqs_params(&cfg, c);
  run<0>("convert 64 values", out, cyc, c);
  run<5>("32 x F2FP pack", out, cyc, c);
  run<6>("64 x max |v|", out, cyc, c);
  run<9>("64 x max |v| (integer)", out, cyc, c);
  run<10>("64 x max v >= 0 (integer)", out, cyc, c);
  run<7>("64 x relu FMNMX + FADD", out, cyc, c);
  run<8>("64 x relu integer + FADD", out, cyc, c);
  run<11>("64 x FADD", out, cyc, c);
  run<1>("64 x FFMA", out, cyc, c);
  run<2>("32 x FFMA2", out, cyc, c);
  run<3>("1 spline evaluation", out, cyc, c);
  run<4>("2 interleaved spline evaluations", out, cyc, c);
  return 0;
}
