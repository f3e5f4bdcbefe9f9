// Micro-benchmarks behind the round-2 design of the fused conditioner kernel:
//   1. tcgen05.ld / tcgen05.st throughput per SM (how expensive is draining partial accumulators?)
//   2. issue rate of tcgen05.mma.kind::f16 (128 x N x 16), A from tensor memory as packed f16x2 columns
//   3. numeric check of the "3xFP16" product a = hi + lo (both fp16) with A in TMEM and B in a 128-byte-swizzled
//      K-major shared-memory tile, against an fp64 product on the host
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flowconductor_b200/csrc -o scripts/microbench/tmem_rate \
//        scripts/microbench/tmem_rate.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>
#include <vector>
#include "../../flowconductor_b200/csrc/fc_tc.cuh"
using namespace fc::tc;

__device__ __forceinline__ bool wait_bounded(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < 20000000; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

// ---------------------------------------------------------------- 1. LDTM / STTM throughput
template <int X, bool STORE>
__global__ void __launch_bounds__(512, 1) ldst_kernel(long long* out, float* sink, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(s32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (uint32_t)(lane + i);
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = (uint32_t)(((it * X) + (warp >> 2) * 64) & 255);
      if (STORE) {
        tmem_st16(tmem + lane_sel + col, v);
        if (X == 32) tmem_st16(tmem + lane_sel + col + 16, v + 16);
        tmem_wait_st();
      } else {
        if (X == 32) {
          tmem_ld32(tmem + lane_sel + col, v);
        } else {
          tmem_ld16(tmem + lane_sel + col, v);
        }
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < X; ++i) acc += __uint_as_float(v[i]);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int X, bool STORE>
static void run_ldst(const char* name, long long* d_out, float* d_sink, int nwarps, int grid) {
  const int iters = 4096;
  ldst_kernel<X, STORE><<<grid, 512>>>(d_out, d_sink, iters, nwarps);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * nwarps * 32 * X * 4;
  printf("%-10s x%-2d warps=%2d grid=%3d : %8.1f cycles/instr/warp, %7.1f B/cycle/SM  (%s)\n", name, X, nwarps, grid,
         (double)h / iters, bytes / (double)h, cudaGetErrorString(e));
}

// ---------------------------------------------------------------- 2. kind::f16 MMA rate
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {  // A, B = f16 (format 0), D = f32
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_s = s32(smem_raw);
  const uint32_t base = (raw_s + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, slot = bar + 16;
  volatile uint32_t* slot_g = reinterpret_cast<volatile uint32_t*>(smem_raw + (base - raw_s) + 96 * 1024 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_g;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t a = make_smem_desc(base, 128), b = make_smem_desc(base + 32 * 1024, 128);
    long long t0 = 0, t1 = 0;
    bool ok = true;
    for (int rep = 0; rep < 2 && ok; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t d = tmem + (uint32_t)((i % NACC) * N);
          if (TS) {
            umma_f16_ts(d, tmem + 448 + (i & 1) * 8, b, idesc, i >= NACC ? 1u : 0u);
          } else {
            umma_f16_ss(d, a, b, idesc, i >= NACC ? 1u : 0u);
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
      ok = wait_bounded(bar, rep & 1);
      t1 = clock64();
    }
    if (lane == 0) out[0] = ok ? t1 - t0 : -1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, bool TS, int NACC>
static void run_rate(const char* name, long long* d_out) {
  const int iters = 2048;
  auto k = rate_kernel<N, TS, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<1, 128, 100 * 1024>>>(d_out, iters);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("f16 %-34s N=%3d  %7.1f cycles / MMA  (%s)\n", name, N, (double)h / iters, cudaGetErrorString(e));
}

// ---------------------------------------------------------------- 3. numeric check of the 3xFP16 product
// C[128, N] = A[128, K] * B[N, K]^T, K = 64, one CTA.  a = a_hi + a_lo, b = b_hi + b_lo (fp16, round to nearest);
// A goes to tensor memory as packed f16x2 columns (k even in the low half), B to shared memory as two K-major
// tiles of N rows x 64 fp16 (= one 128-byte swizzle span per row).
constexpr int kK = 64;
template <int N>
__global__ void __launch_bounds__(128, 1) check_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ C, int* flag) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_s = s32(smem_raw);
  const uint32_t base = (raw_s + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - raw_s);
  constexpr int TILE = N * 128;  // bytes of one fp16 plane of B
  const uint32_t bar = base + 2 * TILE, slot = bar + 16;
  volatile uint32_t* slot_g = reinterpret_cast<volatile uint32_t*>(gbase + 2 * TILE + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  if (t == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_g;
  // B planes: row n, 16-byte chunk c (8 fp16) lands at n*128 + ((c ^ (n & 7)) << 4)
  for (int n = t; n < N; n += 128) {
    for (int c = 0; c < 8; ++c) {
      __half hi[8], lo[8];
      for (int i = 0; i < 8; ++i) {
        const float v = B[n * kK + c * 8 + i];
        hi[i] = __float2half_rn(v);
        lo[i] = __float2half_rn(v - __half2float(hi[i]));
      }
      const uint32_t off = (uint32_t)(n * 128 + ((c ^ (n & 7)) << 4));
      *reinterpret_cast<uint4*>(gbase + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(gbase + TILE + off) = *reinterpret_cast<uint4*>(lo);
    }
  }
  fence_proxy_async_smem();
  // A: thread t = row t = TMEM lane t; columns [256, 256+32) hold hi (K = 64 -> 32 packed columns), [288, 320) lo
  {
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    uint32_t hi[32], lo[32];
    for (int j = 0; j < 32; ++j) {
      const float v0 = A[t * kK + 2 * j], v1 = A[t * kK + 2 * j + 1];
      const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
      const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
      hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    tmem_st16(tmem + lane_sel + 256, hi);
    tmem_st16(tmem + lane_sel + 272, hi + 16);
    tmem_st16(tmem + lane_sel + 288, lo);
    tmem_st16(tmem + lane_sel + 304, lo + 16);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t b_hi = make_smem_desc(base, 128), b_lo = make_smem_desc(base + TILE, 128);
    if (elect_one()) {
      for (int kk = 0; kk < kK / 16; ++kk) {
        const uint64_t o = (uint64_t)(kk * 2);  // 16 fp16 = 32 bytes inside the swizzle span
        const uint32_t a_hi = tmem + 256 + kk * 8, a_lo = tmem + 288 + kk * 8;
        umma_f16_ts(tmem, a_lo, b_hi + o, idesc, kk > 0 ? 1u : 0u);
        umma_f16_ts(tmem, a_hi, b_lo + o, idesc, 1u);
        umma_f16_ts(tmem, a_hi, b_hi + o, idesc, 1u);
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  const bool ok = wait_bounded(bar, 0);
  if (!ok && t == 0) *flag = 1;
  tc_fence_after();
  if (ok) {
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    for (int j = 0; j < N; j += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + lane_sel + j, v);
      tmem_wait_ld();
      for (int i = 0; i < 16; ++i) C[t * N + j + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N>
static void run_check(float scale_a, float scale_b) {
  std::vector<float> A(128 * kK), B(N * kK), C(128 * N);
  srand(1234);
  auto rnd = [] { return (float)((rand() / (double)RAND_MAX) * 2.0 - 1.0); };
  for (auto& v : A) v = rnd() * scale_a;
  for (auto& v : B) v = rnd() * scale_b;
  float *dA, *dB, *dC;
  int* dflag;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dC, C.size() * 4);
  cudaMalloc(&dflag, 4);
  cudaMemset(dflag, 0, 4);
  cudaMemset(dC, 0, C.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  auto k = check_kernel<N>;
  const int smem = 2 * N * 128 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(dA, dB, dC, dflag);
  int flag = 0;
  cudaError_t e = cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
  double max_err = 0, rms = 0, rms32 = 0, ref_rms = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      float r32 = 0.f;
      for (int kx = 0; kx < kK; ++kx) {
        ref += (double)A[m * kK + kx] * (double)B[n * kK + kx];
        r32 = fmaf(A[m * kK + kx], B[n * kK + kx], r32);
      }
      const double d = C[m * N + n] - ref;
      max_err = fmax(max_err, fabs(d));
      rms += d * d;
      rms32 += (r32 - ref) * (r32 - ref);
      ref_rms += ref * ref;
    }
  const double cnt = 128.0 * N;
  printf("3xFP16 check N=%3d scale %g x %g: rms err %.3e (fp32 FMA chain %.3e), max %.3e, rms value %.3e, timeout=%d (%s)\n", N,
         scale_a, scale_b, sqrt(rms / cnt), sqrt(rms32 / cnt), max_err, sqrt(ref_rms / cnt), flag, cudaGetErrorString(e));
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dC);
  cudaFree(dflag);
}

int main() {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, sizeof(long long));
  cudaMalloc(&d_sink, sizeof(float));
  run_check<128>(1.f, 1.f);
  run_check<256>(1.f, 1.f);
  run_check<128>(100.f, 0.01f);
  run_check<128>(1e-3f, 1.f);
  for (int nw : {1, 4, 8, 16}) run_ldst<16, false>("tcgen05.ld", d_out, d_sink, nw, 1);
  for (int nw : {4, 8, 16}) run_ldst<32, false>("tcgen05.ld", d_out, d_sink, nw, 1);
  run_ldst<32, false>("tcgen05.ld", d_out, d_sink, 8, 148);
  for (int nw : {1, 4, 8, 16}) run_ldst<16, true>("tcgen05.st", d_out, d_sink, nw, 1);
  for (int nw : {4, 8}) run_ldst<32, true>("tcgen05.st", d_out, d_sink, nw, 1);
  run_rate<128, true, 1>("TS, one accumulator", d_out);
  run_rate<256, true, 1>("TS, one accumulator", d_out);
  run_rate<128, true, 2>("TS, two alternating accumulators", d_out);
  run_rate<192, true, 2>("TS, two alternating accumulators", d_out);
  run_rate<128, false, 1>("SS, one accumulator", d_out);
  run_rate<256, false, 1>("SS, one accumulator", d_out);
  return 0;
}
