// Micro-benchmark: issue rate of tcgen05.mma.kind::tf32 (128 x N x 8) on one SM — dependent chain into one accumulator
// vs two alternating accumulators, A from shared memory (SS) vs tensor memory (TS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flowconductor_b200/csrc -o /tmp/mma_rate scripts/microbench/mma_rate.cu
#include <cstdio>
#include <cuda.h>
#include "../../flowconductor_b200/csrc/fc_tc.cuh"
using namespace fc::tc;

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_s = s32(smem_raw);
  const uint32_t base = (raw_s + 1023u) & ~1023u;
  const uint32_t bar = base + 64 * 1024, slot = bar + 16;
  volatile uint32_t* slot_g = reinterpret_cast<volatile uint32_t*>(smem_raw + (base - raw_s) + 64 * 1024 + 16);
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
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint64_t a = make_smem_desc(base, 64), b = make_smem_desc(base + 16 * 1024, 64);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 = warm-up
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t d = tmem + (uint32_t)((i % NACC) * N);
          if (TS) {
            umma_tf32_ts(d, tmem + 448 + (i & 1) * 8, b, idesc, i >= NACC ? 1u : 0u);
          } else {
            umma_tf32_ss(d, a, b, idesc, i >= NACC ? 1u : 0u);
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      t1 = clock64();
    }
    if (lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, bool TS, int NACC>
void run(const char* name, long long* d_out) {
  const int iters = 2048;
  auto k = rate_kernel<N, TS, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  k<<<1, 128, 80 * 1024>>>(d_out, iters);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d  %7.1f cycles / MMA  (%s)\n", name, N, (double)h / iters, cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long));
  run<64, false, 1>("SS, one accumulator", d_out);
  run<128, false, 1>("SS, one accumulator", d_out);
  run<192, false, 1>("SS, one accumulator", d_out);
  run<256, false, 1>("SS, one accumulator", d_out);
  run<64, false, 2>("SS, two alternating accumulators", d_out);
  run<128, false, 2>("SS, two alternating accumulators", d_out);
  run<192, false, 2>("SS, two alternating accumulators", d_out);
  run<64, true, 1>("TS, one accumulator", d_out);
  run<128, true, 1>("TS, one accumulator", d_out);
  run<192, true, 1>("TS, one accumulator", d_out);
  run<192, true, 2>("TS, two alternating accumulators", d_out);
  return 0;
}
