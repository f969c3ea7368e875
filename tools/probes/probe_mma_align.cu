// Microbenchmark (one CTA per SM): cycles per tcgen05.mma (bf16, M = 128, K = 16) as a function of N, of how the no-swizzle
// K-major A operand sits in shared memory (stride between 8-row groups = SBO, 16-byte shift of the start address: the in-plane
// taps of the implicit-GEMM convs in csrc/conv3d_tc.cu are exactly such shifts), and of how many distinct TMEM accumulators
// consecutive MMAs rotate over (1 = every MMA accumulates into the columns the previous one wrote).
// The issue loop is unrolled x16 with precomputed descriptors so that the issuing thread is not the bottleneck.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I semstereo_b200/csrc -I include tools/probes/probe_mma_align.cu -o /tmp/probe
#include <cstdio>
#include "tc_common.cuh"

template <int NACC>
__global__ void __launch_bounds__(128, 1) probe(int N, uint32_t sbo_a, uint32_t shift_a, uint32_t lbo_a, int iters, int nmma, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 32) {
    const uint32_t a_lo = tc::desc_lo(tc::smem_u32(smem) + shift_a, lbo_a), a_hi = tc::desc_hi(sbo_a);
    const uint32_t b_lo = tc::desc_lo(tc::smem_u32(smem + 96 * 1024), (uint32_t)N * 16), b_hi = tc::desc_hi(128);
    const uint32_t idesc = tc::make_idesc_bf16(128, N);
    const uint32_t astep = sbo_a >> 4;
    long long best = 1ll << 60;
    for (int it = 0; it < iters; ++it) {
      const long long t0 = clock64();
      for (int i = 0; i < nmma; i += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
          tc::mma_bf16_lohi(tmem_base + (uint32_t)((u % NACC) * (512 / NACC)), a_lo + (uint32_t)(u & 7) * astep, a_hi, b_lo, b_hi, idesc, 1u);
      }
      tc::mma_commit(&bar);
      tc::mbar_wait(&bar, it & 1);
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (blockIdx.x == 0) *cycles = best;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int nmma = 4096;
  const int Ns[] = {32, 64, 96, 128, 192, 256};
  const uint32_t geo[][2] = {{128, 0}, {160, 0}, {160, 16}};
  for (int nacc : {1, 2, 4})
    for (int N : Ns)
      for (auto& g : geo) {
        if (nacc == 4 && N > 128) continue;
        if (nacc == 1) probe<1><<<sms, 128, 200 * 1024>>>(N, g[0], g[1], 2880, 5, nmma, d);
        else if (nacc == 2) probe<2><<<sms, 128, 200 * 1024>>>(N, g[0], g[1], 2880, 5, nmma, d);
        else probe<4><<<sms, 128, 200 * 1024>>>(N, g[0], g[1], 2880, 5, nmma, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const double per = (double)cyc / nmma, ideal = 128.0 * N * 16 / 4096.0;
        printf("accumulators %d  N=%3d  SBO=%3u shift %2u : %7.1f cycles/MMA  (ideal %5.1f, efficiency %.2f)  %s\n", nacc, N, g[0], g[1], per, ideal,
               ideal / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
        fflush(stdout);
      }
  return 0;
}
