// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, bf16, SS mode, M = 256 over a CTA pair) as a function
// of N, operands resident in shared memory (SW128 K-major, contents irrelevant), one CTA per SM, 74 clusters of 2.
// MODE bit 0: one accumulator chain (every MMA accumulates into the same columns) instead of alternating between two;
// MODE bit 1: a multicast tcgen05.commit after every 12 MMAs; MODE bit 2: A start address walks kh * 1024 B.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../include -I ../../arbitrarystyletransfer_b200/csrc mma_rate2.cu -o mma_rate2
#include <cstdio>
#include <cuda_runtime.h>
#include "tc.cuh"
using namespace ast::tc;

template <int N, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar, dummy[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&dummy[i]), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(smem_u32(&slot));
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = *(volatile uint32_t*)&slot;
  if (warp == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(256, N);
    const uint64_t ad0 = make_sdesc_k128(base), bd0 = make_sdesc_k128(base + 65536);
    long long t0 = clock64();
    int cs = 0;
    for (int i = 0; i < iters; ++i) {
      if (elect_one_sync()) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ad = ad0 + ((MODE & 4) ? (uint64_t)(kh * 64) : 0) + j * 2;
            const uint64_t bd = bd0 + ((MODE & 4) ? (uint64_t)(kh * ((N / 2) * 128 >> 4)) : 0) + j * 2;
            umma_bf16_2sm(tm + (uint32_t)(((MODE & 1) ? 0 : (kh & 1)) * N), ad, bd, idesc, 1u);
          }
        if (MODE & 2) umma_commit_2sm(smem_u32(&dummy[cs]));
      }
      __syncwarp();
      cs = (cs + 1) & 7;
    }
    if (lane == 0) {
      umma_commit_2sm(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  if (rank == 1 && warp == 0 && lane == 0) mbar_wait(smem_u32(&bar), 0);
  tc_fence_before(); cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm<512>(tm); }
}

template <int N, int MODE>
void run(long long* d, int nblk) {
  const int iters = 1024;
  const int smem = 160 * 1024;
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, MODE><<<nblk, 128, smem>>>(d, iters);
  k<N, MODE><<<nblk, 128, smem>>>(d, iters);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = (double)h / (iters * 12.0);
  printf("2-CTA M=256 N=%3d mode=%d (%s%s%s) blocks=%3d : %7.1f cycles/MMA -> %6.1f%% of the 8192 flop/clk/SM dense rate (%s)\n", N, MODE,
         MODE & 1 ? "chain " : "", MODE & 2 ? "commit/12 " : "", MODE & 4 ? "walk " : "", nblk, cyc,
         100.0 * (2.0 * 128 * N * 16 / cyc) / 8192.0, cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int nblk : {2, 148}) {
    run<32, 0>(d, nblk); run<64, 0>(d, nblk); run<128, 0>(d, nblk); run<256, 0>(d, nblk);
  }
  run<64, 1>(d, 148); run<64, 2>(d, 148); run<64, 4>(d, 148); run<64, 7>(d, 148);
  run<128, 1>(d, 148); run<128, 2>(d, 148); run<128, 7>(d, 148);
  run<256, 7>(d, 148);
  return 0;
}
