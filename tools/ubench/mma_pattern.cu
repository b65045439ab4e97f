// Microbenchmark: what makes a tcgen05.mma (M128 x N x K16, bf16, SS) in the conv kernels cost more than the 48 / 64 / 128
// cycles of tools/ubench/mma_rate.cu?  One CTA per SM, one issuing thread, operands resident; factors switched on one
// at a time: (1) A start address walks through kh * 1024 B offsets and ring slots like the kw-box conv kernel,
// (2) a tcgen05.commit after every 12 MMAs, (4) chains of 36 MMAs into one of four accumulators (instead of 4 MMAs
// alternating between two), (8) issue under elect.sync in a converged warp that also waits on an (already complete)
// mbarrier per group of 12.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../include -I ../../arbitrarystyletransfer_b200/csrc mma_pattern.cu -o mma_pattern
#include <cstdio>
#include <cuda_runtime.h>
#include "tc.cuh"
using namespace ast::tc;

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int groups) {
  extern __shared__ uint8_t raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar, dummy[8], ready;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&ready), 1);
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&dummy[i]), 1);
    fence_barrier_init();
    mbar_arrive(smem_u32(&ready));      // phase 0 complete: waits on parity 0 return at once
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *(volatile uint32_t*)&slot;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    constexpr int A_SLOT = 18432, B_BYTES = N * 128;
    const uint32_t a_base = base, b_base = base + 8 * A_SLOT;      // 8 A slots (144 KB) + 4 B tiles
    const uint64_t ad0 = make_sdesc_k128(a_base), bd0 = make_sdesc_k128(b_base);
    constexpr bool walk = MODE & 1, commit = MODE & 2, chain = MODE & 4, elect = MODE & 8, commit2 = MODE & 16, per4 = MODE & 32;
    long long t0 = clock64();
    int sa = 0, acc = 0, inchain = 0, cslot = 0;
    for (int g = 0; g < groups; ++g) {
      if (elect) { mbar_wait(smem_u32(&ready), 0u); tc_fence_after(); }
      const bool issue = elect ? elect_one_sync() : (lane == 0);
      if (issue) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ad = (walk ? ad0 + (uint64_t)sa * (A_SLOT >> 4) + (uint64_t)(kh * 64) : ad0) + j * 2;
            const uint64_t bd = (walk ? bd0 + (uint64_t)(kh * (B_BYTES >> 4)) : bd0) + j * 2;
            const uint32_t d = tm + (uint32_t)((chain ? acc : (kh & 1)) * N);
            umma_bf16(d, ad, bd, idesc, 1u);
          }
          if (per4) umma_commit(smem_u32(&dummy[kh]));
        }
        if (commit) umma_commit(smem_u32(&dummy[cslot]));
        if (commit2) umma_commit(smem_u32(&dummy[(cslot + 4) & 7]));
      }
      if (elect) __syncwarp();
      cslot = (cslot + 1) & 7;
      if (++sa == 8) sa = 0;
      if (++inchain == 3) { inchain = 0; acc = (acc + 1) & 3; }
    }
    if (lane == 0) {
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N, int mode>
void run(long long* d) {
  const int groups = 2048;
  const int smem = 8 * 18432 + 4 * N * 128 + 2048;
  cudaFuncSetAttribute(k<N, mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, mode><<<148, 128, smem>>>(d, groups);
  k<N, mode><<<148, 128, smem>>>(d, groups);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d mode=%2d (%s%s%s%s%s%s): %6.1f cycles/MMA (%s)\n", N, mode, mode & 1 ? "walk " : "", mode & 2 ? "commit/12 " : "",
         mode & 4 ? "chain36 " : "", mode & 8 ? "elect+wait " : "", mode & 16 ? "2nd-commit/12 " : "", mode & 32 ? "commit/4 " : "",
         (double)h / (groups * 12.0), cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<64, 0>(d); run<64, 1>(d); run<64, 2>(d); run<64, 4>(d); run<64, 8>(d); run<64, 10>(d); run<64, 18>(d); run<64, 32>(d);
  run<64, 15>(d); run<64, 31>(d);
  run<128, 0>(d); run<128, 8>(d); run<128, 10>(d); run<128, 15>(d); run<128, 31>(d);
  run<256, 0>(d); run<256, 8>(d); run<256, 32>(d); run<256, 47>(d);
  return 0;
}
