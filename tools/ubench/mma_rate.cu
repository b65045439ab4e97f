// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, SS mode, cta_group::1) as a function of
// (M, N), operands resident in shared memory (SW128 K-major, contents irrelevant), one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../include -I ../../arbitrarystyletransfer_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc.cuh"
using namespace ast::tc;

template <int M, int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *(volatile uint32_t*)&slot;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(M, N);
    const uint64_t ad = make_sdesc_k128(base), bd = make_sdesc_k128(base + 32768);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j)   // 4 K-slices of one 64-wide K block, as in the conv kernel
        umma_bf16(tm + (uint32_t)((i & 1) * N), ad + j * 2, bd + j * 2, idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int M, int N>
void run(long long* d, int nblk) {
  const int iters = 4096;
  cudaFuncSetAttribute(k<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<M, N><<<nblk, 128, 100 * 1024>>>(d, iters);
  k<M, N><<<nblk, 128, 100 * 1024>>>(d, iters);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = (double)h / (iters * 4.0);
  printf("M=%3d N=%3d blocks=%3d : %7.1f cycles/MMA  -> %6.1f%% of the 8192 flop/clk/SM dense rate  (%s)\n", M, N,
         nblk, cyc, 100.0 * (2.0 * M * N * 16 / cyc) / 8192.0, cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int nblk : {1, 148}) {
    run<128, 16>(d, nblk); run<128, 32>(d, nblk); run<128, 64>(d, nblk); run<128, 128>(d, nblk); run<128, 256>(d, nblk);
    run<64, 64>(d, nblk); run<64, 128>(d, nblk); run<64, 256>(d, nblk);
  }
  return 0;
}
