// Probe: does tcgen05.mma kind::f16 accept DIFFERENT 16-bit formats for A and B (fp16 x bf16)?  The instruction
// descriptor has separate a_format / b_format fields (bits 7-9 / 10-12; 0 = f16, 1 = bf16).  One M128 x N64 x K16 MMA
// per format pair, operands in the no-swizzle canonical K-major layout, result checked against the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../include -I ../../arbitrarystyletransfer_b200/csrc mixed_fmt.cu -o mixed_fmt
#include <cstdio>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "tc.cuh"
using namespace ast::tc;

constexpr int M = 128, N = 64, K = 16, LBO = 128, SBO = (K / 8) * 128;

__global__ void __launch_bounds__(128, 1) probe(const float* a, const float* b, float* d, int afmt, int bfmt) {
  __shared__ __align__(128) uint8_t sa[M * K * 2];
  __shared__ __align__(128) uint8_t sb[N * K * 2];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    const uint32_t off = (r >> 3) * SBO + (k >> 3) * LBO + (r & 7) * 16 + (k & 7) * 2;
    if (afmt == 0) *reinterpret_cast<__half*>(sa + off) = __float2half_rn(a[i]);
    else *reinterpret_cast<__nv_bfloat16*>(sa + off) = __float2bfloat16_rn(a[i]);
  }
  for (int i = threadIdx.x; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    const uint32_t off = (r >> 3) * SBO + (k >> 3) * LBO + (r & 7) * 16 + (k & 7) * 2;
    if (bfmt == 0) *reinterpret_cast<__half*>(sb + off) = __float2half_rn(b[i]);
    else *reinterpret_cast<__nv_bfloat16*>(sb + off) = __float2bfloat16_rn(b[i]);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<64>(smem_u32(&slot));
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *(volatile uint32_t*)&slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    umma_bf16(tm, make_sdesc_k_noswizzle(smem_u32(sa), LBO, SBO), make_sdesc_k_noswizzle(smem_u32(sb), LBO, SBO), idesc, 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[(warp * 32 + lane) * N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<64>(tm); }
}

static float rnd(float x, int fmt) {
  return fmt == 0 ? __half2float(__float2half_rn(x)) : __bfloat162float(__float2bfloat16_rn(x));
}

int main() {
  float ha[M * K], hb[N * K], hd[M * N];
  srand(1);
  for (float& x : ha) x = (rand() % 2001 - 1000) / 1000.f * 1.37f;
  for (float& x : hb) x = (rand() % 2001 - 1000) / 1000.f * 0.83f;
  float *a, *b, *d;
  cudaMalloc(&a, sizeof(ha)); cudaMalloc(&b, sizeof(hb)); cudaMalloc(&d, sizeof(hd));
  cudaMemcpy(a, ha, sizeof(ha), cudaMemcpyHostToDevice);
  cudaMemcpy(b, hb, sizeof(hb), cudaMemcpyHostToDevice);
  const char* nm[2] = {"f16", "bf16"};
  for (int af = 0; af < 2; ++af)
    for (int bf = 0; bf < 2; ++bf) {
      cudaMemset(d, 0, sizeof(hd));
      probe<<<1, 128>>>(a, b, d, af, bf);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(hd, d, sizeof(hd), cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
          double r = 0;
          for (int k = 0; k < K; ++k) r += (double)rnd(ha[i * K + k], af) * (double)rnd(hb[j * K + k], bf);
          worst = fmax(worst, fabs(r - hd[i * N + j]));
        }
      printf("A %-4s x B %-4s : max abs err %.3e  (%s)\n", nm[af], nm[bf], worst, cudaGetErrorString(e));
    }
  return 0;
}
