// Shared device/host helpers for libast_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/ast_b200.h"

#define AST_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

#define AST_CUDA(call)                                       \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

namespace ast {

constexpr int kWarp = 32;

// ---- Welford / Chan running moments (count, mean, M2), fp32 --------------------------------
struct Moments {
  float n, mean, m2;
};

__device__ __forceinline__ Moments moments_merge(Moments a, Moments b) {
  // Chan et al. pairwise combination; safe for empty operands.
  float n = a.n + b.n;
  if (n == 0.f) return Moments{0.f, 0.f, 0.f};
  float d = b.mean - a.mean;
  float fb = b.n / n;
  Moments r;
  r.n = n;
  r.mean = a.mean + d * fb;
  r.m2 = a.m2 + b.m2 + d * d * a.n * fb;
  return r;
}

__device__ __forceinline__ Moments moments_warp_reduce(Moments m) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Moments o;
    o.n = __shfl_xor_sync(0xffffffffu, m.n, off);
    o.mean = __shfl_xor_sync(0xffffffffu, m.mean, off);
    o.m2 = __shfl_xor_sync(0xffffffffu, m.m2, off);
    m = moments_merge(m, o);
  }
  return m;
}

// Per-thread streaming Welford over V interleaved lanes that share one count: one reciprocal
// per vector, three FMAs-class ops per element.
template <int V>
struct WelfordLanes {
  float n;
  float mean[V];
  float m2[V];
  __device__ __forceinline__ void init() {
    n = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) { mean[j] = 0.f; m2[j] = 0.f; }
  }
  __device__ __forceinline__ void push(const float (&x)[V]) {
    n += 1.f;
    float rn = __frcp_rn(n);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float d = x[j] - mean[j];
      mean[j] = fmaf(d, rn, mean[j]);
      m2[j] = fmaf(d, x[j] - mean[j], m2[j]);
    }
  }
  __device__ __forceinline__ Moments fold() const {
    Moments r{n, mean[0], m2[0]};
#pragma unroll
    for (int j = 1; j < V; ++j) r = moments_merge(r, Moments{n, mean[j], m2[j]});
    return r;
  }
};

// Cheaper per-thread streaming moments for HBM-bound kernels: sums of d = x - shift and d^2 with packed
// fp32x2 arithmetic (FADD2 / FFMA2: 1.5 instructions per element instead of ~5 for the Welford update, which
// made the statistics kernels issue-bound).  `shift` is any value of the same row (callers pass its first
// element): it removes a common offset, so m2 = q - s^2/n cancels at most (range/sigma)^2 ulps over the few
// hundred elements one thread sees; threads are then combined with the Chan merge as before.
template <int V>
struct ShiftedLanes {
  static_assert(V % 2 == 0, "pairs");
  float n, shift;
  float2 s[V / 2], q[V / 2];
  __device__ __forceinline__ void init(float shift_) {
    n = 0.f;
    shift = shift_;
#pragma unroll
    for (int j = 0; j < V / 2; ++j) { s[j] = make_float2(0.f, 0.f); q[j] = make_float2(0.f, 0.f); }
  }
  __device__ __forceinline__ void push(const float (&x)[V]) {
    n += 1.f;
    const float2 nk = make_float2(-shift, -shift);
#pragma unroll
    for (int j = 0; j < V / 2; ++j) {
      const float2 d = __fadd2_rn(make_float2(x[2 * j], x[2 * j + 1]), nk);
      s[j] = __fadd2_rn(s[j], d);
      q[j] = __ffma2_rn(d, d, q[j]);
    }
  }
  __device__ __forceinline__ Moments fold() const {
    if (n == 0.f) return Moments{0.f, 0.f, 0.f};
    const float rn = 1.f / n;
    Moments r{0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < V / 2; ++j) {
      r = moments_merge(r, Moments{n, shift + s[j].x * rn, fmaxf(fmaf(-s[j].x, s[j].x * rn, q[j].x), 0.f)});
      r = moments_merge(r, Moments{n, shift + s[j].y * rn, fmaxf(fmaf(-s[j].y, s[j].y * rn, q[j].y), 0.f)});
    }
    return r;
  }
};

// ---- vector I/O ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// 256-bit store (sm_100+): one full 32-byte sector per lane per instruction; p must be 32 B aligned
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// ReLU + round + pack in ONE instruction (cvt.rn.relu: negative -> +0, NaN stays NaN like torch.relu)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Element-type traits for 16-byte vectors: fp32 -> 4 lanes, bf16 -> 8 lanes.
template <bool BF16>
struct Vec16;
template <>
struct Vec16<false> {
  static constexpr int V = 4;
  using elem = float;
  __device__ static __forceinline__ void unpack(uint4 u, float (&x)[4]) {
    x[0] = __uint_as_float(u.x); x[1] = __uint_as_float(u.y);
    x[2] = __uint_as_float(u.z); x[3] = __uint_as_float(u.w);
  }
  __device__ static __forceinline__ uint4 pack(const float (&x)[4]) {
    return make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]),
                      __float_as_uint(x[3]));
  }
  __device__ static __forceinline__ float load1(const void* p, int64_t i) {
    return reinterpret_cast<const float*>(p)[i];
  }
  __device__ static __forceinline__ void store1(void* p, int64_t i, float v) {
    reinterpret_cast<float*>(p)[i] = v;
  }
};
template <>
struct Vec16<true> {
  static constexpr int V = 8;
  using elem = __nv_bfloat16;
  __device__ static __forceinline__ void unpack(uint4 u, float (&x)[8]) {
    x[0] = bf16lo(u.x); x[1] = bf16hi(u.x); x[2] = bf16lo(u.y); x[3] = bf16hi(u.y);
    x[4] = bf16lo(u.z); x[5] = bf16hi(u.z); x[6] = bf16lo(u.w); x[7] = bf16hi(u.w);
  }
  __device__ static __forceinline__ uint4 pack(const float (&x)[8]) {
    return make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                      pack_bf16(x[6], x[7]));
  }
  __device__ static __forceinline__ float load1(const void* p, int64_t i) {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  }
  __device__ static __forceinline__ void store1(void* p, int64_t i, float v) {
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  }
};

// ---- 16-bit storage formats of the MobileNet-style path (K4) -------------------------------------------------
// Forward activations (and the weights they meet in the tensor cores) are IEEE fp16: an 11-bit significand.  With
// bf16's 8 bits the ~150 roundings between the image and the deepest encoder tap added up to 5 % (eval) / 7 %
// (train) relative error and to 25 % error on the stem's gradients -- all of it caused by the FORWARD roundings, none
// by the gradients' (oracle simulation); fp16 brings that to < 1 % / 7 %.  Gradients stay bf16:
// they need range, not precision.  Conversions to fp16 saturate (+-65504) instead of overflowing to infinity.
// tcgen05.mma kind::f16 takes fp16 x fp16 or bf16 x bf16 operands, never a mix (measured: an fp16 x bf16 instruction
// descriptor is an illegal instruction), so weight-gradient GEMMs convert their activation operand to bf16 first.
typedef __half act_t;            // the DEFAULT activation type; kernels are templates on AT in {__half, __nv_bfloat16}
typedef __nv_bfloat16 grad_t;
int act_format();                // AST_DT_F16 (default) or AST_DT_BF16: misc.cu, set by ast_set_act_format
// Launch for the process-wide activation format (fp16 by default, bf16 when range matters more than precision:
// ast_set_act_format): inside the macro `AT` is the activation element type.
#define AST_ACT_DISPATCH(...)                                                    \
  do {                                                                           \
    if (ast::act_format() == AST_DT_F16) { using AT = __half; __VA_ARGS__; }     \
    else { using AT = __nv_bfloat16; __VA_ARGS__; }                              \
  } while (0)


__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}

template <typename T>
struct H16;
template <>
struct H16<__nv_bfloat16> : Vec16<true> {
  __device__ static __forceinline__ float2 un2(uint32_t u) { return make_float2(bf16lo(u), bf16hi(u)); }
  __device__ static __forceinline__ uint32_t pk2(float lo, float hi) { return pack_bf16(lo, hi); }
};
template <>
struct H16<__half> {
  static constexpr int V = 8;
  using elem = __half;
  __device__ static __forceinline__ float2 un2(uint32_t u) { return unpack_f16(u); }
  __device__ static __forceinline__ uint32_t pk2(float lo, float hi) { return pack_f16(lo, hi); }
  __device__ static __forceinline__ void unpack(uint4 u, float (&x)[8]) {
    const float2 a = unpack_f16(u.x), b = unpack_f16(u.y), c = unpack_f16(u.z), d = unpack_f16(u.w);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
  }
  __device__ static __forceinline__ uint4 pack(const float (&x)[8]) {
    return make_uint4(pack_f16(x[0], x[1]), pack_f16(x[2], x[3]), pack_f16(x[4], x[5]), pack_f16(x[6], x[7]));
  }
  __device__ static __forceinline__ float load1(const void* p, int64_t i) {
    return __half2float(reinterpret_cast<const __half*>(p)[i]);
  }
  __device__ static __forceinline__ void store1(void* p, int64_t i, float v) {
    reinterpret_cast<uint16_t*>(p)[i] = (uint16_t)(pack_f16(v, 0.f) & 0xffffu);
  }
};
// typed 16-byte loads / stores: the pointer type (act_t / grad_t) selects the conversion
template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&x)[8]) {
  H16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(p)), x);
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&x)[8]) {
  *reinterpret_cast<uint4*>(p) = H16<T>::pack(x);
}
// runtime-selected format (kernels that serve both roles: pointwise GEMM epilogue, layout converters)
__device__ __forceinline__ float2 un2_dt(uint32_t u, int f16) { return f16 ? unpack_f16(u) : make_float2(bf16lo(u), bf16hi(u)); }
__device__ __forceinline__ uint32_t pk2_dt(float lo, float hi, int f16) { return f16 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }
__device__ __forceinline__ void unpack8_dt(uint4 u, float (&x)[8], int f16) {
  if (f16) H16<__half>::unpack(u, x); else Vec16<true>::unpack(u, x);
}
__device__ __forceinline__ uint4 pack8_dt(const float (&x)[8], int f16) {
  return f16 ? H16<__half>::pack(x) : Vec16<true>::pack(x);
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

}  // namespace ast
