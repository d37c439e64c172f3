// Shared helpers for libmudiff_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/mudiff_b200.h"

#define MUDIFF_NUM_SMS 148

extern int64_t g_mudiff_launches;   // defined in abi.cu

static inline int mudiff_launch_status() {
  ++g_mudiff_launches;
  cudaError_t e = cudaPeekAtLastError();   // do not clear sticky state of other libs
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  return 0;
}

// ---- storage <-> fp32 -----------------------------------------------------------
template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

// A 16-byte vector of T viewed as floats.
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[16 / sizeof(T)]) {
  constexpr int N = 16 / sizeof(T);
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = Cvt<T>::to_f(e[i]);
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[16 / sizeof(T)]) {
  constexpr int N = 16 / sizeof(T);
  uint4 raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = Cvt<T>::from_f(v[i]);
  *reinterpret_cast<uint4*>(p) = raw;
}

// fast variants for bf16 outputs: silu(x) = h + h*tanh(h), h = x/2 -> ONE MUFU op (tanh.approx, rel err ~2^-11)
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16 specialisations: paired conversions (F2FP pack on the ALU pipe instead of eight F2F on the XU pipe)
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}

// packed fp32 pairs: Blackwell's FFMA2 (fma.rn.f32x2) issues two fp32 FMAs per instruction - the FFMA-issue-bound
// kernels (1-channel stems, head conv, FIR) keep their accumulators and weights as register pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

__device__ __forceinline__ float silu_f(float x) { const float h = 0.5f * x; return fmaf(h, tanh_approx(h), h); }
__device__ __forceinline__ float sigmoid_f(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
// exact-ish variants for the fp32 parity path
__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case MUDIFF_ACT_SILU: return silu_exact(v);
    case MUDIFF_ACT_SIGMOID: return sigmoid_exact(v);
    case MUDIFF_ACT_TANH: return tanhf(v);
    case MUDIFF_ACT_LRELU: return v > 0.f ? v : 0.2f * v;
    default: return v;
  }
}

static inline int grid_for(int64_t work, int block, int max_blocks = MUDIFF_NUM_SMS * 16) {
  int64_t g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}
