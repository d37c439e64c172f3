// Memory-bound glue kernels: fused bias+activation, posterior update, gating, concat copy,
// global average pool, tanh, small dense layers, timestep embedding, PixelNorm, row softmax.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------
// fused_bias_act  (semantics: utils/op/fused_bias_act_kernel.cu:20-51)
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void bias_act_kernel(const T* __restrict__ x, const T* __restrict__ b, const T* __restrict__ ref,
                                T* __restrict__ out, int64_t n, int size_b, int64_t step_b,
                                int act, int grad, float alpha, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = Cvt<T>::to_f(x[i]);
    if (b) v += Cvt<T>::to_f(b[(i / step_b) % size_b]);
    float r = ref ? Cvt<T>::to_f(ref[i]) : 0.f;
    float y;
    int code = act * 10 + grad;
    switch (code) {
      default:
      case 10: y = v; break;
      case 11: y = v; break;
      case 12: y = 0.f; break;
      case 30: y = v > 0.f ? v : v * alpha; break;
      case 31: y = r > 0.f ? v : v * alpha; break;
      case 32: y = 0.f; break;
    }
    out[i] = Cvt<T>::from_f(y * scale);
  }
}

// vectorised forward-only fast path: NCHW contiguous planes, step_b % vec == 0
template <typename T>
__global__ void bias_lrelu_vec_kernel(const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ out,
                                      int64_t nvec, int size_b, int64_t step_b_vec, int act, float alpha, float scale) {
  constexpr int V = 16 / sizeof(T);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[V];
    load_vec<T>(x + i * V, v);
    float bb = b ? Cvt<T>::to_f(b[(i / step_b_vec) % size_b]) : 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float t = v[k] + bb;
      if (act == 3) t = t > 0.f ? t : t * alpha;
      v[k] = t * scale;
    }
    store_vec<T>(out + i * V, v);
  }
}

template <typename T>
int launch_bias_act(const void* x, const void* b, const void* ref, void* out, int64_t n, int size_b, int64_t step_b,
                    int act, int grad, float alpha, float scale, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
  if (grad == 0 && !ref && aligned && n % V == 0 && (!b || step_b % V == 0)) {
    int64_t nvec = n / V;
    bias_lrelu_vec_kernel<T><<<grid_for(nvec, 256), 256, 0, st>>>((const T*)x, (const T*)b, (T*)out, nvec,
                                                                 b ? size_b : 1, b ? step_b / V : 1, act, alpha, scale);
  } else {
    bias_act_kernel<T><<<grid_for(n, 256), 256, 0, st>>>((const T*)x, (const T*)b, (const T*)ref, (T*)out, n,
                                                          b ? size_b : 1, b ? step_b : 1, act, grad, alpha, scale);
  }
  return mudiff_launch_status();
}

// ---------------------------------------------------------------------------------
// posterior update (engine/test.py:150-177), operation order kept for fp32 parity
// ---------------------------------------------------------------------------------
__global__ void posterior_kernel(const float* __restrict__ x01, int64_t s01, const float* __restrict__ x02, int64_t s02,
                                 const float* __restrict__ xt, const float* __restrict__ noise,
                                 const int64_t* __restrict__ t, const float* __restrict__ c1t,
                                 const float* __restrict__ c2t, const float* __restrict__ lvt, int n_steps,
                                 float* __restrict__ out, int batch, int64_t per) {
  const int64_t total = (int64_t)batch * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / per);
    int64_t j = i - (int64_t)b * per;
    int64_t ti = t[b];
    ti = ti < 0 ? 0 : (ti >= n_steps ? n_steps - 1 : ti);
    float c1 = c1t[ti], c2 = c2t[ti];
    float x = xt[i];
    float m1 = __fadd_rn(__fmul_rn(c1, x01[b * s01 + j]), __fmul_rn(c2, x));
    float m2 = __fadd_rn(__fmul_rn(c1, x02[b * s02 + j]), __fmul_rn(c2, x));
    float mean = __fdiv_rn(__fadd_rn(m1, m2), 2.0f);
    float mask = (ti == 0) ? 0.f : 1.f;
    float sig = expf(0.5f * lvt[ti]);
    out[i] = __fadd_rn(mean, __fmul_rn(__fmul_rn(mask, sig), noise[i]));
  }
}

// ---------------------------------------------------------------------------------
// gating / residual / concat glue (NHWC, 16-byte channel vectors)
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void gate_mul_kernel(const T* __restrict__ a, int a_ld, const T* __restrict__ b, int b_ld,
                                T* __restrict__ out, int out_ld, int64_t pixels, int c) {
  constexpr int V = 16 / sizeof(T);
  const int cv = c / V;
  const int64_t total = pixels * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / cv; int cc = (int)(i % cv) * V;
    float va[V], vb[V];
    load_vec<T>(a + p * a_ld + cc, va);
    load_vec<T>(b + p * b_ld + cc, vb);
#pragma unroll
    for (int k = 0; k < V; ++k) va[k] *= vb[k];
    store_vec<T>(out + p * out_ld + cc, va);
  }
}

template <typename T>
__global__ void gate_blend_kernel(const T* __restrict__ g, int g_ld, const T* __restrict__ a, int a_ld,
                                  const T* __restrict__ b, int b_ld, T* __restrict__ out, int out_ld,
                                  int64_t pixels, int c) {
  constexpr int V = 16 / sizeof(T);
  const int cv = c / V;
  const int64_t total = pixels * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / cv; int cc = (int)(i % cv) * V;
    float vg[V], va[V], vb[V];
    load_vec<T>(g + p * g_ld + cc, vg);
    load_vec<T>(a + p * a_ld + cc, va);
    load_vec<T>(b + p * b_ld + cc, vb);
#pragma unroll
    for (int k = 0; k < V; ++k) va[k] = vg[k] * va[k] + (1.0f - vg[k]) * vb[k];
    store_vec<T>(out + p * out_ld + cc, va);
  }
}

template <typename T>
__global__ void add_scale_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                 int64_t nvec, float scale) {
  constexpr int V = 16 / sizeof(T);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float va[V], vb[V];
    load_vec<T>(a + i * V, va);
    load_vec<T>(b + i * V, vb);
#pragma unroll
    for (int k = 0; k < V; ++k) va[k] = (va[k] + vb[k]) * scale;
    store_vec<T>(out + i * V, va);
  }
}

template <typename TS, typename TD>
__global__ void copy_channels_kernel(const TS* __restrict__ src, int src_ld, TD* __restrict__ dst, int dst_ld,
                                     int64_t pixels, int c) {
  const int64_t total = pixels * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / c; int cc = (int)(i % c);
    dst[p * dst_ld + cc] = Cvt<TD>::from_f(Cvt<TS>::to_f(src[p * src_ld + cc]));
  }
}

template <typename T>
__global__ void copy_channels_vec_kernel(const T* __restrict__ src, int src_ld, T* __restrict__ dst, int dst_ld,
                                         int64_t pixels, int c) {
  constexpr int V = 16 / sizeof(T);
  const int cv = c / V;
  const int64_t total = pixels * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / cv; int cc = (int)(i % cv) * V;
    *reinterpret_cast<uint4*>(dst + p * dst_ld + cc) = *reinterpret_cast<const uint4*>(src + p * src_ld + cc);
  }
}

template <typename TS, typename TD>
__global__ void tanh_kernel(const TS* __restrict__ x, TD* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = Cvt<TD>::from_f(tanhf(Cvt<TS>::to_f(x[i])));
}

// ---------------------------------------------------------------------------------
// small dense layers: one warp per (output feature j, chunk of 8 batch rows): J * ceil(B/8) independent warps
// (the layers are a serial chain of tiny GEMMs, so what matters is the latency of one launch: short
// dependent-load chains and enough warps to cover the machine), 4 k-steps of loads in flight per lane.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ in, int in_ld, const float* __restrict__ w,
                              const float* __restrict__ bias, float* __restrict__ out, int out_ld,
                              int batch, int k, int j_total, int chunks, int act_in, int act_out) {
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (int64_t)j_total * chunks) return;
  const int j = (int)(gw / chunks);
  const int b0 = (int)(gw % chunks) * 8;
  const float* wr = w + (int64_t)j * k;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int kk = lane; kk < k; kk += 128) {
    float wv[4], xv[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kx = kk + 32 * u;
      wv[u] = kx < k ? wr[kx] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[u][i] = (kx < k && b0 + i < batch) ? in[(int64_t)(b0 + i) * in_ld + kx] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(apply_act(xv[u][i], act_in), wv[u], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (lane == 0) {
    const float bj = bias ? bias[j] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (b0 + i < batch) out[(int64_t)(b0 + i) * out_ld + j] = apply_act(acc[i] + bj, act_out);
  }
}

__global__ void temb_kernel(const int64_t* __restrict__ t, float* __restrict__ out, int batch, int dim, float max_pos) {
  const int half = dim / 2;
  const int total = batch * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int b = i / dim, d = i % dim;
    float v = 0.f;
    if (d < 2 * half) {
      int f = d < half ? d : d - half;
      // layers.py:469-475: exp(arange(half) * -(log(max)/(half-1))) * t
      float freq = expf((float)f * -(logf(max_pos) / (float)(half - 1)));
      float a = (float)t[b] * freq;
      v = d < half ? sinf(a) : cosf(a);
    }
    out[i] = v;
  }
}

__global__ void pixelnorm_kernel(const float* __restrict__ z, float* __restrict__ out, int batch, int dim) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) { float v = z[(int64_t)b * dim + i]; s += v * v; }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float inv = 1.0f / sqrtf(red[0] / (float)dim + 1e-8f);
  for (int i = threadIdx.x; i < dim; i += blockDim.x) out[(int64_t)b * dim + i] = z[(int64_t)b * dim + i] * inv;
}

// ---------------------------------------------------------------------------------
// row softmax.  Fast path: cols == 256 * VEC * R (R <= 4): the row lives in registers (16-byte loads),
// two block reductions, one pass over HBM each way.  Generic path: row staged in shared memory.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < (blockDim.x >> 5) ? red[threadIdx.x & 31] : -INFINITY;
  for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < (blockDim.x >> 5) ? red[threadIdx.x & 31] : 0.f;
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  __syncthreads();
  return r;
}

template <typename T, int R>
__global__ void __launch_bounds__(256) softmax_rows_reg_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                               int64_t rows, int cols, float scale) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float red[32];
  const float sl2 = scale * 1.4426950408889634f;      // exp(s*x - m) = exp2(s*log2e*x - m')
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* xr = x + r * (int64_t)cols;
    T* yr = y + r * (int64_t)cols;
    float v[R][V];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      load_vec<T>(xr + (j * 256 + threadIdx.x) * V, v[j]);
#pragma unroll
      for (int k = 0; k < V; ++k) { v[j][k] *= sl2; mx = fmaxf(mx, v[j][k]); }
    }
    mx = block_reduce_max(mx, red);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int k = 0; k < V; ++k) { v[j][k] = exp2f(v[j][k] - mx); s += v[j][k]; }
    s = block_reduce_sum(s, red);
    const float inv = 1.0f / s;
#pragma unroll
    for (int j = 0; j < R; ++j) {
#pragma unroll
      for (int k = 0; k < V; ++k) v[j][k] *= inv;
      store_vec<T>(yr + (j * 256 + threadIdx.x) * V, v[j]);
    }
  }
}

template <typename T>
__global__ void softmax_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int cols, float scale) {
  extern __shared__ float row[];            // cols floats
  __shared__ float red[32];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* xr = x + r * (int64_t)cols;
    T* yr = y + r * (int64_t)cols;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float v = Cvt<T>::to_f(xr[i]) * scale;
      row[i] = v;
      mx = fmaxf(mx, v);
    }
    mx = block_reduce_max(mx, red);
    float s = 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float e = expf(row[i] - mx);
      row[i] = e;
      s += e;
    }
    s = block_reduce_sum(s, red);
    const float inv = 1.0f / s;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) yr[i] = Cvt<T>::from_f(row[i] * inv);
    __syncthreads();
  }
}

template <typename T>
int launch_softmax(const void* x, void* y, int64_t rows, int cols, float scale, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  const bool al = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
  int grid = (int)(rows < MUDIFF_NUM_SMS * 16 ? rows : MUDIFF_NUM_SMS * 16);
  if (al && cols % (256 * V) == 0 && cols / (256 * V) <= 4) {
    switch (cols / (256 * V)) {
      case 1: softmax_rows_reg_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)x, (T*)y, rows, cols, scale); break;
      case 2: softmax_rows_reg_kernel<T, 2><<<grid, 256, 0, st>>>((const T*)x, (T*)y, rows, cols, scale); break;
      case 3: softmax_rows_reg_kernel<T, 3><<<grid, 256, 0, st>>>((const T*)x, (T*)y, rows, cols, scale); break;
      default: softmax_rows_reg_kernel<T, 4><<<grid, 256, 0, st>>>((const T*)x, (T*)y, rows, cols, scale); break;
    }
    return mudiff_launch_status();
  }
  size_t smem = sizeof(float) * (size_t)cols;
  if (smem > 48 * 1024) cudaFuncSetAttribute(softmax_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  softmax_rows_kernel<T><<<grid, 256, smem, st>>>((const T*)x, (T*)y, rows, cols, scale);
  return mudiff_launch_status();
}

}  // namespace

#define DISPATCH3(dtype, FN, ...)                                   \
  switch (dtype) {                                                   \
    case MUDIFF_F32: return FN<float>(__VA_ARGS__);                  \
    case MUDIFF_BF16: return FN<__nv_bfloat16>(__VA_ARGS__);         \
    case MUDIFF_F16: return FN<__half>(__VA_ARGS__);                 \
    default: return MUDIFF_EINVAL;                                   \
  }

extern "C" int mudiff_fused_bias_act(const void* x, const void* bias, const void* ref, void* out, int dtype,
                                     int64_t n, int size_b, int64_t step_b, int act, int grad,
                                     float alpha, float scale, void* stream) {
  if (n < 0 || (act != 1 && act != 3) || grad < 0 || grad > 2) return MUDIFF_EINVAL;
  if (bias && (size_b < 1 || step_b < 1)) return MUDIFF_EINVAL;
  if (n == 0) return 0;
  if (!x || !out) return MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH3(dtype, launch_bias_act, x, bias, ref, out, n, size_b, step_b, act, grad, alpha, scale, st);
}

// minibatch standard deviation (backbones/discriminator.py:243-250): one block per sub-batch index m
template <typename T>
__global__ void __launch_bounds__(256) mbstd_kernel(const T* __restrict__ x, int ld, T* __restrict__ out, int out_ld, int out_c,
                                                    int n_sub, int group, int channels, int hw) {
  const int m = blockIdx.x;
  const int64_t per = (int64_t)channels * hw;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < per; i += blockDim.x) {
    const int64_t pix = i / channels;
    const int c = (int)(i - pix * channels);
    float v[8], mean = 0.f;
    for (int g = 0; g < group; ++g) {
      v[g] = Cvt<T>::to_f(x[((int64_t)(g * n_sub + m) * hw + pix) * ld + c]);
      mean += v[g];
    }
    mean /= (float)group;
    float var = 0.f;
    for (int g = 0; g < group; ++g) var += (v[g] - mean) * (v[g] - mean);
    acc += sqrtf(var / (float)group + 1e-8f);
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {                       // fixed-order tree: deterministic
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const T sv = Cvt<T>::from_f(red[0] / (float)per);
  for (int64_t i = threadIdx.x; i < (int64_t)group * hw; i += blockDim.x) {
    const int g = (int)(i / hw);
    const int64_t pix = i - (int64_t)g * hw;
    out[((int64_t)(g * n_sub + m) * hw + pix) * out_ld + out_c] = sv;
  }
}

template <typename T>
static int launch_mbstd(const void* x, int ld, void* out, int out_ld, int out_c, int n_sub, int group, int channels, int hw,
                        cudaStream_t st) {
  mbstd_kernel<T><<<n_sub, 256, 0, st>>>((const T*)x, ld, (T*)out, out_ld, out_c, n_sub, group, channels, hw);
  return mudiff_launch_status();
}

extern "C" int mudiff_minibatch_stddev(const void* x, int ld, void* out, int out_ld, int out_c, int dtype,
                                       int batch, int group, int channels, int hw, void* stream) {
  if (!x || !out || batch <= 0 || group <= 0 || group > 8 || channels <= 0 || hw <= 0 || batch % group) return MUDIFF_EINVAL;
  if (ld < channels || out_c < 0 || out_c >= out_ld) return MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH3(dtype, launch_mbstd, x, ld, out, out_ld, out_c, batch / group, group, channels, hw, st);
}

extern "C" int mudiff_posterior_update(const float* x01, int64_t x01_bstride, const float* x02, int64_t x02_bstride,
                                       const float* xt, const float* noise, const int64_t* t,
                                       const float* coef1, const float* coef2, const float* logvar, int n_steps,
                                       float* out, int batch, int64_t per_sample, void* stream) {
  if (batch < 0 || per_sample < 0 || n_steps < 1) return MUDIFF_EINVAL;
  if (batch == 0 || per_sample == 0) return 0;
  if (!x01 || !x02 || !xt || !noise || !t || !coef1 || !coef2 || !logvar || !out) return MUDIFF_EINVAL;
  int64_t total = (int64_t)batch * per_sample;
  posterior_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      x01, x01_bstride, x02, x02_bstride, xt, noise, t, coef1, coef2, logvar, n_steps, out, batch, per_sample);
  return mudiff_launch_status();
}

template <typename T>
static int launch_gate_mul(const void* a, int a_ld, const void* b, int b_ld, void* out, int out_ld,
                           int64_t pixels, int c, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  if (c % V || a_ld % V || b_ld % V || out_ld % V) return MUDIFF_EUNSUPPORTED;
  gate_mul_kernel<T><<<grid_for(pixels * (c / V), 256), 256, 0, st>>>((const T*)a, a_ld, (const T*)b, b_ld, (T*)out, out_ld, pixels, c);
  return mudiff_launch_status();
}
extern "C" int mudiff_gate_mul(const void* a, int a_ld, const void* b, int b_ld, void* out, int out_ld,
                               int dtype, int64_t pixels, int c, void* stream) {
  if (pixels <= 0 || c <= 0) return pixels == 0 ? 0 : MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH3(dtype, launch_gate_mul, a, a_ld, b, b_ld, out, out_ld, pixels, c, st);
}

template <typename T>
static int launch_gate_blend(const void* g, int g_ld, const void* a, int a_ld, const void* b, int b_ld, void* out,
                             int out_ld, int64_t pixels, int c, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  if (c % V || g_ld % V || a_ld % V || b_ld % V || out_ld % V) return MUDIFF_EUNSUPPORTED;
  gate_blend_kernel<T><<<grid_for(pixels * (c / V), 256), 256, 0, st>>>((const T*)g, g_ld, (const T*)a, a_ld, (const T*)b, b_ld,
                                                                       (T*)out, out_ld, pixels, c);
  return mudiff_launch_status();
}
extern "C" int mudiff_gate_blend(const void* g, int g_ld, const void* a, int a_ld, const void* b, int b_ld,
                                 void* out, int out_ld, int dtype, int64_t pixels, int c, void* stream) {
  if (pixels <= 0 || c <= 0) return pixels == 0 ? 0 : MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH3(dtype, launch_gate_blend, g, g_ld, a, a_ld, b, b_ld, out, out_ld, pixels, c, st);
}

template <typename T>
static int launch_add_scale(const void* a, const void* b, void* out, int64_t n, float scale, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  if (n % V) return MUDIFF_EUNSUPPORTED;
  add_scale_kernel<T><<<grid_for(n / V, 256), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n / V, scale);
  return mudiff_launch_status();
}
extern "C" int mudiff_add_scale(const void* a, const void* b, void* out, int dtype, int64_t n, float scale, void* stream) {
  if (n <= 0) return n == 0 ? 0 : MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH3(dtype, launch_add_scale, a, b, out, n, scale, st);
}

// fp32 -> three bf16 planes (hi | mid | lo along the channel axis): x == hi + mid + lo exactly (24 mantissa bits = 3 x 8).
// Feeds the tcgen05 conv on the fp32 parity path (ops.conv: six bf16 products per fp32 product, fp32 accumulation).
// layout 0: [lo | mid | hi] (3c per pixel, the activation side); layout 1: [lo | mid mid | hi hi hi] (6c per pixel: the
// "weight" side when the second operand is an activation too - attention scores / PV / V^T).  SMALL terms first: the conv
// walks K in this order, so the products of relative size 2^-16 and 2^-8 are accumulated while the TMEM accumulator is
// still small and the hi x hi products come last - the tensor core's fp32 accumulation truncates, and its error is
// proportional to the accumulator's magnitude at every step.
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, int ld, __nv_bfloat16* __restrict__ out,
                                                    int64_t pixels, int c, int layout) {
  const int cv = c / 8;
  const int64_t total = pixels * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / cv;
    const int c0 = (int)(i - pix * cv) * 8;
    const float4 a = *reinterpret_cast<const float4*>(x + pix * ld + c0);
    const float4 b = *reinterpret_cast<const float4*>(x + pix * ld + c0 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, mid, lo;
    __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(&hi);
    __nv_bfloat16* pm = reinterpret_cast<__nv_bfloat16*>(&mid);
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(&lo);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
      const float r1 = v[j] - __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(m);
      ph[j] = h; pm[j] = m; pl[j] = __float2bfloat16_rn(r2);
    }
    if (layout == 0) {
      __nv_bfloat16* o = out + pix * (3 * (int64_t)c) + c0;
      *reinterpret_cast<uint4*>(o) = lo;
      *reinterpret_cast<uint4*>(o + c) = mid;
      *reinterpret_cast<uint4*>(o + 2 * c) = hi;
    } else {
      __nv_bfloat16* o = out + pix * (6 * (int64_t)c) + c0;
      *reinterpret_cast<uint4*>(o) = lo;
      *reinterpret_cast<uint4*>(o + c) = mid;
      *reinterpret_cast<uint4*>(o + 2 * c) = mid;
      *reinterpret_cast<uint4*>(o + 3 * c) = hi;
      *reinterpret_cast<uint4*>(o + 4 * c) = hi;
      *reinterpret_cast<uint4*>(o + 5 * c) = hi;
    }
  }
}

extern "C" int mudiff_split3_bf16(const float* x, int ld, void* out, int64_t pixels, int c, int layout, void* stream) {
  if (!x || !out || pixels <= 0 || c <= 0 || (layout != 0 && layout != 1)) return MUDIFF_EINVAL;
  if (c % 8 || ld % 4 || ((uintptr_t)x % 16) || ((uintptr_t)out % 16)) return MUDIFF_EUNSUPPORTED;
  split3_kernel<<<grid_for(pixels * (c / 8), 256), 256, 0, (cudaStream_t)stream>>>(x, ld, (__nv_bfloat16*)out, pixels, c, layout);
  return mudiff_launch_status();
}

extern "C" int mudiff_copy_channels(const void* src, int src_ld, int src_dtype, void* dst, int dst_ld, int dst_dtype,
                                    int64_t pixels, int c, void* stream) {
  if (pixels <= 0 || c <= 0) return pixels == 0 ? 0 : MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = pixels * c;
  const bool al = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  if (src_dtype == dst_dtype && al) {
    if (src_dtype == MUDIFF_F32 && c % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0) {
      copy_channels_vec_kernel<float><<<grid_for(total / 4, 256), 256, 0, st>>>((const float*)src, src_ld, (float*)dst, dst_ld, pixels, c);
      return mudiff_launch_status();
    }
    if (src_dtype == MUDIFF_BF16 && c % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0) {
      copy_channels_vec_kernel<__nv_bfloat16><<<grid_for(total / 8, 256), 256, 0, st>>>((const __nv_bfloat16*)src, src_ld, (__nv_bfloat16*)dst, dst_ld, pixels, c);
      return mudiff_launch_status();
    }
  }
  int grid = grid_for(total, 256);
#define CC(TS, TD) copy_channels_kernel<TS, TD><<<grid, 256, 0, st>>>((const TS*)src, src_ld, (TD*)dst, dst_ld, pixels, c)
  if (src_dtype == MUDIFF_F32 && dst_dtype == MUDIFF_F32) CC(float, float);
  else if (src_dtype == MUDIFF_F32 && dst_dtype == MUDIFF_BF16) CC(float, __nv_bfloat16);
  else if (src_dtype == MUDIFF_BF16 && dst_dtype == MUDIFF_F32) CC(__nv_bfloat16, float);
  else if (src_dtype == MUDIFF_BF16 && dst_dtype == MUDIFF_BF16) CC(__nv_bfloat16, __nv_bfloat16);
  else return MUDIFF_EUNSUPPORTED;
#undef CC
  return mudiff_launch_status();
}

extern "C" int mudiff_tanh(const void* x, void* out, int dtype_in, int dtype_out, int64_t n, void* stream) {
  if (n <= 0) return n == 0 ? 0 : MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n, 256);
  if (dtype_in == MUDIFF_F32 && dtype_out == MUDIFF_F32) tanh_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (float*)out, n);
  else if (dtype_in == MUDIFF_BF16 && dtype_out == MUDIFF_F32) tanh_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)out, n);
  else if (dtype_in == MUDIFF_BF16 && dtype_out == MUDIFF_BF16) tanh_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, n);
  else return MUDIFF_EUNSUPPORTED;
  return mudiff_launch_status();
}

extern "C" int mudiff_linear(const float* in, int in_ld, const float* w, const float* bias, float* out, int out_ld,
                             int batch, int k, int j, int act_in, int act_out, void* stream) {
  if (batch <= 0 || k <= 0 || j <= 0) return MUDIFF_EINVAL;
  const int chunks = (batch + 7) / 8;
  const int64_t warps = (int64_t)j * chunks;
  const int64_t grid = (warps + 7) / 8;
  if (grid >= (1LL << 31)) return MUDIFF_EUNSUPPORTED;
  linear_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(in, in_ld, w, bias, out, out_ld, batch, k, j, chunks, act_in, act_out);
  return mudiff_launch_status();
}

extern "C" int mudiff_timestep_embedding(const int64_t* t, float* out, int batch, int dim, float max_positions, void* stream) {
  if (batch <= 0 || dim < 4) return MUDIFF_EINVAL;
  temb_kernel<<<grid_for((int64_t)batch * dim, 128), 128, 0, (cudaStream_t)stream>>>(t, out, batch, dim, max_positions);
  return mudiff_launch_status();
}

extern "C" int mudiff_pixelnorm(const float* z, float* out, int batch, int dim, void* stream) {
  if (batch <= 0 || dim <= 0) return MUDIFF_EINVAL;
  pixelnorm_kernel<<<batch, 128, 0, (cudaStream_t)stream>>>(z, out, batch, dim);
  return mudiff_launch_status();
}

extern "C" int mudiff_softmax_rows(const void* x, void* y, int dtype, int64_t rows, int cols, float scale, void* stream) {
  if (rows <= 0 || cols <= 0) return rows == 0 ? 0 : MUDIFF_EINVAL;
  if (cols > 49152) return MUDIFF_EUNSUPPORTED;      // generic path stages the row in <= 192 KB smem
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MUDIFF_F32) return launch_softmax<float>(x, y, rows, cols, scale, st);
  if (dtype == MUDIFF_BF16) return launch_softmax<__nv_bfloat16>(x, y, rows, cols, scale, st);
  return MUDIFF_EUNSUPPORTED;
}

extern "C" int mudiff_zero(void* p, int64_t nbytes, void* stream) {
  if (nbytes < 0) return MUDIFF_EINVAL;
  if (nbytes == 0) return 0;
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)nbytes, (cudaStream_t)stream);
  ++g_mudiff_launches;
  return (int)e;
}
