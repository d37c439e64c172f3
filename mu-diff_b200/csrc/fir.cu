// FIR resampling kernels: upfirdn2d (pad -> zero-insert upsample -> FIR -> decimate).
// Semantics follow utils/op/upfirdn2d.py:201-242 / upfirdn2d_kernel.cu:211-371 of the
// reference; the implementation is new:
//   * NHWC (minor % vec == 0): one 16-byte channel vector per thread, polyphase tap walk
//     (only taps that hit a non-inserted sample are visited), fully coalesced;
//   * NCHW (minor == 1): shared-memory staged input tile + flipped taps, 32x32 output tile;
//   * anything else: generic per-element gather.
#include <stdlib.h>
#include "common.cuh"

namespace {

__host__ __device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (q * b > a) ? q - 1 : q;
}

struct FirP {
  int64_t major;
  int in_h, in_w, minor, kh, kw, up_x, up_y, down_x, down_y, px0, py0, out_h, out_w;
};

#define FIR_MAX_TAPS 1024

// first tap index k >= 0 such that (base + k) % up == 0, where base may be negative
__device__ __forceinline__ int first_tap(int base, int up) {
  int r = base % up;
  if (r < 0) r += up;
  return r == 0 ? 0 : up - r;
}

template <typename T>
__global__ void fir_nhwc_vec_kernel(const T* __restrict__ in, T* __restrict__ out,
                                    const float* __restrict__ kern, FirP p) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float sk[FIR_MAX_TAPS];
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    int ky = i / p.kw, kx = i % p.kw;
    sk[i] = kern[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];   // flip: true convolution
  }
  __syncthreads();
  const int cv = p.minor / V;
  const int64_t total = p.major * p.out_h * p.out_w * cv;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % cv);
    int64_t r = idx / cv;
    int ox = (int)(r % p.out_w); r /= p.out_w;
    int oy = (int)(r % p.out_h);
    int64_t m = r / p.out_h;
    const int by = oy * p.down_y - p.py0, bx = ox * p.down_x - p.px0;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int ky = first_tap(by, p.up_y); ky < p.kh; ky += p.up_y) {
      int iy = (by + ky) / p.up_y;           // exact: (by+ky) % up == 0 (may be negative)
      if (by + ky < 0 || iy >= p.in_h) continue;
      for (int kx = first_tap(bx, p.up_x); kx < p.kw; kx += p.up_x) {
        int ix = (bx + kx) / p.up_x;
        if (bx + kx < 0 || ix >= p.in_w) continue;
        float w = sk[ky * p.kw + kx];
        float v[V];
        load_vec<T>(in + ((m * p.in_h + iy) * p.in_w + ix) * p.minor + c * V, v);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(w, v[i], acc[i]);
      }
    }
    store_vec<T>(out + ((m * p.out_h + oy) * p.out_w + ox) * p.minor + c * V, acc);
  }
}

template <typename T>
__global__ void fir_generic_kernel(const T* __restrict__ in, T* __restrict__ out,
                                   const float* __restrict__ kern, FirP p) {
  __shared__ float sk[FIR_MAX_TAPS];
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    int ky = i / p.kw, kx = i % p.kw;
    sk[i] = kern[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  __syncthreads();
  const int64_t total = p.major * p.out_h * p.out_w * p.minor;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % p.minor);
    int64_t r = idx / p.minor;
    int ox = (int)(r % p.out_w); r /= p.out_w;
    int oy = (int)(r % p.out_h);
    int64_t m = r / p.out_h;
    const int by = oy * p.down_y - p.py0, bx = ox * p.down_x - p.px0;
    float acc = 0.f;
    for (int ky = first_tap(by, p.up_y); ky < p.kh; ky += p.up_y) {
      int iy = (by + ky) / p.up_y;
      if (by + ky < 0 || iy >= p.in_h) continue;
      for (int kx = first_tap(bx, p.up_x); kx < p.kw; kx += p.up_x) {
        int ix = (bx + kx) / p.up_x;
        if (bx + kx < 0 || ix >= p.in_w) continue;
        acc = fmaf(sk[ky * p.kw + kx],
                   Cvt<T>::to_f(in[((m * p.in_h + iy) * p.in_w + ix) * p.minor + c]), acc);
      }
    }
    out[idx] = Cvt<T>::from_f(acc);
  }
}

// NCHW plane kernel: one block = one 32x32 output tile of one (n,c) plane.  The input
// footprint of the tile is staged in shared memory with coalesced row loads.
#define FIR_TILE 32
template <typename T>
__global__ void fir_nchw_tiled_kernel(const T* __restrict__ in, T* __restrict__ out,
                                      const float* __restrict__ kern, FirP p, int tin_h, int tin_w) {
  extern __shared__ float smem[];
  float* sk = smem;                         // kh*kw flipped taps
  float* sx = smem + p.kh * p.kw;           // tin_h * tin_w staged input
  const int64_t m = blockIdx.z;
  const int oy0 = blockIdx.y * FIR_TILE, ox0 = blockIdx.x * FIR_TILE;
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    int ky = i / p.kw, kx = i % p.kw;
    sk[i] = kern[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  // lowest input row/col that any tap of this tile can touch (ceil division, may be < 0)
  const int iy_lo = -floordiv(-(oy0 * p.down_y - p.py0), p.up_y);
  const int ix_lo = -floordiv(-(ox0 * p.down_x - p.px0), p.up_x);
  const T* plane = in + m * (int64_t)p.in_h * p.in_w;
  for (int i = threadIdx.x; i < tin_h * tin_w; i += blockDim.x) {
    int ry = i / tin_w, rx = i % tin_w;
    int iy = iy_lo + ry, ix = ix_lo + rx;
    float v = 0.f;
    if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v = Cvt<T>::to_f(plane[(int64_t)iy * p.in_w + ix]);
    sx[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FIR_TILE * FIR_TILE; i += blockDim.x) {
    int oy = oy0 + i / FIR_TILE, ox = ox0 + i % FIR_TILE;
    if (oy >= p.out_h || ox >= p.out_w) continue;
    const int by = oy * p.down_y - p.py0, bx = ox * p.down_x - p.px0;
    float acc = 0.f;
    for (int ky = first_tap(by, p.up_y); ky < p.kh; ky += p.up_y) {
      int ry = floordiv(by + ky, p.up_y) - iy_lo;   // staged zeros cover out-of-range samples
      for (int kx = first_tap(bx, p.up_x); kx < p.kw; kx += p.up_x) {
        int rx = floordiv(bx + kx, p.up_x) - ix_lo;
        acc = fmaf(sk[ky * p.kw + kx], sx[ry * tin_w + rx], acc);
      }
    }
    out[(m * p.out_h + oy) * (int64_t)p.out_w + ox] = Cvt<T>::from_f(acc);
  }
}

// Specialised NHWC kernels for the three modes the generators use (4x4 FIR, up_or_down_sampling.py:200-262):
//   UP=2   (upsample_2d, pad (2,1))      DOWN=2 (downsample_2d, pad (1,1))      UP=DOWN=1 (pre-filter, pad (2,2))
// A thread owns one 16-byte channel vector and computes a 2x2 block of outputs from the NRxNR input window that
// block touches (3x3 / 6x6 / 5x5): 2.25 / 9 / 6.25 vector loads per output instead of 4 / 16 / 16, no
// data-dependent branches (the polyphase structure is resolved at compile time), one window row (NR independent
// 16-byte loads) in flight at a time, YB vertically adjacent blocks per thread so re-read rows hit L1.
// Requires even pads for UP=2 (polyphase alignment); anything else goes to the generic kernels above.
template <int UP, int DOWN> struct Fir4Geom {
  static constexpr int NR = UP == 2 ? 3 : (DOWN == 2 ? 6 : 5);
  // tap index of window row/col r for output offset d in {0,1}, or -1
  static __host__ __device__ constexpr int tap(int d, int r) {
    if (UP == 2) { const int t = r - d; return (t == 0 || t == 1) ? 2 * t + d : -1; }
    if (DOWN == 2) { const int t = r - 2 * d; return (t >= 0 && t < 4) ? t : -1; }
    { const int t = r - d; return (t >= 0 && t < 4) ? t : -1; }
  }
};

template <typename T, int UP, int DOWN, int YB>
__global__ void __launch_bounds__(256, 4) fir4_quad_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                        const float* __restrict__ kern, FirP p, int bxs, int bygs) {
  using G = Fir4Geom<UP, DOWN>;
  constexpr int V = 16 / sizeof(T);
  constexpr int NR = G::NR;
  __shared__ float sk[16];
  if (threadIdx.x < 16) sk[threadIdx.x] = kern[15 - threadIdx.x];        // flipped: true convolution
  __syncthreads();
  float kf[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i] = sk[i];
  const int cv = p.minor / V;
  const int64_t total = p.major * bygs * (int64_t)bxs * cv;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % cv) * V;
  int64_t r_ = idx / cv;
  const int bx = (int)(r_ % bxs); r_ /= bxs;
  const int byg = (int)(r_ % bygs);
  const int64_t m = r_ / bygs;
  const T* inm = in + m * p.in_h * (int64_t)p.in_w * p.minor + c;
  T* outm = out + m * p.out_h * (int64_t)p.out_w * p.minor + c;
  const int ox0 = bx * 2;
  // first input column of the window
  const int ix0 = UP == 2 ? (ox0 - p.px0) / 2 : ox0 * DOWN - p.px0;      // UP=2: px0 even -> exact
#pragma unroll 1
  for (int yb = 0; yb < YB; ++yb) {
    const int oy0 = (byg * YB + yb) * 2;
    if (oy0 >= p.out_h) break;
    const int iy0 = UP == 2 ? (oy0 - p.py0) / 2 : oy0 * DOWN - p.py0;
    f32x2 acc[2][2][V / 2];                    // channel pairs: FFMA2
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int i = 0; i < V / 2; ++i) acc[a][b][i] = pack2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int iy = iy0 + r;
      const bool oky = iy >= 0 && iy < p.in_h;
      uint4 raw[NR];
#pragma unroll
      for (int s2 = 0; s2 < NR; ++s2) {
        const int ix = ix0 + s2;
        raw[s2] = make_uint4(0u, 0u, 0u, 0u);
        if (oky && ix >= 0 && ix < p.in_w) raw[s2] = *reinterpret_cast<const uint4*>(inm + ((int64_t)iy * p.in_w + ix) * p.minor);
      }
#pragma unroll
      for (int s2 = 0; s2 < NR; ++s2) {
        f32x2 v2[V / 2];
        const T* e = reinterpret_cast<const T*>(&raw[s2]);
#pragma unroll
        for (int i = 0; i < V / 2; ++i) v2[i] = pack2(Cvt<T>::to_f(e[2 * i]), Cvt<T>::to_f(e[2 * i + 1]));
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int ky = G::tap(dy, r);
          if (ky < 0) continue;
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int kx = G::tap(dx, s2);
            if (kx < 0) continue;
            const float w = kf[ky * 4 + kx];
            const f32x2 ww = pack2(w, w);
#pragma unroll
            for (int i = 0; i < V / 2; ++i) acc[dy][dx][i] = fma2(ww, v2[i], acc[dy][dx][i]);
          }
        }
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      if (oy0 + dy >= p.out_h) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        if (ox0 + dx >= p.out_w) continue;
        float o[V];
#pragma unroll
        for (int i = 0; i < V / 2; ++i) unpack2(acc[dy][dx][i], o[2 * i], o[2 * i + 1]);
        store_vec<T>(outm + ((int64_t)(oy0 + dy) * p.out_w + ox0 + dx) * p.minor, o);
      }
    }
  }
}

template <typename T, int UP, int DOWN>
int launch_fir4_quad(const void* in, void* out, const float* kern, const FirP& p, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  constexpr int YB = 4;
  const int cv = p.minor / V;
  const int bxs = (p.out_w + 1) / 2;
  const int bygs = ((p.out_h + 1) / 2 + YB - 1) / YB;
  const int64_t total = p.major * bygs * (int64_t)bxs * cv;
  const int64_t blocks = (total + 255) / 256;
  if (blocks >= (1LL << 31)) return MUDIFF_EUNSUPPORTED;
  fir4_quad_kernel<T, UP, DOWN, YB><<<(unsigned)blocks, 256, 0, st>>>((const T*)in, (T*)out, kern, p, bxs, bygs);
  return mudiff_launch_status();
}

// Fused AdaGN + SiLU + FIR for the resample ResBlocks (backbones/layerspp.py:293-305): h = act(AdaGN(x)) is FIR-resampled
// and so is x itself - the reference runs GroupNorm, scale/shift, SiLU and two upfirdn2d calls, i.e. the full-resolution
// normalised tensor is written and read back and x is read twice more.  Here ONE kernel reads x once and writes both
// out_x = FIR(x) and out_h = FIR(act(x * scale[b][c] + shift[b][c])) with the folded AdaGN parameters of
// mudiff_gn_scale_shift.  Zero padding applies AFTER the activation (taps outside the image contribute nothing to
// either output).  Same geometry as fir4_quad_kernel (2x2 output quad per thread) with V = 4 channels per thread so that
// the two accumulator sets fit in registers.
template <typename T, int V> struct VecIO;
template <> struct VecIO<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
    h[0] = __floats2bfloat162_rn(v[0], v[1]); h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = raw;
  }
};
template <> struct VecIO<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <typename T, int UP, int DOWN, int YB>
__global__ void __launch_bounds__(256, 2) fir4_quad_gn_kernel(const T* __restrict__ in, T* __restrict__ out_h, T* __restrict__ out_x,
                                                           const float* __restrict__ kern, const float* __restrict__ table,
                                                           int table_ld, int act, FirP p, int bxs, int bygs) {
  using G = Fir4Geom<UP, DOWN>;
  constexpr int V = 4;
  constexpr int NR = G::NR;
  __shared__ float sk[16];
  if (threadIdx.x < 16) sk[threadIdx.x] = kern[15 - threadIdx.x];        // flipped: true convolution
  __syncthreads();
  float kf[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i] = sk[i];
  const int cv = p.minor / V;
  const int64_t total = p.major * bygs * (int64_t)bxs * cv;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % cv) * V;
  int64_t r_ = idx / cv;
  const int bx = (int)(r_ % bxs); r_ /= bxs;
  const int byg = (int)(r_ % bygs);
  const int64_t m = r_ / bygs;
  float sc[V], sh[V];
  {
    const float* tp = table + ((int64_t)m * table_ld + c) * 2;
#pragma unroll
    for (int i = 0; i < V; ++i) { sc[i] = tp[2 * i]; sh[i] = tp[2 * i + 1]; }
  }
  const T* inm = in + m * p.in_h * (int64_t)p.in_w * p.minor + c;
  const int64_t obase = m * p.out_h * (int64_t)p.out_w * p.minor + c;
  const int ox0 = bx * 2;
  const int ix0 = UP == 2 ? (ox0 - p.px0) / 2 : ox0 * DOWN - p.px0;      // UP=2: px0 even -> exact
#pragma unroll 1
  for (int yb = 0; yb < YB; ++yb) {
    const int oy0 = (byg * YB + yb) * 2;
    if (oy0 >= p.out_h) break;
    const int iy0 = UP == 2 ? (oy0 - p.py0) / 2 : oy0 * DOWN - p.py0;
    float ax[2][2][V], ah[2][2][V];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int i = 0; i < V; ++i) { ax[a][b][i] = 0.f; ah[a][b][i] = 0.f; }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int iy = iy0 + r;
      const bool oky = iy >= 0 && iy < p.in_h;
      float xv[NR][V];
      bool okx[NR];
#pragma unroll
      for (int s2 = 0; s2 < NR; ++s2) {
        const int ix = ix0 + s2;
        okx[s2] = oky && ix >= 0 && ix < p.in_w;
#pragma unroll
        for (int i = 0; i < V; ++i) xv[s2][i] = 0.f;
        if (okx[s2]) VecIO<T, V>::load(inm + ((int64_t)iy * p.in_w + ix) * p.minor, xv[s2]);
      }
#pragma unroll
      for (int s2 = 0; s2 < NR; ++s2) {
        float hv[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float t = fmaf(xv[s2][i], sc[i], sh[i]);
          if (act == MUDIFF_ACT_SILU) t = (sizeof(T) == 4) ? silu_exact(t) : silu_f(t);
          hv[i] = okx[s2] ? t : 0.f;
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int ky = G::tap(dy, r);
          if (ky < 0) continue;
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int kx = G::tap(dx, s2);
            if (kx < 0) continue;
            const float w = kf[ky * 4 + kx];
#pragma unroll
            for (int i = 0; i < V; ++i) {
              ax[dy][dx][i] = fmaf(w, xv[s2][i], ax[dy][dx][i]);
              ah[dy][dx][i] = fmaf(w, hv[i], ah[dy][dx][i]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      if (oy0 + dy >= p.out_h) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        if (ox0 + dx >= p.out_w) continue;
        const int64_t o = obase + ((int64_t)(oy0 + dy) * p.out_w + ox0 + dx) * p.minor;
        VecIO<T, V>::store(out_x + o, ax[dy][dx]);
        VecIO<T, V>::store(out_h + o, ah[dy][dx]);
      }
    }
  }
}

template <typename T, int UP, int DOWN>
int launch_fir4_quad_gn(const void* in, void* out_h, void* out_x, const float* kern, const float* table, int table_ld, int act,
                        const FirP& p, cudaStream_t st) {
  constexpr int V = 4, YB = 4;
  const int cv = p.minor / V;
  const int bxs = (p.out_w + 1) / 2;
  const int bygs = ((p.out_h + 1) / 2 + YB - 1) / YB;
  const int64_t total = p.major * bygs * (int64_t)bxs * cv;
  const int64_t blocks = (total + 255) / 256;
  if (blocks >= (1LL << 31)) return MUDIFF_EUNSUPPORTED;
  fir4_quad_gn_kernel<T, UP, DOWN, YB><<<(unsigned)blocks, 256, 0, st>>>((const T*)in, (T*)out_h, (T*)out_x, kern, table, table_ld,
                                                                     act, p, bxs, bygs);
  return mudiff_launch_status();
}

// Shared-memory tiled variant of the fused AdaGN + SiLU + FIR kernel (bf16, channels % 32 == 0): the register-window kernel above
// evaluates the activation once per WINDOW position (2.25x per input element when downsampling, 9x when upsampling) and
// re-reads every input 2.25 / 9 times through L1, which left it MUFU / LSU bound at a third of the HBM roofline.  Here a CTA
// stages the input tile (+ FIR halo) of a 32-channel chunk ONCE: raw x and h = act(x * scale + shift) rounded to bf16 (exactly
// what the stand-alone GroupNorm pass stores), then every thread filters 2 (down) / 2 x 4 (up) outputs of one 16-byte channel
// vector from shared memory.  DOWN: 8 x 16 outputs from 18 x 34 inputs (76.5 KB, two CTAs per SM; pixel slots swizzled so
// that the stride-2 window reads are bank-conflict free); UP: 16 x 32 outputs from 10 x 18 inputs, three CTAs per SM (write-bound:
// 618 -> 520 us for 128 channels at 128^2 against two CTAs; filtering x and h in two passes over ONE tile buffer - four CTAs per
// SM - was slower, 579 us, and did not move the instruction-bound down-sampling case).
constexpr int kFirCC = 32;
template <int UP, int DOWN> struct FirTile {
  static constexpr int OTH = DOWN == 2 ? 8 : 16, OTW = DOWN == 2 ? 16 : 32;
  static constexpr int ITH = DOWN == 2 ? 18 : 10, ITW = DOWN == 2 ? 34 : 18;
  static constexpr int SMEM = ITH * ITW * (kFirCC * 2) * 2;
  static __device__ __forceinline__ int slot(int q) { return DOWN == 2 ? (q ^ ((q >> 1) & 1)) : q; }
};

__device__ __forceinline__ void unpack_bf16x8(const uint4& raw, f32x2 (&v)[4]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[i] = pack2(f.x, f.y); }
}
__device__ __forceinline__ uint4 pack_bf16x8(const f32x2 (&v)[4]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float a, b; unpack2(v[i], a, b); h[i] = __floats2bfloat162_rn(a, b); }
  return raw;
}

template <int UP, int DOWN>
__global__ void __launch_bounds__(256, UP == 2 ? 3 : 2) fir4_tile_gn_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out_h,
                                                           __nv_bfloat16* __restrict__ out_x, const float* __restrict__ kern,
                                                           const float* __restrict__ table, int table_ld, int act, FirP p,
                                                           int tiles_x, int tiles_y, int cchunks) {
  using G = Fir4Geom<UP, DOWN>;
  using TL = FirTile<UP, DOWN>;
  extern __shared__ uint4 fir_smem[];
  uint4* sx = fir_smem;                                   // [ITH * ITW pixel slots][4 vectors] raw x
  uint4* sh = fir_smem + TL::ITH * TL::ITW * 4;           // ... activated, bf16-rounded
  __shared__ float sk[16];
  if (threadIdx.x < 16) sk[threadIdx.x] = kern[15 - threadIdx.x];        // flipped: true convolution
  int bid = blockIdx.x;
  const int cc = bid % cchunks; bid /= cchunks;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int64_t m = bid / tiles_y;
  const int oy0 = ty * TL::OTH, ox0 = tx * TL::OTW;
  const int iy0 = UP == 2 ? (oy0 - p.py0) / 2 : oy0 * DOWN - p.py0;     // UP: oy0, py0 even -> exact
  const int ix0 = UP == 2 ? (ox0 - p.px0) / 2 : ox0 * DOWN - p.px0;
  const int v = threadIdx.x & 3;
  const int c0 = cc * kFirCC + v * 8;
  {
    // ---- stage: one 16-byte vector per thread per trip; (scale, shift) of this thread's 8 channels live in registers
    f32x2 sc[4], sf[4];
    const float4* tp = reinterpret_cast<const float4*>(table + ((int64_t)m * table_ld + c0) * 2);
    const float hf = act == MUDIFF_ACT_SILU ? 0.5f : 1.f;             // silu(t) = h + h tanh(h), h = t / 2
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t4 = __ldg(tp + i);
      sc[i] = pack2(t4.x * hf, t4.z * hf); sf[i] = pack2(t4.y * hf, t4.w * hf);
    }
    const __nv_bfloat16* inm = in + m * p.in_h * (int64_t)p.in_w * p.minor + c0;
    constexpr int ITEMS = TL::ITH * TL::ITW * 4;
    constexpr int TRIPS = (ITEMS + 255) / 256;
    uint4 raw[TRIPS];
    bool ok[TRIPS];
#pragma unroll
    for (int j = 0; j < TRIPS; ++j) {
      const int it = threadIdx.x + 256 * j;
      const int q = it >> 2;
      const int r = q / TL::ITW, c = q - r * TL::ITW;
      const int iy = iy0 + r, ix = ix0 + c;
      ok[j] = it < ITEMS && iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w;
      raw[j] = make_uint4(0u, 0u, 0u, 0u);
      if (ok[j]) raw[j] = __ldg(reinterpret_cast<const uint4*>(inm + ((int64_t)iy * p.in_w + ix) * p.minor));
    }
#pragma unroll
    for (int j = 0; j < TRIPS; ++j) {
      const int it = threadIdx.x + 256 * j;
      if (it < ITEMS) {
        uint4 hv = make_uint4(0u, 0u, 0u, 0u);
        if (ok[j]) {                                                  // zero padding applies AFTER the activation
          f32x2 x2[4];
          unpack_bf16x8(raw[j], x2);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            x2[i] = fma2(x2[i], sc[i], sf[i]);
            if (act == MUDIFF_ACT_SILU) {
              float a, b; unpack2(x2[i], a, b);
              x2[i] = fma2(x2[i], pack2(tanh_approx(a), tanh_approx(b)), x2[i]);
            }
          }
          hv = pack_bf16x8(x2);
        }
        const int s = TL::slot(it >> 2) * 4 + v;
        sx[s] = raw[j]; sh[s] = hv;
      }
    }
  }
  __syncthreads();
  float kf[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i] = sk[i];
  const int64_t obase = m * p.out_h * (int64_t)p.out_w * p.minor + c0;
  if (DOWN == 2) {
    // ---- one vertical pair of outputs per thread: window rows 4 * oyp .. + 5, columns 2 * ox .. + 3 of the staged tile
    const int ox = (threadIdx.x >> 2) & 15, oyp = threadIdx.x >> 6;
    f32x2 ax[2][4], ah[2][4];
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
      for (int i = 0; i < 4; ++i) { ax[d][i] = pack2(0.f, 0.f); ah[d][i] = pack2(0.f, 0.f); }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int s = TL::slot((4 * oyp + r) * TL::ITW + 2 * ox + c) * 4 + v;
        f32x2 xv[4], hv[4];
        unpack_bf16x8(sx[s], xv);
        unpack_bf16x8(sh[s], hv);
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int ky = r - 2 * d;
          if (ky < 0 || ky > 3) continue;
          const float w = kf[ky * 4 + c];
          const f32x2 ww = pack2(w, w);
#pragma unroll
          for (int i = 0; i < 4; ++i) { ax[d][i] = fma2(ww, xv[i], ax[d][i]); ah[d][i] = fma2(ww, hv[i], ah[d][i]); }
        }
      }
    }
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const int oy = oy0 + 2 * oyp + d, oxx = ox0 + ox;
      if (oy < p.out_h && oxx < p.out_w) {
        const int64_t o = obase + ((int64_t)oy * p.out_w + oxx) * p.minor;
        *reinterpret_cast<uint4*>(out_x + o) = pack_bf16x8(ax[d]);
        *reinterpret_cast<uint4*>(out_h + o) = pack_bf16x8(ah[d]);
      }
    }
  } else {
    // ---- 2 x 2 output quads from the 3 x 3 window at (qy, qx) of the staged tile; two quads per thread
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int qi = (threadIdx.x >> 2) + 64 * half;
      const int qx = qi & 15, qy = qi >> 4;
      f32x2 ax[2][2][4], ah[2][2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int i = 0; i < 4; ++i) { ax[a][b][i] = pack2(0.f, 0.f); ah[a][b][i] = pack2(0.f, 0.f); }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int s = ((qy + r) * TL::ITW + qx + c) * 4 + v;
          f32x2 xv[4], hv[4];
          unpack_bf16x8(sx[s], xv);
          unpack_bf16x8(sh[s], hv);
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            const int ky = G::tap(dy, r);
            if (ky < 0) continue;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const int kx = G::tap(dx, c);
              if (kx < 0) continue;
              const float w = kf[ky * 4 + kx];
              const f32x2 ww = pack2(w, w);
#pragma unroll
              for (int i = 0; i < 4; ++i) { ax[dy][dx][i] = fma2(ww, xv[i], ax[dy][dx][i]); ah[dy][dx][i] = fma2(ww, hv[i], ah[dy][dx][i]); }
            }
          }
        }
      }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int oy = oy0 + 2 * qy + dy;
        if (oy >= p.out_h) continue;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int oxx = ox0 + 2 * qx + dx;
          if (oxx >= p.out_w) continue;
          const int64_t o = obase + ((int64_t)oy * p.out_w + oxx) * p.minor;
          *reinterpret_cast<uint4*>(out_x + o) = pack_bf16x8(ax[dy][dx]);
          *reinterpret_cast<uint4*>(out_h + o) = pack_bf16x8(ah[dy][dx]);
        }
      }
    }
  }
}

template <int UP, int DOWN>
int launch_fir4_tile_gn(const void* in, void* out_h, void* out_x, const float* kern, const float* table, int table_ld, int act,
                        const FirP& p, cudaStream_t st) {
  using TL = FirTile<UP, DOWN>;
  const int tiles_x = (p.out_w + TL::OTW - 1) / TL::OTW, tiles_y = (p.out_h + TL::OTH - 1) / TL::OTH;
  const int cchunks = p.minor / kFirCC;
  const int64_t blocks = p.major * tiles_y * (int64_t)tiles_x * cchunks;
  if (blocks >= (1LL << 31)) return MUDIFF_EUNSUPPORTED;
  static bool attr_set[16] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(fir4_tile_gn_kernel<UP, DOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TL::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  fir4_tile_gn_kernel<UP, DOWN><<<(unsigned)blocks, 256, TL::SMEM, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out_h,
                                                                          (__nv_bfloat16*)out_x, kern, table, table_ld, act, p,
                                                                          tiles_x, tiles_y, cchunks);
  return mudiff_launch_status();
}

template <typename T>
int launch_fir(const void* in, void* out, const float* kern, const FirP& p, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  const int64_t total = p.major * p.out_h * p.out_w * p.minor;
  if (total == 0) return 0;
  const bool aligned = ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
  if (p.minor % V == 0 && aligned && p.kh == 4 && p.kw == 4 && p.up_x == p.up_y && p.down_x == p.down_y &&
      p.up_x <= 2 && p.down_x <= 2 && !(p.up_x == 2 && p.down_x == 2) &&
      !(p.up_x == 2 && ((p.px0 | p.py0) & 1)) && p.px0 >= 0 && p.py0 >= 0) {
    if (p.up_x == 2) return launch_fir4_quad<T, 2, 1>(in, out, kern, p, st);
    if (p.down_x == 2) return launch_fir4_quad<T, 1, 2>(in, out, kern, p, st);
    return launch_fir4_quad<T, 1, 1>(in, out, kern, p, st);
  }
  if (p.minor % V == 0 && aligned) {
    int grid = grid_for(total / V, 256);
    fir_nhwc_vec_kernel<T><<<grid, 256, 0, st>>>((const T*)in, (T*)out, kern, p);
    return mudiff_launch_status();
  }
  if (p.minor == 1 && p.major <= 65535LL * 32768) {
    // staged footprint of a 32x32 output tile
    int tin_h = ((FIR_TILE - 1) * p.down_y + p.kh - 1) / p.up_y + 2;
    int tin_w = ((FIR_TILE - 1) * p.down_x + p.kw - 1) / p.up_x + 2;
    size_t smem = sizeof(float) * ((size_t)p.kh * p.kw + (size_t)tin_h * tin_w);
    if (smem <= 48 * 1024 && p.major <= 65535) {
      dim3 grid((p.out_w + FIR_TILE - 1) / FIR_TILE, (p.out_h + FIR_TILE - 1) / FIR_TILE, (unsigned)p.major);
      fir_nchw_tiled_kernel<T><<<grid, 256, smem, st>>>((const T*)in, (T*)out, kern, p, tin_h, tin_w);
      return mudiff_launch_status();
    }
  }
  int grid = grid_for(total, 256);
  fir_generic_kernel<T><<<grid, 256, 0, st>>>((const T*)in, (T*)out, kern, p);
  return mudiff_launch_status();
}

}  // namespace

extern "C" int mudiff_upfirdn2d(const void* in, void* out, const float* kernel, int dtype,
                                int64_t major, int in_h, int in_w, int minor, int kh, int kw,
                                int up_x, int up_y, int down_x, int down_y,
                                int pad_x0, int pad_x1, int pad_y0, int pad_y1, void* stream) {
  if (up_x < 1 || up_y < 1 || down_x < 1 || down_y < 1 || kh < 1 || kw < 1 || kh * kw > FIR_MAX_TAPS)
    return MUDIFF_EINVAL;
  if (major < 0 || in_h < 0 || in_w < 0 || minor < 1) return MUDIFF_EINVAL;
  FirP p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.px0 = pad_x0; p.py0 = pad_y0;
  int full_h = in_h * up_y + pad_y0 + pad_y1 - kh;
  int full_w = in_w * up_x + pad_x0 + pad_x1 - kw;
  if (full_h < 0 || full_w < 0) return MUDIFF_EINVAL;
  p.out_h = full_h / down_y + 1;
  p.out_w = full_w / down_x + 1;
  if (major == 0 || in_h == 0 || in_w == 0) return 0;
  if (!in || !out || !kernel) return MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MUDIFF_F32: return launch_fir<float>(in, out, kernel, p, st);
    case MUDIFF_BF16: return launch_fir<__nv_bfloat16>(in, out, kernel, p, st);
    case MUDIFF_F16: return launch_fir<__half>(in, out, kernel, p, st);
    default: return MUDIFF_EINVAL;
  }
}

// out_h = FIR(act(x * scale + shift)), out_x = FIR(x) from ONE read of x: the two resampled branches of a resample ResBlock
// (backbones/layerspp.py:293-305).  x, out_h, out_x: channels-last [batch, H, W, C] dense; kernel: 4x4 fp32;
// table: float[batch][table_ld][2] = (scale, shift) from mudiff_gn_scale_shift; act: MUDIFF_ACT_NONE | MUDIFF_ACT_SILU.
// Only the generators' two modes: up = 2 with pad (2, 1) or down = 2 with pad (1, 1).
extern "C" int mudiff_upfirdn2d_gn(const void* x, void* out_h, void* out_x, const float* kernel, const float* table,
                                   int table_ld, int act, int dtype, int batch, int in_h, int in_w, int channels,
                                   int up, int down, int pad0, int pad1, void* stream) {
  if (!x || !out_h || !out_x || !kernel || !table || batch <= 0 || in_h <= 0 || in_w <= 0 || channels <= 0) return MUDIFF_EINVAL;
  if (act != MUDIFF_ACT_NONE && act != MUDIFF_ACT_SILU) return MUDIFF_EINVAL;
  if (!((up == 2 && down == 1 && pad0 == 2 && pad1 == 1) || (up == 1 && down == 2 && pad0 == 1 && pad1 == 1))) return MUDIFF_EUNSUPPORTED;
  if (channels % 4 || table_ld < channels) return MUDIFF_EUNSUPPORTED;
  if (((uintptr_t)x % 16) || ((uintptr_t)out_h % 16) || ((uintptr_t)out_x % 16)) return MUDIFF_EUNSUPPORTED;
  FirP p;
  p.major = batch; p.in_h = in_h; p.in_w = in_w; p.minor = channels; p.kh = 4; p.kw = 4;
  p.up_x = p.up_y = up; p.down_x = p.down_y = down; p.px0 = p.py0 = pad0;
  p.out_h = (in_h * up + pad0 + pad1 - 4) / down + 1;
  p.out_w = (in_w * up + pad0 + pad1 - 4) / down + 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MUDIFF_BF16) {
    static int tiled = -1;                     // MUDIFF_FIR_GN_TILED=0: the register-window kernel for every shape
    if (tiled < 0) { const char* e = getenv("MUDIFF_FIR_GN_TILED"); tiled = (e && e[0] == '0') ? 0 : 1; }
    if (tiled && channels % kFirCC == 0)
      return up == 2 ? launch_fir4_tile_gn<2, 1>(x, out_h, out_x, kernel, table, table_ld, act, p, st)
                     : launch_fir4_tile_gn<1, 2>(x, out_h, out_x, kernel, table, table_ld, act, p, st);
    return up == 2 ? launch_fir4_quad_gn<__nv_bfloat16, 2, 1>(x, out_h, out_x, kernel, table, table_ld, act, p, st)
                   : launch_fir4_quad_gn<__nv_bfloat16, 1, 2>(x, out_h, out_x, kernel, table, table_ld, act, p, st);
  }
  if (dtype == MUDIFF_F32) {
    return up == 2 ? launch_fir4_quad_gn<float, 2, 1>(x, out_h, out_x, kernel, table, table_ld, act, p, st)
                   : launch_fir4_quad_gn<float, 1, 2>(x, out_h, out_x, kernel, table, table_ld, act, p, st);
  }
  return MUDIFF_EUNSUPPORTED;
}
