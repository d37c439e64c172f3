// CUDA-core implicit-GEMM convolution / contraction on NHWC tensors (fp32 math).
// This is (1) the fp32 parity path (max-abs <= 1e-4 vs the fp32 reference needs true
// fp32 products, which tcgen05 does not offer), and (2) the path for shapes the
// tensor-core kernel does not take: Cin=1 stem convs, the Cout=1 head conv, the
// stride-2 input-pyramid conv.  Same mudiff_conv_desc contract as mudiff_conv_tc.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

struct SimtP {
  const void* a[3];
  int a_c[3], a_ld[3], a_taps[3], a_koff[3];
  int nseg, a_batched;
  int batch, h, w, ho, wo, stride, pad;
  const void* wt; int64_t w_bstride; int ktot; int w_ld;
  int n;
  const float* bias; const float* rowbias; int rowbias_ld;
  const void* residual; int res_ld;
  float alpha, beta; int act;
  void* out; int out_ld, out_coff;
  int rpb;                                   // stem kernels: image rows per block (launch shape only, see stem_rpb)
};

template <typename TO>
__device__ __forceinline__ void epilogue_store(const SimtP& p, int b, int64_t pix, int n, float acc) {
  float v = acc;
  if (p.bias) v += p.bias[n];
  if (p.rowbias) v += p.rowbias[(int64_t)b * p.rowbias_ld + n];
  v *= p.alpha;
  if (p.residual) v = fmaf(p.beta, Cvt<TO>::to_f(((const TO*)p.residual)[pix * p.res_ld + n]), v);
  v = apply_act(v, p.act);
  ((TO*)p.out)[pix * p.out_ld + p.out_coff + n] = Cvt<TO>::from_f(v);
}

// ---------------------------------------------------------------------------------
// generic tile kernel: 64 pixels x 64 channels per block, 256 threads, 4x4 per thread
// ---------------------------------------------------------------------------------
#define BM 64
#define BN 64
#define BK 16
template <typename TA, typename TO>
__global__ void __launch_bounds__(256) conv_simt_kernel(SimtP p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  __shared__ int s_oy[BM], s_ox[BM];
  const int b = blockIdx.z;
  const int hw_o = p.ho * p.wo;
  const int pix0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  if (tid < BM) {
    int pp = pix0 + tid;
    s_oy[tid] = pp < hw_o ? pp / p.wo : -100000;
    s_ox[tid] = pp < hw_o ? pp % p.wo : 0;
  }
  __syncthreads();
  const TA* wbase = (const TA*)p.wt + (int64_t)b * p.w_bstride;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tx = tid % 16, ty = tid / 16;      // tx -> 4 channels, ty -> 4 pixels
  const int lp = tid / 4, lk = (tid % 4) * 4;  // loader: pixel / weight-row lp, 4 consecutive k
  for (int s = 0; s < p.nseg; ++s) {
    const TA* abase = (const TA*)p.a[s] + (p.a_batched ? (int64_t)b * p.h * p.w * p.a_ld[s] : 0);
    const int C = p.a_c[s];
    for (int tap = 0; tap < p.a_taps[s]; ++tap) {
      const int dy = p.a_taps[s] == 9 ? tap / 3 - p.pad : 0;
      const int dx = p.a_taps[s] == 9 ? tap % 3 - p.pad : 0;
      const int iy = s_oy[lp] * p.stride + dy, ix = s_ox[lp] * p.stride + dx;
      const bool inb = iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
      const TA* arow = abase + ((int64_t)iy * p.w + ix) * p.a_ld[s];
      for (int c0 = 0; c0 < C; c0 += BK) {
        // A tile: 64 pixels x 16 channels
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int c = c0 + lk + i;
          As[lk + i][lp] = (inb && c < C) ? Cvt<TA>::to_f(arow[c]) : 0.f;
        }
        // W tile: 64 out-channels x 16 k
        {
          int n = n0 + lp;
          const TA* wrow = wbase + (int64_t)n * p.w_ld + p.a_koff[s] + tap * C + c0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int c = c0 + lk + i;
            Ws[lk + i][lp] = (n < p.n && c < C) ? Cvt<TA>::to_f(wrow[lk + i]) : 0.f;
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
          float a4[4] = {av.x, av.y, av.z, av.w};
          float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], w4[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int pp = pix0 + ty * 4 + i;
    if (pp >= hw_o) continue;
    int64_t pix = (int64_t)b * hw_o + pp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < p.n) epilogue_store<TO>(p, b, pix, n, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------
// small-N kernel (N <= 4, e.g. the 64->1 head conv): 8 lanes per output pixel, each
// lane owns one 8-channel slice of every 64-channel block, shuffle-reduce at the end.
// Weights staged in shared memory as fp32.
// ---------------------------------------------------------------------------------
template <typename TA, typename TO>
__global__ void __launch_bounds__(256) conv_small_n_kernel(SimtP p) {
  extern __shared__ float sw[];                 // [n][ktot]
  for (int i = threadIdx.x; i < p.n * p.ktot; i += blockDim.x) sw[i] = Cvt<TA>::to_f(((const TA*)p.wt)[i]);
  __syncthreads();
  const int hw_o = p.ho * p.wo;
  const int64_t total = (int64_t)p.batch * hw_o;
  const int sub = threadIdx.x & 7;
  for (int64_t pix = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; pix < total;
       pix += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int b = (int)(pix / hw_o);
    const int pp = (int)(pix % hw_o);
    const int oy = pp / p.wo, ox = pp % p.wo;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < p.nseg; ++s) {
      const TA* abase = (const TA*)p.a[s] + (int64_t)b * p.h * p.w * p.a_ld[s];
      const int C = p.a_c[s];
      for (int tap = 0; tap < p.a_taps[s]; ++tap) {
        const int dy = p.a_taps[s] == 9 ? tap / 3 - p.pad : 0;
        const int dx = p.a_taps[s] == 9 ? tap % 3 - p.pad : 0;
        const int iy = oy * p.stride + dy, ix = ox * p.stride + dx;
        if (iy < 0 || iy >= p.h || ix < 0 || ix >= p.w) continue;
        const TA* arow = abase + ((int64_t)iy * p.w + ix) * p.a_ld[s];
        const int kb = p.a_koff[s] + tap * C;
        for (int c = sub; c < C; c += 8) {
          float av = Cvt<TA>::to_f(arow[c]);
#pragma unroll
          for (int n = 0; n < 4; ++n)
            if (n < p.n) acc[n] = fmaf(av, sw[n * p.ktot + kb + c], acc[n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 1);
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 2);
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 4);
    }
    if (sub == 0)
      for (int n = 0; n < p.n; ++n) epilogue_store<TO>(p, b, pix, n, acc[n]);
  }
}

// ---------------------------------------------------------------------------------
// small-K kernel (Ktot <= 36, e.g. the 1->64 stem convs): one thread per (pixel, 8 output
// channels); input taps gathered once per pixel, weights in shared memory.
// ---------------------------------------------------------------------------------
template <typename TA, typename TO>
__global__ void __launch_bounds__(256) conv_small_k_kernel(SimtP p) {
  extern __shared__ float sw[];                 // [ktot][n]  (n innermost)
  for (int i = threadIdx.x; i < p.n * p.ktot; i += blockDim.x) {
    int n = i / p.ktot, k = i % p.ktot;
    sw[k * p.n + n] = Cvt<TA>::to_f(((const TA*)p.wt)[i]);
  }
  __syncthreads();
  const int hw_o = p.ho * p.wo;
  const int nv = p.n / 8;
  const int64_t total = (int64_t)p.batch * hw_o * nv;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = idx / nv;
    const int n0 = (int)(idx % nv) * 8;
    const int b = (int)(pix / hw_o);
    const int pp = (int)(pix % hw_o);
    const int oy = pp / p.wo, ox = pp % p.wo;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int s = 0; s < p.nseg; ++s) {
      const TA* abase = (const TA*)p.a[s] + (int64_t)b * p.h * p.w * p.a_ld[s];
      const int C = p.a_c[s];
      for (int tap = 0; tap < p.a_taps[s]; ++tap) {
        const int dy = p.a_taps[s] == 9 ? tap / 3 - p.pad : 0;
        const int dx = p.a_taps[s] == 9 ? tap % 3 - p.pad : 0;
        const int iy = oy * p.stride + dy, ix = ox * p.stride + dx;
        if (iy < 0 || iy >= p.h || ix < 0 || ix >= p.w) continue;
        const TA* arow = abase + ((int64_t)iy * p.w + ix) * p.a_ld[s];
        for (int c = 0; c < C; ++c) {
          float av = Cvt<TA>::to_f(arow[c]);
          const float* wr = sw + (p.a_koff[s] + tap * C + c) * p.n + n0;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(av, wr[j], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) epilogue_store<TO>(p, b, pix, n0 + j, acc[j]);
  }
}


// ---------------------------------------------------------------------------------
#define STEM_ROWS 8
// rows per block of the stem kernels: STEM_ROWS for large launches; fewer when the launch would not fill the SMs (a batch-1
// 256^2 image is 32 blocks of 8 rows - 28 us for a kernel whose work is 2 - 3 us; with one row per block it is 256 blocks).
// Every output pixel is computed by the same instruction sequence whatever the block shape: results are identical.
static inline int stem_rpb(int64_t rows) {
  int64_t r = rows / (2 * MUDIFF_NUM_SMS);
  return (int)(r < 1 ? 1 : (r > STEM_ROWS ? STEM_ROWS : r));
}
// stem kernel: Cin == 1, 3x3 pad 1, stride 1, N % 8 == 0 (the 1->nf ConvFeatBlock convs).
// HBM-write bound in principle (9 MACs per output) but FFMA/issue bound in practice, so every thread
// computes TWO adjacent pixels x 8 output channels from one 3x4 input patch (12 loads for 144 FMAs; the 72
// weights + 8 biases live in registers), STEM_ROWS rows per block; a warp stores 8 pixels x 128 B contiguous.
template <typename TO>
__global__ void __launch_bounds__(256) conv_stem_kernel(SimtP p) {
  const int nv = p.n / 8;
  const int n0 = (threadIdx.x % nv) * 8;
  const int lane = threadIdx.x / nv, lanes = blockDim.x / nv;
  if (lane >= lanes) return;
  float w[9][8], bs[8];
  const float* wt = (const float*)p.wt;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int t = 0; t < 9; ++t) w[t][j] = wt[(n0 + j) * 9 + t];
    bs[j] = p.bias ? p.bias[n0 + j] : 0.f;
  }
  const int ld = p.a_ld[0];
  const int W = p.w, H = p.h;
  const int64_t rows = (int64_t)p.batch * H;
  for (int64_t row = (int64_t)blockIdx.x * p.rpb; row < rows && row < (int64_t)(blockIdx.x + 1) * p.rpb; ++row) {
    const int b = (int)(row / H), y = (int)(row - (int64_t)b * H);
    float rb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) rb[j] = bs[j] + (p.rowbias ? p.rowbias[(int64_t)b * p.rowbias_ld + n0 + j] : 0.f);
    const float* in = (const float*)p.a[0] + (int64_t)b * H * W * ld;
    const float* r0 = y > 0 ? in + (int64_t)(y - 1) * W * ld : nullptr;
    const float* r1 = in + (int64_t)y * W * ld;
    const float* r2 = y + 1 < H ? in + (int64_t)(y + 1) * W * ld : nullptr;
    for (int x = lane * 2; x < W; x += lanes * 2) {
      float v[3][4];                               // rows y-1..y+1, columns x-1..x+2
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ix = x - 1 + c;
        const bool okx = ix >= 0 && ix < W;
        v[0][c] = (okx && r0) ? __ldg(r0 + (int64_t)ix * ld) : 0.f;
        v[1][c] = okx ? __ldg(r1 + (int64_t)ix * ld) : 0.f;
        v[2][c] = (okx && r2) ? __ldg(r2 + (int64_t)ix * ld) : 0.f;
      }
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        if (x + px >= W) break;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = rb[j];
#pragma unroll
          for (int t = 0; t < 9; ++t) a = fmaf(v[t / 3][t % 3 + px], w[t][j], a);
          acc[j] = a * p.alpha;
        }
        if (p.act != MUDIFF_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = apply_act(acc[j], p.act);
        }
        const int64_t pix = row * W + x + px;
        TO* op = (TO*)p.out + pix * p.out_ld + p.out_coff + n0;
        if constexpr (sizeof(TO) == 2) {
          store_vec<TO>(op, acc);
        } else {
          float lo[4] = {acc[0], acc[1], acc[2], acc[3]}, hi[4] = {acc[4], acc[5], acc[6], acc[7]};
          store_vec<float>((float*)op, lo);
          store_vec<float>((float*)op + 4, hi);
        }
      }
    }
  }
}

// stride-2 VALID variant of the stem kernel (Cin == 1, pad 0): the 3x3 stride-2 conv of the input-pyramid branch
// (conv_downsample_2d, up_or_down_sampling.py:183) after its FIR pre-filter.  One output pixel x 8 channels per
// thread, all nine taps in bounds by construction, channel pairs on FFMA2.
template <typename TO>
__global__ void __launch_bounds__(256) conv_stem_s2_kernel(SimtP p) {
  const int nv = p.n / 8;
  const int n0 = (threadIdx.x % nv) * 8;
  const int lane = threadIdx.x / nv, lanes = blockDim.x / nv;
  if (lane >= lanes) return;
  const float* wt = (const float*)p.wt;
  f32x2 w2[9][4], bs2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = n0 + 2 * j;
#pragma unroll
    for (int t = 0; t < 9; ++t) w2[t][j] = pack2(wt[c * 9 + t], wt[(c + 1) * 9 + t]);
    bs2[j] = pack2(p.bias ? p.bias[c] : 0.f, p.bias ? p.bias[c + 1] : 0.f);
  }
  const int ld = p.a_ld[0];
  const int W = p.w, H = p.h, WO = p.wo, HO = p.ho;
  const int64_t rows = (int64_t)p.batch * HO;
  for (int64_t row = (int64_t)blockIdx.x * p.rpb; row < rows && row < (int64_t)(blockIdx.x + 1) * p.rpb; ++row) {
    const int b = (int)(row / HO), y = (int)(row - (int64_t)b * HO);
    const float* in = (const float*)p.a[0] + ((int64_t)b * H + 2 * y) * W * ld;
    for (int x = lane; x < WO; x += lanes) {
      float v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = __ldg(in + ((int64_t)(t / 3) * W + 2 * x + t % 3) * ld);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f32x2 a = bs2[j];
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fma2(pack2(v[t], v[t]), w2[t][j], a);
        unpack2(a, acc[2 * j], acc[2 * j + 1]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float r = acc[j] + (p.rowbias ? p.rowbias[(int64_t)b * p.rowbias_ld + n0 + j] : 0.f);
        acc[j] = apply_act(r * p.alpha, p.act);
      }
      const int64_t pix = row * WO + x;
      TO* op = (TO*)p.out + pix * p.out_ld + p.out_coff + n0;
      if constexpr (sizeof(TO) == 2) {
        store_vec<TO>(op, acc);
      } else {
        float lo[4] = {acc[0], acc[1], acc[2], acc[3]}, hi[4] = {acc[4], acc[5], acc[6], acc[7]};
        store_vec<float>((float*)op, lo);
        store_vec<float>((float*)op + 4, hi);
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Fused stem: conv3x3(1 -> N) -> GroupNorm / AdaGN -> SiLU without ever storing the raw conv output
// (ConvFeatBlock / ConvBlock / ConvBlock_GAP, backbones/layerspp.py:394-501).  The GroupNorm statistics of a
// Cin == 1 convolution follow from second moments of the INPUT: with the 9-vector patch x_t(p) (zero padded),
//   sum_p y_c   = n b_c + sum_t w_ct s_t,                       s_t  = sum_p x_t(p)
//   sum_p y_c^2 = n b_c^2 + 2 b_c sum_t w_ct s_t + w_c^T R w_c,  R_tu = sum_p x_t(p) x_u(p)
// so one pass over the 1-channel image (4 B / pixel) replaces the statistics pass over the N-channel tensor
// (2 N B / pixel), and the conv kernel applies scale/shift/activation in its epilogue: 1 write instead of
// write + read + read + write.
// ---------------------------------------------------------------------------------
#define MOM_N 54                       // 9 sums + 45 upper-triangular products
#define MOM_ROWS 16
__global__ void __launch_bounds__(256) stem_moments_kernel(const float* __restrict__ x, int ld, int H, int W,
                                                           double* __restrict__ moments, double* __restrict__ partial,
                                                           unsigned int* __restrict__ tickets) {
  const int b = blockIdx.y, chunks = gridDim.x;
  const int y0 = blockIdx.x * MOM_ROWS, y1 = min(y0 + MOM_ROWS, H);
  const float* img = x + (int64_t)b * H * W * ld;
  float acc[MOM_N];
#pragma unroll
  for (int i = 0; i < MOM_N; ++i) acc[i] = 0.f;
  for (int xx = threadIdx.x; xx < W; xx += blockDim.x) {
    float v[3][3];                              // rolling window: rows y-1, y, y+1 at columns xx-1..xx+1
    auto load_row = [&](int yy, float (&r)[3]) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int ix = xx - 1 + c;
        r[c] = (yy >= 0 && yy < H && ix >= 0 && ix < W) ? __ldg(img + ((int64_t)yy * W + ix) * ld) : 0.f;
      }
    };
    load_row(y0 - 1, v[0]);
    load_row(y0, v[1]);
    for (int y = y0; y < y1; ++y) {
      load_row(y + 1, v[2]);
      float pt[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) pt[t] = v[t / 3][t % 3];
      int k = 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        acc[t] += pt[t];
#pragma unroll
        for (int u2 = t; u2 < 9; ++u2) { acc[k] = fmaf(pt[t], pt[u2], acc[k]); ++k; }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { v[0][c] = v[1][c]; v[1][c] = v[2][c]; }
    }
  }
  // warp reduce (fp32, fixed order), then across the warps in double
  __shared__ double sred[8][MOM_N];
  __shared__ bool s_last;
#pragma unroll
  for (int i = 0; i < MOM_N; ++i) {
    float a = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5][i] = (double)a;
  }
  __syncthreads();
  double* mine = partial + ((int64_t)b * chunks + blockIdx.x) * MOM_N;
  if (threadIdx.x < MOM_N) {
    double a = 0.0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) a += sred[w2][threadIdx.x];
    mine[threadIdx.x] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[b], 1u);
    s_last = (t == (unsigned int)chunks - 1);
    if (s_last) tickets[b] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < MOM_N) {
    double a = 0.0;
    for (int c = 0; c < chunks; ++c) a += partial[((int64_t)b * chunks + c) * MOM_N + threadIdx.x];
    moments[(int64_t)b * MOM_N + threadIdx.x] = a;
  }
}

// per (image, channel) folded GroupNorm scale / shift of the conv output, from the input moments:
// one block per image, one thread per channel.
__global__ void stem_scale_shift_kernel(const double* __restrict__ moments, const float* __restrict__ wt,
                                        const float* __restrict__ bias, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, int64_t gb_bstride, int n, int groups, float eps,
                                        double cnt, float* __restrict__ scale_shift) {
  __shared__ double s_cs[256][2];
  __shared__ float s_mean[64], s_rstd[64];
  const int b = blockIdx.x, c = threadIdx.x;
  const int cpg = n / groups;
  if (c < n) {
    const double* m = moments + (int64_t)b * MOM_N;
    double wc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wc[t] = (double)wt[c * 9 + t];
    const double bc = bias ? (double)bias[c] : 0.0;
    double lin = 0.0, quad = 0.0;
    int k = 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      lin += wc[t] * m[t];
#pragma unroll
      for (int u2 = t; u2 < 9; ++u2) { quad += (u2 == t ? 1.0 : 2.0) * wc[t] * wc[u2] * m[k]; ++k; }
    }
    s_cs[c][0] = cnt * bc + lin;
    s_cs[c][1] = cnt * bc * bc + 2.0 * bc * lin + quad;
  }
  __syncthreads();
  if (c < groups) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < cpg; ++i) { a += s_cs[c * cpg + i][0]; q += s_cs[c * cpg + i][1]; }
    const double nn = cnt * (double)cpg;
    const double mu = a / nn;
    double var = q / nn - mu * mu;
    if (var < 0.0) var = 0.0;
    s_mean[c] = (float)mu;
    s_rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (c < n) {
    const int g = c / cpg;
    const float ga = gamma ? gamma[(int64_t)b * gb_bstride + c] : 1.f;
    const float be = beta ? beta[(int64_t)b * gb_bstride + c] : 0.f;
    const float sc = ga * s_rstd[g];
    scale_shift[((int64_t)b * n + c) * 2 + 0] = sc;
    scale_shift[((int64_t)b * n + c) * 2 + 1] = be - s_mean[g] * sc;
  }
}

// conv3x3(1 -> N) * scale[b][c] + shift[b][c] -> activation, per-image folded weights (a thread owns 8 output channels: 72
// weights + 8 biases as FFMA2 pairs in registers) and walks pixel pairs of a row.  The input rows live in a shared-memory RING
// of four zero-padded rows (one new row per output row, one __syncthreads per row): the first version gathered its 3 x 4 window
// per pixel pair from global memory with bounds checks and 64-bit address arithmetic - 288 instructions per pair of which only
// 126 did arithmetic, loads or stores (cuobjdump); here the window is six 8-byte LDS with immediate offsets.  For SiLU outputs
// the 1/2 of silu(t) = h + h tanh(h), h = t / 2, is folded into the weights (bf16 path).
template <typename TO>
__global__ void __launch_bounds__(256, 2) conv_stem_gn_kernel(SimtP p, const float* __restrict__ scale_shift) {
  extern __shared__ float s_rows[];                       // [4][wp]: row slot (y & 3); s[1 + x] = x-th pixel, s[0] = s[W + 1] = 0
  const int nv = p.n / 8;
  const int n0 = (threadIdx.x % nv) * 8;
  const int lane = threadIdx.x / nv, lanes = blockDim.x / nv;
  const bool worker = lane < lanes;
  const float* wt = (const float*)p.wt;
  const int ld = p.a_ld[0];
  const int W = p.w, H = p.h;
  const int wp = (W + 2 + 3) & ~3;
  const int64_t rows = (int64_t)p.batch * H;
  const bool fold = p.act == MUDIFF_ACT_SILU && sizeof(TO) == 2;
  const float hf = fold ? 0.5f : 1.f;
  f32x2 w2[9][4], bs2[4];                                 // channel pairs (n0 + 2j, n0 + 2j + 1): FFMA2
  int cur_b = -1;
  const int64_t row0 = (int64_t)blockIdx.x * p.rpb;
  for (int64_t row = row0; row < rows && row < row0 + p.rpb; ++row) {
    const int b = (int)(row / H), y = (int)(row - (int64_t)b * H);
    const bool fresh = b != cur_b;                        // first row of this block in image b: stage rows y - 1, y, y + 1
    if (fresh) {
      if (cur_b >= 0) __syncthreads();                    // image boundary inside the block: all three slots are rewritten
      cur_b = b;
      if (worker) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = n0 + 2 * j;
          const float4 ss = *reinterpret_cast<const float4*>(scale_shift + ((int64_t)b * p.n + c) * 2);   // sc0 sh0 sc1 sh1
#pragma unroll
          for (int t = 0; t < 9; ++t) w2[t][j] = pack2(wt[c * 9 + t] * ss.x * hf, wt[(c + 1) * 9 + t] * ss.z * hf);
          bs2[j] = pack2(((p.bias ? p.bias[c] : 0.f) * ss.x + ss.y) * hf, ((p.bias ? p.bias[c + 1] : 0.f) * ss.z + ss.w) * hf);
        }
      }
    }
    const float* in = (const float*)p.a[0] + (int64_t)b * H * W * ld;
    for (int yy = fresh ? y - 1 : y + 1; yy <= y + 1; ++yy) {
      float* dst = s_rows + ((yy + 4) & 3) * wp;
      const bool inside = yy >= 0 && yy < H;
      const float* src = in + (int64_t)yy * W * ld;
      for (int i = threadIdx.x; i < wp; i += blockDim.x)
        dst[i] = (inside && i >= 1 && i <= W) ? __ldg(src + (int64_t)(i - 1) * ld) : 0.f;
    }
    __syncthreads();          // the ring has four slots: the slot written for row y + 2 is not among the three read for row y + 1
    if (!worker) continue;
    const float* r0 = s_rows + ((y + 3) & 3) * wp;        // row y - 1
    const float* r1 = s_rows + (y & 3) * wp;
    const float* r2 = s_rows + ((y + 1) & 3) * wp;
    for (int x = lane * 2; x < W; x += lanes * 2) {
      // window columns x - 1 .. x + 2 = s[x .. x + 3] (x even: two aligned float2 per row)
      float v[3][4];
      {
        const float2 a0 = *reinterpret_cast<const float2*>(r0 + x), a1 = *reinterpret_cast<const float2*>(r0 + x + 2);
        const float2 b0 = *reinterpret_cast<const float2*>(r1 + x), b1 = *reinterpret_cast<const float2*>(r1 + x + 2);
        const float2 c0 = *reinterpret_cast<const float2*>(r2 + x), c1 = *reinterpret_cast<const float2*>(r2 + x + 2);
        v[0][0] = a0.x; v[0][1] = a0.y; v[0][2] = a1.x; v[0][3] = a1.y;
        v[1][0] = b0.x; v[1][1] = b0.y; v[1][2] = b1.x; v[1][3] = b1.y;
        v[2][0] = c0.x; v[2][1] = c0.y; v[2][2] = c1.x; v[2][3] = c1.y;
      }
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        if (x + px >= W) break;
        f32x2 a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = bs2[j];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float xv = v[t / 3][t % 3 + px];
          const f32x2 xx = pack2(xv, xv);
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = fma2(xx, w2[t][j], a[j]);
        }
        float acc[8];
        if (fold) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float h0, h1; unpack2(a[j], h0, h1);
            unpack2(fma2(a[j], pack2(tanh_approx(h0), tanh_approx(h1)), a[j]), acc[2 * j], acc[2 * j + 1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) unpack2(a[j], acc[2 * j], acc[2 * j + 1]);
          if (p.act == MUDIFF_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = silu_exact(acc[j]);
          }
        }
        const int64_t pix = row * W + x + px;
        TO* op = (TO*)p.out + pix * p.out_ld + p.out_coff + n0;
        if constexpr (sizeof(TO) == 2) {
          store_vec<TO>(op, acc);
        } else {
          float lo[4] = {acc[0], acc[1], acc[2], acc[3]}, hi[4] = {acc[4], acc[5], acc[6], acc[7]};
          store_vec<float>((float*)op, lo);
          store_vec<float>((float*)op + 4, hi);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// head kernel: N == 1, 3x3 pad 1, stride 1, single segment, Cin % 64 == 0 (the nf->1 output conv, with the
// tanh of ncsnpp_generator_adagn_feat.py:445 in the epilogue).  HBM-read bound: every input pixel is read from
// global memory ONCE per strip.  A warp owns 30 output columns x HEAD_YR rows: lane l holds input column
// x0 - 1 + l, walks down the rows, and turns each loaded pixel into the nine per-tap dot products over its
// channels; vertical taps roll through registers, horizontal taps are exchanged with the neighbour lanes by
// shuffle (lanes 0 and 31 are halo columns).  Rows are fetched with fully coalesced 16-byte loads (one row
// ahead, in registers) and re-distributed pixel-per-lane through a padded per-warp shared-memory tile.
// ---------------------------------------------------------------------------------
#define HEAD_YR 32
#define HEAD_COLS 30
template <typename TA, typename TO>
__global__ void __launch_bounds__(256, 2) conv_head_kernel(SimtP p, int xchunks, int ystrips, int yr) {
  constexpr int V = 16 / sizeof(TA);          // 8 (bf16) or 4 (fp32) channels per 16-byte load
  constexpr int PB = 64 * sizeof(TA);         // bytes of one 64-channel block of a pixel: 128 / 256
  constexpr int NV = PB / 16;                 // 16-byte vectors per pixel block: 8 / 16
  constexpr int PITCH = PB + 16;              // padded pixel pitch -> conflict-free 16-byte accesses
  extern __shared__ __align__(16) uint8_t head_smem[];
  const int C = p.a_c[0];
  float* sw = reinterpret_cast<float*>(head_smem);                    // [9][C] fp32 weights
  uint8_t* tile = head_smem + (size_t)9 * C * 4 + (threadIdx.x >> 5) * (32 * PITCH);
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) sw[i] = Cvt<TA>::to_f(((const TA*)p.wt)[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t unit = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t units = (int64_t)p.batch * ystrips * xchunks;
  if (unit >= units) return;
  const int xc = (int)(unit % xchunks);
  const int ys = (int)((unit / xchunks) % ystrips);
  const int b = (int)(unit / ((int64_t)xchunks * ystrips));
  const int x0 = xc * HEAD_COLS;                 // first output column of this warp
  const int y0 = ys * yr;
  const int y1 = min(y0 + yr, p.h);
  const int ld = p.a_ld[0];
  const TA* in = (const TA*)p.a[0] + (int64_t)b * p.h * p.w * ld;
  const int nblk = C / 64;
  const float bias = (p.bias ? p.bias[0] : 0.f) + (p.rowbias ? p.rowbias[(int64_t)b * p.rowbias_ld] : 0.f);
  float r0[3] = {0.f, 0.f, 0.f}, r1[3] = {0.f, 0.f, 0.f};
  // coalesced fetch of one 64-channel block of row yy: vector j of the 32-pixel segment, j = lane + 32 i
  auto fetch = [&](int yy, int cb, uint4 (&regs)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = lane + 32 * i;
      const int px = j / NV, vv = j % NV;
      const int x = x0 - 1 + px;
      regs[i] = make_uint4(0u, 0u, 0u, 0u);
      if (yy >= 0 && yy < p.h && x >= 0 && x < p.w)
        regs[i] = *reinterpret_cast<const uint4*>(in + ((int64_t)yy * p.w + x) * ld + cb * 64 + vv * V);
    }
  };
  uint4 nxt[NV];
  fetch(y0 - 1, 0, nxt);
  for (int yy = y0 - 1; yy <= y1; ++yy) {
    f32x2 P2[9];                                 // per-tap dot products, even / odd channels in the two halves (FFMA2)
#pragma unroll
    for (int t = 0; t < 9; ++t) P2[t] = pack2(0.f, 0.f);
    for (int cb = 0; cb < nblk; ++cb) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int j = lane + 32 * i;
        *reinterpret_cast<uint4*>(tile + (j / NV) * PITCH + (j % NV) * 16) = nxt[i];
      }
      __syncwarp();
      // prefetch the next block / row while this one is consumed
      if (cb + 1 < nblk) fetch(yy, cb + 1, nxt);
      else if (yy + 1 <= y1) fetch(yy + 1, 0, nxt);
      const float* wb = sw + cb * 64;
#pragma unroll 2
      for (int vv = 0; vv < NV; ++vv) {
        const uint4 raw = *reinterpret_cast<const uint4*>(tile + lane * PITCH + vv * 16);
        float v[V];
        const TA* e = reinterpret_cast<const TA*>(&raw);
#pragma unroll
        for (int k = 0; k < V; ++k) v[k] = Cvt<TA>::to_f(e[k]);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
          for (int k4 = 0; k4 < V; k4 += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wb + t * C + vv * V + k4);     // warp-uniform: broadcast
            P2[t] = fma2(pack2(v[k4 + 0], v[k4 + 1]), pack2(w4.x, w4.y), P2[t]);
            P2[t] = fma2(pack2(v[k4 + 2], v[k4 + 3]), pack2(w4.z, w4.w), P2[t]);
          }
        }
      }
    }
    float P[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) { float lo, hi; unpack2(P2[t], lo, hi); P[t] = lo + hi; }
    // P[(dy+1)*3 + (dx+1)] is this pixel's contribution to out(yy - dy, x - dx)
    float q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      q[d] = r1[d] + P[6 + d];                 // dy = +1 completes output row yy - 1
      r1[d] = r0[d] + P[3 + d];                // dy =  0 -> output row yy
      r0[d] = P[d];                            // dy = -1 -> output row yy + 1
    }
    // out(y, x) = q[dx=-1] of column x-1  +  q[dx=0] of column x  +  q[dx=+1] of column x+1
    // tap index d = dx + 1 is the tap the INPUT column applies: input column x-1 feeds out(x) with dx = -1,
    // i.e. the weight column index 0 ... careful: in(y+dy, x+dx) * w[dy][dx]; input column xi = x+dx -> dx = xi - x
    const float from_left = __shfl_up_sync(0xffffffffu, q[0], 1);     // column x-1 holds dx = -1 -> d = 0
    const float from_right = __shfl_down_sync(0xffffffffu, q[2], 1);  // column x+1 holds dx = +1 -> d = 2
    const int y = yy - 1, x = x0 - 1 + lane;
    if (y >= y0 && y < y1 && lane >= 1 && lane <= HEAD_COLS && x < p.w) {
      const float acc = from_left + q[1] + from_right;
      float v = (acc + bias) * p.alpha;
      const int64_t pix = ((int64_t)b * p.h + y) * p.w + x;
      if (p.residual) v = fmaf(p.beta, Cvt<TO>::to_f(((const TO*)p.residual)[pix * p.res_ld]), v);
      v = apply_act(v, p.act);
      ((TO*)p.out)[pix * p.out_ld + p.out_coff] = Cvt<TO>::from_f(v);
    }
  }
}

template <typename TA, typename TO>
int launch_simt(const SimtP& p, cudaStream_t st) {
  const int hw_o = p.ho * p.wo;
  const bool plain = p.a_batched && p.w_bstride == 0 && p.w_ld == p.ktot;
  const bool s1 = p.stride == 1 && p.nseg == 1 && p.a_taps[0] == 9 && p.pad == 1;
  const int64_t rows = (int64_t)p.batch * p.h;
  if constexpr (sizeof(TA) == 4) {
    if (plain && s1 && p.a_c[0] == 1 && p.n % 8 == 0 && p.n <= 256 && !p.residual && rows < (1LL << 31) &&
        (p.out_ld % 8 == 0) && (p.out_coff % 8 == 0)) {
      const int nv = p.n / 8;
      int block = (256 / nv) * nv;
      SimtP q = p; q.rpb = stem_rpb(rows);
      conv_stem_kernel<TO><<<(unsigned)((rows + q.rpb - 1) / q.rpb), block, 0, st>>>(q);
      return mudiff_launch_status();
    }
  }
  if constexpr (sizeof(TA) == 4) {
    if (plain && p.stride == 2 && p.pad == 0 && p.nseg == 1 && p.a_taps[0] == 9 && p.a_c[0] == 1 && p.n % 8 == 0 && p.n <= 256 &&
        !p.residual && (p.out_ld % 8 == 0) && (p.out_coff % 8 == 0) && p.ho > 0 && p.wo > 0) {
      const int nv = p.n / 8;
      const int block = (256 / nv) * nv;
      const int64_t orows = (int64_t)p.batch * p.ho;
      if (orows < (1LL << 31)) {
        SimtP q = p; q.rpb = stem_rpb(orows);
        conv_stem_s2_kernel<TO><<<(unsigned)((orows + q.rpb - 1) / q.rpb), block, 0, st>>>(q);
        return mudiff_launch_status();
      }
    }
  }
  if (plain && s1 && p.n == 1 && p.a_c[0] % 64 == 0 && p.a_c[0] <= 512 && p.a_ld[0] % (16 / (int)sizeof(TA)) == 0 &&
      ((uintptr_t)p.a[0] % 16 == 0)) {
    // strip height: HEAD_YR rows per warp for large launches (a strip re-reads two halo rows), fewer when the launch would not
    // fill the SMs (batch 1 at 256^2: 9 blocks, 99 us; with 4-row strips 72 blocks).  Identical results for any height.
    const int xchunks = (p.w + HEAD_COLS - 1) / HEAD_COLS;
    int yr = (int)(((int64_t)p.batch * p.h * xchunks) / (8 * 2 * MUDIFF_NUM_SMS));
    yr = yr < 2 ? 2 : (yr > HEAD_YR ? HEAD_YR : yr);
    const int ystrips = (p.h + yr - 1) / yr;
    const int64_t units = (int64_t)p.batch * ystrips * xchunks;
    const int64_t blocks = (units + 7) / 8;
    if (blocks < (1LL << 31)) {
      const size_t smem = (size_t)9 * p.a_c[0] * 4 + 8 * 32 * (64 * sizeof(TA) + 16);
      static bool attr_done[2][2] = {};
      bool& done = attr_done[sizeof(TA) == 4][sizeof(TO) == 4];
      if (smem > 48 * 1024 && !done) {
        if (cudaFuncSetAttribute(conv_head_kernel<TA, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
          return (int)cudaGetLastError();
        done = true;
      }
      conv_head_kernel<TA, TO><<<(unsigned)blocks, 256, smem, st>>>(p, xchunks, ystrips, yr);
      return mudiff_launch_status();
    }
  }
  if (plain && p.n <= 4 && (size_t)p.n * p.ktot * 4 <= 48 * 1024) {
    int64_t threads = (int64_t)p.batch * hw_o * 8;
    conv_small_n_kernel<TA, TO><<<grid_for(threads, 256), 256, sizeof(float) * p.n * p.ktot, st>>>(p);
    return mudiff_launch_status();
  }
  if (plain && p.ktot <= 36 && p.n % 8 == 0 && (size_t)p.n * p.ktot * 4 <= 48 * 1024) {
    int64_t threads = (int64_t)p.batch * hw_o * (p.n / 8);
    conv_small_k_kernel<TA, TO><<<grid_for(threads, 256), 256, sizeof(float) * p.n * p.ktot, st>>>(p);
    return mudiff_launch_status();
  }
  if (p.batch > 65535) return MUDIFF_EUNSUPPORTED;
  dim3 grid((hw_o + BM - 1) / BM, (p.n + BN - 1) / BN, p.batch);
  conv_simt_kernel<TA, TO><<<grid, 256, 0, st>>>(p);
  return mudiff_launch_status();
}

}  // namespace

extern "C" int mudiff_conv_simt(const mudiff_conv_desc* d, int dtype, void* stream) {
  if (!d || d->nseg < 1 || d->nseg > 3 || d->batch <= 0 || d->h <= 0 || d->w <= 0 || d->n <= 0) return MUDIFF_EINVAL;
  if (d->stride != 1 && d->stride != 2) return MUDIFF_EINVAL;
  if (!d->wt || !d->out) return MUDIFF_EINVAL;
  if (d->stats) return MUDIFF_EUNSUPPORTED;
  if (d->a_xform[0] || d->a_xform[1] || d->a_xform[2]) return MUDIFF_EUNSUPPORTED;   // tcgen05 kernel only
  SimtP p;
  int koff = 0;
  bool any9 = false;
  for (int s = 0; s < 3; ++s) {
    if (s < d->nseg) {
      if (!d->a[s] || d->a_c[s] <= 0 || (d->a_taps[s] != 1 && d->a_taps[s] != 9)) return MUDIFF_EINVAL;
      p.a[s] = d->a[s]; p.a_c[s] = d->a_c[s]; p.a_ld[s] = d->a_ld[s]; p.a_taps[s] = d->a_taps[s]; p.a_koff[s] = koff;
      koff += d->a_taps[s] * d->a_c[s];
      any9 |= d->a_taps[s] == 9;
    } else { p.a[s] = nullptr; p.a_c[s] = 0; p.a_ld[s] = 0; p.a_taps[s] = 0; p.a_koff[s] = 0; }
  }
  p.nseg = d->nseg; p.a_batched = d->a_batched;
  p.batch = d->batch; p.h = d->h; p.w = d->w; p.stride = d->stride; p.pad = d->pad;
  if (d->stride == 1) { p.ho = d->h; p.wo = d->w; if (any9 && d->pad != 1) return MUDIFF_EUNSUPPORTED; }
  else {
    int k = any9 ? 3 : 1;
    p.ho = (d->h + 2 * d->pad - k) / 2 + 1; p.wo = (d->w + 2 * d->pad - k) / 2 + 1;
  }
  p.wt = d->wt; p.w_bstride = d->w_bstride; p.ktot = koff; p.w_ld = d->w_ld > 0 ? d->w_ld : koff; p.n = d->n;
  p.bias = d->bias; p.rowbias = d->rowbias; p.rowbias_ld = d->rowbias_ld;
  p.residual = d->residual; p.res_ld = d->res_ld; p.alpha = d->alpha; p.beta = d->beta; p.act = d->act;
  p.out = d->out; p.out_ld = d->out_ld; p.out_coff = d->out_coff;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MUDIFF_F32 && d->out_dtype == MUDIFF_F32) return launch_simt<float, float>(p, st);
  if (dtype == MUDIFF_F32 && d->out_dtype == MUDIFF_BF16) return launch_simt<float, __nv_bfloat16>(p, st);
  if (dtype == MUDIFF_BF16 && d->out_dtype == MUDIFF_BF16) return launch_simt<__nv_bfloat16, __nv_bfloat16>(p, st);
  if (dtype == MUDIFF_BF16 && d->out_dtype == MUDIFF_F32) return launch_simt<__nv_bfloat16, float>(p, st);
  return MUDIFF_EUNSUPPORTED;
}

// scratch of the moments kernel (block partials + tickets), one stream of use per device
static double* g_mom_partial[16] = {nullptr};
static size_t g_mom_cap[16] = {0};
static unsigned int* g_mom_tickets[16] = {nullptr};
static int g_mom_tcap[16] = {0};

// moments[b] = {s_t (9), R_tu for t <= u (45)} of the zero-padded 3x3 patches of the 1-channel image x[b] (fp32).
extern "C" int mudiff_stem_moments(const float* x, int ld, int batch, int h, int w, double* moments, void* stream) {
  if (!x || !moments || batch <= 0 || h <= 0 || w <= 0 || ld < 1) return MUDIFF_EINVAL;
  if (batch > 65535) return MUDIFF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  const int chunks = (h + MOM_ROWS - 1) / MOM_ROWS;
  const size_t need = (size_t)batch * chunks * MOM_N;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (need > g_mom_cap[dev]) {
    if (cs != cudaStreamCaptureStatusNone) return MUDIFF_EUNSUPPORTED;     // must be sized by a warm-up run
    double* pnew = nullptr;
    if (cudaMalloc(&pnew, need * 2 * sizeof(double)) != cudaSuccess) return (int)cudaGetLastError();
    g_mom_partial[dev] = pnew; g_mom_cap[dev] = need * 2;
  }
  if (batch > g_mom_tcap[dev]) {
    if (cs != cudaStreamCaptureStatusNone) return MUDIFF_EUNSUPPORTED;
    const int cap = batch * 2 < 4096 ? 4096 : batch * 2;
    unsigned int* t = nullptr;
    if (cudaMalloc(&t, cap * sizeof(unsigned int)) != cudaSuccess) return (int)cudaGetLastError();
    cudaMemset(t, 0, cap * sizeof(unsigned int));
    g_mom_tickets[dev] = t; g_mom_tcap[dev] = cap;
  }
  stem_moments_kernel<<<dim3(chunks, batch), 256, 0, st>>>(x, ld, h, w, moments, g_mom_partial[dev], g_mom_tickets[dev]);
  return mudiff_launch_status();
}

// out = act(GroupNorm_groups(conv3x3(x; wt, bias)) * gamma + beta) for a 1-channel fp32 input x [B, H, W] (pixel
// stride ld), wt fp32 [n][9]; gamma/beta fp32 [B][...] with row stride gb_bstride (NULL = plain GroupNorm);
// scale_shift = caller-provided float[batch][n][2] workspace (the folded per-image scale / shift).
extern "C" int mudiff_stem_conv_gn_act(const float* x, int ld, const float* wt, const float* bias, const double* moments,
                                       const float* gamma, const float* beta, int64_t gb_bstride, int groups, float eps,
                                       int act, float* scale_shift, void* out, int out_ld, int out_coff, int out_dtype,
                                       int batch, int h, int w, int n, void* stream) {
  if (!x || !wt || !moments || !out || !scale_shift || batch <= 0 || h <= 0 || w <= 0 || n <= 0 || groups <= 0) return MUDIFF_EINVAL;
  if (act != MUDIFF_ACT_NONE && act != MUDIFF_ACT_SILU) return MUDIFF_EINVAL;
  if (n % 8 || n > 256 || n % groups || groups > 64 || out_ld % 8 || out_coff % 8 || ((uintptr_t)out % 16)) return MUDIFF_EUNSUPPORTED;
  const int64_t rows = (int64_t)batch * h;
  if (rows >= (1LL << 31)) return MUDIFF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  stem_scale_shift_kernel<<<batch, 256, 0, st>>>(moments, wt, bias, gamma, beta, gb_bstride, n, groups, eps,
                                                 (double)h * (double)w, scale_shift);
  int rc = mudiff_launch_status();
  if (rc) return rc;
  // tensor-core stem (stem_tc.cuh): correct, but measured SLOWER than the CUDA-core kernel below (355 vs 308 us per
  // launch at B = 64, 256^2: its 128-thread epilogue is the bottleneck), so it is opt-in: MUDIFF_STEM_TC=1
  static int use_tc = -1;
  if (use_tc < 0) { const char* e = getenv("MUDIFF_STEM_TC"); use_tc = (e && e[0] == '1') ? 1 : 0; }
  if (use_tc && ld == 1 && n % 32 == 0 && out_dtype == MUDIFF_BF16) {      // the fp32 parity path stays on exact fp32 FMAs
    rc = mudiff_stem_conv_tc(x, wt, bias, scale_shift, act, out, out_ld, out_coff, out_dtype, batch, h, w, n, stream);
    if (rc != MUDIFF_EUNSUPPORTED) return rc;
  }
  SimtP p;
  memset(&p, 0, sizeof(p));
  p.a[0] = x; p.a_c[0] = 1; p.a_ld[0] = ld; p.a_taps[0] = 9; p.nseg = 1; p.a_batched = 1;
  p.batch = batch; p.h = h; p.w = w; p.ho = h; p.wo = w; p.stride = 1; p.pad = 1;
  p.wt = wt; p.ktot = 9; p.w_ld = 9; p.n = n; p.bias = bias; p.alpha = 1.f; p.act = act;
  p.out = out; p.out_ld = out_ld; p.out_coff = out_coff;
  const int nv = n / 8;
  const int block = (256 / nv) * nv;
  p.rpb = stem_rpb(rows);
  const unsigned grid = (unsigned)((rows + p.rpb - 1) / p.rpb);
  const size_t smem = 4 * (size_t)((w + 2 + 3) & ~3) * sizeof(float);          // ring of four padded input rows
  if (smem > 48 * 1024) return MUDIFF_EUNSUPPORTED;
  if (out_dtype == MUDIFF_BF16) conv_stem_gn_kernel<__nv_bfloat16><<<grid, block, smem, st>>>(p, scale_shift);
  else if (out_dtype == MUDIFF_F32) conv_stem_gn_kernel<float><<<grid, block, smem, st>>>(p, scale_shift);
  else return MUDIFF_EUNSUPPORTED;
  return mudiff_launch_status();
}
