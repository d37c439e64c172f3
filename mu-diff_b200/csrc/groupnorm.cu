// GroupNorm statistics + (AdaGN scale/shift) + activation on NHWC tensors, with the
// channel-concat of two sources folded in (no materialised torch.cat).
//   stats : per (batch, group) (sum, sumsq) in fp64, block-partial -> atomicAdd(double)
//   apply : y = act(x * scale[b,c] + shift[b,c]),  scale = gamma*rstd, shift = beta - mean*scale
// HBM-bound: stats reads the tensor once, apply reads once + writes once, 16-byte vectors.
#include <stdlib.h>
#include "common.cuh"

namespace {

#define GN_MAX_C 1024

// Deterministic and batch-invariant: a block always owns GN_PPB consecutive pixels of one image,
// threads reduce in a fixed order, block partials go to `partial[b][chunk][g]` and the LAST block of
// each image (self-resetting ticket counter) adds them up in chunk order.  No floating-point atomics.
// pixels per block: a function of (hw, C) only (=> batch-invariant): ~256 KB of bf16 per block (MUDIFF_GN_CHUNK_KB; 64 KB
// chunks cost the stand-alone statistics pass 28 -> 36 ms per bench step), at least 32 blocks per image for large images
static inline int gn_ppb(int64_t hw, int C) {
  static int chunk_elems = 0;
  if (!chunk_elems) { const char* e = getenv("MUDIFF_GN_CHUNK_KB"); chunk_elems = (e ? atoi(e) : 256) * 512; if (chunk_elems < 8192) chunk_elems = 8192; }
  int64_t a = chunk_elems / C; if (a < 128) a = 128;
  int64_t b = hw / 32; if (b < 128) b = 128;
  return (int)(a < b ? a : b);
}
// Statistics of one block's pixel chunk (blockIdx.x) of image blockIdx.y; returns true in the LAST block of the image, which
// has then written the image's per-channel-group totals to `stats`.  Shared by gn_stats_kernel and gn_l2_kernel (bit-identical).
template <typename T>
__device__ __forceinline__ bool gn_stats_block(const T* __restrict__ x0, int c0, int ld0, const T* __restrict__ x1, int c1, int ld1,
                                               int64_t hw, int groups, double* stats, int st_ld, int st_off,
                                               double* partial, unsigned int* tickets, int GN_PPB, float* s_part, bool* s_last,
                                               double* s_tot = nullptr) {
  constexpr int V = 16 / sizeof(T);
  const int C = c0 + c1;
  const int cv = C / V;
  const int b = blockIdx.y;
  const int chunks = gridDim.x;
  // blockDim.x is a multiple of cv: each thread keeps one channel vector for the whole loop
  const int my_cv = threadIdx.x % cv;
  const int lane = threadIdx.x / cv;
  const int lanes = blockDim.x / cv;
  const int ch = my_cv * V;
  const T* base; int ld; int cc;
  if (ch < c0) { base = x0 + (int64_t)b * hw * ld0; ld = ld0; cc = ch; }
  else { base = x1 + (int64_t)b * hw * ld1; ld = ld1; cc = ch - c0; }
  float sum[V], sq[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { sum[i] = 0.f; sq[i] = 0.f; }
  const int64_t p0 = (int64_t)blockIdx.x * GN_PPB;
  int64_t p1 = p0 + GN_PPB; if (p1 > hw) p1 = hw;
  constexpr int U = 8;                        // 8 x 16 B in flight per thread
  int64_t p = p0 + lane;
  for (; p + (int64_t)(U - 1) * lanes < p1; p += (int64_t)lanes * U) {       // full groups: no predication
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) raw[u] = *reinterpret_cast<const uint4*>(base + (p + (int64_t)u * lanes) * ld + cc);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const T* e = reinterpret_cast<const T*>(&raw[u]);
#pragma unroll
      for (int i = 0; i < V; ++i) { const float f = Cvt<T>::to_f(e[i]); sum[i] += f; sq[i] = fmaf(f, f, sq[i]); }
    }
  }
  for (; p < p1; p += lanes) {                                               // tail
    float v[V];
    load_vec<T>(base + p * ld + cc, v);
#pragma unroll
    for (int i = 0; i < V; ++i) { sum[i] += v[i]; sq[i] = fmaf(v[i], v[i], sq[i]); }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    s_part[((size_t)lane * C + ch + i) * 2 + 0] = sum[i];
    s_part[((size_t)lane * C + ch + i) * 2 + 1] = sq[i];
  }
  __syncthreads();
  const int cpg = C / groups;
  double* my_partial = partial + ((int64_t)b * chunks + blockIdx.x) * groups * 2;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < cpg; ++i)
      for (int l = 0; l < lanes; ++l) {
        a += (double)s_part[((size_t)l * C + g * cpg + i) * 2 + 0];
        q += (double)s_part[((size_t)l * C + g * cpg + i) * 2 + 1];
      }
    my_partial[g * 2 + 0] = a;
    my_partial[g * 2 + 1] = q;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(&tickets[b], 1u);
    *s_last = (t == (unsigned int)chunks - 1);
    if (*s_last) tickets[b] = 0;                // self-reset for the next launch on this stream
  }
  __syncthreads();
  if (!*s_last) return false;
  __threadfence();
  // totals = chunk partials added in a FIXED order (deterministic): `segs` threads per value, each adds a contiguous run of
  // chunks (8 independent L2 loads in flight), the run sums are then added in run order.  In gn_l2_kernel every other block
  // of the image waits for this, so it must not be a chain of `chunks` dependent L2 round trips.
  const double* pb = partial + (int64_t)b * chunks * groups * 2;
  const int nvals = groups * 2;
  int segs = (int)blockDim.x / nvals;
  if (segs < 1) segs = 1;
  if (segs > 8) segs = 8;
  const int cps = (chunks + segs - 1) / segs;
  double* s_run = reinterpret_cast<double*>(s_part);      // [segs][nvals] (the float partials are dead by now)
  __syncthreads();
  for (int idx = threadIdx.x; idx < nvals * segs; idx += blockDim.x) {
    const int i = idx % nvals, sg = idx / nvals;
    const int k0 = sg * cps, k1 = (k0 + cps < chunks) ? k0 + cps : chunks;
    double a = 0.0;
    int k = k0;
    for (; k + 16 <= k1; k += 16) {                        // one L2 round trip for the usual 32 chunks x 2 runs (same order of adds)
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = __ldcg(pb + (int64_t)(k + u) * nvals + i);
#pragma unroll
      for (int u = 0; u < 16; ++u) a += v[u];
    }
    for (; k + 8 <= k1; k += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(pb + (int64_t)(k + u) * nvals + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) a += v[u];
    }
    for (; k < k1; ++k) a += __ldcg(pb + (int64_t)k * nvals + i);
    s_run[(size_t)sg * nvals + i] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nvals; i += blockDim.x) {
    double a = 0.0;
    for (int sg = 0; sg < segs; ++sg) a += s_run[(size_t)sg * nvals + i];
    stats[((int64_t)b * st_ld + st_off) * 2 + i] = a;
    if (s_tot) s_tot[i] = a;                               // gn_stats_table_kernel goes on in this block: no global round trip
  }
  return true;
}

template <typename T>
__global__ void __launch_bounds__(256, 3) gn_stats_kernel(const T* __restrict__ x0, int c0, int ld0, const T* __restrict__ x1, int c1, int ld1,
                                int64_t hw, int groups, double* __restrict__ stats, int st_ld, int st_off,
                                double* __restrict__ partial, unsigned int* __restrict__ tickets, int GN_PPB) {
  extern __shared__ float s_part[];            // [lanes][C][2]
  __shared__ bool s_last;
  gn_stats_block<T>(x0, c0, ld0, x1, c1, ld1, hw, groups, stats, st_ld, st_off, partial, tickets, GN_PPB, s_part, &s_last);
}

// apply: each thread owns ONE channel vector (scale/shift live in registers) and walks pixels with a
// fixed stride -> no integer division in the streaming loop, 4 independent 16-byte loads in flight.
#define GN_APPLY_PPB 1024
// Apply phase of one block on the pixels [p0, p1) of image blockIdx.y (shared by gn_apply_kernel and gn_l2_kernel).  The
// statistics are read with ld.global.cg: in gn_l2_kernel another block of the SAME launch has just written them.
template <typename TI, typename TO>
__device__ __forceinline__ void gn_apply_block(const TI* __restrict__ x0, int c0, int ld0, const TI* __restrict__ x1, int c1, int ld1,
                                const double* st0, int st0_ld, const double* st1, int st1_ld,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int64_t gb_bstride,
                                TO* __restrict__ out, int ld_out, int64_t hw, int groups, float eps, int act,
                                const int64_t p0, const int64_t p1, float* s_mean, float* s_rstd, const int b) {
  // vector width is chosen on the WIDER of the two element types so both sides stay <= 16 bytes
  constexpr int V = (sizeof(TI) >= sizeof(TO)) ? 16 / sizeof(TI) : 16 / sizeof(TO);
  const int C = c0 + c1;
  const int cv = C / V;
  const int cpg = C / groups;
  // per-channel (sum, sumsq) of the two sources -> per-group mean / rstd (double math, once per block)
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < cpg; ++i) {
      const int c = g * cpg + i;
      const double* sp = c < c0 ? st0 + ((int64_t)b * st0_ld + c) * 2 : st1 + ((int64_t)b * st1_ld + (c - c0)) * 2;
      a += __ldcg(sp); q += __ldcg(sp + 1);
    }
    const double cnt = (double)hw * (double)cpg;
    const double m = a / cnt;
    double var = q / cnt - m * m;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)m;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int my_cv = threadIdx.x % cv;
  const int lane = threadIdx.x / cv;
  const int lanes = blockDim.x / cv;
  if (lane >= lanes) return;                  // gn_l2_kernel: the block size follows x0's channel count (phase 1)
  const int ch = my_cv * V;
  float sc[V], sh[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int c = ch + k;
    const int g = c / cpg;
    const float ga = gamma ? gamma[(int64_t)b * gb_bstride + c] : 1.f;
    const float be = beta ? beta[(int64_t)b * gb_bstride + c] : 0.f;
    sc[k] = ga * s_rstd[g];
    sh[k] = be - s_mean[g] * sc[k];
  }
  const TI* src; int ld;
  if (ch < c0) { src = x0 + (int64_t)b * hw * ld0 + ch; ld = ld0; }
  else { src = x1 + (int64_t)b * hw * ld1 + (ch - c0); ld = ld1; }
  TO* dst = out + (int64_t)b * hw * ld_out + ch;
  if constexpr (sizeof(TI) == sizeof(TO)) {
    // same element size on both sides: 8 raw 16-byte loads in flight per thread (latency-bound otherwise: the
    // kernel sits at ~3 blocks/SM), then convert -> scale/shift -> activation -> pack -> store one at a time
    constexpr int U = 8;
    int64_t p = p0 + lane;
    for (; p + (int64_t)(U - 1) * lanes < p1; p += (int64_t)lanes * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = *reinterpret_cast<const uint4*>(src + (p + (int64_t)u * lanes) * ld);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[V];
        const TI* e = reinterpret_cast<const TI*>(&raw[u]);
        if constexpr (sizeof(TI) == 2) {
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
#pragma unroll
          for (int k = 0; k < V / 2; ++k) { const float2 f = __bfloat1622float2(h2[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) v[k] = Cvt<TI>::to_f(e[k]);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float t = fmaf(v[k], sc[k], sh[k]);
          if (act == MUDIFF_ACT_SILU) t = (sizeof(TO) == 4) ? silu_exact(t) : silu_f(t);
          v[k] = t;
        }
        store_vec<TO>(dst + (p + (int64_t)u * lanes) * ld_out, v);
      }
    }
    for (; p < p1; p += lanes) {
      float v[V];
      load_vec<TI>(src + p * ld, v);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float t = fmaf(v[k], sc[k], sh[k]);
        if (act == MUDIFF_ACT_SILU) t = (sizeof(TO) == 4) ? silu_exact(t) : silu_f(t);
        v[k] = t;
      }
      store_vec<TO>(dst + p * ld_out, v);
    }
  } else {
  constexpr int U = 4;
  for (int64_t p = p0 + lane; p < p1; p += (int64_t)lanes * U) {
    float v[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pp = p + (int64_t)u * lanes;
      if (pp < p1) {
        if constexpr (V * sizeof(TI) == 16) {
          load_vec<TI>(src + pp * ld, *reinterpret_cast<float(*)[16 / sizeof(TI)]>(v[u]));
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) v[u][k] = Cvt<TI>::to_f(src[pp * ld + k]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pp = p + (int64_t)u * lanes;
      if (pp < p1) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float t = fmaf(v[u][k], sc[k], sh[k]);
          if (act == MUDIFF_ACT_SILU) t = (sizeof(TO) == 4) ? silu_exact(t) : silu_f(t);
          v[u][k] = t;
        }
        if constexpr (V * sizeof(TO) == 16) {
          store_vec<TO>(dst + pp * ld_out, *reinterpret_cast<float(*)[16 / sizeof(TO)]>(v[u]));
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) dst[pp * ld_out + k] = Cvt<TO>::from_f(v[u][k]);
        }
      }
    }
  }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256, 3) gn_apply_kernel(const TI* __restrict__ x0, int c0, int ld0, const TI* __restrict__ x1, int c1, int ld1,
                                const double* __restrict__ st0, int st0_ld, const double* __restrict__ st1, int st1_ld,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int64_t gb_bstride,
                                TO* __restrict__ out, int ld_out, int64_t hw, int groups, float eps, int act) {
  __shared__ float s_mean[GN_MAX_C / 4], s_rstd[GN_MAX_C / 4];
  // (walking the batch backwards, so that the first reads hit what the statistics pass left in L2, was measured: 43.6 -> 43.2 ms
  // per step at B = 64 - the tensors are 4 - 8x the L2; not kept)
  const int b = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * GN_APPLY_PPB;
  int64_t p1 = p0 + GN_APPLY_PPB; if (p1 > hw) p1 = hw;
  gn_apply_block<TI, TO>(x0, c0, ld0, x1, c1, ld1, st0, st0_ld, st1, st1_ld, gamma, beta, gb_bstride, out, ld_out, hw, groups, eps, act,
                         p0, p1, s_mean, s_rstd, b);
}

// ---------------------------------------------------------------------------------
// Statistics + apply in ONE launch with the second read served by L2 ("gn_l2"): block (chunk, image) computes the partial
// statistics of its chunk exactly like gn_stats_kernel, the last block of the image publishes the totals and raises the
// image's flag, every block of the image then normalises ITS OWN chunk - which it read a few microseconds ago, so the re-read
// hits L2 when the chunks of all resident blocks fit there: 2 HBM passes per GroupNorm instead of 3 (see ops.GN_L2 for when
// that pays).
// Blocks of one image are consecutive in launch order, so they are co-resident (the host only takes this path when an image
// has at most kGnL2MaxChunks chunks); the flag / arrival counters reset themselves (CUDA-graph replays need no host reset).
// Results are bit-identical to gn_stats_kernel + gn_apply_kernel (same device functions, same chunking).
// x0: statistics unknown (computed here, stored to st0); optional x1 (channel-concat partner) with known statistics st1.
// ---------------------------------------------------------------------------------
constexpr int kGnL2MaxChunks = 256;
template <typename T>
__global__ void __launch_bounds__(256, 3) gn_l2_kernel(const T* __restrict__ x0, int c0, int ld0, const T* __restrict__ x1, int c1, int ld1,
                                double* st0, int st0_ld, const double* st1, int st1_ld, int groups0,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int64_t gb_bstride,
                                T* __restrict__ out, int ld_out, int64_t hw, int groups, float eps, int act,
                                double* partial, unsigned int* tickets, unsigned int* flags, unsigned int* done, int GN_PPB, unsigned int poll_ns) {
  extern __shared__ float s_part[];            // [lanes][c0][2]
  __shared__ float s_mean[GN_MAX_C / 4], s_rstd[GN_MAX_C / 4];
  __shared__ bool s_last;
  const int b = blockIdx.y;
  // phase 1: per-channel statistics of x0 (groups0 == c0: one "group" per channel, as mudiff_gn_stats)
  const bool last = gn_stats_block<T>(x0, c0, ld0, nullptr, 0, 0, hw, groups0, st0, st0_ld, 0, partial, tickets, GN_PPB, s_part, &s_last);
  if (last) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(&flags[b * 32], 1u);
  } else if (threadIdx.x == 0) {
    // one poller per block, one 128-byte line per image, sleeping between polls: hundreds of blocks hammering one L2
    // sector with back-to-back acquire loads starved the very atomics that release them
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + b * 32) : "memory");
      if (!v) __nanosleep(poll_ns);
    } while (!v);
  }
  __syncthreads();
  // phase 2: normalise this block's own chunk (L2-resident)
  const int64_t p0 = (int64_t)blockIdx.x * GN_PPB;
  int64_t p1 = p0 + GN_PPB; if (p1 > hw) p1 = hw;
  gn_apply_block<T, T>(x0, c0, ld0, x1, c1, ld1, st0, st0_ld, st1, st1_ld, gamma, beta, gb_bstride, out, ld_out, hw, groups, eps, act,
                       p0, p1, s_mean, s_rstd, b);
  // every block of the image has read the flag and the statistics once it arrives here: the last arrival resets them
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&done[b], 1u);
    if (t == gridDim.x - 1) { done[b] = 0; atomicExch(&flags[b * 32], 0u); }
  }
}

// Per-tile per-channel (sum, sumsq) float partials written by the conv epilogue -> per-channel doubles.
// One block per image, ordered summation over the tiles (deterministic, batch-invariant).
// rows = partial rows per image (tiles x 4 lane quadrants).  Block (b, column block of 32): thread (col, seg) adds the rows
// r = seg, seg + 32, ... of its column in order (8 independent loads in flight), the 32 segment sums are then added in
// segment order in double precision: the order depends on `rows` only -> deterministic and batch-invariant.
__global__ void __launch_bounds__(1024) stats_finalize_kernel(const float* __restrict__ partial, int rows, int n2,
                                                              double* __restrict__ chstats, int st_ld, int st_off) {
  __shared__ double red[32][33];
  const int b = blockIdx.x;
  const int col = blockIdx.y * 32 + (threadIdx.x & 31);
  const int seg = threadIdx.x >> 5;
  double a = 0.0;
  if (col < n2) {
    const float* pb = partial + (int64_t)b * rows * n2 + col;
    int r = seg;
    for (; r + 7 * 32 < rows; r += 8 * 32) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = pb[(int64_t)(r + 32 * i) * n2];
#pragma unroll
      for (int i = 0; i < 8; ++i) a += (double)v[i];
    }
    for (; r < rows; r += 32) a += (double)pb[(int64_t)r * n2];
  }
  red[seg][threadIdx.x & 31] = a;
  __syncthreads();
  if (seg == 0 && col < n2) {
    double t = 0.0;
#pragma unroll 8
    for (int s2 = 0; s2 < 32; ++s2) t += red[s2][threadIdx.x];
    chstats[((int64_t)b * st_ld + st_off) * 2 + col] = t;
  }
}

// dynamic shared memory of gn_stats_block: float partials [lanes][C][2], reused as double run sums [segs][groups * 2]
static inline size_t gn_stats_smem(int lanes, int C, int groups, int block, bool with_totals = false) {
  int segs = block / (groups * 2); if (segs < 1) segs = 1; if (segs > 8) segs = 8;
  const size_t a = sizeof(float) * 2 * (size_t)lanes * C, r = sizeof(double) * (size_t)segs * groups * 2;
  const size_t m = ((a > r ? a : r) + 15) & ~(size_t)15;
  return m + (with_totals ? sizeof(double) * (size_t)groups * 2 : 0);      // + [groups * 2] totals (gn_stats_table_kernel)
}
// scratch for block partials + ticket counters, grown on demand (single stream of use per device)
struct StatsScratch { double* partial = nullptr; size_t cap = 0; unsigned int* tickets = nullptr; int tcap = 0; };
static StatsScratch g_scratch[16];

static int ensure_scratch(int dev, size_t need_partial, int need_tickets, cudaStream_t st) {
  StatsScratch& s = g_scratch[dev];
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (need_partial > s.cap) {
    if (cs != cudaStreamCaptureStatusNone) return MUDIFF_EUNSUPPORTED;   // must be sized by a warm-up run
    size_t cap = need_partial * 2; if (cap < (1u << 20)) cap = 1u << 20;
    double* p = nullptr;
    if (cudaMalloc(&p, cap * sizeof(double)) != cudaSuccess) return (int)cudaGetLastError();
    // old buffer is leaked on purpose while kernels of earlier launches may still read it (tiny)
    s.partial = p; s.cap = cap;
  }
  if (need_tickets > s.tcap) {
    if (cs != cudaStreamCaptureStatusNone) return MUDIFF_EUNSUPPORTED;
    int cap = need_tickets * 2; if (cap < 4096) cap = 4096;
    unsigned int* t = nullptr;
    if (cudaMalloc(&t, cap * sizeof(unsigned int)) != cudaSuccess) return (int)cudaGetLastError();
    cudaMemset(t, 0, cap * sizeof(unsigned int));
    s.tickets = t; s.tcap = cap;
  }
  return 0;
}

template <typename T>
int launch_stats(const void* x0, int c0, int ld0, const void* x1, int c1, int ld1, int batch, int64_t hw,
                 int groups, double* stats, int st_ld, int st_off, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  const int C = c0 + c1;
  if (C % V || c0 % V || ld0 % V || (x1 && ld1 % V) || C > GN_MAX_C || C % groups) return MUDIFF_EUNSUPPORTED;
  if (((uintptr_t)x0 % 16) || (x1 && ((uintptr_t)x1 % 16))) return MUDIFF_EUNSUPPORTED;
  const int cv = C / V;
  int block = (256 / cv) * cv;
  if (block < cv) block = cv;                // cv <= 256 since C <= 1024, V >= 4
  const int lanes = block / cv;
  const int GN_PPB = gn_ppb(hw, C);
  int chunks = (int)((hw + GN_PPB - 1) / GN_PPB);
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  int rc = ensure_scratch(dev, (size_t)batch * chunks * groups * 2, batch, st);
  if (rc) return rc;
  dim3 grid(chunks, batch);
  size_t smem = gn_stats_smem(lanes, C, groups, block);
  gn_stats_kernel<T><<<grid, block, smem, st>>>((const T*)x0, c0, ld0, (const T*)x1, c1, ld1, hw, groups, stats, st_ld, st_off,
                                               g_scratch[dev].partial, g_scratch[dev].tickets, GN_PPB);
  return mudiff_launch_status();
}

template <typename TI, typename TO>
int launch_apply(const void* x0, int c0, int ld0, const void* x1, int c1, int ld1, const double* st0, int st0_ld,
                 const double* st1, int st1_ld, const float* gamma, const float* beta, int64_t gbs, void* out,
                 int ld_out, int batch, int64_t hw, int groups, float eps, int act, cudaStream_t st) {
  constexpr int V = (sizeof(TI) >= sizeof(TO)) ? 16 / sizeof(TI) : 16 / sizeof(TO);
  const int C = c0 + c1;
  if (C % V || c0 % V || ld0 % V || (x1 && ld1 % V) || ld_out % V || C > GN_MAX_C || C % groups || groups > GN_MAX_C / 4)
    return MUDIFF_EUNSUPPORTED;
  if (((uintptr_t)x0 % 16) || (x1 && ((uintptr_t)x1 % 16)) || ((uintptr_t)out % 16)) return MUDIFF_EUNSUPPORTED;
  const int cv = C / V;
  int block = (256 / cv) * cv;
  if (block < cv) block = cv;
  int chunks = (int)((hw + GN_APPLY_PPB - 1) / GN_APPLY_PPB);
  dim3 grid(chunks, batch);
  gn_apply_kernel<TI, TO><<<grid, block, 0, st>>>((const TI*)x0, c0, ld0, (const TI*)x1, c1, ld1, st0, st0_ld, st1, st1_ld,
                                                  gamma, beta, gbs, (TO*)out, ld_out, hw, groups, eps, act);
  return mudiff_launch_status();
}

// ---------------------------------------------------------------------------------
// Single-pass GroupNorm + AdaGN + activation ("gn_fused"): statistics AND apply with ONE read of the tensor.
// The two-kernel path reads x twice from HBM (the tensors are 0.5 - 2 GB at B = 64, far beyond the 126 MB L2).
// Here a persistent cooperative grid (one CTA per SM) walks the images in order; CTA c owns pixel chunk c of EVERY
// image and keeps it in shared memory between the two phases:
//   L(b)  bulk-async copy (cp.async.bulk + mbarrier) of its chunk of image b into one of D stages
//   S(b)  per-channel (sum, sumsq) of the chunk -> partial[b][c]; the LAST CTA to arrive for image b (ticket) adds the
//         G partials in a fixed order, publishes chstats[b] (same [B, C, 2] doubles as gn_stats) and sets flag[b]
//   A(a)  a = b - (D-1): wait for flag[a], fold mean / rstd / gamma / beta, apply + activation from the shared-memory
//         copy, store; the stage is then refilled with image a + D
// so the latency of the cross-CTA reduction (a few us) is covered by the D - 1 images in flight.  Deterministic and
// batch-invariant: chunking depends on (HW, grid) only, all reductions have a fixed order.  The last CTA to finish
// the launch resets the flags / tickets (CUDA-graph replays need no host-side reset).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gnf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct GnFusedP {
  const __nv_bfloat16* x; __nv_bfloat16* out;
  int C, cv, groups, batch, D;
  long long hw; int ppc;                         // pixels per chunk
  const float* gamma; const float* beta; long long gb_bstride;
  float eps; int act;
  double* chstats; int st_ld, st_off;            // [B][st_ld][2]
  double* partial;                               // [B][G][C][2]
  unsigned int* tickets;                         // [B]
  unsigned int* flags;                           // [B]
  unsigned int* done;                            // [1]
  uint32_t stage_bytes;
};

__global__ void __launch_bounds__(256, 1) gn_fused_kernel(const GnFusedP p) {
  extern __shared__ __align__(128) uint8_t gnf_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(gnf_smem);                 // [D] mbarriers
  float* s_part = reinterpret_cast<float*>(gnf_smem + 64);                // [lanes][C][2] cross-lane reduction
  __shared__ float s_mean[GN_MAX_C / 4], s_rstd[GN_MAX_C / 4];
  __shared__ bool s_last;
  const int C = p.C, cv = p.cv;
  const int lanes = blockDim.x / cv;
  uint8_t* stages = gnf_smem + 64 + ((size_t)lanes * C * 2 * sizeof(float) + 127) / 128 * 128;
  const int G = gridDim.x, c = blockIdx.x;
  const long long px0 = (long long)c * p.ppc;
  long long npx = p.hw - px0; if (npx > p.ppc) npx = p.ppc; if (npx < 0) npx = 0;
  const uint32_t bytes = (uint32_t)(npx * C * 2);
  const int my_cv = threadIdx.x % cv, lane = threadIdx.x / cv, ch = my_cv * 8;
  const int cpg = C / p.groups;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.D; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gnf_smem_u32(&full[i])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue_load = [&](int b) {                 // thread 0 only
    if (bytes == 0) return;
    const int st = b % p.D;
    const uint32_t bar = gnf_smem_u32(&full[st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    const __nv_bfloat16* src = p.x + ((long long)b * p.hw + px0) * C;
    uint8_t* dst = stages + (size_t)st * p.stage_bytes;
    // chunks of <= 64 KB per bulk copy
    for (uint32_t off = 0; off < bytes; off += 65536u) {
      const uint32_t n = bytes - off < 65536u ? bytes - off : 65536u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(gnf_smem_u32(dst + off)), "l"((const uint8_t*)src + off), "r"(n), "r"(bar) : "memory");
    }
  };
  if (threadIdx.x == 0)
    for (int b = 0; b < p.D && b < p.batch; ++b) issue_load(b);

  for (int b = 0; b < p.batch + p.D - 1; ++b) {
    if (b < p.batch) {
      // ------------------------------ S(b) ------------------------------
      const int st = b % p.D;
      const uint32_t parity = (uint32_t)((b / p.D) & 1);
      if (bytes) {
        uint32_t ok = 0; long long t0 = clock64();
        while (!ok) {
          asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                       : "=r"(ok) : "r"(gnf_smem_u32(&full[st])), "r"(parity) : "memory");
          if (!ok && clock64() - t0 > 4000000000LL) __trap();
        }
      }
      const uint8_t* sm = stages + (size_t)st * p.stage_bytes;
      float sum[8], sq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { sum[i] = 0.f; sq[i] = 0.f; }
      if (lane < lanes)
        for (long long q2 = lane; q2 < npx; q2 += lanes) {
          const uint4 raw = *reinterpret_cast<const uint4*>(sm + (q2 * C + ch) * 2);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h2[i]);
            sum[2 * i] += f.x; sq[2 * i] = fmaf(f.x, f.x, sq[2 * i]);
            sum[2 * i + 1] += f.y; sq[2 * i + 1] = fmaf(f.y, f.y, sq[2 * i + 1]);
          }
        }
      __syncthreads();                           // s_part free (previous image's readers are done)
      if (lane < lanes) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s_part[((size_t)lane * C + ch + i) * 2 + 0] = sum[i];
          s_part[((size_t)lane * C + ch + i) * 2 + 1] = sq[i];
        }
      }
      __syncthreads();
      double* mine = p.partial + (((long long)b * G + c) * C) * 2;
      for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        double a = 0.0;
        for (int l = 0; l < lanes; ++l) a += (double)s_part[(size_t)l * C * 2 + i];
        mine[i] = a;
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&p.tickets[b], 1u);
        s_last = (t == (unsigned int)G - 1);
      }
      __syncthreads();
      if (s_last) {                              // last CTA of image b: fixed-order sum of the G partials -> chstats[b]
        __threadfence();
        const double* pb = p.partial + (long long)b * G * C * 2;
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          int k = 0;
          for (; k + 4 <= G; k += 4) {
            a0 += pb[(long long)(k + 0) * C * 2 + i]; a1 += pb[(long long)(k + 1) * C * 2 + i];
            a2 += pb[(long long)(k + 2) * C * 2 + i]; a3 += pb[(long long)(k + 3) * C * 2 + i];
          }
          for (; k < G; ++k) a0 += pb[(long long)k * C * 2 + i];
          p.chstats[((long long)b * p.st_ld + p.st_off) * 2 + i] = (a0 + a1) + (a2 + a3);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) { p.tickets[b] = 0; atomicExch(&p.flags[b], 1u); }
      }
    }
    const int a = b - (p.D - 1);
    if (a >= 0 && a < p.batch) {
      // ------------------------------ A(a) ------------------------------
      if (threadIdx.x == 0) {
        long long t0 = clock64();
        while (atomicAdd(&p.flags[a], 0u) == 0u) { if (clock64() - t0 > 4000000000LL) __trap(); }
      }
      __syncthreads();
      __threadfence();
      for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < cpg; ++i) {
          const double* sp = p.chstats + ((long long)a * p.st_ld + p.st_off + g * cpg + i) * 2;
          s1 += __ldcg(sp); s2 += __ldcg(sp + 1);
        }
        const double cnt = (double)p.hw * (double)cpg;
        const double m = s1 / cnt;
        double var = s2 / cnt - m * m;
        if (var < 0.0) var = 0.0;
        s_mean[g] = (float)m;
        s_rstd[g] = (float)(1.0 / sqrt(var + (double)p.eps));
      }
      __syncthreads();
      const int st = a % p.D;
      if (lane < lanes && npx > 0) {
        float sc[8], sh[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int cc = ch + k, g = cc / cpg;
          const float ga = p.gamma ? p.gamma[(long long)a * p.gb_bstride + cc] : 1.f;
          const float be = p.beta ? p.beta[(long long)a * p.gb_bstride + cc] : 0.f;
          sc[k] = ga * s_rstd[g];
          sh[k] = be - s_mean[g] * sc[k];
        }
        const uint8_t* sm = stages + (size_t)st * p.stage_bytes;
        __nv_bfloat16* dst = p.out + ((long long)a * p.hw + px0) * C + ch;
        for (long long q2 = lane; q2 < npx; q2 += lanes) {
          const uint4 raw = *reinterpret_cast<const uint4*>(sm + (q2 * C + ch) * 2);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
          float v[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h2[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float t = fmaf(v[k], sc[k], sh[k]);
            if (p.act == MUDIFF_ACT_SILU) t = silu_f(t);
            v[k] = t;
          }
          store_vec<__nv_bfloat16>(dst + q2 * C, v);
        }
      }
      __syncthreads();                           // every reader of the stage is done -> refill it
      if (threadIdx.x == 0 && a + p.D < p.batch) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of the stage before the async refill
        issue_load(a + p.D);
      }
    }
  }
  // the last CTA of the launch resets the flags for the next launch (graph replays)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(p.done, 1u);
    if (t == (unsigned int)G - 1) {
      for (int b = 0; b < p.batch; ++b) p.flags[b] = 0;
      __threadfence();
      *p.done = 0;
    }
  }
}

// Folded GroupNorm / AdaGN parameters for consumers that apply the normalisation themselves (mudiff_conv_tc's
// A-operand transform): table[b][c] = (scale, shift) with scale = gamma * rstd, shift = beta - mean * scale.
// One block per image, one thread per channel; same double-precision group statistics as gn_apply_kernel.
// table[b][c] = (scale, shift) of image b (all threads of the block; shared by gn_scale_shift_kernel and gn_stats_table_kernel)
template <bool kSt0Shared>
__device__ __forceinline__ void gn_table_block(const double* st0, int st0_ld, int c0, const double* st1, int st1_ld, int c1,
                                               const float* __restrict__ gamma, const float* __restrict__ beta, int64_t gb_bstride,
                                               double hw, int groups, float eps, float* __restrict__ table, int b,
                                               float* s_mean, float* s_rstd) {
  // kSt0Shared: st0 points at THIS image's [c0][2] totals in shared memory (written by this block a moment ago)
  const int C = c0 + c1, cpg = C / groups;
  // gamma / beta of this thread's first four channels: issued before the statistics math, consumed after it
  float ga_r[4], be_r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = threadIdx.x + k * blockDim.x;
    ga_r[k] = (gamma && c < C) ? gamma[(int64_t)b * gb_bstride + c] : 1.f;
    be_r[k] = (beta && c < C) ? beta[(int64_t)b * gb_bstride + c] : 0.f;
  }
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < cpg; ++i) {
      const int c = g * cpg + i;
      if (c < c0) {
        const double* sp = kSt0Shared ? st0 + (int64_t)c * 2 : st0 + ((int64_t)b * st0_ld + c) * 2;
        if (kSt0Shared) { a += sp[0]; q += sp[1]; } else { a += __ldcg(sp); q += __ldcg(sp + 1); }
      } else {
        const double* sp = st1 + ((int64_t)b * st1_ld + (c - c0)) * 2;
        a += __ldcg(sp); q += __ldcg(sp + 1);
      }
    }
    const double cnt = hw * (double)cpg;
    const double m = a / cnt;
    double var = q / cnt - m * m;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)m;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = threadIdx.x + k * blockDim.x;
    if (c < C) {
      const int g = c / cpg;
      const float sc = ga_r[k] * s_rstd[g];
      table[((int64_t)b * C + c) * 2 + 0] = sc;
      table[((int64_t)b * C + c) * 2 + 1] = be_r[k] - s_mean[g] * sc;
    }
  }
  for (int c = threadIdx.x + 4 * blockDim.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float ga = gamma ? gamma[(int64_t)b * gb_bstride + c] : 1.f;
    const float be = beta ? beta[(int64_t)b * gb_bstride + c] : 0.f;
    const float sc = ga * s_rstd[g];
    table[((int64_t)b * C + c) * 2 + 0] = sc;
    table[((int64_t)b * C + c) * 2 + 1] = be - s_mean[g] * sc;
  }
}

__global__ void gn_scale_shift_kernel(const double* __restrict__ st0, int st0_ld, int c0, const double* __restrict__ st1,
                                      int st1_ld, int c1, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      int64_t gb_bstride, double hw, int groups, float eps, float* __restrict__ table) {
  __shared__ float s_mean[GN_MAX_C / 4], s_rstd[GN_MAX_C / 4];
  gn_table_block<false>(st0, st0_ld, c0, st1, st1_ld, c1, gamma, beta, gb_bstride, hw, groups, eps, table, blockIdx.x, s_mean, s_rstd);
}

// Statistics of x0 AND the folded (scale, shift) table of GroupNorm/AdaGN over [x0 | x1] in one launch: the last block of an
// image (the one that adds up the chunk partials) goes on to write the image's table rows.  Same values as mudiff_gn_stats +
// mudiff_gn_scale_shift; one launch less per fused GroupNorm (the sampling loop at batch 1 is launch-latency bound).
template <typename T>
__global__ void __launch_bounds__(256, 3) gn_stats_table_kernel(const T* __restrict__ x0, int c0, int ld0, double* st0, int st0_ld,
                                int c1, const double* st1, int st1_ld, const float* __restrict__ gamma, const float* __restrict__ beta,
                                int64_t gb_bstride, int64_t hw, int groups, float eps, float* __restrict__ table,
                                double* partial, unsigned int* tickets, int GN_PPB, unsigned int tot_off) {
  extern __shared__ float s_part[];
  __shared__ float s_mean[GN_MAX_C / 4], s_rstd[GN_MAX_C / 4];
  __shared__ bool s_last;
  double* s_tot = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(s_part) + tot_off);     // [c0 * 2] totals of this image
  if (!gn_stats_block<T>(x0, c0, ld0, nullptr, 0, 0, hw, c0, st0, st0_ld, 0, partial, tickets, GN_PPB, s_part, &s_last, s_tot)) return;
  __syncthreads();                               // the totals (shared memory) are visible to all threads of this block
  gn_table_block<true>(s_tot, 0, c0, st1, st1_ld, c1, gamma, beta, gb_bstride, (double)hw, groups, eps, table, blockIdx.y, s_mean, s_rstd);
}

__global__ void gap_mean_kernel(const double* __restrict__ stats, float* __restrict__ out, int n, double inv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)(stats[2 * i] * inv);
}

}  // namespace

// Global average pool = per-channel sums of the deterministic statistics kernel (groups == C).
extern "C" int mudiff_gap(const void* x, int ld, int dtype, float* out, int batch, int64_t hw, int c, void* stream) {
  if (batch <= 0 || hw <= 0 || c <= 0 || c > GN_MAX_C || !x || !out) return MUDIFF_EINVAL;
  if (batch > 65535) return MUDIFF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  static double* gap_stats[16] = {nullptr};
  static size_t gap_cap[16] = {0};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  size_t need = (size_t)batch * c * 2;
  if (need > gap_cap[dev]) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) return MUDIFF_EUNSUPPORTED;
    size_t cap = need * 2; if (cap < 65536) cap = 65536;
    double* p = nullptr;
    if (cudaMalloc(&p, cap * sizeof(double)) != cudaSuccess) return (int)cudaGetLastError();
    gap_stats[dev] = p; gap_cap[dev] = cap;
  }
  int rc;
  if (dtype == MUDIFF_F32) rc = launch_stats<float>(x, c, ld, nullptr, 0, 0, batch, hw, c, gap_stats[dev], c, 0, st);
  else if (dtype == MUDIFF_BF16) rc = launch_stats<__nv_bfloat16>(x, c, ld, nullptr, 0, 0, batch, hw, c, gap_stats[dev], c, 0, st);
  else return MUDIFF_EUNSUPPORTED;
  if (rc) return rc;
  int n = batch * c;
  gap_mean_kernel<<<(n + 255) / 256, 256, 0, st>>>(gap_stats[dev], out, n, 1.0 / (double)hw);
  return mudiff_launch_status();
}

extern "C" int mudiff_gn_stats(const void* x, int c, int ld, int dtype, int batch, int64_t hw,
                               double* chstats, int st_ld, int st_off, void* stream) {
  if (batch <= 0 || hw <= 0 || c <= 0 || !x || !chstats || st_ld < c + st_off) return MUDIFF_EINVAL;
  if (batch > 65535) return MUDIFF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MUDIFF_F32) return launch_stats<float>(x, c, ld, nullptr, 0, 0, batch, hw, c, chstats, st_ld, st_off, st);
  if (dtype == MUDIFF_BF16) return launch_stats<__nv_bfloat16>(x, c, ld, nullptr, 0, 0, batch, hw, c, chstats, st_ld, st_off, st);
  return MUDIFF_EUNSUPPORTED;
}

extern "C" int mudiff_stats_finalize(const float* partial, int tiles_per_image, int n, double* chstats, int st_ld,
                                     int st_off, int batch, void* stream) {
  if (!partial || !chstats || tiles_per_image <= 0 || n <= 0 || batch <= 0 || st_ld < n + st_off) return MUDIFF_EINVAL;
  stats_finalize_kernel<<<dim3(batch, (2 * n + 31) / 32), 1024, 0, (cudaStream_t)stream>>>(partial, tiles_per_image, 2 * n, chstats, st_ld, st_off);
  return mudiff_launch_status();
}

extern "C" int mudiff_gn_apply(const void* x0, int c0, int ld0, const double* st0, int st0_ld,
                               const void* x1, int c1, int ld1, const double* st1, int st1_ld, int dtype_in,
                               const float* gamma, const float* beta, int64_t gb_bstride,
                               void* out, int ld_out, int dtype_out, int batch, int64_t hw, int groups,
                               float eps, int act, void* stream) {
  if (batch <= 0 || hw <= 0 || groups <= 0 || c0 <= 0 || c1 < 0 || !x0 || !st0 || !out) return MUDIFF_EINVAL;
  if (act != MUDIFF_ACT_NONE && act != MUDIFF_ACT_SILU) return MUDIFF_EINVAL;
  if (!x1) c1 = 0;
  if (c1 > 0 && !st1) return MUDIFF_EINVAL;
  if (batch > 65535) return MUDIFF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
#define AP(TI, TO) return launch_apply<TI, TO>(x0, c0, ld0, x1, c1, ld1, st0, st0_ld, st1, st1_ld, gamma, beta, gb_bstride, out, ld_out, batch, hw, groups, eps, act, st)
  if (dtype_in == MUDIFF_F32 && dtype_out == MUDIFF_F32) AP(float, float);
  if (dtype_in == MUDIFF_BF16 && dtype_out == MUDIFF_BF16) AP(__nv_bfloat16, __nv_bfloat16);
  if (dtype_in == MUDIFF_F32 && dtype_out == MUDIFF_BF16) AP(float, __nv_bfloat16);
  if (dtype_in == MUDIFF_BF16 && dtype_out == MUDIFF_F32) AP(__nv_bfloat16, float);
#undef AP
  return MUDIFF_EUNSUPPORTED;
}

// mudiff_gn_stats(x0 -> st0) + mudiff_gn_scale_shift([st0 | st1] -> table) in one launch (gn_stats_table_kernel); identical values.
extern "C" int mudiff_gn_stats_table(const void* x0, int c0, int ld0, int dtype, double* st0, int st0_ld,
                                     int c1, const double* st1, int st1_ld, const float* gamma, const float* beta,
                                     int64_t gb_bstride, int batch, int64_t hw, int groups, float eps, float* table, void* stream) {
  if (!x0 || !st0 || !table || batch <= 0 || hw <= 0 || groups <= 0 || c0 <= 0 || c1 < 0 || st0_ld < c0) return MUDIFF_EINVAL;
  if (!st1) c1 = 0;
  const int C = c0 + c1;
  if (C > GN_MAX_C || C % groups || groups > GN_MAX_C / 4 || batch > 65535) return MUDIFF_EUNSUPPORTED;
  if (dtype != MUDIFF_BF16 && dtype != MUDIFF_F32) return MUDIFF_EUNSUPPORTED;
  const int V = dtype == MUDIFF_BF16 ? 8 : 4;
  if (c0 % V || ld0 % V || ((uintptr_t)x0 % 16)) return MUDIFF_EUNSUPPORTED;
  const int cv = c0 / V;
  int block = (256 / cv) * cv;
  if (block < cv) block = cv;
  const int lanes = block / cv;
  const int ppb = gn_ppb(hw, c0);
  const int chunks = (int)((hw + ppb - 1) / ppb);
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  int rc = ensure_scratch(dev, (size_t)batch * chunks * c0 * 2, batch, st);
  if (rc) return rc;
  const size_t smem = gn_stats_smem(lanes, c0, c0, block, true);
  const unsigned int tot_off = (unsigned int)gn_stats_smem(lanes, c0, c0, block);
  if (smem > 48 * 1024) return MUDIFF_EUNSUPPORTED;
  if (dtype == MUDIFF_BF16)
    gn_stats_table_kernel<__nv_bfloat16><<<dim3(chunks, batch), block, smem, st>>>(
        (const __nv_bfloat16*)x0, c0, ld0, st0, st0_ld, c1, st1, st1_ld, gamma, beta, gb_bstride, hw, groups, eps, table,
        g_scratch[dev].partial, g_scratch[dev].tickets, ppb, tot_off);
  else
    gn_stats_table_kernel<float><<<dim3(chunks, batch), block, smem, st>>>(
        (const float*)x0, c0, ld0, st0, st0_ld, c1, st1, st1_ld, gamma, beta, gb_bstride, hw, groups, eps, table,
        g_scratch[dev].partial, g_scratch[dev].tickets, ppb, tot_off);
  return mudiff_launch_status();
}

// Statistics of x0 + GroupNorm/AdaGN (+SiLU) of [x0 | x1] in ONE launch, the re-read of x0 served by L2 (gn_l2_kernel).
// x0's per-channel (sum, sumsq) go to st0 (same values as mudiff_gn_stats); x1 (optional) comes with known statistics st1.
// bf16 in/out.  MUDIFF_EUNSUPPORTED when the shape does not qualify (callers then run mudiff_gn_stats + mudiff_gn_apply,
// which give bit-identical results).
extern "C" int mudiff_gn_stats_apply(const void* x0, int c0, int ld0, double* st0, int st0_ld,
                                     const void* x1, int c1, int ld1, const double* st1, int st1_ld, int dtype,
                                     const float* gamma, const float* beta, int64_t gb_bstride,
                                     void* out, int ld_out, int batch, int64_t hw, int groups, float eps, int act, void* stream) {
  if (batch <= 0 || hw <= 0 || groups <= 0 || c0 <= 0 || c1 < 0 || !x0 || !st0 || !out || st0_ld < c0) return MUDIFF_EINVAL;
  if (act != MUDIFF_ACT_NONE && act != MUDIFF_ACT_SILU) return MUDIFF_EINVAL;
  if (!x1) c1 = 0;
  if (c1 > 0 && !st1) return MUDIFF_EINVAL;
  if (dtype != MUDIFF_BF16 || batch > 65535) return MUDIFF_EUNSUPPORTED;
  constexpr int V = 8;
  const int C = c0 + c1;
  if (C % V || c0 % V || ld0 % V || (x1 && ld1 % V) || ld_out % V || C > GN_MAX_C || C % groups || groups > GN_MAX_C / 4)
    return MUDIFF_EUNSUPPORTED;
  if (((uintptr_t)x0 % 16) || (x1 && ((uintptr_t)x1 % 16)) || ((uintptr_t)out % 16)) return MUDIFF_EUNSUPPORTED;
  const int cv0 = c0 / V, cv = C / V;
  const int block = (256 / cv0) * cv0;            // phase 1 geometry == launch_stats(x0 alone): bit-identical statistics
  if (block < cv0 || block / cv < 1 || (block / cv) * cv * 2 < block) return MUDIFF_EUNSUPPORTED;
  const int lanes = block / cv0;
  const int ppb = gn_ppb(hw, c0);
  const int chunks = (int)((hw + ppb - 1) / ppb);
  if (chunks > kGnL2MaxChunks) return MUDIFF_EUNSUPPORTED;   // all blocks of an image must be co-resident (they wait for each other)
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  int rc = ensure_scratch(dev, (size_t)batch * chunks * c0 * 2, 34 * batch, st);
  if (rc) return rc;
  unsigned int* ctr = g_scratch[dev].tickets;
  static int poll_ns = -1;
  if (poll_ns < 0) { const char* e = getenv("MUDIFF_GN_POLL_NS"); poll_ns = e ? atoi(e) : 1000; }
  const size_t smem = gn_stats_smem(lanes, c0, c0, block);
  gn_l2_kernel<__nv_bfloat16><<<dim3(chunks, batch), block, smem, st>>>(
      (const __nv_bfloat16*)x0, c0, ld0, (const __nv_bfloat16*)x1, c1, ld1, st0, st0_ld, st1, st1_ld, c0, gamma, beta, gb_bstride,
      (__nv_bfloat16*)out, ld_out, hw, groups, eps, act, g_scratch[dev].partial, ctr, ctr + 2 * batch, ctr + batch, ppb, (unsigned int)poll_ns);
  return mudiff_launch_status();
}

extern "C" int mudiff_gn_scale_shift(const double* st0, int st0_ld, int c0, const double* st1, int st1_ld, int c1,
                                     const float* gamma, const float* beta, int64_t gb_bstride, int batch, int64_t hw,
                                     int groups, float eps, float* table, void* stream) {
  if (!st0 || !table || batch <= 0 || hw <= 0 || groups <= 0 || c0 <= 0 || c1 < 0) return MUDIFF_EINVAL;
  if (!st1) c1 = 0;
  const int C = c0 + c1;
  if (C > GN_MAX_C || C % groups || groups > GN_MAX_C / 4) return MUDIFF_EUNSUPPORTED;
  gn_scale_shift_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(st0, st0_ld, c0, st1, st1_ld, c1, gamma, beta, gb_bstride,
                                                                   (double)hw, groups, eps, table);
  return mudiff_launch_status();
}

// Single-pass GroupNorm (+AdaGN scale/shift, +SiLU): out = act(GN(x) * gamma + beta) and chstats[b][st_off + c] = per-channel
// (sum, sumsq) of x, with ONE read of x (gn_fused_kernel).  bf16 in/out, dense NHWC (pixel stride == c), c % 8 == 0.
// Returns MUDIFF_EUNSUPPORTED when an image chunk does not fit the shared-memory stages (callers use gn_stats + gn_apply).
struct GnFusedScratch { double* partial = nullptr; size_t pcap = 0; unsigned int* ctr = nullptr; int ccap = 0; };
static GnFusedScratch g_gnf[16];

extern "C" int mudiff_gn_fused(const void* x, void* out, int c, int batch, int64_t hw, int groups, const float* gamma,
                               const float* beta, int64_t gb_bstride, float eps, int act, double* chstats, int st_ld,
                               int st_off, void* stream) {
  if (!x || !out || !chstats || batch <= 0 || hw <= 0 || groups <= 0 || c <= 0 || st_ld < c + st_off) return MUDIFF_EINVAL;
  if (act != MUDIFF_ACT_NONE && act != MUDIFF_ACT_SILU) return MUDIFF_EINVAL;
  if (c % 8 || c > GN_MAX_C || c % groups || groups > GN_MAX_C / 4 || c / 8 > 256) return MUDIFF_EUNSUPPORTED;
  if (((uintptr_t)x % 16) || ((uintptr_t)out % 16)) return MUDIFF_EUNSUPPORTED;
  const int cv = c / 8;
  const int block = (256 / cv) * cv;
  const int lanes = block / cv;
  const size_t fixed = 64 + ((size_t)lanes * c * 2 * sizeof(float) + 127) / 128 * 128;
  // two CTAs per SM when the stages fit (twice the latency hiding: an iteration is a chain of dependent global round
  // trips - ticket, flag, statistics, gamma/beta), else one; three stages when they fit, else two
  int G = 0, D = 0, ppc = 0;
  size_t stage_bytes = 0;
  for (int R = 1; R >= 1 && !G; --R) {     // R = 2 (two CTAs per SM) measured slower: more partials, same dependent chain
    const size_t budget = R == 2 ? 108 * 1024 : 200 * 1024;
    const int g = MUDIFF_NUM_SMS * R;
    const int pp = (int)((hw + g - 1) / g);
    const size_t sb = ((size_t)pp * c * 2 + 127) / 128 * 128;
    for (int d = 3; d >= 2; --d)
      if (fixed + (size_t)d * sb <= budget) { G = g; D = d; ppc = pp; stage_bytes = sb; break; }
  }
  if (!G) return MUDIFF_EUNSUPPORTED;
  if (batch < D) D = batch < 1 ? 1 : batch;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  GnFusedScratch& sc = g_gnf[dev];
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  const size_t need = (size_t)batch * G * c * 2;      // G <= 2 * MUDIFF_NUM_SMS
  if (need > sc.pcap) {
    if (cs != cudaStreamCaptureStatusNone) return MUDIFF_EUNSUPPORTED;     // must be sized by a warm-up run
    double* pnew = nullptr;
    if (cudaMalloc(&pnew, need * 2 * sizeof(double)) != cudaSuccess) return (int)cudaGetLastError();
    sc.partial = pnew; sc.pcap = need * 2;
  }
  if (2 * batch + 1 > sc.ccap) {
    if (cs != cudaStreamCaptureStatusNone) return MUDIFF_EUNSUPPORTED;
    const int cap = 2 * batch + 1 < 4096 ? 4096 : 2 * (2 * batch + 1);
    unsigned int* t = nullptr;
    if (cudaMalloc(&t, cap * sizeof(unsigned int)) != cudaSuccess) return (int)cudaGetLastError();
    cudaMemset(t, 0, cap * sizeof(unsigned int));
    sc.ctr = t; sc.ccap = cap;
  }
  static bool attr_set[16] = {};
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) != cudaSuccess)
      return (int)cudaGetLastError();
    attr_set[dev] = true;
  }
  GnFusedP p;
  p.x = (const __nv_bfloat16*)x; p.out = (__nv_bfloat16*)out;
  p.C = c; p.cv = cv; p.groups = groups; p.batch = batch; p.D = D; p.hw = hw; p.ppc = ppc;
  p.gamma = gamma; p.beta = beta; p.gb_bstride = gb_bstride; p.eps = eps; p.act = act;
  p.chstats = chstats; p.st_ld = st_ld; p.st_off = st_off;
  p.partial = sc.partial;
  const int half = sc.ccap / 2;
  p.tickets = sc.ctr; p.flags = sc.ctr + half; p.done = sc.ctr + sc.ccap - 1;
  p.stage_bytes = (uint32_t)stage_bytes;
  const size_t smem = fixed + (size_t)D * stage_bytes;
  void* args[] = {(void*)&p};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)gn_fused_kernel, dim3(G), dim3(block), args, smem, st);
  ++g_mudiff_launches;
  if (e == cudaErrorCooperativeLaunchTooLarge) { cudaGetLastError(); return MUDIFF_EUNSUPPORTED; }
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  return 0;
}
