// Volume-prediction front / back end on the GPU (engine/test_volume.py:135-191, 269-294 is the
// numpy reference): robust [pmin, pmax] percentile window over the non-zero voxels -> [-1, 1], centre-slice extraction
// ([H, W, Z] -> [n, 1, S, S], bilinear resize when S != H, W), and the (x + 1) / 2 clamp + re-stack of the predictions.
//
// Percentiles are EXACT order statistics (numpy 'linear' method) found by a two-level radix select on the
// order-preserving 32-bit key of the fp32 voxels: a 2^16-bin histogram of the high halves, then - for each of the six
// wanted ranks - a 2^16-bin histogram of the low halves inside the selected bin.  Two reads of the volume, no sort.
#include "common.cuh"

namespace {

constexpr int kBins = 65536;
constexpr int kTargets = 6;        // ranks: floor/ceil of the pmin and pmax virtual indices, minimum, maximum

__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}

struct SelState {
  unsigned long long m;            // number of valid (non-zero) voxels
  unsigned long long rank[kTargets];
  unsigned long long below[kTargets];   // valid voxels in bins below the selected one
  uint32_t bin[kTargets];          // selected high half
  uint32_t key[kTargets];          // final 32-bit keys
  double gamma_lo, gamma_hi;       // interpolation fractions of the two percentiles
  float lo, hi;                    // result window
  int status;                      // 0 ok, 1 degenerate (all zero output)
};

__global__ void hist_hi_kernel(const float* __restrict__ v, long long n, unsigned int* __restrict__ hist) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float f = v[i];
    if (f != 0.0f) atomicAdd(&hist[f2key(f) >> 16], 1u);       // reference mask: data != 0 (test_volume.py:143)
  }
}

// one block: total count, the six ranks (numpy 'linear': virtual index (m-1) * q / 100), and for each rank the bin of
// the histogram that holds it
__global__ void select_hi_kernel(const unsigned int* __restrict__ hist, SelState* st, float pmin, float pmax) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x;
  unsigned long long s = 0;
  for (int i = 0; i < kBins / 1024; ++i) s += hist[t * (kBins / 1024) + i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    unsigned long long tot = 0;
    for (int i = 0; i < 1024; ++i) { const unsigned long long c = part[i]; part[i] = tot; tot += c; }   // exclusive scan
    st->m = tot;
    st->status = tot == 0 ? 1 : 0;
    if (tot) {
      // numpy >= 2 evaluates the quantile in the ARRAY dtype (fp32 here): q = pmin / float32(100), virtual index
      // = (m - 1) * q in fp32 (np.lib._function_base_impl._quantile), indices >= m - 1 select the last element
      const float nm1 = __ull2float_rn(tot - 1);
      const float vlo = __fmul_rn(nm1, __fdiv_rn(pmin, 100.0f)), vhi = __fmul_rn(nm1, __fdiv_rn(pmax, 100.0f));
      unsigned long long klo = (unsigned long long)floorf(vlo), khi = (unsigned long long)floorf(vhi);
      st->gamma_lo = (double)(vlo - floorf(vlo)); st->gamma_hi = (double)(vhi - floorf(vhi));
      unsigned long long klo1 = klo + 1, khi1 = khi + 1;
      if (vlo >= nm1 || klo1 > tot - 1) { klo = tot - 1; klo1 = tot - 1; }
      if (vhi >= nm1 || khi1 > tot - 1) { khi = tot - 1; khi1 = tot - 1; }
      st->rank[0] = klo; st->rank[1] = klo1;
      st->rank[2] = khi; st->rank[3] = khi1;
      st->rank[4] = 0; st->rank[5] = tot - 1;
    }
  }
  __syncthreads();
  if (st->status) return;
  // thread t owns bins [t*64, t*64+64): find the ranks that fall inside
  const unsigned long long base = part[t];
  unsigned long long run = base;
  for (int i = 0; i < kBins / 1024; ++i) {
    const unsigned long long c = hist[t * (kBins / 1024) + i];
    for (int k = 0; k < kTargets; ++k) {
      const unsigned long long r = st->rank[k];
      if (r >= run && r < run + c) { st->bin[k] = (uint32_t)(t * (kBins / 1024) + i); st->below[k] = run; }
    }
    run += c;
  }
}

__global__ void hist_lo_kernel(const float* __restrict__ v, long long n, const SelState* __restrict__ st,
                               unsigned int* __restrict__ hist2) {
  if (st->status) return;
  uint32_t bins[kTargets];
#pragma unroll
  for (int k = 0; k < kTargets; ++k) bins[k] = st->bin[k];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float f = v[i];
    if (f == 0.0f) continue;
    const uint32_t key = f2key(f), hi = key >> 16;
#pragma unroll
    for (int k = 0; k < kTargets; ++k)
      if (hi == bins[k]) atomicAdd(&hist2[(size_t)k * kBins + (key & 0xFFFFu)], 1u);
  }
}

// grid = kTargets blocks: exact key of each rank; block 0 then folds the window (numpy _lerp in fp32)
__global__ void select_lo_kernel(const unsigned int* __restrict__ hist2, SelState* st, unsigned int* done) {
  __shared__ unsigned long long part[1024];
  __shared__ bool s_last;
  if (st->status) return;
  const int k = blockIdx.x, t = threadIdx.x;
  const unsigned int* h = hist2 + (size_t)k * kBins;
  unsigned long long s = 0;
  for (int i = 0; i < kBins / 1024; ++i) s += h[t * (kBins / 1024) + i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    unsigned long long tot = st->below[k];
    for (int i = 0; i < 1024; ++i) { const unsigned long long c = part[i]; part[i] = tot; tot += c; }
  }
  __syncthreads();
  unsigned long long run = part[t];
  const unsigned long long r = st->rank[k];
  for (int i = 0; i < kBins / 1024; ++i) {
    const unsigned long long c = h[t * (kBins / 1024) + i];
    if (r >= run && r < run + c) st->key[k] = (st->bin[k] << 16) | (uint32_t)(t * (kBins / 1024) + i);
    run += c;
  }
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const unsigned int d = atomicAdd(done, 1u);
    s_last = d == kTargets - 1;
    if (s_last) *done = 0;
  }
  __syncthreads();
  if (!s_last || t != 0) return;
  __threadfence();
  volatile SelState* vs = st;
  float a[kTargets];
  for (int i = 0; i < kTargets; ++i) a[i] = key2f(vs->key[i]);
  // numpy.percentile(..., method='linear') on a float32 array: gamma cast to float32, _lerp(a, b, t) =
  //   a + (b - a) * t, replaced by b - (b - a) * (1 - t) where t >= 0.5 (separately rounded fp32 operations)
  auto lerp = [](float lo, float hi, double g) {
    const float t = (float)g;
    const float d = __fsub_rn(hi, lo);
    float r = __fadd_rn(lo, __fmul_rn(d, t));
    if (t >= 0.5f) r = __fsub_rn(hi, __fmul_rn(d, __fsub_rn(1.0f, t)));
    if (d == 0.0f) r = lo;
    return r;
  };
  float lo = lerp(a[0], a[1], vs->gamma_lo), hi = lerp(a[2], a[3], vs->gamma_hi);
  int status = 0;
  if (!isfinite(lo) || !isfinite(hi) || !(hi > lo)) {          // test_volume.py:148-151: fall back to the full range
    lo = a[4]; hi = a[5];
    if (!(hi > lo)) status = 1;
  }
  st->lo = lo; st->hi = hi; st->status = status;
}

// out[i, 0, y, x] = bilinear(normalised slice s0 + i)(y, x);  normalise = clip((v - lo) / (hi - lo), 0, 1) * 2 - 1 in fp32.
// vol is [H, W, Z] (Z fastest, the numpy layout of nibabel volumes).  When S == H == W the sample is the voxel itself.
__global__ void slices_kernel(const float* __restrict__ vol, int H, int W, int Z, int s0, int n, int SH, int SW,
                              const SelState* __restrict__ st, float* __restrict__ out) {
  const long long total = (long long)n * SH * SW;
  const bool degenerate = st->status != 0;
  const float lo = st->lo, den = __fsub_rn(st->hi, st->lo);
  auto norm = [&](float v) {
    if (degenerate) return 0.0f;
    float x = __fdiv_rn(__fsub_rn(v, lo), den);
    x = fminf(fmaxf(x, 0.0f), 1.0f);
    return __fsub_rn(__fmul_rn(x, 2.0f), 1.0f);
  };
  const float sh = (float)H / (float)SH, sw = (float)W / (float)SW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % SW);
    const int y = (int)((idx / SW) % SH);
    const int i = (int)(idx / ((long long)SW * SH));
    const int z = s0 + i;
    auto at = [&](int yy, int xx) { return norm(vol[((long long)yy * W + xx) * Z + z]); };
    float r;
    if (SH == H && SW == W) {
      r = at(y, x);
    } else {
      // F.interpolate(mode='bilinear', align_corners=False) (test_volume.py:274-275): ATen area_pixel_compute_source_index
      float fy = __fsub_rn(__fmul_rn(sh, (float)y + 0.5f), 0.5f); if (fy < 0.f) fy = 0.f;
      float fx = __fsub_rn(__fmul_rn(sw, (float)x + 0.5f), 0.5f); if (fx < 0.f) fx = 0.f;
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      const float hy = 1.f - ly, hx = 1.f - lx;
      r = hy * (hx * at(y0, x0) + lx * at(y0, x1)) + ly * (hx * at(y1, x0) + lx * at(y1, x1));
    }
    out[idx] = r;
  }
}

// vol[y, x, s0 + i] = clamp((pred[i, 0, y, x] + 1) / 2, 0, 1) (test_volume.py:285), all other voxels zero
__global__ void restack_kernel(const float* __restrict__ pred, int H, int W, int Z, int s0, int n, int to01,
                               float* __restrict__ vol) {
  const long long total = (long long)H * W * Z;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(idx % Z);
    const long long yx = idx / Z;
    float r = 0.f;
    const int i = z - s0;
    if (i >= 0 && i < n) {
      r = pred[(long long)i * H * W + yx];
      if (to01) r = fminf(fmaxf(__fdiv_rn(__fadd_rn(r, 1.0f), 2.0f), 0.0f), 1.0f);
    }
    vol[idx] = r;
  }
}

}  // namespace

extern "C" int mudiff_volume_workspace_bytes(void) {
  return (int)(sizeof(unsigned int) * (size_t)kBins * (1 + kTargets) + 1024);
}

// Robust percentile window of a volume (fp32, n voxels): state->lo / hi as engine/test_volume.py:135-157 computes them
// (np.percentile 'linear' over the voxels != 0, fall back to min / max, degenerate -> all-zero output).  `workspace`
// = mudiff_volume_workspace_bytes() bytes; the window stays on the device inside it and is consumed by
// mudiff_volume_to_slices.  `window_out` (optional, device float[3]) receives lo, hi, status.
extern "C" int mudiff_volume_window(const float* vol, int64_t n, float pmin, float pmax, void* workspace, void* stream) {
  if (!vol || !workspace || n <= 0 || !(pmin >= 0.f && pmax <= 100.f && pmin <= pmax)) return MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* hist = (unsigned int*)workspace;
  unsigned int* hist2 = hist + kBins;
  SelState* state = (SelState*)(hist2 + (size_t)kBins * kTargets);
  unsigned int* done = (unsigned int*)((uint8_t*)state + 512);
  cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)mudiff_volume_workspace_bytes(), st);
  if (e != cudaSuccess) return (int)e;
  const int grid = grid_for(n, 256, MUDIFF_NUM_SMS * 8);
  hist_hi_kernel<<<grid, 256, 0, st>>>(vol, n, hist);
  select_hi_kernel<<<1, 1024, 0, st>>>(hist, state, pmin, pmax);
  hist_lo_kernel<<<grid, 256, 0, st>>>(vol, n, state, hist2);
  select_lo_kernel<<<kTargets, 1024, 0, st>>>(hist2, state, done);
  g_mudiff_launches += 3;
  return mudiff_launch_status();
}

// window[0..2] <- lo, hi, status of the last mudiff_volume_window on this workspace (device -> device copy, tests / logs)
extern "C" int mudiff_volume_window_read(const void* workspace, float* window, void* stream) {
  if (!workspace || !window) return MUDIFF_EINVAL;
  const uint8_t* state = (const uint8_t*)workspace + sizeof(unsigned int) * (size_t)kBins * (1 + kTargets);
  const SelState* s = (const SelState*)state;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(window, &s->lo, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(window + 2, &s->status, sizeof(int), cudaMemcpyDeviceToDevice, st);
  ++g_mudiff_launches;
  return (int)e;
}

// out [n, 1, sh, sw] <- normalised (window in `workspace`) axial slices s0 .. s0+n-1 of vol [H, W, Z], resized
// bilinearly (align_corners = False) when (sh, sw) != (H, W)
extern "C" int mudiff_volume_to_slices(const float* vol, int h, int w, int z, int s0, int n, int sh, int sw,
                                       const void* workspace, float* out, void* stream) {
  if (!vol || !workspace || !out || h <= 0 || w <= 0 || z <= 0 || n <= 0 || s0 < 0 || s0 + n > z || sh <= 0 || sw <= 0)
    return MUDIFF_EINVAL;
  const SelState* state = (const SelState*)((const uint8_t*)workspace + sizeof(unsigned int) * (size_t)kBins * (1 + kTargets));
  const long long total = (long long)n * sh * sw;
  slices_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(vol, h, w, z, s0, n, sh, sw, state, out);
  return mudiff_launch_status();
}

// vol [H, W, Z] <- zeros, with slices s0 .. s0+n-1 = pred [n, 1, H, W] (mapped (x+1)/2 and clamped to [0, 1] if to01)
extern "C" int mudiff_slices_to_volume(const float* pred, int h, int w, int z, int s0, int n, int to01, float* vol,
                                       void* stream) {
  if (!pred || !vol || h <= 0 || w <= 0 || z <= 0 || n <= 0 || s0 < 0 || s0 + n > z) return MUDIFF_EINVAL;
  const long long total = (long long)h * w * z;
  restack_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(pred, h, w, z, s0, n, to01, vol);
  return mudiff_launch_status();
}

// ---------------------------------------------------------------------------------
// Slice-test driver helpers (engine/test.py:265-400, dataset/dataset_brats.py:73-92):
//   z-score slices -> [-1, 1]   (clamp(x, -3, 3) / 3),
//   global min / max over predictions and ground truth, and the [0, 255] uint8 export with that global window.
// ---------------------------------------------------------------------------------
namespace {

__global__ void zscore_unit_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn(fminf(fmaxf(in[i], -3.0f), 3.0f), 3.0f);
}

__global__ void minmax_init_kernel(unsigned int* keys) { keys[0] = 0xFFFFFFFFu; keys[1] = 0u; }
__global__ void minmax_read_kernel(const unsigned int* keys, float* w) { w[0] = key2f(keys[0]); w[1] = key2f(keys[1]); }

// exact and order-independent: min / max on the order-preserving integer keys
__global__ void minmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ keys) {
  unsigned int lo = 0xFFFFFFFFu, hi = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int k = f2key(x[i]);
    lo = min(lo, k); hi = max(hi, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&keys[0], lo); atomicMax(&keys[1], hi); }
}

// engine/test.py:378-388: np.clip((x - gmin) / (gmax - gmin) * 255.0, 0, 255).astype(np.uint8) with python-float gmin /
// gmax (fp64 difference, then fp32 array arithmetic); constant images fall back to the window [0, 1] (:373-374)
__global__ void to_u8_kernel(const float* __restrict__ x, long long n, const unsigned int* __restrict__ keys,
                             unsigned char* __restrict__ out) {
  float gmin = key2f(keys[0]), gmax = key2f(keys[1]);
  if (!(gmax > gmin)) { gmin = 0.0f; gmax = 1.0f; }
  const float den = (float)((double)gmax - (double)gmin);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = __fmul_rn(__fdiv_rn(__fsub_rn(x[i], gmin), den), 255.0f);
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    out[i] = (unsigned char)v;                   // truncation, like ndarray.astype(np.uint8) on [0, 255]
  }
}

}  // namespace

// out = clamp(in, -3, 3) / 3  (dataset/dataset_brats.py:83,91), fp32, n elements; in == out allowed
extern "C" int mudiff_zscore_to_unit(const float* in, float* out, int64_t n, void* stream) {
  if (!in || !out || n < 0) return MUDIFF_EINVAL;
  if (n == 0) return 0;
  zscore_unit_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  return mudiff_launch_status();
}

// keys[0..1] (device uint32[2]) <- order-preserving keys of min / max over x (n > 0).  `accumulate` != 0 folds x into
// the keys already there (global window over several tensors: predictions and ground truth, engine/test.py:368-371).
extern "C" int mudiff_minmax_keys(const float* x, int64_t n, int accumulate, unsigned int* keys, void* stream) {
  if (!x || !keys || n <= 0) return MUDIFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) { minmax_init_kernel<<<1, 1, 0, st>>>(keys); ++g_mudiff_launches; }
  minmax_kernel<<<grid_for(n, 256, MUDIFF_NUM_SMS * 8), 256, 0, st>>>(x, n, keys);
  return mudiff_launch_status();
}

// window[0..1] (device float[2]) <- min, max decoded from the keys
extern "C" int mudiff_minmax_read(const unsigned int* keys, float* window, void* stream) {
  if (!keys || !window) return MUDIFF_EINVAL;
  minmax_read_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(keys, window);
  return mudiff_launch_status();
}

// out (uint8, n) <- the reference's global-window 8-bit export of x with the window in `keys`
extern "C" int mudiff_scale_to_u8(const float* x, int64_t n, const unsigned int* keys, unsigned char* out, void* stream) {
  if (!x || !keys || !out || n < 0) return MUDIFF_EINVAL;
  if (n == 0) return 0;
  to_u8_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, keys, out);
  return mudiff_launch_status();
}
