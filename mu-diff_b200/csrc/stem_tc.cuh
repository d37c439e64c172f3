// Stem convolution conv3x3(1 -> N) + folded GroupNorm/AdaGN scale-shift + activation on the tensor cores.
// Included by conv_tc.cu (same PTX wrappers / bounded waits / debug channel).
//
// The CUDA-core kernel (conv_stem_gn_kernel) needs ~11 issue slots per output element and sits at 28 % of the HBM
// write roofline.  Here the 9-tap contraction is ONE K = 32 UMMA pair per 128 pixels:
//   A[128 px][64]  bf16, built by four warps from the fp32 image (im2col in shared memory, 128-byte swizzle):
//                  k 0..8 = hi(x_t), 9..17 = lo(x_t), 18..26 = hi(x_t), rest 0      (x = hi + lo, both bf16)
//   B[N][64]       bf16, stationary: k 0..8 = hi(w_t), 9..17 = hi(w_t), 18..26 = lo(w_t), rest 0
//   D = x_hi w_hi + x_lo w_hi + x_hi w_lo  (fp32 accumulate in TMEM; the dropped lo*lo term is ~2^-18 relative)
// and the epilogue (tcgen05.ld -> acc * scale[b][c] + shift[b][c] -> SiLU -> bf16/fp32 16-byte stores) is all that is
// left on the CUDA cores: ~3 issue slots per output element.
//   warps 0-3  epilogue (one per TMEM lane quadrant)      warps 4-7  A builders      warp 8  TMEM alloc + MMA issuer

struct StemTcP {
  const float* x; int H, W, batch;
  const float* wt; const float* bias;            // [N][9], [N]
  const float* scale_shift;                      // [B][N][2] or NULL (plain conv + bias)
  int N, act;
  void* out; int out_ld, out_coff;
  int tiles_per_image; long long total_units;
  uint32_t idesc;
};

constexpr int kStemStages = 4;
constexpr uint32_t kStemA = 0;                                   // 4 x 16 KB
constexpr uint32_t kStemB = kStemStages * 16384;                 // N x 128 B (<= 32 KB)
constexpr uint32_t kStemTab = kStemB + 32768;                    // [N][2] floats (<= 2 KB)
constexpr uint32_t kStemBar = kStemTab + 2048;
constexpr uint32_t kStemSmem = kStemBar + 1024 + 1024;
enum { SB_AFULL = 0, SB_AEMPTY = 4, SB_TFULL = 8, SB_TEMPTY = 10, SB_COUNT = 12 };

__device__ __forceinline__ uint32_t bf16_hi_lo_pack(float v, float& lo_out) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  lo_out = v - __bfloat162float(h);
  return (uint32_t)__bfloat16_as_ushort(h);
}

template <bool kOutF32>
__global__ void __launch_bounds__(288, 2) stem_tc_kernel(const __grid_constant__ StemTcP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint8_t* bar_block = smem + kStemBar;
  uint64_t* bar = (uint64_t*)(smem + kStemBar);
  uint32_t* tmem_slot = (uint32_t*)(bar + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N;
  const int tmem_cols = N <= 64 ? 128 : (N <= 128 ? 256 : 512);       // two accumulator stages
  const int acc_stride = tmem_cols / 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStemStages; ++i) { mbar_init(&bar[SB_AFULL + i], 4); mbar_init(&bar[SB_AEMPTY + i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar[SB_TFULL + i], 1); mbar_init(&bar[SB_TEMPTY + i], 4); }
    for (int i = 0; i < 48; ++i) ((int*)(bar_block + kDbgRecOff))[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // stationary weights: row n = 128 bytes, chunk c of the row at (c ^ (n & 7)); 64 bf16 per row
  for (int idx = threadIdx.x; idx < N * 8; idx += blockDim.x) {
    const int n = idx >> 3, c = idx & 7;
    uint32_t w32[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t pk = 0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = c * 8 + j * 2 + e;
        float v = 0.f;
        if (k < 27) {
          const float w = p.wt[n * 9 + k % 9];
          float lo;
          const uint32_t hi = bf16_hi_lo_pack(w, lo);
          v = k < 18 ? __uint_as_float(hi << 16) : lo;
        }
        pk |= (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)) << (16 * e);
      }
      w32[j] = pk;
    }
    *reinterpret_cast<uint4*>(smem + kStemB + n * 128 + ((c ^ (n & 7)) << 4)) = make_uint4(w32[0], w32[1], w32[2], w32[3]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long u_begin = (p.total_units * (long long)blockIdx.x) / gridDim.x;
  const long long u_end = (p.total_units * (long long)(blockIdx.x + 1)) / gridDim.x;
  const long long hw = (long long)p.H * p.W;

  if (warp >= 4 && warp < 8) {
    // ===================== A builders: one pixel row of the tile per thread =====================
    const int r = (warp - 4) * 32 + lane;
    uint32_t it = 0;
    for (long long u = u_begin; u < u_end; ++u, ++it) {
      const int b = (int)(u / p.tiles_per_image);
      const long long pix = (u % p.tiles_per_image) * 128 + r;
      const uint32_t s = it % kStemStages;
      mbar_wait(&bar[SB_AEMPTY + s], ((it / kStemStages) & 1u) ^ 1u, bar_block, (int)it);
      float v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = 0.f;
      if (pix < hw) {
        const int y = (int)(pix / p.W), x = (int)(pix - (long long)y * p.W);
        const float* img = p.x + (long long)b * hw;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
          if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) v[t] = __ldg(img + (long long)yy * p.W + xx);
        }
      }
      uint16_t e[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) e[k] = 0;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float lo;
        const uint16_t hi = (uint16_t)bf16_hi_lo_pack(v[t], lo);
        e[t] = hi; e[18 + t] = hi;
        e[9 + t] = __bfloat16_as_ushort(__float2bfloat16_rn(lo));
      }
      uint8_t* rowp = smem + kStemA + s * 16384 + r * 128;
      const int rs = r & 7;
#pragma unroll
      for (int c = 0; c < 4; ++c) {               // k 0..31 only: the two UMMAs never read chunks 4..7 of the row
        uint4 q;
        q.x = e[c * 8 + 0] | ((uint32_t)e[c * 8 + 1] << 16); q.y = e[c * 8 + 2] | ((uint32_t)e[c * 8 + 3] << 16);
        q.z = e[c * 8 + 4] | ((uint32_t)e[c * 8 + 5] << 16); q.w = e[c * 8 + 6] | ((uint32_t)e[c * 8 + 7] << 16);
        *reinterpret_cast<uint4*>(rowp + ((c ^ rs) << 4)) = q;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[SB_AFULL + s]);
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t sb = smem_u32(smem);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint64_t bd = ((uint64_t)hi << 32) | ((((sb + kStemB) & 0x3FFFFu) >> 4) | (1u << 16));
    uint32_t it = 0;
    for (long long u = u_begin; u < u_end; ++u, ++it) {
      const uint32_t s = it % kStemStages, acc = it & 1u;
      mbar_wait(&bar[SB_TEMPTY + acc], ((it >> 1) & 1u) ^ 1u, bar_block, (int)it);
      mbar_wait(&bar[SB_AFULL + s], (it / kStemStages) & 1u, bar_block, (int)it);
      tc_fence_after();
      if (leader) {
        const uint64_t ad = ((uint64_t)hi << 32) | ((((sb + kStemA + s * 16384) & 0x3FFFFu) >> 4) | (1u << 16));
        tc_mma_f16(tmem_base + acc * acc_stride, ad, bd, p.idesc, 0u);
        tc_mma_f16(tmem_base + acc * acc_stride, ad + 2, bd + 2, p.idesc, 1u);       // k 16..31 (only 27 slots are non-zero)
        tc_commit(&bar[SB_AEMPTY + s]);
        tc_commit(&bar[SB_TFULL + acc]);
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    // ===================== epilogue =====================
    const int q = warp;
    const int row = q * 32 + lane;
    float* tab = (float*)(smem + kStemTab);        // [N][2]: scale, shift (bias folded)
    int tab_b = -1;
    uint32_t it = 0;
    for (long long u = u_begin; u < u_end; ++u, ++it) {
      const int b = (int)(u / p.tiles_per_image);
      const long long pix = (u % p.tiles_per_image) * 128 + row;
      if (b != tab_b) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int c = threadIdx.x; c < N; c += 128) {
          float sc = 1.f, sh = 0.f;
          if (p.scale_shift) { sc = p.scale_shift[((long long)b * N + c) * 2]; sh = p.scale_shift[((long long)b * N + c) * 2 + 1]; }
          tab[2 * c] = sc;
          tab[2 * c + 1] = (p.bias ? p.bias[c] : 0.f) * sc + sh;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        tab_b = b;
      }
      const uint32_t acc = it & 1u;
      mbar_wait(&bar[SB_TFULL + acc], (it >> 1) & 1u, bar_block, (int)it);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_stride;
      for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr0 + (uint32_t)c, v);
        tmem_ld_wait();
        if (pix < hw) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 ss = *reinterpret_cast<const float4*>(tab + 2 * (c + 2 * j));      // sc0 sh0 sc1 sh1 (broadcast)
            f[2 * j] = fmaf(__uint_as_float(v[2 * j]), ss.x, ss.y);
            f[2 * j + 1] = fmaf(__uint_as_float(v[2 * j + 1]), ss.z, ss.w);
          }
          if (p.act == MUDIFF_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = kOutF32 ? silu_exact(f[j]) : silu_f(f[j]);
          }
          const long long o = ((long long)b * hw + pix) * p.out_ld + p.out_coff + c;
          if (kOutF32) {
            float4* op = reinterpret_cast<float4*>((float*)p.out + o);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>((__nv_bfloat16*)p.out + o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 raw;
              __nv_bfloat162* e2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
              for (int i = 0; i < 4; ++i) e2[i] = __floats2bfloat162_rn(f[8 * j + 2 * i], f[8 * j + 2 * i + 1]);
              op[j] = raw;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[SB_TEMPTY + acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}
