// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a  (bf16 x bf16 -> fp32).
//
//   D[128 pixels, n_tile] += A[128 pixels, 64 ch] * W[n_tile, 64 ch]^T        per K block
//
// Persistent, warp-specialised CTA (192 threads, one per SM):
//   warp 0      TMA producer: NHWC activation boxes (4-D tensor map; conv padding = TMA
//               out-of-bounds zero fill) + K-major weight boxes (3-D map) into a ring of
//               128B-swizzled shared-memory stages, mbarrier complete_tx signalling
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x n_tile x 16),
//               tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns) -> bias / temb row-bias /
//               alpha / residual / activation -> bf16|fp32 NHWC stores; double-buffered
//               TMEM accumulators let tile i+1's MMAs overlap tile i's epilogue
//
// Two A-staging modes per 3x3 segment:
//   per-tap : one {64ch, tw, th} box per (tap, channel block); 9x re-read of the input from L2
//   halo    : one {64ch, tw+2, th+2} box per channel block, the nine taps are nine UMMA
//             descriptors into the SAME staged tile (start address shifted by whole 128-byte
//             pixel rows, stride-byte-offset = halo row pitch).  Needs tw == 8 so that one
//             8-row swizzle atom is one tile row.  Cuts L2->smem traffic for A by 6.25x.
#include <cuda.h>
#include "common.cuh"

namespace {

constexpr int kMaxStages = 8;
constexpr int kThreads = 192;
constexpr uint32_t kSmemBudget = 225 * 1024;

struct TcParams {
  int batch, H, W;
  int tile_h, tile_w, tiles_x, tiles_y, tiles_per_img;
  int n_tiles, n_tile, total_tiles;
  int nseg;
  int seg_cblk[3], seg_taps[3], seg_halo[3], seg_koff[3], seg_c[3];
  int a_batched, w_batched;
  uint32_t stage_bytes, a_region_bytes, b_sub_bytes;
  int num_stages;
  uint32_t idesc;
  int acc_stride;       // TMEM columns between the two accumulator stages
  int tmem_cols;
  int base_off_variant; // debug: fill descriptor base-offset field from the address
  // epilogue
  const float* bias; const float* rowbias; int rowbias_ld;
  const void* residual; int res_ld;
  float alpha, beta; int act;
  void* out; int out_ld, out_coff;
};

// ---------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled UMMA shared-memory descriptor.
//   rows are 128 B (64 bf16), 8 rows = one swizzle atom, atoms `sbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes, int base_off_variant) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);                 // [0,14)  start address / 16
  d |= (uint64_t)1 << 16;                                    // [16,30) LBO (ignored for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;        // [32,46) SBO / 16
  d |= (uint64_t)1 << 46;                                    // [46,48) descriptor version (sm_100)
  if (base_off_variant) d |= (uint64_t)((saddr >> 7) & 7u) << 49;   // [49,52) matrix base offset
  d |= (uint64_t)2 << 61;                                    // [61,64) SWIZZLE_128B
  return d;
}

struct TileCoord { int b, y0, x0, n0; };
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile) {
  TileCoord t;
  int nt = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  t.b = mt / p.tiles_per_img;
  int r = mt - t.b * p.tiles_per_img;
  int ty = r / p.tiles_x;
  t.y0 = ty * p.tile_h;
  t.x0 = (r - ty * p.tiles_x) * p.tile_w;
  t.n0 = nt * p.n_tile;
  return t;
}

template <bool kOutF32>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW,
               const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.num_stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.num_stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    if (p.nseg > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA1) : "memory");
    if (p.nseg > 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA2) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int ab = p.a_batched ? t.b : 0;
        const int wb = p.w_batched ? t.b : 0;
        for (int s = 0; s < p.nseg; ++s) {
          const CUtensorMap* mapA = s == 0 ? &mapA0 : (s == 1 ? &mapA1 : &mapA2);
          const int taps = p.seg_taps[s];
          const bool halo = p.seg_halo[s] != 0;
          const int items = halo ? p.seg_cblk[s] : taps * p.seg_cblk[s];
          for (int it = 0; it < items; ++it) {
            int tap, cb;
            if (halo) { tap = 0; cb = it; } else { tap = it / p.seg_cblk[s]; cb = it - tap * p.seg_cblk[s]; }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
            uint8_t* sb = sa + p.a_region_bytes;
            if (halo) {
              const uint32_t a_bytes = (uint32_t)(p.tile_w + 2) * (p.tile_h + 2) * 128u;
              mbar_expect_tx(&full_bar[stage], a_bytes + 9u * p.b_sub_bytes);
              tma_load_4d(sa, mapA, &full_bar[stage], cb * 64, t.x0 - 1, t.y0 - 1, ab);
              for (int j = 0; j < 9; ++j)
                tma_load_3d(sb + (size_t)j * p.b_sub_bytes, &mapW, &full_bar[stage],
                            p.seg_koff[s] + j * p.seg_c[s] + cb * 64, t.n0, wb);
            } else {
              const int dy = taps == 9 ? tap / 3 - 1 : 0;
              const int dx = taps == 9 ? tap % 3 - 1 : 0;
              mbar_expect_tx(&full_bar[stage], 128u * 128u + p.b_sub_bytes);
              tma_load_4d(sa, mapA, &full_bar[stage], cb * 64, t.x0 + dx, t.y0 + dy, ab);
              tma_load_3d(sb, &mapW, &full_bar[stage], p.seg_koff[s] + tap * p.seg_c[s] + cb * 64, t.n0, wb);
            }
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        uint32_t accumulate = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const bool halo = p.seg_halo[s] != 0;
          const int items = halo ? p.seg_cblk[s] : p.seg_taps[s] * p.seg_cblk[s];
          const int nb = halo ? 9 : 1;
          const uint32_t sbo_a = halo ? (uint32_t)(p.tile_w + 2) * 128u : 1024u;
          for (int it = 0; it < items; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
            const uint32_t sb = sa + p.a_region_bytes;
            for (int j = 0; j < nb; ++j) {
              const uint32_t a_off = halo ? (uint32_t)((j / 3) * (p.tile_w + 2) + (j % 3)) * 128u : 0u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = umma_desc(sa + a_off + k * 32u, sbo_a, p.base_off_variant);
                const uint64_t bd = umma_desc(sb + (uint32_t)j * p.b_sub_bytes + k * 32u, 1024u, 0);
                tc_mma_f16(d_tmem, ad, bd, p.idesc, accumulate);
                accumulate = 1;
              }
            }
            tc_commit(&empty_bar[stage]);          // frees this smem stage when the MMAs retire
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        }
        tc_commit(&tfull_bar[acc]);                // accumulator complete -> epilogue
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // =========================== epilogue ===============================
    const int q = warp & 3;                        // TMEM lane quadrant this warp may read
    const int r = q * 32 + lane;                   // accumulator row == pixel within the tile
    const int ty = r / p.tile_w, tx = r - ty * p.tile_w;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int y = t.y0 + ty, x = t.x0 + tx;
      const bool valid = (y < p.H) && (x < p.W);
      const int64_t pix = ((int64_t)t.b * p.H + y) * p.W + x;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
      for (int c = 0; c < p.n_tile; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr0 + (uint32_t)c, v);
        tmem_ld_wait();
        if (valid) {
          const int n = t.n0 + c;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + n + j);
          }
          if (p.rowbias) {
            const float* rb = p.rowbias + (int64_t)t.b * p.rowbias_ld + n;
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(rb + j);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
          if (p.residual) {
            if (kOutF32) {
              const float4* rp = reinterpret_cast<const float4*>((const float*)p.residual + pix * p.res_ld + n);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 rv = rp[j];
                f[4 * j + 0] = fmaf(p.beta, rv.x, f[4 * j + 0]); f[4 * j + 1] = fmaf(p.beta, rv.y, f[4 * j + 1]);
                f[4 * j + 2] = fmaf(p.beta, rv.z, f[4 * j + 2]); f[4 * j + 3] = fmaf(p.beta, rv.w, f[4 * j + 3]);
              }
            } else {
              const uint4* rp = reinterpret_cast<const uint4*>((const __nv_bfloat16*)p.residual + pix * p.res_ld + n);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 raw = rp[j];
                const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[8 * j + i] = fmaf(p.beta, __bfloat162float(e[i]), f[8 * j + i]);
              }
            }
          }
          if (p.act == MUDIFF_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = sigmoid_f(f[j]);
          } else if (p.act == MUDIFF_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
          } else if (p.act == MUDIFF_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = tanhf(f[j]);
          }
          if (kOutF32) {
            float4* op = reinterpret_cast<float4*>((float*)p.out + pix * p.out_ld + p.out_coff + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>((__nv_bfloat16*)p.out + pix * p.out_ld + p.out_coff + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 raw;
              __nv_bfloat162* e = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
              for (int i = 0; i < 4; ++i) e[i] = __floats2bfloat162_rn(f[8 * j + 2 * i], f[8 * j + 2 * i + 1]);
              op[j] = raw;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// NHWC activation map: dims (C, W, H, B), box (64, bw, bh, 1), bf16, 128B swizzle, zero OOB fill
int make_map_a(CUtensorMap* m, const void* ptr, int C, int ld, int W, int H, int B, int bw, int bh) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return MUDIFF_EUNSUPPORTED;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MUDIFF_EINVAL;
}

// packed weights: dims (Ktot, N, WB), box (64, n_tile, 1)
int make_map_w(CUtensorMap* m, const void* ptr, int ktot, int w_ld, int n, int wb, int64_t w_bstride, int n_tile) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return MUDIFF_EUNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)n, (cuuint64_t)wb};
  cuuint64_t strides[2] = {(cuuint64_t)w_ld * 2, wb > 1 ? (cuuint64_t)w_bstride * 2 : (cuuint64_t)w_ld * 2 * n};
  cuuint32_t box[3] = {64, (cuuint32_t)n_tile, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MUDIFF_EINVAL;
}

}  // namespace

extern "C" int mudiff_conv_tc(const mudiff_conv_desc* d, void* stream) {
  if (!d || d->nseg < 1 || d->nseg > 3 || d->batch <= 0 || d->h <= 0 || d->w <= 0 || d->n <= 0) return MUDIFF_EINVAL;
  if (!d->wt || !d->out) return MUDIFF_EINVAL;
  if (d->stride != 1 || d->stats) return MUDIFF_EUNSUPPORTED;
  if (d->n % 32 || d->out_ld % 8 || d->out_coff % 8 || (d->residual && d->res_ld % 8)) return MUDIFF_EUNSUPPORTED;
  if (d->out_dtype != MUDIFF_BF16 && d->out_dtype != MUDIFF_F32) return MUDIFF_EUNSUPPORTED;
  bool any9 = false;
  int ktot = 0;
  for (int s = 0; s < d->nseg; ++s) {
    if (!d->a[s] || d->a_c[s] <= 0 || d->a_c[s] % 64 || d->a_ld[s] % 8 || ((uintptr_t)d->a[s] % 16)) return MUDIFF_EUNSUPPORTED;
    if (d->a_taps[s] != 1 && d->a_taps[s] != 9) return MUDIFF_EINVAL;
    any9 |= d->a_taps[s] == 9;
    ktot += d->a_taps[s] * d->a_c[s];
  }
  if (any9 && d->pad != 1) return MUDIFF_EUNSUPPORTED;
  if (((uintptr_t)d->wt % 16) || ((uintptr_t)d->out % 16) || (d->w_ld % 8) || (d->w_bstride % 8)) return MUDIFF_EUNSUPPORTED;

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.batch = d->batch; p.H = d->h; p.W = d->w;
  // N tiling: largest multiple of 32 that divides n and is <= 256
  int n_tile = 0;
  for (int c = 256; c >= 32; c -= 32) if (d->n % c == 0) { n_tile = c; break; }
  if (!n_tile) return MUDIFF_EUNSUPPORTED;
  p.n_tile = n_tile; p.n_tiles = d->n / n_tile;
  // A staging mode
  // halo staging by default whenever it applies (N <= 64: the A operand dominates L2->smem traffic)
  bool want_halo = true;
  if (d->flags & 2) want_halo = false;
  if (!any9 || d->w < 8 || d->h < 2 || n_tile > 64) want_halo = false;
  if (want_halo) { p.tile_w = 8; p.tile_h = 16; }
  else if (d->h == 1) { p.tile_w = 128; p.tile_h = 1; }
  else if (d->w >= 16) { p.tile_w = 16; p.tile_h = 8; }
  else { p.tile_w = 8; p.tile_h = 16; }
  p.tiles_x = (d->w + p.tile_w - 1) / p.tile_w;
  p.tiles_y = (d->h + p.tile_h - 1) / p.tile_h;
  p.tiles_per_img = p.tiles_x * p.tiles_y;
  p.total_tiles = p.tiles_per_img * d->batch * p.n_tiles;
  p.nseg = d->nseg;
  p.a_batched = d->a_batched ? 1 : 0;
  p.w_batched = d->w_bstride != 0 ? 1 : 0;
  p.b_sub_bytes = (uint32_t)n_tile * 128u;
  uint32_t a_region = 128u * 128u;
  int nb_max = 1;
  int koff = 0;
  for (int s = 0; s < d->nseg; ++s) {
    p.seg_c[s] = d->a_c[s]; p.seg_cblk[s] = d->a_c[s] / 64; p.seg_taps[s] = d->a_taps[s];
    p.seg_halo[s] = (want_halo && d->a_taps[s] == 9) ? 1 : 0;
    p.seg_koff[s] = koff; koff += d->a_taps[s] * d->a_c[s];
    if (p.seg_halo[s]) {
      uint32_t hb = (uint32_t)(p.tile_w + 2) * (p.tile_h + 2) * 128u;
      if (hb > a_region) a_region = hb;
      nb_max = 9;
    }
  }
  p.a_region_bytes = (a_region + 1023u) & ~1023u;
  p.stage_bytes = p.a_region_bytes + (uint32_t)nb_max * p.b_sub_bytes;   // b_sub_bytes is a multiple of 1024 (n_tile % 8 == 0)
  p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
  const uint32_t tail = 1024;   // barriers + tmem slot
  int stages = (int)((kSmemBudget - tail - 1024) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return MUDIFF_EUNSUPPORTED;
  p.num_stages = stages;
  // UMMA instruction descriptor: D=f32, A=B=bf16, both K-major, N = n_tile, M = 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_tile >> 3) << 17) | ((128u >> 4) << 24);
  p.acc_stride = n_tile <= 64 ? 64 : (n_tile <= 128 ? 128 : 256);
  p.tmem_cols = 2 * p.acc_stride;
  p.base_off_variant = (d->flags & 4) ? 1 : 0;
  p.bias = d->bias; p.rowbias = d->rowbias; p.rowbias_ld = d->rowbias_ld;
  p.residual = d->residual; p.res_ld = d->res_ld; p.alpha = d->alpha; p.beta = d->beta; p.act = d->act;
  p.out = d->out; p.out_ld = d->out_ld; p.out_coff = d->out_coff;

  CUtensorMap maps[3], mapw;
  memset(maps, 0, sizeof(maps));
  for (int s = 0; s < 3; ++s) {
    int ss = s < d->nseg ? s : 0;
    int bw = p.seg_halo[ss] ? p.tile_w + 2 : p.tile_w;
    int bh = p.seg_halo[ss] ? p.tile_h + 2 : p.tile_h;
    int rc = make_map_a(&maps[s], d->a[ss], d->a_c[ss], d->a_ld[ss], d->w, d->h, d->a_batched ? d->batch : 1, bw, bh);
    if (rc) return rc;
  }
  {
    int rc = make_map_w(&mapw, d->wt, ktot, d->w_ld > 0 ? d->w_ld : ktot, d->n, p.w_batched ? d->batch : 1, d->w_bstride, n_tile);
    if (rc) return rc;
  }
  const size_t smem_bytes = (size_t)stages * p.stage_bytes + tail + 1024;
  int grid = p.total_tiles < MUDIFF_NUM_SMS ? p.total_tiles : MUDIFF_NUM_SMS;
  cudaStream_t st = (cudaStream_t)stream;
  // opt in to the full 227 KB of dynamic shared memory once per kernel instantiation (per device)
  static bool attr_set[16][2] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  const int which = d->out_dtype == MUDIFF_F32 ? 1 : 0;
  if (!attr_set[dev][which]) {
    cudaError_t e = which ? cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)
                          : cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev][which] = true;
  }
  if (which) conv_tc_kernel<true><<<grid, kThreads, smem_bytes, st>>>(maps[0], maps[1], maps[2], mapw, p);
  else conv_tc_kernel<false><<<grid, kThreads, smem_bytes, st>>>(maps[0], maps[1], maps[2], mapw, p);
  return mudiff_launch_status();
}
