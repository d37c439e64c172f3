// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a  (bf16 x bf16 -> fp32), version 3.
//
//   D[128 pixels, n_tile] += A[128 pixels, 64 ch] * W[n_tile, 64 ch]^T        per (tap, channel block)
//
// Persistent, warp-specialised CTA (256 threads; 384 with the operand transform), one per SM; each CTA owns a
// CONTIGUOUS range of work units (unit = MT consecutive 128-pixel tiles of one image x one N tile):
//   warp 0      TMA producer of the A rings (one ring per MMA-issuing warp): NHWC activation boxes through a 4-D
//               tensor map (conv padding = TMA out-of-bounds zero fill).  3x3 segments are staged as ONE halo box
//               {64ch, tw+2, th+2} per channel block; the nine taps are nine UMMA descriptors into that tile (start
//               address shifted by whole 128-byte pixel rows, SBO = halo row pitch).
//   warp 1      TMA producer of the B ring: K-major weight sub-tiles {64, n_tile} through a 3-D map,
//               or - when all of K x n_tile fits - ONE stationary load of the whole weight matrix.
//   warps 2, 3  TMEM allocator (warp 2) + the two single-thread tcgen05.mma issuers (UMMA 128 x n_tile x 16): warp 2
//               owns pixel tile 0 of every unit and its accumulator, warp 3 tile 1 (idle when MT == 1); a weight
//               sub-tile is released when both have committed (tcgen05.commit).
//   warps 4-7   epilogue, one per TMEM lane quadrant: tcgen05.ld -> bias / temb row-bias / alpha / residual /
//               activation -> bf16|fp32 NHWC stores, plus optional per-tile per-channel (sum, sum^2) partials for
//               the GroupNorm that consumes the output (butterfly shuffles, no atomics -> deterministic).
//   warps 8-11  (only with a_xform) GroupNorm/AdaGN scale + shift + SiLU applied to the staged A tile in shared
//               memory between its TMA load and its MMAs.
// TMEM holds acc_stages x MT accumulators so that unit i+1's MMAs overlap unit i's epilogue.
// Every mbarrier has exactly ONE in-order waiter role, and every wait is bounded (deadlock dump + trap).
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kMaxSlots = 16;
constexpr int kThreads = 256;               // warps 0-3: A producer, B producer, MMA issuer of tile 0, MMA issuer of tile 1; warps 4-7: epilogue.
constexpr int kThreadsXform = 384;          // + warps 8-11: A-operand transform (GroupNorm scale/shift + SiLU applied in shared memory)
constexpr int kAReadyOff = 832;             // byte offset of the a_ready barriers inside the barrier block
// One epilogue warp per TMEM lane quadrant.  Two warps per quadrant (8 epilogue warps, aligned or not) made
// epilogue-bound GEMMs (N = 4096, K = 256) fault intermittently on B200 (~1 launch in 10, tools/stress_conv.py).
constexpr uint32_t kSmemMax = 232448;       // 227 KB opt-in limit per CTA

struct TcParams {
  int batch, H, W;
  int tile_h, tile_w, tiles_x, tpi;           // tpi = pixel tiles per image
  int MT, gpi;                                // tiles per unit, unit groups per image
  int n_tiles, n_tile, n_total;
  long long total_units;
  int nseg;
  int seg_cblk[3], seg_taps[3], seg_halo[3], seg_koff[3], seg_c[3];
  int a_batched, w_batched;
  uint32_t a_slot_bytes, b_sub_bytes, off_b, off_stats, off_bar;
  int a_slots, a_ring, b_slots, stationary, b_total_subs;   // a_slots = MT * a_ring
  int issuers;                                // conv_tc2_kernel: issuing warps in the leader CTA (1 or 2), a_slots = issuers * a_ring
  uint32_t idesc;
  int acc_stride, acc_stages, tmem_cols;
  int base_off_variant;
  int round_robin;                            // debug: interleaved instead of contiguous unit assignment
  int dbg_dry, dbg_noepi, dbg_nostore, dbg_noldtm;
  int dbg_xf;                                 // timing ablations of the operand transform: 1 = barrier hop only, 2 = no activation
  int dbg_skew;                               // adversarial schedules: bit0 delay MMA warp 2, bit1 MMA warp 3, bit2 epilogue, bit3 A producer
  int dec2, Ho, Wo;                           // stride-2 VALID conv as a decimated 'same' conv: keep odd (y, x) only                     // debug: no operand traffic / no epilogue work (timing only)
  // epilogue
  const float* bias; const float* rowbias; int rowbias_ld;
  const void* residual; int res_ld;
  float alpha, beta; int act;
  void* out; int out_ld, out_coff;
  float* stats_partial;                       // [batch*tpi*4][n_total][2] (one row per image, tile, TMEM lane quadrant) or NULL
  // A-operand transform: y = act(x * scale[b][c] + shift[b][c]) applied to the staged tile of segment s
  const float* xform[3]; int xform_ld[3]; int xform_any, xform_act;
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
// Bounded wait: a protocol bug must surface as a launch error, never as a hung GPU.  Before trapping,
// the waiter copies the CTA's whole 1 KB barrier block (all mbarrier words + one progress record per
// warp) into a mapped host buffer (readable after the context died): tools/decode_timeout.py.
__device__ int* g_dbg_host = nullptr;         // device pointer of a pinned, mapped host int[1024]
__device__ int g_dbg_lock = 0;                // first timed-out waiter writes the record
constexpr int kDbgRecOff = 640;               // byte offset of the per-warp records inside the barrier block
__device__ int g_dbg_done = 0;
// Bound of every mbarrier wait in clock64 cycles (default 4e9 ~ 2 s); 0 = wait forever.  Settable through
// mudiff_set_wait_timeout() / the MUDIFF_WAIT_CYCLES environment variable: a time-sliced or preempted GPU, or a long
// profiler replay, can legitimately stall a CTA for longer than any fixed bound.
__device__ long long g_wait_cycles = 4000000000LL;
__device__ __noinline__ void mbar_timeout(uint64_t* bar, uint32_t parity, const uint8_t* bar_block) {
  int* d = g_dbg_host;
  if (d) {
    if (atomicCAS(&g_dbg_lock, 0, 1) == 0) {
      d[1] = (int)blockIdx.x; d[2] = (int)(threadIdx.x >> 5); d[3] = (int)(threadIdx.x & 31);
      d[4] = (int)smem_u32(bar); d[5] = (int)parity; d[6] = (int)gridDim.x; d[7] = (int)smem_u32(bar_block);
      const int* src = reinterpret_cast<const int*>(bar_block);
      for (int i = 0; i < 256; ++i) d[16 + i] = src[i];
      __threadfence_system();
      d[0] = 1;
      __threadfence_system();
      atomicExch(&g_dbg_done, 1);
    } else {                                   // let the first waiter finish its record before the trap kills the grid
      const long long t0 = clock64();
      while (atomicAdd(&g_dbg_done, 0) == 0 && clock64() - t0 < 200000000LL) {}
    }
  }
  __trap();
}
// rec = this warp's int[4] progress record {barrier byte offset in the block, parity, tag, waiting?}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, const uint8_t* bar_block, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  int* rec = (int*)(bar_block + kDbgRecOff) + (threadIdx.x >> 5) * 4;
  if ((threadIdx.x & 31) == 0) {
    rec[0] = (int)(smem_u32(bar) - smem_u32(bar_block)); rec[1] = (int)parity; rec[2] = tag; rec[3] = 1;
  }
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    const long long lim = *(volatile long long*)&g_wait_cycles;
    if (lim > 0 && clock64() - t0 > lim) mbar_timeout(bar, parity, bar_block);
  }
  if ((threadIdx.x & 31) == 0) rec[3] = 0;
}
// progress mark outside of mbarrier waits (named barriers, TMEM loads): code < 0
__device__ __forceinline__ void dbg_mark(const uint8_t* bar_block, int code, int tag) {
  if ((threadIdx.x & 31) == 0) {
    int* rec = (int*)(bar_block + kDbgRecOff) + (threadIdx.x >> 5) * 4;
    rec[0] = code; rec[2] = tag; rec[3] = 2;
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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


struct Unit { int b, r0, count, n0; };
__device__ __forceinline__ Unit decode_unit(const TcParams& p, long long u) {
  Unit t;
  const int nt = (int)(u % p.n_tiles);
  const long long g = u / p.n_tiles;
  t.b = (int)(g / p.gpi);
  const int gi = (int)(g - (long long)t.b * p.gpi);
  t.r0 = gi * p.MT;
  t.count = p.tpi - t.r0 < p.MT ? p.tpi - t.r0 : p.MT;
  t.n0 = nt * p.n_tile;
  return t;
}

// Enumerates the "A groups" of one unit in the order every role walks them:
//   halo segment    : one group per channel block, 9 weight sub-tiles (taps) each
//   plain segment   : one group per (tap, channel block), 1 weight sub-tile each
template <typename F>
__device__ __forceinline__ void for_each_group(const TcParams& p, F&& f) {
  for (int s = 0; s < p.nseg; ++s) {
    if (p.seg_halo[s]) {
      for (int cb = 0; cb < p.seg_cblk[s]; ++cb) f(s, cb, 0, 9);
    } else {
      for (int tap = 0; tap < p.seg_taps[s]; ++tap)
        for (int cb = 0; cb < p.seg_cblk[s]; ++cb) f(s, cb, tap, 1);
    }
  }
}

// sum over the 32 lanes of 32 per-lane values; lane j ends up with the total of v[j]
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = upper ? v[i] : v[i + n];
      const float keep = upper ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Epilogue arithmetic of one 32-column chunk of one accumulator row (shared by conv_tc_kernel and conv_tc2_kernel, so that the
// two kernels stay bit-identical): f = (acc + bias) * alpha [+ beta * residual] -> activation -> bf16 | fp32 store.
// Packed fp32x2 instructions (add / mul / fma on register pairs): half the issue slots of the scalar form, identical rounding.
// On return f[] holds the values as stored (bf16-rounded for bf16 outputs) for the optional statistics.
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <bool kOutF32>
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, const uint32_t (&v)[32], float (&f)[32], const float* sbias_c,
                                               int64_t pix, int n, bool want_rounded) {
  f32x2 t[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) t[j] = pack2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
  if (p.bias || p.rowbias) {
    const float4* sb4 = reinterpret_cast<const float4*>(sbias_c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bv = sb4[j];
      t[2 * j] = add2(t[2 * j], pack2(bv.x, bv.y));
      t[2 * j + 1] = add2(t[2 * j + 1], pack2(bv.z, bv.w));
    }
  }
  const f32x2 al2 = pack2(p.alpha, p.alpha);
#pragma unroll
  for (int j = 0; j < 16; ++j) t[j] = mul2(t[j], al2);
  if (p.residual) {
    const f32x2 be2 = pack2(p.beta, p.beta);
    if (kOutF32) {
      const float4* rp = reinterpret_cast<const float4*>((const float*)p.residual + pix * p.res_ld + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 rv = rp[j];
        t[2 * j] = fma2(be2, pack2(rv.x, rv.y), t[2 * j]);
        t[2 * j + 1] = fma2(be2, pack2(rv.z, rv.w), t[2 * j + 1]);
      }
    } else {
      const uint4* rp = reinterpret_cast<const uint4*>((const __nv_bfloat16*)p.residual + pix * p.res_ld + n);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 raw = rp[j];
        const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 r2 = __bfloat1622float2(e[i]);
          t[4 * j + i] = fma2(be2, pack2(r2.x, r2.y), t[4 * j + i]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) unpack2(t[j], f[2 * j], f[2 * j + 1]);
  if (p.act == MUDIFF_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = sigmoid_f(f[j]);
  } else if (p.act == MUDIFF_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
  } else if (p.act == MUDIFF_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = tanhf(f[j]);
  } else if (p.act == MUDIFF_ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : 0.2f * f[j];
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
      if (want_rounded) {             // statistics of what was actually stored (bf16-rounded)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 rr = __bfloat1622float2(e[i]);
          f[8 * j + 2 * i] = rr.x; f[8 * j + 2 * i + 1] = rr.y;
        }
      }
    }
  }
}

template <bool kOutF32>
__global__ void __launch_bounds__(kThreadsXform, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW,
               const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint8_t* bar_block = smem + p.off_bar;
  uint64_t* a_full = (uint64_t*)(smem + p.off_bar);
  uint64_t* a_empty = a_full + kMaxSlots;
  uint64_t* b_full = a_empty + kMaxSlots;
  uint64_t* b_empty = b_full + kMaxSlots;
  uint64_t* w_full = b_empty + kMaxSlots;
  uint64_t* tfull_bar = w_full + 1;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
  uint64_t* a_ready = (uint64_t*)(smem + p.off_bar + kAReadyOff);   // transformed A tile ready (4 arrivals: one per transform warp)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_slots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_ready[i], 4); }
    for (int i = 0; i < p.b_slots; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], p.MT); }     // one arrival per issuing warp
    for (int i = 0; i < 48; ++i) ((int*)(bar_block + kDbgRecOff))[i] = 0;      // progress records of up to 12 warps
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], p.MT); mbar_init(&tempty_bar[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    if (p.nseg > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA1) : "memory");
    if (p.nseg > 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA2) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // contiguous, balanced range of units for this CTA
  const long long u_begin = p.round_robin ? (long long)blockIdx.x : (p.total_units * (long long)blockIdx.x) / gridDim.x;
  const long long u_end = p.round_robin ? p.total_units : (p.total_units * (long long)(blockIdx.x + 1)) / gridDim.x;
  const long long u_step = p.round_robin ? (long long)gridDim.x : 1;

  if (warp == 0) {
    // =========================== A producer ===========================
    if (lane == 0 && !p.dbg_dry) {
      // One A ring PER ISSUING WARP (tile m of every group goes to ring m, `a_ring` slots each): every ring has
      // exactly one in-order consumer.  A single ring shared by the two MMA warps let slots alternate between
      // them whenever the ring size was odd; the faster warp could then be two phases ahead on a slot and
      // its parity wait passed on the previous phase (ABA) -> stale operands / deadlock (seen on B200).
      uint32_t a_cnt0 = 0, a_cnt1 = 0;
      const uint32_t ra = (uint32_t)p.a_ring;
      for (long long u = u_begin; u < u_end; u += u_step) {
        const Unit un = decode_unit(p, u);
        const int ab = p.a_batched ? un.b : 0;
        for_each_group(p, [&](int s, int cb, int tap, int nb) {
          const CUtensorMap* mapA = s == 0 ? &mapA0 : (s == 1 ? &mapA1 : &mapA2);
          for (int m = 0; m < un.count; ++m) {
            const int r = un.r0 + m;
            const int ty = r / p.tiles_x;
            const int y0 = ty * p.tile_h, x0 = (r - ty * p.tiles_x) * p.tile_w;
            if (p.dbg_skew & 8) __nanosleep(1500);
            const uint32_t k = m == 0 ? a_cnt0 : a_cnt1;
            const uint32_t slot = (uint32_t)m * ra + k % ra;
            const uint32_t par = ((k / ra) & 1u) ^ 1u;
            mbar_wait(&a_empty[slot], par, bar_block, (int)k);
            uint8_t* sa = smem + (size_t)slot * p.a_slot_bytes;
            if (nb == 9) {
              mbar_expect_tx(&a_full[slot], (uint32_t)(p.tile_w + 2) * (p.tile_h + 2) * 128u);
              tma_load_4d(sa, mapA, &a_full[slot], cb * 64, x0 - 1, y0 - 1, ab);
            } else {
              const int dy = p.seg_taps[s] == 9 ? tap / 3 - 1 : 0;
              const int dx = p.seg_taps[s] == 9 ? tap % 3 - 1 : 0;
              mbar_expect_tx(&a_full[slot], 128u * 128u);
              tma_load_4d(sa, mapA, &a_full[slot], cb * 64, x0 + dx, y0 + dy, ab);
            }
            if (m == 0) ++a_cnt0; else ++a_cnt1;
          }
        });
      }
    }
  } else if (warp == 1) {
    // =========================== B producer ===========================
    if (lane == 0 && !p.dbg_dry) {
      if (p.stationary) {
        mbar_expect_tx(w_full, (uint32_t)p.b_total_subs * p.b_sub_bytes);
        for (int i = 0; i < p.b_total_subs; ++i)
          tma_load_3d(smem + p.off_b + (size_t)i * p.b_sub_bytes, &mapW, w_full, i * 64, 0, 0);
      } else {
        uint32_t b_item = 0;
        for (long long u = u_begin; u < u_end; u += u_step) {
          const Unit un = decode_unit(p, u);
          const int wb = p.w_batched ? un.b : 0;
          for_each_group(p, [&](int s, int cb, int tap, int nb) {
            for (int j = 0; j < nb; ++j) {
              const int tp = nb == 9 ? j : tap;
              const uint32_t slot = b_item % (uint32_t)p.b_slots;
              const uint32_t par = ((b_item / (uint32_t)p.b_slots) & 1u) ^ 1u;
              mbar_wait(&b_empty[slot], par, bar_block, (int)b_item);
              mbar_expect_tx(&b_full[slot], p.b_sub_bytes);
              tma_load_3d(smem + p.off_b + (size_t)slot * p.b_sub_bytes, &mapW, &b_full[slot],
                          p.seg_koff[s] + tp * p.seg_c[s] + cb * 64, un.n0, wb);
              ++b_item;
            }
          });
        }
      }
    }
  } else if (warp == 2 || (warp == 3 && p.MT == 2)) {
    // =========================== MMA issuers =============================
    // The single-thread issue rate bounds every N <= 192 shape (UMMA 128xNx16 needs < 100 clk of tensor
    // time), so (1) the loop is kept lean: all lanes poll the mbarriers (warp-uniform control flow; polling
    // by the elected lane alone + __syncwarp deadlocked / faulted intermittently on B200), descriptor low
    // words are precomputed per group and advanced with immediates; (2) with two pixel tiles per unit
    // there are TWO issuing warps: warp 2 owns tile 0 and its accumulator, warp 3 owns tile 1.  A weight
    // sub-tile is released when both have committed (b_empty / tfull expect two arrivals).
    const int me = warp - 2;                               // which pixel tile of the unit this warp issues
    const bool leader = elect_one();
    uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t hi_b = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t hi_a_halo = ((10u * 128u) >> 4) | (1u << 14) | (2u << 29);     // halo tiles are 8 wide: pitch 10 px
    const uint32_t b_lo_base = (((smem_base + p.off_b) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_step = p.b_sub_bytes >> 4;
    const uint32_t a_lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_step = p.a_slot_bytes >> 4;
    const bool dry = p.dbg_dry != 0;
    if (p.stationary && !dry) { mbar_wait(w_full, 0, bar_block, -1); tc_fence_after(); }
    for (long long u = u_begin; u < u_end; u += u_step) {
      const Unit un = decode_unit(p, u);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, bar_block, (int)(u - u_begin));
      tc_fence_after();
      uint32_t accum = 0;                                  // 0 only for the first MMA of the accumulator
      const uint32_t d_mine = tmem_base + (uint32_t)((acc * p.MT + me) * p.acc_stride);
      const bool active = me < un.count;                   // partial last group of an image: tile 1 may not exist
      for_each_group(p, [&](int s, int cb, int tap, int nb) {
        if (p.dbg_skew & (1 << me)) __nanosleep(2000);
        // my A tile of this group: next slot of MY ring (ring `me`, in order)
        const uint32_t sl = (uint32_t)me * (uint32_t)p.a_ring + a_slot, ph = a_phase;
        if (active && !dry) mbar_wait(p.xform_any ? &a_ready[sl] : &a_full[sl], ph, bar_block, (int)(u - u_begin));
        tc_fence_after();
        const uint32_t alo = a_lo_base + sl * a_step;
        const uint32_t hi_a = nb == 9 ? hi_a_halo : hi_b;
        const int kbase = (p.seg_koff[s] + (nb == 9 ? 0 : tap) * p.seg_c[s]) / 64 + cb;   // stationary sub-tile index of tap 0
        const int kstep = p.seg_c[s] / 64;                                                   // ... and its stride per tap
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          if (j < nb) {
            uint32_t blo;
            if (p.stationary) {
              blo = b_lo_base + (uint32_t)(kbase + j * kstep) * b_step;
            } else {
              if (!dry) mbar_wait(&b_full[b_slot], b_phase, bar_block, (int)(u - u_begin));
              tc_fence_after();
              blo = b_lo_base + b_slot * b_step;
            }
            const uint32_t toff = (uint32_t)((j / 3) * 10 + (j % 3)) * 8u;     // tap offset in 16-byte units (halo only; j == 0 otherwise)
            if (leader) {
              if (active) {
                const uint64_t bd = ((uint64_t)hi_b << 32) | blo;
                const uint64_t ad = ((uint64_t)hi_a << 32) | (alo + toff);
                tc_mma_f16(d_mine, ad, bd, p.idesc, accum);
                tc_mma_f16(d_mine, ad + 2, bd + 2, p.idesc, 1u);
                tc_mma_f16(d_mine, ad + 4, bd + 4, p.idesc, 1u);
                tc_mma_f16(d_mine, ad + 6, bd + 6, p.idesc, 1u);
              }
              if (!p.stationary) tc_commit(&b_empty[b_slot]);       // one arrival per issuing warp
            }
            if (!p.stationary) { if (++b_slot == (uint32_t)p.b_slots) { b_slot = 0; b_phase ^= 1u; } }
            accum = 1u;
          }
        }
        if (active) {
          if (leader) tc_commit(&a_empty[sl]);                      // my A tile is free once its last tap retires
          if (++a_slot == (uint32_t)p.a_ring) { a_slot = 0; a_phase ^= 1u; }
        }
      });
      if (leader) tc_commit(&tfull_bar[acc]);                // my accumulator is complete -> epilogue
      if (p.acc_stages == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else { acc_phase ^= 1; }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // =========================== epilogue ===============================
    const int q = warp & 3;                        // TMEM lane quadrant this warp may read
    const int et = (warp - 4) * 32 + lane;         // epilogue thread index 0..127
    const int row = q * 32 + lane;                 // accumulator row == pixel within the tile
    const int ty_in = row / p.tile_w, tx_in = row - ty_in * p.tile_w;
    float* sbias = (float*)(smem + p.off_stats);   // [n_tile] bias + row-bias of the current image
    int bias_b = -1, bias_n0 = -1;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long u = u_begin; u < u_end; u += u_step) {
      const Unit un = decode_unit(p, u);
      if ((p.bias || p.rowbias) && (un.b != bias_b || un.n0 != bias_n0)) {
        // stage bias[n] + rowbias[b][n] once per (image, N tile); all 128 epilogue threads take part
        dbg_mark(bar_block, -2, (int)(u - u_begin));
        asm volatile("bar.sync 2, 128;" ::: "memory");
        for (int col = et; col < p.n_tile; col += 128) {
          float bv = p.bias ? __ldg(p.bias + un.n0 + col) : 0.f;
          if (p.rowbias) bv += __ldg(p.rowbias + (int64_t)un.b * p.rowbias_ld + un.n0 + col);
          sbias[col] = bv;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        bias_b = un.b; bias_n0 = un.n0;
      }
      if (p.dbg_skew & 4) __nanosleep(20000);
      mbar_wait(&tfull_bar[acc], acc_phase, bar_block, (int)(u - u_begin));
      tc_fence_after();
      dbg_mark(bar_block, -3, (int)(u - u_begin));
      for (int m = 0; m < (p.dbg_noepi ? 0 : un.count); ++m) {
        const int r = un.r0 + m;
        const int tyt = r / p.tiles_x;
        const int y = tyt * p.tile_h + ty_in, x = (r - tyt * p.tiles_x) * p.tile_w + tx_in;
        bool valid = (y < p.H) && (x < p.W);
        int64_t pix = ((int64_t)un.b * p.H + y) * p.W + x;
        if (p.dec2) {
          valid = valid && (y & 1) && (x & 1) && (y >> 1) < p.Ho && (x >> 1) < p.Wo;
          pix = ((int64_t)un.b * p.Ho + (y >> 1)) * p.Wo + (x >> 1);
        }
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * p.MT + m) * p.acc_stride);
        for (int c = 0; c < p.n_tile; c += 32) {
          uint32_t v[32];
          if (!p.dbg_noldtm) { tmem_ld32(taddr0 + (uint32_t)c, v); tmem_ld_wait(); }
          else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          const int n = un.n0 + c;
          float f[32];
          if (valid && !p.dbg_nostore) {
            epilogue_chunk<kOutF32>(p, v, f, sbias + c, pix, n, p.stats_partial != nullptr);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;
          }
          if (p.stats_partial) {
            // per-channel (sum, sum^2) of this warp's 32 pixel rows: lane j ends up with the totals of column c + j and
            // stores them straight to the partial buffer (row = (image, tile, lane quadrant)): no shared-memory staging,
            // no named barrier; mudiff_stats_finalize adds the rows in a fixed order (deterministic, batch-invariant)
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = f[j] * f[j];
            const float cs = warp_transpose_reduce(f, lane);
            const float cq = warp_transpose_reduce(sq, lane);
            float* dst = p.stats_partial + (((((int64_t)un.b * p.tpi + r) * 4 + q) * p.n_total) + un.n0 + c + lane) * 2;
            *reinterpret_cast<float2*>(dst) = make_float2(cs, cq);
          }
        }
      }
      dbg_mark(bar_block, -4, (int)(u - u_begin));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (p.acc_stages == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else { acc_phase ^= 1; }
    }
  }

  if (warp >= 8 && p.xform_any && !p.dbg_dry) {
    // =========================== A-operand transform ===========================
    // GroupNorm / AdaGN scale + shift and SiLU applied to the TMA-staged activation tile IN shared memory, between
    // the TMA load and the MMAs: the normalised tensor is never written to HBM (it was a full read + write pass).
    // 128 threads; thread t owns the logical 16-byte channel chunk c8 = t % 8 (its 8 (scale, shift) pairs live in
    // registers) of rows t / 8 + 16 i; the physical chunk is c8 ^ (row & 7) (128-byte swizzle).  Pixels outside the
    // image stay zero (the convolution pads AFTER the activation).
    const int tx = threadIdx.x - 256;
    const int c8 = tx & 7, r_first = tx >> 3;
    uint32_t cnt0 = 0, cnt1 = 0;
    const uint32_t ra = (uint32_t)p.a_ring;
    for (long long u = u_begin; u < u_end; u += u_step) {
      const Unit un = decode_unit(p, u);
      for_each_group(p, [&](int s, int cb, int tap, int nb) {
        for (int m = 0; m < un.count; ++m) {
          const uint32_t k = m == 0 ? cnt0 : cnt1;
          const uint32_t slot = (uint32_t)m * ra + k % ra;
          mbar_wait(&a_full[slot], (k / ra) & 1u, bar_block, (int)k);
          const float* tab = p.xform[s];
          if (tab && p.dbg_xf != 1) {
            const int r = un.r0 + m;
            const int ty = r / p.tiles_x;
            int oy = ty * p.tile_h, ox = (r - ty * p.tiles_x) * p.tile_w;
            int rows, wpx;
            if (nb == 9) { rows = (p.tile_w + 2) * (p.tile_h + 2); wpx = p.tile_w + 2; oy -= 1; ox -= 1; }
            else {
              rows = 128; wpx = p.tile_w;
              if (p.seg_taps[s] == 9) { oy += tap / 3 - 1; ox += tap % 3 - 1; }
            }
            // this thread's 8 channels: (scale, shift) of the pre-activation, halved for silu(t) = h + h tanh(h), h = t/2
            const float4* tp = reinterpret_cast<const float4*>(tab + ((int64_t)un.b * p.xform_ld[s] + cb * 64 + c8 * 8) * 2);
            float sc[8], sh[8];
            const bool do_silu = p.xform_act == MUDIFF_ACT_SILU && p.dbg_xf != 2;
            const float hf = do_silu ? 0.5f : 1.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 v4 = __ldg(tp + i);
              sc[2 * i] = v4.x * hf; sh[2 * i] = v4.y * hf; sc[2 * i + 1] = v4.z * hf; sh[2 * i + 1] = v4.w * hf;
            }
            uint8_t* base = smem + (size_t)slot * p.a_slot_bytes;
            // six rows per trip: independent load -> convert -> FMA -> MUFU -> pack -> store chains (one warp per
            // scheduler has no other way to hide the latencies); (hy, hx) advance incrementally (no divisions)
            constexpr int RU = 6;
            int hy = r_first / wpx, hx = r_first - hy * wpx;
            const int dyy = 16 / wpx, dxx = 16 - dyy * wpx;          // advance of 16 rows in (hy, hx)
            for (int rr0 = r_first; rr0 < rows; rr0 += 16 * RU) {
              uint4 raw[RU];
              bool ok[RU];
#pragma unroll
              for (int j = 0; j < RU; ++j) {
                const int rr = rr0 + 16 * j;
                const int y = oy + hy, x = ox + hx;
                ok[j] = rr < rows && (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
                raw[j] = make_uint4(0u, 0u, 0u, 0u);
                if (ok[j]) raw[j] = *reinterpret_cast<const uint4*>(base + rr * 128 + ((c8 ^ (rr & 7)) << 4));
                hy += dyy; hx += dxx;
                if (hx >= wpx) { hx -= wpx; ++hy; }
              }
#pragma unroll
              for (int j = 0; j < RU; ++j) {
                __nv_bfloat162* e = reinterpret_cast<__nv_bfloat162*>(&raw[j]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = __bfloat1622float2(e[i]);
                  float h0 = fmaf(f.x, sc[2 * i], sh[2 * i]), h1 = fmaf(f.y, sc[2 * i + 1], sh[2 * i + 1]);
                  if (do_silu) { h0 = fmaf(h0, tanh_approx(h0), h0); h1 = fmaf(h1, tanh_approx(h1), h1); }
                  e[i] = __floats2bfloat162_rn(h0, h1);
                }
              }
#pragma unroll
              for (int j = 0; j < RU; ++j) {
                const int rr = rr0 + 16 * j;
                if (ok[j]) *reinterpret_cast<uint4*>(base + rr * 128 + ((c8 ^ (rr & 7)) << 4)) = raw[j];
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[slot]);
          if (m == 0) ++cnt0; else ++cnt1;
        }
      });
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
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


// ---------------------------------------------------------------------------------
// planning (shared by the launcher and mudiff_conv_tc_query)
// ---------------------------------------------------------------------------------
int plan_conv(const mudiff_conv_desc* d, TcParams& p, int& ktot_out) {
  if (!d || d->nseg < 1 || d->nseg > 3 || d->batch <= 0 || d->h <= 0 || d->w <= 0 || d->n <= 0) return MUDIFF_EINVAL;
  if (!d->wt || !d->out) return MUDIFF_EINVAL;
  if (d->stride != 1) return MUDIFF_EUNSUPPORTED;
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
  ktot_out = ktot;
  if (any9 && d->pad != 1) return MUDIFF_EUNSUPPORTED;
  if (((uintptr_t)d->wt % 16) || ((uintptr_t)d->out % 16) || (d->w_ld % 8) || (d->w_bstride % 8)) return MUDIFF_EUNSUPPORTED;

  memset(&p, 0, sizeof(p));
  p.batch = d->batch; p.H = d->h; p.W = d->w;
  int n_tile = 0;
  for (int c = 256; c >= 32; c -= 32) if (d->n % c == 0) { n_tile = c; break; }
  // N tiles between 128 and 256 columns (N = 384 -> 192) leave room for ONE accumulator stage only (2 tiles x 256 TMEM columns):
  // the epilogue is then not overlapped with the next unit's MMAs (tools/conv_bench.py, N = 384, K = 1728: 1243 TFLOP/s, 1755
  // without the epilogue).  128-column tiles keep two stages: gate conv 20.3 -> 15.6 ms per bench step.  Flag 0x400000: old plan.
  if (!(d->flags & 0x400000) && n_tile > 128 && n_tile < 256 && d->n % 128 == 0) n_tile = 128;
  p.n_tile = n_tile; p.n_tiles = d->n / n_tile; p.n_total = d->n;
  // A staging: halo whenever a 3x3 segment exists and the image is at least one 8x2 patch
  bool halo = any9 && d->w >= 8 && d->h >= 2 && !(d->flags & 2);
  if (halo) { p.tile_w = 8; p.tile_h = 16; }
  else if (d->h == 1) { p.tile_w = 128; p.tile_h = 1; }
  else if (d->w >= 16) { p.tile_w = 16; p.tile_h = 8; }
  else { p.tile_w = 8; p.tile_h = 16; }
  p.tiles_x = (d->w + p.tile_w - 1) / p.tile_w;
  const int tiles_y = (d->h + p.tile_h - 1) / p.tile_h;
  p.tpi = p.tiles_x * tiles_y;
  p.nseg = d->nseg;
  p.a_batched = d->a_batched ? 1 : 0;
  p.w_batched = d->w_bstride != 0 ? 1 : 0;
  p.b_sub_bytes = (uint32_t)n_tile * 128u;
  int koff = 0;
  p.a_slot_bytes = 128u * 128u;
  for (int s = 0; s < d->nseg; ++s) {
    p.seg_c[s] = d->a_c[s]; p.seg_cblk[s] = d->a_c[s] / 64; p.seg_taps[s] = d->a_taps[s];
    p.seg_halo[s] = (halo && d->a_taps[s] == 9) ? 1 : 0;
    p.seg_koff[s] = koff; koff += d->a_taps[s] * d->a_c[s];
    if (p.seg_halo[s]) {
      uint32_t hb = ((uint32_t)(p.tile_w + 2) * (p.tile_h + 2) * 128u + 1023u) & ~1023u;
      if (hb > p.a_slot_bytes) p.a_slot_bytes = hb;
    }
  }
  // TMEM plan: MT tiles per unit x acc_stages accumulator sets
  int pow2 = 32; while (pow2 < n_tile) pow2 <<= 1;
  int MT = (p.tpi >= 2 && !(d->flags & 16)) ? 2 : 1;
  int stages = 2;
  if (MT * stages * pow2 > 512) {
    if (n_tile > 128 && n_tile < 256 && p.tpi >= 2 && !(d->flags & 16)) { MT = 2; stages = 1; }   // e.g. n_tile = 192
    else { MT = 1; stages = 2; }
  }
  if (MT * stages * pow2 > 512) stages = 1;
  if (d->flags & 256) stages = 1;                      // debug
  if (MT * stages * pow2 > 512) return MUDIFF_EUNSUPPORTED;
  p.MT = MT; p.acc_stages = stages; p.acc_stride = pow2;
  int cols = MT * stages * pow2; int tc = 32; while (tc < cols) tc <<= 1;
  p.tmem_cols = tc;
  p.gpi = (p.tpi + MT - 1) / MT;
  p.total_units = (long long)d->batch * p.gpi * p.n_tiles;
  // shared memory plan
  const uint32_t bar_bytes = 1024;
  const uint32_t stats_bytes = (uint32_t)n_tile * 4u;        // [n_tile] bias staging of the epilogue
  const uint32_t fixed = bar_bytes + ((stats_bytes + 1023u) & ~1023u) + 1024u /*alignment slack*/;
  const uint32_t avail = kSmemMax - fixed;
  const uint32_t b_total = (uint32_t)(ktot / 64) * p.b_sub_bytes;
  // A ring depth: two groups of tiles for halo staging; plain 16 KB tiles (GEMM mode, 1x1 segments only)
  // get four groups (16 KB tiles: the extra depth is free there, and epilogue-bound GEMMs such as N = 4096, K = 256
  // need it to keep the loads ahead of the MMAs).
  // With the A-operand transform a tile spends an extra ~0.5 us between its TMA load and its MMAs: the A rings
  // must be three to four tiles deep per issuing warp (measured: two-deep rings made the fused launches 60 %
  // slower, three-deep 13 %), so stationary weights are only taken when they leave that much room.
  const bool xf = d->a_xform[0] || (d->nseg > 1 && d->a_xform[1]) || (d->nseg > 2 && d->a_xform[2]);
  const int a_min = xf ? (MT == 2 ? 6 : 4) : ((!halo || (d->flags & 512)) ? 4 * MT : 2 * MT);
  p.stationary = 0;
  if (p.n_tiles == 1 && !p.w_batched && !(d->flags & 8) && (d->w_ld == 0 || d->w_ld == ktot) &&
      b_total + (uint32_t)a_min * p.a_slot_bytes <= avail && b_total < (1u << 20)) {
    p.stationary = 1;
    p.b_total_subs = ktot / 64;
    int as = (int)((avail - b_total) / p.a_slot_bytes);
    p.a_slots = as > 8 ? 8 : as;
    p.a_slots -= p.a_slots % MT;
    p.b_slots = 0;
    p.off_b = (uint32_t)p.a_slots * p.a_slot_bytes;
    p.off_stats = p.off_b + b_total;
  } else {
    p.a_slots = a_min;
    int bs = (int)((avail - (uint32_t)a_min * p.a_slot_bytes) / p.b_sub_bytes);
    if (bs < 2) {                       // not enough room: fall back to fewer A slots
      if (MT == 2) return MUDIFF_EUNSUPPORTED;
      return MUDIFF_EUNSUPPORTED;
    }
    // 8 weight sub-tiles in flight are enough to cover the TMA latency and leave room for a third A tile per issuing
    // warp (+2 % on the streaming N = 64 launches, tools/bcap_ab.sh); flag 1024 restores the older 12
    const int b_cap = (d->flags & 1024) ? 12 : 8;
    if (bs > b_cap) {                   // spend the surplus on a deeper A ring
      int extra = (int)(((uint32_t)(bs - b_cap) * p.b_sub_bytes) / p.a_slot_bytes);
      p.a_slots += extra; if (p.a_slots > 8) p.a_slots = 8;
      p.a_slots -= p.a_slots % MT;
      bs = b_cap;
    }
    p.b_slots = bs;
    p.off_b = (uint32_t)p.a_slots * p.a_slot_bytes;
    p.off_stats = p.off_b + (uint32_t)bs * p.b_sub_bytes;
  }
  p.a_ring = p.a_slots / MT;
  if (p.a_ring < 1 || p.a_ring * MT != p.a_slots) return MUDIFF_EUNSUPPORTED;
  p.off_bar = p.off_stats + ((stats_bytes + 1023u) & ~1023u);
  if (p.off_bar + bar_bytes + 1024u > kSmemMax) return MUDIFF_EUNSUPPORTED;
  // UMMA instruction descriptor: D=f32, A=B=bf16, both K-major, N = n_tile, M = 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_tile >> 3) << 17) | ((128u >> 4) << 24);
  p.base_off_variant = (d->flags & 4) ? 1 : 0;
  p.round_robin = (d->flags & 32) ? 1 : 0;
  p.dbg_dry = (d->flags & 64) ? 1 : 0;
  p.dbg_noepi = (d->flags & 128) ? 1 : 0;
  p.dbg_nostore = (d->flags & 2048) ? 1 : 0;
  p.dbg_noldtm = (d->flags & 4096) ? 1 : 0;
  p.dbg_skew = (d->flags >> 16) & 15;
  p.dbg_xf = (d->flags >> 20) & 3;
  p.dec2 = (d->flags & 32768) ? 1 : 0;
  p.Ho = p.dec2 ? (d->h - 1) / 2 : d->h;
  p.Wo = p.dec2 ? (d->w - 1) / 2 : d->w;
  p.bias = d->bias; p.rowbias = d->rowbias; p.rowbias_ld = d->rowbias_ld;
  p.residual = d->residual; p.res_ld = d->res_ld; p.alpha = d->alpha; p.beta = d->beta; p.act = d->act;
  p.out = d->out; p.out_ld = d->out_ld; p.out_coff = d->out_coff;
  p.stats_partial = (float*)d->stats;
  p.xform_any = 0; p.xform_act = d->a_xform_act;
  for (int s = 0; s < 3; ++s) {
    p.xform[s] = s < d->nseg ? d->a_xform[s] : nullptr;
    p.xform_ld[s] = s < d->nseg ? d->a_xform_ld[s] : 0;
    if (p.xform[s]) {
      if (((uintptr_t)p.xform[s] % 16) || p.xform_ld[s] % 2 || p.xform_ld[s] < d->a_c[s] || !d->a_batched) return MUDIFF_EUNSUPPORTED;
      p.xform_any = 1;
    }
  }
  if (p.xform_any && p.xform_act != MUDIFF_ACT_NONE && p.xform_act != MUDIFF_ACT_SILU) return MUDIFF_EINVAL;
  return 0;
}

#include "attn_tc.cuh"
#include "stem_tc.cuh"
#include "conv_tc2.cuh"

}  // namespace

__global__ void dbg_selftest_kernel() { if (g_dbg_host) { g_dbg_host[8] = 12345; __threadfence_system(); } }
static int* g_dbg_host_ptr = nullptr;
static void ensure_dbg() {
  static bool done = false;
  if (done) return;
  done = true;
  int* h = nullptr;
  if (cudaHostAlloc((void**)&h, 4096 + 64, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return; }
  memset(h, 0, 4096 + 64);
  int* dptr = nullptr;
  if (cudaHostGetDevicePointer((void**)&dptr, h, 0) != cudaSuccess) { cudaGetLastError(); return; }
  if (cudaMemcpyToSymbol(g_dbg_host, &dptr, sizeof(dptr)) != cudaSuccess) { cudaGetLastError(); return; }
  g_dbg_host_ptr = h;
}

// out[0] = 1 if a barrier wait timed out in the last conv_tc kernels; out[1..6] = block, warp, lane,
// barrier smem address, parity, grid.  Readable even after the CUDA context reported a launch failure.
extern "C" int mudiff_debug_selftest(void) {
  ensure_dbg();
  if (!g_dbg_host_ptr) return -1;
  dbg_selftest_kernel<<<1, 1>>>();
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  return ((volatile int*)g_dbg_host_ptr)[8] == 12345 ? 1 : 0;
}

extern "C" int mudiff_set_wait_timeout(long long cycles) {
  if (cycles < 0) return MUDIFF_EINVAL;
  return (int)cudaMemcpyToSymbol(g_wait_cycles, &cycles, sizeof(cycles));
}

extern "C" int mudiff_debug_last_timeout(int32_t* out) {
  if (!g_dbg_host_ptr) { for (int i = 0; i < 8; ++i) out[i] = 0; return 0; }
  for (int i = 0; i < 8; ++i) out[i] = ((volatile int*)g_dbg_host_ptr)[i];
  return 0;
}

// Full record of the last timeout: out[0..15] header as above (out[7] = shared address of the barrier block),
// out[16..271] = the CTA's 1 KB barrier block (mbarrier words, then per-warp progress records at byte 640).
extern "C" int mudiff_debug_dump(int32_t* out, int n) {
  for (int i = 0; i < n; ++i) out[i] = (g_dbg_host_ptr && i < 1024 + 16) ? ((volatile int*)g_dbg_host_ptr)[i] : 0;
  return 0;
}

extern "C" int mudiff_conv_tc_query(const mudiff_conv_desc* d, int32_t* out) {
  TcParams p; int ktot = 0;
  int rc = plan_conv(d, p, ktot);
  if (rc) return rc;
  out[0] = p.tile_h; out[1] = p.tile_w; out[2] = p.tpi; out[3] = p.n_tile; out[4] = p.MT;
  out[5] = p.stationary; out[6] = p.a_slots; out[7] = p.b_slots; out[8] = p.acc_stages; out[9] = p.seg_halo[0];
  return 0;
}

// Stem conv3x3(1 -> n) (+ folded GroupNorm/AdaGN scale-shift, + activation) on the tensor cores, see stem_tc.cuh.
// x fp32 [B, H, W] dense; wt fp32 [n][9]; scale_shift fp32 [B][n][2] or NULL; out NHWC bf16 / fp32.
extern "C" int mudiff_stem_conv_tc(const float* x, const float* wt, const float* bias, const float* scale_shift, int act,
                                   void* out, int out_ld, int out_coff, int out_dtype, int batch, int h, int w, int n,
                                   void* stream) {
  if (!x || !wt || !out || batch <= 0 || h <= 0 || w <= 0 || n <= 0) return MUDIFF_EINVAL;
  if (n % 32 || n > 256 || out_ld % 8 || out_coff % 8 || ((uintptr_t)out % 16)) return MUDIFF_EUNSUPPORTED;
  if (act != MUDIFF_ACT_NONE && act != MUDIFF_ACT_SILU) return MUDIFF_EUNSUPPORTED;
  if (out_dtype != MUDIFF_BF16 && out_dtype != MUDIFF_F32) return MUDIFF_EUNSUPPORTED;
  ensure_dbg();
  StemTcP p;
  p.x = x; p.H = h; p.W = w; p.batch = batch; p.wt = wt; p.bias = bias; p.scale_shift = scale_shift; p.N = n; p.act = act;
  p.out = out; p.out_ld = out_ld; p.out_coff = out_coff;
  p.tiles_per_image = (int)(((long long)h * w + 127) / 128);
  p.total_units = (long long)batch * p.tiles_per_image;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
  static bool attr_set[16][2] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  const int which = out_dtype == MUDIFF_F32 ? 1 : 0;
  if (!attr_set[dev][which]) {
    cudaError_t e = which ? cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStemSmem)
                          : cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStemSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev][which] = true;
  }
  static int per_sm = 0;
  if (!per_sm) { const char* e = getenv("MUDIFF_STEM_TC_CTAS"); per_sm = e ? atoi(e) : 2; if (per_sm < 1) per_sm = 1; }
  const long long gmax = (long long)MUDIFF_NUM_SMS * per_sm;
  const int grid = (int)(p.total_units < gmax ? p.total_units : gmax);
  cudaStream_t st = (cudaStream_t)stream;
  if (which) stem_tc_kernel<true><<<grid, 288, kStemSmem, st>>>(p);
  else stem_tc_kernel<false><<<grid, 288, kStemSmem, st>>>(p);
  return mudiff_launch_status();
}

// Fused attention O = softmax(Q K^T * scale) V (AttnBlockpp, backbones/layerspp.py:118-122), see attn_tc.cuh.
extern "C" int mudiff_attention_tc(const void* qk, const void* vt, void* out, int batch, int L, int C, float scale,
                                   void* stream) {
  if (!qk || !vt || !out || batch <= 0 || L <= 0 || C <= 0) return MUDIFF_EINVAL;
  if (C != 256 || L % 128) return MUDIFF_EUNSUPPORTED;
  if (((uintptr_t)qk % 16) || ((uintptr_t)vt % 16) || ((uintptr_t)out % 16)) return MUDIFF_EUNSUPPORTED;
  ensure_dbg();
  CUtensorMap mqk, mv;
  int rc = make_map_w(&mqk, qk, 2 * C, 2 * C, L, batch, (int64_t)L * 2 * C, 128);
  if (rc) return rc;
  rc = make_map_w(&mv, vt, L, L, C, batch, (int64_t)C * L, 256);
  if (rc) return rc;
  AttnP p;
  p.batch = batch; p.L = L; p.C = C; p.qtiles = L / 128; p.nk = L / 128;
  p.total_units = (long long)batch * p.qtiles;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = (__nv_bfloat16*)out;
  p.idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  p.idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  static bool attr_set[16] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const int grid = (int)(p.total_units < MUDIFF_NUM_SMS ? p.total_units : MUDIFF_NUM_SMS);
  attn_tc_kernel<<<grid, 256, kAttnSmem, (cudaStream_t)stream>>>(mqk, mv, p);
  return mudiff_launch_status();
}

extern "C" int mudiff_conv_tc(const mudiff_conv_desc* d, void* stream) {
  TcParams p; int ktot = 0;
  int rc = plan_conv(d, p, ktot);
  if (rc) return rc;
  ensure_dbg();
  CUtensorMap maps[3], mapw;
  memset(maps, 0, sizeof(maps));
  for (int s = 0; s < 3; ++s) {
    int ss = s < d->nseg ? s : 0;
    int bw = p.seg_halo[ss] ? p.tile_w + 2 : p.tile_w;
    int bh = p.seg_halo[ss] ? p.tile_h + 2 : p.tile_h;
    rc = make_map_a(&maps[s], d->a[ss], d->a_c[ss], d->a_ld[ss], d->w, d->h, d->a_batched ? d->batch : 1, bw, bh);
    if (rc) return rc;
  }
  // CTA-pair kernel (cta_group::2) for the shared-memory-bound N = 64 / 128 launches with stationary weights (conv_tc2.cuh).
  // MUDIFF_CONV_PAIR=0 disables it.
  static int pair_on = -1;
  if (pair_on < 0) { const char* e = getenv("MUDIFF_CONV_PAIR"); pair_on = (e && e[0] == '0') ? 0 : 1; }
  TcParams p2;
  if (pair_on && !(d->flags & 0x4000) && plan_pair(d, p, ktot, p2)) {
    rc = make_map_w(&mapw, d->wt, ktot, ktot, d->n, 1, 0, p2.n_tile / 2);
    if (rc) return rc;
    const size_t smem2 = (size_t)p2.off_bar + 1024 + 1024;
    static bool attr2[16][2] = {};
    int dev2 = 0; cudaGetDevice(&dev2);
    if (dev2 < 0 || dev2 >= 16) return MUDIFF_EUNSUPPORTED;
    const int w2 = d->out_dtype == MUDIFF_F32 ? 1 : 0;
    if (!attr2[dev2][w2]) {
      cudaError_t e = w2 ? cudaFuncSetAttribute(conv_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)
                         : cudaFuncSetAttribute(conv_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
      if (e != cudaSuccess) return (int)e;
      attr2[dev2][w2] = true;
    }
    long long pairs = p2.total_units < MUDIFF_NUM_SMS / 2 ? p2.total_units : MUDIFF_NUM_SMS / 2;
    const int grid2 = (int)pairs * 2;
    if (w2) conv_tc2_kernel<true><<<grid2, kThreads, smem2, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], mapw, p2);
    else conv_tc2_kernel<false><<<grid2, kThreads, smem2, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], mapw, p2);
    return mudiff_launch_status();
  }
  rc = make_map_w(&mapw, d->wt, ktot, d->w_ld > 0 ? d->w_ld : ktot, d->n, p.w_batched ? d->batch : 1, d->w_bstride, p.n_tile);
  if (rc) return rc;
  const size_t smem_bytes = (size_t)p.off_bar + 1024 + 1024;
  long long g = p.total_units < MUDIFF_NUM_SMS ? p.total_units : MUDIFF_NUM_SMS;
  const int grid = (int)g;
  cudaStream_t st = (cudaStream_t)stream;
  // opt in to the full 227 KB of dynamic shared memory once per kernel instantiation (per device)
  static bool attr_set[16][2] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return MUDIFF_EUNSUPPORTED;
  const int which = d->out_dtype == MUDIFF_F32 ? 1 : 0;
  if (!attr_set[dev][which]) {
    cudaError_t e = which ? cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)
                          : cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev][which] = true;
  }
  const int threads = p.xform_any ? kThreadsXform : kThreads;
  if (which) conv_tc_kernel<true><<<grid, threads, smem_bytes, st>>>(maps[0], maps[1], maps[2], mapw, p);
  else conv_tc_kernel<false><<<grid, threads, smem_bytes, st>>>(maps[0], maps[1], maps[2], mapw, p);
  return mudiff_launch_status();
}
