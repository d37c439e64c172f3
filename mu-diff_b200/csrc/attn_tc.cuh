// Fused attention for AttnBlockpp (backbones/layerspp.py:111-137):  O = softmax(Q K^T * scale) V  in ONE kernel,
// flash-style (online softmax), tcgen05 + TMEM + TMA.  Included by conv_tc.cu (same PTX wrappers / debug channel).
//
//   inputs   qk  [B, L, 2C] bf16  (q = channels [0,C), k = channels [C,2C) of one GEMM output)
//            vt  [B, C, L]  bf16  (V^T, K-major operand of the P V product)
//   output   o   [B, L, C]  bf16
//   C == 256 (one head of dimension 256: nf = 64, ch_mult[-1] = 4), L % 128 == 0 (L = 4096 at 256^2).
//
// Persistent CTA per SM, 256 threads, unit = 128 query rows of one image:
//   warp 0   TMA producer: Q tile (4 x [128 x 64ch], once per unit) and K tiles ([128 keys x 256 ch] per step)
//   warp 2   TMA producer: V^T tiles ([256 ch x 128 keys] per step)
//   warp 1   TMEM allocator + MMA issuer:  S_j = Q K_j^T (16 UMMA 128x128x16, fp32 in TMEM, double-buffered)
//            issued one step AHEAD of  O += P_j V_j (8 UMMA 128x256x16 into a 128x256 fp32 TMEM accumulator)
//   warps 4-7 softmax, one query row per thread: row max / exp2 / row sum from TMEM (two passes over tcgen05.ld),
//            P_j -> bf16 -> 128B-swizzled shared memory (A operand of the P V product).  The accumulator is
//            rescaled lazily: only when some row maximum of the warp grows by more than 2^8 over the maximum the
//            accumulator is currently expressed in (tcgen05.ld -> multiply -> tcgen05.st), which after the
//            first few key tiles practically never happens.  Final O / rowsum -> bf16 stores.
// TMEM: S0 [0,128) S1 [128,256) O [256,512).  Shared memory: Q 64 KB, K 64 KB, V 64 KB, P 32 KB.
// Every mbarrier has exactly one in-order waiter role (see the A-ring note in conv_tc.cu).

struct AttnP {
  int batch, L, C, qtiles, nk;
  long long total_units;
  float scale_log2;                      // softmax scale * log2(e)
  __nv_bfloat16* out;
  uint32_t idesc_qk, idesc_pv;
};

constexpr uint32_t kAttnQ = 0, kAttnK = 65536, kAttnV = 131072, kAttnPs = 196608, kAttnBar = 229376;
constexpr uint32_t kAttnSmem = kAttnBar + 1024 + 1024;
// barrier indices inside the barrier block
enum { AB_QFULL = 0, AB_QEMPTY, AB_KFULL, AB_KEMPTY, AB_VFULL, AB_VEMPTY, AB_SFULL0, AB_SFULL1, AB_SFREE0, AB_SFREE1,
       AB_PREADY, AB_PVDONE, AB_COUNT };

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float exp2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap mapQK, const __grid_constant__ CUtensorMap mapV,
               const __grid_constant__ AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint8_t* bar_block = smem + kAttnBar;
  uint64_t* bar = (uint64_t*)(smem + kAttnBar);
  uint32_t* tmem_slot = (uint32_t*)(bar + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < AB_COUNT; ++i) mbar_init(&bar[i], (i == AB_SFREE0 || i == AB_SFREE1 || i == AB_PREADY) ? 4 : 1);
    for (int i = 0; i < 32; ++i) ((int*)(bar_block + kDbgRecOff))[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapV) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // strided unit assignment: at any time the grid works on ~gridDim.x consecutive query tiles = 4-5 images, so the K / V^T
  // tiles every query tile of an image re-reads (4 MB per image) stay in L2.  (Contiguous ranges spread the CTAs over all
  // 64 images at once: 256 MB of K / V^T, half of the re-reads missed L2 - 4.0 GB of DRAM traffic for 0.54 GB of operands.)
  const long long u_begin = blockIdx.x, u_end = p.total_units, u_step = gridDim.x;
  const int nk = p.nk;

  if (warp == 0) {
    // ===================== Q / K producer =====================
    if (lane == 0) {
      uint32_t qc = 0, kc = 0;
      for (long long u = u_begin; u < u_end; u += u_step) {
        const int b = (int)(u / p.qtiles), q0 = (int)(u % p.qtiles) * 128;
        mbar_wait(&bar[AB_QEMPTY], (qc & 1u) ^ 1u, bar_block, (int)qc);
        mbar_expect_tx(&bar[AB_QFULL], 65536u);
        for (int cb = 0; cb < 4; ++cb) tma_load_3d(smem + kAttnQ + cb * 16384, &mapQK, &bar[AB_QFULL], cb * 64, q0, b);
        ++qc;
        for (int j = 0; j < nk; ++j) {
          mbar_wait(&bar[AB_KEMPTY], (kc & 1u) ^ 1u, bar_block, (int)kc);
          mbar_expect_tx(&bar[AB_KFULL], 65536u);
          for (int cb = 0; cb < 4; ++cb)
            tma_load_3d(smem + kAttnK + cb * 16384, &mapQK, &bar[AB_KFULL], p.C + cb * 64, j * 128, b);
          ++kc;
        }
      }
    }
  } else if (warp == 2) {
    // ===================== V^T producer =====================
    if (lane == 0) {
      uint32_t vc = 0;
      for (long long u = u_begin; u < u_end; u += u_step) {
        const int b = (int)(u / p.qtiles);
        for (int j = 0; j < nk; ++j) {
          mbar_wait(&bar[AB_VEMPTY], (vc & 1u) ^ 1u, bar_block, (int)vc);
          mbar_expect_tx(&bar[AB_VFULL], 65536u);
          for (int t = 0; t < 2; ++t) tma_load_3d(smem + kAttnV + t * 32768, &mapV, &bar[AB_VFULL], j * 128 + t * 64, 0, b);
          ++vc;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t sb = smem_u32(smem);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);                 // SBO 1024, version 1, SWIZZLE_128B
    auto lo = [&](uint32_t off) { return (((sb + off) & 0x3FFFFu) >> 4) | (1u << 16); };
    uint32_t qc = 0, kc = 0, vc = 0, sc = 0, pc = 0;
    auto issue_qk = [&]() {
      const uint32_t sbuf = sc & 1u;
      mbar_wait(&bar[AB_SFREE0 + sbuf], ((sc >> 1) & 1u) ^ 1u, bar_block, (int)sc);
      mbar_wait(&bar[AB_KFULL], kc & 1u, bar_block, (int)kc);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          const uint64_t ad = ((uint64_t)hi << 32) | lo(kAttnQ + cb * 16384);
          const uint64_t bd = ((uint64_t)hi << 32) | lo(kAttnK + cb * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_base + sbuf * 128u, ad + 2 * k, bd + 2 * k, p.idesc_qk, (cb | k) ? 1u : 0u);
        }
        tc_commit(&bar[AB_KEMPTY]);
        tc_commit(&bar[AB_SFULL0 + sbuf]);
      }
      ++kc; ++sc;
    };
    for (long long u = u_begin; u < u_end; u += u_step) {
      mbar_wait(&bar[AB_QFULL], qc & 1u, bar_block, (int)qc);
      tc_fence_after();
      issue_qk();
      for (int j = 0; j < nk; ++j) {
        if (j + 1 < nk) issue_qk();
        else if (leader) tc_commit(&bar[AB_QEMPTY]);               // all S products of this unit are issued
        mbar_wait(&bar[AB_PREADY], pc & 1u, bar_block, (int)pc);
        mbar_wait(&bar[AB_VFULL], vc & 1u, bar_block, (int)vc);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint64_t ad = ((uint64_t)hi << 32) | lo(kAttnPs + t * 16384);
            const uint64_t bd = ((uint64_t)hi << 32) | lo(kAttnV + t * 32768);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16(tmem_base + 256u, ad + 2 * k, bd + 2 * k, p.idesc_pv, (j | t | k) ? 1u : 0u);
          }
          tc_commit(&bar[AB_VEMPTY]);
          tc_commit(&bar[AB_PVDONE]);
        }
        ++pc; ++vc;
      }
      ++qc;
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== softmax / epilogue =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* prow = smem + kAttnPs + row * 128;
    const int rsw = row & 7;
    uint32_t tcount = 0;                       // key tiles processed by this CTA so far (== S / P / PV sequence number)
    for (long long u = u_begin; u < u_end; u += u_step) {
      const int b = (int)(u / p.qtiles), q0 = (int)(u % p.qtiles) * 128;
      float m_used = 0.f, l = 0.f;
      for (int j = 0; j < nk; ++j, ++tcount) {
        const uint32_t sbuf = tcount & 1u;
        mbar_wait(&bar[AB_SFULL0 + sbuf], (tcount >> 1) & 1u, bar_block, (int)tcount);
        tc_fence_after();
        const uint32_t s_addr = lane_base + sbuf * 128u;
        // pass 1: row maximum
        float mx = -3.0e38f;
#pragma unroll
        for (int c = 0; c < 128; c += 32) {
          uint32_t v[32];
          tmem_ld32(s_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float mt = mx * p.scale_log2;
        float alpha = 1.f;
        bool rescale = false;
        if (j == 0) {
          m_used = mt;
        } else {
          rescale = __any_sync(0xffffffffu, mt > m_used + 8.f);
          if (rescale) {
            const float m_new = fmaxf(m_used, mt);
            alpha = exp2_approx(m_used - m_new);
            m_used = m_new;
            l *= alpha;
          }
        }
        // pass 2: P = exp2(s * scale - m_used), packed to bf16; row sum of the ROUNDED values
        uint32_t pk[64];
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < 128; c += 32) {
          uint32_t v[32];
          tmem_ld32(s_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float e0 = exp2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_used));
            const float e1 = exp2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_used));
            const __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
            const float2 r = __bfloat1622float2(h);
            rs += r.x + r.y;
            pk[(c + i) >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          }
        }
        l += rs;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[AB_SFREE0 + sbuf]);          // S buffer may be overwritten by Q K_{j+2}^T
        // the P buffer is free and the accumulator is stable once the previous P V product has retired
        if (j > 0) { mbar_wait(&bar[AB_PVDONE], (tcount - 1) & 1u, bar_block, (int)tcount); tc_fence_after(); }
        if (rescale) {
#pragma unroll 1
          for (int c = 0; c < 256; c += 32) {
            uint32_t v[32];
            tmem_ld32(lane_base + 256u + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(lane_base + 256u + c, v);
          }
          tmem_st_wait();
        }
        // P row -> 128B-swizzled K-major tiles (two sub-tiles of 64 keys): chunk ch of the row at (ch ^ (row & 7))
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          const int t = ch >> 3, cc = ch & 7;
          *reinterpret_cast<uint4*>(prow + t * 16384 + ((cc ^ rsw) << 4)) =
              make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[AB_PREADY]);
      }
      // epilogue: O / l -> bf16
      mbar_wait(&bar[AB_PVDONE], (tcount - 1) & 1u, bar_block, (int)tcount);
      tc_fence_after();
      const float inv = 1.f / l;
      __nv_bfloat16* orow = p.out + ((int64_t)b * p.L + q0 + row) * p.C;
#pragma unroll 1
      for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        tmem_ld32(lane_base + 256u + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 raw;
          __nv_bfloat162* e = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            e[k] = __floats2bfloat162_rn(__uint_as_float(v[8 * i + 2 * k]) * inv, __uint_as_float(v[8 * i + 2 * k + 1]) * inv);
          *reinterpret_cast<uint4*>(orow + c + 8 * i) = raw;
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
