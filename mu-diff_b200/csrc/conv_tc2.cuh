// CTA-pair variant of the tcgen05 implicit-GEMM convolution (cta_group::2), included by conv_tc.cu.
//
// Why: a single-CTA UMMA 128 x N x 16 reads A (4 KB) + B (N x 32 B) from shared memory; at N = 64 that is 6 KB per 32 tensor
// clocks = 192 B/clk against a 128 B/clk shared-memory port, so the N = 64 launches (44 % of the conv time of the bench
// workload) are bound by operand reads at 2/3 of the tensor peak, and N = 128 sits exactly at the port limit.  With a CTA pair
// (two SMs of one TPC, cluster of 2) one UMMA 256 x N x 16 uses each CTA's own 128-pixel A tile and HALF of the weight rows
// from each CTA: 4 KB + N x 16 B per CTA per instruction - 5 KB at N = 64 (port time 40 clk instead of 48), 6 KB at N = 128
// (48 clk, below the 64-clk tensor time).  The stationary weights also halve per CTA, which leaves room for a deeper A ring.
//
// Structure (256 threads per CTA, one cluster = one "pair unit" stream):
//   unit        = tiles (2g, 2g+1) of one image: CTA rank r stages / stores tile 2g + r.
//   warp 0      A producer of EACH CTA: TMA halo boxes of its own tile into its own ring; the load completes on the LEADER's
//               a_full barrier (cp.async.bulk.tensor ... cta_group::2 with a shared::cluster mbarrier address); the leader's
//               producer posts the expect_tx for both CTAs' bytes.  Slots are freed by a multicast tcgen05.commit.
//   warp 1      weight producer of each CTA: ONE stationary load of its half of the rows (completes on the leader's w_full).
//   warp 2      leader CTA only: the single-thread issuer of tcgen05.mma.cta_group::2 (M = 256); both CTAs: TMEM alloc / dealloc.
//   warps 4-7   epilogue of each CTA on its own TMEM (its 128 accumulator rows), waits the local tfull (multicast commit),
//               releases the accumulator stage with a remote arrive on the leader's tempty (8 arrivals: 4 warps x 2 CTAs).
//               Four accumulator stages (4 x N TMEM columns) decouple the MMA stream from the hand-over latency.
// Every mbarrier keeps exactly one in-order waiter role per CTA.  Restrictions (the host falls back to conv_tc_kernel otherwise):
// stationary weights, one N tile (N = 64 or 128), even number of pixel tiles per image, no operand transform, no fused
// statistics, no decimation.  (A streamed-weight variant - ring of half sub-tiles completing on the leader's barrier - was
// built and is correct, but with ONE issuing thread per pair the per-tap wait + commit makes it slower than the single-CTA
// kernel with its two issuers: N = 64, K = 2880: 23.3 vs 13.2 ms; and the extra branches in the issue loop cost the
// stationary launches 20-35 %.  Those launches stay on conv_tc_kernel.)

constexpr int kPairStagesMax = 8;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Remote arrive with the DEFAULT (.release.cta) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id) issues it.  A
// `.release.cluster` arrive made every epilogue warp drain its global stores before releasing the accumulator stage (the
// epilogue then took 4 300 clk per tile instead of < 2 000: 540 vs 1 213 TFLOP/s without the epilogue at N = 64, K = 576);
// what the MMA warp must observe is only that the tcgen05.ld of the stage have completed (tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync), not the stores.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// arrive (count 1) on the barrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs retired
__device__ __forceinline__ void tc2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc2_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <bool kOutF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW,
                const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint8_t* bar_block = smem + p.off_bar;
  uint64_t* a_full = (uint64_t*)(smem + p.off_bar);
  uint64_t* a_empty = a_full + kMaxSlots;
  uint64_t* w_full = a_empty + kMaxSlots;                 // (the b_full / b_empty rings of conv_tc_kernel are not used)
  uint64_t* tfull_bar = w_full + 1;                       // [kPairStagesMax]
  uint64_t* tempty_bar = tfull_bar + kPairStagesMax;      // [kPairStagesMax]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + kPairStagesMax);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_slots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 48; ++i) ((int*)(bar_block + kDbgRecOff))[i] = 0;
    mbar_init(w_full, 1);
    for (int i = 0; i < p.acc_stages; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    if (p.nseg > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA1) : "memory");
    if (p.nseg > 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA2) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                      // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // contiguous, balanced range of pair units for this cluster
  const long long cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const long long u_begin = (p.total_units * cid) / ncl, u_end = (p.total_units * (cid + 1)) / ncl;

  if (warp == 0) {
    // =========================== A producer (each CTA, its own tile) ===========================
    if (lane == 0 && !p.dbg_dry) {
      // one A ring per issuing warp (units alternate between the issuers): every ring keeps exactly one in-order consumer
      uint32_t cnt0 = 0u, cnt1 = 0u;
      const uint32_t nis = (uint32_t)p.issuers, ar = (uint32_t)p.a_ring;
      for (long long u = u_begin; u < u_end; ++u) {
        const int b = (int)(u / p.gpi);
        const int r = (int)(u - (long long)b * p.gpi) * 2 + (int)rank;
        const int ty = r / p.tiles_x;
        const int y0 = ty * p.tile_h, x0 = (r - ty * p.tiles_x) * p.tile_w;
        const uint32_t ring = nis == 2 ? (uint32_t)((u - u_begin) & 1) : 0u;
        for_each_group(p, [&](int s, int cb, int tap, int nb) {
          const CUtensorMap* mapA = s == 0 ? &mapA0 : (s == 1 ? &mapA1 : &mapA2);
          const uint32_t c = ring ? cnt1 : cnt0;
          const uint32_t slot = ring * ar + c % ar;
          const uint32_t par = ((c / ar) & 1u) ^ 1u;
          mbar_wait(&a_empty[slot], par, bar_block, (int)c);
          uint8_t* sa = smem + (size_t)slot * p.a_slot_bytes;
          const uint32_t bytes = nb == 9 ? (uint32_t)(p.tile_w + 2) * (p.tile_h + 2) * 128u : 128u * 128u;
          if (rank == 0) mbar_expect_tx(&a_full[slot], 2u * bytes);           // both CTAs' tiles complete on the leader's barrier
          const uint32_t bar = mapa_rank(smem_u32(&a_full[slot]), 0);
          if (nb == 9) {
            tma2_load_4d(sa, mapA, bar, cb * 64, x0 - 1, y0 - 1, b);
          } else {
            const int dy = p.seg_taps[s] == 9 ? tap / 3 - 1 : 0;
            const int dx = p.seg_taps[s] == 9 ? tap % 3 - 1 : 0;
            tma2_load_4d(sa, mapA, bar, cb * 64, x0 + dx, y0 + dy, b);
          }
          if (ring) cnt1 = c + 1; else cnt0 = c + 1;
        });
      }
    }
  } else if (warp == 1) {
    // =========================== weight producer: this CTA's half of the rows, once ===========================
    if (lane == 0 && !p.dbg_dry) {
      if (rank == 0) mbar_expect_tx(w_full, 2u * (uint32_t)p.b_total_subs * p.b_sub_bytes);
      const uint32_t bar = mapa_rank(smem_u32(w_full), 0);
      for (int i = 0; i < p.b_total_subs; ++i)
        tma2_load_3d(smem + p.off_b + (size_t)i * p.b_sub_bytes, &mapW, bar, i * 64, (int)rank * (p.n_tile / 2), 0);
    }
  } else if (warp == 2 || (warp == 3 && p.issuers == 2)) {
    // =========================== MMA issuers (leader CTA only) =============================
    // One thread cannot issue a UMMA faster than every ~54 clocks (profiles/r02_umma_collector_probe.md), above the 40-clock
    // operand-read time of a 256 x 64 x 16 pair UMMA.  Optional second issuer (p.issuers == 2, opt-in): warp 2 takes the even
    // units of the cluster's range (accumulator stages 0, 2), warp 3 the odd ones (stages 1, 3), each from its own A ring.
    if (rank == 0) {
      const int me = warp - 2;
      const int nis = p.issuers;
      const bool leader = elect_one();
      const uint32_t ar = (uint32_t)p.a_ring;
      uint32_t a_slot = 0, a_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t hi_b = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t hi_a_halo = ((10u * 128u) >> 4) | (1u << 14) | (2u << 29);
      const uint32_t b_lo_base = (((smem_base + p.off_b) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_step = p.b_sub_bytes >> 4;
      const uint32_t a_lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_step = p.a_slot_bytes >> 4;
      const bool dry = p.dbg_dry != 0;                      // timing ablation: no operand traffic
      if (!dry) mbar_wait(w_full, 0, bar_block, -1);
      tc_fence_after();
      for (long long u = u_begin + me; u < u_end; u += nis) {
        const long long iu = u - u_begin;
        const int acc = (int)(iu % p.acc_stages);
        const uint32_t acc_phase = (uint32_t)((iu / p.acc_stages) & 1);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, bar_block, (int)iu);
        tc_fence_after();
        uint32_t accum = 0;
        const uint32_t d_mine = tmem_base + (uint32_t)(acc * p.acc_stride);
        for_each_group(p, [&](int s, int cb, int tap, int nb) {
          const uint32_t sl = (uint32_t)me * ar + a_slot, ph = a_phase;
          if (!dry) mbar_wait(&a_full[sl], ph, bar_block, (int)iu);
          tc_fence_after();
          const uint32_t alo = a_lo_base + sl * a_step;
          const uint32_t hi_a = nb == 9 ? hi_a_halo : hi_b;
          const int kbase = (p.seg_koff[s] + (nb == 9 ? 0 : tap) * p.seg_c[s]) / 64 + cb;
          const int kstep = p.seg_c[s] / 64;
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            if (j < nb) {
              const uint32_t blo = b_lo_base + (uint32_t)(kbase + j * kstep) * b_step;
              const uint32_t toff = (uint32_t)((j / 3) * 10 + (j % 3)) * 8u;
              if (leader) {
                const uint64_t bd = ((uint64_t)hi_b << 32) | blo;
                const uint64_t ad = ((uint64_t)hi_a << 32) | (alo + toff);
                tc2_mma_f16(d_mine, ad, bd, p.idesc, accum);
                tc2_mma_f16(d_mine, ad + 2, bd + 2, p.idesc, 1u);
                tc2_mma_f16(d_mine, ad + 4, bd + 4, p.idesc, 1u);
                tc2_mma_f16(d_mine, ad + 6, bd + 6, p.idesc, 1u);
              }
              accum = 1u;
            }
          }
          if (leader && !dry) tc2_commit_mc(&a_empty[sl]);    // frees the slot in BOTH CTAs
          if (++a_slot == ar) { a_slot = 0; a_phase ^= 1u; }
        });
        if (leader) tc2_commit_mc(&tfull_bar[acc]);           // both CTAs' epilogues
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // =========================== epilogue (each CTA, its own 128 accumulator rows) ===============================
    const int q = warp & 3;
    const int et = (warp - 4) * 32 + lane;
    const int row = q * 32 + lane;
    const int ty_in = row / p.tile_w, tx_in = row - ty_in * p.tile_w;
    float* sbias = (float*)(smem + p.off_stats);
    int bias_b = -1;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long u = u_begin; u < u_end; ++u) {
      const int b = (int)(u / p.gpi);
      const int r = (int)(u - (long long)b * p.gpi) * 2 + (int)rank;
      if ((p.bias || p.rowbias) && b != bias_b) {
        asm volatile("bar.sync 2, 128;" ::: "memory");
        for (int col = et; col < p.n_tile; col += 128) {
          float bv = p.bias ? __ldg(p.bias + col) : 0.f;
          if (p.rowbias) bv += __ldg(p.rowbias + (int64_t)b * p.rowbias_ld + col);
          sbias[col] = bv;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        bias_b = b;
      }
      mbar_wait(&tfull_bar[acc], acc_phase, bar_block, (int)(u - u_begin));
      tc_fence_after();
      if (!p.dbg_noepi) {
        const int tyt = r / p.tiles_x;
        const int y = tyt * p.tile_h + ty_in, x = (r - tyt * p.tiles_x) * p.tile_w + tx_in;
        const bool valid = (y < p.H) && (x < p.W);
        const int64_t pix = ((int64_t)b * p.H + y) * p.W + x;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
        for (int c = 0; c < p.n_tile; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr0 + (uint32_t)c, v);
          tmem_ld_wait();
          float f[32];
          if (valid) epilogue_chunk<kOutF32>(p, v, f, sbias + c, pix, c, false);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty_bar[acc]), 0));     // the leader's MMA warp owns the stage
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();                                      // the peer's shared memory / TMEM stay alive until the pair is done
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// Plan of the pair kernel on top of plan_conv's result (same tiling / segments); returns false when the launch does not
// qualify (the caller then runs conv_tc_kernel).
static bool plan_pair(const mudiff_conv_desc* d, const TcParams& p1, int ktot, TcParams& p) {
  if (p1.xform_any || p1.n_tiles != 1 || (p1.n_tile != 64 && p1.n_tile != 128) || (p1.tpi & 1) || p1.dec2 || p1.stats_partial ||
      p1.w_batched || !p1.a_batched || p1.dbg_skew || p1.round_robin)
    return false;
  if (d->w_ld != 0 && d->w_ld != ktot) return false;
  p = p1;
  // One tile per CTA per unit, FOUR accumulator stages: with two, the fixed latency of the MMA -> epilogue -> MMA hand-over
  // (multicast commit, wake-up, tcgen05.ld, remote arrive) bounds short-K units (K = 576: 546 vs 941 TFLOP/s single-CTA).
  p.MT = 1; p.acc_stages = 4;
  int pow2 = 32; while (pow2 < p.n_tile) pow2 <<= 1;
  p.acc_stride = pow2;
  p.tmem_cols = p.acc_stages * pow2;
  p.gpi = p.tpi / 2;
  p.total_units = (long long)d->batch * p.gpi;
  if (p.total_units < MUDIFF_NUM_SMS / 2) return false;
  p.b_sub_bytes = (uint32_t)(p.n_tile / 2) * 128u;
  p.b_total_subs = ktot / 64;
  p.stationary = 1; p.b_slots = 0;
  const uint32_t bar_bytes = 1024, stats_bytes = (uint32_t)p.n_tile * 4u;
  const uint32_t fixed = bar_bytes + ((stats_bytes + 1023u) & ~1023u) + 1024u;
  const uint32_t b_total = (uint32_t)p.b_total_subs * p.b_sub_bytes;
  static int min_slots = 0;
  if (!min_slots) { const char* e = getenv("MUDIFF_PAIR_MIN_SLOTS"); min_slots = e ? atoi(e) : 3; if (min_slots < 2) min_slots = 2; }
  if (b_total + (uint32_t)min_slots * p.a_slot_bytes + fixed > kSmemMax) return false;
  int as = (int)((kSmemMax - fixed - b_total) / p.a_slot_bytes);
  p.a_slots = as > 8 ? 8 : as;
  // MUDIFF_PAIR_ISSUERS=2: two issuing warps for N = 64.  Lifts the issue floor (no operands, no epilogue: 1280 -> 1548 TFLOP/s
  // at K = 576) but not the full kernel (964 -> 959, bound by epilogue + operand traffic), and halves the A ring per issuer,
  // which costs the three-segment fused-shortcut launches 20 % (780 -> 629): off by default (profiles/r02_pair_issuers.md).
  static int two_issuers = -1;
  if (two_issuers < 0) { const char* e = getenv("MUDIFF_PAIR_ISSUERS"); two_issuers = (e && e[0] == '2') ? 1 : 0; }
  p.issuers = (two_issuers && p.n_tile == 64 && p.a_slots >= 4 && p.total_units >= MUDIFF_NUM_SMS) ? 2 : 1;
  if (p.issuers == 2) p.a_slots &= ~1;
  p.a_ring = p.a_slots / p.issuers;
  p.off_b = (uint32_t)p.a_slots * p.a_slot_bytes;
  p.off_stats = p.off_b + b_total;
  p.off_bar = p.off_stats + ((stats_bytes + 1023u) & ~1023u);
  // UMMA instruction descriptor: D = f32, A = B = bf16, K-major, N = n_tile, M = 256 (the pair)
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((256u >> 4) << 24);
  return true;
}
