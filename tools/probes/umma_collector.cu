// Probe: does tcgen05.mma's A-collector reuse (.collector::a::fill / ::use / ::lastuse) remove the shared-memory read of A?
// One CTA, one issuing thread, UMMA 128 x N x 16 (bf16) on operands already in shared memory.  Variants:
//   0  distinct A per MMA, default (discard)                       -> baseline, operand-read bound for N <= 128
//   1  groups of 3 MMAs with the SAME A, 3 different B / accumulators, default (discard)
//   2  same as 1 with fill / use / lastuse
// Prints SM clocks per MMA and a checksum of the three accumulators (1 and 2 must agree).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define MMA(suffix)                                                                                          \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                             \
               "tcgen05.mma.cta_group::1.kind::f16" suffix " [%0], %1, %2, %3, p;\n\t}"                      \
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory")
__device__ __forceinline__ void mma_plain(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { MMA(""); }
__device__ __forceinline__ void mma_fill(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { MMA(".collector::a::fill"); }
__device__ __forceinline__ void mma_use(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { MMA(".collector::a::use"); }
__device__ __forceinline__ void mma_last(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { MMA(".collector::a::lastuse"); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int kATiles = 8;
template <int N, int VARIANT, int ISSUERS>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* clocks, float* sums) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* A = (__nv_bfloat16*)smem;                           // kATiles x [128][64]
  __nv_bfloat16* B = (__nv_bfloat16*)(smem + kATiles * 16384);       // 3 x [N][64]
  uint64_t* bar = (uint64_t*)(smem + kATiles * 16384 + 3 * N * 128);
  uint32_t* slot = (uint32_t*)(bar + 2);
  for (int i = threadIdx.x; i < kATiles * 128 * 64; i += 128) A[i] = __float2bfloat16((float)((i * 7 + (i >> 9)) % 5 - 2));
  for (int i = threadIdx.x; i < 3 * N * 64; i += 128) B[i] = __float2bfloat16((float)((i * 3 + (i >> 7)) % 3 - 1));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < ISSUERS) {
    const int me = threadIdx.x >> 5;
    // lean issue loop: descriptors precomputed, 12 MMAs per trip fully unrolled (the issuing thread must not be the bound)
    const uint64_t a0 = desc(smem_u32(A), 1024);
    uint64_t bd[3];
    for (int j = 0; j < 3; ++j) bd[j] = desc(smem_u32(B) + (uint32_t)j * N * 128u, 1024);
    const uint32_t dbase = tmem + (uint32_t)((me * 3 * N) % 512);
    const uint32_t d0 = dbase, d1 = tmem + (me * 3 * N + N) % 512, d2 = tmem + (me * 3 * N + 2 * N) % 512;
    // prologue MMAs zero-initialise the accumulators (not timed separately; same for every variant)
    mma_plain(d0, a0, bd[0], idesc, 0u); mma_plain(d1, a0, bd[1], idesc, 0u); mma_plain(d2, a0, bd[2], idesc, 0u);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t at = a0 + (uint64_t)((it & (kATiles - 1)) * 1024);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = at + 2 * kk;
        if (VARIANT == 0) {
          mma_plain(d0, ad, bd[0] + 2 * kk, idesc, 1u);
          mma_plain(d1, a0 + (uint64_t)(((it + 1) & (kATiles - 1)) * 1024) + 2 * kk, bd[1] + 2 * kk, idesc, 1u);
          mma_plain(d2, a0 + (uint64_t)(((it + 2) & (kATiles - 1)) * 1024) + 2 * kk, bd[2] + 2 * kk, idesc, 1u);
        } else if (VARIANT == 1) {
          mma_plain(d0, ad, bd[0] + 2 * kk, idesc, 1u);
          mma_plain(d1, ad, bd[1] + 2 * kk, idesc, 1u);
          mma_plain(d2, ad, bd[2] + 2 * kk, idesc, 1u);
        } else {
          mma_fill(d0, ad, bd[0] + 2 * kk, idesc, 1u);
          mma_use(d1, ad, bd[1] + 2 * kk, idesc, 1u);
          mma_last(d2, ad, bd[2] + 2 * kk, idesc, 1u);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + me)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar + me)), "r"(0) : "memory");
    }
    clocks[me] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // checksum: each thread reads its accumulator row (lane = threadIdx.x) of the three accumulators
  float s = 0.f;
  const int q = threadIdx.x >> 5;
  for (int c = 0; c < (3 * N > 512 ? 512 : 3 * N); c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) s += __uint_as_float(v[i]) * (float)((c + i) % 7 + 1);
  }
  sums[threadIdx.x] = s;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N, int VARIANT, int ISSUERS>
void run1(int iters, long long* dclk, float* dsum) {
  const int smem = kATiles * 16384 + 3 * N * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe<N, VARIANT, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    probe<N, VARIANT, ISSUERS><<<1, 128, smem>>>(iters, dclk, dsum);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d variant %d: %s\n", N, VARIANT, cudaGetErrorString(e)); return; }
  }
  long long clk2[2]; float sums[128];
  cudaMemcpy(clk2, dclk, 16, cudaMemcpyDeviceToHost);
  const long long clk = ISSUERS == 2 && clk2[1] > clk2[0] ? clk2[1] : clk2[0];
  cudaMemcpy(sums, dsum, sizeof(sums), cudaMemcpyDeviceToHost);
  double cs = 0; for (int i = 0; i < 128; ++i) cs += sums[i] * (double)(i % 11 + 1);
  printf("N=%3d variant %d issuers %d: %8.2f clk per UMMA 128x%dx16 per SM (%d MMAs)  checksum %.6e\n", N, VARIANT, ISSUERS, (double)clk / (iters * 12.0 * ISSUERS), N, iters * 12 * ISSUERS, cs);
}
template <int N>
void run(int iters) {
  long long* dclk; float* dsum;
  cudaMalloc(&dclk, 16); cudaMalloc(&dsum, 128 * 4);
  run1<N, 0, 1>(iters, dclk, dsum);
  run1<N, 1, 1>(iters, dclk, dsum);
  run1<N, 2, 1>(iters, dclk, dsum);
  run1<N, 0, 2>(iters, dclk, dsum);
  run1<N, 1, 2>(iters, dclk, dsum);
  run1<N, 2, 2>(iters, dclk, dsum);
}

int main() {
  run<16>(2000);
  run<32>(2000);
  run<64>(2000);
  run<96>(2000);
  run<128>(2000);
  run<256>(2000);
  return 0;
}
