// Micro-benchmark: tcgen05.mma issue-to-retire throughput as a function of the A/B shared-memory descriptor
// geometry (start alignment, SBO, LBO, swizzle mode).  One CTA per SM, one thread issues NITER MMAs back to back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
struct Cfg { int n; int a_off, a_lbo, a_sbo, a_layout, a_kadv; int b_lbo, b_sbo, b_layout, b_kadv; int ntaps, tap_stride; };
__global__ void __launch_bounds__(128, 1) bench(Cfg c, int niter, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
    const uint64_t a_hi = (uint64_t)(((c.a_sbo >> 4) & 0x3FFF) | (1u << 14) | ((uint32_t)c.a_layout << 29)) << 32;
    const uint64_t b_hi = (uint64_t)(((c.b_sbo >> 4) & 0x3FFF) | (1u << 14) | ((uint32_t)c.b_layout << 29)) << 32;
    const uint32_t a_lo = (((uint32_t)c.a_lbo >> 4) << 16) | ((a0 + c.a_off) >> 4);
    const uint32_t b_lo = (((uint32_t)c.b_lbo >> 4) << 16) | (b0 >> 4);
    const uint32_t ts = c.tap_stride >> 4, ak = c.a_kadv >> 4, bk = c.b_kadv >> 4;
    long long t0 = clock64();
    for (int i = 0; i < niter; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint64_t ad = a_hi | (a_lo + (u >> 2) * ts + (u & 3) * ak);
        const uint64_t bd = b_hi | (b_lo + (u & 3) * bk);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tbase), "l"(ad), "l"(bd), "r"(idesc), "r"(i + u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int niter = 4096;
  struct Named { const char* name; Cfg c; };
  const int P = 2896;  // conv patch plane (181 * 16)
  Named cfgs[] = {
    // n, a_off, a_lbo, a_sbo, a_layout, a_kadv, b_lbo, b_sbo, b_layout, b_kadv, ntaps, tap_stride
    {"N64  A canonical nosw (SBO128,LBO2048)      ", {64, 0, 2048, 128, 0, 4096, 1024, 128, 0, 2048, 1, 0}},
    {"N128 A canonical nosw                        ", {128, 0, 2048, 128, 0, 4096, 2048, 128, 0, 4096, 1, 0}},
    {"N64  A patch SBO160 LBO2896 aligned start    ", {64, 0, P, 160, 0, 2 * P, 1024, 128, 0, 2048, 1, 0}},
    {"N64  A patch SBO160 start+16                 ", {64, 16, P, 160, 0, 2 * P, 1024, 128, 0, 2048, 1, 0}},
    {"N64  A patch SBO160 9 taps (conv pattern)    ", {64, 0, P, 160, 0, 2 * P, 1024, 128, 0, 2048, 9, 16}},
    {"N128 A patch SBO160 9 taps                   ", {128, 0, P, 160, 0, 2 * P, 2048, 128, 0, 4096, 9, 16}},
    {"N64  A patch SBO256 LBO4624 start+16         ", {64, 16, 4624, 256, 0, 2 * 4624, 1024, 128, 0, 2048, 1, 0}},
    {"N64  A patch SBO128(no halo) start+16        ", {64, 16, 2048, 128, 0, 4096, 1024, 128, 0, 2048, 1, 0}},
    {"N64  A SW128 canonical (SBO1024) B SW128     ", {64, 0, 16, 1024, 2, 32, 16, 1024, 2, 32, 1, 0}},
    {"N128 A SW128 canonical B SW128               ", {128, 0, 16, 1024, 2, 32, 16, 1024, 2, 32, 1, 0}},
    {"N64  A SW128 SBO1280 (10-row pitch)          ", {64, 0, 16, 1280, 2, 32, 16, 1024, 2, 32, 1, 0}},
    {"N64  A SW128 SBO1280 start+128               ", {64, 128, 16, 1280, 2, 32, 16, 1024, 2, 32, 1, 0}},
    {"N256 A SW128 canonical B SW128               ", {256, 0, 16, 1024, 2, 32, 16, 1024, 2, 32, 1, 0}},
    {"N256 A canonical nosw                        ", {256, 0, 2048, 128, 0, 4096, 4096, 128, 0, 8192, 1, 0}},
    {"N64  A patch SBO160 B SW128                  ", {64, 16, P, 160, 0, 2 * P, 16, 1024, 2, 32, 1, 0}},
    {"N64  A SW32 (SBO 256: 8 rows x 32B) K16      ", {64, 0, 16, 256, 6, 0, 1024, 128, 0, 2048, 1, 0}},
  };
  for (auto& nc : cfgs) {
    bench<<<148, 128, 200 * 1024>>>(nc.c, niter, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("%s : %7.1f cyc/MMA  (%s)\n", nc.name, (double)cyc / niter, cudaGetErrorString(e));
    if (e != cudaSuccess) break;
  }
  return 0;
}
