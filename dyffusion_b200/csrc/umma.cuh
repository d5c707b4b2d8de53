// PTX wrappers shared by the tcgen05 kernels (conv_umma.cu, conv_up.cu): mbarriers, cp.async / bulk-TMA copies,
// tcgen05.mma issue blocks, TMEM loads, UMMA shared-memory descriptors, and the tensor-map encoder.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dyf {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {  // arrive when this thread's prior cp.async land
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One filter tap = KSTEPS back-to-back MMAs (K = 16 each) whose descriptors differ by compile-time constants; issued
// from ONE asm block by the elected lane so that no per-MMA election / convergence code is generated.
template <int KSTEPS, int AK, int BK>
__device__ __forceinline__ void umma_tap(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc_first,
                                         uint32_t leader) {
  static_assert(KSTEPS == 2 || KSTEPS == 4, "unsupported chunk depth");
  if constexpr (KSTEPS == 4) {
    asm volatile(
        "{\n\t"
        ".reg .pred pl, pa, pt;\n\t"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "setp.ne.b32 pl, %5, 0;\n\t"
        "setp.ne.b32 pa, %4, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "add.u64 a1, %1, %6;\n\t add.u64 a2, a1, %6;\n\t add.u64 a3, a2, %6;\n\t"
        "add.u64 b1, %2, %7;\n\t add.u64 b2, b1, %7;\n\t add.u64 b3, b2, %7;\n\t"
        "@pl tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t"
        "@pl tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n\t"
        "@pl tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, pt;\n\t"
        "@pl tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, pt;\n\t"
        "}" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc_first), "r"(leader), "n"((long long)AK), "n"((long long)BK)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred pl, pa, pt;\n\t"
        ".reg .b64 a1, b1;\n\t"
        "setp.ne.b32 pl, %5, 0;\n\t"
        "setp.ne.b32 pa, %4, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "add.u64 a1, %1, %6;\n\t"
        "add.u64 b1, %2, %7;\n\t"
        "@pl tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t"
        "@pl tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n\t"
        "}" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc_first), "r"(leader), "n"((long long)AK), "n"((long long)BK)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_if(uint32_t bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred pl;\n\tsetp.ne.b32 pl, %1, 0;\n\t"
      "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(leader)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
// K-major, no-swizzle shared-memory matrix descriptor (sm_100 version bits = 1):
//   addr(row, k16half) = start + (row % 8) * 16 + (row / 8) * SBO + k16half * LBO
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// 32 consecutive accumulator columns of this thread's TMEM lane in one instruction (the wait is separate so that the
// caller can batch loads before consuming them).
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}
// 4-D tensor map over a bf16 NHWC activation tensor [rows, H, W, ld] (ld = channel stride, >= C): box = 64 channels x
// bw x bh pixels x bn images, 128-B swizzle, zero fill out of bounds (= the convolution's zero padding).
inline int make_nhwc_tmap(const act_t* base, int rows, int H, int W, int C, int ld, int bw, int bh, int bn,
                          CUtensorMap* out) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return -1;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)rows};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t est[4] = {1, 1, 1, 1};
  return fn(out, DYF_TMAP_DTYPE, 4, const_cast<act_t*>(base), gdim, gstr, box, est,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

// General 4-D bf16 tensor map: dims (innermost first, dims[0] = channels with unit stride), byte strides of dims 1..3,
// box extents; 128-B swizzle, zero fill out of bounds.  Dimension order is free (e.g. (C, image, W, H)).
inline int make_tmap4(const act_t* base, const cuuint64_t dims[4], const cuuint64_t strides[3], const cuuint32_t box[4],
                      CUtensorMap* out, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return -1;
  cuuint32_t est[4] = {1, 1, 1, 1};
  return fn(out, DYF_TMAP_DTYPE, 4, const_cast<act_t*>(base), dims, strides, box, est,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace
}  // namespace dyf
