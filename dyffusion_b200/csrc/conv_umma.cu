// Implicit-GEMM 3x3 (stride 1) convolution on the 5th-generation tensor cores: tcgen05.mma (UMMA) with the
// accumulator in TMEM, weights streamed by bulk-TMA (cp.async.bulk + mbarrier complete_tx) and the activations
// staged ONCE per channel chunk as a halo patch that all nine taps re-read through shifted UMMA descriptors.
//
// Tile: 16 x 8 output pixels (M = 128) x BN output channels.  For a 64-channel chunk the input patch is
// 18 x 10 pixels; it is stored K-major / no-swizzle as eight 16-byte channel planes [chunk8][pixel], so that
//   * 8 consecutive M rows (one UMMA core matrix) = 8 consecutive pixels of a patch row  (16 B apart),
//   * consecutive 8-row groups = consecutive patch rows                                   (SBO = 10 * 16 B),
//   * the two 16-byte K halves of a K=16 step = consecutive channel planes                 (LBO = plane stride),
// and tap (ky, kx) is nothing but a start-address offset of (ky * 10 + kx) * 16 B.  Activations therefore cross
// L2 -> SM once (x1.4 halo) instead of nine times, which is what lets the N = 64 / 128 layers (60 % of the
// Navier-Stokes FLOPs) feed the tensor pipe.  Zero padding = zero-filled cp.async.
//
// Warp roles (192 threads): warps 0-3 gather patches (cp.async) and later run the epilogue (tcgen05.ld ->
// fused affine/activation/dropout -> 128-bit stores), warp 4 allocates TMEM and issues the MMAs (one thread),
// warp 5 streams weight tiles.  Pipelines: A patches 2 stages, B tiles 4 stages, all on mbarriers.
#include "conv.cuh"

namespace dyf {
namespace {

constexpr int TILE_H = 16, TILE_W = 8;          // output pixels per CTA tile (M = 128)
constexpr int PATCH_H = TILE_H + 2, PATCH_W = TILE_W + 2;
constexpr int PATCH_PIX = PATCH_H * PATCH_W;    // 180
constexpr int A_PLANE = (PATCH_PIX + 1) * 16;   // bytes per 8-channel plane; odd pixel pitch => conflict-free fills
constexpr int A_STAGE = 8 * A_PLANE;            // one 64-channel chunk
constexpr int A_STAGES = 2, B_STAGES = 3;
constexpr int THREADS = 192;

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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct __align__(8) Barriers {
  uint64_t a_full[A_STAGES], a_empty[A_STAGES], b_full[B_STAGES], b_empty[B_STAGES], acc_full;
  uint32_t tmem_base;
};

template <int BN>
__global__ void __launch_bounds__(THREADS) conv3x3_umma_kernel(const ConvParams p, const __nv_bfloat16* __restrict__ wblob,
                                                              int tiles_x, int tiles_y) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int B_STAGE = BN * 128;  // [kchunk 8][BN rows][16 B]
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_STAGES * A_STAGE;
  Barriers* bars = reinterpret_cast<Barriers*>(sB + B_STAGES * B_STAGE);
  float* s_tab = reinterpret_cast<float*>(bars + 1);  // [2][BN]: epilogue multiplier / offset of this (row, n_tile)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.y;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int row = t / tiles_y;           // batch row
  const int oy0 = ty * TILE_H, ox0 = tx * TILE_W;
  const int nchunks = p.Cin >> 6;        // 64-channel chunks

  if (tid == 0) {
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(smem_u32(&bars->a_full[i]), 128); mbar_init(smem_u32(&bars->a_empty[i]), 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {  // TMEM allocation (power of two >= 32 columns), owned by this warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (warp == 5) {  // per-(row, channel) epilogue tables: one batch row per tile, so a plain smem copy suffices
    const size_t tb = (size_t)row * p.Cout + n_tile * BN;
    for (int i = lane; i < BN; i += 32) { s_tab[i] = __ldg(p.tabA + tb + i); s_tab[BN + i] = __ldg(p.tabB + tb + i); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = bars->tmem_base;

  if (warp < 4) {
    // =============================== A producer: halo patches via cp.async (zero fill = padding) ===============
    constexpr int ITEMS = PATCH_PIX * 8;               // 16-byte items per chunk
    constexpr int ITERS = (ITEMS + 127) / 128;
    const int chunk8 = tid & 7;                        // fixed 8-channel plane of this thread (128 % 8 == 0)
    int src_off[ITERS];                                // element offset of the patch pixel, -1 = outside the image
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int pix = (tid >> 3) + it * 16;
      int off = -1;
      if (pix < PATCH_PIX) {
        const int pr = pix / PATCH_W, pc = pix - pr * PATCH_W;
        const int iy = oy0 - 1 + pr, ix = ox0 - 1 + pc;
        if ((unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi) off = (iy * p.Wi + ix) * p.Cin;
      }
      src_off[it] = off;
    }
    const __nv_bfloat16* in_row = p.in + (size_t)row * p.Hi * p.Wi * p.Cin + chunk8 * 8;
    const __nv_bfloat16* const in_base = p.in;
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % A_STAGES;
      mbar_wait(smem_u32(&bars->a_empty[st]), ((c / A_STAGES) & 1) ^ 1);
      const uint32_t dst0 = smem_u32(sA + st * A_STAGE + chunk8 * A_PLANE);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int pix = (tid >> 3) + it * 16;
        if (pix < PATCH_PIX) {
          const bool v = src_off[it] >= 0;
          cp_async16(dst0 + pix * 16, v ? (const void*)(in_row + src_off[it] + c * 64) : (const void*)in_base, v ? 16 : 0);
        }
      }
      // hardware arrives on a_full[st] once this thread's copies have landed (no wait, no proxy fence needed:
      // same pattern as CUTLASS's sm100 cp.async mainloop); the producer immediately moves on to the next chunk
      cp_async_arrive_noinc(smem_u32(&bars->a_full[st]));
    }

    // =============================== epilogue: TMEM -> registers -> fused math -> global ===========================
    mbar_wait(smem_u32(&bars->acc_full), 0);
    tc_fence_after();
    const int m_local = warp * 32 + lane;               // accumulator row = TMEM lane
    const int oy = oy0 + (m_local >> 3), ox = ox0 + (m_local & 7);
    const bool valid = oy < p.Ho && ox < p.Wo;
    const long long m = ((long long)row * p.Ho + oy) * p.Wo + ox;
    __nv_bfloat16* const orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.out_ld + p.out_coff + n_tile * BN;
    const __nv_bfloat16* const rrow = p.res ? p.res + (size_t)m * p.res_ld + n_tile * BN : nullptr;
    const int act = p.act;
    const uint32_t thresh = p.drop.thresh;
    const float dscale = p.drop.scale;
#pragma unroll 2
    for (int c0 = 0; c0 < BN; c0 += 8) {
      {
        uint32_t v[8];
        tmem_ld8(tmem_acc + ((uint32_t)(warp * 32) << 16) + c0, v);
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = fmaf(__uint_as_float(v[j]), s_tab[c0 + j], s_tab[BN + c0 + j]);
        switch (act) {
          case ACT_RELU:
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
            break;
          case ACT_LEAKY:
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = y[j] > 0.f ? y[j] : 0.2f * y[j];
            break;
          case ACT_SILU:
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = y[j] / (1.f + __expf(-y[j]));
            break;
          case ACT_GELU:
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = 0.5f * y[j] * (1.f + erff(y[j] * 0.70710678118654752f));
            break;
          default: break;
        }
        if (thresh) {
          const uint32_t keep = drop_keep_bits8(p.drop, (uint64_t)m * p.Cout + n_tile * BN + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = ((keep >> j) & 1u) ? y[j] * dscale : 0.f;
        }
        if (valid) {
          if (rrow) {
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(rrow + c0)), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] += f[j];
          }
          *reinterpret_cast<uint4*>(orow + c0) = pack8(y);
        }
      }
    }
  } else if (warp == 4) {
    // =============================== MMA issuer (one elected thread) ==============================================
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      int ib = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % A_STAGES;
        mbar_wait(smem_u32(&bars->a_full[st]), (c / A_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + st * A_STAGE);
        for (int tap = 0; tap < 9; ++tap, ++ib) {
          const int sb = ib % B_STAGES;
          mbar_wait(smem_u32(&bars->b_full[sb]), (ib / B_STAGES) & 1);
          tc_fence_after();
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint32_t a_tap = a_base + (ky * PATCH_W + kx) * 16;
          const uint32_t b_base = smem_u32(sB + sb * B_STAGE);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 64 channels = 4 x (K = 16)
            const uint64_t ad = smem_desc(a_tap + ks * 2 * A_PLANE, A_PLANE, PATCH_W * 16);
            const uint64_t bd = smem_desc(b_base + ks * 2 * (BN * 16), BN * 16, 128);
            umma_bf16(tmem_acc, ad, bd, idesc, (c | tap | ks) != 0);
          }
          umma_commit(smem_u32(&bars->b_empty[sb]));   // weight stage free once these MMAs retire
        }
        umma_commit(smem_u32(&bars->a_empty[st]));     // patch stage free
      }
      umma_commit(smem_u32(&bars->acc_full));          // accumulator complete -> epilogue
    }
  } else {
    // =============================== B producer: bulk-TMA weight tiles =============================================
    if (lane == 0) {
      const int total = nchunks * 9;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(wblob) + (size_t)n_tile * total * B_STAGE;
      for (int ib = 0; ib < total; ++ib) {
        const int sb = ib % B_STAGES;
        mbar_wait(smem_u32(&bars->b_empty[sb]), ((ib / B_STAGES) & 1) ^ 1);
        mbar_expect_tx(smem_u32(&bars->b_full[sb]), B_STAGE);
        bulk_g2s(smem_u32(sB + sb * B_STAGE), src + (size_t)ib * B_STAGE, B_STAGE, smem_u32(&bars->b_full[sb]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(BN));
  }
}

// weights fp32 [O, I, 3, 3] -> bf16 blobs [n_tile][chunk][tap][k8 (8)][n (BN)][8]  (one contiguous tile per MMA stage)
__global__ void __launch_bounds__(256) repack_umma_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                         int O, int I, int BN, int standardize) {
  const int nchunks = I >> 6;
  const long long total = (long long)O * I * 9;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // decode the destination index
  long long r = idx;
  const int e = (int)(r % 8); r /= 8;
  const int n = (int)(r % BN); r /= BN;
  const int k8 = (int)(r % 8); r /= 8;
  const int tap = (int)(r % 9); r /= 9;
  const int chunk = (int)(r % nchunks); r /= nchunks;
  const int n_tile = (int)r;
  const int o = n_tile * BN + n, ci = chunk * 64 + k8 * 8 + e;
  float v = w[((size_t)o * I + ci) * 9 + tap];
  if (standardize) {  // WeightStandardizedConv2d (reference unet.py:32-40), recomputed per element (load-time only)
    const float* wo = w + (size_t)o * I * 9;
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < I * 9; ++i) s1 += wo[i];
    const float mean = s1 / (I * 9);
    for (int i = 0; i < I * 9; ++i) { const float d = wo[i] - mean; s2 += d * d; }
    v = (v - mean) * rsqrtf(s2 / (I * 9) + 1e-5f);
  }
  out[idx] = __float2bfloat16_rn(v);
}

template <int BN>
int launch_t(const ConvParams& p, cudaStream_t stream) {
  constexpr int smem = A_STAGES * A_STAGE + B_STAGES * BN * 128 + (int)sizeof(Barriers) + 2 * BN * 4 + 64;
  static bool configured = false;
  if (!configured) {
    DYF_CUDA_OK(cudaFuncSetAttribute(conv3x3_umma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles_x = (p.Wo + TILE_W - 1) / TILE_W, tiles_y = (p.Ho + TILE_H - 1) / TILE_H;
  dim3 grid((unsigned)(tiles_x * tiles_y * p.rows), (unsigned)(p.Cout / BN));
  const double flops = 2.0 * (double)p.M * p.Cout * 9.0 * p.Cin_real;
  const double bytes = 2.0 * ((double)p.rows * p.Hi * p.Wi * p.Cin + (double)p.M * p.Cout + (double)p.Cout * p.Kpad);
  ProfScope prof(stream, KC_CONV_UMMA, flops, bytes);
  conv3x3_umma_kernel<BN><<<grid, THREADS, smem, stream>>>(p, p.w_umma, tiles_x, tiles_y);
  DYF_LAUNCH_OK("conv3x3_umma_kernel");
  return 1;
}

}  // namespace

int umma_tile_n(int Cout) { return Cout == 64 ? 64 : 128; }

bool conv_umma_shape_ok(int Cin_pad, int Cout, int k, int stride, int pad) {
  return k == 3 && stride == 1 && pad == 1 && Cin_pad % 64 == 0 && (Cout == 64 || Cout % 128 == 0);
}

bool conv_umma_eligible(const ConvParams& p) {
  return p.w_umma != nullptr && conv_umma_shape_ok(p.Cin, p.Cout, p.KH, p.stride, p.pad) && p.KW == 3 &&
         p.Ho == p.Hi && p.Wo == p.Wi && p.out_fp32 != 2;
}

int launch_conv_umma(const ConvParams& p, cudaStream_t stream) {
  if (!conv_umma_eligible(p)) return 0;
  return umma_tile_n(p.Cout) == 64 ? launch_t<64>(p, stream) : launch_t<128>(p, stream);
}

int launch_repack_umma(const float* w, __nv_bfloat16* out, int O, int I, int standardize, cudaStream_t s) {
  const long long total = (long long)O * I * 9;
  repack_umma_kernel<<<cdiv(total, 256), 256, 0, s>>>(w, out, O, I, umma_tile_n(O), standardize);
  DYF_LAUNCH_OK("repack_umma_kernel");
  return 0;
}

}  // namespace dyf
