// Implicit-GEMM convolution, generic shapes: cp.async 4-stage pipeline + ldmatrix + mma.sync.m16n8k16 (bf16 -> fp32).
// GEMM view: M = rows*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin gathered on the fly from the NHWC input
// (zero padding = zero-filled cp.async).  Epilogue: accumulators -> smem (fp32) -> fused affine/activation/dropout/
// residual -> 128-bit coalesced stores.  This is the any-shape pipeline; the heavy layers go to conv_umma.cu.
#include "conv.cuh"

namespace dyf {
namespace {

constexpr int BM = 128, BK = 32, STAGES = 4, THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." DYF_MMA_T "." DYF_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 16-byte chunk swizzle inside a 64-byte (32 x bf16) tile row: conflict-free for ldmatrix and cp.async.
__device__ __forceinline__ int swz(int row, int chunk) { return (chunk ^ ((row >> 1) & 3)) << 3; }

template <int BN>
__global__ void __launch_bounds__(THREADS) conv_mma_kernel(const ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  act_t* sA = reinterpret_cast<act_t*>(smem_raw);
  act_t* sB = sA + STAGES * BM * BK;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int HoWo = p.Ho * p.Wo;

  // ---- per-thread gather state: each thread owns one 16-byte K segment of two A rows and BN/64 B rows
  const int seg = tid & 3, lrow = tid >> 2;
  long long abase[2];
  int iy0[2], ix0[2];
  bool mval[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    long long m = m0 + lrow + 64 * i;
    mval[i] = m < p.M;
    long long mm = mval[i] ? m : 0;
    int r = (int)(mm / HoWo);
    int rem = (int)(mm - (long long)r * HoWo);
    int oy = rem / p.Wo, ox = rem - oy * p.Wo;
    iy0[i] = oy * p.stride - p.pad;
    ix0[i] = ox * p.stride - p.pad;
    abase[i] = (((long long)r * p.Hi + iy0[i]) * p.Wi + ix0[i]) * p.Cin;
  }
  int kc = seg * 8, ky = 0, kx = 0;  // position of this thread's segment inside the (ky, kx, c) ordering of K
  while (kc >= p.Cin) { kc -= p.Cin; if (++kx == p.KW) { kx = 0; ++ky; } }

  auto load_stage = [&](int stage, int kb) {
    act_t* a = sA + stage * BM * BK;
    act_t* b = sB + stage * BN * BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = lrow + 64 * i;
      const bool v = mval[i] && ky < p.KH && (unsigned)(iy0[i] + ky) < (unsigned)p.Hi &&
                     (unsigned)(ix0[i] + kx) < (unsigned)p.Wi;
      const act_t* src = v ? p.in + abase[i] + ((long long)ky * p.Wi + kx) * p.Cin + kc : p.in;
      cp_async16(smem_u32(a + row * BK + swz(row, seg)), src, v ? 16 : 0);
    }
#pragma unroll
    for (int i = 0; i < BN / 64; ++i) {
      const int row = lrow + 64 * i;
      const int n = n0 + row;
      const bool v = n < p.Cout;
      const act_t* src = v ? p.w + (size_t)n * p.Kpad + kb * BK + seg * 8 : p.w;
      cp_async16(smem_u32(b + row * BK + swz(row, seg)), src, v ? 16 : 0);
    }
    kc += BK;
    while (kc >= p.Cin) { kc -= p.Cin; if (++kx == p.KW) { kx = 0; ++ky; } }
  };

  constexpr int NI = BN / 16;  // n8 blocks per warp (warp tile = 32 x BN/2)
  float acc[2][NI][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < NI; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;

  const int nkb = p.Kpad / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kb + STAGES - 1 < nkb) load_stage((kb + STAGES - 1) % STAGES, kb + STAGES - 1);
    cp_async_commit();
    const act_t* a = sA + (kb % STAGES) * BM * BK;
    const act_t* b = sB + (kb % STAGES) * BN * BK;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t af[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int row = warp_m * 32 + mi * 16 + (lane & 15);
        ldmatrix_x4(af[mi], smem_u32(a + row * BK + swz(row, ks * 2 + (lane >> 4))));
      }
      uint32_t bf[NI][2];
#pragma unroll
      for (int nj = 0; nj < NI / 2; ++nj) {
        const int row = warp_n * (BN / 2) + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
        uint32_t r[4];
        ldmatrix_x4(r, smem_u32(b + row * BK + swz(row, ks * 2 + ((lane >> 3) & 1))));
        bf[2 * nj][0] = r[0]; bf[2 * nj][1] = r[1]; bf[2 * nj + 1][0] = r[2]; bf[2 * nj + 1][1] = r[3];
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) mma_bf16(acc[mi][ni], af[mi], bf[ni][0], bf[ni][1]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: registers -> smem (fp32, padded rows) -> fused math -> coalesced global stores
  constexpr int LDC = BN + 8;
  float* sC = reinterpret_cast<float*>(smem_raw);
  {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int row = warp_m * 32 + mi * 16 + g;
        const int col = warp_n * (BN / 2) + ni * 8 + 2 * q;
        *reinterpret_cast<float2*>(sC + row * LDC + col) = make_float2(acc[mi][ni][0], acc[mi][ni][1]);
        *reinterpret_cast<float2*>(sC + (row + 8) * LDC + col) = make_float2(acc[mi][ni][2], acc[mi][ni][3]);
      }
  }
  __syncthreads();
  constexpr int GROUPS = BN / 8;
  for (int u = tid; u < BM * GROUPS; u += THREADS) {
    const int row = u / GROUPS, cg = u - row * GROUPS;
    const long long m = m0 + row;
    const int c0 = n0 + cg * 8;
    if (m >= p.M || c0 >= p.Cout) continue;
    const float4 lo = *reinterpret_cast<const float4*>(sC + row * LDC + cg * 8);
    const float4 hi = *reinterpret_cast<const float4*>(sC + row * LDC + cg * 8 + 4);
    const float a8[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    conv_epilogue8(p, m, c0, a8);
  }
}

template <int BN>
int launch(const ConvParams& p, cudaStream_t stream) {
  constexpr int pipe = STAGES * (BM + BN) * BK * 2;
  constexpr int epi = BM * (BN + 8) * 4;
  constexpr int smem = pipe > epi ? pipe : epi;
  static bool configured = false;
  if (!configured) {
    DYF_CUDA_OK(cudaFuncSetAttribute(conv_mma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.Cout + BN - 1) / BN));
  const double flops = 2.0 * (double)p.M * p.Cout * p.KH * p.KW * p.Cin_real;
  const double bytes = 2.0 * ((double)p.rows * p.Hi * p.Wi * p.Cin + (double)p.M * p.Cout + (double)p.Cout * p.Kpad);
  ProfScope prof(stream, KC_CONV_MMA, flops, bytes);
  conv_mma_kernel<BN><<<grid, THREADS, smem, stream>>>(p);
  DYF_LAUNCH_OK("conv_mma_kernel");
  return 0;
}

}  // namespace

int launch_conv_mma(const ConvParams& p, cudaStream_t stream) {
  if (p.Cin % 8 != 0 || p.Kpad % BK != 0 || p.Kpad < p.K) {
    set_error("conv_mma: Cin must be a multiple of 8 and Kpad a multiple of 32");
    return -1;
  }
  return p.Cout <= 64 ? launch<64>(p, stream) : launch<128>(p, stream);
}

}  // namespace dyf
