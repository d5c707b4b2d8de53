// Spatial attention blocks of the SST backbone (reference: src/models/unet.py:43-52,187-191,209 and
// src/models/modules/attention.py:7-73): channel LayerNorm, linear attention (softmax over the head dimension for q,
// over the n = H*W positions for k) and the full bottleneck attention.  These are a few percent of the network's
// FLOPs (0.18 of 10.7 GF), so they run as fp32 SIMT kernels on bf16 NHWC activations; the 1x1 qkv / out
// projections around them go through the convolution kernels.
#include "aux.cuh"

namespace dyf {
namespace {

constexpr int DH = 32;  // dim_head (attention.py:8,52) -- also the warp width, which the kernels below exploit

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ bool keep1(const DropCfg& d, const DropRow& dr, uint64_t elem) {
  return (drop_keep_bits8(d, dr, elem & ~7ull) >> (elem & 7)) & 1u;
}

// ---------------------------------------------------------------------------------------------- channel LayerNorm
// y = (x - mean_c) * rsqrt(var_c + 1e-5) * g  per pixel (biased variance, gain only), then the dropout that sits on
// the input of the qkv projection.  LANES = C / 8 lanes per pixel, 8 channels (one 128-bit access, one Philox draw) per
// lane; the channel reduction is a butterfly inside the LANES-wide lane group.
template <int LANES>
__global__ void __launch_bounds__(256) channel_ln_kernel(const ChannelLNParams p) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pix = t / LANES;
  const int c0 = (int)(t % LANES) * 8;
  const bool active = pix < p.M;  // (inactive lanes still take part in the shuffles)
  float v[8];
  if (active) unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + (size_t)pix * p.C + c0)), v);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)p.C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q += d * d; }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)p.C + 1e-5f);
  if (!active) return;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.g + c0)), g1 = __ldg(reinterpret_cast<const float4*>(p.g + c0) + 1);
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * g[j];
  if (p.drop.thresh) {
    const int r = (int)(pix / p.HW);
    const uint32_t keep = drop_keep_bits8(p.drop, drop_row(p.drop, r, (uint64_t)p.HW * p.C),
                                          (uint64_t)(pix - (long long)r * p.HW) * p.C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * p.drop.scale : 0.f;
  }
  *reinterpret_cast<uint4*>(p.y + (size_t)pix * p.C + c0) = pack8(v);
}

// C = 512 (one warp per pixel, 16 channels per lane): the shape outside the shipped configs, kept simple
__global__ void __launch_bounds__(256) channel_ln_wide_kernel(const ChannelLNParams p) {
  constexpr int PER = 16;
  const int lane = threadIdx.x & 31;
  const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pix >= p.M) return;
  const act_t* x = p.x + (size_t)pix * p.C + lane * PER;
  float v[PER];
  unpack8(__ldg(reinterpret_cast<const uint4*>(x)), v);
  unpack8(__ldg(reinterpret_cast<const uint4*>(x) + 1), v + 8);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PER; ++j) s += v[j];
  const float mean = warp_sum(s) / (float)p.C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < PER; ++j) { const float d = v[j] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.C + 1e-5f);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[8 * half + j] - mean) * rstd * __ldg(p.g + lane * PER + 8 * half + j);
    if (p.drop.thresh) {
      const int r = (int)(pix / p.HW);
      const uint32_t keep = drop_keep_bits8(p.drop, drop_row(p.drop, r, (uint64_t)p.HW * p.C),
                                            (uint64_t)(pix - (long long)r * p.HW) * p.C + lane * PER + 8 * half);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ((keep >> j) & 1u) ? o[j] * p.drop.scale : 0.f;
    }
    reinterpret_cast<uint4*>(p.y + (size_t)pix * p.C + lane * PER)[half] = pack8(o);
  }
}

// ---------------------------------------------------------------------------------------------- linear attention
// qkv: [rows, n, 3*heads*DH] with q | k | v blocks, heads-major inside each block ("b (h c) x y -> b h c (x y)").
// Pass 1 (one block per (head, row)): ctx[d][e] = sum_n softmax_n(k)[d, n] * v[e, n] / n.
// Phase A: max over n of k[d, n].  Phase B: every warp walks its own chunks of 16 positions: stage exp(k - max) and v as
// fp32 in a private shared-memory tile (128-bit loads, 8 channels per lane), then accumulate a 4 (d) x 8 (e) register
// tile per lane from broadcast float4 reads (3 LDS per 32 FMA).  Phase C: the 8 per-warp partial sums are combined in
// warp order (fixed order => bit-reproducible).
constexpr int CTX_POS = 16;  // positions per warp chunk
__global__ void __launch_bounds__(256) linattn_ctx_kernel(const AttnParams p) {
  __shared__ __align__(16) float s_tile[8][2][CTX_POS][DH];  // per warp: exp(k - max) | v ; reused for the final combine
  __shared__ float s_red[8][DH];
  __shared__ float s_max[DH], s_den[8][DH];
  const int h = blockIdx.x, r = blockIdx.y;
  const int ld = 3 * p.heads * DH;
  const act_t* kbase = p.qkv + (size_t)r * p.n * ld + p.heads * DH + h * DH;
  const act_t* vbase = kbase + p.heads * DH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {  // ---- phase A
    float mx = -INFINITY;
    for (int n = warp; n < p.n; n += 8) mx = fmaxf(mx, act2f(kbase[(size_t)n * ld + lane]));
    s_red[warp][lane] = mx;
    __syncthreads();
    if (warp == 0) {
      float m = s_red[0][lane];
#pragma unroll
      for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i][lane]);
      s_max[lane] = m;
    }
    __syncthreads();
  }
  // ---- phase B
  const int sp = lane >> 2, sc = (lane & 3) * 8;  // staging: this lane loads channels [sc, sc+8) of positions sp, sp+8
  float kmax[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) kmax[j] = s_max[sc + j];
  const int dg = (lane >> 2) * 4, eg = (lane & 3) * 8;  // accumulation tile: d in [dg, dg+4), e in [eg, eg+8)
  float acc[4][8], den[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
  float (*tk)[DH] = s_tile[warp][0];
  float (*tv)[DH] = s_tile[warp][1];
  for (int n0 = warp * CTX_POS; n0 < p.n; n0 += 8 * CTX_POS) {
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = sp + 8 * half, n = n0 + i;
      float fk[8], fv[8];
      if (n < p.n) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(kbase + (size_t)n * ld + sc)), fk);
        unpack8(__ldg(reinterpret_cast<const uint4*>(vbase + (size_t)n * ld + sc)), fv);
#pragma unroll
        for (int j = 0; j < 8; ++j) fk[j] = __expf(fk[j] - kmax[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { fk[j] = 0.f; fv[j] = 0.f; }
      }
      *reinterpret_cast<float4*>(&tk[i][sc]) = make_float4(fk[0], fk[1], fk[2], fk[3]);
      *reinterpret_cast<float4*>(&tk[i][sc + 4]) = make_float4(fk[4], fk[5], fk[6], fk[7]);
      *reinterpret_cast<float4*>(&tv[i][sc]) = make_float4(fv[0], fv[1], fv[2], fv[3]);
      *reinterpret_cast<float4*>(&tv[i][sc + 4]) = make_float4(fv[4], fv[5], fv[6], fv[7]);
    }
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < CTX_POS; ++i) {
      const float4 kd = *reinterpret_cast<const float4*>(&tk[i][dg]);
      const float4 v0 = *reinterpret_cast<const float4*>(&tv[i][eg]), v1 = *reinterpret_cast<const float4*>(&tv[i][eg + 4]);
      const float ka[4] = {kd.x, kd.y, kd.z, kd.w}, va[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        den[a] += ka[a];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[a][e] = fmaf(ka[a], va[e], acc[a][e]);
      }
    }
  }
  // ---- phase C: per-warp partials -> shared memory (the warp's own tile: 2 * 16 * 32 floats = one 32 x 32 matrix)
  __syncwarp();
  float* part = &s_tile[warp][0][0][0];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int e = 0; e < 8; ++e) part[(dg + a) * DH + eg + e] = acc[a][e];
    if ((lane & 3) == 0) s_den[warp][dg + a] = den[a];
  }
  __syncthreads();
  float* ctx = p.ctx + ((size_t)r * p.heads + h) * DH * DH;
  for (int idx = threadIdx.x; idx < DH * DH; idx += blockDim.x) {
    const int d = idx / DH;
    float sum = 0.f, dsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { sum += (&s_tile[w][0][0][0])[idx]; dsum += s_den[w][d]; }
    ctx[idx] = sum / (dsum * (float)p.n);  // softmax normalisation and the v / (h*w) rescale (attention.py:42)
  }
}

// Pass 2: out[e, n] = sum_d ctx[d][e] * softmax_d(q)[d, n] * DH^-0.5.  One THREAD per position: its 32 q values (64
// contiguous bytes) are soft-maxed in registers, the 32 x 32 context matrix is read from shared memory as broadcast
// float4 rows, and the 32 outputs leave as four 128-bit stores -- no cross-lane traffic at all.
__global__ void __launch_bounds__(256) linattn_out_kernel(const AttnParams p) {
  __shared__ __align__(16) float s_ctx[DH][DH];
  const int h = blockIdx.y, r = blockIdx.z;
  const float* ctx = p.ctx + ((size_t)r * p.heads + h) * DH * DH;
  for (int i = threadIdx.x; i < DH * DH; i += blockDim.x) s_ctx[i / DH][i % DH] = ctx[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.n) return;
  const int ld = 3 * p.heads * DH;
  const uint4* qp = reinterpret_cast<const uint4*>(p.qkv + ((size_t)r * p.n + n) * ld + h * DH);
  float q[DH];
#pragma unroll
  for (int i = 0; i < 4; ++i) unpack8(__ldg(qp + i), q + 8 * i);
  float m = q[0];
#pragma unroll
  for (int d = 1; d < DH; ++d) m = fmaxf(m, q[d]);
  float sum = 0.f;
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = __expf(q[d] - m); sum += q[d]; }
  const float norm = rsqrtf((float)DH) / sum;
  float o[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) o[e] = 0.f;
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    const float qs = q[d] * norm;
#pragma unroll
    for (int e4 = 0; e4 < DH / 4; ++e4) {
      const float4 c = *reinterpret_cast<const float4*>(&s_ctx[d][4 * e4]);
      o[4 * e4 + 0] = fmaf(c.x, qs, o[4 * e4 + 0]); o[4 * e4 + 1] = fmaf(c.y, qs, o[4 * e4 + 1]);
      o[4 * e4 + 2] = fmaf(c.z, qs, o[4 * e4 + 2]); o[4 * e4 + 3] = fmaf(c.w, qs, o[4 * e4 + 3]);
    }
  }
  uint4* op = reinterpret_cast<uint4*>(p.out + ((size_t)r * p.n + n) * (p.heads * DH) + h * DH);
#pragma unroll
  for (int i = 0; i < 4; ++i) op[i] = pack8(o + 8 * i);
}

// ---------------------------------------------------------------------------------------------- full attention
// One block per (head, row): K^T and V of the head live in shared memory (n <= 1024 positions).  Each warp handles FOUR
// queries at a time, so that every shared-memory read of K^T / V feeds four FMAs: scores (lane = key position; the four
// scaled queries are read as one broadcast float4 per channel), softmax, dropout on the probabilities, then P.V (lane =
// channel; the four probabilities of a key are one broadcast float4).
constexpr int AQ = 4;  // queries per warp pass
__global__ void __launch_bounds__(256) attention_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float s_mem[];
  const int n = p.n, np = n | 1;
  float* s_kt = s_mem;                       // [DH][np]
  float* s_v = s_mem + DH * np;              // [n][DH]
  float* s_p = s_v + n * DH;                 // [8 warps][n][AQ]   (scores, then probabilities)
  float* s_q = s_p + 8 * n * AQ;             // [8 warps][DH][AQ]  (scaled queries)
  const int h = blockIdx.x, r = blockIdx.y;
  const int ld = 3 * p.heads * DH;
  const act_t* base = p.qkv + (size_t)r * n * ld + h * DH;
  for (int i = threadIdx.x; i < n * DH; i += blockDim.x) {
    const int j = i / DH, d = i - j * DH;
    s_kt[d * np + j] = act2f(base[(size_t)j * ld + p.heads * DH + d]);
    s_v[j * DH + d] = act2f(base[(size_t)j * ld + 2 * p.heads * DH + d]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = rsqrtf((float)DH);
  float4* pw = reinterpret_cast<float4*>(s_p + (size_t)warp * n * AQ);
  float4* qw = reinterpret_cast<float4*>(s_q + (size_t)warp * DH * AQ);
  for (int i0 = warp * AQ; i0 < n; i0 += 8 * AQ) {
    {  // the four queries, lane = channel
      float q[AQ];
#pragma unroll
      for (int t = 0; t < AQ; ++t) q[t] = i0 + t < n ? act2f(base[(size_t)(i0 + t) * ld + lane]) * scale : 0.f;
      qw[lane] = make_float4(q[0], q[1], q[2], q[3]);
    }
    __syncwarp();
    float mx[AQ] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int j = lane; j < n; j += 32) {  // lane = key position
      float sc[AQ] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        const float k = s_kt[d * np + j];
        const float4 q = qw[d];
        sc[0] = fmaf(q.x, k, sc[0]); sc[1] = fmaf(q.y, k, sc[1]); sc[2] = fmaf(q.z, k, sc[2]); sc[3] = fmaf(q.w, k, sc[3]);
      }
      pw[j] = make_float4(sc[0], sc[1], sc[2], sc[3]);
#pragma unroll
      for (int t = 0; t < AQ; ++t) mx[t] = fmaxf(mx[t], sc[t]);
    }
    float sum[AQ] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < AQ; ++t) mx[t] = warp_max(mx[t]);
    for (int j = lane; j < n; j += 32) {
      float4 e = pw[j];
      e.x = __expf(e.x - mx[0]); e.y = __expf(e.y - mx[1]); e.z = __expf(e.z - mx[2]); e.w = __expf(e.w - mx[3]);
      sum[0] += e.x; sum[1] += e.y; sum[2] += e.z; sum[3] += e.w;
      if (p.drop.thresh) {  // dropout on the probabilities (attention.py:59,70); element = ((row, head, query), key)
        const DropRow dr = drop_row(p.drop, r, (((uint64_t)p.heads * n * n) + 7) & ~7ull);
        const uint64_t e0 = ((uint64_t)h * n + i0) * n + j;
        e.x = keep1(p.drop, dr, e0) ? e.x * p.drop.scale : 0.f;
        e.y = keep1(p.drop, dr, e0 + n) ? e.y * p.drop.scale : 0.f;
        e.z = keep1(p.drop, dr, e0 + 2 * (uint64_t)n) ? e.z * p.drop.scale : 0.f;
        e.w = keep1(p.drop, dr, e0 + 3 * (uint64_t)n) ? e.w * p.drop.scale : 0.f;
      }
      pw[j] = e;
    }
    float inv[AQ];
#pragma unroll
    for (int t = 0; t < AQ; ++t) inv[t] = 1.f / warp_sum(sum[t]);
    __syncwarp();
    float o[AQ] = {0.f, 0.f, 0.f, 0.f};  // lane = channel
    for (int j = 0; j < n; ++j) {
      const float v = s_v[j * DH + lane];
      const float4 pr = pw[j];
      o[0] = fmaf(pr.x, v, o[0]); o[1] = fmaf(pr.y, v, o[1]); o[2] = fmaf(pr.z, v, o[2]); o[3] = fmaf(pr.w, v, o[3]);
    }
#pragma unroll
    for (int t = 0; t < AQ; ++t)
      if (i0 + t < n) p.out[((size_t)r * n + i0 + t) * (p.heads * DH) + h * DH + lane] = f2act(o[t] * inv[t]);
    __syncwarp();
  }
}

}  // namespace

int launch_channel_ln(const ChannelLNParams& p, cudaStream_t s) {
  ProfScope prof(s, KC_ATTENTION);
  const int lanes = p.C / 8;
  const int grid = cdiv(p.M * lanes, 256);
  switch (p.C) {
    case 64: channel_ln_kernel<8><<<grid, 256, 0, s>>>(p); break;
    case 128: channel_ln_kernel<16><<<grid, 256, 0, s>>>(p); break;
    case 256: channel_ln_kernel<32><<<grid, 256, 0, s>>>(p); break;
    case 512: channel_ln_wide_kernel<<<cdiv(p.M * 32, 256), 256, 0, s>>>(p); break;
    default: set_error("channel LayerNorm: channel count must be 64, 128, 256 or 512"); return -1;
  }
  DYF_LAUNCH_OK("channel_ln_kernel");
  return 0;
}

int launch_linear_attention(const AttnParams& p, cudaStream_t s) {
  ProfScope prof(s, KC_ATTENTION, 4.0 * p.rows * p.heads * (double)p.n * DH * DH);
  linattn_ctx_kernel<<<dim3(p.heads, p.rows), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("linattn_ctx_kernel");
  linattn_out_kernel<<<dim3(cdiv(p.n, 256), p.heads, p.rows), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("linattn_out_kernel");
  return 0;
}

int launch_attention_mma(const AttnParams& p, cudaStream_t s);  // attn_fused.cu

int launch_attention(const AttnParams& p, cudaStream_t s) {
  {  // tensor-core kernel for bottleneck grids (n <= 256); larger grids keep the shared-memory SIMT kernel below
    const int rc = launch_attention_mma(p, s);
    if (rc != 0) return rc < 0 ? rc : 0;
  }
  if (p.n > 576) {
    set_error("full attention is built for bottleneck grids (n <= 576 positions = 24 x 24, K/V/P resident in shared memory); "
              "keep_spatial_dims on large grids needs the streaming-softmax variant");
    return -4;
  }
  const size_t smem = ((size_t)DH * (p.n | 1) + (size_t)p.n * DH + 8 * (size_t)p.n * AQ + 8 * DH * AQ) * sizeof(float) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    DYF_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  ProfScope prof(s, KC_ATTENTION, 4.0 * p.rows * p.heads * (double)p.n * p.n * DH);
  attention_kernel<<<dim3(p.heads, p.rows), 256, smem, s>>>(p);
  DYF_LAUNCH_OK("attention_kernel");
  return 0;
}

}  // namespace dyf
