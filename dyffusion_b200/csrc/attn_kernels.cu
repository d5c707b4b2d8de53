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
__device__ __forceinline__ bool keep1(const DropCfg& d, uint64_t elem) {
  return (drop_keep_bits8(d, elem & ~7ull) >> (elem & 7)) & 1u;
}

// ---------------------------------------------------------------------------------------------- channel LayerNorm
// y = (x - mean_c) * rsqrt(var_c + 1e-5) * g  per pixel (biased variance, gain only), then the dropout that sits on
// the input of the qkv projection.  One warp per pixel, C/32 channels per lane.
template <int PER>
__global__ void __launch_bounds__(256) channel_ln_kernel(const ChannelLNParams p) {
  const int lane = threadIdx.x & 31;
  const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pix >= p.M) return;
  const __nv_bfloat16* x = p.x + (size_t)pix * p.C + lane * PER;
  float v[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) v[j] = __bfloat162float(x[j]);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PER; ++j) s += v[j];
  const float mean = warp_sum(s) / (float)p.C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < PER; ++j) { const float d = v[j] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.C + 1e-5f);
  __nv_bfloat16* y = p.y + (size_t)pix * p.C + lane * PER;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    float o = (v[j] - mean) * rstd * __ldg(p.g + lane * PER + j);
    if (p.drop.thresh) o = keep1(p.drop, (uint64_t)pix * p.C + lane * PER + j) ? o * p.drop.scale : 0.f;
    y[j] = __float2bfloat16_rn(o);
  }
}

// ---------------------------------------------------------------------------------------------- linear attention
// qkv: [rows, n, 3*heads*DH] with q | k | v blocks, heads-major inside each block ("b (h c) x y -> b h c (x y)").
// Pass 1 (one block per (head, row)): ctx[d][e] = sum_n softmax_n(k)[d, n] * v[e, n] / n.
__global__ void __launch_bounds__(256) linattn_ctx_kernel(const AttnParams p) {
  __shared__ float s_red[8][DH];
  __shared__ float s_max[DH], s_sum[DH];
  __shared__ float s_k[32][DH + 1], s_v[32][DH + 1];
  const int h = blockIdx.x, r = blockIdx.y;
  const int ld = 3 * p.heads * DH;
  const __nv_bfloat16* kbase = p.qkv + (size_t)r * p.n * ld + p.heads * DH + h * DH;
  const __nv_bfloat16* vbase = kbase + p.heads * DH;
  const int d = threadIdx.x & 31, g = threadIdx.x >> 5;  // 8 groups of 32 threads
  // max over n of k[d, n]
  float mx = -INFINITY;
  for (int n = g; n < p.n; n += 8) mx = fmaxf(mx, __bfloat162float(kbase[(size_t)n * ld + d]));
  s_red[g][d] = mx;
  __syncthreads();
  if (g == 0) {
    float m = s_red[0][d];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i][d]);
    s_max[d] = m;
  }
  __syncthreads();
  const float kmax = s_max[d];
  // thread (d, g) accumulates ctx[d][4g .. 4g+3] and (g == 0) the softmax denominator of row d
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, den = 0.f;
  for (int n0 = 0; n0 < p.n; n0 += 32) {
    __syncthreads();
    for (int i = g; i < 32; i += 8) {  // stage 32 positions: exp(k - max) and v
      const int n = n0 + i;
      const bool ok = n < p.n;
      s_k[i][d] = ok ? __expf(__bfloat162float(kbase[(size_t)n * ld + d]) - s_max[d]) : 0.f;
      s_v[i][d] = ok ? __bfloat162float(vbase[(size_t)n * ld + d]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const float ek = s_k[i][d];
      den += ek;
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = fmaf(ek, s_v[i][4 * g + e], acc[e]);
    }
  }
  (void)kmax;
  if (g == 0) s_sum[d] = den;
  __syncthreads();
  const float inv = 1.f / (s_sum[d] * (float)p.n);  // softmax normalisation and the v / (h*w) rescale (attention.py:42)
  float* ctx = p.ctx + (((size_t)r * p.heads + h) * DH + d) * DH + 4 * g;
#pragma unroll
  for (int e = 0; e < 4; ++e) ctx[e] = acc[e] * inv;
}

// Pass 2: out[e, n] = sum_d ctx[d][e] * softmax_d(q)[d, n] * DH^-0.5 ; one warp per position, lane = d then e.
__global__ void __launch_bounds__(256) linattn_out_kernel(const AttnParams p) {
  __shared__ float s_ctx[DH][DH + 1];
  const int h = blockIdx.y, r = blockIdx.z;
  const float* ctx = p.ctx + ((size_t)r * p.heads + h) * DH * DH;
  for (int i = threadIdx.x; i < DH * DH; i += blockDim.x) s_ctx[i / DH][i % DH] = ctx[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = 3 * p.heads * DH;
  const float scale = rsqrtf((float)DH);
  for (int n = blockIdx.x * 8 + warp; n < p.n; n += gridDim.x * 8) {
    const float q = __bfloat162float(p.qkv[((size_t)r * p.n + n) * ld + h * DH + lane]);
    const float m = warp_max(q);
    const float e = __expf(q - m);
    const float qs = e / warp_sum(e) * scale;
    float o = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) o = fmaf(s_ctx[d][lane], __shfl_sync(0xffffffffu, qs, d), o);
    p.out[((size_t)r * p.n + n) * (p.heads * DH) + h * DH + lane] = __float2bfloat16_rn(o);
  }
}

// ---------------------------------------------------------------------------------------------- full attention
// One block per (head, row): K^T and V of the head live in shared memory (n <= 1024 positions), each warp walks over
// query positions: scores (lane = key position), softmax, dropout on the probabilities, then P.V (lane = channel).
__global__ void __launch_bounds__(256) attention_kernel(const AttnParams p) {
  extern __shared__ float s_mem[];
  const int n = p.n, np = n | 1;
  float* s_kt = s_mem;            // [DH][np]
  float* s_v = s_mem + DH * np;   // [n][DH]
  float* s_p = s_v + n * DH;      // [8 warps][n]
  const int h = blockIdx.x, r = blockIdx.y;
  const int ld = 3 * p.heads * DH;
  const __nv_bfloat16* base = p.qkv + (size_t)r * n * ld + h * DH;
  for (int i = threadIdx.x; i < n * DH; i += blockDim.x) {
    const int j = i / DH, d = i - j * DH;
    s_kt[d * np + j] = __bfloat162float(base[(size_t)j * ld + p.heads * DH + d]);
    s_v[j * DH + d] = __bfloat162float(base[(size_t)j * ld + 2 * p.heads * DH + d]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = rsqrtf((float)DH);
  float* pw = s_p + warp * n;
  for (int i = warp; i < n; i += 8) {
    const float q = __bfloat162float(base[(size_t)i * ld + lane]) * scale;  // lane = d
    float qv[DH];  // every lane needs the whole query vector: broadcast it once (all lanes participate)
#pragma unroll
    for (int d = 0; d < DH; ++d) qv[d] = __shfl_sync(0xffffffffu, q, d);
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) {  // lane = key position
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) sc = fmaf(qv[d], s_kt[d * np + j], sc);
      pw[j] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float e = __expf(pw[j] - mx);
      pw[j] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    if (p.drop.thresh) {
      const uint64_t e0 = (((uint64_t)r * p.heads + h) * n + i) * n;
      for (int j = lane; j < n; j += 32) pw[j] = keep1(p.drop, e0 + j) ? pw[j] * p.drop.scale : 0.f;
    }
    __syncwarp();
    float o = 0.f;  // lane = d
    for (int j = 0; j < n; ++j) o = fmaf(pw[j], s_v[j * DH + lane], o);
    p.out[((size_t)r * n + i) * (p.heads * DH) + h * DH + lane] = __float2bfloat16_rn(o * inv);
    __syncwarp();
  }
}

}  // namespace

int launch_channel_ln(const ChannelLNParams& p, cudaStream_t s) {
  ProfScope prof(s, KC_ATTENTION);
  const int grid = cdiv(p.M * 32, 256);
  switch (p.C) {
    case 64: channel_ln_kernel<2><<<grid, 256, 0, s>>>(p); break;
    case 128: channel_ln_kernel<4><<<grid, 256, 0, s>>>(p); break;
    case 256: channel_ln_kernel<8><<<grid, 256, 0, s>>>(p); break;
    case 512: channel_ln_kernel<16><<<grid, 256, 0, s>>>(p); break;
    default: set_error("channel LayerNorm: channel count must be 64, 128, 256 or 512"); return -1;
  }
  DYF_LAUNCH_OK("channel_ln_kernel");
  return 0;
}

int launch_linear_attention(const AttnParams& p, cudaStream_t s) {
  ProfScope prof(s, KC_ATTENTION, 4.0 * p.rows * p.heads * (double)p.n * DH * DH);
  linattn_ctx_kernel<<<dim3(p.heads, p.rows), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("linattn_ctx_kernel");
  const int gx = cdiv(p.n, 8 * 8);
  linattn_out_kernel<<<dim3(gx, p.heads, p.rows), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("linattn_out_kernel");
  return 0;
}

int launch_attention(const AttnParams& p, cudaStream_t s) {
  if (p.n > 1024) {
    set_error("full attention is built for bottleneck grids (n <= 1024 positions); keep_spatial_dims on large grids "
              "needs the streaming-softmax variant");
    return -4;
  }
  const size_t smem = ((size_t)DH * (p.n | 1) + (size_t)p.n * DH + 8 * (size_t)p.n) * sizeof(float);
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
