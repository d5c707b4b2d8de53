// Memory-bound kernels around the convolutions (see aux.cuh).  All activations are bf16 NHWC with 128-bit accesses
// (8 channels per thread); network inputs/outputs are the reference's fp32 NCHW tensors.
#include <cooperative_groups.h>

#include <cstdlib>

#include "aux.cuh"

namespace dyf {
namespace {

// PyTorch area_pixel_compute_source_index for align_corners=False (bilinear resize).
__device__ __forceinline__ void bilinear_coord(int dst, int in_size, float scale, int& i0, int& i1, float& l1) {
  float src = scale * (dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

__device__ __forceinline__ float gauss_from(uint32_t a, uint32_t b) {
  float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
  float u2 = (float)b * 2.3283064365386963e-10f;
  return sqrtf(-2.f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}

// Movers index their work as (batch row, position inside the row) over grids of rows x ceil(per_row / 256) blocks, so
// the per-thread index arithmetic stays 32-bit: a 64-bit division by a run-time value is ~100 instructions per thread,
// which made these copy kernels issue-bound (pack_s2d: 535 instructions per 32 output bytes).
struct RowPos { unsigned r, i; };
__device__ __forceinline__ RowPos row_pos(unsigned per_row) {
  const unsigned bpr = (per_row + 255u) >> 8;
  const unsigned r = blockIdx.x / bpr;
  return {r, (blockIdx.x - r * bpr) * 256u + threadIdx.x};
}
static inline unsigned row_grid(long long rows, long long per_row) { return (unsigned)(rows * ((per_row + 255) / 256)); }

// ------------------------------------------------------------------------------------------------ pack
__global__ void __launch_bounds__(256) pack_kernel(const PackParams p) {
  const unsigned per_row = (unsigned)(p.Ho * p.Wo);
  const RowPos rp = row_pos(per_row);
  if (rp.i >= per_row) return;
  const int oy = (int)(rp.i / (unsigned)p.Wo), ox = (int)(rp.i - (unsigned)oy * p.Wo);
  const int r = (int)rp.r;
  const size_t idx = (size_t)rp.r * per_row + rp.i;
  int y0 = oy, y1 = oy, x0 = ox, x1 = ox;
  float ly = 0.f, lx = 0.f;
  if (p.bilinear) {
    bilinear_coord(oy, p.Hi, (float)p.Hi / (float)p.Ho, y0, y1, ly);
    bilinear_coord(ox, p.Wi, (float)p.Wi / (float)p.Wo, x0, x1, lx);
  }
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const size_t plane = (size_t)p.Hi * p.Wi;
  act_t* o = p.out + idx * p.Cpad;
  float v[8];
  int oc = 0;
  for (int s = 0; s < p.nsrc; ++s) {
    const float* base = p.src[s] + (size_t)(r % p.src_rows) * p.C[s] * plane;
    for (int c = 0; c < p.C[s]; ++c) {
      const float* pl = base + (size_t)c * plane;
      float val;
      if (p.bilinear) {
        val = w00 * __ldg(pl + (size_t)y0 * p.Wi + x0) + w01 * __ldg(pl + (size_t)y0 * p.Wi + x1) +
              w10 * __ldg(pl + (size_t)y1 * p.Wi + x0) + w11 * __ldg(pl + (size_t)y1 * p.Wi + x1);
      } else {
        val = __ldg(pl + (size_t)oy * p.Wi + ox);
        if (s == p.noise_src) {
          // per OUTPUT row: batch row r = (logical call r / rng_rows, row r % rng_rows + row_off of the un-sharded job)
          const uint32_t jc = (uint32_t)r / p.rng_rows, rr = (uint32_t)r - jc * p.rng_rows + p.row_off;
          const uint64_t e = ((uint64_t)rr * p.C[s] + c) * plane + (size_t)oy * p.Wi + ox;
          Philox ph(rng_seed(p.seed, p.seed_ptr));
          uint4 rn = ph((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)p.stream + jc, 0x4e4f4953u);
          val = p.noise_w * val + (1.f - p.noise_w) * gauss_from(rn.x, rn.y);
        }
      }
      v[oc & 7] = val;
      if ((++oc & 7) == 0) *reinterpret_cast<uint4*>(o + oc - 8) = pack8(v);
    }
  }
  while (oc < p.Cpad) {
    v[oc & 7] = oc == p.ones_channel ? 1.f : 0.f;
    if ((++oc & 7) == 0) *reinterpret_cast<uint4*>(o + oc - 8) = pack8(v);
  }
}

// ------------------------------------------------------------------------------------------------ pack, space-to-depth
// One thread per (block pixel, sub-position): 16 channel slots = C_in resized input channels, a ones slot, zeros.
// The concat is flattened on the host into per-channel plane pointers (S2dChannels), so the per-thread work is the four
// bilinear taps per channel and two 128-bit stores.
struct S2dChannels {
  const float* plane[16];  // channel plane of source row 0
  int row_stride[16];      // floats between consecutive source rows of that tensor
  int n;                   // real channels (slot n carries 1.0)
};
template <int N>  // N = real channels (compile time: the 4 N gathers of a thread are all in flight before the first use)
__global__ void __launch_bounds__(256) pack_s2d_kernel(const PackParams p, const S2dChannels ch) {
  const unsigned per_row = (unsigned)(p.Ho * p.Wo) * 4u;
  const RowPos rp = row_pos(per_row);
  if (rp.i >= per_row) return;
  const int sub = (int)(rp.i & 3u);
  const unsigned pb = rp.i >> 2;  // block pixel inside the row
  const int by = (int)(pb / (unsigned)p.Wo), bx = (int)(pb - (unsigned)by * p.Wo);
  const size_t blk = (size_t)rp.r * (per_row >> 2) + pb;
  const int oy = 2 * by + (sub >> 1), ox = 2 * bx + (sub & 1);  // pixel of the full (resized) grid
  const int Hf = 2 * p.Ho, Wf = 2 * p.Wo;
  int y0 = oy, y1 = oy, x0 = ox, x1 = ox;
  float ly = 0.f, lx = 0.f;
  if (p.bilinear) {
    bilinear_coord(oy, p.Hi, (float)p.Hi / (float)Hf, y0, y1, ly);
    bilinear_coord(ox, p.Wi, (float)p.Wi / (float)Wf, x0, x1, lx);
  }
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const int o00 = y0 * p.Wi + x0, o01 = y0 * p.Wi + x1, o10 = y1 * p.Wi + x0, o11 = y1 * p.Wi + x1;
  const unsigned rs = rp.r % (unsigned)p.src_rows;  // block-uniform
  float g[N][4];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float* pl = ch.plane[i] + (size_t)rs * ch.row_stride[i];
    g[i][0] = __ldg(pl + o00); g[i][1] = __ldg(pl + o01); g[i][2] = __ldg(pl + o10); g[i][3] = __ldg(pl + o11);
  }
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
    v[i] = i < N ? (w00 * g[i][0] + w01 * g[i][1] + w10 * g[i][2] + w11 * g[i][3]) : i == N ? 1.f : 0.f;
  uint4* o = reinterpret_cast<uint4*>(p.out + blk * 64 + sub * 16);
  st_global_256(o, pack8(v), pack8(v + 8));  // 32 bytes per thread, 32-byte aligned
}

// ------------------------------------------------------------------------------------------------ pack, x-im2col
// fp32 NCHW sources -> 16-bit NHWC [rows, H, W, 64] whose channel kx * 8 + c holds source channel c of pixel (y, x + kx - 3)
// (zero outside the image, c >= C_in, kx = 7): the horizontal taps of a 7x7 stem live in the channel axis, so the conv is
// seven VERTICAL taps over 64 channels on the tcgen05 halo-patch kernel (conv_umma.cu, S1K7V).  "data+noise" conditioning is
// applied per source element with the element-keyed Philox stream of pack_kernel (the seven copies of a pixel agree).
// One CTA per image row: every source pixel is gathered, noised and packed to 8 x 16 bit ONCE into shared memory (with the
// three zero pixels of padding on either side); an output pixel is then seven 128-bit copies of its neighbours + one of zeros.
__global__ void __launch_bounds__(128) pack_xim2col_kernel(const PackParams p, const S2dChannels ch) {
  extern __shared__ __align__(16) uint4 s_px[];  // [Wi + 6]
  const unsigned y = blockIdx.x % (unsigned)p.Ho, r = blockIdx.x / (unsigned)p.Ho;
  const unsigned rs = r % (unsigned)p.src_rows;
  const size_t plane = (size_t)p.Hi * p.Wi;
  for (int x = threadIdx.x; x < p.Wi + 6; x += blockDim.x) {
    const int xs = x - 3;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float val = 0.f;
      if (c < ch.n && xs >= 0 && xs < p.Wi) {
        val = __ldg(ch.plane[c] + (size_t)rs * ch.row_stride[c] + (size_t)y * p.Wi + xs);
        if (c >= p.noise_c0 && c < p.noise_c1) {
          const uint32_t jc = r / p.rng_rows, rr = r - jc * p.rng_rows + p.row_off;
          const uint64_t e = ((uint64_t)rr * (p.noise_c1 - p.noise_c0) + (c - p.noise_c0)) * plane + (size_t)y * p.Wi + xs;
          Philox ph(rng_seed(p.seed, p.seed_ptr));
          uint4 rn = ph((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)p.stream + jc, 0x4e4f4953u);
          val = p.noise_w * val + (1.f - p.noise_w) * gauss_from(rn.x, rn.y);
        }
      }
      v[c] = val;
    }
    s_px[x] = pack8(v);
  }
  __syncthreads();
  uint4* const orow = reinterpret_cast<uint4*>(p.out + ((size_t)blockIdx.x * p.Wo) * 64);
  for (int i = threadIdx.x; i < p.Wo * 8; i += blockDim.x) {  // 16-byte unit i of the row: pixel i / 8, tap i % 8 (coalesced)
    const int x = i >> 3, kx = i & 7;
    orow[i] = kx < 7 ? s_px[x + kx] : make_uint4(0u, 0u, 0u, 0u);
  }
}

// 7x7 stem filter fp32 [O][C][7][7] -> [O][64][7]: (kx * 8 + c, ky), zero for the unused slots (the weight of the conv over
// the x-im2col'd input)
__global__ void stem_xim2col_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= O * 64 * 7) return;
  const int ky = idx % 7, ci = (idx / 7) % 64, o = idx / (7 * 64);
  const int kx = ci >> 3, c = ci & 7;
  out[idx] = (kx < 7 && c < C) ? w[(((size_t)o * C + c) * 7 + ky) * 7 + kx] : 0.f;
}

// 1x1 head with a handful of output channels (SST final_conv 64 -> 1): y[row][oc][pixel] = w[oc] . x[row][pixel] + b[oc],
// fp32 NCHW out.  One thread per pixel: a 128-bit load per 8 channels, weights broadcast from shared memory.
__global__ void __launch_bounds__(256) head1x1_kernel(const Head1x1Params p) {
  extern __shared__ float s_w[];  // [OC][C] + [OC]
  for (int i = threadIdx.x; i < p.OC * p.C; i += blockDim.x) s_w[i] = p.w[i];
  for (int i = threadIdx.x; i < p.OC; i += blockDim.x) s_w[p.OC * p.C + i] = p.bias ? p.bias[i] : 0.f;
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = o < p.OC ? s_w[p.OC * p.C + o] : 0.f;
  const uint4* xp = reinterpret_cast<const uint4*>(p.x + (size_t)m * p.C);
  for (int c8 = 0; c8 < p.C / 8; c8 += 2) {  // (C % 16 == 0, checked by the launcher: 256-bit loads)
    uint4 u0, u1;
    ld_global_nc_256(xp + c8, u0, u1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float f[8];
      unpack8(h ? u1 : u0, f);
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o < p.OC) {
          const float* w = s_w + o * p.C + (c8 + h) * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[o] = fmaf(f[j], w[j], acc[o]);
        }
    }
  }
  const long long r = m / p.HW, pix = m - r * p.HW;
  for (int o = 0; o < p.OC; ++o) p.y[((size_t)r * p.OC + o) * p.HW + pix] = acc[o];
}

// ------------------------------------------------------------------------------------------------ fused stem
// One warp = 32 consecutive output pixels.  Phase 1: lane p gathers (bilinear) the C_in input values of pixel p --
// neighbouring lanes read neighbouring source addresses of the same channel plane, so the gathers coalesce.
// Phase 2: lane p computes the 64 output channels of its pixel (fp32), the warp stages them in shared memory and
// writes the 32 x 128-byte pixel lines with fully coalesced 128-bit stores.
__global__ void __launch_bounds__(256) stem_kernel(const StemParams p) {
  __shared__ __align__(16) float s_w[16 * 64];      // [c][o]
  __shared__ float s_b[64];
  __shared__ __align__(16) uint4 s_out[8][32 * 9];  // per warp: 32 pixels x 8 chunks (+1 pad)
  for (int i = threadIdx.x; i < p.Cin * 64; i += blockDim.x) s_w[i] = p.w[(size_t)(i & 63) * p.Cin + (i >> 6)];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_b[i] = p.bias[i];
  __syncthreads();
  const PackParams& k = p.pk;
  const long long total = (long long)k.rows * k.Ho * k.Wo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pix0 = ((long long)blockIdx.x * 8 + warp) * 32;
  const long long pix = pix0 + lane;
  const bool active = pix < total;
  const long long pp = active ? pix : 0;
  const int ox = (int)(pp % k.Wo);
  const int oy = (int)((pp / k.Wo) % k.Ho);
  const int r = (int)(pp / ((long long)k.Wo * k.Ho));
  int y0 = oy, y1 = oy, x0 = ox, x1 = ox;
  float ly = 0.f, lx = 0.f;
  if (k.bilinear) {
    bilinear_coord(oy, k.Hi, (float)k.Hi / (float)k.Ho, y0, y1, ly);
    bilinear_coord(ox, k.Wi, (float)k.Wi / (float)k.Wo, x0, x1, lx);
  }
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const size_t plane = (size_t)k.Hi * k.Wi;
  const size_t o00 = (size_t)y0 * k.Wi + x0, o01 = (size_t)y0 * k.Wi + x1, o10 = (size_t)y1 * k.Wi + x0,
               o11 = (size_t)y1 * k.Wi + x1;
  float acc[64];
#pragma unroll
  for (int o = 0; o < 64; ++o) acc[o] = s_b[o];
  int c = 0;
  for (int s = 0; s < k.nsrc; ++s) {
    const float* base = k.src[s] + (size_t)(r % k.src_rows) * k.C[s] * plane;
    for (int cc = 0; cc < k.C[s]; ++cc, ++c) {
      const float* pl = base + (size_t)cc * plane;
      const float v = k.bilinear ? w00 * __ldg(pl + o00) + w01 * __ldg(pl + o01) + w10 * __ldg(pl + o10) + w11 * __ldg(pl + o11)
                                 : __ldg(pl + o00);
      const float4* wr = reinterpret_cast<const float4*>(s_w + c * 64);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 w4 = wr[q];
        acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    if (p.drop.thresh) {
      const uint32_t keep = drop_keep_bits8(p.drop, drop_row(p.drop, r, (uint64_t)k.Ho * k.Wo * 64),
                                            (uint64_t)(oy * k.Wo + ox) * 64 + g * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[g * 8 + e] = ((keep >> e) & 1u) ? acc[g * 8 + e] * p.drop.scale : 0.f;
    }
    s_out[warp][lane * 9 + g] = pack8(acc + g * 8);
  }
  __syncwarp();
  // coalesced write-out: 32 pixels x 128 B are contiguous in the NHWC output
  uint4* out = reinterpret_cast<uint4*>(p.out + (size_t)pix0 * 64);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int q = it * 32 + lane;  // 16-byte unit inside the warp's 4 KB span
    const int pl_ = q >> 3, g = q & 7;
    if (pix0 + pl_ < total) out[q] = s_out[warp][pl_ * 9 + g];
  }
}

// ------------------------------------------------------------------------------------------------ upsample x2 + concat
__global__ void __launch_bounds__(256) upsample_kernel(const UpsampleParams p) {
  const int Ho = p.H * p.scale, Wo = p.W * p.scale;
  const int Ct = p.C[0] + p.C[1];
  const int chunks = Ct >> 3;
  const unsigned per_row = (unsigned)(Ho * Wo * chunks);
  const RowPos rp = row_pos(per_row);
  if (rp.i >= per_row) return;
  const unsigned pr = rp.i / (unsigned)chunks;  // pixel inside the row
  const int ch = (int)(rp.i - pr * chunks);
  const int oy = (int)(pr / (unsigned)Wo), ox = (int)(pr - (unsigned)oy * Wo);
  const int r = (int)rp.r;
  const size_t pix = (size_t)rp.r * ((size_t)Ho * Wo) + pr;
  const int c = ch << 3;
  const int s = c < p.C[0] ? 0 : 1;
  const int cs = s ? c - p.C[0] : c;
  const act_t* src = p.src[s] + (size_t)r * p.H * p.W * p.ld[s] + cs;
  uint4 outv;
  if (p.scale == 1) {
    outv = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)oy * p.W + ox) * p.ld[s]));
  } else if (!p.bilinear) {
    outv = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)(oy >> 1) * p.W + (ox >> 1)) * p.ld[s]));
  } else {
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_coord(oy, p.H, 0.5f, y0, y1, ly);
    bilinear_coord(ox, p.W, 0.5f, x0, x1, lx);
    float a[8], b[8], cc[8], d[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + ((size_t)y0 * p.W + x0) * p.ld[s])), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + ((size_t)y0 * p.W + x1) * p.ld[s])), b);
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + ((size_t)y1 * p.W + x0) * p.ld[s])), cc);
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + ((size_t)y1 * p.W + x1) * p.ld[s])), d);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = w00 * a[j] + w01 * b[j] + w10 * cc[j] + w11 * d[j];
    outv = pack8(o);
  }
  *reinterpret_cast<uint4*>(p.out + pix * Ct + c) = outv;
}

// Bilinear x2 (align_corners=False) as a fixed separable stencil: with source indices clamped to the image,
//   u[2i] = 0.25 x[i-1] + 0.75 x[i],   u[2i+1] = 0.75 x[i] + 0.25 x[i+1]
// (identical to PyTorch's coordinate rule, including the clamped first/last rows).  One thread turns the 3 x 3 source
// neighbourhood of pixel (i, j) into the 2 x 2 output quad for 8 channels: 9 independent 128-bit loads, 4 stores.
__global__ void __launch_bounds__(256) upsample2x_quad_kernel(const UpsampleParams p) {
  const int Ct = p.C[0] + p.C[1];
  const int chunks = Ct >> 3;
  const unsigned per_row = (unsigned)(p.H * p.W * chunks);
  const RowPos rp = row_pos(per_row);
  if (rp.i >= per_row) return;
  const unsigned pr = rp.i / (unsigned)chunks;  // source pixel inside the row
  const int ch = (int)(rp.i - pr * chunks);
  const int i = (int)(pr / (unsigned)p.W), j = (int)(pr - (unsigned)i * p.W);
  const int r = (int)rp.r;
  const int c = ch << 3;
  const int s = c < p.C[0] ? 0 : 1;
  const int cs = s ? c - p.C[0] : c;
  const int ld = p.ld[s];
  const act_t* src = p.src[s] + (size_t)r * p.H * p.W * ld + cs;
  const int ym = max(i - 1, 0), yp = min(i + 1, p.H - 1), xm = max(j - 1, 0), xp = min(j + 1, p.W - 1);
  const int ys[3] = {ym, i, yp};
  uint4 raw[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const act_t* rowp = src + (size_t)ys[a] * p.W * ld;
    raw[a][0] = __ldg(reinterpret_cast<const uint4*>(rowp + (size_t)xm * ld));
    raw[a][1] = __ldg(reinterpret_cast<const uint4*>(rowp + (size_t)j * ld));
    raw[a][2] = __ldg(reinterpret_cast<const uint4*>(rowp + (size_t)xp * ld));
  }
  float h0[3][8], h1[3][8];  // horizontally interpolated: columns 2j and 2j+1
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float L[8], C[8], R[8];
    unpack8(raw[a][0], L); unpack8(raw[a][1], C); unpack8(raw[a][2], R);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      h0[a][e] = 0.25f * L[e] + 0.75f * C[e];
      h1[a][e] = 0.75f * C[e] + 0.25f * R[e];
    }
  }
  float o[8];
  const int Wo = 2 * p.W;
  act_t* out = p.out + (((size_t)r * 2 * p.H + 2 * i) * Wo + 2 * j) * Ct + c;
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.25f * h0[0][e] + 0.75f * h0[1][e];
  *reinterpret_cast<uint4*>(out) = pack8(o);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.25f * h1[0][e] + 0.75f * h1[1][e];
  *reinterpret_cast<uint4*>(out + Ct) = pack8(o);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.75f * h0[1][e] + 0.25f * h0[2][e];
  *reinterpret_cast<uint4*>(out + (size_t)Wo * Ct) = pack8(o);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.75f * h1[1][e] + 0.25f * h1[2][e];
  *reinterpret_cast<uint4*>(out + (size_t)Wo * Ct + Ct) = pack8(o);
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// pass 1: per (row, group, pixel-slab) sum / sum of squares, reduced in a FIXED order (no float atomics), so that the
// statistics -- and with them every downstream bf16 rounding -- are bit-reproducible from run to run.
// Each thread owns one fixed 8-channel chunk => one fixed group.  Partials land in stats[row][group][slab][2].
constexpr int GN_SLABS = 32;
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const GroupNormParams p, int pix_per_block) {
  __shared__ float s_part[256][2];
  const int r = blockIdx.y;
  const int chunks = p.C >> 3;
  const int cpg = p.C / p.G;  // channels per group (multiple of 8)
  const int ch = threadIdx.x % chunks;
  const int pstep = blockDim.x / chunks;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, p.HW);
  float s1 = 0.f, s2 = 0.f;
  const act_t* x = p.x + (size_t)r * p.HW * p.C + (ch << 3);
  for (int px = p0 + threadIdx.x / chunks; px < p1; px += pstep) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + (size_t)px * p.C)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 += f[j] * f[j]; }
  }
  s_part[threadIdx.x][0] = s1;
  s_part[threadIdx.x][1] = s2;
  __syncthreads();
  // tree over the pixel lanes of every channel chunk (thread = (pixel lane, chunk), pstep = 256 / chunks is a power of two),
  // then thread g adds the chunks of group g in index order: a fixed order, so the statistics are bit-reproducible
  for (int half = pstep >> 1; half > 0; half >>= 1) {
    if ((int)threadIdx.x < half * chunks) {
      s_part[threadIdx.x][0] += s_part[threadIdx.x + half * chunks][0];
      s_part[threadIdx.x][1] += s_part[threadIdx.x + half * chunks][1];
    }
    __syncthreads();
  }
  if (threadIdx.x < p.G) {
    const int g = threadIdx.x, cpc = cpg >> 3;  // chunks per group
    float a = 0.f, b = 0.f;
    for (int c = g * cpc; c < (g + 1) * cpc; ++c) { a += s_part[c][0]; b += s_part[c][1]; }
    float* o = p.stats + (((size_t)r * p.G + g) * GN_SLABS + blockIdx.x) * 2;
    o[0] = a;
    o[1] = b;
  }
}

// pass 2 (tiny): per (row, channel) fold statistics, affine and the time scale/shift into y = x * A + B:
//   A = rstd * gamma * (scale + 1),  B = (beta - mean * rstd * gamma) * (scale + 1) + shift
// (slab partials are combined in a fixed order).  ab: [rows][2][C] behind the slab partials.
__global__ void __launch_bounds__(256) groupnorm_fold_kernel(const GroupNormParams p, float* __restrict__ ab) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int cpg = p.C / p.G, g = c / cpg;
    const float n = (float)p.HW * (float)cpg;
    float sum1 = 0.f, sum2 = 0.f;
    const float* st = p.stats + ((size_t)r * p.G + g) * GN_SLABS * 2;
    for (int sl = 0; sl < p.slabs; ++sl) { sum1 += st[2 * sl]; sum2 += st[2 * sl + 1]; }  // fixed order
    const float mean = sum1 / n;
    const float var = fmaxf(sum2 / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    float A = rstd * __ldg(p.gamma + c);
    float B = __ldg(p.beta + c) - mean * A;
    if (p.tabA) {
      const float tA = __ldg(p.tabA + (size_t)(r / p.tab_div) * p.C + c), tB = __ldg(p.tabB + (size_t)(r / p.tab_div) * p.C + c);
      A *= tA;
      B = B * tA + tB;
    }
    ab[((size_t)r * 2 + 0) * p.C + c] = A;
    ab[((size_t)r * 2 + 1) * p.C + c] = B;
  }
}

// pass 3: y = act(x * A + B) -> dropout -> + residual
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const GroupNormParams p, const float* __restrict__ ab) {
  const int chunks = p.C >> 3;
  const long long total = (long long)p.rows * p.HW * chunks;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % chunks);
  const long long m = idx / chunks;
  const int r = (int)(m / p.HW);
  const int c0 = ch << 3;
  const float4* A = reinterpret_cast<const float4*>(ab + ((size_t)r * 2 + 0) * p.C + c0);
  const float4* B = reinterpret_cast<const float4*>(ab + ((size_t)r * 2 + 1) * p.C + c0);
  const float4 a0 = __ldg(A), a1 = __ldg(A + 1), b0 = __ldg(B), b1 = __ldg(B + 1);
  const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float f[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + (size_t)m * p.C + c0)), f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = apply_act(fmaf(f[j], av[j], bv[j]), p.act);
  if (p.drop.thresh) {
    const uint32_t keep = drop_keep_bits8(p.drop, drop_row(p.drop, r, (uint64_t)p.HW * p.C),
                                          (uint64_t)(m - (long long)r * p.HW) * p.C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = ((keep >> j) & 1u) ? f[j] * p.drop.scale : 0.f;
  }
  if (p.res) {
    float rr[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p.res + (size_t)m * p.res_ld + c0)), rr);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += rr[j];
  }
  *reinterpret_cast<uint4*>(p.y + (size_t)m * p.C + c0) = pack8(f);
}

// One-kernel GroupNorm: a CLUSTER of CS CTAs per batch row, each walking its slice of the row twice -- first pass: per-thread
// sums of a fixed 8-channel chunk, reduced in a fixed order inside the CTA, then across the cluster through distributed
// shared memory (rank order: bit-reproducible); second pass: y = act(x * A + B) -> dropout -> + residual.  The slice
// (<= 64 KB) was just streamed by the same CTA, so the second read hits L2: DRAM traffic is one read and one write of the
// tensor instead of two reads and one write, and one launch replaces three.
constexpr int GNF_THREADS = 256, GNF_UNROLL = 4, GNF_SLABS = 8;
#ifndef GNF_MINB
#define GNF_MINB 8
#endif
#ifndef GNF_AU
#define GNF_AU 1
#endif
template <int AU, int MINB>  // AU = pixels in flight per thread in the apply pass, MINB = CTAs per SM the register budget allows
__global__ void __launch_bounds__(GNF_THREADS, MINB) groupnorm_fused_kernel(const GroupNormParams p, int CS, int slab_pix) {
  pdl_trigger();
  pdl_wait();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float s_part[GNF_THREADS][2];
  __shared__ float s_sum[GNF_SLABS][64][2];  // per-group partial sums of this CTA's slabs (read by the whole cluster)
  __shared__ float s_grp[64][2];             // mean, rstd
  extern __shared__ __align__(16) float s_ab[];  // [2][C]
  const int r = blockIdx.x / CS, rank = blockIdx.x - r * CS;
  const int chunks = p.C >> 3, cpg = p.C / p.G;
  const int ch = threadIdx.x % chunks, lane_px = threadIdx.x / chunks, pstep = GNF_THREADS / chunks;
  // The row is cut into GNF_SLABS fixed pixel slabs whatever the cluster size; a CTA owns GNF_SLABS / CS consecutive slabs,
  // reduces each one separately, and the row totals are the slab sums added in slab order: the statistics (and with them
  // every output bit) do not depend on how many CTAs share a row -- a rank of a sharded job computes what the full batch does.
  const int spc = GNF_SLABS / CS;  // slabs per CTA
  const int p_beg = rank * spc * slab_pix, p_end = min(p.HW, p_beg + spc * slab_pix);
  const act_t* x = p.x + (size_t)r * p.HW * p.C + (ch << 3);
  int px;
  for (int sl = 0; sl < spc; ++sl) {
    const int s_beg = p_beg + sl * slab_pix, s_end = min(p.HW, s_beg + slab_pix);
    float s1 = 0.f, s2 = 0.f;
    px = s_beg + lane_px;
    for (; px + (GNF_UNROLL - 1) * pstep < s_end; px += GNF_UNROLL * pstep) {
      uint4 u[GNF_UNROLL];
#pragma unroll
      for (int k = 0; k < GNF_UNROLL; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)(px + k * pstep) * p.C));
#pragma unroll
      for (int k = 0; k < GNF_UNROLL; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 += f[j] * f[j]; }
      }
    }
    for (; px < s_end; px += pstep) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (size_t)px * p.C)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 += f[j] * f[j]; }
    }
    s_part[threadIdx.x][0] = s1;
    s_part[threadIdx.x][1] = s2;
    __syncthreads();
    for (int half = pstep >> 1; half > 0; half >>= 1) {  // tree over the pixel lanes of every chunk (pstep is a power of two)
      if ((int)threadIdx.x < half * chunks) {
        s_part[threadIdx.x][0] += s_part[threadIdx.x + half * chunks][0];
        s_part[threadIdx.x][1] += s_part[threadIdx.x + half * chunks][1];
      }
      __syncthreads();
    }
    if ((int)threadIdx.x < p.G) {  // chunks of a group in index order
      const int g = threadIdx.x, cpc = cpg >> 3;
      float a = 0.f, b = 0.f;
      for (int c = g * cpc; c < (g + 1) * cpc; ++c) { a += s_part[c][0]; b += s_part[c][1]; }
      s_sum[sl][g][0] = a;
      s_sum[sl][g][1] = b;
    }
    __syncthreads();  // s_part is rewritten by the next slab
  }
  cluster.sync();  // every CTA's slab sums are in place
  if ((int)threadIdx.x < p.G) {
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int sl = 0; sl < GNF_SLABS; ++sl) {  // slab order, whatever CTA holds the slab
      const float* remote = cluster.map_shared_rank(&s_sum[0][0][0], sl / spc);
      a += remote[((sl % spc) * 64 + g) * 2];
      b += remote[((sl % spc) * 64 + g) * 2 + 1];
    }
    const float n = (float)p.HW * (float)cpg;
    const float mean = a / n;
    s_grp[g][0] = mean;
    s_grp[g][1] = rsqrtf(fmaxf(b / n - mean * mean, 0.f) + p.eps);
  }
  cluster.sync();  // remote reads done before any CTA of the cluster may exit; s_grp visible
  for (int c = threadIdx.x; c < p.C; c += GNF_THREADS) {  // A = rstd gamma (scale + 1), B = (beta - mean rstd gamma) (scale + 1) + shift
    const int g = c / cpg;
    float A = s_grp[g][1] * __ldg(p.gamma + c);
    float B = __ldg(p.beta + c) - s_grp[g][0] * A;
    if (p.tabA) {
      const float tA = __ldg(p.tabA + (size_t)(r / p.tab_div) * p.C + c), tB = __ldg(p.tabB + (size_t)(r / p.tab_div) * p.C + c);
      A *= tA;
      B = B * tA + tB;
    }
    s_ab[c] = A;
    s_ab[p.C + c] = B;
  }
  __syncthreads();
  // (the per-channel A / B stay in shared memory: registers buy occupancy here -- the pass is latency- and issue-bound)
  const float4* const sA4 = reinterpret_cast<const float4*>(s_ab + (ch << 3));
  const float4* const sB4 = reinterpret_cast<const float4*>(s_ab + p.C + (ch << 3));
  const DropRow dr = drop_row(p.drop, r, (uint64_t)p.HW * p.C);
  act_t* y = p.y + (size_t)r * p.HW * p.C + (ch << 3);
  const act_t* res = p.res ? p.res + (size_t)r * p.HW * p.res_ld + (ch << 3) : nullptr;
  for (px = p_beg + lane_px; px < p_end; px += AU * pstep) {
    uint4 u[AU], rr[AU];
#pragma unroll
    for (int k = 0; k < AU; ++k) {
      const int q = px + k * pstep;
      if (q < p_end) {
        u[k] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)q * p.C));
        if (res) rr[k] = __ldg(reinterpret_cast<const uint4*>(res + (size_t)q * p.res_ld));
      }
    }
#pragma unroll
    for (int k = 0; k < AU; ++k) {
      const int q = px + k * pstep;
      if (q >= p_end) break;
      float f[8];
      unpack8(u[k], f);
      {
        const float4 a0 = sA4[0], a1 = sA4[1], b0 = sB4[0], b1 = sB4[1];
        f[0] = fmaf(f[0], a0.x, b0.x); f[1] = fmaf(f[1], a0.y, b0.y); f[2] = fmaf(f[2], a0.z, b0.z); f[3] = fmaf(f[3], a0.w, b0.w);
        f[4] = fmaf(f[4], a1.x, b1.x); f[5] = fmaf(f[5], a1.y, b1.y); f[6] = fmaf(f[6], a1.z, b1.z); f[7] = fmaf(f[7], a1.w, b1.w);
      }
      if (p.act == ACT_SILU) {  // (the activation switch stays outside the element loop: the pass is issue-bound)
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] / (1.f + __expf(-f[j]));
      } else if (p.act != ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = apply_act(f[j], p.act);
      }
      if (p.drop.thresh) drop_apply8(p.drop, dr, (uint64_t)q * p.C + (ch << 3), f);
      if (res) {
        float g[8];
        unpack8(rr[k], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += g[j];
      }
      *reinterpret_cast<uint4*>(y + (size_t)q * p.C) = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------------ NS readout
constexpr int RO_MAXC = 4;
__global__ void __launch_bounds__(256) readout_kernel(const ReadoutParams p) {
  extern __shared__ float s_w[];  // [ky][kx][co][ci]
  const int nW = 16 * p.Cout * p.Cin;
  for (int i = threadIdx.x; i < nW; i += blockDim.x) {
    const int ci = i % p.Cin;
    const int co = (i / p.Cin) % p.Cout;
    const int kk = i / (p.Cin * p.Cout);
    s_w[i] = p.w[((size_t)ci * p.Cout + co) * 16 + kk];
  }
  __syncthreads();
  const int lanes = p.Cin >> 3;  // threads cooperating on one output pixel (8 for Cin = 64)
  const long long total = (long long)p.rows * p.Ho * p.Wo;
  const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / lanes;
  const int sub = threadIdx.x % lanes;
  const bool active = pix < total;
  const long long pp = active ? pix : 0;
  const int ox = (int)(pp % p.Wo);
  const int oy = (int)((pp / p.Wo) % p.Ho);
  const int r = (int)(pp / ((long long)p.Wo * p.Ho));
  const int H2 = 2 * p.Hs, W2 = 2 * p.Ws;
  int Y[2], X[2];
  float ly, lx;
  bilinear_coord(oy, H2, (float)H2 / (float)p.Ho, Y[0], Y[1], ly);
  bilinear_coord(ox, W2, (float)W2 / (float)p.Wo, X[0], X[1], lx);
  const float wy[2] = {1.f - ly, ly}, wx[2] = {1.f - lx, lx};
  float acc[RO_MAXC] = {0.f, 0.f, 0.f, 0.f};
  const act_t* xr = p.x + (size_t)r * p.Hs * p.Ws * p.Cin + (sub << 3);
  // 2 x 2 bilinear corners x 2 x 2 transposed-conv taps, all 16 combinations executed by every lane (no divergence):
  // invalid taps get weight 0 and a clamped address.
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int yy = Y[a], xx = X[b];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int ky = ((yy + 1) & 1) + 2 * dy, kx = ((xx + 1) & 1) + 2 * dx;
          const int iy = (yy + 1 - ky) >> 1, ix = (xx + 1 - kx) >> 1;
          const bool ok = iy >= 0 && iy < p.Hs && ix >= 0 && ix < p.Ws;
          const float wgt = ok ? wy[a] * wx[b] : 0.f;
          const int cy = min(max(iy, 0), p.Hs - 1), cx = min(max(ix, 0), p.Ws - 1);
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(xr + ((size_t)cy * p.Ws + cx) * p.Cin)), f);
          const float* wk = s_w + (size_t)(ky * 4 + kx) * p.Cout * p.Cin + (sub << 3);
          for (int co = 0; co < p.Cout; ++co) {
            const float4 w0 = *reinterpret_cast<const float4*>(wk + co * p.Cin);
            const float4 w1 = *reinterpret_cast<const float4*>(wk + co * p.Cin + 4);
            float d = f[0] * w0.x;
            d = fmaf(f[1], w0.y, d); d = fmaf(f[2], w0.z, d); d = fmaf(f[3], w0.w, d);
            d = fmaf(f[4], w1.x, d); d = fmaf(f[5], w1.y, d); d = fmaf(f[6], w1.z, d); d = fmaf(f[7], w1.w, d);
            acc[co] = fmaf(wgt, d, acc[co]);
          }
        }
    }
  for (int co = 0; co < p.Cout; ++co) {
    float v = acc[co];
    for (int o = 1; o < lanes; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (active && sub == 0) p.y[(((size_t)r * p.Cout + co) * p.Ho + oy) * p.Wo + ox] = v + __ldg(p.bias + co);
  }
}

// readout, step 2: gather (one thread per output pixel and channel)
__global__ void __launch_bounds__(256) readout_gather_kernel(const ReadoutGatherParams p) {
  const unsigned per_row = (unsigned)(p.Ho * p.Wo * p.Cout);
  const RowPos rp = row_pos(per_row);
  if (rp.i >= per_row) return;
  const unsigned pr = rp.i / (unsigned)p.Cout;  // output pixel inside the row
  const int co = (int)(rp.i - pr * p.Cout);
  const int oy = (int)(pr / (unsigned)p.Wo), ox = (int)(pr - (unsigned)oy * p.Wo);
  const int r = (int)rp.r;
  const int H2 = 2 * p.Hs, W2 = 2 * p.Ws, ldz = 16 * p.Cout;
  int Y[2], X[2];
  float ly, lx;
  bilinear_coord(oy, H2, (float)H2 / (float)p.Ho, Y[0], Y[1], ly);
  bilinear_coord(ox, W2, (float)W2 / (float)p.Wo, X[0], X[1], lx);
  const float wy[2] = {1.f - ly, ly}, wx[2] = {1.f - lx, lx};
  const act_t* zr = p.z + (size_t)r * p.Hs * p.Wz * ldz + co;
  // 2 x 2 bilinear corners x 2 x 2 transposed-conv taps: the 16 addresses are formed first (clamped, invalid taps masked
  // afterwards) so that the 16 loads of a thread are in flight together
  int rowo[2][2], colo[2][2];  // [corner][tap]: element offset of the source row / of (column, tap channel)
  bool rok[2][2], cok[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int dd = 0; dd < 2; ++dd) {
      const int ky = ((Y[a] + 1) & 1) + 2 * dd, iy = (Y[a] + 1 - ky) >> 1;
      rok[a][dd] = iy >= 0 && iy < p.Hs;
      rowo[a][dd] = min(max(iy, 0), p.Hs - 1) * p.Wz * ldz + ky * 4 * p.Cout;
      const int kx = ((X[a] + 1) & 1) + 2 * dd, ix = (X[a] + 1 - kx) >> 1;
      cok[a][dd] = ix >= 0 && ix < p.Ws;
      const int cx = min(max(ix, 0), p.Ws - 1);
      colo[a][dd] = (p.xinv ? __ldg(p.xinv + cx) : cx) * ldz + kx * p.Cout;
    }
  act_t raw[2][2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) raw[a][b][dy][dx] = zr[rowo[a][dy] + colo[b][dx]];
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      float v = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          if (rok[a][dy] && cok[b][dx]) v += act2f(raw[a][b][dy][dx]);
      acc = fmaf(wy[a] * wx[b], v, acc);
    }
  p.y[(((size_t)r * p.Cout + co) * p.Ho + oy) * p.Wo + ox] = acc + __ldg(p.bias + co);
}

__global__ void convt_to_conv1x1_kernel(const float* __restrict__ wt, float* __restrict__ out, int Cin, int Cout) {
  const int j = blockIdx.x;  // output row (ky*4+kx)*Cout + co
  const int co = j % Cout, kk = j / Cout;
  for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) out[(size_t)j * Cin + ci] = wt[((size_t)ci * Cout + co) * 16 + kk];
}

// ------------------------------------------------------------------------------------------------ time tables
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Stage 1: one block per distinct time value: sinusoidal embedding -> Linear -> GELU -> Linear -> SiLU, kept in global
// memory ([tab_rows, time_dim]); the SiLU belongs to every per-block time MLP (SiLU -> Linear) and is hoisted here.
__global__ void __launch_bounds__(256) time_embed_kernel(const TimeParams p) {
  __shared__ float s_emb[256], s_h[512];
  const int r = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float t = p.time[r];
  const int half = p.dim / 2;
  const double k = -log(10000.0) / (double)(half - 1);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = (float)exp((double)i * k);   // frequencies in fp32 like the reference (misc.py:27-28)
    const double a = (double)(t * f);            // fp32 product, then an accurately reduced sin / cos
    s_emb[i] = (float)sin(a);
    s_emb[half + i] = (float)cos(a);
  }
  __syncthreads();
  for (int o = warp; o < p.time_dim; o += nwarp) {  // Linear(dim, time_dim) + exact GELU
    const float* w = p.packed + p.w1_off + (size_t)o * p.dim;
    float s = 0.f;
    for (int i = lane; i < p.dim; i += 32) s = fmaf(w[i], s_emb[i], s);
    s = warp_sum(s);
    if (lane == 0) {
      const float v = s + p.packed[p.b1_off + o];
      s_h[o] = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    }
  }
  __syncthreads();
  for (int o = warp; o < p.time_dim; o += nwarp) {  // Linear(time_dim, time_dim), then SiLU
    const float* w = p.packed + p.w2_off + (size_t)o * p.time_dim;
    float s = 0.f;
    for (int i = lane; i < p.time_dim; i += 32) s = fmaf(w[i], s_h[i], s);
    s = warp_sum(s);
    if (lane == 0) {
      const float v = s + p.packed[p.b2_off + o];
      p.temb[(size_t)r * p.time_dim + o] = v / (1.f + expf(-v));
    }
  }
}

// Stage 2: one warp per (block of TT_ROWS time rows, layer, channel): (scale, shift) = Linear(time_dim, 2C)(SiLU(temb)) folded
// with the layer's norm/bias affine into the epilogue tables A, B.  The two weight rows of the channel are loaded once into
// registers and reused for every time row of the block (per-row weight reads were 1.8 GB of L2 traffic for 304 rows of the SST
// Unet: 2.0 ms per forward whose times are not cached); the per-row arithmetic and its order are unchanged.
constexpr int TT_ROWS = 8;
__global__ void __launch_bounds__(256) time_tables_kernel(const TimeParams p, int total_ch) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int row_blocks = (p.rows + TT_ROWS - 1) / TT_ROWS;
  if (gw >= (long long)total_ch * row_blocks) return;
  const int rb = (int)(gw / total_ch);
  const int f = (int)(gw - (long long)rb * total_ch);  // flat (padded) channel index = tab_off + c
  int li = 0;
  for (int j = 1; j < p.n_layers; ++j)
    if (p.layers[j].tab_off <= f) li = j;  // layers are ordered by tab_off
  const TimeLayer L = p.layers[li];
  const int c = f - (int)L.tab_off;
  if (c >= L.C) return;  // padding slot
  const bool timed = p.time != nullptr && L.w_off >= 0;
  float ws[16], wh[16];  // time_dim <= 512: 16 values per lane
  float bs = 0.f, bh = 0.f;
  if (timed) {
    const float* w0 = p.packed + L.w_off + (size_t)c * p.time_dim;
    const float* w1 = p.packed + L.w_off + (size_t)(L.C + c) * p.time_dim;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = lane + 32 * k;
      ws[k] = i < p.time_dim ? w0[i] : 0.f;
      wh[k] = i < p.time_dim ? w1[i] : 0.f;
    }
    bs = p.packed[L.b_off + c];
    bh = p.packed[L.b_off + L.C + c];
  }
  const float na = L.na_off >= 0 ? p.packed[L.na_off + c] : 1.f;
  const float nb = L.nb_off >= 0 ? p.packed[L.nb_off + c] : 0.f;
  for (int r = rb * TT_ROWS; r < min(p.rows, (rb + 1) * TT_ROWS); ++r) {
    float scale = 0.f, shift = 0.f;
    if (timed) {
      const float* st = p.temb + (size_t)r * p.time_dim;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int i = lane + 32 * k;
        if (i < p.time_dim) {
          const float tv = st[i];
          s1 = fmaf(ws[k], tv, s1);
          s2 = fmaf(wh[k], tv, s2);
        }
      }
      scale = warp_sum(s1) + bs;
      shift = warp_sum(s2) + bh;
    }
    if (lane == 0) {
      float* A = p.tabA + (size_t)L.tab_off * p.rows + (size_t)r * L.C;
      float* B = p.tabB + (size_t)L.tab_off * p.rows + (size_t)r * L.C;
      if (L.mode == 1) {
        A[c] = scale + 1.f;
        B[c] = shift;
      } else {
        A[c] = na * (scale + 1.f);
        B[c] = nb * (scale + 1.f) + shift;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ weight re-packing
__global__ void __launch_bounds__(256) repack_conv_kernel(const float* __restrict__ w, act_t* __restrict__ out,
                                                         int I, int KH, int KW, int Cpad, int Kpad, int standardize) {
  __shared__ float s_red[2][8];
  __shared__ float s_stat[2];
  const int o = blockIdx.x;
  const int n = I * KH * KW;
  const float* wo = w + (size_t)o * n;
  float mean = 0.f, rstd = 1.f;
  if (standardize) {  // WeightStandardizedConv2d (unet.py:32-40): biased variance, eps = 1e-5 (fp32 inputs)
    float s1 = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s1 += wo[i];
    s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) s_red[0][threadIdx.x >> 5] = s1;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[0][i];
      s_stat[0] = t / n;
    }
    __syncthreads();
    mean = s_stat[0];
    float s2 = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = wo[i] - mean; s2 += d * d; }
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) s_red[1][threadIdx.x >> 5] = s2;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[1][i];
      s_stat[1] = rsqrtf(t / n + 1e-5f);
    }
    __syncthreads();
    rstd = s_stat[1];
  }
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    const int tap = k / Cpad, c = k - tap * Cpad;
    float v = 0.f;
    if (tap < KH * KW && c < I) {
      const int ky = tap / KW, kx = tap - ky * KW;
      v = (wo[((size_t)c * KH + ky) * KW + kx] - mean) * rstd;
    }
    out[(size_t)o * Kpad + k] = f2act(v);
  }
}

__global__ void __launch_bounds__(256) compose_conv_kernel(const float* __restrict__ w0, const float* __restrict__ wi,
                                                          const float* __restrict__ bi, act_t* __restrict__ out,
                                                          int Cm, int Cs, int taps, int Cpad, int Kpad) {
  const int o = blockIdx.x;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    const int tap = k / Cpad, ch = k - tap * Cpad;
    float v = 0.f;
    if (tap < taps && ch <= Cs) {
      for (int m = 0; m < Cm; ++m) {
        const float a = w0[((size_t)o * Cm + m) * taps + tap];
        v = fmaf(a, ch < Cs ? wi[(size_t)m * Cs + ch] : bi[m], v);
      }
    }
    out[(size_t)o * Kpad + k] = f2act(v);
  }
}

__global__ void __launch_bounds__(256) compose_s2d_kernel(const float* __restrict__ w0, const float* __restrict__ wi,
                                                         const float* __restrict__ bi, float* __restrict__ out, int Cm,
                                                         int Cs) {
  const int o = blockIdx.x;
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) {
    const int slot = i / 9, t = i - slot * 9;
    const int sub = slot >> 4, c = slot & 15, dy = sub >> 1, dx = sub & 1, ty = t / 3, tx = t - ty * 3;
    const int ky = 2 * (ty - 1) + dy + 1, kx = 2 * (tx - 1) + dx + 1;  // tap of the original 4x4 / stride-2 conv
    float v = 0.f;
    if (ky >= 0 && ky < 4 && kx >= 0 && kx < 4 && c <= Cs) {
      for (int m = 0; m < Cm; ++m) {
        const float a = w0[(((size_t)o * Cm + m) * 4 + ky) * 4 + kx];
        v = fmaf(a, c < Cs ? wi[(size_t)m * Cs + c] : bi[m], v);
      }
    }
    out[(size_t)o * 576 + i] = v;
  }
}

__global__ void fold_norm_kernel(const float* bias, const float* g, const float* beta, const float* mean,
                                 const float* var, float eps, float* na, float* nb, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float b = bias ? bias[c] : 0.f;
  if (g) {
    const float a = g[c] * rsqrtf(var[c] + eps);
    na[c] = a;
    nb[c] = (b - mean[c]) * a + beta[c];
  } else {
    na[c] = 1.f;
    nb[c] = b;
  }
}

// ------------------------------------------------------------------------------------------------ sampler elementwise
__global__ void cold_update_kernel(float* x_s, const float* a, const float* b, float* out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = a ? (x_s[i] - a[i]) + b[i] : b[i];  // x_s - D(x0_hat, s) + D(x0_hat, s_next)  (dyffusion.py:386-388)
  x_s[i] = v;
  if (out) out[i] = v;
}
__global__ void fill_kernel(float* p, float v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void dropout_mask_kernel(DropCfg d, long long n, uint8_t* mask) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g * 8 >= n) return;
  const uint32_t keep = d.thresh ? drop_keep_bits8(d, drop_row(d, 0, 0), (uint64_t)g * 8) : 0xFFu;
  for (int j = 0; j < 8 && g * 8 + j < n; ++j) mask[g * 8 + j] = (keep >> j) & 1u;
}

}  // namespace

int launch_head1x1(const Head1x1Params& p, cudaStream_t s) {
  if (p.OC > 8 || p.C % 16) { set_error("head1x1: needs <= 8 output channels and C % 16 == 0"); return -1; }
  ProfScope prof(s, KC_ELEMENTWISE, 2.0 * (double)p.M * p.C * p.OC, (double)p.M * (2.0 * p.C + 4.0 * p.OC));
  head1x1_kernel<<<cdiv(p.M, 256), 256, (size_t)(p.OC * p.C + p.OC) * sizeof(float), s>>>(p);
  DYF_LAUNCH_OK("head1x1_kernel");
  return 0;
}

int launch_stem_xim2col_weight(const float* w, float* out, int O, int C, cudaStream_t s) {
  stem_xim2col_weight_kernel<<<cdiv((long long)O * 64 * 7, 256), 256, 0, s>>>(w, out, O, C);
  DYF_LAUNCH_OK("stem_xim2col_weight_kernel");
  return 0;
}

int launch_pack(const PackParams& p, cudaStream_t s) {
  if (p.s2d == 2) {  // x-im2col for a 7x7 stem (<= 8 input channels)
    S2dChannels ch{};
    PackParams q = p;
    q.noise_c0 = q.noise_c1 = 0;
    const size_t plane = (size_t)p.Hi * p.Wi;
    for (int i = 0; i < p.nsrc; ++i) {
      if (i == p.noise_src) { q.noise_c0 = ch.n; q.noise_c1 = ch.n + p.C[i]; }
      for (int c = 0; c < p.C[i]; ++c, ++ch.n) {
        if (ch.n >= 8) { set_error("pack x-im2col: needs <= 8 input channels"); return -1; }
        ch.plane[ch.n] = p.src[i] + (size_t)c * plane;
        ch.row_stride[ch.n] = (int)(p.C[i] * plane);
      }
    }
    if (p.bilinear || p.Hi != p.Ho || p.Wi != p.Wo) { set_error("pack x-im2col: no resize"); return -1; }
    ProfScope prof(s, KC_PACK);
    pack_xim2col_kernel<<<(unsigned)(p.rows * p.Ho), 128, (size_t)(p.Wi + 6) * sizeof(uint4), s>>>(q, ch);
    DYF_LAUNCH_OK("pack_xim2col_kernel");
    return 0;
  }
  if (p.s2d) {
    int ctot = 0;
    for (int i = 0; i < p.nsrc; ++i) ctot += p.C[i];
    if (ctot > 15 || p.ones_channel != ctot || p.noise_src >= 0) { set_error("pack s2d: needs <= 15 input channels"); return -1; }
    S2dChannels ch{};
    const size_t plane = (size_t)p.Hi * p.Wi;
    for (int i = 0; i < p.nsrc; ++i)
      for (int c = 0; c < p.C[i]; ++c, ++ch.n) {
        ch.plane[ch.n] = p.src[i] + (size_t)c * plane;
        ch.row_stride[ch.n] = (int)(p.C[i] * plane);
      }
    ProfScope prof(s, KC_PACK);
    const unsigned grid = row_grid(p.rows, (long long)p.Ho * p.Wo * 4);
    switch (ch.n) {
#define DYF_S2D_CASE(N) case N: pack_s2d_kernel<N><<<grid, 256, 0, s>>>(p, ch); break;
      DYF_S2D_CASE(1) DYF_S2D_CASE(2) DYF_S2D_CASE(3) DYF_S2D_CASE(4) DYF_S2D_CASE(5) DYF_S2D_CASE(6) DYF_S2D_CASE(7)
      DYF_S2D_CASE(8) DYF_S2D_CASE(9) DYF_S2D_CASE(10) DYF_S2D_CASE(11) DYF_S2D_CASE(12) DYF_S2D_CASE(13) DYF_S2D_CASE(14)
      DYF_S2D_CASE(15)
#undef DYF_S2D_CASE
      default: set_error("pack s2d: needs 1..15 input channels"); return -1;
    }
    DYF_LAUNCH_OK("pack_s2d_kernel");
    return 0;
  }
  if (p.Cpad % 8 != 0) { set_error("pack: Cpad must be a multiple of 8"); return -1; }
  ProfScope prof(s, KC_PACK);
  pack_kernel<<<row_grid(p.rows, (long long)p.Ho * p.Wo), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("pack_kernel");
  return 0;
}

int launch_stem(const StemParams& p, cudaStream_t s) {
  if (p.Cout != 64 || p.Cin > 16) { set_error("stem: needs dim == 64 and at most 16 input channels"); return -1; }
  const long long total = (long long)p.pk.rows * p.pk.Ho * p.pk.Wo;
  ProfScope prof(s, KC_PACK, 2.0 * p.Cin * 64 * (double)total, 2.0 * 64 * (double)total);
  stem_kernel<<<cdiv(total, 256), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("stem_kernel");
  return 0;
}

int launch_upsample(const UpsampleParams& p, cudaStream_t s) {
  if ((p.C[0] | p.C[1] | p.ld[0] | p.ld[1]) & 7) { set_error("upsample: channel counts must be multiples of 8"); return -1; }
  if (p.scale == 2 && p.bilinear) {
    const long long quads = (long long)p.rows * p.H * p.W * ((p.C[0] + p.C[1]) >> 3);
    ProfScope prof(s, KC_UPSAMPLE, 0.0, 2.0 * 5.0 * (double)quads * 8);
    upsample2x_quad_kernel<<<row_grid(p.rows, (long long)p.H * p.W * ((p.C[0] + p.C[1]) >> 3)), 256, 0, s>>>(p);
    DYF_LAUNCH_OK("upsample2x_quad_kernel");
    return 0;
  }
  const long long total = (long long)p.rows * p.H * p.scale * p.W * p.scale * ((p.C[0] + p.C[1]) >> 3);
  ProfScope prof(s, KC_UPSAMPLE, 0.0, 2.0 * 1.25 * (double)total * 8);
  upsample_kernel<<<row_grid(p.rows, (long long)p.H * p.scale * p.W * p.scale * ((p.C[0] + p.C[1]) >> 3)), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("upsample_kernel");
  return 0;
}

int gn_scratch_floats(int C, int G) { return 2 * G * GN_SLABS + 2 * C; }  // per row: slab partials + folded (A, B)

int launch_groupnorm(const GroupNormParams& p, cudaStream_t s) {
  const int chunks = p.C >> 3;
  if (p.C % (8 * p.G) != 0 || chunks > 256 || 256 % chunks != 0 || p.G > 64) {
    set_error("groupnorm: unsupported channel/group configuration");
    return -1;
  }
  static const bool no_fuse = getenv("DYF_DISABLE_GN_FUSE") != nullptr;
  if (!no_fuse && GNF_THREADS % chunks == 0 && chunks <= GNF_THREADS / 2 && (size_t)p.HW * p.C * sizeof(act_t) <= (1u << 20)) {
    // rows of <= 1 MB: a cluster of up to 8 CTAs per row (~64 KB each), the second pass re-reads the slice from L2
    const size_t row_bytes = (size_t)p.HW * p.C * sizeof(act_t);
    int CS = 1;
    while (CS < 8 && row_bytes / CS > (64u << 10)) CS *= 2;
    // few rows (a rank of a sharded job): more CTAs per row, so that the two dependent passes of a CTA are short and at least
    // two CTAs per SM are in flight
    while (CS < 8 && (long long)p.rows * CS < 2 * 148 && row_bytes / CS > (8u << 10)) CS *= 2;
    const int pstep = GNF_THREADS / chunks;
    const int per = cdiv(cdiv(p.HW, GNF_SLABS), pstep) * pstep;  // pixels per slab (whole pixel-lane rounds), independent of CS
    ProfScope prof(s, KC_GROUPNORM);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.rows * CS)); cfg.blockDim = dim3(GNF_THREADS);
    cfg.dynamicSmemBytes = 2 * p.C * sizeof(float); cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // (see pdl_wait in common.cuh)
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled(2) ? 2 : 1;
    // one pixel in flight per thread and 40 registers (6 CTAs / SM, no spills): measured against 2 and 4 pixels in flight
    // (2.36 - 2.72 ms per 304-row SST forward: spills) and against 32 registers / 8 CTAs (2.07 ms, 24 bytes of spills): 2.04 ms
    DYF_CUDA_OK(cudaLaunchKernelEx(&cfg, groupnorm_fused_kernel<1, 6>, p, CS, per));
    count_launch();
    return 0;
  }
  const int pstep = 256 / chunks;
  int pix_per_block = pstep * 16;
  if (cdiv(p.HW, pix_per_block) > GN_SLABS) pix_per_block = cdiv(cdiv(p.HW, GN_SLABS), pstep) * pstep;
  GroupNormParams q = p;
  q.slabs = cdiv(p.HW, pix_per_block);
  dim3 grid(q.slabs, p.rows);
  ProfScope prof(s, KC_GROUPNORM);
  groupnorm_stats_kernel<<<grid, 256, 0, s>>>(q, pix_per_block);
  DYF_LAUNCH_OK("groupnorm_stats_kernel");
  float* ab = p.stats + (size_t)p.rows * p.G * GN_SLABS * 2;  // [rows][2][C] behind the slab partials (see gn_scratch_floats)
  groupnorm_fold_kernel<<<p.rows, 256, 0, s>>>(q, ab);
  DYF_LAUNCH_OK("groupnorm_fold_kernel");
  const long long total = (long long)p.rows * p.HW * chunks;
  groupnorm_apply_kernel<<<cdiv(total, 256), 256, 0, s>>>(q, ab);
  DYF_LAUNCH_OK("groupnorm_apply_kernel");
  return 0;
}

int launch_readout(const ReadoutParams& p, cudaStream_t s) {
  const int lanes = p.Cin >> 3;
  if (p.Cout > RO_MAXC || p.Cin % 8 != 0 || lanes > 32 || (lanes & (lanes - 1)) != 0) {
    set_error("readout: unsupported channel configuration");
    return -1;
  }
  const long long total = (long long)p.rows * p.Ho * p.Wo * lanes;
  const size_t smem = (size_t)16 * p.Cout * p.Cin * sizeof(float);
  ProfScope prof(s, KC_READOUT);
  readout_kernel<<<cdiv(total, 256), 256, smem, s>>>(p);
  DYF_LAUNCH_OK("readout_kernel");
  return 0;
}

int launch_readout_gather(const ReadoutGatherParams& p, cudaStream_t s) {
  ProfScope prof(s, KC_READOUT);
  readout_gather_kernel<<<row_grid(p.rows, (long long)p.Ho * p.Wo * p.Cout), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("readout_gather_kernel");
  return 0;
}

// Conv2d weight [O, C, 2, 2] -> 1x1-conv weight [O, 4C] over the space-to-depth view, K index = (ky*2 + kx)*C + ci
__global__ void k2s2_to_conv1x1_kernel(const float* __restrict__ w, float* __restrict__ out, int C) {
  const int o = blockIdx.x;
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    const int t = i / C, ci = i - t * C;
    out[(size_t)o * 4 * C + i] = w[((size_t)o * C + ci) * 4 + t];
  }
}
int launch_k2s2_to_conv1x1(const float* w, float* out, int O, int C, cudaStream_t s) {
  k2s2_to_conv1x1_kernel<<<O, 256, 0, s>>>(w, out, C);
  DYF_LAUNCH_OK("k2s2_to_conv1x1_kernel");
  return 0;
}

int launch_convt_to_conv1x1(const float* wt, float* out, int Cin, int Cout, cudaStream_t s) {
  convt_to_conv1x1_kernel<<<16 * Cout, 64, 0, s>>>(wt, out, Cin, Cout);
  DYF_LAUNCH_OK("convt_to_conv1x1_kernel");
  return 0;
}

int launch_time_tables(const TimeParams& p, cudaStream_t s) {
  if (p.dim > 256 || p.time_dim > 512) { set_error("time tables: dim too large"); return -1; }
  if (p.n_layers == 0) return 0;
  ProfScope prof(s, KC_TIME);
  if (p.time != nullptr) {
    time_embed_kernel<<<p.rows, 256, 0, s>>>(p);
    DYF_LAUNCH_OK("time_embed_kernel");
  }
  const long long warps = (long long)p.total_ch * ((p.rows + TT_ROWS - 1) / TT_ROWS);
  time_tables_kernel<<<cdiv(warps * 32, 256), 256, 0, s>>>(p, p.total_ch);
  DYF_LAUNCH_OK("time_tables_kernel");
  return 0;
}

int launch_repack_conv(const float* w, act_t* out, int O, int I, int KH, int KW, int Cpad, int Kpad,
                       int standardize, cudaStream_t s) {
  repack_conv_kernel<<<O, 256, 0, s>>>(w, out, I, KH, KW, Cpad, Kpad, standardize);
  DYF_LAUNCH_OK("repack_conv_kernel");
  return 0;
}

int launch_compose_conv(const float* w0, const float* wi, const float* bi, act_t* out, int O, int Cm, int Cs,
                        int KH, int KW, int Cpad, int Kpad, cudaStream_t s) {
  compose_conv_kernel<<<O, 256, 0, s>>>(w0, wi, bi, out, Cm, Cs, KH * KW, Cpad, Kpad);
  DYF_LAUNCH_OK("compose_conv_kernel");
  return 0;
}

int launch_compose_s2d(const float* w0, const float* wi, const float* bi, float* out, int O, int Cm, int Cs,
                       cudaStream_t s) {
  compose_s2d_kernel<<<O, 256, 0, s>>>(w0, wi, bi, out, Cm, Cs);
  DYF_LAUNCH_OK("compose_s2d_kernel");
  return 0;
}

int launch_fold_norm(const float* bias, const float* g, const float* beta, const float* mean, const float* var,
                     float eps, float* na, float* nb, int C, cudaStream_t s) {
  fold_norm_kernel<<<cdiv(C, 128), 128, 0, s>>>(bias, g, beta, mean, var, eps, na, nb, C);
  DYF_LAUNCH_OK("fold_norm_kernel");
  return 0;
}

int launch_cold_update(float* x_s, const float* a, const float* b, float* out, long long n, cudaStream_t s) {
  ProfScope prof(s, KC_ELEMENTWISE);
  cold_update_kernel<<<cdiv(n, 256), 256, 0, s>>>(x_s, a, b, out, n);
  DYF_LAUNCH_OK("cold_update_kernel");
  return 0;
}
int launch_fill(float* p, float v, long long n, cudaStream_t s) {
  ProfScope prof(s, KC_ELEMENTWISE);
  fill_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, v, n);
  DYF_LAUNCH_OK("fill_kernel");
  return 0;
}
int launch_dropout_mask(DropCfg d, long long n, uint8_t* mask, cudaStream_t s) {
  dropout_mask_kernel<<<cdiv((n + 7) / 8, 256), 256, 0, s>>>(d, n, mask);
  DYF_LAUNCH_OK("dropout_mask_kernel");
  return 0;
}

}  // namespace dyf
