// Linear-attention block of the SST backbone as TWO tensor-core kernels (reference: Residual(PreNorm(LayerNorm,
// LinearAttention(rescale="qkv"))), src/models/unet.py:43-52,183-191 and src/models/modules/attention.py:7-50):
//
//   y = x + to_out( ctx^T . (softmax_d(q) * dh^-0.5) ),   ctx[d][e] = sum_n softmax_n(k)[d, n] * v[e, n] / n,
//   (q, k, v) = to_qkv( Dropout( LayerNorm_c(x) ) )        per row (image) and head; n = H*W positions, dh = 32
//
// The un-fused path runs LayerNorm, the 1x1 qkv conv, two attention kernels and the 1x1 output conv separately and moves
// 2.9 GB per 60x60 block of a 304-row batch (the [n, 384] qkv tensor alone is written and re-read); here the block reads x
// twice and writes y once (0.42 GB), and every contraction runs on the tensor cores (mma.sync.m16n8k16, fp32 accumulate):
//
//   pass 1 (linattn_ctx_fused): per (row, pixel chunk): LayerNorm + dropout -> shared memory; k^T = W_k . y^T and
//          v^T = W_v . y^T with the PIXELS on the MMA N axis, so the accumulator fragments of k^T / v^T are exactly the A / B
//          fragments of the context product P . v^T (K axis = pixels) -- no shared-memory round trip; streaming softmax
//          over the pixels (running max, rescaled accumulators) -> per-chunk partial (max, sum, ctx) in global memory
//   pass 2 (linattn_out_fused): per (row, 128-pixel tile): combine the chunk partials -> ctx (fp16, shared memory);
//          LayerNorm + dropout again (same Philox masks); q = y . W_q^T -> softmax over the head dimension in registers ->
//          out = q~ . ctx -> to_out GEMM, each accumulator fragment re-packed as the next A fragment -> + bias + x -> y
//
// Tensor-core generation: these are register-level chains of four small GEMMs (K = 32 ... 256) with softmaxes in between;
// tcgen05 keeps accumulators in TMEM and would need a TMEM -> register -> shared-memory round trip per link of the chain,
// so the chain uses mma.sync; the block is 1.7 % of the network's FLOPs and, fused, bound by its three passes over x.
#include <algorithm>
#include <cstdlib>

#include "aux.cuh"

namespace dyf {
namespace {

constexpr int DH = 32, HEADS = 4, HID = HEADS * DH;
constexpr int TP = 128;           // pixels per tile (8 warps x 16)
constexpr int FTHREADS = 256;
constexpr int PART = 2 * DH + DH * DH;  // floats of one (row, chunk, head) partial: max[32], sum[32], ctx[32][32]

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." DYF_MMA_T "." DYF_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// address of this lane's row for a 16 x 16 A block (rows r0.., cols k0..) of a row-major [rows][ld] b16 tile:
// matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15) = a0..a3
__device__ __forceinline__ uint32_t a_addr(const act_t* tile, int ld, int r0, int k0, int lane) {
  return s_u32(tile + (size_t)(r0 + (lane & 15)) * ld + k0 + ((lane >> 4) << 3));
}
// ... and for a 16(n) x 16(k) block of B stored as [n][ld] (k contiguous): matrices (n 0-7, k 0-7), (n 0-7, k 8-15),
// (n 8-15, k 0-7), (n 8-15, k 8-15) = (b0, b1) of n-block 0, (b0, b1) of n-block 1
__device__ __forceinline__ uint32_t b_addr(const act_t* tile, int ld, int n0, int k0, int lane) {
  return s_u32(tile + (size_t)(n0 + ((lane >> 4) << 3) + (lane & 7)) * ld + k0 + (((lane >> 3) & 1) << 3));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// LayerNorm over the channels (biased variance, eps 1e-5, gain only) + dropout of one pixel tile -> shared memory
// [TP][C + 8] (rows past the end of the image are zero).  C / 32 threads per pixel, 32 channels each.
template <int C>
__device__ __forceinline__ void load_ln_tile(const LinAttnFusedParams& p, int row, int pix0, int pix_end, act_t* sX) {
  constexpr int LD = C + 8, TPP = C / 32, PIX_PER_PASS = FTHREADS / TPP;
  const int tid = threadIdx.x, part = tid % TPP;
  const DropRow dr = drop_row(p.drop, row, (uint64_t)p.n * C);
#pragma unroll 1
  for (int px = tid / TPP; px < TP; px += PIX_PER_PASS) {
    const int pix = pix0 + px;
    const bool ok = pix < pix_end;
    float v[32];
    if (ok) {
      const uint4* src = reinterpret_cast<const uint4*>(p.x + ((size_t)row * p.n + pix) * C + part * 32);
#pragma unroll
      for (int i = 0; i < 4; i += 2) {  // 64 contiguous bytes per thread: two 256-bit loads
        uint4 a, b;
        ld_global_nc_256(src + i, a, b);
        unpack8(a, v + 8 * i);
        unpack8(b, v + 8 * i + 8);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
#pragma unroll
    for (int o = TPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = TPP / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / C) + 1e-5f);
    uint4* dst = reinterpret_cast<uint4*>(sX + (size_t)px * LD + part * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float o[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.g + part * 32 + 8 * i));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.g + part * 32 + 8 * i) + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ok ? (v[8 * i + j] - mean) * rstd * g[j] : 0.f;
      if (p.drop.thresh && ok) {
        const uint32_t keep = drop_keep_bits8(p.drop, dr, (uint64_t)pix * C + part * 32 + 8 * i);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = ((keep >> j) & 1u) ? o[j] * p.drop.scale : 0.f;
      }
      dst[i] = pack8(o);
    }
  }
}

template <int ROWS, int K>
__device__ __forceinline__ void load_weight_rows(const act_t* __restrict__ w, int ldw, act_t* sW) {
  // [ROWS][K] (row stride ldw in global) -> shared memory [ROWS][K + 8]
  constexpr int LD = K + 8, PER = K / 8;
  for (int i = threadIdx.x; i < ROWS * PER; i += FTHREADS) {
    const int r = i / PER, c = i - r * PER;
    *reinterpret_cast<uint4*>(sW + (size_t)r * LD + c * 8) = __ldg(reinterpret_cast<const uint4*>(w + (size_t)r * ldw) + c);
  }
}

// ------------------------------------------------------------------------------------------------ pass 1
// warp = (head, pixel half of the tile).  Shared memory: W_k | W_v of all heads [256][C + 8], pixel tile [TP][C + 8].
template <int C>
__global__ void __launch_bounds__(FTHREADS) linattn_ctx_fused_kernel(const LinAttnFusedParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int LD = C + 8;
  act_t* sW = reinterpret_cast<act_t*>(smem_raw);          // rows 0-127: W_k (head-major), rows 128-255: W_v
  act_t* sX = sW + 2 * HID * LD;
  const int row = blockIdx.y, chunk = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = warp & 3, ph = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  pdl_trigger();
  load_weight_rows<2 * HID, C>(p.w_qkv + (size_t)HID * p.ldw, p.ldw, sW);  // k and v rows of to_qkv.weight [384][C]
  pdl_wait();  // (the weights are static: staged while the previous kernel drains; x is its output)
  const int c_beg = chunk * p.chunk_pix, c_end = min(p.n, c_beg + p.chunk_pix);

  float ctx[2][4][4];       // [d block of 16][e block of 8][frag]: rows d = 16 mb + g (+8), cols e = 8 eb + 2t (+1)
  float m_run[2][2], l_run[2][2];  // running max / sum of rows d = 16 mb + g + 8 j
#pragma unroll
  for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
    for (int j = 0; j < 2; ++j) { m_run[mb][j] = -INFINITY; l_run[mb][j] = 0.f; }
#pragma unroll
    for (int eb = 0; eb < 4; ++eb)
#pragma unroll
      for (int f = 0; f < 4; ++f) ctx[mb][eb][f] = 0.f;
  }
  const act_t* const sWk = sW + (size_t)(h * DH) * LD;
  const act_t* const sWv = sW + (size_t)(HID + h * DH) * LD;

  for (int pix0 = c_beg; pix0 < c_end; pix0 += TP) {
    __syncthreads();  // previous tile consumed (and, first time, weights visible after the next barrier)
    load_ln_tile<C>(p, row, pix0, c_end, sX);
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {      // this warp's 64 pixels in steps of 16
      const int pl = ph * 64 + sub * 16;     // first pixel of the step inside the tile
      if (pix0 + pl >= c_end) break;         // (warp-uniform)
      float kT[2][2][4], vT[2][2][4];        // [d / e block of 16][pixel block of 8][frag]
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int f = 0; f < 4; ++f) { kT[mb][nb][f] = 0.f; vT[mb][nb][f] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        uint32_t bx[4];
        ldsm4(bx, b_addr(sX, LD, pl, ks * 16, lane));  // y^T: (b0, b1) of pixel blocks 0 and 1
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          uint32_t ak[4], av[4];
          ldsm4(ak, a_addr(sWk, LD, mb * 16, ks * 16, lane));
          ldsm4(av, a_addr(sWv, LD, mb * 16, ks * 16, lane));
          mma16816(kT[mb][0], ak, bx[0], bx[1]);
          mma16816(kT[mb][1], ak, bx[2], bx[3]);
          mma16816(vT[mb][0], av, bx[0], bx[1]);
          mma16816(vT[mb][1], av, bx[2], bx[3]);
        }
      }
      // pixels past the end of the chunk do not take part in the softmax
      bool okp[2][2];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) okp[nb][e] = pix0 + pl + nb * 8 + 2 * t + e < c_end;
      uint32_t pa[2][4];  // P = exp(k - max) as A fragments [d block][a0..a3]
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        float alpha[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // rows d = 16 mb + g + 8 j hold frags [2 j], [2 j + 1] of both pixel blocks
          float mx = -INFINITY;
#pragma unroll
          for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) if (okp[nb][e]) mx = fmaxf(mx, kT[mb][nb][2 * j + e]);
          mx = quad_max(mx);
          const float m_new = fmaxf(m_run[mb][j], mx);
          alpha[j] = m_run[mb][j] == -INFINITY ? 0.f : __expf(m_run[mb][j] - m_new);
          float sum = 0.f;
#pragma unroll
          for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float pv = okp[nb][e] ? __expf(kT[mb][nb][2 * j + e] - m_new) : 0.f;
              kT[mb][nb][2 * j + e] = pv;
              sum += pv;
            }
          l_run[mb][j] = l_run[mb][j] * alpha[j] + quad_sum(sum);
          m_run[mb][j] = m_new;
        }
        pa[mb][0] = pack_act2(kT[mb][0][0], kT[mb][0][1]);
        pa[mb][1] = pack_act2(kT[mb][0][2], kT[mb][0][3]);
        pa[mb][2] = pack_act2(kT[mb][1][0], kT[mb][1][1]);
        pa[mb][3] = pack_act2(kT[mb][1][2], kT[mb][1][3]);
#pragma unroll
        for (int eb = 0; eb < 4; ++eb) {
          ctx[mb][eb][0] *= alpha[0]; ctx[mb][eb][1] *= alpha[0];
          ctx[mb][eb][2] *= alpha[1]; ctx[mb][eb][3] *= alpha[1];
        }
      }
      // ctx[d][e] += sum_pix P[d][pix] v[e][pix]: B fragments straight from the v^T accumulators
#pragma unroll
      for (int eb = 0; eb < 4; ++eb) {
        const int mbv = eb >> 1, jr = (eb & 1) * 2;  // e block = rows g (+8) of v^T block mbv
        const uint32_t b0 = pack_act2(vT[mbv][0][jr], vT[mbv][0][jr + 1]);
        const uint32_t b1 = pack_act2(vT[mbv][1][jr], vT[mbv][1][jr + 1]);
        mma16816(ctx[0][eb], pa[0], b0, b1);
        mma16816(ctx[1][eb], pa[1], b0, b1);
      }
    }
  }
  // ---- the two pixel halves of every head are combined through shared memory, then one partial per (row, chunk, head)
  __syncthreads();
  float* sP = reinterpret_cast<float*>(sX);  // [8 warps][PART]
  float* mine = sP + (size_t)warp * PART;
#pragma unroll
  for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int d = mb * 16 + g + 8 * j;
      if (t == 0) { mine[d] = m_run[mb][j]; mine[DH + d] = l_run[mb][j]; }
#pragma unroll
      for (int eb = 0; eb < 4; ++eb) {
        mine[2 * DH + d * DH + eb * 8 + 2 * t] = ctx[mb][eb][2 * j];
        mine[2 * DH + d * DH + eb * 8 + 2 * t + 1] = ctx[mb][eb][2 * j + 1];
      }
    }
  }
  __syncthreads();
  float* out = p.part + (((size_t)row * p.chunks + chunk) * HEADS) * PART;
  for (int i = threadIdx.x; i < HEADS * DH * (DH + 2); i += FTHREADS) {
    const int hh = i / (DH * (DH + 2)), r = i - hh * DH * (DH + 2);
    const int d = r / (DH + 2), c = r - d * (DH + 2);  // c = 0: max, 1: sum, 2..: ctx[d][c - 2]
    const float* a = sP + (size_t)hh * PART;           // pixel half 0 (warp hh)
    const float* b = sP + (size_t)(4 + hh) * PART;     // pixel half 1 (warp 4 + hh)
    const float ma = a[d], mb2 = b[d], mm = fmaxf(ma, mb2);
    const float wa = ma == -INFINITY ? 0.f : __expf(ma - mm), wb = mb2 == -INFINITY ? 0.f : __expf(mb2 - mm);
    float val;
    if (c == 0) val = mm;
    else if (c == 1) val = a[DH + d] * wa + b[DH + d] * wb;
    else val = a[2 * DH + d * DH + c - 2] * wa + b[2 * DH + d * DH + c - 2] * wb;
    float* o = out + (size_t)hh * PART;
    if (c == 0) o[d] = val;
    else if (c == 1) o[DH + d] = val;
    else o[2 * DH + d * DH + c - 2] = val;
  }
}

// ------------------------------------------------------------------------------------------------ pass 2
// ctx = sum_chunks e^(m_c - m) ctx_c / (sum_chunks e^(m_c - m) l_c) / n per (row, head), stored transposed ([e][d], d
// contiguous: the B operand of pass 2) as 16-bit values in the first bytes of the row's partial area
__global__ void __launch_bounds__(256) linattn_ctx_combine_kernel(const LinAttnFusedParams p) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.y, hh = blockIdx.x;
  const float* base = p.part + ((size_t)row * p.chunks * HEADS + hh) * PART;
  act_t* out = p.ctx16 + ((size_t)row * HEADS + hh) * DH * DH;
  constexpr int MAXC = 8;  // linattn_fused_chunks() <= 8: all loads of an element are issued before the first use
  for (int i = threadIdx.x; i < DH * DH; i += blockDim.x) {
    const int d = i / DH, e = i - d * DH;
    float mc[MAXC], lc[MAXC], xc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const float* pc = base + (size_t)c * HEADS * PART;
      const bool on = c < p.chunks;
      mc[c] = on ? __ldg(pc + d) : -INFINITY;
      lc[c] = on ? __ldg(pc + DH + d) : 0.f;
      xc[c] = on ? __ldg(pc + 2 * DH + d * DH + e) : 0.f;
    }
    float mm = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) mm = fmaxf(mm, mc[c]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {  // (same order and arithmetic as a loop over p.chunks: absent chunks add exact zeros)
      if (c < p.chunks) {
        const float wgt = mc[c] == -INFINITY ? 0.f : __expf(mc[c] - mm);
        num += wgt * xc[c];
        den += wgt * lc[c];
      }
    }
    out[e * DH + d] = f2act(num / (den * (float)p.n));
  }
}

// warp = 16 pixels of a tile; a CTA walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ... of one row, so the weights and
// the context are staged once per CTA.  Shared memory: W_q [128][C + 8] | W_out [C][128 + 8] | ctx^T [4][32 e][32 d + 8] | tile.
template <int C>
__global__ void __launch_bounds__(FTHREADS) linattn_out_fused_kernel(const LinAttnFusedParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int LD = C + 8, LDO = HID + 8, LDC = DH + 8;
  act_t* sWq = reinterpret_cast<act_t*>(smem_raw);
  act_t* sWo = sWq + HID * LD;
  act_t* sC = sWo + C * LDO;
  act_t* sX = sC + HEADS * DH * LDC;
  const int row = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  pdl_trigger();
  load_weight_rows<HID, C>(p.w_qkv, p.ldw, sWq);       // q rows of to_qkv.weight
  load_weight_rows<C, HID>(p.w_out, p.ldw_out, sWo);   // to_out.weight [C][128]
  pdl_wait();  // (static weights staged before the wait; the context and x are outputs of the previous kernels)
  {  // ctx^T of the row (linattn_ctx_combine_kernel): [4 * 32 e][32 d] -> padded rows
    const act_t* src = p.ctx16 + (size_t)row * HEADS * DH * DH;
    for (int i = threadIdx.x; i < HEADS * DH * (DH / 8); i += FTHREADS) {
      const int r = i / (DH / 8), c = i - r * (DH / 8);
      *reinterpret_cast<uint4*>(sC + (size_t)r * LDC + c * 8) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * DH) + c);
    }
  }
  const int ntiles = (p.n + TP - 1) / TP;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
  const int pix0 = tile * TP;
  __syncthreads();  // the previous tile's rows have been copied out
  load_ln_tile<C>(p, row, pix0, p.n, sX);
  __syncthreads();
  const int pl = warp * 16;
  if (pix0 + pl < p.n) {
    // ---- q = y . W_q^T  [16 pixels x 128]
    float q[16][4];
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int f = 0; f < 4; ++f) q[j][f] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C / 16; ++ks) {
      uint32_t a[4];
      ldsm4(a, a_addr(sX, LD, pl, ks * 16, lane));
#pragma unroll
      for (int jb = 0; jb < 8; ++jb) {
        uint32_t b[4];
        ldsm4(b, b_addr(sWq, LD, jb * 16, ks * 16, lane));
        mma16816(q[2 * jb], a, b[0], b[1]);
        mma16816(q[2 * jb + 1], a, b[2], b[3]);
      }
    }
    // ---- softmax over the 32 channels of every head (rows g and g + 8), times dh^-0.5; out = q~ . ctx per head
    float o[16][4];
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int f = 0; f < 4; ++f) o[j][f] = 0.f;
    const float scale = rsqrtf((float)DH);
#pragma unroll
    for (int hh = 0; hh < HEADS; ++hh) {
#pragma unroll
      for (int jr = 0; jr < 2; ++jr) {  // row g (frags 0, 1) / row g + 8 (frags 2, 3)
        float mx = -INFINITY;
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) mx = fmaxf(mx, fmaxf(q[4 * hh + jb][2 * jr], q[4 * hh + jb][2 * jr + 1]));
        mx = quad_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int jb = 0; jb < 4; ++jb)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float ev = __expf(q[4 * hh + jb][2 * jr + e] - mx);
            q[4 * hh + jb][2 * jr + e] = ev;
            sum += ev;
          }
        const float inv = scale / quad_sum(sum);
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) { q[4 * hh + jb][2 * jr] *= inv; q[4 * hh + jb][2 * jr + 1] *= inv; }
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {  // K = d in two steps of 16
        uint32_t a[4];
        a[0] = pack_act2(q[4 * hh + 2 * ks][0], q[4 * hh + 2 * ks][1]);
        a[1] = pack_act2(q[4 * hh + 2 * ks][2], q[4 * hh + 2 * ks][3]);
        a[2] = pack_act2(q[4 * hh + 2 * ks + 1][0], q[4 * hh + 2 * ks + 1][1]);
        a[3] = pack_act2(q[4 * hh + 2 * ks + 1][2], q[4 * hh + 2 * ks + 1][3]);
#pragma unroll
        for (int eb = 0; eb < 2; ++eb) {  // N = e in two blocks of 16
          uint32_t b[4];
          ldsm4(b, b_addr(sC + (size_t)hh * DH * LDC, LDC, eb * 16, ks * 16, lane));
          mma16816(o[4 * hh + 2 * eb], a, b[0], b[1]);
          mma16816(o[4 * hh + 2 * eb + 1], a, b[2], b[3]);
        }
      }
    }
    // ---- A fragments of `out` for the output projection (K = 128 in 8 steps)
    uint32_t oa[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      oa[ks][0] = pack_act2(o[2 * ks][0], o[2 * ks][1]);
      oa[ks][1] = pack_act2(o[2 * ks][2], o[2 * ks][3]);
      oa[ks][2] = pack_act2(o[2 * ks + 1][0], o[2 * ks + 1][1]);
      oa[ks][3] = pack_act2(o[2 * ks + 1][2], o[2 * ks + 1][3]);
    }
    __syncwarp();  // every lane has read its q operands from the tile rows this warp now overwrites with y
    // ---- y = out . W_out^T + bias, 64 output channels at a time, staged in this warp's rows of the tile
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 64) {
      float yv[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int f = 0; f < 4; ++f) yv[j][f] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
          uint32_t b[4];
          ldsm4(b, b_addr(sWo, LDO, c0 + jb * 16, ks * 16, lane));
          mma16816(yv[2 * jb], oa[ks], b[0], b[1]);
          mma16816(yv[2 * jb + 1], oa[ks], b[2], b[3]);
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j * 8 + 2 * t;
        const float b0 = __ldg(p.b_out + c), b1 = __ldg(p.b_out + c + 1);
        *reinterpret_cast<uint32_t*>(sX + (size_t)(pl + g) * LD + c) = pack_act2(yv[j][0] + b0, yv[j][1] + b1);
        *reinterpret_cast<uint32_t*>(sX + (size_t)(pl + g + 8) * LD + c) = pack_act2(yv[j][2] + b0, yv[j][3] + b1);
      }
    }
    __syncwarp();
    // ---- + residual x, 128-bit coalesced stores of this warp's 16 pixels
    constexpr int PER = C / 8;
    for (int i = lane; i < 16 * PER; i += 32) {
      const int r = i / PER, c = (i - r * PER) * 8;
      const int pix = pix0 + pl + r;
      if (pix >= p.n) continue;
      float a[8], b[8];
      unpack8(*reinterpret_cast<const uint4*>(sX + (size_t)(pl + r) * LD + c), a);
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + ((size_t)row * p.n + pix) * C + c)), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      *reinterpret_cast<uint4*>(p.y + ((size_t)row * p.n + pix) * C + c) = pack8(a);
    }
  }
  }
}

// ------------------------------------------------------------------------------------------------ full attention
// Bottleneck attention (reference: src/models/modules/attention.py:52-73) on the tensor cores: one CTA per (head, row);
// Q, K ([n][32]) and V^T ([32][n]) of the head in shared memory; a warp owns 16 queries at a time: S = Q K^T (accumulator
// fragments = the whole score row block in registers), softmax over the keys with quad shuffles, dropout on the
// probabilities, then the probabilities re-packed as A fragments of O = P V.  NP = keys rounded up to a multiple of 16.
template <int NP>
__global__ void __launch_bounds__(FTHREADS) attention_mma_kernel(const AttnParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int LDQ = DH + 8, LDV = NP + 8, NB = NP / 8;
  act_t* sQ = reinterpret_cast<act_t*>(smem_raw);   // [NP][LDQ], pre-scaled by dh^-0.5
  act_t* sK = sQ + NP * LDQ;                        // [NP][LDQ]
  act_t* sV = sK + NP * LDQ;                        // [DH][LDV] (V transposed: keys contiguous)
  uint8_t* sMask = reinterpret_cast<uint8_t*>(sV + DH * LDV);  // [8 warps][2 rows][40]: keep bits of 8-element groups
  const int h = blockIdx.x, r = blockIdx.y, n = p.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int ld = 3 * p.heads * DH;
  const act_t* base = p.qkv + (size_t)r * n * ld + h * DH;
  const float scale = rsqrtf((float)DH);
  for (int i = threadIdx.x; i < NP * (DH / 8); i += FTHREADS) {
    const int j = i / (DH / 8), c = (i - j * (DH / 8)) * 8;
    float q[8], k[8], v[8];
    if (j < n) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)j * ld + c)), q);
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)j * ld + p.heads * DH + c)), k);
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)j * ld + 2 * p.heads * DH + c)), v);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) { q[e] = 0.f; k[e] = 0.f; v[e] = 0.f; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] *= scale;
    *reinterpret_cast<uint4*>(sQ + (size_t)j * LDQ + c) = pack8(q);
    *reinterpret_cast<uint4*>(sK + (size_t)j * LDQ + c) = pack8(k);
#pragma unroll
    for (int e = 0; e < 8; ++e) sV[(size_t)(c + e) * LDV + j] = f2act(v[e]);
  }
  __syncthreads();
  const uint64_t row_elems = ((uint64_t)p.heads * n * n + 7) & ~7ull;
  const DropRow dr = drop_row(p.drop, r, row_elems);
  for (int mb = warp; mb < NP / 16; mb += FTHREADS / 32) {
    // ---- S = Q K^T for queries 16 mb .. 16 mb + 15 and all keys
    float sc[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int f = 0; f < 4; ++f) sc[nb][f] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[4];
      ldsm4(a, a_addr(sQ, LDQ, mb * 16, ks * 16, lane));
#pragma unroll
      for (int nb2 = 0; nb2 < NB / 2; ++nb2) {
        uint32_t b[4];
        ldsm4(b, b_addr(sK, LDQ, nb2 * 16, ks * 16, lane));
        mma16816(sc[2 * nb2], a, b[0], b[1]);
        mma16816(sc[2 * nb2 + 1], a, b[2], b[3]);
      }
    }
    // ---- softmax over the keys (rows g and g + 8 of the block), dropout on the probabilities
    float inv[2];
#pragma unroll
    for (int jr = 0; jr < 2; ++jr) {
      float mx = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) if (nb * 8 + 2 * t + e < n) mx = fmaxf(mx, sc[nb][2 * jr + e]);
      mx = quad_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float ev = nb * 8 + 2 * t + e < n ? __expf(sc[nb][2 * jr + e] - mx) : 0.f;
          sc[nb][2 * jr + e] = ev;
          sum += ev;
        }
      inv[jr] = 1.f / quad_sum(sum);
    }
    if (p.drop.thresh) {
      // element = ((head, query), key) of the row's [heads][n][n] probability tensor.  One Philox draw yields the keep bits of
      // 8 consecutive elements: the (<= NP / 8 + 1) draws a query row needs are spread over the lanes of the quad-octet that
      // owns the row (lane = 4 g' + t computes groups t, t + 4, ... of row g'), parked in shared memory, then picked up per
      // element -- instead of one draw per element pair
      uint8_t* const mk = sMask + warp * 2 * 8 * 40;  // [row half][g][40 groups]
      __syncwarp();
#pragma unroll
      for (int jr = 0; jr < 2; ++jr) {
        const uint64_t e0 = ((uint64_t)h * n + (mb * 16 + g + 8 * jr)) * n;
        for (int gi = t; gi <= NP / 8; gi += 4)
          mk[(jr * 8 + g) * 40 + gi] = (uint8_t)drop_keep_bits8(p.drop, dr, (e0 & ~7ull) + (uint64_t)gi * 8);
      }
      __syncwarp();
#pragma unroll
      for (int jr = 0; jr < 2; ++jr) {
        const uint32_t off = (uint32_t)((((uint64_t)h * n + (mb * 16 + g + 8 * jr)) * n) & 7);
        const uint8_t* row = mk + (jr * 8 + g) * 40;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const uint32_t el = off + nb * 8 + 2 * t;  // element offset from the row's first (8-aligned) group
          const uint32_t k0 = row[el >> 3], k1 = row[(el + 1) >> 3];
          sc[nb][2 * jr] = ((k0 >> (el & 7)) & 1u) ? sc[nb][2 * jr] * p.drop.scale : 0.f;
          sc[nb][2 * jr + 1] = ((k1 >> ((el + 1) & 7)) & 1u) ? sc[nb][2 * jr + 1] * p.drop.scale : 0.f;
        }
      }
    }
    // ---- O = P V  (K axis = keys in steps of 16, B = V^T rows)
    float o[4][4];
#pragma unroll
    for (int eb = 0; eb < 4; ++eb)
#pragma unroll
      for (int f = 0; f < 4; ++f) o[eb][f] = 0.f;
#pragma unroll
    for (int ks = 0; ks < NP / 16; ++ks) {
      uint32_t a[4];
      a[0] = pack_act2(sc[2 * ks][0], sc[2 * ks][1]);
      a[1] = pack_act2(sc[2 * ks][2], sc[2 * ks][3]);
      a[2] = pack_act2(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
      a[3] = pack_act2(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
#pragma unroll
      for (int eb2 = 0; eb2 < 2; ++eb2) {
        uint32_t b[4];
        ldsm4(b, b_addr(sV, LDV, eb2 * 16, ks * 16, lane));
        mma16816(o[2 * eb2], a, b[0], b[1]);
        mma16816(o[2 * eb2 + 1], a, b[2], b[3]);
      }
    }
#pragma unroll
    for (int jr = 0; jr < 2; ++jr) {
      const int i = mb * 16 + g + 8 * jr;
      if (i >= n) continue;
      act_t* op = p.out + ((size_t)r * n + i) * (p.heads * DH) + h * DH;
#pragma unroll
      for (int eb = 0; eb < 4; ++eb)
        *reinterpret_cast<uint32_t*>(op + eb * 8 + 2 * t) = pack_act2(o[eb][2 * jr] * inv[jr], o[eb][2 * jr + 1] * inv[jr]);
    }
  }
}

template <int NP>
int launch_attention_mma_t(const AttnParams& p, cudaStream_t s) {
  const size_t smem = ((size_t)2 * NP * (DH + 8) + (size_t)DH * (NP + 8)) * sizeof(act_t) + 8 * 2 * 8 * 40;
  static bool configured = false;
  if (!configured) {
    DYF_CUDA_OK(cudaFuncSetAttribute(attention_mma_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  ProfScope prof(s, KC_ATTENTION, 4.0 * p.rows * p.heads * (double)p.n * p.n * DH);
  DYF_LAUNCH_PDL(3, "attention_mma_kernel", attention_mma_kernel<NP>, dim3(p.heads, p.rows), dim3(FTHREADS), smem, s, p);
  return 0;
}

template <int C>
int launch_fused_t(const LinAttnFusedParams& p, cudaStream_t s) {
  constexpr int LD = C + 8;
  const size_t smem1 = ((size_t)2 * HID * LD + (size_t)TP * LD) * sizeof(act_t);
  const size_t smem1b = std::max(smem1, (size_t)2 * HID * LD * sizeof(act_t) + (size_t)8 * PART * sizeof(float));
  const size_t smem2 = ((size_t)HID * LD + (size_t)C * (HID + 8) + (size_t)HEADS * DH * (DH + 8) + (size_t)TP * LD) * sizeof(act_t);
  static bool configured = false;
  if (!configured) {
    DYF_CUDA_OK(cudaFuncSetAttribute(linattn_ctx_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1b));
    DYF_CUDA_OK(cudaFuncSetAttribute(linattn_out_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    configured = true;
  }
  const double flops = 2.0 * p.rows * (double)p.n * (3.0 * HID * C + 2.0 * HEADS * DH * DH + (double)HID * C);
  const double bytes = 2.0 * p.rows * (double)p.n * C * 3.0;
  ProfScope prof(s, KC_ATTENTION, flops, bytes);
  DYF_LAUNCH_PDL(3, "linattn_ctx_fused_kernel", linattn_ctx_fused_kernel<C>, dim3(p.chunks, p.rows), dim3(FTHREADS), smem1b, s, p);
  DYF_LAUNCH_PDL(3, "linattn_ctx_combine_kernel", linattn_ctx_combine_kernel, dim3(HEADS, p.rows), dim3(256), 0, s, p);
  // CTAs per row: enough of them for ~4 resident waves, each walking several tiles with the weights staged once
  const int ntiles = cdiv(p.n, TP);
  const int per_row = std::max(1, std::min(ntiles, cdiv(4 * 2 * 148, p.rows)));
  DYF_LAUNCH_PDL(3, "linattn_out_fused_kernel", linattn_out_fused_kernel<C>, dim3(per_row, p.rows), dim3(FTHREADS), smem2, s, p);
  return 0;
}

}  // namespace

// tensor-core bottleneck attention for n <= 256 keys (score row blocks live in registers); 1 = launched, 0 = not eligible
int launch_attention_mma(const AttnParams& p, cudaStream_t s) {
  static const bool off = getenv("DYF_DISABLE_ATTN_MMA") != nullptr;
  if (off || p.heads != HEADS || p.n > 256) return 0;
  int rc;
  if (p.n <= 64) rc = launch_attention_mma_t<64>(p, s);
  else if (p.n <= 144) rc = launch_attention_mma_t<144>(p, s);
  else if (p.n <= 240) rc = launch_attention_mma_t<240>(p, s);
  else rc = launch_attention_mma_t<256>(p, s);
  return rc ? rc : 1;
}

// pixel chunks per row of pass 1 (a partial per chunk): one 128-pixel tile each, at most 8.  The chunking depends on the grid
// only -- never on the row count -- so a rank of a sharded job computes bit for bit what the full batch computes.
int linattn_fused_chunks(int n) { return std::max(1, std::min(8, (n + TP - 1) / TP)); }
size_t linattn_fused_scratch_floats(int) { return (size_t)8 * HEADS * PART + HEADS * DH * DH / 2; }
bool linattn_fused_shape_ok(int C, int heads) { return heads == HEADS && (C == 64 || C == 128 || C == 256); }

int launch_linattn_fused(const LinAttnFusedParams& p0, cudaStream_t s) {
  LinAttnFusedParams p = p0;
  p.chunks = linattn_fused_chunks(p.n);
  p.chunk_pix = (((p.n + p.chunks - 1) / p.chunks) + 15) & ~15;
  p.ctx16 = reinterpret_cast<act_t*>(p.part + (size_t)p.rows * p.chunks * HEADS * PART);
  switch (p.C) {
    case 64: return launch_fused_t<64>(p, s);
    case 128: return launch_fused_t<128>(p, s);
    case 256: return launch_fused_t<256>(p, s);
    default: set_error("fused linear attention: channel count must be 64, 128 or 256"); return -1;
  }
}

}  // namespace dyf
