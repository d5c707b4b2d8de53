// Decoder block of the Navier-Stokes backbone as ONE tensor-core kernel: bilinear x2 upsample (align_corners=False) of
// the concatenated [x, skip] maps followed by the 3x3 / pad-1 convolution (reference: src/models/unet_simple.py:41-51,
// :133-140), without ever materialising the upsampled tensor.
//
// Algebra.  Bilinear x2 is a fixed separable stencil, u[2i] = .25 x[i-1] + .75 x[i], u[2i+1] = .75 x[i] + .25 x[i+1]
// (indices clamped), so a 3x3 conv over u is, per output parity class (a, b) = (row & 1, col & 1), a 3x3 conv over the
// LOW-resolution map x with composite weights  Wc[a,b][di,dj] = sum_{ky,kx} V_a[di,ky] H_b[dj,kx] w[ky,kx].
// Stacking the four classes along the GEMM N axis turns the block into a plain 3x3 conv Cin -> 4*Cout on the low-res
// grid whose epilogue stores depth-to-space: same FLOPs as the reference's conv, a quarter of its activation traffic,
// N = 128 tiles even for the Cout = 64 layer, and no upsample kernel.
//
// Borders.  The reference zero-pads the UPSAMPLED map and clamps the interpolation; on the first/last low-res row or
// column this gives different 1-D maps (V_top, V_bottom, H_left, H_right -- computed on the host by simulating the
// stencil), i.e. the border pixels are the same conv with other weights.  They are therefore computed by their own
// tiles: ROW tiles (1 x 128 pixels of the first/last row), COL tiles (128 x 1), and the four corners per image by a
// small CUDA-core kernel; MAIN tiles skip the border ring.  Every output pixel is written exactly once.
//
// Pipeline = conv_umma.cu's: persistent CTAs, TMA box loads of the low-res patch (zero fill = the low-res taps that
// fall outside the image, whose composite weights are irrelevant), weights by bulk TMA, tcgen05.mma into two
// ping-pong TMEM accumulator sets, 8 epilogue warps.
#include <cstdlib>
#include <map>
#include <tuple>

#include "conv.cuh"
#include "umma.cuh"

namespace dyf {
namespace {

constexpr int BN = 128;                                  // GEMM N tile (columns of 4*Cout)
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 1 + 1 + 1) * 32;    // epilogue + A producer + MMA + weight producer warps
constexpr int B_TAP = BN * 64 * 2;                       // one tap of one 64-channel chunk: [k8][BN][8] bf16

// Tile kinds.  A GEMM row group (8 rows = one 1024-byte swizzle atom of the patch) is
//   MAIN    8 consecutive pixels of one image      (16 x 16-pixel super-tile = two 16 x 8 M-tiles per weight stage)
//   ROW/COL 8 consecutive IMAGES at one pixel       (M-tile = 16 border pixels x 8 images: the patch is loaded through
//           a tensor map whose dimension order is (C, image, W, H) resp. (C, image, H, W), so that small images fill the
//           tile as well as large ones)
//   CORNER  8 consecutive images at the corner pixel (M-tile = 128 images; only the 2 x 2 in-bounds taps are loaded)
enum Kind { K_MAIN = 0, K_ROW = 1, K_COL = 2, K_CORNER = 3, N_KINDS = 4 };

template <int KIND> struct UGeo;
template <> struct UGeo<K_MAIN> {
  static constexpr int T = 2, NTAP = 9, A_BYTES = 18 * 18 * 128, A_STAGE = ((A_BYTES + 1023) / 1024) * 1024;
  static constexpr int SBO = 18 * 128, AS = 2, BS = 8;
  __device__ static constexpr int tap_off(int t) { return ((t / 3) * 18 + t % 3) * 128; }
};
template <> struct UGeo<K_ROW> {   // patch [3 rows][18 pixels][8 images] x 128 B
  static constexpr int T = 1, NTAP = 9, A_BYTES = 3 * 18 * 8 * 128, A_STAGE = A_BYTES;
  static constexpr int SBO = 1024, AS = 2, BS = 7;
  __device__ static constexpr int tap_off(int t) { return ((t / 3) * 18 + t % 3) * 1024; }
};
template <> struct UGeo<K_COL> {   // patch [3 columns][18 pixels][8 images] x 128 B
  static constexpr int T = 1, NTAP = 9, A_BYTES = 3 * 18 * 8 * 128, A_STAGE = A_BYTES;
  static constexpr int SBO = 1024, AS = 2, BS = 7;
  __device__ static constexpr int tap_off(int t) { return ((t % 3) * 18 + t / 3) * 1024; }
};
template <> struct UGeo<K_CORNER> {  // patch [2 rows][2 columns][128 images] x 128 B
  static constexpr int T = 1, NTAP = 4, A_BYTES = 4 * 128 * 128, A_STAGE = A_BYTES;
  static constexpr int SBO = 1024, AS = 2, BS = 5;
  __device__ static constexpr int tap_off(int t) { return t * 128 * 128; }
};

// One launch runs the four tile kinds back to back as phases of the same persistent CTAs (a CTA that runs out of MAIN
// tiles moves on to the border tiles without waiting for the others).  The barrier block and the epilogue tables sit at
// a fixed offset behind the largest A/B ring; barriers are re-initialised between phases.
constexpr int MAX_AS = 2, MAX_BS = 8;
struct __align__(16) UBarriers {
  uint64_t a_full[MAX_AS], a_empty[MAX_AS], b_full[MAX_BS], b_empty[MAX_BS], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
constexpr int cmax(int a, int b) { return a > b ? a : b; }
template <int KIND> constexpr int ring_bytes() { return UGeo<KIND>::AS * UGeo<KIND>::A_STAGE + UGeo<KIND>::BS * B_TAP; }
constexpr int BAR_OFF = cmax(cmax(ring_bytes<K_MAIN>(), ring_bytes<K_ROW>()), cmax(ring_bytes<K_COL>(), ring_bytes<K_CORNER>()));
constexpr int TAB_OFF = BAR_OFF + (((int)sizeof(UBarriers) + 15) & ~15);
constexpr int SMEM_BYTES = TAB_OFF + 4 * BN * 4 + 64;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
struct UpWork { int tiles_a[N_KINDS], num_work[N_KINDS]; };  // per kind: tile blocks along the tiled axis, work items
struct UpMaps { CUtensorMap m[N_KINDS][2]; };                // per kind, per source

// Fused epilogue of NG groups of 8 consecutive GEMM columns starting at column n (each group = one parity class, channels
// co..co+7) of low-res pixel (i, j).  tA / tB point at the table entries of column n (global memory, or the per-item copy
// in shared memory).  Straight-line code: the NG chains interleave.
template <int NG>
__device__ __forceinline__ void up_store(const UpConvParams& p, int img, int i, int j, int n, float* y, const float* tA,
                                         const float* tB) {
#pragma unroll
  for (int q = 0; q < 2 * NG; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(tA + 4 * q), b = *reinterpret_cast<const float4*>(tB + 4 * q);
    y[4 * q + 0] = fmaf(y[4 * q + 0], a.x, b.x); y[4 * q + 1] = fmaf(y[4 * q + 1], a.y, b.y);
    y[4 * q + 2] = fmaf(y[4 * q + 2], a.z, b.z); y[4 * q + 3] = fmaf(y[4 * q + 3], a.w, b.w);
  }
  if (p.act <= ACT_LEAKY) {  // identity / ReLU / LeakyReLU(0.2) = max(y, slope * y)
    const float slope = p.act == ACT_NONE ? 1.f : p.act == ACT_RELU ? 0.f : 0.2f;
#pragma unroll
    for (int e = 0; e < 8 * NG; ++e) y[e] = fmaxf(y[e], slope * y[e]);
  } else {
#pragma unroll
    for (int e = 0; e < 8 * NG; ++e) y[e] = apply_act(y[e], p.act);
  }
  long long m[NG];
  int co[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int ng = n + 8 * g, cls = ng / p.Cout;
    co[g] = ng - cls * p.Cout;
    // hi-res pixel index (also the dropout element order of the two-kernel path)
    m[g] = ((long long)img * (2 * p.H) + 2 * i + (cls >> 1)) * (2 * p.W) + 2 * j + (cls & 1);
  }
  if (p.drop.thresh) {
    const uint64_t per_img = (uint64_t)(4 * p.H * p.W) * p.Cout;
    const DropRow dr = drop_row(p.drop, img, per_img);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const uint32_t keep = drop_keep_bits8(p.drop, dr, (uint64_t)(m[g] - (long long)img * (4 * p.H * p.W)) * p.Cout + co[g]);
#pragma unroll
      for (int e = 0; e < 8; ++e) y[8 * g + e] = ((keep >> e) & 1u) ? y[8 * g + e] * p.drop.scale : 0.f;
    }
  }
  const bool wide_st = (p.out_ld & 15) == 0;
#pragma unroll
  for (int g = 0; g < NG; g += 2) {  // groups g, g + 1 are 16 consecutive channels of one output pixel unless a class ends between them
    act_t* const o0 = p.out + (size_t)m[g] * p.out_ld + co[g];
    if (g + 1 < NG && wide_st && m[g + 1] == m[g] && co[g + 1] == co[g] + 8 && (co[g] & 15) == 0) {
      st_global_256(o0, pack8(y + 8 * g), pack8(y + 8 * g + 8));
    } else {
      *reinterpret_cast<uint4*>(o0) = pack8(y + 8 * g);
      if (g + 1 < NG) *reinterpret_cast<uint4*>(p.out + (size_t)m[g + 1] * p.out_ld + co[g + 1]) = pack8(y + 8 * g + 8);
    }
  }
}

template <int KIND>
__device__ __forceinline__ void up_phase(const UpConvParams& p, uint8_t* smem, uint32_t tmem_base, int tiles_a, int n_tiles,
                                         int num_work, int first, const CUtensorMap* tmap0, const CUtensorMap* tmap1) {
  using G = UGeo<KIND>;
  constexpr int AS = G::AS, BS = G::BS, T = G::T, A_STAGE = G::A_STAGE, NTAP = G::NTAP;
  static_assert(AS <= MAX_AS && BS <= MAX_BS, "barrier block too small");
  uint8_t* sA = smem;
  uint8_t* sB = smem + AS * A_STAGE;
  UBarriers* bars = reinterpret_cast<UBarriers*>(smem + BAR_OFF);
  float* sTab = reinterpret_cast<float*>(smem + TAB_OFF);  // [2 acc][A | B][BN] epilogue tables (MAIN tiles)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nch0 = p.C[0] >> 6, nchunks = (p.C[0] + p.C[1]) >> 6;

  // work item -> (n_tile, first image, pixel origin, side).  tiles_a = 16-pixel blocks along the tiled axis; `side`:
  // ROW / COL 0 = first row / column, 1 = last; CORNER bit 1 = bottom, bit 0 = right.
  auto decode = [&](int w, int& n_tile, int& img, int& i0, int& j0, int& side) {
    n_tile = w % n_tiles;
    int t = w / n_tiles;
    if constexpr (KIND == K_MAIN) {
      const int tiles_y = (p.H + 15) >> 4, per_img = tiles_a * tiles_y;
      img = t / per_img;
      t -= img * per_img;
      const int ty = t / tiles_a;
      i0 = ty * 16;
      j0 = (t - ty * tiles_a) * 16;
      side = 0;
    } else if constexpr (KIND == K_CORNER) {
      side = t & 3;
      img = (t >> 2) * 128;
      i0 = (side & 2) ? p.H - 1 : 0;
      j0 = (side & 1) ? p.W - 1 : 0;
    } else {
      const int blk = (t % tiles_a) * 16;
      t /= tiles_a;
      side = t & 1;
      img = (t >> 1) * 8;
      if constexpr (KIND == K_ROW) { i0 = side ? p.H - 1 : 0; j0 = blk; }
      else { j0 = side ? p.W - 1 : 0; i0 = blk; }
    }
  };

  if (warp < EPI_WARPS) {
    // =============================== epilogue: TMEM -> affine/activation/dropout -> depth-to-space stores ==========
    int it = 0, tab_key = -1, tab_buf = 0;
    const int quarter = warp & 3, m_local = quarter * 32 + lane;
    constexpr int COLS = BN / (EPI_WARPS / 4);
    const int cbeg = (warp >> 2) * COLS;
    for (int w = first; w < num_work; w += gridDim.x, ++it) {
      int n_tile, img, i0, j0, side;
      decode(w, n_tile, img, i0, j0, side);
      const int acc = it & 1;
      if constexpr (KIND == K_MAIN) {  // one image per item: its epilogue tables are staged in shared memory, re-staged
        const int tkey = (img / p.tab_div) * n_tiles + n_tile;  // (into the other buffer) only when (table row, n-tile) changes
        if (tkey != tab_key) {
          tab_key = tkey;
          tab_buf ^= 1;
          float* const dst = sTab + tab_buf * 2 * BN;
          if (tid < BN) {
            const size_t off = (size_t)(img / p.tab_div) * p.Cout + (n_tile * BN + tid) % p.Cout;
            dst[tid] = __ldg(p.tabA + off);
            dst[BN + tid] = __ldg(p.tabB + off);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        }
      }
      const float* const sA_tab = sTab + tab_buf * 2 * BN;
      mbar_wait(smem_u32(&bars->acc_full[acc]), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int tile = 0; tile < T; ++tile) {
        int im, i, j;
        bool valid;
        if constexpr (KIND == K_MAIN) {
          im = img; i = i0 + (m_local >> 3); j = j0 + tile * 8 + (m_local & 7);
          valid = i >= 1 && i <= p.H - 2 && j >= 1 && j <= p.W - 2;
        } else if constexpr (KIND == K_ROW) {
          im = img + (m_local & 7); i = i0; j = j0 + (m_local >> 3);
          valid = im < p.rows && j >= 1 && j <= p.W - 2;
        } else if constexpr (KIND == K_COL) {
          im = img + (m_local & 7); j = j0; i = i0 + (m_local >> 3);
          valid = im < p.rows && i >= 1 && i <= p.H - 2;
        } else {
          im = img + m_local; i = i0; j = j0;
          valid = im < p.rows;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc * T + tile) * BN;
#pragma unroll 1
        for (int cg = cbeg; cg < cbeg + COLS; cg += 32) {
          uint32_t v[32];
          tmem_ld32_nowait(taddr + cg, v);
          tmem_ld_wait();
          if (valid) {
            float y[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) y[e] = __uint_as_float(v[e]);
            const int n = n_tile * BN + cg;
            if constexpr (KIND == K_MAIN) {
              up_store<4>(p, im, i, j, n, y, sA_tab + cg, sA_tab + BN + cg);
            } else {  // several images per tile: tables straight from global memory (32 columns = one class: contiguous)
              const size_t off = (size_t)(im / p.tab_div) * p.Cout + n % p.Cout;
              up_store<4>(p, im, i, j, n, y, p.tabA + off, p.tabB + off);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->acc_empty[acc]));
    }
  } else if (warp == EPI_WARPS) {
    // =============================== A producer: one TMA box of the low-res [x | skip] patch per chunk ==============
    if (lane == 0) {
      int ca = 0;
      for (int w = first; w < num_work; w += gridDim.x) {
        int n_tile, img, i0, j0, side;
        decode(w, n_tile, img, i0, j0, side);
        // box origin in the coordinate order of this kind's tensor map (after the channel coordinate)
        int c1, c2, c3;
        if constexpr (KIND == K_MAIN) { c1 = j0 - 1; c2 = i0 - 1; c3 = img; }                 // (C, W, H, image)
        else if constexpr (KIND == K_ROW) { c1 = img; c2 = j0 - 1; c3 = i0 - 1; }               // (C, image, W, H)
        else if constexpr (KIND == K_COL) { c1 = img; c2 = i0 - 1; c3 = j0 - 1; }               // (C, image, H, W)
        else { c1 = img; c2 = (side & 1) ? p.W - 2 : 0; c3 = (side & 2) ? p.H - 2 : 0; }        // (C, image, W, H), 2 x 2 box
        for (int c = 0; c < nchunks; ++c, ++ca) {
          const int st = ca % AS;
          mbar_wait(smem_u32(&bars->a_empty[st]), ((ca / AS) & 1) ^ 1);
          const uint32_t bar = smem_u32(&bars->a_full[st]);
          mbar_expect_tx(bar, G::A_BYTES);
          const uint64_t tm = reinterpret_cast<uint64_t>(c < nch0 ? tmap0 : tmap1);
          const int cc = (c < nch0 ? c : c - nch0) * 64;
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
              ::"r"(smem_u32(sA + st * A_STAGE)), "l"(tm), "r"(cc), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // =============================== MMA issuer ====================================================================
    const uint32_t leader = elect_one();
    const uint32_t idesc = (1u << 4) | (DYF_UMMA_FMT << 7) | (DYF_UMMA_FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hi = (uint32_t)((G::SBO >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);  // SBO | version | SWIZZLE_128B
    const uint32_t b_hi = (uint32_t)((128 >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t a_lo0 = (1u << 16) | (smem_u32(sA) >> 4);
    const uint32_t b_lo0 = ((uint32_t)((BN * 16) >> 4) << 16) | (smem_u32(sB) >> 4);
    const uint32_t bar_a_full = smem_u32(&bars->a_full[0]), bar_a_empty = smem_u32(&bars->a_empty[0]);
    const uint32_t bar_b_full = smem_u32(&bars->b_full[0]), bar_b_empty = smem_u32(&bars->b_empty[0]);
    constexpr int AK = 32 >> 4, BK = (2 * BN * 16) >> 4;
    int sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
    for (int w = first; w < num_work; w += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(smem_u32(&bars->acc_empty[acc]), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * T * BN;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar_a_full + sa * 8, pa);
        tc_fence_after();
        const uint64_t a_st = ((uint64_t)a_hi << 32) | (a_lo0 + sa * (A_STAGE >> 4));
#pragma unroll
        for (int tap = 0; tap < NTAP; ++tap) {
          mbar_wait(bar_b_full + sb * 8, pb);
          tc_fence_after();
          const uint64_t b_st = ((uint64_t)b_hi << 32) | (b_lo0 + sb * (B_TAP >> 4));
          const int a_off = G::tap_off(tap);
#pragma unroll
          for (int tile = 0; tile < T; ++tile)
            umma_tap<4, AK, BK>(tmem_acc + tile * BN, a_st + (uint64_t)((a_off + tile * 8 * 128) >> 4), b_st, idesc,
                                tap ? 1u : (uint32_t)(c != 0), leader);
          umma_commit_if(bar_b_empty + sb * 8, leader);
          if (++sb == BS) { sb = 0; pb ^= 1; }
        }
        umma_commit_if(bar_a_empty + sa * 8, leader);
        if (++sa == AS) { sa = 0; pa ^= 1; }
      }
      umma_commit_if(smem_u32(&bars->acc_full[acc]), leader);
    }
  } else {
    // =============================== B producer: one bulk-TMA copy per (chunk, tap) weight tile ====================
    if (lane == 0) {
      int sb = 0, pb = 1;
      for (int w = first; w < num_work; w += gridDim.x) {
        int n_tile, img, i0, j0, side;
        decode(w, n_tile, img, i0, j0, side);
        const act_t* wv = KIND == K_MAIN ? p.w[0] : KIND == K_ROW ? p.w[1 + side] : KIND == K_COL ? p.w[3 + side] : p.w[5 + side];
        const uint8_t* src = reinterpret_cast<const uint8_t*>(wv) + (size_t)n_tile * nchunks * 9 * B_TAP;
        for (int c = 0; c < nchunks; ++c) {
#pragma unroll 1
          for (int t = 0; t < NTAP; ++t) {
            // CORNER: box tap (by, bx) is tap (by + !bottom, bx + !right) of the 3 x 3 composite
            const int tap9 = KIND == K_CORNER ? ((t >> 1) + ((side & 2) ? 0 : 1)) * 3 + (t & 1) + ((side & 1) ? 0 : 1) : t;
            mbar_wait(smem_u32(&bars->b_empty[sb]), pb);
            mbar_expect_tx(smem_u32(&bars->b_full[sb]), B_TAP);
            bulk_g2s(smem_u32(sB + sb * B_TAP), src + (size_t)(c * 9 + tap9) * B_TAP, B_TAP, smem_u32(&bars->b_full[sb]));
            if (++sb == BS) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) conv_up_kernel(const UpConvParams p, const UpWork wk, int n_tiles,
                                                             const __grid_constant__ UpMaps maps) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  UBarriers* bars = reinterpret_cast<UBarriers*>(smem + BAR_OFF);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == EPI_WARPS + 1) {  // TMEM: two accumulator sets of (up to) two 128-column tiles, owned by the MMA warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  unsigned done = 0;  // work items of the earlier phases
#pragma unroll 1
  for (int phase = 0; phase < N_KINDS; ++phase) {
    if (tid == 0) {
      for (int i = 0; i < MAX_AS; ++i) { mbar_init(smem_u32(&bars->a_full[i]), 1); mbar_init(smem_u32(&bars->a_empty[i]), 1); }
      for (int i = 0; i < MAX_BS; ++i) { mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), EPI_WARPS * 32); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    if (phase == 0) pdl_wait();  // TMEM allocation and barrier init overlapped the previous kernel's tail
    const CUtensorMap* m0 = &maps.m[phase][0];
    const CUtensorMap* m1 = &maps.m[phase][1];
    // The work items of all phases form one round-robin sequence over the CTAs (concurrent CTAs work on neighbouring
    // tiles: shared halos hit in L2 and the accesses spread over the DRAM partitions -- contiguous per-CTA slices were
    // measured 50 % slower); a phase starts at the CTA after the one that took the previous phase's last item, so CTAs with
    // one MAIN tile less pick up the border tiles first.
    const int first = (int)((blockIdx.x + gridDim.x - done % gridDim.x) % gridDim.x);
    if (phase == K_MAIN) up_phase<K_MAIN>(p, smem, tmem_base, wk.tiles_a[phase], n_tiles, wk.num_work[phase], first, m0, m1);
    else if (phase == K_ROW) up_phase<K_ROW>(p, smem, tmem_base, wk.tiles_a[phase], n_tiles, wk.num_work[phase], first, m0, m1);
    else if (phase == K_COL) up_phase<K_COL>(p, smem, tmem_base, wk.tiles_a[phase], n_tiles, wk.num_work[phase], first, m0, m1);
    else up_phase<K_CORNER>(p, smem, tmem_base, wk.tiles_a[phase], n_tiles, wk.num_work[phase], first, m0, m1);
    done += (unsigned)wk.num_work[phase];
    tc_fence_before();
    __syncthreads();  // every role of this CTA has drained the phase: barriers can be re-initialised
  }
  if (warp == EPI_WARPS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(bars->tmem_base), "r"(512));
  }
}

// Composite weights of one border variant: out[n = cls*Cout + co][ci][di][dj] = sum V[a][di][ky] H[b][dj][kx] w[co][ci][ky][kx]
struct Maps { float V[2][3][3], H[2][3][3]; };
__global__ void __launch_bounds__(256) compose_up_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout,
                                                        int Cin, Maps m) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)4 * Cout * Cin) return;
  const int ci = (int)(idx % Cin), n = (int)(idx / Cin);
  const int cls = n / Cout, co = n - cls * Cout, a = cls >> 1, b = cls & 1;
  float k[3][3];
#pragma unroll
  for (int t = 0; t < 9; ++t) k[t / 3][t % 3] = w[((size_t)co * Cin + ci) * 9 + t];
#pragma unroll
  for (int di = 0; di < 3; ++di)
#pragma unroll
    for (int dj = 0; dj < 3; ++dj) {
      float s = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) s = fmaf(m.V[a][di][ky] * m.H[b][dj][kx], k[ky][kx], s);
      out[((size_t)n * Cin + ci) * 9 + di * 3 + dj] = s;
    }
}
// 1-D composite maps by simulating the reference ops on basis vectors: M[a][d][k] = coefficient of w[k] * x[i + d - 1]
// in output 2i + a of (zero-padded 3-tap conv) o (clamped bilinear x2), for i at the start / interior / end of the axis.
void axis_maps(int where /*0 first, 1 interior, 2 last*/, float M[2][3][3], int nearest) {
  const int L = 5, i = where == 0 ? 0 : where == 1 ? 2 : L - 1;
  for (int a = 0; a < 2; ++a)
    for (int d = 0; d < 3; ++d)
      for (int k = 0; k < 3; ++k) {
        const int src = i + d - 1, r = 2 * i + a + k - 1;  // source pixel, upsampled position read by tap k
        double v = 0.0;
        if (nearest) {  // upsampled position r is a copy of source pixel r >> 1
          if (src >= 0 && src < L && r >= 0 && r < 2 * L && (r >> 1) == src) v = 1.0;
        } else if (src >= 0 && src < L && r >= 0 && r < 2 * L) {
          const int q = r >> 1;
          const int lo = r & 1 ? q : (q > 0 ? q - 1 : 0), hi = r & 1 ? (q + 1 < L ? q + 1 : L - 1) : q;
          const double wlo = r & 1 ? 0.75 : 0.25, whi = 1.0 - wlo;
          v = (lo == src ? wlo : 0.0) + (hi == src ? whi : 0.0);
        }
        M[a][d][k] = (float)v;
      }
}

// Tensor maps of one source for the four tile kinds (cached per buffer / geometry: the workspace carving is stable).
int source_maps(const UpConvParams& p, int s, CUtensorMap out[N_KINDS]) {
  using Key = std::tuple<const void*, int, int, int, int, int>;
  struct Entry { CUtensorMap m[N_KINDS]; };
  static std::map<Key, Entry> cache;
  const Key key{p.src[s], p.rows, p.H, p.W, p.C[s], p.ld[s]};
  auto it = cache.find(key);
  if (it == cache.end()) {
    Entry e;
    const cuuint64_t C = (cuuint64_t)p.C[s], N = (cuuint64_t)p.rows, W = (cuuint64_t)p.W, H = (cuuint64_t)p.H;
    const cuuint64_t pix = (cuuint64_t)p.ld[s] * 2, row = W * pix, img = H * row;
    const cuuint64_t d_main[4] = {C, W, H, N}, s_main[3] = {pix, row, img};
    const cuuint64_t d_row[4] = {C, N, W, H}, s_row[3] = {img, pix, row};
    const cuuint64_t d_col[4] = {C, N, H, W}, s_col[3] = {img, row, pix};
    const cuuint32_t b_main[4] = {64, 18, 18, 1}, b_edge[4] = {64, 8, 18, 3}, b_corner[4] = {64, 128, 2, 2};
    if (make_tmap4(p.src[s], d_main, s_main, b_main, &e.m[K_MAIN]) || make_tmap4(p.src[s], d_row, s_row, b_edge, &e.m[K_ROW]) ||
        make_tmap4(p.src[s], d_col, s_col, b_edge, &e.m[K_COL]) || make_tmap4(p.src[s], d_row, s_row, b_corner, &e.m[K_CORNER])) {
      set_error("conv_up: cuTensorMapEncodeTiled failed");
      return -1;
    }
    if (cache.size() > 1024) cache.clear();
    it = cache.emplace(key, e).first;
  }
  for (int k = 0; k < N_KINDS; ++k) out[k] = it->second.m[k];
  return 0;
}

}  // namespace

bool conv_up_shape_ok(int C0, int C1, int Cout, int H, int W) {
  return C0 > 0 && C0 % 64 == 0 && C1 % 64 == 0 && Cout % 32 == 0 && H >= 16 && W >= 16;
}
size_t conv_up_weight_elems(int Cin, int Cout) { return (size_t)4 * Cout * Cin * 9; }  // per variant (bf16 stage tiles)

// w: conv weight fp32 [Cout, Cin, 3, 3].  Fills the nine composite variants as tcgen05 stage tiles: interior, first row,
// last row, first column, last column, then the corners (top-left, top-right, bottom-left, bottom-right).
// `scratch` must hold 4*Cout*Cin*9 floats.
int launch_compose_up(const float* w, int Cout, int Cin, act_t* const* w_variants, float* scratch, cudaStream_t s, int nearest) {
  float ax[3][2][3][3];
  for (int k = 0; k < 3; ++k) axis_maps(k, ax[k], nearest);
  const long long total = (long long)4 * Cout * Cin;
  const int vks[DYF_UP_VARIANTS] = {1, 0, 2, 1, 1, 0, 0, 2, 2}, hks[DYF_UP_VARIANTS] = {1, 1, 1, 0, 2, 0, 2, 0, 2};
  for (int v = 0; v < DYF_UP_VARIANTS; ++v) {
    Maps m;
    memcpy(m.V, ax[vks[v]], sizeof(m.V));
    memcpy(m.H, ax[hks[v]], sizeof(m.H));
    compose_up_kernel<<<cdiv(total, 256), 256, 0, s>>>(w, scratch, Cout, Cin, m);
    DYF_LAUNCH_OK("compose_up_kernel");
    int rc = launch_repack_umma(scratch, w_variants[v], 4 * Cout, Cin, 3, 1, 1, 0, s);
    if (rc) return rc;
  }
  return 0;
}

int launch_conv_up(const UpConvParams& p, cudaStream_t stream) {
  if (!conv_up_shape_ok(p.C[0], p.C[1], p.Cout, p.H, p.W) || (p.out_ld & 7) || (p.ld[0] & 7) || (p.C[1] && (p.ld[1] & 7))) {
    set_error("conv_up: unsupported shape");
    return -1;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    DYF_CUDA_OK(cudaGetDevice(&dev));
    DYF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    DYF_CUDA_OK(cudaFuncSetAttribute(conv_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  UpMaps maps;
  CUtensorMap tm[2][N_KINDS];
  for (int s = 0; s < 2; ++s)
    if (source_maps(p, p.C[s] ? s : 0, tm[s])) return -1;  // single-source layers: the second map is never used
  for (int k = 0; k < N_KINDS; ++k) { maps.m[k][0] = tm[0][k]; maps.m[k][1] = tm[1][k]; }
  const int n_tiles = 4 * p.Cout / BN;
  UpWork wk;
  wk.tiles_a[K_MAIN] = (p.W + 15) / 16; wk.tiles_a[K_ROW] = (p.W + 15) / 16; wk.tiles_a[K_COL] = (p.H + 15) / 16; wk.tiles_a[K_CORNER] = 1;
  const long long work[N_KINDS] = {(long long)wk.tiles_a[K_MAIN] * ((p.H + 15) / 16) * p.rows * n_tiles,
                                   (long long)2 * wk.tiles_a[K_ROW] * ((p.rows + 7) / 8) * n_tiles,
                                   (long long)2 * wk.tiles_a[K_COL] * ((p.rows + 7) / 8) * n_tiles,
                                   (long long)4 * ((p.rows + 127) / 128) * n_tiles};
  for (int k = 0; k < N_KINDS; ++k) {
    if (work[k] > 0x7fffffffLL) { set_error("conv_up: too many tiles"); return -1; }
    wk.num_work[k] = (int)work[k];
  }
  const int grid = (int)(work[0] < num_sms ? work[0] : num_sms);
  const int Cin = p.C[0] + p.C[1];
  const double flops = 2.0 * 4.0 * p.rows * p.H * p.W * p.Cout * 9.0 * Cin;  // = the reference conv on the upsampled grid
  const double bytes = 2.0 * ((double)p.rows * p.H * p.W * (Cin + 4.0 * p.Cout) + 36.0 * p.Cout * Cin);
  ProfScope prof(stream, KC_CONV_UP, flops, bytes);
  DYF_LAUNCH_PDL(1, "conv_up_kernel", conv_up_kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, stream, p, wk, n_tiles, maps);
  return 0;
}

}  // namespace dyf
