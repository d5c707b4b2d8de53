// Implicit-GEMM convolutions on the 5th-generation tensor cores: tcgen05.mma (UMMA) with the accumulator in TMEM,
// weights streamed by bulk-TMA (cp.async.bulk + mbarrier complete_tx) and the activations staged ONCE per channel
// chunk as a halo patch that every filter tap re-reads through shifted UMMA shared-memory descriptors.
//
// Tile: 16 x 8 output pixels (M = 128) x BN output channels; T tiles side by side share every weight stage (L2 -> SM
// bandwidth, ~42 B/cycle/SM, is what a 128-row tile cannot afford to spend on its own weight stream).  Modes:
//   S1K3  3x3, stride 1, pad 1:  64-channel chunks; the patch of the 16 x 8T super-tile is ONE 4-D tensor-map box
//         (cp.async.bulk.tensor, 128-byte swizzle, hardware zero fill = padding); a tap is a start-address shift of
//         (ky * patch_width + kx) * 128 B of a SWIZZLE_128B K-major descriptor whose SBO is one patch row
//   S2K4  4x4, stride 2, pad 1:  64-channel chunks; one box per (row parity, column parity) VIEW of the input (tensor
//         maps with pixel/row pitch 2), 128-byte swizzle: inside a view the stride-2 walk is unit stride and tap (ky, kx)
//         is a shift of (ky>>1, kx>>1) inside view ((ky+1)&1, (kx+1)&1); the patch ring is view-granular (one stage = one
//         view of one chunk = 4 taps)
//   S1K1  1x1: the 16 x 8T pixel box itself, one "tap"
//   S2K2  2x2, stride 2, pad 0:  runs on the S1K1 kernel over two strided tensor-map views of the input (a free
//         space-to-depth): K = (ky, kx, ci)
// A cp.async gather path (K-major / no-swizzle 16-byte channel planes [plane][pixel]; 8 gather warps; 32-channel chunks for
// stride 2) is kept for every mode as the fallback when no tensor map can be built and as the A/B reference
// (DYF_UMMA_A=cpasync: bit-exact for stride 1, same sums in another order for stride 2).
//
// Persistent, warp-specialised CTA (one per SM): warps 0-7 run the epilogue (tcgen05.ld.x32 -> per-item tables from
// shared memory -> fused affine / activation / dropout / residual as straight-line 32-column blocks -> 128-bit stores),
// then the patch producer (one TMA-issuing thread, or 8 cp.async warps), the MMA warp (owns TMEM, one elected thread
// issues) and the weight-stream warp.  Two TMEM accumulator sets ping-pong, so the epilogue of item i overlaps the MMAs
// of item i+1 and the A/B rings (2-6 patch stages, 4-9 weight stages, all on mbarriers) keep streaming across items.
// Work items go round-robin over the CTAs (concurrent CTAs on neighbouring tiles), in runs of 8 for single-chunk layers.
#include <cuda.h>

#include <cstdlib>
#include <map>
#include <tuple>

#include "conv.cuh"
#include "umma.cuh"

namespace dyf {
namespace {

constexpr int TILE_H = 16, TILE_W = 8;  // output pixels per CTA tile (M = 128)
constexpr int EPI_WARPS = 8;    // two warps per TMEM lane quarter, each draining half of the accumulator columns
// epilogue + patch-producer warps (1 issuing thread with TMA, 8 gathering warps with cp.async) + MMA warp + weight warp
// (the 1x1 gather kernel -- the NS readout over a column subset -- runs 4 gather warps: 448 threads leave its epilogue 144
// registers per thread; with 8 gather warps the 96-register cap spilled the epilogue's column block to local memory)
__host__ __device__ constexpr int prod_warps(bool tma, int mode) { return tma ? 1 : mode == 2 /* S1K1 */ ? 4 : 8; }
__host__ __device__ constexpr int cta_threads(bool tma, int mode) { return (EPI_WARPS + prod_warps(tma, mode) + 2) * 32; }

enum Mode { S1K3 = 0, S2K4 = 1, S1K1 = 2, S2K2 = 3, S1K7V = 4 };  // S2K2 (2x2 / stride 2 / pad 0) runs on the S1K1 kernel, see launch_s2k2
// S1K7V: 7 VERTICAL taps (7x1 filter, pad 3 vertically, none horizontally) over 64 channels -- the SST stem (7x7, C_in <= 8)
// after its horizontal taps were packed into the channel axis (pack_xim2col_kernel): TMA path only

template <int MODE, int T = 1> struct Geo;  // T = pixel tiles side by side per work item (gathered patches: stride-2 mode only)
template <int T> struct Geo<S1K3, T> {
  static constexpr int CH = 64;                     // channels per chunk
  static constexpr int PLANES = 8;                  // 16-byte planes per chunk
  static constexpr int CPL = 8;                     // 8-channel groups per pixel in a chunk
  static constexpr int PH = TILE_H + 2, PW = TILE_W + 2;
  static constexpr int PIX = PH * PW;               // gathered pixels
  static constexpr int SLOTS = PIX;                 // pixel slots per plane
  static constexpr int TAPS = 9, KW = 3, HALO = 2;
  static constexpr int GT = 3;                      // taps per weight stage (one filter row)
  static constexpr int SBO = PW * 16;
  __device__ static int tap_offset(int ky, int kx, int /*plane_bytes*/) { return (ky * PW + kx) * 16; }
  __device__ static int slot(int pr, int pc) { return pr * PW + pc; }
  __device__ static int plane_of(int g, int /*pc*/) { return g; }
  __device__ static int in_y(int oy0, int pr) { return oy0 - 1 + pr; }
  __device__ static int in_x(int ox0, int pc) { return ox0 - 1 + pc; }
};
template <int T> struct Geo<S1K1, T> {  // 1x1 convolution = plain GEMM over the 16 x 8T pixel tile (no halo, one "tap")
  static constexpr int CH = 64, PLANES = 8, CPL = 8;
  static constexpr int PH = TILE_H, PW = TILE_W * T;
  static constexpr int PIX = PH * PW, SLOTS = PIX;
  static constexpr int TAPS = 1, KW = 1, GT = 1, HALO = 0;
  static constexpr int SBO = PW * 16;
  __device__ static int tap_offset(int, int, int) { return 0; }
  __device__ static int slot(int pr, int pc) { return pr * PW + pc; }
  __device__ static int plane_of(int g, int) { return g; }
  __device__ static int in_y(int oy0, int pr) { return oy0 + pr; }
  __device__ static int in_x(int ox0, int pc) { return ox0 + pc; }
};
template <int T> struct Geo<S1K7V, T> {  // (the gather-path members are placeholders: this mode only runs with TMA patches)
  static constexpr int CH = 64, PLANES = 8, CPL = 8;
  static constexpr int PH = TILE_H + 6, PW = TILE_W;
  static constexpr int PIX = PH * PW, SLOTS = PIX;
  static constexpr int TAPS = 7, KW = 1, GT = 1, HALO = 0;
  static constexpr int SBO = PW * 16;
  __device__ static int tap_offset(int ky, int, int) { return ky * PW * 16; }
  __device__ static int slot(int pr, int pc) { return pr * PW + pc; }
  __device__ static int plane_of(int g, int) { return g; }
  __device__ static int in_y(int oy0, int pr) { return oy0 - 3 + pr; }
  __device__ static int in_x(int ox0, int pc) { return ox0 + pc; }
};
template <int T> struct Geo<S2K4, T> {
  static constexpr int CH = 32;
  static constexpr int PLANES = 8;                  // 2 column parities x 4 channel groups
  static constexpr int CPL = 4;
  static constexpr int PH = 2 * TILE_H + 2, PW = 2 * TILE_W * T + 2, PWH = PW / 2;
  static constexpr int PIX = PH * PW;
  static constexpr int SLOTS = PH * PWH;
  static constexpr int TAPS = 16, KW = 4, HALO = 2;
  static constexpr int GT = 4;
  static constexpr int SBO = 2 * PWH * 16;          // output row r reads patch row 2r + ky
  __device__ static int tap_offset(int ky, int kx, int plane_bytes) {
    return (kx & 1) * CPL * plane_bytes + (ky * PWH + (kx >> 1)) * 16;
  }
  __device__ static int slot(int pr, int pc) { return pr * PWH + (pc >> 1); }
  __device__ static int plane_of(int g, int pc) { return (pc & 1) * CPL + g; }
  __device__ static int in_y(int oy0, int pr) { return 2 * oy0 - 1 + pr; }
  __device__ static int in_x(int ox0, int pc) { return 2 * ox0 - 1 + pc; }
};
template <int MODE, int T = 1> struct Sizes {
  using G = Geo<MODE, T>;
  static constexpr int PLANE = (G::SLOTS | 1) * 16;  // odd slot pitch => conflict-free 16-byte fills across planes
  static constexpr int A_STAGE = G::PLANES * PLANE;
  static constexpr int KSTEPS = G::CH / 16;
};

struct WorkDiv { FastDiv n_tiles, per_row, tiles_x, tab, grp; };  // divisors of the work-item decoding, table row, RNG row group

template <int AS, int BS>
struct __align__(8) Barriers {
  uint64_t a_full[AS], a_empty[AS], b_full[BS], b_empty[BS], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// Persistent kernel: one CTA per SM walks the work list (n_tile fastest, so CTAs sharing an activation tile run
// together and hit it in L2).  All pipelines run across tile boundaries: the A/B rings keep streaming into the next
// tile while the epilogue warps drain the previous accumulator (two TMEM accumulators, ping-pong).
// TMA = true (S1K3 only): the halo patch of a chunk is ONE 4-D tensor-map copy (cp.async.bulk.tensor, box
// 64 ch x 10 x 18 px, hardware zero fill outside the image) landing pixel-major with the 128-byte swizzle; taps are
// still pure start-address shifts (+128 B per pixel) of a SWIZZLE_128B K-major descriptor with SBO = one patch row.
// T (TMA mode only): pixel tiles per work item, side by side along x.  All T tiles consume every weight stage before it
// is released, so a weight byte pulled from L2 feeds T x 128 GEMM rows; the patch of the 16 x 8T super-tile is one TMA
// box and each tile's taps are start-address shifts inside it.  GT = filter taps per weight stage.
template <int BN, int MODE, int AS, int BS, bool TMA, int T, int GT>
__global__ void __launch_bounds__(cta_threads(TMA, MODE), 1) conv_umma_kernel(const ConvParams p, const act_t* __restrict__ wblob,
                                                              int tiles_x, int tiles_y, int n_tiles, int num_work,
                                                              const __grid_constant__ CUtensorMap tmap,
                                                              const __grid_constant__ CUtensorMap tmap2,
                                                              const __grid_constant__ CUtensorMap tmap3,
                                                              const __grid_constant__ CUtensorMap tmap4, int nch_split,
                                                              const WorkDiv wd) {
  using G = Geo<MODE, TMA ? 1 : T>;
  using S = Sizes<MODE, TMA ? 1 : T>;
  constexpr int PROD_WARPS = prod_warps(TMA, MODE), PROD_THREADS = PROD_WARPS * 32, MMA_WARP = EPI_WARPS + PROD_WARPS;
  constexpr bool S2TMA = TMA && MODE == S2K4;
  constexpr int CHK = S2TMA ? 64 : G::CH;   // channels per chunk (the stride-2 TMA path uses 64: 128-byte pixel rows)
  constexpr int PLANE = S::PLANE, KSTEPS = CHK / 16;
  static_assert(TMA || T == 1 || MODE == S2K4 || MODE == S1K1, "multi-tile gathered patches: stride-2 and 1x1 modes only");
  static_assert(G::TAPS % GT == 0, "weight stages must tile the filter");
  constexpr int PWT = TILE_W * T + G::HALO;  // patch width of the super-tile (TMA mode)
  constexpr int PIXT = G::PH * PWT;
  // Stride-2 TMA mode: the patch is four boxes, one per (row parity, column parity) VIEW of the input (tensor maps with
  // pixel pitch 2 and row pitch 2: quarter-resolution images; zero fill outside = the conv's padding).  Inside a view
  // the stride-2 walk of a tap is a unit-stride one, so every tap is a start-address shift inside one view: tap (ky, kx)
  // reads view ((ky+1)&1, (kx+1)&1) at box offset (ky>>1, kx>>1).  The TMA engine sustains about one box segment
  // (one pixel of one view) per 6 cycles per SM: 32-channel chunks (64-byte segments) were measured segment-rate bound
  // at 60 % of the tensor pipe, hence 64-channel chunks = 128-byte pixel rows with the 128-byte swizzle -- and, to keep
  // the footprint at one chunk, a VIEW-granular ring: one stage = one parity view of one chunk (4 taps), the producer
  // runs three views ahead of the MMAs.
  constexpr int VW = TILE_W * T + 1, VH = TILE_H + 1;                      // view box: 17 rows x (8T+1) pixels
  constexpr int VREGION = ((VH * VW * 128 + 1023) / 1024) * 1024;
  constexpr int A_STAGE = S2TMA ? VREGION : TMA ? ((PIXT * 128 + 1023) / 1024) * 1024 : S::A_STAGE;
  constexpr int B_TAP = BN * CHK * 2;      // one tap of one chunk: [k8][BN rows][16 B]
  constexpr int B_STAGE = GT * B_TAP;      // a weight stage carries GT taps
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + AS * A_STAGE;
  using Bars = Barriers<AS, BS>;
  Bars* bars = reinterpret_cast<Bars*>(sB + BS * B_STAGE);
  float* sTab = reinterpret_cast<float*>(sB + BS * B_STAGE + ((sizeof(Bars) + 15) & ~15));  // [2 acc][A | B][BN] epilogue tables

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = p.Cin / CHK;
  const int tiles_per_row = tiles_x * tiles_y;
  // whole filter fits the weight ring and every work item uses the same n-tile: load it once, never release it
  const bool b_resident = n_tiles == 1 && nchunks * (G::TAPS / GT) == BS;
  // Work items go round-robin over the CTAs in runs of `cw` consecutive items.  Multi-chunk layers use cw = 1 (concurrent
  // CTAs on neighbouring tiles: halos hit in L2, accesses spread over the DRAM partitions; contiguous per-CTA slices were
  // measured 10-50 % slower); short single-chunk layers (stem, readout, 1x1) are bound by per-item latencies, and runs of 8
  // tiles of one image let them keep their epilogue tables.
  const int cw = (nchunks * G::TAPS <= 9 && num_work >= 16 * (int)gridDim.x) ? 8 : 1;  // (only with plenty of items per CTA)

  if (tid == 0) {
    for (int i = 0; i < AS; ++i) { mbar_init(smem_u32(&bars->a_full[i]), TMA ? 1 : PROD_THREADS); mbar_init(smem_u32(&bars->a_empty[i]), 1); }
    for (int i = 0; i < BS; ++i) { mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {  // TMEM: two accumulators of BN fp32 columns, owned by the MMA warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(2 * T * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();  // the prologue above overlapped the previous kernel's tail; everything below touches its output

  // work item -> (n_tile, batch row, tile origin)
  auto decode = [&](int w, int& n_tile, int& row, int& oy0, int& ox0) {
    int t = (int)wd.n_tiles.div((uint32_t)w);
    n_tile = w - t * n_tiles;
    row = (int)wd.per_row.div((uint32_t)t);
    t -= row * tiles_per_row;
    const int ty = (int)wd.tiles_x.div((uint32_t)t);
    oy0 = ty * TILE_H;
    ox0 = (t - ty * tiles_x) * (TILE_W * T);
  };

  if (warp < EPI_WARPS) {
    // =============================== epilogue warps: TMEM -> registers -> fused math -> global ====================
    const int act = p.act;
    const float slope = act == ACT_NONE ? 1.f : act == ACT_RELU ? 0.f : 0.2f;
    const uint32_t thresh = p.drop.thresh;
    const float dscale = p.drop.scale;
    const bool wide_st = ((p.out_ld | p.out_coff) & 15) == 0;  // output rows and slices 32-byte aligned: 256-bit stores
    const bool wide_res = p.res && (p.res_ld & 15) == 0;
    int it = 0, tab_key = -1, tab_buf = 0;
    for (int w0 = blockIdx.x * cw; w0 < num_work; w0 += gridDim.x * cw)
    for (int w = w0; w < min(w0 + cw, num_work); ++w, ++it) {
      int n_tile, row, oy0, ox0;
      decode(w, n_tile, row, oy0, ox0);
      const int acc = it & 1;
      const uint32_t gj = wd.grp.div((uint32_t)row), gr = (uint32_t)row - gj * p.drop.group_rows;  // = drop_row()
      const DropRow dr{p.drop.stream_lo + gj, (uint64_t)(gr + p.drop.row_off) * ((uint64_t)p.Ho * p.Wo * p.Cout)};
      const int quarter = warp & 3;          // TMEM lanes 32*quarter.. are the ones this warp may read
      const int m_local = quarter * 32 + lane;  // accumulator row = TMEM lane
      const int oy = oy0 + (m_local >> 3);
      // epilogue tables of (table row, n-tile) staged in shared memory; re-staged only when the key changes (contiguous
      // work ranges: once per image), into the other buffer so that one barrier per change suffices
      const int trow = (int)wd.tab.div((uint32_t)row);  // table row (rows of one logical call share it)
      const int tkey = trow * n_tiles + n_tile;
      if (tkey != tab_key) {
        tab_key = tkey;
        tab_buf ^= 1;
        float* const dst = sTab + tab_buf * 2 * BN;
        if (tid < BN) {
          const int col = n_tile * BN + tid;
          const size_t off = (size_t)trow * p.Cout + col;
          dst[tid] = col < p.Cout ? __ldg(p.tabA + off) : 0.f;
          dst[BN + tid] = col < p.Cout ? __ldg(p.tabB + off) : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      }
      const float* const tA = sTab + tab_buf * 2 * BN;
      const float* const tB = tA + BN;
      mbar_wait(smem_u32(&bars->acc_full[acc]), (it >> 1) & 1);
      tc_fence_after();
      constexpr int COLS = BN / (EPI_WARPS / 4);  // columns drained by this warp
      const int cbeg = (warp >> 2) * COLS;
#pragma unroll 1
      for (int tile = 0; tile < T; ++tile) {
      const int ox = ox0 + tile * TILE_W + (m_local & 7);
      const bool valid = oy < p.Ho && ox < p.Wo;
      const long long m = ((long long)row * p.Ho + oy) * p.Wo + ox;
      act_t* const orow = reinterpret_cast<act_t*>(p.out) + (size_t)m * p.out_ld + p.out_coff + n_tile * BN;
      const act_t* const rrow = p.res ? p.res + (size_t)m * p.res_ld + n_tile * BN : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc * T + tile) * BN;
#pragma unroll 1
      for (int cg = cbeg; cg < cbeg + COLS; cg += 32) {
        if (n_tile * BN + cg >= p.Cout) break;  // zero-padded output channels of a ragged last n-tile (warp-uniform)
        uint32_t v[32];
        tmem_ld32_nowait(taddr + cg, v);
        tmem_ld_wait();
        // 32 columns as one straight-line block (no per-8-column control flow: the scheduler interleaves the chains)
        float y[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 a = *reinterpret_cast<const float4*>(tA + cg + 4 * q), b = *reinterpret_cast<const float4*>(tB + cg + 4 * q);
          y[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), a.x, b.x);
          y[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), a.y, b.y);
          y[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), a.z, b.z);
          y[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), a.w, b.w);
        }
        if (act == ACT_NONE) {
        } else if (act <= ACT_LEAKY) {  // ReLU / LeakyReLU(0.2) = max(y, slope * y) with slope 0 / 0.2
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], slope * y[j]);
        } else if (act == ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = y[j] / (1.f + __expf(-y[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = gelu_fast(y[j]);
        }
        if (thresh) {
#pragma unroll
          for (int cs = 0; cs < 32; cs += 8) {
            const uint32_t keep = drop_keep_bits8(p.drop, dr, (uint64_t)(oy * p.Wo + ox) * p.Cout + n_tile * BN + cg + cs);
#pragma unroll
            for (int j = 0; j < 8; ++j) y[cs + j] = ((keep >> j) & 1u) ? y[cs + j] * dscale : 0.f;
          }
        }
        if (valid) {
          if (rrow) {
#pragma unroll
            for (int cs = 0; cs < 32; cs += 16) {
              const int c0 = n_tile * BN + cg + cs;
              uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
              if (wide_res && c0 + 16 <= p.Cout) {
                ld_global_nc_256(rrow + cg + cs, r0, r1);
              } else {
                if (c0 < p.Cout) r0 = __ldg(reinterpret_cast<const uint4*>(rrow + cg + cs));
                if (c0 + 8 < p.Cout) r1 = __ldg(reinterpret_cast<const uint4*>(rrow + cg + cs + 8));
              }
              float f[8];
              if (c0 < p.Cout) {
                unpack8(r0, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[cs + j] += f[j];
              }
              if (c0 + 8 < p.Cout) {
                unpack8(r1, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[cs + 8 + j] += f[j];
              }
            }
          }
#pragma unroll
          for (int cs = 0; cs < 32; cs += 16) {
            const int c0 = n_tile * BN + cg + cs;
            if (wide_st && c0 + 16 <= p.Cout) {
              st_global_256(orow + cg + cs, pack8(y + cs), pack8(y + cs + 8));
            } else {
              if (c0 < p.Cout) *reinterpret_cast<uint4*>(orow + cg + cs) = pack8(y + cs);
              if (c0 + 8 < p.Cout) *reinterpret_cast<uint4*>(orow + cg + cs + 8) = pack8(y + cs + 8);
            }
          }
        }
      }
      }
      tc_fence_before();                                    // all tcgen05.ld of this accumulator have completed
      mbar_arrive(smem_u32(&bars->acc_empty[acc]));         // hand the accumulator back to the MMA warp
    }
  } else if (warp < MMA_WARP) {
    // =============================== A producers: halo patches via cp.async (zero fill = padding) ================
    const int ptid = tid - EPI_WARPS * 32;
    if constexpr (TMA) {
      if (ptid == 0) {
        int ca = 0;
        for (int w0 = blockIdx.x * cw; w0 < num_work; w0 += gridDim.x * cw)
      for (int w = w0; w < min(w0 + cw, num_work); ++w) {
          int n_tile, row, oy0, ox0;
          decode(w, n_tile, row, oy0, ox0);
          if constexpr (S2TMA) {
            for (int c = 0; c < nchunks; ++c) {
#pragma unroll
              for (int v = 0; v < 4; ++v, ++ca) {  // v = row parity * 2 + column parity; odd views start one view pixel earlier
                const int st = ca % AS;
                mbar_wait(smem_u32(&bars->a_empty[st]), ((ca / AS) & 1) ^ 1);
                const uint32_t bar = smem_u32(&bars->a_full[st]);
                mbar_expect_tx(bar, VH * VW * 128);
                const uint64_t tmv = reinterpret_cast<uint64_t>(v == 0 ? &tmap : v == 1 ? &tmap2 : v == 2 ? &tmap3 : &tmap4);
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                    ::"r"(smem_u32(sA + st * A_STAGE)), "l"(tmv), "r"(c * CHK), "r"(ox0 - (v & 1)),
                      "r"(oy0 - (v >> 1)), "r"(row), "r"(bar) : "memory");
              }
            }
            continue;
          }
          for (int c = 0; c < nchunks; ++c, ++ca) {
            const int st = ca % AS;
            mbar_wait(smem_u32(&bars->a_empty[st]), ((ca / AS) & 1) ^ 1);
            const uint32_t bar = smem_u32(&bars->a_full[st]);
            mbar_expect_tx(bar, PIXT * 128);
            // nch_split > 0: the K axis is the concatenation of two tensor views (chunks [0, nch_split) / the rest)
            const bool second = nch_split > 0 && c >= nch_split;
            const uint64_t tm = reinterpret_cast<uint64_t>(second ? &tmap2 : &tmap);
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                ::"r"(smem_u32(sA + st * A_STAGE)), "l"(tm), "r"((second ? c - nch_split : c) * G::CH),
                  "r"(G::in_x(ox0, 0)), "r"(G::in_y(oy0, 0)), "r"(row), "r"(bar) : "memory");
          }
        }
      }
    } else {
    constexpr int PSTEP = PROD_THREADS / G::CPL;            // patch pixels covered per pass of the producer threads
    constexpr int ITERS = (G::PIX + PSTEP - 1) / PSTEP;
    const int g8 = ptid % G::CPL;                           // fixed 8-channel group of this thread
    const act_t* const in_base = p.in;
    int ca = 0;                                             // running chunk counter (A ring position)
    for (int w0 = blockIdx.x * cw; w0 < num_work; w0 += gridDim.x * cw)
      for (int w = w0; w < min(w0 + cw, num_work); ++w) {
      int n_tile, row, oy0, ox0;
      decode(w, n_tile, row, oy0, ox0);
      int src_off[ITERS];                                   // element offset of the patch pixel, -1 = outside the image
      int dst_off[ITERS];                                   // byte offset inside a stage, -1 = no such pixel
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int pix = ptid / G::CPL + it * PSTEP;
        int so = -1, dof = -1;
        if (pix < G::PIX) {
          const int pr = pix / G::PW, pc = pix - pr * G::PW;
          const int iy = G::in_y(oy0, pr);
          int ix = G::in_x(ox0, pc);
          bool ok = (unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)(p.in_xmap ? p.Wo : p.Wi);
          if (ok && p.in_xmap) ix = __ldg(p.in_xmap + ix);  // column subset: virtual output column -> input column
          if (ok) so = (iy * p.Wi + ix) * p.Cin;
          dof = G::plane_of(g8, pc) * PLANE + G::slot(pr, pc) * 16;
        }
        src_off[it] = so;
        dst_off[it] = dof;
      }
      const act_t* const in_row = in_base + (size_t)row * p.Hi * p.Wi * p.Cin + g8 * 8;
      for (int c = 0; c < nchunks; ++c, ++ca) {
        const int st = ca % AS;
        mbar_wait(smem_u32(&bars->a_empty[st]), ((ca / AS) & 1) ^ 1);
        const uint32_t dst0 = smem_u32(sA + st * A_STAGE);
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          if (dst_off[it] >= 0) {
            const bool v = src_off[it] >= 0;
            cp_async16(dst0 + dst_off[it], v ? (const void*)(in_row + src_off[it] + c * G::CH) : (const void*)in_base,
                       v ? 16 : 0);
          }
        }
        // hardware arrives on a_full[st] once this thread's copies have landed (no wait, no proxy fence: the same
        // pattern as CUTLASS's sm100 cp.async mainloop); the producer immediately moves on to the next chunk
        cp_async_arrive_noinc(smem_u32(&bars->a_full[st]));
      }
    }
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer (one elected thread) ==============================================
    // The issue loop is latency-critical (one thread feeds the whole tensor pipe): descriptors are advanced by adding
    // compile-time constants to pre-built low words, taps are fully unrolled, and barriers are touched once per
    // filter row (GT taps x KSTEPS MMAs) rather than once per tap.
    // The whole warp runs the (warp-uniform) loop; only the elected lane's tcgen05 instructions are predicated on.
    {
      const uint32_t leader = elect_one();
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (DYF_UMMA_FMT << 7) | (DYF_UMMA_FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      // A descriptor high word: SBO | version 1 (bit 46) [| SWIZZLE_128B (bits 61-63) with SBO = one 128-B-pixel patch row]
      const uint32_t a_hi = S2TMA ? ((uint32_t)(((VW * 128) >> 4) & 0x3FFF) | (1u << 14) | (2u << 29))  // SWIZZLE_128B
                            : TMA ? ((uint32_t)(((PWT * 128) >> 4) & 0x3FFF) | (1u << 14) | (2u << 29))
                                  : ((uint32_t)((G::SBO >> 4) & 0x3FFF) | (1u << 14));
      const uint32_t b_hi = (uint32_t)((128 >> 4) & 0x3FFF) | (1u << 14);
      const uint32_t a_lo0 = ((uint32_t)(TMA ? 1 : (PLANE >> 4)) << 16) | (smem_u32(sA) >> 4);  // LBO | start address
      const uint32_t b_lo0 = ((uint32_t)((BN * 16) >> 4) << 16) | (smem_u32(sB) >> 4);
      const uint32_t bar_a_full = smem_u32(&bars->a_full[0]), bar_a_empty = smem_u32(&bars->a_empty[0]);
      const uint32_t bar_b_full = smem_u32(&bars->b_full[0]), bar_b_empty = smem_u32(&bars->b_empty[0]);
      constexpr int AK = (TMA ? 32 : 2 * PLANE) >> 4;  // descriptor step between the K = 16 slices of a chunk
      constexpr int BK = (2 * BN * 16) >> 4;
      int sa = 0, pa = 0, sb = 0, pb = 0, it = 0;  // ring positions / phase parities
      for (int w0 = blockIdx.x * cw; w0 < num_work; w0 += gridDim.x * cw)
    for (int w = w0; w < min(w0 + cw, num_work); ++w, ++it) {
        const int acc = it & 1;
        mbar_wait(smem_u32(&bars->acc_empty[acc]), ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * T * BN;
        if constexpr (S2TMA) {
          // view-major tap order: view v = (py, px) holds taps ky = 1 - py + 2a, kx = 1 - px + 2b (a, b in {0, 1}) at box
          // offset (a, b); the weight producer streams the taps in the same order
          static_assert(!S2TMA || GT == 1, "stride-2 TMA path: one tap per weight stage");
          for (int c = 0; c < nchunks; ++c) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              mbar_wait(bar_a_full + sa * 8, pa);
              tc_fence_after();
              const uint64_t a_st = ((uint64_t)a_hi << 32) | (a_lo0 + sa * (A_STAGE >> 4));
#pragma unroll
              for (int ab = 0; ab < 4; ++ab) {
                mbar_wait(bar_b_full + sb * 8, pb);
                tc_fence_after();
                const uint64_t b_st = ((uint64_t)b_hi << 32) | (b_lo0 + sb * (B_STAGE >> 4));
                const int a_off = ((ab >> 1) * VW + (ab & 1)) * 128;
#pragma unroll
                for (int tile = 0; tile < T; ++tile)
                  umma_tap<KSTEPS, AK, BK>(tmem_acc + tile * BN, a_st + (uint64_t)((a_off + tile * TILE_W * 128) >> 4), b_st, idesc,
                                           (v | ab) ? 1u : (uint32_t)(c != 0), leader);
                umma_commit_if(bar_b_empty + sb * 8, leader);
                if (++sb == BS) { sb = 0; pb ^= 1; }
              }
              umma_commit_if(bar_a_empty + sa * 8, leader);
              if (++sa == AS) { sa = 0; pa ^= 1; }
            }
          }
        } else
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(bar_a_full + sa * 8, pa);
          tc_fence_after();
          const uint64_t a_st = ((uint64_t)a_hi << 32) | (a_lo0 + sa * (A_STAGE >> 4));
#pragma unroll
          for (int g = 0; g < G::TAPS / GT; ++g) {
            if (!b_resident || it == 0) {
              mbar_wait(bar_b_full + sb * 8, pb);
              tc_fence_after();
            }
            const uint64_t b_st = ((uint64_t)b_hi << 32) | (b_lo0 + sb * (B_STAGE >> 4));
#pragma unroll
            for (int t = 0; t < GT; ++t) {
              const int tap = g * GT + t, ky = tap / G::KW, kx = tap - ky * G::KW;
              const int a_off = TMA ? (ky * PWT + kx) * 128 : G::tap_offset(ky, kx, PLANE);
              constexpr int TILE_STEP = TILE_W * (TMA ? 128 : 16);  // 8 pixels further along the patch row
#pragma unroll
              for (int tile = 0; tile < T; ++tile)
                umma_tap<KSTEPS, AK, BK>(tmem_acc + tile * BN, a_st + (uint64_t)((a_off + tile * TILE_STEP) >> 4),
                                         b_st + (uint64_t)((t * B_TAP) >> 4), idesc, tap ? 1u : (uint32_t)(c != 0), leader);
            }
            if (!b_resident) umma_commit_if(bar_b_empty + sb * 8, leader);  // weight stage free once these MMAs retire
            if (++sb == BS) { sb = 0; pb ^= 1; }
          }
          umma_commit_if(bar_a_empty + sa * 8, leader);    // patch stage free
          if (++sa == AS) { sa = 0; pa ^= 1; }
        }
        umma_commit_if(smem_u32(&bars->acc_full[acc]), leader);  // accumulator complete -> epilogue
      }
    }
  } else {
    // =============================== B producer: bulk-TMA weight tiles =============================================
    if (lane == 0) {
      const int per_tile = nchunks * (G::TAPS / GT);
      int sb = 0, pb = 1;
      bool loaded = false;  // resident filter: loaded by the first work item only
      for (int w0 = blockIdx.x * cw; w0 < num_work && !(b_resident && loaded); w0 += gridDim.x * cw)
      for (int w = w0; w < min(w0 + cw, num_work) && !(b_resident && loaded); ++w) {
        loaded = true;
        const int n_tile = w - (int)wd.n_tiles.div((uint32_t)w) * n_tiles;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(wblob) + (size_t)n_tile * per_tile * B_STAGE;
        for (int i = 0; i < per_tile; ++i) {
          int j = i;
          if constexpr (S2TMA) {  // view-major tap order of the MMA loop: (chunk, view (py, px), a, b) -> tap ky*4 + kx
            const int c = i >> 4, v = (i >> 2) & 3, ab = i & 3;
            j = c * 16 + (1 - (v >> 1) + 2 * (ab >> 1)) * 4 + (1 - (v & 1) + 2 * (ab & 1));
          }
          mbar_wait(smem_u32(&bars->b_empty[sb]), pb);
          mbar_expect_tx(smem_u32(&bars->b_full[sb]), B_STAGE);
          bulk_g2s(smem_u32(sB + sb * B_STAGE), src + (size_t)j * B_STAGE, B_STAGE, smem_u32(&bars->b_full[sb]));
          if (++sb == BS) { sb = 0; pb ^= 1; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * T * BN));
  }
}

// weights fp32 [O, I, KH, KW] -> bf16 stage tiles [n_tile][chunk][tap][k8][n (BN)][8]: one contiguous blob per MMA stage
__global__ void __launch_bounds__(256) repack_umma_kernel(const float* __restrict__ w, act_t* __restrict__ out,
                                                         int O, int I, int BN, int taps, int ch, int standardize) {
  const int nchunks = I / ch, k8n = ch / 8;
  const long long total = (long long)((O + BN - 1) / BN * BN) * I * taps;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long r = idx;
  const int e = (int)(r % 8); r /= 8;
  const int n = (int)(r % BN); r /= BN;
  const int k8 = (int)(r % k8n); r /= k8n;
  const int tap = (int)(r % taps); r /= taps;
  const int chunk = (int)(r % nchunks); r /= nchunks;
  const int n_tile = (int)r;
  const int o = n_tile * BN + n, ci = chunk * ch + k8 * 8 + e;
  if (o >= O) { out[idx] = f2act(0.f); return; }  // zero rows of a ragged last n-tile
  float v = w[((size_t)o * I + ci) * taps + tap];
  if (standardize) {  // WeightStandardizedConv2d (reference unet.py:32-40), recomputed per element (load-time only)
    const float* wo = w + (size_t)o * I * taps;
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < I * taps; ++i) s1 += wo[i];
    const float mean = s1 / (I * taps);
    for (int i = 0; i < I * taps; ++i) { const float d = wo[i] - mean; s2 += d * d; }
    v = (v - mean) * rsqrtf(s2 / (I * taps) + 1e-5f);
  }
  out[idx] = f2act(v);
}

// Pipeline shapes that fill the 227 KB of one SM.  T = pixel tiles per work item (TMA mode), GT = taps per weight stage,
// A / B = patch / weight ring depths.
// V = shape variant: 0 = default; 1 = single-chunk layers (Cin == one chunk, one n-tile, e.g. the NS stem): the whole
// filter stays resident in the weight ring and the freed L2 bandwidth + a deeper patch ring feed the short tiles; also the
// single-tile shape of 1x1 convs on grids <= 8 pixels wide.
template <int MODE, int BN, bool TMA, int V = 0> struct Stages;
template <> struct Stages<S1K3, 64, true> { static constexpr int T = 2, GT = 1, A = 3, B = 9; };    // 123 KB patches +  72 KB weights
template <> struct Stages<S1K3, 128, true> { static constexpr int T = 2, GT = 1, A = 2, B = 8; };   //  82 KB patches + 128 KB weights
template <> struct Stages<S1K3, 64, true, 1> { static constexpr int T = 2, GT = 1, A = 3, B = 9; };
template <> struct Stages<S1K3, 128, true, 1> { static constexpr int T = 1, GT = 1, A = 3, B = 9; };  //  69 KB patches + 144 KB resident filter
template <> struct Stages<S1K3, 64, false> { static constexpr int T = 1, GT = 3, A = 6, B = 3; };
template <> struct Stages<S1K3, 128, false> { static constexpr int T = 1, GT = 3, A = 3, B = 3; };
template <> struct Stages<S1K7V, 64, true> { static constexpr int T = 2, GT = 1, A = 3, B = 7; };    // 135 KB patches + 56 KB resident filter
template <> struct Stages<S1K7V, 128, true> { static constexpr int T = 2, GT = 1, A = 2, B = 7; };   //  90 KB patches + 112 KB resident filter
template <> struct Stages<S1K1, 64, true> { static constexpr int T = 2, GT = 1, A = 5, B = 4; };
template <> struct Stages<S1K1, 128, true> { static constexpr int T = 2, GT = 1, A = 4, B = 4; };
template <> struct Stages<S1K1, 64, true, 1> { static constexpr int T = 1, GT = 1, A = 6, B = 6; };   // grids <= 8 pixels wide
template <> struct Stages<S1K1, 128, true, 1> { static constexpr int T = 1, GT = 1, A = 6, B = 6; };
template <> struct Stages<S1K1, 64, false> { static constexpr int T = 1, GT = 1, A = 6, B = 4; };
template <> struct Stages<S1K1, 128, false> { static constexpr int T = 1, GT = 1, A = 6, B = 4; };
template <> struct Stages<S1K1, 64, false, 2> { static constexpr int T = 1, GT = 1, A = 10, B = 1; };  // column-subset readout: resident filter, deep gather ring
template <> struct Stages<S2K4, 64, false> { static constexpr int T = 2, GT = 2, A = 2, B = 8; };   // 145 KB patches +  64 KB weights
template <> struct Stages<S2K4, 128, false> { static constexpr int T = 2, GT = 1, A = 2, B = 9; };  // 145 KB patches +  72 KB weights
template <> struct Stages<S2K4, 64, true> { static constexpr int T = 2, GT = 1, A = 4, B = 8; };    // 148 KB (4 view stages) + 64 KB weights
template <> struct Stages<S2K4, 128, true> { static constexpr int T = 2, GT = 1, A = 4, B = 4; };   // 148 KB (4 view stages) + 64 KB weights

// Patch tensor map of a layer input, cached per (buffer, geometry): the workspace carving is stable across forwards.
static int make_patch_tmap(const ConvParams& p, const act_t* src, int C, int pw, int ph, CUtensorMap* out) {
  using Key = std::tuple<const void*, int, int, int, int, int, int>;
  static std::map<Key, CUtensorMap> cache;
  const Key key{src, p.rows, p.Hi, p.Wi, C, pw, ph};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  CUtensorMap m;
  if (make_nhwc_tmap(src, p.rows, p.Hi, p.Wi, C, C, pw, ph, 1, &m) != 0) return -1;
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return 0;
}

// views: optional pair of pre-built tensor maps splitting the K axis (2x2 / stride-2 layers); nch_split = chunks of the first
template <int BN, int MODE, bool TMA, int V = 0>
int launch_t(const ConvParams& p, cudaStream_t stream, const CUtensorMap* views = nullptr, int nch_split = 0) {
  using St = Stages<MODE, BN, TMA, V>;
  constexpr int AS = St::A, BS = St::B, T = St::T, GT = St::GT;
  using G = Geo<MODE, TMA ? 1 : T>;
  constexpr int CHK = (TMA && MODE == S2K4) ? 64 : G::CH;
  constexpr int a_stage = (TMA && MODE == S2K4) ? ((((TILE_H + 1) * (TILE_W * T + 1) * 128 + 1023) / 1024) * 1024)
                          : TMA ? (((TILE_W * T + G::HALO) * G::PH * 128 + 1023) / 1024) * 1024 : Sizes<MODE, T>::A_STAGE;
  constexpr int smem = AS * a_stage + BS * GT * BN * CHK * 2 + (((int)sizeof(Barriers<AS, BS>) + 15) & ~15) + 4 * BN * 4 + 64;
  static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
  static_assert(2 * T * BN <= 512, "TMEM budget exceeded");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    DYF_CUDA_OK(cudaGetDevice(&dev));
    DYF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    DYF_CUDA_OK(cudaFuncSetAttribute(conv_umma_kernel<BN, MODE, AS, BS, TMA, T, GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  CUtensorMap tmap{}, tmap2{}, tmap3{}, tmap4{};
  if (views) { tmap = views[0]; tmap2 = views[1]; if (MODE == S2K4) { tmap3 = views[2]; tmap4 = views[3]; } }
  else if (TMA && p.in2) {  // channel concat of two sources: chunks [0, Cin0/64) from the first map, the rest from the second
    if (make_patch_tmap(p, p.in, p.Cin0, TILE_W * T + G::HALO, G::PH, &tmap) != 0 ||
        make_patch_tmap(p, p.in2, p.Cin - p.Cin0, TILE_W * T + G::HALO, G::PH, &tmap2) != 0)
      return 0;
    nch_split = p.Cin0 / G::CH;
  } else if (TMA && make_patch_tmap(p, p.in, p.Cin, TILE_W * T + G::HALO, G::PH, &tmap) != 0) return 0;  // caller falls back
  const int tiles_x = (p.Wo + TILE_W * T - 1) / (TILE_W * T), tiles_y = (p.Ho + TILE_H - 1) / TILE_H;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const long long work = (long long)tiles_x * tiles_y * p.rows * n_tiles;
  if (work > 0x7fffffffLL) { set_error("conv_umma: too many tiles"); return -1; }
  const int grid = (int)(work < num_sms ? work : num_sms);
  const WorkDiv wd{FastDiv((uint32_t)n_tiles), FastDiv((uint32_t)(tiles_x * tiles_y)), FastDiv((uint32_t)tiles_x),
                   FastDiv((uint32_t)(p.tab_div > 0 ? p.tab_div : 1)), FastDiv(p.drop.group_rows ? p.drop.group_rows : 1u)};
  const double flops = 2.0 * (double)p.M * p.Cout * G::TAPS * p.Cin_real;
  const double bytes = 2.0 * ((double)p.rows * p.Hi * (p.in_xmap ? p.Wo : p.Wi) * p.Cin + (double)p.M * p.Cout + (double)p.Cout * p.Kpad);
  ProfScope prof(stream, KC_CONV_UMMA, flops, bytes);
  DYF_LAUNCH_PDL(0, "conv_umma_kernel", (conv_umma_kernel<BN, MODE, AS, BS, TMA, T, GT>), dim3(grid), dim3(cta_threads(TMA, MODE)), smem,
                 stream, p, p.w_umma, tiles_x, tiles_y, n_tiles, (int)work, tmap, tmap2, tmap3, tmap4, nch_split, wd);
  return 1;
}

// channels per chunk of the stride-2 weight tiles: 64 for the TMA path, 32 for the cp.async gather (DYF_UMMA_A=cpasync);
// process-wide, so that finalize (re-packing) and launch agree
int s2k4_weights_chunk() {
  static const char* env_a = getenv("DYF_UMMA_A");
  return (env_a && env_a[0] == 'c') ? 32 : 64;
}

int mode_of(int k, int stride, int pad, int kw = -1) {
  if (k == 7 && kw == 1 && stride == 1 && pad == 3) return S1K7V;
  if (kw >= 0 && kw != k) return -1;
  if (k == 3 && stride == 1 && pad == 1) return S1K3;
  if (k == 4 && stride == 2 && pad == 1) return S2K4;
  if (k == 1 && stride == 1 && pad == 0) return S1K1;
  if (k == 2 && stride == 2 && pad == 0) return S2K2;
  return -1;
}

}  // namespace

int umma_tile_n(int Cout) { return Cout <= 64 ? 64 : 128; }
int umma_padded_cout(int Cout) { const int bn = umma_tile_n(Cout); return (Cout + bn - 1) / bn * bn; }

bool conv_umma_shape_ok(int Cin_pad, int Cout, int k, int stride, int pad) {
  const int mode = mode_of(k, stride, pad);
  if (mode < 0 || Cout % 8 != 0) return false;
  if (mode == S1K1) return Cin_pad % 64 == 0 && (Cout <= 64 || Cout % 128 == 0);
  if (mode == S2K2) return Cin_pad % 32 == 0 && (Cout == 64 || Cout % 128 == 0);
  if (!(Cout == 64 || Cout % 128 == 0)) return false;
  return Cin_pad % (mode == S1K3 ? 64 : 32) == 0;
}

bool conv_umma_eligible(const ConvParams& p) {
  if (p.in2) {  // two-source input: stride-1 TMA modes only, both parts whole chunks
    const int mode = mode_of(p.KH, p.stride, p.pad);
    if ((mode != S1K3 && mode != S1K1) || p.Cin0 <= 0 || p.Cin0 % 64 || (p.Cin - p.Cin0) % 64) return false;
  }
  if (p.in_xmap && (p.in2 || mode_of(p.KH, p.stride, p.pad, p.KW) != S1K1)) return false;
  if (mode_of(p.KH, p.stride, p.pad, p.KW) == S1K7V)
    return p.w_umma != nullptr && !p.in2 && p.Cin == 64 && (p.Cout == 64 || p.Cout % 128 == 0) && p.out_fp32 == 0 &&
           ((p.out_ld | p.out_coff) & 7) == 0 && (!p.res || (p.res_ld & 7) == 0);
  return p.w_umma != nullptr && p.KH == p.KW && conv_umma_shape_ok(p.Cin, p.Cout, p.KH, p.stride, p.pad) &&
         p.out_fp32 == 0 && ((p.out_ld | p.out_coff) & 7) == 0 && (!p.res || (p.res_ld & 7) == 0);
}

// 2x2 / stride-2 / pad-0 conv = a 1x1 conv over the space-to-depth view of the input, and that view needs no copy: for
// filter row ky, the NHWC tensor read with pixel pitch 2*Cin and row pitch 2*Wi*Cin from base + ky*Wi*Cin IS
// [rows, Hi/2, Wi/2, (kx, ci)].  The K axis is the two views back to back: K index = (ky*2 + kx)*Cin + ci (weights are
// staged in that order by launch_k2s2_to_conv1x1).
static int launch_s2k2(const ConvParams& p, cudaStream_t stream) {
  if ((p.Hi | p.Wi) & 1) return 0;
  const bool n64 = umma_tile_n(p.Cout) == 64, small = p.Wo <= TILE_W;
  const int bw = TILE_W * (small ? 1 : 2);
  using Key = std::tuple<const void*, int, int, int, int, int>;
  struct Pair { CUtensorMap m[2]; };
  static std::map<Key, Pair> cache;
  const Key key{p.in, p.rows, p.Hi, p.Wi, p.Cin, bw};
  auto it = cache.find(key);
  if (it == cache.end()) {
    Pair pr;
    const cuuint64_t dims[4] = {(cuuint64_t)2 * p.Cin, (cuuint64_t)p.Wi / 2, (cuuint64_t)p.Hi / 2, (cuuint64_t)p.rows};
    const cuuint64_t strides[3] = {(cuuint64_t)4 * p.Cin, (cuuint64_t)4 * p.Wi * p.Cin, (cuuint64_t)2 * p.Hi * p.Wi * p.Cin};
    const cuuint32_t box[4] = {64, (cuuint32_t)bw, TILE_H, 1};
    for (int ky = 0; ky < 2; ++ky)
      if (make_tmap4(p.in + (size_t)ky * p.Wi * p.Cin, dims, strides, box, &pr.m[ky]) != 0) return 0;
    if (cache.size() > 1024) cache.clear();
    it = cache.emplace(key, pr).first;
  }
  ConvParams q = p;
  q.Cin = 4 * p.Cin;       // K of the equivalent 1x1 conv
  q.Cin_real = 4 * p.Cin_real;  // FLOP accounting (launch_t counts taps x Cin_real)
  q.KH = q.KW = 1; q.stride = 1; q.pad = 0;
  q.Hi = p.Ho; q.Wi = p.Wo;
  const int split = 2 * p.Cin / 64;
  int rc;
  if (small) rc = n64 ? launch_t<64, S1K1, true, 1>(q, stream, it->second.m, split) : launch_t<128, S1K1, true, 1>(q, stream, it->second.m, split);
  else rc = n64 ? launch_t<64, S1K1, true>(q, stream, it->second.m, split) : launch_t<128, S1K1, true>(q, stream, it->second.m, split);
  return rc;
}

int launch_conv_umma(const ConvParams& p, cudaStream_t stream) {
  if (!conv_umma_eligible(p)) return 0;
  if (mode_of(p.KH, p.stride, p.pad, p.KW) == S2K2) return launch_s2k2(p, stream);
  const bool n64 = umma_tile_n(p.Cout) == 64;
  const int mode = mode_of(p.KH, p.stride, p.pad, p.KW);
  if (mode == S1K7V) return n64 ? launch_t<64, S1K7V, true>(p, stream) : launch_t<128, S1K7V, true>(p, stream);
  static const char* env_a = getenv("DYF_UMMA_A");  // "cpasync" forces the cp.async patch gather
  const bool want_tma = !(env_a && env_a[0] == 'c');
  // (descriptor base_offset stays 0: the 128-B swizzle is a function of absolute smem address bits for TMA writes
  //  and UMMA reads alike -- verified on hardware)
  if (p.in2 && !want_tma) return 0;
  if (mode == S1K3) {
    if (want_tma) {
      const bool single = p.Cin == 64 && p.Cout <= 128;  // one chunk, one n-tile: resident filter
      const int rc = n64 ? launch_t<64, S1K3, true>(p, stream)
                         : single ? launch_t<128, S1K3, true, 1>(p, stream) : launch_t<128, S1K3, true>(p, stream);
      if (rc != 0 || p.in2) return rc;
    }
    return n64 ? launch_t<64, S1K3, false>(p, stream) : launch_t<128, S1K3, false>(p, stream);
  }
  if (mode == S1K1) {
    if (want_tma && !p.in_xmap) {  // (a column subset is not a tensor-map box: cp.async gather)
      const int rc = n64 ? launch_t<64, S1K1, true>(p, stream) : launch_t<128, S1K1, true>(p, stream);
      if (rc != 0 || p.in2) return rc;
    }
    if (p.in_xmap && n64 && p.Cin == 64) return launch_t<64, S1K1, false, 2>(p, stream);
    return n64 ? launch_t<64, S1K1, false>(p, stream) : launch_t<128, S1K1, false>(p, stream);
  }
  // The stride-2 weight tiles were packed (finalize) for 64-channel chunks iff Cin % 64 == 0 and the TMA path is on: those
  // layers MUST take the TMA kernel -- when it cannot run (odd input size, no tensor map) the layer is handed back to the
  // mma.sync pipeline rather than to the 32-channel gather kernel.
  const bool w64 = p.Cin % 64 == 0 && s2k4_weights_chunk() == 64;
  if (w64 && !(want_tma && !((p.Hi | p.Wi) & 1))) return 0;
  if (w64) {  // four parity views of the input
    constexpr int T2 = Stages<S2K4, 128, true>::T;
    using Key = std::tuple<const void*, int, int, int, int>;
    struct Quad { CUtensorMap m[4]; };
    static std::map<Key, Quad> cache;
    const Key key{p.in, p.rows, p.Hi, p.Wi, p.Cin};
    auto it = cache.find(key);
    bool ok = true;
    if (it == cache.end()) {
      Quad q;
      const cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi / 2, (cuuint64_t)p.Hi / 2, (cuuint64_t)p.rows};
      const cuuint64_t strides[3] = {(cuuint64_t)4 * p.Cin, (cuuint64_t)4 * p.Wi * p.Cin, (cuuint64_t)2 * p.Hi * p.Wi * p.Cin};
      const cuuint32_t box[4] = {64, TILE_W * T2 + 1, TILE_H + 1, 1};
      for (int v = 0; v < 4 && ok; ++v)
        ok = make_tmap4(p.in + ((size_t)(v >> 1) * p.Wi + (v & 1)) * p.Cin, dims, strides, box, &q.m[v]) == 0;
      if (ok) {
        if (cache.size() > 1024) cache.clear();
        it = cache.emplace(key, q).first;
      }
    }
    if (!ok) return 0;
    return n64 ? launch_t<64, S2K4, true>(p, stream, it->second.m) : launch_t<128, S2K4, true>(p, stream, it->second.m);
  }
  return n64 ? launch_t<64, S2K4, false>(p, stream) : launch_t<128, S2K4, false>(p, stream);
}

// weights [O][64][7] (x-im2col'd stem filter, see launch_stem_xim2col_weight) -> S1K7V stage tiles
int launch_repack_umma_k7v(const float* w, act_t* out, int O, cudaStream_t s) {
  const long long total = (long long)umma_padded_cout(O) * 64 * 7;
  repack_umma_kernel<<<cdiv(total, 256), 256, 0, s>>>(w, out, O, 64, umma_tile_n(O), 7, 64, 0);
  DYF_LAUNCH_OK("repack_umma_kernel");
  return 0;
}

int launch_repack_umma(const float* w, act_t* out, int O, int I, int k, int stride, int pad, int standardize,
                       cudaStream_t s) {
  const int mode = mode_of(k, stride, pad);
  if (mode < 0) { set_error("repack_umma: unsupported geometry"); return -1; }
  const int taps = k * k, ch = mode == S2K4 ? ((I % 64 == 0) ? s2k4_weights_chunk() : 32) : 64;
  const long long total = (long long)umma_padded_cout(O) * I * taps;
  repack_umma_kernel<<<cdiv(total, 256), 256, 0, s>>>(w, out, O, I, umma_tile_n(O), taps, ch, standardize);
  DYF_LAUNCH_OK("repack_umma_kernel");
  return 0;
}

}  // namespace dyf
