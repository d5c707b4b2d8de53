// Convolutions of the spring-mesh backbone (reference: src/models/simple_conv_net.py:12-56, k = 9 / 7 / 5 / 3 on a 10 x 10
// grid) on the 5th-generation tensor cores, for images far smaller than a 16 x 8-pixel tile.
//
// Layout.  A layer input is a FLAT PADDED RASTER: image i of a logical call sits at positions
//     i * PI + y * S + x,     S = W + p,  PI = (H + p) * S,   p = (k - 1) / 2 of the layer that READS the raster,
// i.e. every image row is followed by p zero positions and every image by p zero rows, 64 channels (128 bytes) per
// position.  The zero gap after a row is the right halo of that row AND the left halo of the next one, the zero rows
// after an image are its bottom halo AND the top halo of the next image: a filter tap (ky, kx) is the plain position
// shift (ky - p) * S + (kx - p) everywhere.  So M = 128 CONSECUTIVE raster positions are one tcgen05 tile whatever the
// image size, the halo patch of a tile is one contiguous run of positions (a 2-D tensor-map box, 128-byte swizzle, zero
// fill before position 0), and every tap is a start-address shift of the same shared-memory patch -- the scheme of
// conv_umma.cu without its 16 x 8 pixel tile, which a 10 x 10 image would fill to 39 %.  Useful fraction of the GEMM
// rows: (W / S) * (H / (H + p)) = 51 % (k 9) ... 83 % (k 3) on 10 x 10 -- against 6.5 % of peak for the mma.sync path.
// Logical calls (rows sharing one time value = one set of epilogue tables) start at multiples of PC = round_up(G * PI + p,
// 256) positions, so a work item (256 positions, two M tiles sharing every weight stage) never straddles two calls.
// The first layer (C_in = 9, k = 9) reads a raster whose 192 "channels" are the 9 horizontal taps x 16 channel slots
// (written by pack_flat_kernel): 9 vertical taps x 3 chunks instead of 81 taps over a 16-channel K.
// The epilogue (BatchNorm + time scale/shift folded into per-call tables, GELU, dropout, residual) writes the valid
// pixels straight into the NEXT layer's raster (its own p); the gaps of every raster are zeroed once and never written.
//
// Pipeline = conv_umma.cu's: persistent warp-specialised CTAs (8 epilogue warps, TMA patch producer, MMA issuer, weight
// producer), weights as 8 KB stage tiles by bulk TMA, two ping-pong TMEM accumulator sets.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <tuple>

#include "conv.cuh"
#include "umma.cuh"

namespace dyf {
namespace {

constexpr int FT = 2;                       // M tiles (128 positions each) per work item
constexpr int ITEM_POS = 128 * FT;          // positions per work item
constexpr int EPI_WARPS = 8;                // 2 per TMEM lane quarter (16 were measured: no gain, the MMA side is the limiter)
constexpr int THREADS = (EPI_WARPS + 3) * 32;
constexpr int BN = 64;
constexpr int B_STAGE = BN * 64 * 2;        // one tap of one 64-channel chunk: [k8][64][8] = 8 KB
constexpr int AS = 2, MAX_BS = 16;          // patch stages; weight-ring depth is chosen at launch (what shared memory allows)

struct __align__(8) FBarriers {
  uint64_t a_full[AS], a_empty[AS], b_full[MAX_BS], b_empty[MAX_BS], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// Launch-side geometry of one layer, and the whole network call: up to FLAT_MAX_LAYERS layers run inside ONE persistent
// kernel (cooperative launch), separated by grid-wide barriers -- a layer reads, through TMA, raster positions that other
// CTAs' epilogues wrote.  All pipelines (patch ring, weight ring, TMEM ping-pong) keep their state across the layers.
struct FlatLayerCfg { int box_rows, nbox, num_work, items_per_call; FastDiv d_items, d_PI, d_S; };  // (d_*: divisions by items_per_call / PI_in / S_in)
struct FlatNetParams {
  int nlayers, BS, a_stage;   // weight-ring depth, patch stage stride (the largest layer's)
  FlatConvParams L[FLAT_MAX_LAYERS];
  FlatLayerCfg G[FLAT_MAX_LAYERS];
};

__global__ void __launch_bounds__(THREADS, 1) conv_flat_net_kernel(const __grid_constant__ FlatNetParams P,
                                                                   const __grid_constant__ CUtensorMap tm0,
                                                                   const __grid_constant__ CUtensorMap tm1,
                                                                   const __grid_constant__ CUtensorMap tm2,
                                                                   const __grid_constant__ CUtensorMap tm3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int a_stage = P.a_stage, BS = P.BS;
  uint8_t* sA = smem;
  uint8_t* sB = smem + AS * a_stage;
  FBarriers* bars = reinterpret_cast<FBarriers*>(sB + BS * B_STAGE);
  float* sTab = reinterpret_cast<float*>(sB + BS * B_STAGE + ((sizeof(FBarriers) + 15) & ~15));  // [2][A | B][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MMA_WARP = EPI_WARPS + 1;

  if (tid == 0) {
    for (int i = 0; i < AS; ++i) { mbar_init(smem_u32(&bars->a_full[i]), 1); mbar_init(smem_u32(&bars->a_empty[i]), 1); }
    for (int i = 0; i < BS; ++i) { mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(2 * FT * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // pipeline state of this thread's role, carried across the layers
  int e_it = 0;                                        // epilogue: items drained so far
  int a_ca = 0;                                        // patch producer: stages issued so far
  int m_sa = 0, m_pa = 0, m_sb = 0, m_pb = 0, m_it = 0;  // MMA issuer: ring positions / parities, items issued
  int w_sb = 0, w_pb = 1;                              // weight producer

  for (int layer = 0; layer < P.nlayers; ++layer) {
  const FlatConvParams& p = P.L[layer];
  const int box_rows = P.G[layer].box_rows, nbox = P.G[layer].nbox, num_work = P.G[layer].num_work,
            items_per_call = P.G[layer].items_per_call;
  const FastDiv d_items = P.G[layer].d_items, d_PI = P.G[layer].d_PI, d_S = P.G[layer].d_S;
  const int nchunks = p.Cin / 64;
  if (warp < EPI_WARPS) {
    // =============================== epilogue: TMEM -> tables / activation / dropout / residual -> next raster ==========
    // Plain layers: warp = (TMEM lane quarter, column half), both tiles.  Last layer with the 1x1 head fused in
    // (head_out != nullptr): warp = (lane quarter, tile), all 64 columns, so that one thread owns a whole pixel and the
    // head's C_out dot products (fp32 activations x fp32 weights) need no cross-warp reduction; the 64-channel map is
    // then never written.  (A 16-warp variant -- one 128 x 32 unit per warp -- was measured: no gain.)
    const bool fused_head = p.head_out != nullptr;
    const int quarter = warp & 3, sub4 = warp >> 2;
    const uint32_t thresh = p.drop.thresh;
    const float dscale = p.drop.scale;
    float* const sHead = sTab + 4 * BN;  // [64][8] head weights (channel-major), [8] bias
    if (fused_head) {
      for (int i = tid; i < BN * 8; i += EPI_WARPS * 32) sHead[i] = (i & 7) < p.head_oc ? __ldg(p.head_w + (size_t)(i & 7) * BN + (i >> 3)) : 0.f;
      if (tid < 8) sHead[BN * 8 + tid] = tid < p.head_oc ? __ldg(p.head_b + tid) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    }
    int tab_key = -1, tab_buf = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++e_it) {
      const int it = e_it;
      const int call = (int)d_items.div((uint32_t)w), t0 = (w - call * items_per_call) * ITEM_POS;
      const int acc = it & 1;
      if (call != tab_key) {  // epilogue tables of this logical call -> shared memory (other buffer: one barrier suffices)
        tab_key = call;
        tab_buf ^= 1;
        float* const dst = sTab + tab_buf * 2 * BN;
        if (tid < BN) {
          dst[tid] = __ldg(p.tabA + (size_t)call * BN + tid);
          dst[BN + tid] = __ldg(p.tabB + (size_t)call * BN + tid);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      }
      const float* const tab = sTab + tab_buf * 2 * BN;
      // units of work of this warp: (tile, first column) pairs
      int u_tile[2], u_col[2];
      int nu;
      if constexpr (EPI_WARPS == 8) {  // warp = (quarter, column half) x both tiles, or (quarter, tile) x both halves
        nu = 2;
        u_tile[0] = fused_head ? sub4 : 0; u_tile[1] = fused_head ? sub4 : 1;
        u_col[0] = fused_head ? 0 : sub4 * 32; u_col[1] = fused_head ? 32 : sub4 * 32;
      } else {                         // 16 warps: one unit each; the fused head uses warps 0-7 with two units
        nu = fused_head ? (sub4 < 2 ? 2 : 0) : 1;
        u_tile[0] = fused_head ? (sub4 & 1) : (sub4 >> 1); u_tile[1] = u_tile[0];
        u_col[0] = fused_head ? 0 : (sub4 & 1) * 32; u_col[1] = 32;
      }
      // geometry + residual of both units before the accumulator is waited for: their latency hides behind the MMAs
      int lp[2], img[2], yy[2], xx[2];
      bool valid[2];
      uint4 rsd[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u >= nu) { valid[u] = false; continue; }
        lp[u] = t0 + u_tile[u] * 128 + quarter * 32 + lane;
        img[u] = (int)d_PI.div((uint32_t)lp[u]);
        const int rem = lp[u] - img[u] * p.PI_in;
        yy[u] = (int)d_S.div((uint32_t)rem); xx[u] = rem - yy[u] * p.S_in;
        valid[u] = img[u] < p.G && yy[u] < p.H && xx[u] < p.W;
        if (p.res && valid[u]) {  // residual = this layer's input at the same pixel = the same raster position
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + ((size_t)call * p.PC_in + lp[u]) * BN + u_col[u]);
#pragma unroll
          for (int cs = 0; cs < 4; cs += 2) ld_global_nc_256(rp + cs, rsd[u][cs], rsd[u][cs + 1]);
        }
      }
      mbar_wait(smem_u32(&bars->acc_full[acc]), (it >> 1) & 1);
      tc_fence_after();
      float hacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u >= nu) break;  // (warp-uniform)
        const int cbeg = u_col[u];
        uint32_t v[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc * FT + u_tile[u]) * BN + cbeg, v);
        tmem_ld_wait();
        if (valid[u]) {  // (gap / padding positions: the next raster keeps its zeros)
          const float* const tA = tab + cbeg;
          const float* const tB = tA + BN;
          float o[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(tA + 4 * q), b = *reinterpret_cast<const float4*>(tB + 4 * q);
            o[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), a.x, b.x);
            o[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), a.y, b.y);
            o[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), a.z, b.z);
            o[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), a.w, b.w);
          }
          if (p.act == ACT_GELU) {  // (branch hoisted out of the element loop: the shipped SimpleConvNet is all GELU)
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = gelu_fast(o[j]);
          } else if (p.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = apply_act(o[j], p.act);
          }
          if (thresh) {  // element order of the NHWC tensor [row][y][x][64]: the masks of the mma.sync path
            const DropRow dr = drop_row(p.drop, call * p.G + img[u], (uint64_t)p.H * p.W * BN);
#pragma unroll
            for (int cs = 0; cs < 32; cs += 8) {
              drop_apply8(p.drop, dr, (uint64_t)(yy[u] * p.W + xx[u]) * BN + cbeg + cs, o + cs);
            }
          }
          if (p.res) {
#pragma unroll
            for (int cs = 0; cs < 4; ++cs) {
              float f[8];
              unpack8(rsd[u][cs], f);
#pragma unroll
              for (int j = 0; j < 8; ++j) o[8 * cs + j] += f[j];
            }
          }
          if (fused_head) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float4 w0 = *reinterpret_cast<const float4*>(sHead + (cbeg + j) * 8), w1 = *reinterpret_cast<const float4*>(sHead + (cbeg + j) * 8 + 4);
              hacc[0] = fmaf(o[j], w0.x, hacc[0]); hacc[1] = fmaf(o[j], w0.y, hacc[1]);
              hacc[2] = fmaf(o[j], w0.z, hacc[2]); hacc[3] = fmaf(o[j], w0.w, hacc[3]);
              hacc[4] = fmaf(o[j], w1.x, hacc[4]); hacc[5] = fmaf(o[j], w1.y, hacc[5]);
              hacc[6] = fmaf(o[j], w1.z, hacc[6]); hacc[7] = fmaf(o[j], w1.w, hacc[7]);
            }
          } else {
            uint4* op = reinterpret_cast<uint4*>(p.out + ((size_t)call * p.PC_out + (size_t)img[u] * p.PI_out + yy[u] * p.S_out + xx[u]) * BN + cbeg);
#pragma unroll
            for (int cs = 0; cs < 4; cs += 2) st_global_256(op + cs, pack8(o + 8 * cs), pack8(o + 8 * cs + 8));  // (rows of 128 B, cbeg % 32 == 0)
          }
        }
        __syncwarp();  // the next tcgen05.ld is warp-collective
      }
      if (fused_head && nu && valid[0]) {  // fp32 NCHW network output [rows][C_out][H][W]
        const int HW = p.H * p.W;
        float* const op = p.head_out + (size_t)(call * p.G + img[0]) * p.head_oc * HW + yy[0] * p.W + xx[0];
#pragma unroll
        for (int oc = 0; oc < 8; ++oc)
          if (oc < p.head_oc) op[(size_t)oc * HW] = hacc[oc] + sHead[BN * 8 + oc];
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->acc_empty[acc]));
    }
  } else if (warp == EPI_WARPS) {
    // =============================== patch producer: nbox TMA boxes of consecutive positions per channel chunk ==========
    if (lane == 0) {
      const CUtensorMap* tm = layer == 0 ? &tm0 : layer == 1 ? &tm1 : layer == 2 ? &tm2 : &tm3;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int call = (int)d_items.div((uint32_t)w), t0 = (w - call * items_per_call) * ITEM_POS;
        const int g0 = call * p.PC_in + t0 - p.halo;  // first patch position (negative before the raster: zero fill)
        for (int c = 0; c < nchunks; ++c, ++a_ca) {
          const int st = a_ca % AS;
          mbar_wait(smem_u32(&bars->a_empty[st]), ((a_ca / AS) & 1) ^ 1);
          const uint32_t bar = smem_u32(&bars->a_full[st]);
          mbar_expect_tx(bar, (uint32_t)(nbox * box_rows * 128));
          for (int b = 0; b < nbox; ++b)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(sA + st * a_stage + b * box_rows * 128)), "l"(reinterpret_cast<uint64_t>(tm)),
                           "r"(c * 64), "r"(g0 + b * box_rows), "r"(bar) : "memory");
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer ==========================================================================
    const uint32_t leader = elect_one();
    const uint32_t idesc = (1u << 4) | (DYF_UMMA_FMT << 7) | (DYF_UMMA_FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hi = (uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);  // SBO = 8 positions, SWIZZLE_128B
    const uint32_t b_hi = (uint32_t)((128 >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t a_lo0 = (1u << 16) | (smem_u32(sA) >> 4);
    const uint32_t b_lo0 = ((uint32_t)((BN * 16) >> 4) << 16) | (smem_u32(sB) >> 4);
    const uint32_t bar_a_full = smem_u32(&bars->a_full[0]), bar_a_empty = smem_u32(&bars->a_empty[0]);
    const uint32_t bar_b_full = smem_u32(&bars->b_full[0]), bar_b_empty = smem_u32(&bars->b_empty[0]);
    constexpr int AK = 32 >> 4, BK = (2 * BN * 16) >> 4;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++m_it) {
      const int acc = m_it & 1;
      mbar_wait(smem_u32(&bars->acc_empty[acc]), ((m_it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * FT * BN;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar_a_full + m_sa * 8, m_pa);
        tc_fence_after();
        const uint64_t a_st = ((uint64_t)a_hi << 32) | (a_lo0 + m_sa * (a_stage >> 4));
#pragma unroll 1
        for (int tap = 0; tap < p.ntaps; ++tap) {
          mbar_wait(bar_b_full + m_sb * 8, m_pb);
          tc_fence_after();
          const uint64_t b_st = ((uint64_t)b_hi << 32) | (b_lo0 + m_sb * (B_STAGE >> 4));
          const int a_off = (p.halo + p.shift[tap]) * 128;  // tap = shift of the patch start by whole positions
#pragma unroll
          for (int tile = 0; tile < FT; ++tile)
            umma_tap<4, AK, BK>(tmem_acc + tile * BN, a_st + (uint64_t)((a_off + tile * 128 * 128) >> 4), b_st, idesc,
                                (tap | c) ? 1u : 0u, leader);
          umma_commit_if(bar_b_empty + m_sb * 8, leader);
          if (++m_sb == BS) { m_sb = 0; m_pb ^= 1; }
        }
        umma_commit_if(bar_a_empty + m_sa * 8, leader);
        if (++m_sa == AS) { m_sa = 0; m_pa ^= 1; }
      }
      umma_commit_if(smem_u32(&bars->acc_full[acc]), leader);
    }
  } else {
    // =============================== weight producer: one 8 KB stage per (chunk, tap) ====================================
    if (lane == 0) {
      const int per_item = nchunks * p.ntaps;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w + (size_t)(blockIdx.x % p.wrep) * p.wrep_stride);
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        for (int i = 0; i < per_item; ++i) {
          mbar_wait(smem_u32(&bars->b_empty[w_sb]), w_pb);
          mbar_expect_tx(smem_u32(&bars->b_full[w_sb]), B_STAGE);
          bulk_g2s(smem_u32(sB + w_sb * B_STAGE), src + (size_t)i * B_STAGE, B_STAGE, smem_u32(&bars->b_full[w_sb]));
          if (++w_sb == BS) { w_sb = 0; w_pb ^= 1; }
        }
      }
    }
  }
  if (layer + 1 < P.nlayers) {
    // the epilogues' global stores (generic proxy) must be visible to the TMA loads (async proxy) of every other CTA
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    cooperative_groups::this_grid().sync();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * FT * BN));
  }
}

// ------------------------------------------------------------------------------------------------ first-layer raster
// fp32 NCHW sources (concat order) -> raster of the first layer: position (call, img, y, x) holds, for every horizontal tap
// kx, the CP channel slots of pixel (y, x + kx - p) (zero outside the image); Cflat = round_up(k * CP, 64) channels.
// One thread per SOURCE pixel: it builds the pixel's slot vector once (one coalesced load per channel plane) and stores it
// into the <= k positions that see it through one of their horizontal taps.
__global__ void __launch_bounds__(256) pack_flat_kernel(const FlatPackParams p) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;   // (row, y, xs): 32-bit (checked by the launcher)
  if (idx >= (unsigned)p.rows * p.H * p.W) return;
  const unsigned t = idx / (unsigned)p.W;
  const int xs = (int)(idx - t * p.W);
  const int r = (int)(t / (unsigned)p.H), y = (int)(t - (unsigned)r * p.H);
  const int call = (int)((unsigned)r / (unsigned)p.G), img = r - call * p.G, rs = (int)((unsigned)r % (unsigned)p.src_rows);
  const int pix = y * p.W + xs;
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = c < p.n_slots ? __ldg(p.slot_ptr[c] + (size_t)rs * p.slot_rstride[c] + pix) : 0.f;
  const uint4 lo = pack8(v), hi = pack8(v + 8);
  act_t* const row0 = p.out + ((size_t)call * p.PC + (size_t)img * p.PI + y * p.S) * p.Cflat;
  const int pad = (p.k - 1) / 2;
  for (int kx = 0; kx < p.k; ++kx) {
    const int x = xs - kx + pad;  // the position whose tap kx reads this pixel
    if (x < 0 || x >= p.W) continue;
    uint4* o = reinterpret_cast<uint4*>(row0 + (size_t)x * p.Cflat + kx * p.CP);
    o[0] = lo;
    if (p.CP > 8) o[1] = hi;
  }
}

// weights fp32 [64, Cin, k, k] -> stage tiles of the first layer: [chunk][tap = ky][k8][n][8] over the raster channels
// ci' = kx * CP + c (zero for the slots without a source channel)
__global__ void __launch_bounds__(256) repack_flat_first_kernel(const float* __restrict__ w, act_t* __restrict__ out, int Cin,
                                                                int k, int CP, int Cflat) {
  const long long total = (long long)(Cflat / 64) * k * 8 * BN * 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long r = idx;
  const int e = (int)(r % 8); r /= 8;
  const int n = (int)(r % BN); r /= BN;
  const int k8 = (int)(r % 8); r /= 8;
  const int ky = (int)(r % k); r /= k;
  const int chunk = (int)r;
  const int ci = chunk * 64 + k8 * 8 + e, kx = ci / CP, c = ci - kx * CP;
  float v = 0.f;
  if (kx < k && c < Cin) v = w[(((size_t)n * Cin + c) * k + ky) * k + kx];
  out[idx] = f2act(v);
}

// weights fp32 [64, Cin, k, k] (Cin a multiple of 64) -> stage tiles [chunk][tap = ky * k + kx][k8][n][8]
__global__ void __launch_bounds__(256) repack_flat_kernel(const float* __restrict__ w, act_t* __restrict__ out, int Cin, int k) {
  const int taps = k * k;
  const long long total = (long long)(Cin / 64) * taps * 8 * BN * 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long r = idx;
  const int e = (int)(r % 8); r /= 8;
  const int n = (int)(r % BN); r /= BN;
  const int k8 = (int)(r % 8); r /= 8;
  const int tap = (int)(r % taps); r /= taps;
  const int ci = (int)r * 64 + k8 * 8 + e;
  out[idx] = f2act(w[((size_t)n * Cin + ci) * taps + tap]);
}

int make_tmap2(const act_t* base, long long positions, int C, int box_rows, CUtensorMap* out) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return -1;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)positions};
  cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t est[2] = {1, 1};
  return fn(out, DYF_TMAP_DTYPE, 2, const_cast<act_t*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace

FlatGeo flat_geo(int H, int W, int k, int G, bool hgap) {
  FlatGeo g;
  g.p = (k - 1) / 2;
  g.S = W + (hgap ? g.p : 0);  // no gap after a row when the layer has no horizontal taps (first layer: they are channels)
  g.PI = (H + g.p) * g.S;
  const long long need = (long long)G * g.PI + (hgap ? g.p : 0);
  g.PC = (int)((need + ITEM_POS - 1) / ITEM_POS * ITEM_POS);
  return g;
}

int flat_weight_replicas() {
  static const int r = getenv("DYF_FLAT_WREP") ? std::max(1, std::min(64, atoi(getenv("DYF_FLAT_WREP")))) : 8;
  return r;
}

bool conv_flat_shape_ok(int H, int W, int k, int Cout) {
  const int p = (k - 1) / 2;
  return (k & 1) && k <= 9 && Cout == 64 && p * (W + p) + p <= 64;  // halo <= 64 positions: the patch fits two 256-row boxes
}

int launch_pack_flat(const FlatPackParams& p, cudaStream_t s) {
  ProfScope prof(s, KC_PACK);
  const long long total = (long long)p.rows * p.H * p.W;
  if (total > 0x7fffffffLL) { set_error("pack_flat: too many source pixels"); return -1; }
  pack_flat_kernel<<<cdiv(total, 256), 256, 0, s>>>(p);
  DYF_LAUNCH_OK("pack_flat_kernel");
  return 0;
}

int launch_repack_flat_first(const float* w, act_t* out, int Cin, int k, int CP, int Cflat, cudaStream_t s) {
  const long long total = (long long)(Cflat / 64) * k * 8 * BN * 8;
  repack_flat_first_kernel<<<cdiv(total, 256), 256, 0, s>>>(w, out, Cin, k, CP, Cflat);
  DYF_LAUNCH_OK("repack_flat_first_kernel");
  return 0;
}

int launch_repack_flat(const float* w, act_t* out, int Cin, int k, cudaStream_t s) {
  const long long total = (long long)(Cin / 64) * k * k * 8 * BN * 8;
  repack_flat_kernel<<<cdiv(total, 256), 256, 0, s>>>(w, out, Cin, k);
  DYF_LAUNCH_OK("repack_flat_kernel");
  return 0;
}

int launch_conv_flat_net(const FlatConvParams* layers, int nlayers, cudaStream_t stream) {
  if (nlayers < 1 || nlayers > FLAT_MAX_LAYERS) { set_error("conv_flat: 1..4 layers per launch"); return -1; }
  static int num_sms = 0, configured = 0;
  if (!num_sms) {
    int dev = 0;
    DYF_CUDA_OK(cudaGetDevice(&dev));
    DYF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  FlatNetParams P{};
  P.nlayers = nlayers;
  CUtensorMap maps[FLAT_MAX_LAYERS]{};
  int max_work = 0;
  double flops = 0.0, bytes = 0.0;
  using Key = std::tuple<const void*, long long, int, int>;
  static std::map<Key, CUtensorMap> cache;
  for (int l = 0; l < nlayers; ++l) {
    const FlatConvParams& p = layers[l];
    if (p.Cin % 64 || p.ntaps < 1 || p.ntaps > 81 || p.halo > 64 || p.PC_in % ITEM_POS) { set_error("conv_flat: bad geometry"); return -1; }
    const int np = ITEM_POS + 2 * p.halo;                       // patch positions of one work item
    const int nbox = (np + 255) / 256;
    const int box_rows = (((np + nbox - 1) / nbox) + 7) & ~7;   // every box starts on a 1024-byte swizzle atom
    P.a_stage = std::max(P.a_stage, nbox * box_rows * 128);
    P.L[l] = p;
    P.G[l].box_rows = box_rows; P.G[l].nbox = nbox;
    P.G[l].items_per_call = p.PC_in / ITEM_POS;
    P.G[l].num_work = p.calls * P.G[l].items_per_call;
    P.G[l].d_items = FastDiv((uint32_t)P.G[l].items_per_call); P.G[l].d_PI = FastDiv((uint32_t)p.PI_in); P.G[l].d_S = FastDiv((uint32_t)p.S_in);
    max_work = std::max(max_work, P.G[l].num_work);
    const long long positions = (long long)p.calls * p.PC_in;
    const Key key{p.in, positions, p.Cin, box_rows};
    auto it = cache.find(key);
    if (it == cache.end()) {
      CUtensorMap m;
      if (make_tmap2(p.in, positions, p.Cin, box_rows, &m) != 0) { set_error("conv_flat: tensor map encoding failed"); return -1; }
      if (cache.size() > 1024) cache.clear();
      it = cache.emplace(key, m).first;
    }
    maps[l] = it->second;
    flops += 2.0 * (double)p.calls * p.G * p.H * p.W * BN * p.flops_k;
    bytes += 2.0 * ((double)positions * p.Cin + (double)p.calls * p.G * p.H * p.W * BN) + 2.0 * p.ntaps * p.Cin * BN;
  }
  const int fixed = AS * P.a_stage + (((int)sizeof(FBarriers) + 15) & ~15) + 4 * BN * 4 + (BN * 8 + 8) * 4 + 64;
  static const int env_bs = getenv("DYF_FLAT_BS") ? atoi(getenv("DYF_FLAT_BS")) : 0;
  int BS = std::min(MAX_BS, (227 * 1024 - fixed) / B_STAGE);
  if (env_bs >= 2 && env_bs < BS) BS = env_bs;
  if (BS < 2) { set_error("conv_flat: patch too large for shared memory"); return -1; }
  P.BS = BS;
  const int smem = fixed + BS * B_STAGE;
  if (smem > configured) {
    DYF_CUDA_OK(cudaFuncSetAttribute(conv_flat_net_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int grid = max_work < num_sms ? max_work : num_sms;
  ProfScope prof(stream, KC_CONV_FLAT, flops, bytes);
  if (nlayers == 1) {
    conv_flat_net_kernel<<<grid, THREADS, smem, stream>>>(P, maps[0], maps[1], maps[2], maps[3]);
  } else {  // grid-wide barriers between the layers: cooperative launch (one CTA per SM, all co-resident)
    void* args[] = {&P, &maps[0], &maps[1], &maps[2], &maps[3]};
    DYF_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(conv_flat_net_kernel), dim3(grid), dim3(THREADS), args,
                                            (size_t)smem, stream));
  }
  DYF_LAUNCH_OK("conv_flat_net_kernel");
  return 0;
}

int launch_conv_flat(const FlatConvParams& p, cudaStream_t stream) { return launch_conv_flat_net(&p, 1, stream); }

}  // namespace dyf
