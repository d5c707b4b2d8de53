// Implicit-GEMM convolution over bf16 NHWC grids: parameters shared by the two implementations
//   conv_mma.cu  -- mma.sync (HMMA) pipeline: any shape, used for odd/small layers
//   conv_umma.cu -- tcgen05.mma (UMMA) + TMEM accumulators + TMA weight tiles: the heavy layers
#pragma once
#include "common.cuh"

namespace dyf {

struct ConvParams {
  const act_t* in;   // [rows, Hi, Wi, Cin]      (Cin % 8 == 0)
  const act_t* in2;  // optional second source [rows, Hi, Wi, Cin - Cin0]: the input is the channel concat
  int Cin0;                  //   [in | in2] without a concat buffer (tcgen05 TMA path only); 0 = single source
  const int* in_xmap;        // optional (1x1, tcgen05 gather path only): output column x reads input column in_xmap[x], Wo
                             //   entries -- a 1x1 conv evaluated on a subset of the input columns (NS readout)
  const act_t* w_umma;  // weights re-packed as UMMA stage tiles (conv_umma.cu) or nullptr
  const act_t* w;    // [Cout, Kpad]  k = (ky*KW + kx)*Cin + c, zero padded to Kpad (multiple of 32)
  void* out;                 // bf16 [M, out_ld] (+out_coff) or fp32 when out_fp32
  const float* tabA;         // [rows, Cout]  y = acc * A + B   (folded norm, conv bias, time scale/shift)
  const float* tabB;         // [rows, Cout]
  int tab_div;               // table row of batch row r is r / tab_div (rows of one logical call share a table)
  const act_t* res;  // optional residual added after activation+dropout, [M, res_ld]
  int rows, Hi, Wi, Cin;
  int Cin_real;              // un-padded input channels (FLOP accounting only)
  int Ho, Wo, Cout;
  int KH, KW, stride, pad;
  int K, Kpad;
  int out_ld, out_coff, res_ld;
  int act;
  int out_fp32;              // 0 = bf16 NHWC, 1 = fp32 NHWC, 2 = fp32 NCHW [rows, Cout, Ho, Wo] (network heads)
  long long M;               // rows * Ho * Wo
  DropCfg drop;
};

// Fused epilogue for 8 consecutive output channels [c0, c0+8) of GEMM row m.  acc[8] are the fp32 accumulators.
__device__ __forceinline__ void conv_epilogue8(const ConvParams& p, long long m, int c0, const float* acc) {
  const int HoWo = p.Ho * p.Wo;
  const int r = (int)(m / HoWo);
  const float* A = p.tabA + (size_t)(r / p.tab_div) * p.Cout + c0;
  const float* B = p.tabB + (size_t)(r / p.tab_div) * p.Cout + c0;
  float v[8];
  const bool full = (c0 + 8 <= p.Cout) && ((p.Cout & 3) == 0);
  if (full) {
    float4 a0 = __ldg(reinterpret_cast<const float4*>(A)), a1 = __ldg(reinterpret_cast<const float4*>(A) + 1);
    float4 b0 = __ldg(reinterpret_cast<const float4*>(B)), b1 = __ldg(reinterpret_cast<const float4*>(B) + 1);
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(fmaf(acc[j], a[j], b[j]), p.act);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = (c0 + j < p.Cout) ? apply_act(fmaf(acc[j], __ldg(A + j), __ldg(B + j)), p.act) : 0.f;
  }
  if (p.drop.thresh) {
    uint32_t keep = drop_keep_bits8(p.drop, drop_row(p.drop, r, (uint64_t)HoWo * p.Cout),
                                    (uint64_t)(m - (long long)r * HoWo) * p.Cout + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * p.drop.scale : 0.f;
  }
  if (p.res) {
    const act_t* rp = p.res + (size_t)m * p.res_ld + c0;
    if (full) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(rp)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += f[j];
    } else {
      for (int j = 0; j < 8 && c0 + j < p.Cout; ++j) v[j] += act2f(rp[j]);
    }
  }
  if (p.out_fp32 == 2) {
    const int rem = (int)(m - (long long)r * HoWo);
    float* o = reinterpret_cast<float*>(p.out) + ((size_t)r * p.Cout + c0) * HoWo + rem;
    for (int j = 0; j < 8 && c0 + j < p.Cout; ++j) o[(size_t)j * HoWo] = v[j];
  } else if (p.out_fp32) {
    float* o = reinterpret_cast<float*>(p.out) + (size_t)m * p.out_ld + p.out_coff + c0;
    if (full && ((p.out_ld | p.out_coff) & 3) == 0) {
      reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(o)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      for (int j = 0; j < 8 && c0 + j < p.Cout; ++j) o[j] = v[j];
    }
  } else {
    act_t* o = reinterpret_cast<act_t*>(p.out) + (size_t)m * p.out_ld + p.out_coff + c0;
    if (full && ((p.out_ld | p.out_coff) & 7) == 0) {
      *reinterpret_cast<uint4*>(o) = pack8(v);
    } else {
      for (int j = 0; j < 8 && c0 + j < p.Cout; ++j) o[j] = f2act(v[j]);
    }
  }
}

// Fused decoder block: bilinear x2 upsample of the concatenated [src0 | src1] low-res maps + 3x3 / pad-1 conv, as a
// composite conv Cin -> 4*Cout on the low-res grid with a depth-to-space epilogue (conv_up.cu).
struct UpConvParams {
  const act_t* src[2];  // low-res bf16 NHWC sources in concat order, [rows, H, W, ld[s]]
  int C[2], ld[2];              // channels taken from each source (multiples of 64; C[1] may be 0), channel strides
  int rows, H, W;               // low-res grid; the output is [rows, 2H, 2W, out_ld]
  int Cout;                     // output channels of the reference conv (multiple of 32); GEMM N = 4 * Cout
  const act_t* w[9];    // composite weights as tcgen05 stage tiles: interior, first row, last row, first col,
                                // last col, corners (top-left, top-right, bottom-left, bottom-right)
  act_t* out;
  int out_ld;
  const float* tabA;            // [rows / tab_div, Cout] epilogue tables (as ConvParams)
  const float* tabB;
  int tab_div;
  int act;
  DropCfg drop;
};
bool conv_up_shape_ok(int C0, int C1, int Cout, int H, int W);
size_t conv_up_weight_elems(int Cin, int Cout);
#define DYF_UP_VARIANTS 9
// nearest = 1: nn.Upsample(scale_factor=2, mode="nearest") instead of the clamped bilinear x2 (the SST Unet's up blocks)
int launch_compose_up(const float* w, int Cout, int Cin, act_t* const* w_variants, float* scratch, cudaStream_t s, int nearest = 0);
int launch_conv_up(const UpConvParams& p, cudaStream_t stream);

// ---- spring-mesh layers on flat padded rasters (conv_flat.cu)
struct FlatGeo { int p, S, PI, PC; };  // halo width, row stride, positions per image, positions per logical call
FlatGeo flat_geo(int H, int W, int k, int G, bool hgap = true);  // raster read by a k x k layer, G rows per logical call
bool conv_flat_shape_ok(int H, int W, int k, int Cout);
struct FlatConvParams {
  const act_t* in;        // raster of this layer [calls * PC_in][Cin]
  const act_t* w;         // stage tiles [chunk][tap][k8][64][8], `wrep` identical copies `wrep_stride` elements apart
  int wrep;
  long long wrep_stride;
  act_t* out;             // raster of the next layer [calls * PC_out][64] (only valid pixels are written)
  const act_t* res;       // residual (= in when Cin == 64) or nullptr
  const float* tabA;      // [calls][64] epilogue tables: y = act(acc * A + B)
  const float* tabB;
  int H, W, G, calls;     // image grid, rows per logical call, logical calls in this launch
  int S_in, PI_in, PC_in, S_out, PI_out, PC_out;
  int Cin;                // raster channels (multiple of 64)
  int ntaps, halo;        // filter taps as position shifts, max |shift|
  int shift[81];
  int act;
  double flops_k;         // reference K of the layer (Cin_real * k * k) for FLOP accounting
  // 1x1 head fused into the epilogue of the last layer (simple_conv_net.py:102): y[row][oc][H][W] = head_w[oc] . act + head_b
  const float* head_w;    // [head_oc][64] fp32, or nullptr
  const float* head_b;
  float* head_out;        // fp32 NCHW network output; when set, `out` is not written
  int head_oc;            // <= 8
  DropCfg drop;
};
struct FlatPackParams {
  const float* slot_ptr[16];      // channel plane of row 0 for every channel slot (concat order over the fp32 NCHW sources)
  long long slot_rstride[16];     // floats between consecutive rows of that source
  int n_slots, src_rows;
  int rows, H, W, k, G;   // k = kernel size of the first layer (horizontal taps packed into the channel axis)
  int CP, Cflat;          // channel slots per horizontal tap, raster channels = round_up(k * CP, 64)
  int S, PI, PC;
  act_t* out;
};
int launch_pack_flat(const FlatPackParams& p, cudaStream_t s);
int launch_conv_flat(const FlatConvParams& p, cudaStream_t stream);
#define FLAT_MAX_LAYERS 4
// up to FLAT_MAX_LAYERS consecutive layers of one network call in ONE persistent kernel (grid-wide barriers in between)
int launch_conv_flat_net(const FlatConvParams* layers, int nlayers, cudaStream_t stream);
int flat_weight_replicas();  // copies of every flat layer's filter kept in global memory (DYF_FLAT_WREP, default 8)
int launch_repack_flat(const float* w, act_t* out, int Cin, int k, cudaStream_t s);
int launch_repack_flat_first(const float* w, act_t* out, int Cin, int k, int CP, int Cflat, cudaStream_t s);

int launch_conv_mma(const ConvParams& p, cudaStream_t stream);
// Returns 1 if the tcgen05 path took the layer, 0 if the shape is not eligible (caller falls back to the mma
// pipeline), <0 on error.
int launch_conv_umma(const ConvParams& p, cudaStream_t stream);
bool conv_umma_eligible(const ConvParams& p);
bool conv_umma_shape_ok(int Cin_pad, int Cout, int k, int stride, int pad);
int umma_padded_cout(int Cout);
int launch_repack_umma(const float* w, act_t* out, int O, int I, int k, int stride, int pad, int standardize,
                       cudaStream_t s);
int launch_repack_umma_k7v(const float* w, act_t* out, int O, cudaStream_t s);  // [O][64][7] -> 7-vertical-tap stage tiles

}  // namespace dyf
