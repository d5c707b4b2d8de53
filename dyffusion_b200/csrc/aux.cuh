// Memory-bound companions of the convolution kernels: input packing/resizing, x2 upsampling + skip concat,
// GroupNorm, the Navier-Stokes readout, time-embedding tables, weight re-packing and the sampler's elementwise steps.
#pragma once
#include "common.cuh"

namespace dyf {

// ---- fp32 NCHW sources (concat order) -> bf16 NHWC [rows, Ho, Wo, Cpad], optional bilinear resize (align_corners=False)
struct PackParams {
  const float* src[6];
  int C[6];
  int nsrc;
  int src_rows;            // sources hold src_rows rows; output row r reads source row r % src_rows (batched calls)
  int rows, Hi, Wi, Ho, Wo, Cpad;
  int bilinear;            // 0 = same size copy, 1 = bilinear resize Hi x Wi -> Ho x Wo
  act_t* out;
  // "data+noise" forecaster conditioning (dyffusion.py:220-227): src[noise_src] <- w*src + (1-w)*N(0,1)
  int ones_channel;        // output channel set to 1.0 inside the image (carries a folded bias), -1 = none
  int s2d;                 // 1: space-to-depth packing -- output pixel (by, bx) holds the 2x2 block of the (resized)
                           //    grid as 4 sub-positions x 16 channel slots (slot `ones_channel` of each = 1.0)
                           // 2: x-im2col for a 7x7 stem -- 64 output channels = 7 horizontal taps x 8 channel slots
  int noise_c0, noise_c1;  // set by the launcher (x-im2col): flattened channel range of the noisy source
  int noise_src;           // -1 = none
  float noise_w;
  uint64_t seed, stream;
  const uint64_t* seed_ptr;  // device copy of the seed (CUDA-graph replays), or nullptr
  uint32_t rng_rows, row_off;  // rows per logical call / global index of the first row (see DropCfg)
};
int launch_pack(const PackParams& p, cudaStream_t s);
int launch_stem_xim2col_weight(const float* w, float* out, int O, int C, cudaStream_t s);  // [O][C][7][7] -> [O][64][7]
struct Head1x1Params {
  const act_t* x;      // [M][C]
  const float* w;      // [OC][C] fp32
  const float* bias;   // [OC] or nullptr
  float* y;            // fp32 NCHW [rows][OC][HW]
  long long M;
  int C, OC, HW;
};
int launch_head1x1(const Head1x1Params& p, cudaStream_t s);

// ---- fused stem of unet_simple: [bilinear resize ->] 1x1 conv (C_in <= 16 -> 64) + bias [+ input dropout],
//      fp32 NCHW sources -> bf16 NHWC, never materialising the resized input (reference unet_simple.py:193 + :113-116)
struct StemParams {
  PackParams pk;           // sources / geometry (pk.out and pk.Cpad unused)
  const float* w;          // [Cout, Cin] fp32 (Conv2d 1x1 weight)
  const float* bias;       // [Cout]
  act_t* out;      // [rows, Ho, Wo, Cout]
  int Cin, Cout;
  DropCfg drop;
};
int launch_stem(const StemParams& p, cudaStream_t s);

// ---- x2 upsample (bilinear align_corners=False | nearest) of up to two bf16 NHWC sources into one concat buffer
struct UpsampleParams {
  const act_t* src[2];
  int C[2];      // channels taken from each source (multiples of 8); C[1] = 0 for a single source
  int ld[2];     // channel stride of each source
  int rows, H, W;  // source spatial size; output is [rows, 2H, 2W, C[0]+C[1]]
  int scale;     // 2 = upsample x2, 1 = plain concat copy
  int bilinear;  // 1 bilinear, 0 nearest
  act_t* out;
};
int launch_upsample(const UpsampleParams& p, cudaStream_t s);

// ---- GroupNorm over bf16 NHWC (+ time scale/shift, activation, dropout, residual)
struct GroupNormParams {
  const act_t* x;   // [rows, HW, C] raw conv output (bias included)
  act_t* y;         // [rows, HW, C]
  const float* gamma;       // [C]
  const float* beta;        // [C]
  const float* tabA;        // [rows, C] (scale + 1) or nullptr
  const float* tabB;        // [rows, C] shift or nullptr
  const act_t* res; // optional residual [rows, HW, res_ld], added last
  float* stats;             // scratch, gn_scratch_floats(C, G) per row: [rows, G, 32 slabs, 2] per-slab partial sums
                            // (combined in fixed order) followed by the folded per-(row, channel) [rows, 2, C] (A, B)
  int slabs;                // set by the launcher
  int rows, HW, C, G, res_ld;
  int tab_div;              // table row = r / tab_div
  int act;
  float eps;
  DropCfg drop;
};
int launch_groupnorm(const GroupNormParams& p, cudaStream_t s);
int gn_scratch_floats(int C, int G);  // floats of `stats` scratch per row

// ---- Navier-Stokes readout: ConvTranspose2d(64->Cout, k4, s2, p1) on the Hs x Ws map followed by the bilinear
//      resize (2Hs x 2Ws) -> (Ho x Wo); only the pixels the resize samples are ever computed.  fp32 NCHW output.
struct ReadoutParams {
  const act_t* x;  // [rows, Hs, Ws, 64]
  const float* w;          // [64, Cout, 4, 4] (ConvTranspose2d layout)
  const float* bias;       // [Cout]
  float* y;                // [rows, Cout, Ho, Wo]
  int rows, Hs, Ws, Cin, Cout, Ho, Wo;
};
int launch_readout(const ReadoutParams& p, cudaStream_t s);
// Two-step readout: the channel contraction of the transposed conv runs on the tensor cores as a 1x1 conv
// 64 -> 16 * Cout (z[pixel][(ky*4+kx)*Cout + co] = sum_ci x[pixel][ci] * Wt[ci][co][ky][kx]); this kernel then gathers, for
// every output pixel, the 2x2 bilinear corners x 2x2 valid taps from z and adds the bias.
struct ReadoutGatherParams {
  const act_t* z;  // [rows, Hs, Wz, 16 * Cout]
  const float* bias;       // [Cout]
  float* y;                // [rows, Cout, Ho, Wo]
  const int* xinv;         // optional: source column -> column of z (z holds only the source columns the resize samples)
  int rows, Hs, Ws, Wz, Cout, Ho, Wo;  // Ws = source grid width, Wz = columns stored in z (= Ws without xinv)
};
int launch_readout_gather(const ReadoutGatherParams& p, cudaStream_t s);
// ConvTranspose2d weight fp32 [Cin, Cout, 4, 4] -> 1x1-conv weight [16 * Cout, Cin] (row (ky*4+kx)*Cout + co)
int launch_convt_to_conv1x1(const float* wt, float* out, int Cin, int Cout, cudaStream_t s);  // -> fp32 [16*Cout, Cin]
// Conv2d(k2, s2, p0) weight fp32 [O, C, 2, 2] -> 1x1-conv weight [O, 4C] over the space-to-depth view of the input
int launch_k2s2_to_conv1x1(const float* w, float* out, int O, int C, cudaStream_t s);

// ---- attention blocks of the SST backbone (attn_kernels.cu)
struct ChannelLNParams {
  const act_t* x;  // [M, C]
  act_t* y;        // [M, C]
  const float* g;          // [C] gain
  long long M;
  int C;
  int HW;                  // pixels per batch row (M = rows * HW)
  DropCfg drop;            // dropout on the qkv-projection input (attention.py:13)
};
int launch_channel_ln(const ChannelLNParams& p, cudaStream_t s);
struct AttnParams {
  const act_t* qkv;  // [rows, n, 3 * heads * 32]
  act_t* out;        // [rows, n, heads * 32]
  float* ctx;                // linear attention scratch [rows, heads, 32, 32]
  int rows, n, heads;
  DropCfg drop;              // dropout on the attention probabilities (full attention, attention.py:59,70)
};
int launch_linear_attention(const AttnParams& p, cudaStream_t s);
// Fused linear-attention block (attn_fused.cu): y = x + to_out(LinearAttention(Dropout(LayerNorm(x)))) in two tensor-core
// kernels; weights are the fp16 [Cout][K] rows of the generic conv weight pack (Net::wq).
struct LinAttnFusedParams {
  const act_t* x;        // [rows, n, C] block input (also the residual)
  act_t* y;              // [rows, n, C]
  const float* g;        // [C] LayerNorm gain
  const act_t* w_qkv;    // to_qkv.1.weight [3 * 128][ldw]: q | k | v rows, heads-major
  const act_t* w_out;    // to_out.weight [C][ldw_out]
  const float* b_out;    // to_out.bias [C]
  float* part;           // scratch: rows * linattn_fused_scratch_floats(n) floats
  int rows, n, C, ldw, ldw_out;
  int chunks, chunk_pix; // set by the launcher
  act_t* ctx16;          // set by the launcher: combined context per row behind the partials
  DropCfg drop;          // dropout on the LayerNorm output (attention.py:13)
};
int launch_linattn_fused(const LinAttnFusedParams& p, cudaStream_t s);
bool linattn_fused_shape_ok(int C, int heads);
size_t linattn_fused_scratch_floats(int n);
int launch_attention(const AttnParams& p, cudaStream_t s);

// ---- time embedding -> per-layer epilogue tables
struct TimeLayer {
  long long w_off;    // offset (floats) of Linear(time_dim, 2C).weight in the packed buffer, -1 = no time MLP
  long long b_off;
  long long na_off;   // folded norm multiplier [C] (or -1 -> 1)
  long long nb_off;   // folded norm/bias offset [C] (or -1 -> 0)
  long long tab_off;  // offset (floats) of this layer's tables inside tabA/tabB, per row stride = C
  int C;
  int mode;           // 0: A = na*(scale+1), B = nb*(scale+1)+shift   1: A = scale+1, B = shift (GroupNorm layers)
};
struct TimeParams {
  const float* time;      // [rows] or nullptr (no time embedding)
  const float* packed;    // all fp32 parameters of the time path + folded norm vectors
  long long w1_off, b1_off, w2_off, b2_off;  // time_emb_mlp.{1,3}
  int dim, time_dim;      // sinusoidal dim, MLP width
  const TimeLayer* layers;  // device array
  int n_layers;
  int rows;
  float* tabA;            // per layer: [rows, C] at rows * tab_off
  float* tabB;
  float* temb;            // scratch [rows, time_dim]: SiLU(time embedding)
  int total_ch;           // padded channel count over all layers (= floats per row of tabA)
};
int launch_time_tables(const TimeParams& p, cudaStream_t s);

// ---- weight re-packing (finalize)
// conv weight fp32 [O, I, KH, KW] -> bf16 [O, Kpad], k = (ky*KW+kx)*Cpad + c; optional weight standardisation
int launch_repack_conv(const float* w, act_t* out, int O, int I, int KH, int KW, int Cpad, int Kpad,
                       int standardize, cudaStream_t s);
// composite weight of (1x1 conv Wi,bi : Cs -> Cm) followed by (conv W0 : Cm -> O, KHxKW): bf16 [O, Kpad] over Cs real
// channels + one "ones" channel carrying bi (zero outside the image, like the padded 1x1 output)
int launch_compose_conv(const float* w0, const float* wi, const float* bi, act_t* out, int O, int Cm, int Cs,
                        int KH, int KW, int Cpad, int Kpad, cudaStream_t s);
// composite of (1x1 conv Wi,bi : Cs -> Cm) and (4x4/s2/p1 conv W0 : Cm -> O) expressed as a 3x3/s1/p1 conv over the
// space-to-depth packed input (4 sub-positions x 16 slots = 64 channels): fp32 [O, 64, 3, 3]
int launch_compose_s2d(const float* w0, const float* wi, const float* bi, float* out, int O, int Cm, int Cs,
                       cudaStream_t s);
// folded affine of conv-bias + eval BatchNorm: na = g*rsqrt(var+eps), nb = (bias-mean)*na + beta (bn may be null)
int launch_fold_norm(const float* bias, const float* g, const float* beta, const float* mean, const float* var,
                     float eps, float* na, float* nb, int C, cudaStream_t s);

// ---- sampler elementwise: x_s <- x_s - a + b (a may be null => x_s <- b), optional copy of the result to `out`
int launch_cold_update(float* x_s, const float* a, const float* b, float* out, long long n, cudaStream_t s);
int launch_fill(float* p, float v, long long n, cudaStream_t s);
int launch_dropout_mask(DropCfg d, long long n, uint8_t* mask, cudaStream_t s);

}  // namespace dyf
