// Host-side engine objects behind the C ABI: a backbone ("net") = parameter store + static layer plan,
// and the DYffusion sampler that drives a forecaster and an interpolator net.
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/dyffusion_b200.h"
#include "aux.cuh"
#include "conv.cuh"

namespace dyf {

struct ParamSlot {
  std::string key;
  std::vector<int64_t> shape;
  long long off = -1;   // offset (floats) inside Net::packed; -1 for ignored integer buffers
  size_t numel = 0;
  bool is_set = false;
  bool ignored = false;  // num_batches_tracked
};

struct ConvLayer {
  int w = -1, b = -1;                         // param slots: weight, bias
  int bn_g = -1, bn_b = -1, bn_m = -1, bn_v = -1;  // eval BatchNorm folded into the epilogue (or -1)
  int tw = -1, tb = -1;                       // per-block time MLP Linear(time_dim, 2*Cout)
  int Cin = 0, Cpad = 0, Cout = 0, KH = 1, KW = 1, stride = 1, pad = 0;
  bool standardize = false;                   // WeightStandardizedConv2d
  int K = 0, Kpad = 0;
  size_t wq_off = 0;                          // offset (elements) into Net::wq (bf16 packed weights)
  int xmap_cols = 0;                            // > 0: 1x1 conv over a column subset of its input (Net::ro_xmap), output width
  int up_nearest = 0;                           // fused upsample + conv (up_off): the upsample is nearest-neighbour, not bilinear
  int convt_z = 0;                              // weight = ConvTranspose2d [Cin, Cout/16, 4, 4] re-laid out as a 1x1 conv
  int comp_s2d = 0;                             // composite expressed as 3x3/s1 over the space-to-depth packed input
  int flops_cin = 0;                            // channels to count per tap in FLOP accounting (0 = Cin)
  int comp_wi = -1, comp_bi = -1, comp_cm = 0;  // composite layer: preceded by a folded 1x1 conv (weights, bias, width)
  long long wu_off = -1;                      // offset into Net::wq_umma (tcgen05 stage tiles) or -1
  long long up_off[DYF_UP_VARIANTS] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};  // fused upsample+conv: composite variants in wq_umma
  long long flat_off = -1, flat_elems = 0;    // flat-raster path (conv_flat.cu): stage tiles in wq_umma (x replicas)
  int stem_xim2col = 0;                       // 7x7 stem run as 7 vertical taps over the x-im2col'd input (64 channels)
  int flat_first = 0;                         // first layer: horizontal taps folded into the channel axis (CP slots per tap)
  int flat_cp = 0, flat_cin = 0;              // channel slots per horizontal tap, raster channels read by the layer
  long long na_off = -1, nb_off = -1;         // folded affine in Net::packed
  int table = -1;                             // index into Net::time_layers
};

struct NormLayer {  // GroupNorm applied by its own kernel
  int g = -1, b = -1, tw = -1, tb = -1;
  int C = 0, G = 8;
  int table = -1;   // (scale+1, shift) table or -1
  long long stats_off = 0;  // floats per row offset in stats scratch
};

enum OpType { OP_STEM, OP_PACK, OP_CONV, OP_CONV_UP, OP_UPSAMPLE, OP_GROUPNORM, OP_READOUT, OP_READOUT_GATHER, OP_LINATTN, OP_ATTN, OP_CHANNEL_LN, OP_FLAT_PACK, OP_FLAT_CONV, OP_LINATTN_FUSED, OP_HEAD1X1 };
constexpr int BUF_NONE = -1;

struct Op {
  OpType type;
  int in0 = BUF_NONE, in1 = BUF_NONE, out = BUF_NONE, res = BUF_NONE;
  int layer = -1;      // ConvLayer / NormLayer index
  int act = ACT_NONE;
  float drop_p = 0.f;
  int site = 0;
  int out_coff = 0;    // channel offset inside the output buffer (concat writes)
  int out_mode = 0;    // 0 bf16 NHWC buffer, 2 fp32 NCHW external output
  int c0 = 0, c1 = 0, scale = 2, bilinear = 1;  // upsample
  int ones_channel = -1;  // pack: channel carrying a folded bias
  int aux = 0;
};

struct Buf {
  int H = 0, W = 0, C = 0;
  size_t row_bytes() const { return (size_t)H * W * C * 2; }
};

struct LNLayer { int g = -1; int C = 0; };
struct FlatBuf { int k = 1, C = 64; bool hgap = true; };  // raster read by a k x k layer, C channels per position (conv_flat.cu)  // channel LayerNorm (gain only) in front of an attention block

struct Net {
  dyf_net_desc d{};
  std::vector<ParamSlot> params;
  std::map<std::string, int> index;
  std::vector<ConvLayer> convs;
  std::vector<NormLayer> norms;
  std::vector<LNLayer> lns;
  std::vector<Op> ops;
  std::vector<Buf> bufs;
  std::vector<FlatBuf> flats;          // rasters of the flat path; device memory below, zeroed when (re)allocated
  act_t* flat_mem = nullptr;
  size_t flat_bytes = 0;
  int flat_G = 0, flat_calls = 0;      // geometry the rasters are currently laid out (and zeroed) for
  std::vector<size_t> flat_offs;
  int ensure_flat(int G, int calls, cudaStream_t s);
  std::vector<TimeLayer> time_layers;
  int t_w1 = -1, t_b1 = -1, t_w2 = -1, t_b2 = -1;
  int ro_w = -1, ro_b = -1;  // readout params
  // NS readout on the source columns the final resize samples: virtual column -> source column (ro_xmap) and back (ro_xinv)
  std::vector<int> ro_xmap, ro_xinv;
  int* ro_tab_dev = nullptr;  // [ro_xmap | ro_xinv] on the device (finalize)
  int ro_src_w = 0;
  int time_dim = 0;
  long long tab_floats_per_row = 0;    // sum of C over time_layers
  long long stats_floats_per_row = 0;
  long long packed_floats = 0, extra_floats = 0;
  size_t wq_elems = 0, wu_elems = 0;
  float* packed = nullptr;             // device: all fp32 params + folded vectors
  act_t* wq = nullptr;         // device: packed bf16 conv weights
  act_t* wq_umma = nullptr;    // device: weights of tcgen05-eligible layers as UMMA stage tiles
  TimeLayer* d_time_layers = nullptr;  // device copy
  bool finalized = false;
  uint64_t generation = 0;             // bumped whenever device buffers a captured graph may point to are released
  int Hin = 0, Win = 0;                // network grid (after the optional outer resize)
  // epilogue tables of time tuples seen before (the sampler's schedule is fixed, so every tuple recurs on every call):
  // key = the times of one forward's logical calls, value = device buffer [tabA | tabB | scratch]
  std::map<std::vector<float>, float*> tab_cache;
  void clear_tab_cache();

  ~Net();
  int add_param(const std::string& key, std::vector<int64_t> shape, bool ignored = false);
  int add_buf(int H, int W, int C);
  int build();                         // dispatch on d.arch
  int build_unet_simple();
  int build_convnet();
  int build_unet_resnet();
  int resnet_block(const std::string& prefix, int x, int Cin, int Cout, int& site, int x2 = BUF_NONE);  // x2: second source of a channel concat [x | x2] that is never materialised
  int attention_block(const std::string& prefix, int x, int C, bool linear, int& site);
  int add_conv(const std::string& wkey, int Cin, int Cout, int k, int stride, int pad, bool bias = true);
  void attach_bn(ConvLayer& c, const std::string& prefix);
  void attach_time(int& tw, int& tb, const std::string& prefix, int C);
  int set_param(const char* key, const void* data, const int64_t* shape, int ndim);
  int finalize(cudaStream_t s);
  size_t workspace_bytes(int rows) const;
  int forward(int rows, const float* const* srcs, const int* src_ch, int nsrc, const float* time, float* y,
              const RngCtx& rng, void* ws, size_t ws_bytes, cudaStream_t s,
              int noise_src = -1, float noise_w = 0.f, int src_rows = 0, int group_rows = 1,
              const float* host_times = nullptr);  // host copy of `time` (rows / group_rows values): enables the table cache
};

struct Sampler {
  Net* F = nullptr;
  Net* I = nullptr;
  dyf_sampler_desc d{};
  std::vector<double> schedule, tau, tF, refine;
  std::vector<double> out_keys;    // reference key number of every output slot
  std::vector<int> step_slot;      // schedule index -> output slot or -1
  std::vector<int> refine_slot;    // refinement index -> output slot
  // CUDA-graph replay of the launch sequence (desc.cuda_graph): one entry per (rows, workspace, row offset)
  struct GraphEntry {
    int runs = 0;                    // plain runs seen for this key (the first one warms every lazily built cache)
    bool failed = false;             // capture / instantiation failed once: stay on plain launches
    cudaGraphExec_t exec = nullptr;
    cudaGraph_t graph = nullptr;
    uint64_t genF = 0, genI = 0;     // Net::generation at capture time
    uint64_t kernels = 0;            // kernel launches inside the graph (for dyf_launch_count)
  };
  std::map<std::tuple<int, const void*, uint64_t>, GraphEntry> graphs;
  // The legacy default stream (what PyTorch's current stream is unless the caller switches) cannot be captured: graph mode
  // then runs on this sampler's own non-blocking stream, fenced against the caller's stream with events on both sides.
  uint64_t graph_replays = 0;      // runs served by cudaGraphLaunch
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  ~Sampler();
  int plan();
  size_t workspace_bytes(int rows) const;
  // the launch sequence of sample_loop on stream s (plain launches; capturable once the caches are warm)
  int enqueue(int rows, const float* ic, const float* stat, float* preds, float* x0_out, uint64_t seed,
              const uint64_t* seed_dev, uint64_t row_offset, void* ws, size_t ws_bytes, cudaStream_t s);
  int run(int rows, const float* ic, const float* stat, float* preds, float* x0_out, uint64_t seed, uint64_t row_offset,
          void* ws, size_t ws_bytes, cudaStream_t s);
  int run_on(int rows, const float* ic, const float* stat, float* preds, float* x0_out, uint64_t seed, uint64_t row_offset,
             void* ws, size_t ws_bytes, cudaStream_t s);  // run() after the stream choice
};

bool profiling_enabled();          // abi.cu: per-launch event bracketing is on (incompatible with stream capture)
uint64_t launch_counter();
struct NvtxRange { bool on; NvtxRange(const char* what, int arch, int rows); ~NvtxRange(); };

}  // namespace dyf
