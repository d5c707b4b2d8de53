// Backbone plans: parameter bookkeeping (reference state-dict keys), weight folding/re-packing, and the per-forward
// kernel sequence of the three hot-path networks.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "engine.hpp"

namespace dyf {

static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static bool stream_is_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone;
}

Net::~Net() {
  if (packed) cudaFree(packed);
  if (wq) cudaFree(wq);
  if (wq_umma) cudaFree(wq_umma);
  if (d_time_layers) cudaFree(d_time_layers);
  if (flat_mem) cudaFree(flat_mem);
  if (ro_tab_dev) cudaFree(ro_tab_dev);
  clear_tab_cache();
}

void Net::clear_tab_cache() {
  for (auto& kv : tab_cache) cudaFree(kv.second);
  tab_cache.clear();
  ++generation;  // captured graphs hold pointers into these buffers
}

int Net::add_param(const std::string& key, std::vector<int64_t> shape, bool ignored) {
  ParamSlot p;
  p.key = key;
  p.shape = std::move(shape);
  p.numel = 1;
  for (auto v : p.shape) p.numel *= (size_t)v;
  p.ignored = ignored;
  if (!ignored) {
    p.off = packed_floats;
    packed_floats += (long long)round_up((int)p.numel, 4);  // keep every slot 16-byte aligned
  }
  index[key] = (int)params.size();
  params.push_back(p);
  return (int)params.size() - 1;
}

int Net::add_buf(int H, int W, int C) {
  Buf b;
  b.H = H; b.W = W; b.C = C;
  bufs.push_back(b);
  return (int)bufs.size() - 1;
}

int Net::add_conv(const std::string& prefix, int Cin, int Cout, int k, int stride, int pad, bool bias) {
  ConvLayer c;
  c.Cin = Cin;
  c.Cpad = round_up(Cin, 8);
  c.Cout = Cout;
  c.KH = c.KW = k;
  c.stride = stride;
  c.pad = pad;
  c.K = k * k * c.Cpad;
  c.Kpad = round_up(c.K, 32);
  c.w = add_param(prefix + ".weight", {Cout, Cin, k, k});
  if (bias) c.b = add_param(prefix + ".bias", {Cout});
  c.wq_off = wq_elems;
  wq_elems += (size_t)Cout * c.Kpad;
  if (conv_umma_shape_ok(c.Cpad, Cout, k, stride, pad) && c.Cpad == Cin) {
    c.wu_off = (long long)wu_elems;
    wu_elems += (size_t)umma_padded_cout(Cout) * Cin * k * k;
  }
  c.na_off = packed_floats + extra_floats;  // resolved against the final packed size in finalize()
  extra_floats += round_up(Cout, 4);
  c.nb_off = packed_floats + extra_floats;
  extra_floats += round_up(Cout, 4);
  convs.push_back(c);
  return (int)convs.size() - 1;
}

void Net::attach_bn(ConvLayer& c, const std::string& prefix) {
  c.bn_g = add_param(prefix + ".weight", {c.Cout});
  c.bn_b = add_param(prefix + ".bias", {c.Cout});
  c.bn_m = add_param(prefix + ".running_mean", {c.Cout});
  c.bn_v = add_param(prefix + ".running_var", {c.Cout});
  add_param(prefix + ".num_batches_tracked", {}, true);
}

void Net::attach_time(int& tw, int& tb, const std::string& prefix, int C) {
  if (!d.with_time_emb) return;
  tw = add_param(prefix + ".weight", {2 * C, time_dim});
  tb = add_param(prefix + ".bias", {2 * C});
}

// NOTE: folded vectors (na/nb) live after all parameters.  add_conv() records their offsets relative to
// "packed_floats at that time + extra so far"; since parameters keep being appended afterwards, the offsets are
// re-based in build() once the parameter region is complete.
int Net::build() {
  time_dim = d.dim * 2;
  if (d.with_time_emb) {  // get_time_embedder (src/models/modules/misc.py:54-67)
    t_w1 = add_param("time_emb_mlp.1.weight", {time_dim, d.dim});
    t_b1 = add_param("time_emb_mlp.1.bias", {time_dim});
    t_w2 = add_param("time_emb_mlp.3.weight", {time_dim, time_dim});
    t_b2 = add_param("time_emb_mlp.3.bias", {time_dim});
  }
  int rc;
  switch (d.arch) {
    case DYF_ARCH_UNET_SIMPLE: rc = build_unet_simple(); break;
    case DYF_ARCH_CONVNET: rc = build_convnet(); break;
    case DYF_ARCH_UNET_RESNET: rc = build_unet_resnet(); break;
    default: set_error("unknown arch"); return DYF_ERR_ARG;
  }
  if (rc) return rc;
  // re-base folded-vector offsets behind the parameter region
  long long cursor = packed_floats;
  for (auto& c : convs) {
    c.na_off = cursor; cursor += round_up(c.Cout, 4);
    c.nb_off = cursor; cursor += round_up(c.Cout, 4);
  }
  extra_floats = cursor - packed_floats;
  // time/epilogue tables: one (A, B) table per conv, one (scale+1, shift) table per timed GroupNorm
  tab_floats_per_row = 0;
  for (auto& c : convs) {
    TimeLayer L{};
    L.w_off = c.tw >= 0 ? params[c.tw].off : -1;
    L.b_off = c.tb >= 0 ? params[c.tb].off : -1;
    L.na_off = c.na_off;
    L.nb_off = c.nb_off;
    L.tab_off = tab_floats_per_row;
    L.C = c.Cout;
    L.mode = 0;
    c.table = (int)time_layers.size();
    time_layers.push_back(L);
    tab_floats_per_row += round_up(c.Cout, 4);
  }
  stats_floats_per_row = 0;
  for (auto& n : norms) {
    n.stats_off = stats_floats_per_row;
    stats_floats_per_row += gn_scratch_floats(n.C, n.G);
    if (n.tw >= 0) {
      TimeLayer L{};
      L.w_off = params[n.tw].off;
      L.b_off = params[n.tb].off;
      L.na_off = L.nb_off = -1;
      L.tab_off = tab_floats_per_row;
      L.C = n.C;
      L.mode = 1;
      n.table = (int)time_layers.size();
      time_layers.push_back(L);
      tab_floats_per_row += round_up(n.C, 4);
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Navier-Stokes backbone (reference: src/models/unet_simple.py:86-197; layer table SURVEY.md A.1)
// ---------------------------------------------------------------------------------------------------------------
int Net::build_unet_simple() {
  const int dim = d.dim;
  const int cin = d.in_channels + d.cond_channels;
  const bool resize = d.upsample_h > 0;
  Hin = resize ? d.upsample_h : d.height;
  Win = resize ? d.upsample_w : d.width;
  if (Hin % 64 || Win % 64) {
    set_error("unet_simple: the network grid (upsample_dims) must be divisible by 64 (six stride-2 stages)");
    return DYF_ERR_UNSUPPORTED;
  }
  if (dim % 8) { set_error("unet_simple: dim must be a multiple of 8"); return DYF_ERR_UNSUPPORTED; }
  int site = 1;
  // stem: [resize ->] 1x1 conv (:113-116)
  // The stem is linear (1x1 conv, no activation): with input_dropout == 0 it is folded into the first encoder conv
  // as a composite 4x4/s2 conv over the C_in resized input channels plus a "ones" channel that carries the stem bias
  // (zero in the padding, exactly like the padded stem output).  The 64-channel full-resolution stem output -- the
  // largest activation of the encoder -- is then never materialised.  Exact in real arithmetic.
  const bool fold_stem = d.input_dropout == 0.f;
  // ... and when the input has <= 15 channels the composite runs on the tcgen05 3x3 kernel: the resized input is packed
  // space-to-depth (2x2 blocks -> 4 x 16 channel slots at half resolution), which turns the 4x4 / stride-2 composite
  // into a 3x3 / stride-1 conv over 64 channels with structurally-zero weights for the unused (block, sub-position) taps.
  const bool stem_s2d = fold_stem && cin + 1 <= 16 && (dim * 2 == 64 || (dim * 2) % 128 == 0);
  int x = BUF_NONE;
  int stem_w = -1, stem_b = -1, b_in = BUF_NONE;
  if (fold_stem) {
    stem_w = add_param("init_conv.weight", {dim, cin, 1, 1});
    stem_b = add_param("init_conv.bias", {dim});
    b_in = stem_s2d ? add_buf(Hin / 2, Win / 2, 64) : add_buf(Hin, Win, round_up(cin + 1, 8));
    Op pk{}; pk.type = OP_PACK; pk.out = b_in; pk.bilinear = resize ? 1 : 0; pk.ones_channel = cin; pk.aux = stem_s2d ? 1 : 0;
    ops.push_back(pk);
  } else {
    int ci = add_conv("init_conv", cin, dim, 1, 1, 0);
    x = add_buf(Hin, Win, dim);
    if (dim == 64 && cin <= 16) {  // fused resize + 1x1 conv: the resized input is never materialised
      Op o{}; o.type = OP_STEM; o.out = x; o.layer = ci; o.bilinear = resize ? 1 : 0; o.drop_p = d.input_dropout; o.site = site++;
      ops.push_back(o);
    } else {
      const int bi = add_buf(Hin, Win, round_up(cin, 8));
      Op pk{}; pk.type = OP_PACK; pk.out = bi; pk.bilinear = resize ? 1 : 0;
      ops.push_back(pk);
      Op o{}; o.type = OP_CONV; o.in0 = bi; o.out = x; o.layer = ci; o.act = ACT_NONE; o.drop_p = d.input_dropout; o.site = site++;
      ops.push_back(o);
    }
  }
  // encoder (:120-129): conv(k, s2) -> BN|GN -> time scale/shift -> LeakyReLU(0.2) -> Dropout
  const int enc_out[6] = {dim * 2, dim * 2, dim * 4, dim * 8, dim * 8, dim * 8};
  const int enc_k[6] = {4, 4, 4, 4, 2, 2}, enc_p[6] = {1, 1, 1, 1, 0, 0};
  int skips[6];
  int C = dim, H = Hin, W = Win;
  for (int i = 0; i < 6; ++i) {
    const std::string p = "input_ops." + std::to_string(i);
    int tw = -1, tb = -1;
    attach_time(tw, tb, p + ".time_mlp.1", enc_out[i]);
    int li = add_conv(p + ".ops.0", C, enc_out[i], enc_k[i], 2, enc_p[i]);
    if (i == 0 && fold_stem) {  // composite layer: reads the packed network input directly
      ConvLayer& c0 = convs[li];
      c0.comp_wi = stem_w; c0.comp_bi = stem_b; c0.comp_cm = dim;
      if (stem_s2d) {
        c0.comp_s2d = 1; c0.flops_cin = (16 * (cin + 1) + 8) / 9;
        c0.Cin = c0.Cpad = 64; c0.KH = c0.KW = 3; c0.stride = 1; c0.pad = 1;
        c0.K = c0.Kpad = 9 * 64;  // fits the 4x4 allocation made by add_conv (wq and the tcgen05 stage tiles)
      } else {
        c0.Cin = cin + 1; c0.Cpad = round_up(cin + 1, 8);
        c0.K = c0.KH * c0.KW * c0.Cpad; c0.Kpad = round_up(c0.K, 32);
        c0.wu_off = -1;
      }
      x = b_in;
    }
    H /= 2; W /= 2;
    if (i < 5) {
      attach_bn(convs[li], p + ".ops.1");
      convs[li].tw = tw; convs[li].tb = tb;
      int y = add_buf(H, W, enc_out[i]);
      Op o{}; o.type = OP_CONV; o.in0 = x; o.out = y; o.layer = li; o.act = ACT_LEAKY; o.drop_p = d.dropout; o.site = site++;
      ops.push_back(o);
      x = y;
    } else {  // last encoder block: GroupNorm(8) (:56, :128)
      int raw = add_buf(H, W, enc_out[i]);
      Op o{}; o.type = OP_CONV; o.in0 = x; o.out = raw; o.layer = li; o.act = ACT_NONE;
      ops.push_back(o);
      NormLayer n; n.C = enc_out[i]; n.G = 8; n.tw = tw; n.tb = tb;
      n.g = add_param(p + ".ops.1.weight", {enc_out[i]});
      n.b = add_param(p + ".ops.1.bias", {enc_out[i]});
      norms.push_back(n);
      int y = add_buf(H, W, enc_out[i]);
      Op g{}; g.type = OP_GROUPNORM; g.in0 = raw; g.out = y; g.layer = (int)norms.size() - 1; g.act = ACT_LEAKY; g.drop_p = d.dropout; g.site = site++;
      ops.push_back(g);
      x = y;
    }
    skips[i] = x;
    C = enc_out[i];
  }
  // decoder (:133-140): bilinear x2 -> conv(k-1) -> BN -> time scale/shift -> ReLU -> Dropout -> cat(skip)
  const int dec_out[6] = {dim * 8, dim * 8, dim * 4, dim * 2, dim * 2, dim};
  const int dec_k[6] = {1, 1, 3, 3, 3, 3}, dec_p[6] = {0, 0, 1, 1, 1, 1};
  for (int i = 0; i < 6; ++i) {
    const std::string p = "output_ops." + std::to_string(i);
    const int skip = i > 0 ? skips[5 - i] : BUF_NONE;
    const int c0 = bufs[x].C, c1 = skip >= 0 ? bufs[skip].C : 0;
    H *= 2; W *= 2;
    // 3x3 decoder blocks on large grids run as ONE kernel: upsample + concat + conv as a composite conv on the low-res
    // grid (conv_up.cu).  Small grids keep the two-kernel path (their border tiles would dominate).
    const char* env_min = getenv("DYF_UPFUSE_MIN");  // smallest upsampled grid side that takes the fused kernel
    const int fuse_min = env_min ? atoi(env_min) : 64;  // measured on the NS step: 64 beats 128 by 2 %; at 32 (one super-tile per
                                                        // image) the border phases + wave quantisation eat the gain
    const bool fuse_up = dec_k[i] == 3 && std::min(H, W) >= fuse_min && conv_up_shape_ok(c0, c1, dec_out[i], H / 2, W / 2) &&
                         !getenv("DYF_DISABLE_UPFUSE");
    int up = BUF_NONE;
    if (!fuse_up) {
      up = add_buf(H, W, c0 + c1);
      Op u{}; u.type = OP_UPSAMPLE; u.in0 = x; u.in1 = skip; u.out = up; u.c0 = c0; u.c1 = c1; u.scale = 2; u.bilinear = 1;
      ops.push_back(u);
    }
    int tw = -1, tb = -1;
    attach_time(tw, tb, p + ".time_mlp.1", dec_out[i]);
    int li = add_conv(p + ".ops.1", c0 + c1, dec_out[i], dec_k[i], 1, dec_p[i]);
    attach_bn(convs[li], p + ".ops.2");
    convs[li].tw = tw; convs[li].tb = tb;
    int y = add_buf(H, W, dec_out[i]);
    Op o{}; o.type = fuse_up ? OP_CONV_UP : OP_CONV; o.in0 = fuse_up ? x : up; o.in1 = fuse_up ? skip : BUF_NONE; o.out = y;
    o.c0 = c0; o.c1 = c1; o.layer = li; o.act = ACT_RELU; o.drop_p = d.dropout; o.site = site++;
    if (fuse_up) {
      for (int v = 0; v < DYF_UP_VARIANTS; ++v) { convs[li].up_off[v] = (long long)wu_elems; wu_elems += conv_up_weight_elems(c0 + c1, dec_out[i]); }
    }
    ops.push_back(o);
    x = y;
  }
  // readout (:141-150) + outer resize back to the data grid (:195)
  // The channel contraction of the ConvTranspose2d (64 -> 16 taps x Cout values per source pixel) runs as a 1x1 conv on
  // the tensor cores; a gather kernel then evaluates only the transposed-conv pixels the final resize samples.
  ro_w = add_param("readout.0.weight", {dim, d.out_channels, 4, 4});
  ro_b = add_param("readout.0.bias", {d.out_channels});
  if (dim % 8 == 0 && (16 * d.out_channels) % 8 == 0) {
    ConvLayer z;
    z.w = ro_w; z.convt_z = 1;
    z.Cin = z.Cpad = dim; z.Cout = 16 * d.out_channels; z.KH = z.KW = 1; z.stride = 1; z.pad = 0;
    z.K = dim; z.Kpad = round_up(dim, 32);
    z.wq_off = wq_elems; wq_elems += (size_t)z.Cout * z.Kpad;
    if (conv_umma_shape_ok(z.Cpad, z.Cout, 1, 1, 0)) { z.wu_off = (long long)wu_elems; wu_elems += (size_t)umma_padded_cout(z.Cout) * z.Cin; }
    // Source columns the outer resize reads (two bilinear corners x two transposed-conv taps per output column): for
    // 512 -> 42 that is 126 of 256 columns, so the 1x1 conv is evaluated on those only (cp.async gather through a column
    // table) and z is stored compactly.  Corner indices are taken for src -+ eps: a superset of what the device's float
    // arithmetic in bilinear_coord can produce.
    ro_src_w = bufs[x].W;
    int zw = bufs[x].W;
    if (z.wu_off >= 0 && !getenv("DYF_DISABLE_READOUT_COLS") && !getenv("DYF_DISABLE_UMMA")) {
      const int Ws = bufs[x].W, W2 = 2 * Ws;
      std::vector<char> need(Ws, 0);
      const double scale = (double)W2 / d.width;
      for (int ox = 0; ox < d.width; ++ox)
        for (int e = -1; e <= 1; e += 2) {
          double src = scale * (ox + 0.5) - 0.5 + e * 1e-3;
          if (src < 0) src = 0;
          int x0 = (int)src;
          if (x0 > W2 - 1) x0 = W2 - 1;
          const int xs[2] = {x0, x0 + (x0 < W2 - 1 ? 1 : 0)};
          for (int xx : xs)
            for (int dx = 0; dx < 2; ++dx) {
              const int kx = ((xx + 1) & 1) + 2 * dx, ix = (xx + 1 - kx) >> 1;
              if (ix >= 0 && ix < Ws) need[ix] = 1;
            }
        }
      std::vector<int> cols;
      for (int i = 0; i < Ws; ++i) if (need[i]) cols.push_back(i);
      const int wv = round_up((int)cols.size(), 8);
      if (!cols.empty() && wv * 4 <= Ws * 3) {
        ro_xinv.assign(Ws, 0);
        for (size_t i = 0; i < cols.size(); ++i) ro_xinv[cols[i]] = (int)i;
        ro_xmap = cols;
        ro_xmap.resize(wv, cols.back());  // padding columns re-read the last one (never gathered)
        z.xmap_cols = zw = wv;
      }
    }
    convs.push_back(z);
    int zb = add_buf(bufs[x].H, zw, 16 * d.out_channels);
    Op c{}; c.type = OP_CONV; c.in0 = x; c.out = zb; c.layer = (int)convs.size() - 1; c.act = ACT_NONE;
    ops.push_back(c);
    Op g{}; g.type = OP_READOUT_GATHER; g.in0 = zb;
    ops.push_back(g);
  } else {
    Op r{}; r.type = OP_READOUT; r.in0 = x;
    ops.push_back(r);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// spring-mesh backbone (reference: src/models/simple_conv_net.py:59-131)
// ---------------------------------------------------------------------------------------------------------------
int Net::build_convnet() {
  const int dim = d.dim;
  const int cin = d.in_channels + d.cond_channels;
  Hin = d.height; Win = d.width;
  if (dim % 8) { set_error("SimpleConvNet: dim must be a multiple of 8"); return DYF_ERR_UNSUPPORTED; }
  int site = 1;
  // Flat-raster tensor-core path (conv_flat.cu): every layer is dim -> 64 on a grid small enough for a <= 64-position halo
  bool flat = dim == 64 && d.n_kernels >= 1 && !getenv("DYF_DISABLE_FLAT") && !getenv("DYF_DISABLE_UMMA");
  for (int i = 0; flat && i < d.n_kernels; ++i) flat = conv_flat_shape_ok(Hin, Win, d.kernel_sizes[i], dim);
  if (flat) {
    const int k0 = d.kernel_sizes[0], cp = round_up(cin, 8), cflat = round_up(k0 * cp, 64);
    flat = cin <= 16 && cflat <= 256 && (k0 - 1) / 2 * (Win + (k0 - 1) / 2) <= 64;
    if (flat) {
      FlatBuf fb; fb.k = k0; fb.C = cflat; fb.hgap = false;  // horizontal taps live in the channel axis
      flats.push_back(fb);
      Op pk{}; pk.type = OP_FLAT_PACK; pk.out = 0;
      ops.push_back(pk);
      int C = cin, last = BUF_NONE;
      for (int i = 0; i < d.n_kernels; ++i) {
        const std::string p = "convs." + std::to_string(i);
        const int k = d.kernel_sizes[i];
        int li = add_conv(p + ".conv", C, dim, k, 1, (k - 1) / 2);
        attach_bn(convs[li], p + ".norm");
        attach_time(convs[li].tw, convs[li].tb, p + ".time_mlp.1", dim);
        ConvLayer& c = convs[li];
        c.flat_first = i == 0; c.flat_cp = cp; c.flat_cin = i == 0 ? cflat : dim;
        c.flat_off = (long long)wu_elems;
        c.flat_elems = (long long)(i == 0 ? k : k * k) * c.flat_cin * 64;
        wu_elems += (size_t)c.flat_elems * flat_weight_replicas();
        Op o{}; o.type = OP_FLAT_CONV; o.in0 = i; o.layer = li; o.act = ACT_GELU; o.drop_p = d.dropout; o.site = site++;
        if (d.residual && C == dim) o.res = i;  // residual only when C_in == C_out (:31, :53-54): the layer's own input raster
        if (i + 1 < d.n_kernels) {
          FlatBuf nb; nb.k = d.kernel_sizes[i + 1]; nb.C = dim;
          flats.push_back(nb);
          o.out = i + 1;
        } else {
          last = add_buf(Hin, Win, dim);  // plain NHWC for the 1x1 head (unused when the head is fused into this layer)
          o.out = last; o.aux = 1;
        }
        ops.push_back(o);
        C = dim;
      }
      int hl = add_conv("head", dim, d.out_channels, 1, 1, 0);
      if (d.out_channels <= 8 && !getenv("DYF_FLAT_NO_HEAD_FUSION")) {
        ops.back().aux = 2;          // last conv layer computes the head in its epilogue
        ops.back().c0 = hl;
      } else {
        Op h{}; h.type = OP_CONV; h.in0 = last; h.layer = hl; h.out_mode = 2;
        ops.push_back(h);
      }
      return 0;
    }
  }
  int x = add_buf(Hin, Win, round_up(cin, 8));
  Op pk{}; pk.type = OP_PACK; pk.out = x; pk.bilinear = 0;
  ops.push_back(pk);
  int C = cin;
  for (int i = 0; i < d.n_kernels; ++i) {
    const std::string p = "convs." + std::to_string(i);
    const int k = d.kernel_sizes[i];
    if (!(k & 1)) { set_error("SimpleConvNet: even kernel sizes change the grid size; unsupported"); return DYF_ERR_UNSUPPORTED; }
    int li = add_conv(p + ".conv", C, dim, k, 1, (k - 1) / 2);
    attach_bn(convs[li], p + ".norm");
    attach_time(convs[li].tw, convs[li].tb, p + ".time_mlp.1", dim);
    int y = add_buf(Hin, Win, dim);
    Op o{}; o.type = OP_CONV; o.in0 = x; o.out = y; o.layer = li; o.act = ACT_GELU; o.drop_p = d.dropout; o.site = site++;
    if (d.residual && C == dim) o.res = x;  // residual only when C_in == C_out (:31, :53-54)
    ops.push_back(o);
    x = y;
    C = dim;
  }
  int hl = add_conv("head", dim, d.out_channels, 1, 1, 0);
  Op h{}; h.type = OP_CONV; h.in0 = x; h.layer = hl; h.out_mode = 2;
  ops.push_back(h);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// SST backbone (reference: src/models/unet.py:114-315; stage table SURVEY.md A.2)
// ---------------------------------------------------------------------------------------------------------------
// ResnetBlock (unet.py:79-109): block1 = WS-conv3x3 -> GroupNorm -> x*(scale+1)+shift -> SiLU -> Dropout(p1),
// block2 = WS-conv3x3 -> GroupNorm -> SiLU -> Dropout(p2), output = block2 + residual_conv(x).
int Net::resnet_block(const std::string& P, int x, int Cin, int Cout, int& site, int x2) {
  const int H = bufs[x].H, W = bufs[x].W;
  int tw = -1, tb = -1;
  attach_time(tw, tb, P + ".mlp.1", Cout);
  auto conv_gn = [&](const std::string& blk, int in, int cin, bool timed, float drop_p, int res) {
    int li = add_conv(blk + ".proj", cin, Cout, 3, 1, 1);
    convs[li].standardize = true;  // WeightStandardizedConv2d (:26-40), folded once at load time
    int raw = add_buf(H, W, Cout);
    Op c{}; c.type = OP_CONV; c.in0 = in; c.out = raw; c.layer = li; c.act = ACT_NONE;
    if (in == x) c.in1 = x2;  // block1 of a concat-fed block reads both sources
    ops.push_back(c);
    NormLayer n; n.C = Cout; n.G = d.groups;
    n.g = add_param(blk + ".norm.weight", {Cout});
    n.b = add_param(blk + ".norm.bias", {Cout});
    if (timed) { n.tw = tw; n.tb = tb; }
    norms.push_back(n);
    int y = add_buf(H, W, Cout);
    Op g{}; g.type = OP_GROUPNORM; g.in0 = raw; g.out = y; g.layer = (int)norms.size() - 1; g.act = ACT_SILU;
    g.drop_p = drop_p; g.site = site++; g.res = res;
    ops.push_back(g);
    return y;
  };
  const int h1 = conv_gn(P + ".block1", x, Cin, true, d.block_dropout1, BUF_NONE);
  int res = x;
  // parameters are registered in the reference's module order: mlp, block1, block2, residual_conv
  const int h2_conv = add_conv(P + ".block2.proj", Cout, Cout, 3, 1, 1);
  convs[h2_conv].standardize = true;
  int raw2 = add_buf(H, W, Cout);
  { Op c{}; c.type = OP_CONV; c.in0 = h1; c.out = raw2; c.layer = h2_conv; c.act = ACT_NONE; ops.push_back(c); }
  NormLayer n2; n2.C = Cout; n2.G = d.groups;
  n2.g = add_param(P + ".block2.norm.weight", {Cout});
  n2.b = add_param(P + ".block2.norm.bias", {Cout});
  norms.push_back(n2);
  const int n2i = (int)norms.size() - 1;
  if (Cin != Cout) {  // 1x1 residual projection (:96)
    int li = add_conv(P + ".residual_conv", Cin, Cout, 1, 1, 0);
    res = add_buf(H, W, Cout);
    Op c{}; c.type = OP_CONV; c.in0 = x; c.in1 = x2; c.out = res; c.layer = li; c.act = ACT_NONE;
    ops.push_back(c);
  }
  int y = add_buf(H, W, Cout);
  Op g{}; g.type = OP_GROUPNORM; g.in0 = raw2; g.out = y; g.layer = n2i; g.act = ACT_SILU; g.drop_p = d.block_dropout;
  g.site = site++; g.res = res;
  ops.push_back(g);
  return y;
}

// Residual(PreNorm(LayerNorm, LinearAttention | Attention)) (unet.py:183-191, :209; attention.py)
int Net::attention_block(const std::string& Q, int x, int C, bool linear, int& site) {
  const int H = bufs[x].H, W = bufs[x].W, heads = 4, hidden = heads * 32;
  LNLayer ln; ln.C = C;
  ln.g = add_param(Q + ".fn.norm.g", {1, C, 1, 1});
  lns.push_back(ln);
  if (linear && linattn_fused_shape_ok(C, heads) && !getenv("DYF_DISABLE_ATTN_FUSE")) {
    // the whole block (LayerNorm, dropout, qkv, both attention contractions, output projection, residual) as two
    // tensor-core kernels (attn_fused.cu); parameters are registered in the reference's module order as below
    const int qi = add_conv(Q + ".fn.fn.to_qkv.1", C, 3 * hidden, 1, 1, 0, /*bias=*/false);
    const int oi = add_conv(Q + ".fn.fn.to_out", hidden, C, 1, 1, 0);
    const int out = add_buf(H, W, C);
    Op f{}; f.type = OP_LINATTN_FUSED; f.in0 = x; f.out = out; f.layer = (int)lns.size() - 1; f.c0 = qi; f.c1 = oi;
    f.drop_p = d.attn_dropout; f.site = site++;
    const size_t fl = linattn_fused_scratch_floats(H * W);
    f.aux = add_buf(1, 1, (int)(fl * 2));  // fp32 partials per row (buffer sizes count 2-byte elements)
    ops.push_back(f);
    return out;
  }
  int y = add_buf(H, W, C);
  Op l{}; l.type = OP_CHANNEL_LN; l.in0 = x; l.out = y; l.layer = (int)lns.size() - 1;
  if (linear) { l.drop_p = d.attn_dropout; l.site = site++; }  // Dropout on the qkv input (attention.py:13)
  ops.push_back(l);
  int qi = add_conv(Q + (linear ? ".fn.fn.to_qkv.1" : ".fn.fn.to_qkv"), C, 3 * hidden, 1, 1, 0, /*bias=*/false);
  int qkv = add_buf(H, W, 3 * hidden);
  { Op c{}; c.type = OP_CONV; c.in0 = y; c.out = qkv; c.layer = qi; c.act = ACT_NONE; ops.push_back(c); }
  int a = add_buf(H, W, hidden);
  Op at{}; at.type = linear ? OP_LINATTN : OP_ATTN; at.in0 = qkv; at.out = a;
  if (linear) at.aux = add_buf(1, 1, heads * 32 * 32 * 2);  // fp32 context matrices [heads, 32, 32]
  else { at.drop_p = d.attn_dropout; at.site = site++; }   // Dropout on the probabilities (attention.py:59,70)
  ops.push_back(at);
  int oi = add_conv(Q + ".fn.fn.to_out", hidden, C, 1, 1, 0);
  int out = add_buf(H, W, C);
  { Op c{}; c.type = OP_CONV; c.in0 = a; c.out = out; c.layer = oi; c.act = ACT_NONE; c.res = x; ops.push_back(c); }
  return out;
}

int Net::build_unet_resnet() {
  const int dim = d.dim, nres = d.n_mults;
  const int cin = d.in_channels + d.cond_channels;
  if (dim % 64 || nres < 1) { set_error("Unet: dim must be a multiple of 64"); return DYF_ERR_UNSUPPORTED; }
  if (d.input_dropout > 0.f) { set_error("Unet: input_dropout > 0 is not built"); return DYF_ERR_UNSUPPORTED; }
  if (d.groups != 8) { set_error("Unet: resnet_block_groups must be 8"); return DYF_ERR_UNSUPPORTED; }
  Hin = d.height; Win = d.width;
  int site = 1;
  // 7x7 stem on the tcgen05 halo-patch kernel: horizontal taps packed into the channel axis (7 x 8 slots = 64 channels),
  // seven vertical taps left (conv_umma.cu S1K7V); otherwise the generic mma.sync pipeline
  const bool stem_umma = d.init_kernel == 7 && d.init_stride == 1 && d.init_padding == 3 && cin <= 8 &&
                         (dim == 64 || dim % 128 == 0) && !getenv("DYF_DISABLE_STEM_UMMA") && !getenv("DYF_DISABLE_UMMA");
  int xin = add_buf(Hin, Win, stem_umma ? 64 : round_up(cin, 8));
  Op pk{}; pk.type = OP_PACK; pk.out = xin; pk.bilinear = 0; pk.aux = stem_umma ? 2 : 0;
  ops.push_back(pk);
  int ci = add_conv("init_conv", cin, dim, d.init_kernel, d.init_stride, d.init_padding);
  if (stem_umma) {
    ConvLayer& c = convs[ci];
    c.stem_xim2col = 1;
    c.flops_cin = cin * 7;                  // FLOP accounting: 7 taps x (7 cin) = the reference's 49 cin per pixel
    c.Cpad = 64; c.KH = 7; c.KW = 1; c.K = 7 * 64; c.Kpad = 7 * 64;
    c.wq_off = wq_elems; wq_elems += (size_t)dim * c.Kpad;
    c.wu_off = (long long)wu_elems; wu_elems += (size_t)umma_padded_cout(dim) * 64 * 7;
  }
  const int H0 = (Hin + 2 * d.init_padding - d.init_kernel) / d.init_stride + 1;
  const int W0 = (Win + 2 * d.init_padding - d.init_kernel) / d.init_stride + 1;
  int x = add_buf(H0, W0, dim);
  { Op o{}; o.type = OP_CONV; o.in0 = xin; o.out = x; o.layer = ci; o.act = ACT_NONE; ops.push_back(o); }
  const int r = x;  // `r = x.clone()` (:276)
  std::vector<int> dims(nres + 1);
  dims[0] = dim;
  for (int i = 0; i < nres; ++i) dims[i + 1] = dim * d.dim_mults[i];
  std::vector<int> hs;
  for (int l = 0; l < nres; ++l) {
    const int din = dims[l], dout = dims[l + 1];
    const std::string P = "downs." + std::to_string(l);
    x = resnet_block(P + ".0", x, din, din, site); hs.push_back(x);
    x = resnet_block(P + ".1", x, din, din, site);
    x = attention_block(P + ".2", x, din, true, site); hs.push_back(x);
    const bool down = l < nres - 1 && !d.keep_spatial_dims;
    int li = down ? add_conv(P + ".3", din, dout, 4, 2, 1) : add_conv(P + ".3", din, dout, 3, 1, 1);
    int y = down ? add_buf((bufs[x].H + 2 - 4) / 2 + 1, (bufs[x].W + 2 - 4) / 2 + 1, dout) : add_buf(bufs[x].H, bufs[x].W, dout);
    Op o{}; o.type = OP_CONV; o.in0 = x; o.out = y; o.layer = li; o.act = ACT_NONE;
    ops.push_back(o);
    x = y;
  }
  const int mid = dims[nres];
  x = resnet_block("mid_block1", x, mid, mid, site);
  x = attention_block("mid_attn", x, mid, false, site);
  x = resnet_block("mid_block2", x, mid, mid, site);
  // cat(a, b) feeding a ResnetBlock needs no buffer when both of its readers (block1's 3x3 conv and the 1x1 residual
  // conv) run on the tcgen05 TMA path, which reads the K axis from two tensor maps (DYF_DISABLE_CATFUSE=1: always copy)
  auto cat_free = [&](int a, int b, int Cout) {
    const int Ca = bufs[a].C, Cb = bufs[b].C;
    return bufs[a].H == bufs[b].H && bufs[a].W == bufs[b].W && Ca % 64 == 0 && Cb % 64 == 0 && Ca + Cb != Cout &&
           conv_umma_shape_ok(Ca + Cb, Cout, 3, 1, 1) && conv_umma_shape_ok(Ca + Cb, Cout, 1, 1, 0) &&
           !getenv("DYF_DISABLE_CATFUSE") && !getenv("DYF_DISABLE_UMMA");
  };
  auto concat = [&](int a, int b) {
    if (bufs[a].H != bufs[b].H || bufs[a].W != bufs[b].W) return -1;
    int y = add_buf(bufs[a].H, bufs[a].W, bufs[a].C + bufs[b].C);
    Op u{}; u.type = OP_UPSAMPLE; u.in0 = a; u.in1 = b; u.out = y; u.c0 = bufs[a].C; u.c1 = bufs[b].C; u.scale = 1; u.bilinear = 0;
    ops.push_back(u);
    return y;
  };
  for (int l = 0; l < nres; ++l) {
    const int din = dims[nres - 1 - l], dout = dims[nres - l];
    const std::string P = "ups." + std::to_string(l);
    for (int j = 0; j < 2; ++j) {
      const int skip = hs.back();
      hs.pop_back();
      if (cat_free(x, skip, dout)) {  // both convs that read cat(x, skip) take the two sources directly
        x = resnet_block(P + "." + std::to_string(j), x, dout + din, dout, site, skip);
      } else {
        int cat = concat(x, skip);
        if (cat < 0) { set_error("Unet: skip connection grid mismatch (odd spatial size; reference fails too, SURVEY.md F4)"); return DYF_ERR_UNSUPPORTED; }
        x = resnet_block(P + "." + std::to_string(j), cat, dout + din, dout, site);
      }
    }
    x = attention_block(P + ".2", x, dout, true, site);
    const bool up = l < nres - 1 && !d.keep_spatial_dims;
    int in = x;
    // nn.Upsample(scale_factor=2, mode="nearest") + Conv3x3 (:16-19): one composite tcgen05 conv on the low-resolution grid
    // (conv_up.cu with nearest-neighbour composite weights: no upsampled tensor) when the grid is >= 16 x 16
    const bool fuse_up = up && conv_up_shape_ok(dout, 0, din, bufs[x].H, bufs[x].W) && !getenv("DYF_DISABLE_UPFUSE") &&
                         !getenv("DYF_DISABLE_UMMA");
    if (up && !fuse_up) {
      in = add_buf(bufs[x].H * 2, bufs[x].W * 2, dout);
      Op u{}; u.type = OP_UPSAMPLE; u.in0 = x; u.in1 = BUF_NONE; u.out = in; u.c0 = dout; u.c1 = 0; u.scale = 2; u.bilinear = 0;
      ops.push_back(u);
    }
    int li = add_conv(up ? P + ".3.1" : P + ".3", dout, din, 3, 1, 1);
    int y = add_buf(bufs[in].H * (fuse_up ? 2 : 1), bufs[in].W * (fuse_up ? 2 : 1), din);
    Op o{}; o.type = fuse_up ? OP_CONV_UP : OP_CONV; o.in0 = in; o.in1 = BUF_NONE; o.out = y; o.layer = li; o.act = ACT_NONE;
    if (fuse_up) {
      o.c0 = dout; o.c1 = 0;
      convs[li].up_nearest = 1;
      for (int v = 0; v < DYF_UP_VARIANTS; ++v) { convs[li].up_off[v] = (long long)wu_elems; wu_elems += conv_up_weight_elems(dout, din); }
    }
    ops.push_back(o);
    x = y;
  }
  if (cat_free(x, r, dim)) {
    x = resnet_block("final_res_block", x, 2 * dim, dim, site, r);
  } else {
    int cat = concat(x, r);
    if (cat < 0) { set_error("Unet: output grid differs from the stem grid"); return DYF_ERR_UNSUPPORTED; }
    x = resnet_block("final_res_block", cat, 2 * dim, dim, site);
  }
  int fl = add_conv("final_conv", dim, d.out_channels, 1, 1, 0);
  if (d.out_channels <= 8 && dim % 16 == 0 && !getenv("DYF_DISABLE_HEAD1X1")) {  // a few output channels: one dot product per pixel, fp32 weights
    Op f{}; f.type = OP_HEAD1X1; f.in0 = x; f.layer = fl;
    ops.push_back(f);
  } else {
    Op f{}; f.type = OP_CONV; f.in0 = x; f.layer = fl; f.out_mode = 2;
    ops.push_back(f);
  }
  if (bufs[x].H != d.height || bufs[x].W != d.width) { set_error("Unet: init_stride != 1 is not built"); return DYF_ERR_UNSUPPORTED; }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
int Net::set_param(const char* key, const void* data, const int64_t* shape, int ndim) {
  auto it = index.find(key);
  if (it == index.end()) { set_error(std::string("unexpected state-dict key: ") + key); return DYF_ERR_ARG; }
  ParamSlot& p = params[it->second];
  if (p.ignored) { p.is_set = true; return 0; }
  bool same = (int)p.shape.size() == ndim;
  for (int i = 0; same && i < ndim; ++i) same = p.shape[i] == shape[i];
  if (!same) { set_error(std::string("shape mismatch for ") + key); return DYF_ERR_ARG; }
  if (!packed) {
    DYF_CUDA_OK(cudaMalloc(&packed, (size_t)(packed_floats + extra_floats) * sizeof(float)));
    DYF_CUDA_OK(cudaMemset(packed, 0, (size_t)(packed_floats + extra_floats) * sizeof(float)));
  }
  DYF_CUDA_OK(cudaMemcpy(packed + p.off, data, p.numel * sizeof(float), cudaMemcpyDeviceToDevice));
  p.is_set = true;
  finalized = false;
  return 0;
}

int Net::finalize(cudaStream_t s) {
  for (auto& p : params)
    if (!p.is_set && !p.ignored) { set_error("missing state-dict key: " + p.key); return DYF_ERR_STATE; }
  if (!wq) DYF_CUDA_OK(cudaMalloc(&wq, wq_elems * sizeof(act_t)));
  if (!wq_umma && wu_elems) DYF_CUDA_OK(cudaMalloc(&wq_umma, wu_elems * sizeof(act_t)));
  if (!ro_xmap.empty() && !ro_tab_dev) {  // readout column tables: [virtual -> source | source -> virtual]
    std::vector<int> tab(ro_xmap);
    tab.insert(tab.end(), ro_xinv.begin(), ro_xinv.end());
    DYF_CUDA_OK(cudaMalloc(&ro_tab_dev, tab.size() * sizeof(int)));
    DYF_CUDA_OK(cudaMemcpy(ro_tab_dev, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  for (auto& c : convs) {
    if (c.up_off[0] >= 0) {  // fused upsample + conv: composite weight variants (conv_up.cu)
      float* scratch = nullptr;
      DYF_CUDA_OK(cudaMalloc(&scratch, conv_up_weight_elems(c.Cin, c.Cout) * sizeof(float)));
      act_t* variants[DYF_UP_VARIANTS];
      for (int v = 0; v < DYF_UP_VARIANTS; ++v) variants[v] = wq_umma + c.up_off[v];
      int ru = launch_compose_up(packed + params[c.w].off, c.Cout, c.Cin, variants, scratch, s, c.up_nearest);
      if (ru) return ru;
      DYF_CUDA_OK(cudaStreamSynchronize(s));
      DYF_CUDA_OK(cudaFree(scratch));
    }
    if (c.flat_off >= 0) {  // flat-raster stage tiles (conv_flat.cu)
      int rf = c.flat_first ? launch_repack_flat_first(packed + params[c.w].off, wq_umma + c.flat_off, c.Cin, c.KH, c.flat_cp, c.flat_cin, s)
                            : launch_repack_flat(packed + params[c.w].off, wq_umma + c.flat_off, c.Cin, c.KH, s);
      if (rf) return rf;
      for (int r = 1; r < flat_weight_replicas(); ++r)
        DYF_CUDA_OK(cudaMemcpyAsync(wq_umma + c.flat_off + (size_t)r * c.flat_elems, wq_umma + c.flat_off,
                                    (size_t)c.flat_elems * sizeof(act_t), cudaMemcpyDeviceToDevice, s));
    }
    if (c.convt_z) {  // ConvTranspose2d weight re-laid out as a 1x1 conv weight [16 * Cout_t, Cin] (fp32 staging)
      float* wz = nullptr;
      DYF_CUDA_OK(cudaMalloc(&wz, (size_t)c.Cout * c.Cin * sizeof(float)));
      int rz = launch_convt_to_conv1x1(packed + params[c.w].off, wz, c.Cin, c.Cout / 16, s);
      if (!rz) rz = launch_repack_conv(wz, wq + c.wq_off, c.Cout, c.Cin, 1, 1, c.Cpad, c.Kpad, 0, s);
      if (!rz && c.wu_off >= 0) rz = launch_repack_umma(wz, wq_umma + c.wu_off, c.Cout, c.Cin, 1, 1, 0, 0, s);
      if (!rz) rz = launch_fold_norm(nullptr, nullptr, nullptr, nullptr, nullptr, 1e-5f, packed + c.na_off, packed + c.nb_off, c.Cout, s);
      if (rz) return rz;
      DYF_CUDA_OK(cudaStreamSynchronize(s));
      DYF_CUDA_OK(cudaFree(wz));
      continue;
    }
    const float* w_src = packed + params[c.w].off;
    float* composed = nullptr;
    if (c.stem_xim2col) {  // [O][cin][7][7] -> [O][64][7]: filter of the conv over the x-im2col'd input
      float* ws = nullptr;
      DYF_CUDA_OK(cudaMalloc(&ws, (size_t)c.Cout * 64 * 7 * sizeof(float)));
      int rs = launch_stem_xim2col_weight(w_src, ws, c.Cout, c.Cin, s);
      if (!rs) rs = launch_repack_umma_k7v(ws, wq_umma + c.wu_off, c.Cout, s);
      if (!rs) rs = launch_repack_conv(ws, wq + c.wq_off, c.Cout, 64, 7, 1, 64, c.Kpad, 0, s);
      const float* bias = c.b >= 0 ? packed + params[c.b].off : nullptr;
      if (!rs) rs = launch_fold_norm(bias, nullptr, nullptr, nullptr, nullptr, 1e-5f, packed + c.na_off, packed + c.nb_off, c.Cout, s);
      if (rs) return rs;
      DYF_CUDA_OK(cudaStreamSynchronize(s));
      DYF_CUDA_OK(cudaFree(ws));
      continue;
    }
    if (c.comp_s2d) {  // composite weights in plain conv layout [Cout, 64, 3, 3] (fp32 staging, freed below)
      DYF_CUDA_OK(cudaMalloc(&composed, (size_t)c.Cout * 576 * sizeof(float)));
      int rcc = launch_compose_s2d(packed + params[c.w].off, packed + params[c.comp_wi].off, packed + params[c.comp_bi].off,
                                   composed, c.Cout, c.comp_cm, params[c.comp_wi].shape[1], s);
      if (rcc) return rcc;
      w_src = composed;
    }
    if (c.wu_off >= 0 && c.KH == 2 && c.stride == 2 && c.pad == 0) {  // runs as a 1x1 conv over the space-to-depth view
      float* w1 = nullptr;
      DYF_CUDA_OK(cudaMalloc(&w1, (size_t)c.Cout * 4 * c.Cin * sizeof(float)));
      int rk = launch_k2s2_to_conv1x1(w_src, w1, c.Cout, c.Cin, s);
      if (!rk) rk = launch_repack_umma(w1, wq_umma + c.wu_off, c.Cout, 4 * c.Cin, 1, 1, 0, 0, s);
      if (rk) return rk;
      DYF_CUDA_OK(cudaStreamSynchronize(s));
      DYF_CUDA_OK(cudaFree(w1));
    } else if (c.wu_off >= 0) {
      int rcu = launch_repack_umma(w_src, wq_umma + c.wu_off, c.Cout, c.Cin, c.KH, c.stride, c.pad, c.standardize ? 1 : 0, s);
      if (rcu) return rcu;
    }
    int rc = (c.comp_wi >= 0 && !c.comp_s2d)
                 ? launch_compose_conv(packed + params[c.w].off, packed + params[c.comp_wi].off,
                                       packed + params[c.comp_bi].off, wq + c.wq_off, c.Cout, c.comp_cm, c.Cin - 1, c.KH,
                                       c.KW, c.Cpad, c.Kpad, s)
                 : launch_repack_conv(w_src, wq + c.wq_off, c.Cout, c.Cin, c.KH, c.KW, c.Cpad, c.Kpad,
                                      c.standardize ? 1 : 0, s);
    if (rc) return rc;
    if (composed) {
      DYF_CUDA_OK(cudaStreamSynchronize(s));
      DYF_CUDA_OK(cudaFree(composed));
    }
    const float* bias = c.b >= 0 ? packed + params[c.b].off : nullptr;
    const bool bn = c.bn_g >= 0;
    rc = launch_fold_norm(bias, bn ? packed + params[c.bn_g].off : nullptr, bn ? packed + params[c.bn_b].off : nullptr,
                          bn ? packed + params[c.bn_m].off : nullptr, bn ? packed + params[c.bn_v].off : nullptr, 1e-5f,
                          packed + c.na_off, packed + c.nb_off, c.Cout, s);
    if (rc) return rc;
  }
  if (!d_time_layers && !time_layers.empty())
    DYF_CUDA_OK(cudaMalloc(&d_time_layers, time_layers.size() * sizeof(TimeLayer)));
  if (!time_layers.empty())
    DYF_CUDA_OK(cudaMemcpyAsync(d_time_layers, time_layers.data(), time_layers.size() * sizeof(TimeLayer),
                                cudaMemcpyHostToDevice, s));
  DYF_CUDA_OK(cudaStreamSynchronize(s));
  clear_tab_cache();  // tables depend on the (re-)loaded parameters
  finalized = true;
  return 0;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Rasters of the flat path: net-owned, laid out for (G rows per logical call, `calls` logical calls) and zeroed whenever that
// geometry changes -- kernels only ever write valid pixels, so the zero gaps (= every layer's padding) persist between calls.
int Net::ensure_flat(int G, int calls, cudaStream_t s) {
  if (flats.empty() || (G == flat_G && calls <= flat_calls)) return 0;
  if (stream_is_capturing(s)) { set_error("internal: raster allocation during graph capture"); return DYF_ERR_STATE; }
  const int cap = G == flat_G ? std::max(calls, flat_calls) : calls;
  std::vector<size_t> offs(flats.size());
  size_t total = 0;
  for (size_t i = 0; i < flats.size(); ++i) {
    offs[i] = total;
    const FlatGeo g = flat_geo(d.height, d.width, flats[i].k, G, flats[i].hgap);
    total += ((size_t)cap * g.PC * flats[i].C * sizeof(act_t) + 1023) & ~(size_t)1023;
  }
  if (total > flat_bytes) {
    DYF_CUDA_OK(cudaStreamSynchronize(s));  // earlier launches may still read the old rasters
    if (flat_mem) DYF_CUDA_OK(cudaFree(flat_mem));
    flat_mem = nullptr; flat_bytes = 0;
    DYF_CUDA_OK(cudaMalloc(&flat_mem, total));
    flat_bytes = total;
  }
  ++generation;  // captured graphs assume the previous layout (and allocation) of the rasters
  DYF_CUDA_OK(cudaMemsetAsync(flat_mem, 0, flat_bytes, s));
  flat_offs = offs;
  flat_G = G; flat_calls = cap;
  return 0;
}

size_t Net::workspace_bytes(int rows) const {
  size_t total = 0;
  total += 2 * align256((size_t)tab_floats_per_row * rows * sizeof(float));
  total += align256((size_t)stats_floats_per_row * rows * sizeof(float));
  total += align256((size_t)time_dim * rows * sizeof(float));
  for (auto& b : bufs) total += align256(b.row_bytes() * rows);
  return total + 256;
}

int Net::forward(int rows, const float* const* srcs, const int* src_ch, int nsrc, const float* time, float* y,
                 const RngCtx& rng_in, void* ws, size_t ws_bytes, cudaStream_t s, int noise_src, float noise_w,
                 int src_rows, int group_rows, const float* host_times) {
  // `group_rows` consecutive rows share one time value (and hence one set of epilogue tables): `time` then holds
  // rows / group_rows entries (the sampler's logical calls); 1 = one time per row (the public forward)
  if (group_rows < 1 || rows % group_rows) { set_error("internal: bad group_rows"); return DYF_ERR_ARG; }
  const int tab_rows = rows / group_rows;
  if (!finalized) { set_error("net not finalized (call dyf_net_finalize after loading parameters)"); return DYF_ERR_STATE; }
  NvtxRange nvtx("dyf.net.forward", d.arch, rows);
  if (rows <= 0) { set_error("rows must be positive"); return DYF_ERR_ARG; }
  if (ws_bytes < workspace_bytes(rows)) { set_error("workspace too small"); return DYF_ERR_ARG; }
  if (d.with_time_emb && !time && !host_times) { set_error("time is required (with_time_emb=True)"); return DYF_ERR_ARG; }
  int ctot = 0;
  for (int i = 0; i < nsrc; ++i) ctot += src_ch[i];
  if (ctot != d.in_channels + d.cond_channels || nsrc > 6) {
    set_error("input/condition channels do not match num_input_channels + num_conditional_channels");
    return DYF_ERR_ARG;
  }
  RngCtx rng = rng_in;
  if (rng.group_rows == 0 || rng.group_rows > (uint32_t)rows) rng.group_rows = (uint32_t)rows;  // one logical call

  // ---- carve the workspace
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* tabA = reinterpret_cast<float*>(base);
  base += align256((size_t)tab_floats_per_row * rows * sizeof(float));
  float* tabB = reinterpret_cast<float*>(base);
  base += align256((size_t)tab_floats_per_row * rows * sizeof(float));
  float* stats = reinterpret_cast<float*>(base);
  base += align256((size_t)stats_floats_per_row * rows * sizeof(float));
  float* temb = reinterpret_cast<float*>(base);
  base += align256((size_t)time_dim * rows * sizeof(float));
  std::vector<act_t*> bp(bufs.size());
  for (size_t i = 0; i < bufs.size(); ++i) {
    bp[i] = reinterpret_cast<act_t*>(base);
    base += align256(bufs[i].row_bytes() * rows);
  }

  // ---- epilogue tables from the time embedding (a5).  With a host copy of the times the tables of a time tuple are
  // computed once and kept (they depend only on the parameters and the times, not on the rows).
  {
    const size_t tab_bytes = align256((size_t)tab_floats_per_row * tab_rows * sizeof(float));
    bool compute = true;
    if (host_times && d.with_time_emb && tab_cache.size() < 4096) {
      std::vector<float> key(host_times, host_times + tab_rows);
      auto it = tab_cache.find(key);
      compute = it == tab_cache.end();
      if (compute && stream_is_capturing(s)) {  // would allocate and copy from host memory inside a capture
        set_error("internal: epilogue tables missing during graph capture");
        return DYF_ERR_STATE;
      }
      if (compute) {
        float* buf = nullptr;
        const size_t scratch = align256((size_t)(time_dim + 1) * tab_rows * sizeof(float));
        DYF_CUDA_OK(cudaMalloc(&buf, 2 * tab_bytes + scratch));
        it = tab_cache.emplace(key, buf).first;
      }
      tabA = it->second;
      tabB = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(it->second) + tab_bytes);
      if (compute) {
        temb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(it->second) + 2 * tab_bytes);
        float* tdev = temb + (size_t)time_dim * tab_rows;
        DYF_CUDA_OK(cudaMemcpyAsync(tdev, host_times, (size_t)tab_rows * sizeof(float), cudaMemcpyHostToDevice, s));
        time = tdev;
      }
    } else if (d.with_time_emb && !time) {
      set_error("time is required (with_time_emb=True)");
      return DYF_ERR_ARG;
    }
    if (compute) {
      TimeParams tp{};
      tp.time = d.with_time_emb ? time : nullptr;
      tp.packed = packed;
      if (d.with_time_emb) {
        tp.w1_off = params[t_w1].off; tp.b1_off = params[t_b1].off;
        tp.w2_off = params[t_w2].off; tp.b2_off = params[t_b2].off;
      }
      tp.dim = d.dim; tp.time_dim = time_dim;
      tp.layers = d_time_layers; tp.n_layers = (int)time_layers.size();
      tp.rows = tab_rows; tp.tabA = tabA; tp.tabB = tabB; tp.temb = temb; tp.total_ch = (int)tab_floats_per_row;
      int rc = launch_time_tables(tp, s);
      if (rc) return rc;
    }
  }

  static const bool flat_persist = !(getenv("DYF_FLAT_PERSIST") && getenv("DYF_FLAT_PERSIST")[0] == '0');
  // parameters of one flat-raster layer (conv_flat.cu)
  auto flat_params = [&](const Op& o) {
    const ConvLayer& c = convs[o.layer];
    const int G = group_rows, calls = rows / group_rows, k = c.KH, pad = (k - 1) / 2;
    const FlatGeo gi = flat_geo(d.height, d.width, k, G, flats[o.in0].hgap);
    FlatConvParams p{};
    p.in = reinterpret_cast<act_t*>(reinterpret_cast<uint8_t*>(flat_mem) + flat_offs[o.in0]);
    p.w = wq_umma + c.flat_off; p.wrep = flat_weight_replicas(); p.wrep_stride = c.flat_elems;
    p.H = d.height; p.W = d.width; p.G = G; p.calls = calls;
    p.S_in = gi.S; p.PI_in = gi.PI; p.PC_in = gi.PC;
    if (o.aux) {  // last layer: plain NHWC [rows][H][W][64]
      p.out = bp[o.out]; p.S_out = d.width; p.PI_out = d.height * d.width; p.PC_out = G * p.PI_out;
      if (o.aux == 2) {  // ... or straight to the network output through the fused 1x1 head
        const ConvLayer& hc = convs[o.c0];
        p.head_w = packed + params[hc.w].off; p.head_b = packed + params[hc.b].off; p.head_out = y; p.head_oc = d.out_channels;
      }
    } else {
      const FlatGeo go = flat_geo(d.height, d.width, flats[o.out].k, G, flats[o.out].hgap);
      p.out = reinterpret_cast<act_t*>(reinterpret_cast<uint8_t*>(flat_mem) + flat_offs[o.out]);
      p.S_out = go.S; p.PI_out = go.PI; p.PC_out = go.PC;
    }
    p.res = o.res >= 0 ? p.in : nullptr;
    p.tabA = tabA + (size_t)time_layers[c.table].tab_off * tab_rows;
    p.tabB = tabB + (size_t)time_layers[c.table].tab_off * tab_rows;
    p.Cin = c.flat_cin;
    if (c.flat_first) {  // horizontal taps live in the channel axis: k vertical taps
      p.ntaps = k; p.halo = pad * gi.S;
      for (int ky = 0; ky < k; ++ky) p.shift[ky] = (ky - pad) * gi.S;
    } else {
      p.ntaps = k * k; p.halo = pad * gi.S + pad;
      for (int t = 0; t < k * k; ++t) p.shift[t] = (t / k - pad) * gi.S + (t % k - pad);
    }
    p.act = o.act; p.flops_k = (double)c.Cin * k * k;
    p.drop = make_drop(rng, (uint32_t)o.site, o.drop_p);
    return p;
  };
  for (size_t oi = 0; oi < ops.size(); ++oi) {
    const Op& o = ops[oi];
    int rc = 0;
    switch (o.type) {
      case OP_STEM: {
        const ConvLayer& c = convs[o.layer];
        StemParams p{};
        for (int i = 0; i < nsrc; ++i) { p.pk.src[i] = srcs[i]; p.pk.C[i] = src_ch[i]; }
        p.pk.nsrc = nsrc; p.pk.src_rows = src_rows > 0 ? src_rows : rows; p.pk.rows = rows;
        p.pk.Hi = d.height; p.pk.Wi = d.width; p.pk.Ho = bufs[o.out].H; p.pk.Wo = bufs[o.out].W;
        p.pk.bilinear = o.bilinear; p.pk.noise_src = -1;
        p.w = packed + params[c.w].off; p.bias = packed + params[c.b].off; p.out = bp[o.out];
        p.Cin = c.Cin; p.Cout = c.Cout;
        p.drop = make_drop(rng, (uint32_t)o.site, o.drop_p);
        if (noise_src >= 0) { set_error("data+noise conditioning is not supported by the fused stem"); return DYF_ERR_UNSUPPORTED; }
        rc = launch_stem(p, s);
        break;
      }
      case OP_PACK: {
        PackParams p{};
        for (int i = 0; i < nsrc; ++i) { p.src[i] = srcs[i]; p.C[i] = src_ch[i]; }
        p.nsrc = nsrc; p.src_rows = src_rows > 0 ? src_rows : rows; p.rows = rows; p.Hi = d.height; p.Wi = d.width;
        p.Ho = bufs[o.out].H; p.Wo = bufs[o.out].W; p.Cpad = bufs[o.out].C;
        p.bilinear = o.bilinear; p.out = bp[o.out];
        p.noise_src = noise_src; p.noise_w = noise_w; p.seed = rng.seed; p.stream = rng.stream; p.seed_ptr = rng.seed_ptr;
        p.rng_rows = rng.group_rows; p.row_off = rng.row_off; p.ones_channel = o.ones_channel;
        p.s2d = o.aux;
        if (o.aux == 2) p.Cpad = 64;
        if (noise_src >= 0 && o.bilinear) { set_error("data+noise conditioning with an outer resize is unsupported"); return DYF_ERR_UNSUPPORTED; }
        rc = launch_pack(p, s);
        break;
      }
      case OP_CONV: {
        const ConvLayer& c = convs[o.layer];
        const Buf& bi = bufs[o.in0];
        ConvParams p{};
        p.in = bp[o.in0]; p.w = wq + c.wq_off;
        p.w_umma = (c.wu_off >= 0 && !getenv("DYF_DISABLE_UMMA")) ? wq_umma + c.wu_off : nullptr;
        p.rows = rows; p.Hi = bi.H; p.Wi = bi.W; p.Cin = bi.C; p.Cin_real = c.flops_cin ? c.flops_cin : c.Cin;
        p.Ho = (bi.H + 2 * c.pad - c.KH) / c.stride + 1;
        p.Wo = (bi.W + 2 * c.pad - c.KW) / c.stride + 1;
        if (c.stem_xim2col) p.Wo = bi.W;  // vertical taps only: no horizontal padding
        if (c.xmap_cols) { p.in_xmap = ro_tab_dev; p.Wo = c.xmap_cols; }  // 1x1 over a column subset (NS readout)
        p.Cout = c.Cout; p.KH = c.KH; p.KW = c.KW; p.stride = c.stride; p.pad = c.pad; p.K = c.K; p.Kpad = c.Kpad;
        p.tabA = tabA + (size_t)time_layers[c.table].tab_off * tab_rows;
        p.tabB = tabB + (size_t)time_layers[c.table].tab_off * tab_rows;
        p.tab_div = group_rows;
        p.act = o.act; p.M = (long long)rows * p.Ho * p.Wo;
        p.drop = make_drop(rng, (uint32_t)o.site, o.drop_p);
        if (o.res >= 0) { p.res = bp[o.res]; p.res_ld = bufs[o.res].C; }
        if (o.out_mode == 2) { p.out = y; p.out_fp32 = 2; p.out_ld = c.Cout; }
        else { p.out = bp[o.out]; p.out_ld = bufs[o.out].C; p.out_coff = o.out_coff; }
        if (o.in1 >= 0) {  // channel concat [in0 | in1] read in place
          p.in2 = bp[o.in1]; p.Cin0 = bi.C; p.Cin = bi.C + bufs[o.in1].C;
          if (!c.flops_cin) p.Cin_real = p.Cin;
        }
        if (p.Cin != c.Cpad) { set_error("internal: conv input channel mismatch"); return DYF_ERR_STATE; }
        rc = launch_conv_umma(p, s);
        if (rc == 0 && c.xmap_cols) { set_error("internal: the column-subset readout needs the tcgen05 gather path"); return DYF_ERR_STATE; }
        if (rc == 0 && c.stem_xim2col) { set_error("internal: the x-im2col stem needs the tcgen05 TMA path"); return DYF_ERR_STATE; }
        if (rc == 0 && p.in2) { set_error("internal: two-source conv needs the tcgen05 TMA path"); return DYF_ERR_STATE; }
        if (rc == 0) rc = launch_conv_mma(p, s);
        else if (rc > 0) rc = 0;
        break;
      }
      case OP_CONV_UP: {
        const ConvLayer& c = convs[o.layer];
        UpConvParams p{};
        p.src[0] = bp[o.in0]; p.C[0] = o.c0; p.ld[0] = bufs[o.in0].C;
        p.src[1] = o.in1 >= 0 ? bp[o.in1] : bp[o.in0]; p.C[1] = o.c1; p.ld[1] = o.in1 >= 0 ? bufs[o.in1].C : bufs[o.in0].C;
        p.rows = rows; p.H = bufs[o.in0].H; p.W = bufs[o.in0].W; p.Cout = c.Cout;
        for (int v = 0; v < DYF_UP_VARIANTS; ++v) p.w[v] = wq_umma + c.up_off[v];
        p.out = bp[o.out]; p.out_ld = bufs[o.out].C;
        p.tabA = tabA + (size_t)time_layers[c.table].tab_off * tab_rows;
        p.tabB = tabB + (size_t)time_layers[c.table].tab_off * tab_rows;
        p.tab_div = group_rows; p.act = o.act;
        p.drop = make_drop(rng, (uint32_t)o.site, o.drop_p);
        rc = launch_conv_up(p, s);
        break;
      }
      case OP_UPSAMPLE: {
        UpsampleParams p{};
        p.src[0] = bp[o.in0]; p.C[0] = o.c0; p.ld[0] = bufs[o.in0].C;
        p.src[1] = o.in1 >= 0 ? bp[o.in1] : bp[o.in0]; p.C[1] = o.c1; p.ld[1] = o.in1 >= 0 ? bufs[o.in1].C : 8;
        p.rows = rows; p.H = bufs[o.in0].H; p.W = bufs[o.in0].W; p.scale = o.scale; p.bilinear = o.bilinear;
        p.out = bp[o.out];
        rc = launch_upsample(p, s);
        break;
      }
      case OP_GROUPNORM: {
        const NormLayer& n = norms[o.layer];
        GroupNormParams p{};
        p.x = bp[o.in0]; p.y = bp[o.out];
        p.gamma = packed + params[n.g].off; p.beta = packed + params[n.b].off;
        if (n.table >= 0) {
          p.tabA = tabA + (size_t)time_layers[n.table].tab_off * tab_rows;
          p.tabB = tabB + (size_t)time_layers[n.table].tab_off * tab_rows;
        }
        if (o.res >= 0) { p.res = bp[o.res]; p.res_ld = bufs[o.res].C; }
        p.stats = stats + (size_t)n.stats_off * rows;
        p.rows = rows; p.HW = bufs[o.in0].H * bufs[o.in0].W; p.C = n.C; p.G = n.G; p.act = o.act; p.eps = 1e-5f;
        p.tab_div = group_rows;
        p.drop = make_drop(rng, (uint32_t)o.site, o.drop_p);
        rc = launch_groupnorm(p, s);
        break;
      }
      case OP_READOUT_GATHER: {
        ReadoutGatherParams p{};
        p.z = bp[o.in0]; p.bias = packed + params[ro_b].off; p.y = y;
        p.rows = rows; p.Hs = bufs[o.in0].H; p.Ws = ro_src_w; p.Wz = bufs[o.in0].W; p.Cout = d.out_channels; p.Ho = d.height; p.Wo = d.width;
        if (!ro_xmap.empty()) p.xinv = ro_tab_dev + ro_xmap.size();
        rc = launch_readout_gather(p, s);
        break;
      }
      case OP_READOUT: {
        ReadoutParams p{};
        p.x = bp[o.in0]; p.w = packed + params[ro_w].off; p.bias = packed + params[ro_b].off; p.y = y;
        p.rows = rows; p.Hs = bufs[o.in0].H; p.Ws = bufs[o.in0].W; p.Cin = bufs[o.in0].C; p.Cout = d.out_channels;
        p.Ho = d.height; p.Wo = d.width;
        rc = launch_readout(p, s);
        break;
      }
      case OP_CHANNEL_LN: {
        const LNLayer& l = lns[o.layer];
        ChannelLNParams p{};
        p.x = bp[o.in0]; p.y = bp[o.out]; p.g = packed + params[l.g].off;
        p.M = (long long)rows * bufs[o.in0].H * bufs[o.in0].W; p.C = l.C; p.HW = bufs[o.in0].H * bufs[o.in0].W;
        p.drop = make_drop(rng, (uint32_t)o.site, o.drop_p);
        rc = launch_channel_ln(p, s);
        break;
      }
      case OP_LINATTN:
      case OP_ATTN: {
        AttnParams p{};
        p.qkv = bp[o.in0]; p.out = bp[o.out];
        p.ctx = o.aux > 0 ? reinterpret_cast<float*>(bp[o.aux]) : nullptr;
        p.rows = rows; p.n = bufs[o.in0].H * bufs[o.in0].W; p.heads = 4;
        p.drop = make_drop(rng, (uint32_t)o.site, o.drop_p);
        rc = o.type == OP_LINATTN ? launch_linear_attention(p, s) : launch_attention(p, s);
        break;
      }
      case OP_HEAD1X1: {
        const ConvLayer& c = convs[o.layer];
        Head1x1Params p{};
        p.x = bp[o.in0]; p.w = packed + params[c.w].off; p.bias = c.b >= 0 ? packed + params[c.b].off : nullptr; p.y = y;
        p.HW = bufs[o.in0].H * bufs[o.in0].W; p.M = (long long)rows * p.HW; p.C = bufs[o.in0].C; p.OC = c.Cout;
        rc = launch_head1x1(p, s);
        break;
      }
      case OP_LINATTN_FUSED: {
        const LNLayer& l = lns[o.layer];
        const ConvLayer& cq = convs[o.c0];
        const ConvLayer& co = convs[o.c1];
        LinAttnFusedParams p{};
        p.x = bp[o.in0]; p.y = bp[o.out]; p.g = packed + params[l.g].off;
        p.w_qkv = wq + cq.wq_off; p.ldw = cq.Kpad;
        p.w_out = wq + co.wq_off; p.ldw_out = co.Kpad; p.b_out = packed + params[co.b].off;
        p.part = reinterpret_cast<float*>(bp[o.aux]);
        p.rows = rows; p.n = bufs[o.in0].H * bufs[o.in0].W; p.C = l.C;
        p.drop = make_drop(rng, (uint32_t)o.site, o.drop_p);
        rc = launch_linattn_fused(p, s);
        break;
      }
      case OP_FLAT_PACK: {
        const int G = group_rows, calls = rows / group_rows;
        rc = ensure_flat(G, calls, s);
        if (rc) break;
        const ConvLayer& c0 = convs[ops[1].layer];
        const FlatGeo g = flat_geo(d.height, d.width, flats[0].k, G, flats[0].hgap);
        FlatPackParams p{};
        const long long plane = (long long)d.height * d.width;
        for (int i = 0; i < nsrc; ++i)
          for (int cc = 0; cc < src_ch[i]; ++cc, ++p.n_slots) {
            if (p.n_slots >= 16) { set_error("internal: flat pack supports <= 16 input channels"); return DYF_ERR_STATE; }
            p.slot_ptr[p.n_slots] = srcs[i] + (size_t)cc * plane;
            p.slot_rstride[p.n_slots] = (long long)src_ch[i] * plane;
          }
        p.src_rows = src_rows > 0 ? src_rows : rows; p.rows = rows; p.H = d.height; p.W = d.width;
        p.k = flats[0].k; p.G = G; p.CP = c0.flat_cp; p.Cflat = flats[0].C; p.S = g.S; p.PI = g.PI; p.PC = g.PC;
        p.out = reinterpret_cast<act_t*>(reinterpret_cast<uint8_t*>(flat_mem) + flat_offs[0]);
        if (noise_src >= 0) { set_error("data+noise conditioning is not built for the flat-raster path"); return DYF_ERR_UNSUPPORTED; }
        rc = launch_pack_flat(p, s);
        break;
      }
      case OP_FLAT_CONV: {
        // consecutive flat layers of the network run in ONE persistent kernel (grid-wide barriers between them)
        FlatConvParams lp[FLAT_MAX_LAYERS];
        int nl = 0;
        while (nl < (flat_persist ? FLAT_MAX_LAYERS : 1) && oi + nl < ops.size() && ops[oi + nl].type == OP_FLAT_CONV) {
          lp[nl] = flat_params(ops[oi + nl]);
          ++nl;
        }
        rc = launch_conv_flat_net(lp, nl, s);
        oi += nl - 1;
        break;
      }
      default: set_error("internal: unknown op"); return DYF_ERR_STATE;
    }
    if (rc) return rc;
  }
  return 0;
}

}  // namespace dyf
