// extern "C" surface of the engine (include/dyffusion_b200.h).
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

#include "engine.hpp"

namespace dyf {
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
bool pdl_enabled(int family) {
  static const bool on = getenv("DYF_DISABLE_PDL") == nullptr;
  static const int mask = getenv("DYF_PDL_MASK") ? atoi(getenv("DYF_PDL_MASK")) : 0xB;  // (GroupNorm: see DESIGN.md)
  return on && ((mask >> family) & 1);
}
uint64_t launch_counter() { return g_launches.load(); }

// ---- NVTX ranges around network calls / sampler runs (nvtx3 is header-only and resolves the tool at run time)
static bool g_nvtx = getenv("DYF_NVTX") != nullptr && getenv("DYF_NVTX")[0] == '1';
NvtxRange::NvtxRange(const char* what, int arch, int rows) : on(g_nvtx) {
  if (!on) return;
  static const char* names[] = {"unet_simple", "unet_resnet", "convnet"};
  char buf[96];
  snprintf(buf, sizeof buf, "%s %s rows=%d", what, arch >= 0 && arch < 3 ? names[arch] : "?", rows);
  nvtxRangePushA(buf);
}
NvtxRange::~NvtxRange() { if (on) nvtxRangePop(); }

// ---- launch profiler
struct ProfRec { cudaEvent_t a, b; int klass; double flops, bytes; };
static bool g_prof_on = false;
static int g_prof_only = -1;  // -1 = every kernel class, else only launches of this class are bracketed by events
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
bool profiling_enabled() { return g_prof_on; }
static cudaEvent_t take_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
ProfScope::ProfScope(cudaStream_t stream, int klass, double flops, double bytes) : s(stream), idx(-1) {
  if (!g_prof_on || (g_prof_only >= 0 && klass != g_prof_only)) return;
  ProfRec r{take_event(), take_event(), klass, flops, bytes};
  cudaEventRecord(r.a, s);
  idx = (int)g_prof.size();
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_prof[idx].b, s);
}
}  // namespace dyf

using namespace dyf;

static int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device: dyffusion_b200 has no CPU fallback");
    return DYF_ERR_CUDA;
  }
  return 0;
}

extern "C" {

int dyf_abi_version(void) { return DYF_ABI_VERSION; }
const char* dyf_act_dtype(void) { return DYF_ACT_NAME; }
int dyf_nvtx_enable(int32_t on) { g_nvtx = on != 0; return 0; }
const char* dyf_last_error(void) { return get_error(); }
uint64_t dyf_launch_count(void) { return g_launches.load(); }

int dyf_profile_filter(int32_t klass) {
  if (klass >= KC_COUNT) { set_error("bad kernel class"); return DYF_ERR_ARG; }
  g_prof_only = klass < 0 ? -1 : klass;
  return 0;
}

int dyf_profile_enable(int32_t on) {
  if (!on) {
    for (auto& r : g_prof) { g_event_pool.push_back(r.a); g_event_pool.push_back(r.b); }
    g_prof.clear();
  }
  g_prof_on = on != 0;
  return 0;
}

int dyf_profile_read(double* ms, double* flops, double* bytes, uint64_t* launches, int32_t n_classes) {
  if (!ms || !flops || !bytes || !launches || n_classes < KC_COUNT) { set_error("bad argument"); return DYF_ERR_ARG; }
  for (int i = 0; i < n_classes; ++i) { ms[i] = flops[i] = bytes[i] = 0.0; launches[i] = 0; }
  for (auto& r : g_prof) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) { set_error("profile: event sync failed"); return DYF_ERR_CUDA; }
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) { set_error("profile: elapsed failed"); return DYF_ERR_CUDA; }
    ms[r.klass] += t; flops[r.klass] += r.flops; bytes[r.klass] += r.bytes; launches[r.klass] += 1;
    g_event_pool.push_back(r.a); g_event_pool.push_back(r.b);
  }
  g_prof.clear();
  return 0;
}

int dyf_net_create(const dyf_net_desc* desc, dyf_net** out) {
  if (!desc || !out) { set_error("null argument"); return DYF_ERR_ARG; }
  Net* n = new (std::nothrow) Net();
  if (!n) { set_error("out of host memory"); return DYF_ERR_STATE; }
  n->d = *desc;
  int rc = n->build();
  if (rc) { delete n; return rc; }
  *out = reinterpret_cast<dyf_net*>(n);
  return 0;
}

void dyf_net_destroy(dyf_net* net) { delete reinterpret_cast<Net*>(net); }

int dyf_net_set_param(dyf_net* net, const char* key, const void* data, const int64_t* shape, int32_t ndim) {
  if (!net || !key || !data) { set_error("null argument"); return DYF_ERR_ARG; }
  if (int rc = require_device()) return rc;
  return reinterpret_cast<Net*>(net)->set_param(key, data, shape, ndim);
}

int dyf_net_finalize(dyf_net* net, void* stream) {
  if (!net) { set_error("null argument"); return DYF_ERR_ARG; }
  if (int rc = require_device()) return rc;
  return reinterpret_cast<Net*>(net)->finalize(reinterpret_cast<cudaStream_t>(stream));
}

int dyf_net_num_params(const dyf_net* net) { return net ? (int)reinterpret_cast<const Net*>(net)->params.size() : 0; }
const char* dyf_net_param_key(const dyf_net* net, int32_t i) {
  const Net* n = reinterpret_cast<const Net*>(net);
  if (!n || i < 0 || i >= (int)n->params.size()) return nullptr;
  return n->params[i].key.c_str();
}

int dyf_net_param_shape(const dyf_net* net, int32_t i, int64_t* shape, int32_t* is_buffer) {
  const Net* n = reinterpret_cast<const Net*>(net);
  if (!n || !shape || i < 0 || i >= (int)n->params.size()) { set_error("bad argument"); return DYF_ERR_ARG; }
  const ParamSlot& p = n->params[i];
  for (size_t k = 0; k < p.shape.size() && k < 4; ++k) shape[k] = p.shape[k];
  if (is_buffer) {
    const std::string& key = p.key;
    auto ends = [&](const char* suf) { std::string s(suf); return key.size() >= s.size() && key.compare(key.size() - s.size(), s.size(), s) == 0; };
    *is_buffer = (ends(".running_mean") || ends(".running_var") || ends(".num_batches_tracked")) ? 1 : 0;
  }
  return (int)p.shape.size();
}

int dyf_net_workspace_bytes(const dyf_net* net, int32_t rows, size_t* bytes) {
  if (!net || !bytes || rows <= 0) { set_error("bad argument"); return DYF_ERR_ARG; }
  *bytes = reinterpret_cast<const Net*>(net)->workspace_bytes(rows);
  return 0;
}

int dyf_net_forward_srcs(dyf_net* net, int32_t rows, const float* const* srcs, const int32_t* src_channels,
                         int32_t nsrc, const float* time, float* y, const dyf_dropout* drop, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!net || !srcs || !src_channels || !y || !workspace) { set_error("null argument"); return DYF_ERR_ARG; }
  if (int rc = require_device()) return rc;
  RngCtx rng;
  if (drop) {
    if (drop->row_offset + (uint64_t)rows > 0xFFFFFFFFull) { set_error("row_offset out of range"); return DYF_ERR_ARG; }
    rng.on = drop->mode == 1; rng.seed = drop->seed; rng.stream = drop->stream; rng.row_off = (uint32_t)drop->row_offset;
  }
  rng.group_rows = (uint32_t)rows;  // one logical call
  return reinterpret_cast<Net*>(net)->forward(rows, srcs, src_channels, nsrc, time, y, rng, workspace, workspace_bytes,
                                              reinterpret_cast<cudaStream_t>(stream));
}

int dyf_net_forward(dyf_net* net, int32_t rows, const float* x, const float* cond, const float* time, float* y,
                    const dyf_dropout* drop, void* workspace, size_t workspace_bytes, void* stream) {
  if (!net || !x) { set_error("null argument"); return DYF_ERR_ARG; }
  Net* n = reinterpret_cast<Net*>(net);
  if ((n->d.cond_channels > 0) != (cond != nullptr)) {
    set_error(cond ? "condition is not None but num_conditional_channels is 0"
                   : "condition is required (num_conditional_channels > 0)");
    return DYF_ERR_ARG;
  }
  const float* srcs[2];
  int32_t ch[2];
  int ns = 0;
  const bool cond_first = n->d.arch == DYF_ARCH_UNET_RESNET;  // unet.py:269 vs unet_simple.py:184 / simple_conv_net.py:121
  if (cond && cond_first) { srcs[ns] = cond; ch[ns++] = n->d.cond_channels; }
  srcs[ns] = x; ch[ns++] = n->d.in_channels;
  if (cond && !cond_first) { srcs[ns] = cond; ch[ns++] = n->d.cond_channels; }
  return dyf_net_forward_srcs(net, rows, srcs, ch, ns, time, y, drop, workspace, workspace_bytes, stream);
}

int dyf_sampler_create(dyf_net* forecaster, dyf_net* interpolator, const dyf_sampler_desc* desc, dyf_sampler** out) {
  if (!forecaster || !interpolator || !desc || !out) { set_error("null argument"); return DYF_ERR_ARG; }
  if (desc->n_schedule <= 0 || !desc->schedule || !desc->tau || !desc->time_forecaster) {
    set_error("sampler: schedule arrays are required");
    return DYF_ERR_ARG;
  }
  Sampler* s = new (std::nothrow) Sampler();
  if (!s) { set_error("out of host memory"); return DYF_ERR_STATE; }
  s->F = reinterpret_cast<Net*>(forecaster);
  s->I = reinterpret_cast<Net*>(interpolator);
  s->d = *desc;
  s->schedule.assign(desc->schedule, desc->schedule + desc->n_schedule);
  s->tau.assign(desc->tau, desc->tau + desc->n_schedule);
  s->tF.assign(desc->time_forecaster, desc->time_forecaster + desc->n_schedule);
  if (desc->n_refine > 0) s->refine.assign(desc->refine_times, desc->refine_times + desc->n_refine);
  s->d.schedule = s->d.tau = s->d.time_forecaster = s->d.refine_times = nullptr;
  // channel bookkeeping must agree with the two backbones
  const int f_cond = (desc->forward_conditioning ? desc->window_channels : 0) + desc->static_channels;
  if (s->F->d.in_channels != desc->channels || s->F->d.cond_channels != f_cond ||
      s->I->d.in_channels != desc->window_channels + desc->channels || s->I->d.cond_channels != desc->static_channels ||
      s->I->d.out_channels != desc->channels || s->F->d.out_channels != desc->channels ||
      s->F->d.height != s->I->d.height || s->F->d.width != s->I->d.width) {
    set_error("sampler: forecaster/interpolator channel configuration does not match the sampler description");
    delete s;
    return DYF_ERR_ARG;
  }
  int rc = s->plan();
  if (rc) { delete s; return rc; }
  *out = reinterpret_cast<dyf_sampler*>(s);
  return 0;
}

void dyf_sampler_destroy(dyf_sampler* s) { delete reinterpret_cast<Sampler*>(s); }

int dyf_sampler_graph_replays(const dyf_sampler* s, uint64_t* replays) {
  if (!s || !replays) { set_error("bad argument"); return DYF_ERR_ARG; }
  *replays = reinterpret_cast<const Sampler*>(s)->graph_replays;
  return DYF_OK;
}

int dyf_sampler_workspace_bytes(const dyf_sampler* s, int32_t rows, size_t* bytes) {
  if (!s || !bytes || rows <= 0) { set_error("bad argument"); return DYF_ERR_ARG; }
  *bytes = reinterpret_cast<const Sampler*>(s)->workspace_bytes(rows);
  return 0;
}

int dyf_sampler_num_outputs(const dyf_sampler* s, int32_t* n_outputs, double* keys, int32_t keys_capacity) {
  if (!s || !n_outputs) { set_error("null argument"); return DYF_ERR_ARG; }
  const Sampler* sm = reinterpret_cast<const Sampler*>(s);
  *n_outputs = (int32_t)sm->out_keys.size();
  if (keys)
    for (int i = 0; i < keys_capacity && i < (int)sm->out_keys.size(); ++i) keys[i] = sm->out_keys[i];
  return 0;
}

int dyf_sampler_run(dyf_sampler* s, int32_t rows, const float* ic, const float* static_cond, float* preds,
                    float* x0_hat_out, uint64_t seed, uint64_t row_offset, void* workspace, size_t workspace_bytes,
                    void* stream) {
  if (!s || !ic || !preds || !workspace || rows <= 0) { set_error("bad argument"); return DYF_ERR_ARG; }
  if (int rc = require_device()) return rc;
  return reinterpret_cast<Sampler*>(s)->run(rows, ic, static_cond, preds, x0_hat_out, seed, row_offset, workspace,
                                            workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int dyf_debug_dropout_mask(uint64_t seed, uint64_t stream, uint32_t site, float p, int64_t n_elems, uint8_t* mask,
                           void* stream_handle) {
  if (!mask || n_elems <= 0) { set_error("bad argument"); return DYF_ERR_ARG; }
  if (int rc = require_device()) return rc;
  RngCtx rng;
  rng.on = true; rng.seed = seed; rng.stream = stream;
  return launch_dropout_mask(make_drop(rng, site, p), n_elems, mask,
                             reinterpret_cast<cudaStream_t>(stream_handle));
}

}  // extern "C"
