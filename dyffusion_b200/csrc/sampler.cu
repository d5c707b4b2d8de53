// The DYffusion sampling loop (reference: BaseDYffusion.sample_loop, src/diffusion/dyffusion.py:335-426) driven
// natively: no Python between network calls, the two interpolator evaluations of a cold-sampling step batched into
// one launch sequence (2R rows), the refinement calls batched as well, all on the caller's stream.
#include <cstdio>
#include <algorithm>
#include <cmath>

#include "engine.hpp"

namespace dyf {

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static bool is_int(double v) { return std::floor(v) == v; }

int Sampler::plan() {
  const int N = d.num_timesteps;
  const int n = (int)schedule.size();
  if (n < 1 || schedule[0] != 0.0) { set_error("sampling schedule must start at 0"); return DYF_ERR_ARG; }
  for (int i = 1; i < n; ++i)
    if (!(schedule[i] > schedule[i - 1])) { set_error("sampling schedule must be strictly increasing"); return DYF_ERR_ARG; }
  // host replay of the key bookkeeping (:365-366, :395-399)
  std::vector<double> keys;
  std::vector<int> is_dyn(n, 0);
  std::vector<double> step_key(n, 0.0);
  double key = 0;
  const double last_plus = schedule[n - 1] + 1;
  for (int i = 0; i < n; ++i) {
    const double s = schedule[i];
    const bool last = s == N - 1;
    const double s_next = i + 1 < n ? schedule[i + 1] : last_plus;
    double t_next = INFINITY;
    if (!last) {
      if (i + 1 < n) t_next = tau[i + 1];
      else { set_error("a sampling schedule that stops before the last diffusion step is not supported natively"); return DYF_ERR_UNSUPPORTED; }
    }
    (void)s_next;
    is_dyn[i] = last || is_int(t_next);
    key = s < N - 1 ? std::floor(t_next) : key + 1;
    step_key[i] = key;
    if (is_dyn[i]) keys.push_back(key);
  }
  for (double r : refine) keys.push_back(r);
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  out_keys = keys;
  auto slot_of = [&](double k) { return (int)(std::lower_bound(out_keys.begin(), out_keys.end(), k) - out_keys.begin()); };
  step_slot.assign(n, -1);
  for (int i = 0; i < n; ++i)
    if (is_dyn[i]) step_slot[i] = slot_of(step_key[i]);
  refine_slot.clear();
  for (double r : refine) {
    if (!(r > 0 && r < d.interpolator_horizon)) { set_error("interpolate time must be in (0, horizon)"); return DYF_ERR_ARG; }
    refine_slot.push_back(slot_of(r));
  }
  for (int i = 1; i < n; ++i)  // interpolator time range check (:484-486)
    if (schedule[i] <= N - 1 && !(tau[i] > 0 && tau[i] < d.interpolator_horizon)) {
      set_error("interpolate time must be in (0, horizon)");
      return DYF_ERR_ARG;
    }
  return 0;
}

static int batch_cap(const dyf_sampler_desc& d, int rows, int cells) {
  // logical interpolator calls per launch sequence; a cold-sampling step needs two (t = s_next and t = s).  Default cap on
  // the rows of one launch sequence: 256, or -- for small grids, whose network calls are launch-latency-bound -- what keeps
  // about 2 M grid cells in flight (spring-mesh 10 x 10: 4096 rows, i.e. the refinement calls of a 200-row job run 20 at
  // a time instead of 2)
  const int by_cells = std::min(4096, (2 << 20) / std::max(1, cells));
  int cap = d.max_rows_per_call > 0 ? std::max(d.max_rows_per_call, 2 * rows) : std::max(2 * rows, std::max(256, by_cells));
  return std::max(2, cap / rows);
}

// workspace layout: [seed | staged ic | staged static | staged forecasts | staged x0_hat] (graph mode only) then the state
// of the loop: x_s, x0_hat, interpolator outputs, network workspace
static size_t stage_bytes(const Sampler& sm, int rows, size_t* off_ic, size_t* off_st, size_t* off_preds, size_t* off_x0) {
  const size_t plane = (size_t)sm.F->d.height * sm.F->d.width * sizeof(float);
  size_t o = 256;  // the seed
  if (off_ic) *off_ic = o;
  o += align256((size_t)rows * sm.d.window_channels * plane);
  if (off_st) *off_st = o;
  o += align256((size_t)rows * sm.d.static_channels * plane);
  if (off_preds) *off_preds = o;
  o += align256(sm.out_keys.size() * (size_t)rows * sm.d.channels * plane);
  if (off_x0) *off_x0 = o;
  o += align256((size_t)rows * sm.d.channels * plane);
  return o;
}

size_t Sampler::workspace_bytes(int rows) const {
  const size_t plane = (size_t)F->d.height * F->d.width;
  const size_t state = align256((size_t)rows * d.channels * plane * sizeof(float));
  const int k = batch_cap(d, rows, F->d.height * F->d.width);
  size_t total = 2 * state;                                                    // x_s, x0_hat
  total += align256((size_t)k * rows * d.channels * plane * sizeof(float));     // interpolator outputs
  total += std::max(F->workspace_bytes(rows), I->workspace_bytes(k * rows));
  total += d.cuda_graph ? stage_bytes(*this, rows, nullptr, nullptr, nullptr, nullptr) : 256;
  return total + 512;
}

Sampler::~Sampler() {
  for (auto& kv : graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
  }
  if (ev_in) cudaEventDestroy(ev_in);
  if (ev_out) cudaEventDestroy(ev_out);
  if (side_stream) cudaStreamDestroy(side_stream);
}

__global__ void store_seed_kernel(uint64_t* dst, uint64_t seed) { *dst = seed; }

int Sampler::enqueue(int rows, const float* ic, const float* stat, float* preds, float* x0_out, uint64_t seed,
                     const uint64_t* seed_dev, uint64_t row_offset, void* ws, size_t ws_bytes, cudaStream_t s) {
  const int N = d.num_timesteps, n = (int)schedule.size(), C = d.channels;
  const size_t plane = (size_t)F->d.height * F->d.width;
  const size_t state_n = (size_t)rows * C * plane;
  const int kmax = batch_cap(d, rows, F->d.height * F->d.width);

  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* x_s = reinterpret_cast<float*>(base); base += align256(state_n * sizeof(float));
  float* x0_hat = reinterpret_cast<float*>(base); base += align256(state_n * sizeof(float));
  float* ybuf = reinterpret_cast<float*>(base); base += align256((size_t)kmax * state_n * sizeof(float));
  void* net_ws = base;
  const size_t net_ws_bytes = ws_bytes - (size_t)(base - reinterpret_cast<uint8_t*>(ws));

  // x_s = initial_condition[:, -C:]  (:348)
  DYF_CUDA_OK(cudaMemcpy2DAsync(x_s, (size_t)C * plane * sizeof(float),
                                ic + (size_t)(d.window_channels - C) * plane,
                                (size_t)d.window_channels * plane * sizeof(float), (size_t)C * plane * sizeof(float),
                                rows, cudaMemcpyDeviceToDevice, s));
  uint64_t call = 0;
  const bool f_cond_first = F->d.arch == DYF_ARCH_UNET_RESNET;  // unet.py:269 concatenates the condition first
  const bool i_cond_first = I->d.arch == DYF_ARCH_UNET_RESNET;
  RngCtx rng;  // masks / noise are keyed by (seed, logical call, site, global row): see DropCfg
  rng.seed = seed; rng.seed_ptr = seed_dev; rng.group_rows = (uint32_t)rows; rng.row_off = (uint32_t)row_offset;

  auto run_F = [&](int idx) -> int {  // predict_x_last (:205-239)
    const float* srcs[4];
    int ch[4];
    int ns = 0, noise_src = -1;
    float noise_w = 0.f;
    auto push_cond = [&]() {
      if (d.forward_conditioning != 0) {
        if (d.forward_conditioning == 2) { noise_src = ns; noise_w = (float)(schedule[idx] / (double)(N - 1)); }
        srcs[ns] = ic; ch[ns++] = d.window_channels;
      }
      if (stat) { srcs[ns] = stat; ch[ns++] = d.static_channels; }
    };
    if (f_cond_first) push_cond();
    srcs[ns] = x_s; ch[ns++] = C;
    if (!f_cond_first) push_cond();
    const float th = (float)tF[idx];  // host copy of the time: the net keeps the epilogue tables of times it has seen
    RngCtx r = rng;
    r.on = d.forecaster_dropout != 0;  // experiment-level inference dropout reaches the forecaster too
    r.stream = call++;
    return F->forward(rows, srcs, ch, ns, nullptr, x0_hat, r, net_ws, net_ws_bytes, s, noise_src, noise_w, rows, rows, &th);
  };
  // k logical interpolator calls at times t[0..k) sharing the inputs (ic, x0_hat); outputs land in ybuf[j]
  auto run_I = [&](const double* t, int k, float* out = nullptr) -> int {  // q_sample (:140-163) + _interpolate (:480-494)
    const float* srcs[4];
    int ch[4];
    int ns = 0;
    if (i_cond_first && stat) { srcs[ns] = stat; ch[ns++] = d.static_channels; }
    srcs[ns] = ic; ch[ns++] = d.window_channels;
    srcs[ns] = x0_hat; ch[ns++] = C;
    if (!i_cond_first && stat) { srcs[ns] = stat; ch[ns++] = d.static_channels; }
    std::vector<float> th(t, t + k);  // host copy of the times: the net keeps the epilogue tables of tuples it has seen
    RngCtx r = rng;
    r.on = d.enable_interpolator_dropout != 0 || d.forecaster_dropout != 0;
    r.stream = call;  // logical call j of the batch draws from stream call + j
    call += k;
    return I->forward(k * rows, srcs, ch, ns, nullptr, out ? out : ybuf, r, net_ws, net_ws_bytes, s, -1, 0.f, rows, rows, th.data());
  };

  for (int i = 0; i < n; ++i) {
    const double sstep = schedule[i];
    const bool last = sstep == N - 1;
    int rc = run_F(i);
    if (rc) return rc;
    const bool has_next = i + 1 < n && schedule[i + 1] <= N - 1;
    float* out = step_slot[i] >= 0 ? preds + (size_t)step_slot[i] * state_n : nullptr;
    if (d.sampling_type == 0) {  // cold (:381-388)
      if (last && !d.use_cold_sampling_for_last_step) {
        rc = launch_cold_update(x_s, nullptr, x0_hat, out, (long long)state_n, s);
      } else {
        double t[2];
        int k = 0;
        if (has_next) t[k++] = tau[i + 1];
        const bool need_cur = sstep > 0;
        if (need_cur) t[k++] = tau[i];
        const float* x_next = x0_hat;
        const float* x_cur = nullptr;
        if (k > 0) {
          rc = run_I(t, k);
          if (rc) return rc;
          if (has_next) { x_next = ybuf; if (need_cur) x_cur = ybuf + state_n; }
          else if (need_cur) x_cur = ybuf;
        }
        if (need_cur) rc = launch_cold_update(x_s, x_cur, x_next, out, (long long)state_n, s);
        else rc = launch_cold_update(x_s, nullptr, x_next, out, (long long)state_n, s);  // s == 0: x_s <- x_next
      }
    } else {  // naive (:390-391)
      const float* x_next = x0_hat;
      if (has_next) {
        double t = tau[i + 1];
        rc = run_I(&t, 1);
        if (rc) return rc;
        x_next = ybuf;
      }
      rc = launch_cold_update(x_s, nullptr, x_next, out, (long long)state_n, s);
    }
    if (rc) return rc;
  }
  // refinement of the intermediate predictions with the final x0_hat (:408-422)
  for (size_t j0 = 0; j0 < refine.size(); j0 += kmax) {
    const int k = (int)std::min<size_t>(kmax, refine.size() - j0);
    bool consecutive = true;  // output slots of this batch back to back: the interpolator writes them in place
    for (int j = 1; j < k; ++j) consecutive = consecutive && refine_slot[j0 + j] == refine_slot[j0] + j;
    float* const dst = consecutive ? preds + (size_t)refine_slot[j0] * state_n : ybuf;
    int rc = run_I(&refine[j0], k, dst);
    if (rc) return rc;
    for (int j = 0; j < k && !consecutive; ++j)
      DYF_CUDA_OK(cudaMemcpyAsync(preds + (size_t)refine_slot[j0 + j] * state_n, ybuf + (size_t)j * state_n,
                                  state_n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  if (x0_out) DYF_CUDA_OK(cudaMemcpyAsync(x0_out, x0_hat, state_n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

int Sampler::run(int rows, const float* ic, const float* stat, float* preds, float* x0_out, uint64_t seed,
                 uint64_t row_offset, void* ws, size_t ws_bytes, cudaStream_t caller) {
  cudaStream_t s = caller;
  const bool graph_mode = d.cuda_graph && !profiling_enabled();
  if (graph_mode && (caller == nullptr || caller == cudaStreamLegacy)) {  // un-capturable stream: run on the side stream
    if (!side_stream) {
      DYF_CUDA_OK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
      DYF_CUDA_OK(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
      DYF_CUDA_OK(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming));
    }
    DYF_CUDA_OK(cudaEventRecord(ev_in, caller));
    DYF_CUDA_OK(cudaStreamWaitEvent(side_stream, ev_in, 0));
    s = side_stream;
  }
  const int rc_run = run_on(rows, ic, stat, preds, x0_out, seed, row_offset, ws, ws_bytes, s);
  if (s != caller) {  // (also after an error: whatever was enqueued on the side stream stays ordered before the caller's next work)
    DYF_CUDA_OK(cudaEventRecord(ev_out, s));
    DYF_CUDA_OK(cudaStreamWaitEvent(caller, ev_out, 0));
  }
  return rc_run;
}

int Sampler::run_on(int rows, const float* ic, const float* stat, float* preds, float* x0_out, uint64_t seed,
                    uint64_t row_offset, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < workspace_bytes(rows)) { set_error("sampler workspace too small"); return DYF_ERR_ARG; }
  if ((d.static_channels > 0) != (stat != nullptr)) { set_error("static_condition does not match static_channels"); return DYF_ERR_ARG; }
  if (row_offset + (uint64_t)rows > 0xFFFFFFFFull) { set_error("row_offset out of range"); return DYF_ERR_ARG; }
  NvtxRange nvtx("dyf.sampler.run", F->d.arch, rows);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  uint64_t* seed_dev = reinterpret_cast<uint64_t*>(base);
  store_seed_kernel<<<1, 1, 0, s>>>(seed_dev, seed);
  DYF_LAUNCH_OK("store_seed_kernel");
  const size_t used = (size_t)(base - reinterpret_cast<uint8_t*>(ws));
  if (!d.cuda_graph || profiling_enabled())
    return enqueue(rows, ic, stat, preds, x0_out, seed, seed_dev, row_offset, base + 256, ws_bytes - used - 256, s);

  // ---- graph mode: inputs / outputs staged through the workspace so that the captured pointers never change
  size_t o_ic, o_st, o_preds, o_x0;
  const size_t staged = stage_bytes(*this, rows, &o_ic, &o_st, &o_preds, &o_x0);
  const size_t plane = (size_t)F->d.height * F->d.width * sizeof(float);
  float* s_ic = reinterpret_cast<float*>(base + o_ic);
  float* s_st = stat ? reinterpret_cast<float*>(base + o_st) : nullptr;
  float* s_preds = reinterpret_cast<float*>(base + o_preds);
  float* s_x0 = reinterpret_cast<float*>(base + o_x0);
  DYF_CUDA_OK(cudaMemcpyAsync(s_ic, ic, (size_t)rows * d.window_channels * plane, cudaMemcpyDeviceToDevice, s));
  if (stat) DYF_CUDA_OK(cudaMemcpyAsync(s_st, stat, (size_t)rows * d.static_channels * plane, cudaMemcpyDeviceToDevice, s));
  void* loop_ws = base + staged;
  const size_t loop_ws_bytes = ws_bytes - used - staged;

  const auto gkey = std::make_tuple(rows, (const void*)base, row_offset);
  if (graphs.size() >= 32 && !graphs.count(gkey)) {  // bound the cache (a caller sweeping batch sizes): drop everything, re-capture
    for (auto& kv : graphs) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    graphs.clear();
  }
  GraphEntry& g = graphs[gkey];
  if (g.exec && (g.genF != F->generation || g.genI != I->generation)) {  // weights / tables were rebuilt: re-capture
    cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.graph);
    g = GraphEntry{};
  }
  int rc = 0;
  if (g.exec) {
    DYF_CUDA_OK(cudaGraphLaunch(g.exec, s));
    count_launch((int)g.kernels);
    ++graph_replays;
  } else if (g.failed || g.runs == 0) {
    // first run of this key: plain launches (builds the epilogue tables, tensor maps and function attributes that the
    // capture below must not have to create)
    rc = enqueue(rows, s_ic, s_st, s_preds, s_x0, seed, seed_dev, row_offset, loop_ws, loop_ws_bytes, s);
    if (rc) return rc;
    ++g.runs;
  } else {
    const uint64_t l0 = launch_counter();
    cudaError_t ce = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    if (ce == cudaSuccess) {
      rc = enqueue(rows, s_ic, s_st, s_preds, s_x0, seed, seed_dev, row_offset, loop_ws, loop_ws_bytes, s);
      cudaGraph_t graph = nullptr;
      ce = cudaStreamEndCapture(s, &graph);
      if (rc == 0 && ce == cudaSuccess && graph) {
        cudaGraphExec_t exec = nullptr;
        ce = cudaGraphInstantiate(&exec, graph, 0);
        if (ce == cudaSuccess) {
          g.exec = exec; g.graph = graph; g.genF = F->generation; g.genI = I->generation;
          g.kernels = launch_counter() - l0;
          DYF_CUDA_OK(cudaGraphLaunch(g.exec, s));
          ++graph_replays;
        } else {
          cudaGraphDestroy(graph);
        }
      } else if (graph) {
        cudaGraphDestroy(graph);
      }
    }
    if (!g.exec) {  // capture refused (e.g. a cache had to allocate): remember, clear the sticky error, launch plainly
      if (getenv("DYF_DEBUG_GRAPH"))
        fprintf(stderr, "[dyf] graph capture failed (rows=%d row_offset=%u rc=%d): %s | %s\n", rows, (unsigned)row_offset, rc, cudaGetErrorString(ce), get_error());
      cudaGetLastError();
      g.failed = true;
      rc = enqueue(rows, s_ic, s_st, s_preds, s_x0, seed, seed_dev, row_offset, loop_ws, loop_ws_bytes, s);
      if (rc) return rc;
    }
  }
  const size_t state = (size_t)rows * d.channels * plane;
  DYF_CUDA_OK(cudaMemcpyAsync(preds, s_preds, out_keys.size() * state, cudaMemcpyDeviceToDevice, s));
  if (x0_out) DYF_CUDA_OK(cudaMemcpyAsync(x0_out, s_x0, state, cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // namespace dyf
