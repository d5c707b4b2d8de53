/* dyffusion_b200 -- C ABI of the B200-native DYffusion sampling engine.
 *
 * This is the drop-in boundary of the hot path.  The reference has no FFI of its own for this
 * path: its "plug-in" interface is two Hydra `_target_` classes (model + diffusion) whose arithmetic runs in
 * PyTorch/ATen.  The Python drop-in classes in `dyffusion_b200/` keep that surface (same constructor kwargs,
 * state-dict keys, `forward/predict_forward/sample/sample_loop`) and forward the arithmetic to the entry points
 * below through ctypes.  Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 *
 * Conventions: plain C, no C++ types or exceptions across the boundary.  Every function returns 0 on success and
 * a negative code on failure; `dyf_last_error()` returns a thread-local message (the Python side raises it as
 * RuntimeError -- the reference signals errors with Python exceptions, e.g. src/diffusion/dyffusion.py:42,47,68).
 * All data pointers are DEVICE pointers owned by the caller unless stated otherwise; tensors are fp32, NCHW,
 * contiguous, exactly as the reference passes them.  The engine owns only its handles (re-packed weights, layer
 * plans).  Workspace is caller-allocated and sized by a query.  All work is enqueued on the caller's stream
 * (`void*` = cudaStream_t); no host synchronisation, no internal threads.  Handles are not thread-safe.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with DYF_ERR_CUDA.
 */
#ifndef DYFFUSION_B200_H
#define DYFFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

#define DYF_ABI_VERSION 2

enum { DYF_OK = 0, DYF_ERR_ARG = -1, DYF_ERR_CUDA = -2, DYF_ERR_STATE = -3, DYF_ERR_UNSUPPORTED = -4 };

/* Backbones on the hot path. */
enum {
  DYF_ARCH_UNET_SIMPLE = 0, /* src/models/unet_simple.py:86-197   (Navier-Stokes)  */
  DYF_ARCH_UNET_RESNET = 1, /* src/models/unet.py:114-315          (SST)            */
  DYF_ARCH_CONVNET = 2      /* src/models/simple_conv_net.py:59-131 (spring-mesh)   */
};

/* Constructor arguments of the three reference backbones that change the arithmetic. */
typedef struct dyf_net_desc {
  int32_t arch;
  int32_t dim;             /* `dim` */
  int32_t in_channels;     /* num_input_channels  (src/models/_base_model.py:44-46)            */
  int32_t cond_channels;   /* num_conditional_channels                                          */
  int32_t out_channels;    /* num_output_channels                                               */
  int32_t height, width;   /* spatial_shape                                                     */
  int32_t with_time_emb;   /* `with_time_emb` */
  /* unet_simple.UNet */
  int32_t upsample_h, upsample_w; /* `upsample_dims` (0,0 = None)   unet_simple.py:91,98-101     */
  float dropout;           /* `dropout` (unet_simple / SimpleConvNet)                            */
  float input_dropout;     /* `input_dropout` (unet_simple / Unet)                               */
  /* unet.Unet */
  int32_t n_mults;
  int32_t dim_mults[8];    /* `dim_mults` */
  int32_t groups;          /* `resnet_block_groups` */
  float block_dropout;     /* dropout of block2 (unet.py:151) */
  float block_dropout1;    /* dropout of block1 (unet.py:150) */
  float attn_dropout;      /* unet.py:187,209 */
  int32_t keep_spatial_dims;
  int32_t init_kernel, init_padding, init_stride;
  /* SimpleConvNet */
  int32_t n_kernels;
  int32_t kernel_sizes[8]; /* `kernel_sizes` */
  int32_t residual;        /* `residual` */
} dyf_net_desc;

typedef struct dyf_net dyf_net;
typedef struct dyf_sampler dyf_sampler;

/* Dropout control of one forward.  mode 0 = off (the reference's eval mode), 1 = on with the engine's
 * counter-based Philox stream keyed by (seed, stream) (the reference's "inference dropout",
 * src/utilities/utils.py:560-574 + src/models/_base_model.py:148-161).  `stream` must differ between calls that
 * must draw independent masks.  `row_offset` is the index of batch row 0 in the un-sharded job: masks (and the
 * "data+noise" noise) are functions of the GLOBAL row, so a job split over several ranks draws exactly the masks of the
 * same job on one rank (rows are the shard axis). */
typedef struct dyf_dropout {
  int32_t mode;
  uint64_t seed;
  uint64_t stream;
  uint64_t row_offset;
} dyf_dropout;

int dyf_abi_version(void);
const char* dyf_last_error(void);
/* Storage type of activations and tensor-core operands the library was built with: "fp16" (default) or "bf16"
 * (accumulation, epilogues, statistics and the sampler state are fp32 either way). */
const char* dyf_act_dtype(void);
/* Number of CUDA kernels the engine has launched in this process (for bench.py's `gpu_launches`). */
uint64_t dyf_launch_count(void);

/* Measurement hooks (bench.py): while enabled, every engine kernel launch is bracketed by CUDA events on its own
 * stream.  `dyf_profile_read` waits for the recorded events and returns, per kernel class (DYF_KC_*), the summed
 * device time [ms], algorithmic FLOPs, algorithmic bytes and launch count since the last read; arrays must hold
 * DYF_KC_COUNT entries. */
enum { DYF_KC_CONV_MMA = 0, DYF_KC_CONV_UMMA, DYF_KC_PACK, DYF_KC_UPSAMPLE, DYF_KC_GROUPNORM, DYF_KC_READOUT,
       DYF_KC_TIME, DYF_KC_ELEMENTWISE, DYF_KC_ATTENTION, DYF_KC_CONV_UP, DYF_KC_CONV_FLAT, DYF_KC_COUNT };
int dyf_profile_enable(int32_t on);
/* Restrict the event bracketing to one kernel class (DYF_KC_*; -1 = all classes): lets bench.py time its dominant kernel
 * live inside the timed region without putting event records between every other pair of launches. */
int dyf_profile_filter(int32_t klass);
int dyf_profile_read(double* ms, double* flops, double* bytes, uint64_t* launches, int32_t n_classes);

/* Replaces: `hydra.utils.instantiate(model_config, ...)` -> backbone constructor
 * (src/experiment_types/_base_experiment.py:180-188). */
int dyf_net_create(const dyf_net_desc* desc, dyf_net** out);
void dyf_net_destroy(dyf_net* net);

/* Replaces: `nn.Module.load_state_dict` for one tensor.  `key` is the reference's state-dict key,
 * `data` a device pointer to fp32 (int64 `num_batches_tracked` entries are accepted and ignored), `shape/ndim`
 * the tensor's shape.  Unknown keys and shape mismatches are errors (strict loading). */
int dyf_net_set_param(dyf_net* net, const char* key, const void* data, const int64_t* shape, int32_t ndim);

/* Folds eval-mode BatchNorm / weight standardisation, re-packs weights to the kernel layouts (bf16, KRSC).
 * Must be called after all parameters are set and again after any parameter changes.  Fails, naming the first
 * missing key, if a parameter was never set. */
int dyf_net_finalize(dyf_net* net, void* stream);

/* Number of state-dict keys the backbone expects and the i-th key (for strictness checks on the host). */
int dyf_net_num_params(const dyf_net* net);
const char* dyf_net_param_key(const dyf_net* net, int32_t i);
/* Shape of the i-th parameter (`shape` receives up to 4 extents); returns ndim, or <0 on error.  `is_buffer` is set
 * for BatchNorm running statistics / num_batches_tracked (nn.Module buffers rather than parameters). */
int dyf_net_param_shape(const dyf_net* net, int32_t i, int64_t* shape, int32_t* is_buffer);

int dyf_net_workspace_bytes(const dyf_net* net, int32_t rows, size_t* bytes);

/* Replaces: backbone `forward(inputs, time=, condition=)` (src/models/unet_simple.py:181-197, src/models/unet.py:266-315,
 * src/models/simple_conv_net.py:112-131).  x: [rows, in_channels, H, W]; cond: [rows, cond_channels, H, W] or NULL when
 * cond_channels == 0; time: [rows] fp32 or NULL when with_time_emb == 0; y: [rows, out_channels, H, W]. */
int dyf_net_forward(dyf_net* net, int32_t rows, const float* x, const float* cond, const float* time, float* y,
                    const dyf_dropout* drop, void* workspace, size_t workspace_bytes, void* stream);

/* Same, with the channel concatenation expressed as a list of source tensors in concat order (each
 * [rows, src_channels[i], H, W]); lets the sampler skip the reference's torch.cat calls
 * (src/diffusion/dyffusion.py:189, :488; src/models/unet_simple.py:184). */
int dyf_net_forward_srcs(dyf_net* net, int32_t rows, const float* const* srcs, const int32_t* src_channels,
                         int32_t nsrc, const float* time, float* y, const dyf_dropout* drop, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Sampler description = the DYffusion hyper-parameters that `sample_loop` reads (src/diffusion/dyffusion.py:18-95,
 * :335-426).  The schedule is passed already parsed: `schedule[i]` = diffusion step s_i (may be fractional),
 * `tau[i]` = diffusion_step_to_interpolation_step(s_i) (:101-138); both computed on the host by the drop-in class. */
typedef struct dyf_sampler_desc {
  int32_t num_timesteps;          /* N (after auxiliary steps) */
  int32_t n_schedule;
  const double* schedule;         /* HOST pointer, n_schedule entries */
  const double* tau;              /* HOST pointer, n_schedule entries */
  const double* time_forecaster;  /* HOST pointer, n_schedule entries: time fed to the forecaster (time_encoding) */
  int32_t forward_conditioning;   /* 0 = "none", 1 = "data", 2 = "data+noise" (:205-239) */
  int32_t sampling_type;          /* 0 = "cold", 1 = "naive" (:381-393) */
  int32_t use_cold_sampling_for_last_step;
  int32_t n_refine;               /* number of refinement times (0 = refine_intermediate_predictions False) */
  const double* refine_times;     /* HOST pointer (:408-422) */
  int32_t enable_interpolator_dropout; /* (:154) */
  int32_t channels;               /* C = forecaster in_channels */
  int32_t window_channels;        /* channels of `initial_condition` (C * window) */
  int32_t static_channels;        /* channels of `static_condition` (0 = None) */
  int32_t interpolator_horizon;   /* for the 0 < t < horizon check (:484-486) */
  int32_t max_rows_per_call;      /* cap on rows per backbone launch when calls are batched (0 = default) */
  int32_t forecaster_dropout;     /* dropout also ON in the forecaster: the reference's experiment-level
                                     `enable_inference_dropout` (src/experiment_types/_base_experiment.py:288-299) */
  int32_t cuda_graph;             /* 1: after one plain run per (rows, workspace, row_offset) the whole launch sequence of
                                     `dyf_sampler_run` is captured into a CUDA graph and replayed (inputs / outputs are staged
                                     through the workspace, the seed lives in device memory) */
} dyf_sampler_desc;

int dyf_sampler_create(dyf_net* forecaster, dyf_net* interpolator, const dyf_sampler_desc* desc, dyf_sampler** out);
void dyf_sampler_destroy(dyf_sampler* s);
int dyf_sampler_workspace_bytes(const dyf_sampler* s, int32_t rows, size_t* bytes);
/* Number of output slots (`preds` holds n_outputs tensors of [rows, C, H, W]); slot j carries the forecast whose
 * reference key is "t{out_key[j]}_preds"; keys[] is filled with n_outputs doubles. */
int dyf_sampler_num_outputs(const dyf_sampler* s, int32_t* n_outputs, double* keys, int32_t keys_capacity);

/* Replaces: `BaseDYffusion.sample_loop` / `sample` (src/diffusion/dyffusion.py:335-431).
 * ic: [rows, window_channels, H, W]; static_cond: [rows, static_channels, H, W] or NULL;
 * preds: [n_outputs, rows, C, H, W] fp32; x0_hat_out (optional): final forecaster output [rows, C, H, W].
 * `row_offset`: index of row 0 in the un-sharded job (0 when the job is not sharded), see dyf_dropout. */
int dyf_sampler_run(dyf_sampler* s, int32_t rows, const float* ic, const float* static_cond, float* preds,
                    float* x0_hat_out, uint64_t seed, uint64_t row_offset, void* workspace, size_t workspace_bytes,
                    void* stream);
/* Measurement hook (no reference counterpart): how many `dyf_sampler_run` calls of this sampler were served by replaying a
 * captured CUDA graph (0 while `cuda_graph` is off, profiling is on, or every capture was refused).  A caller on the legacy
 * default stream -- which cannot be captured -- is served from the sampler's own stream, fenced with events on both sides. */
int dyf_sampler_graph_replays(const dyf_sampler* s, uint64_t* replays);
/* NVTX ranges ("dyf.net.forward <arch> rows=..", "dyf.sampler.run") around every network call / sampler run, for
 * nsys / ncu timelines (off by default; also switched on by the environment variable DYF_NVTX=1). */
int dyf_nvtx_enable(int32_t on);

/* Widening row SURVEY.md 8f-2 -- on-device ensemble evaluation.
 * Replaces: `evaluate_ensemble_prediction` (src/utilities/evaluation.py:10-80), `evaluate_ensemble_crps` (:83-97, i.e.
 * xskillscore.crps_ensemble with equal member weights) and `evaluate_ensemble_spread_skill_ratio` (:100-120), which the
 * reference feeds with numpy copies of every horizon (src/experiment_types/forecasting_multi_horizon.py:185-187, :246).
 * preds: [n_members, n_samples, inner] fp32 (device), targets: [n_samples, inner]; `inner` = product of the remaining
 * dimensions.  per_sample (device, [n_samples][3] doubles) receives, per sample, the SUMS over `inner` of: the CRPS, the
 * squared error of the ensemble mean, the population variance over members -- the host finishes
 * crps = sum/inner, mse = sum/inner, ssr = sqrt(mean var) / sqrt(mse) for either setting of `mean_over_samples`.
 * per_member_mse (device, [n_members] doubles, may be NULL): mean squared error of every member (:47-53).
 * Sums are accumulated in a fixed order (bit-reproducible). */
int dyf_ensemble_metrics_workspace_bytes(int32_t n_members, int64_t n_samples, int64_t inner, size_t* bytes);
int dyf_ensemble_metrics(const float* preds, const float* targets, int32_t n_members, int64_t n_samples, int64_t inner,
                         double* per_sample, double* per_member_mse, void* workspace, size_t workspace_bytes, void* stream);

/* Widening row SURVEY.md 8f-3 (first half) -- the benchmark's boundary conditions as one masked-write kernel.
 * Replaces: `PhysicalSystemsBenchmarkDataModule.boundary_conditions` (src/datamodules/physical_systems_benchmark.py:245-297),
 * a per-sample Python loop.  In place on `preds`.
 * Navier-Stokes (:253-276): preds [batch, channels, H, W]; fixed_mask [batch, channels, H, W] (bytes, non-zero = fixed ->
 * 0); then channel 0 / grid row 0 <- in_velocity[b] * 4 y (0.41 - y) / 0.41^2 * (1 - exp(-5 t)), y = vertex_y[b, w]
 * (vertex_y = metadata["vertices"][:, 1, 0, :]); time: one value (time_per_sample = 0) or [batch].
 * spring-mesh (:277-287): preds [lead, batch, 4, H, W] (lead = 1 or the ensemble size); fixed_mask [batch, 4, H, W];
 * base_q [batch, 2, H, W] (= metadata["features"][:, 0, 2:]); fixed p <- 0, fixed q <- base_q. */
int dyf_boundary_conditions_navier_stokes(float* preds, const uint8_t* fixed_mask, const float* vertex_y, const float* in_velocity,
                                          const float* time, int32_t time_per_sample, int32_t batch, int32_t channels,
                                          int32_t height, int32_t width, void* stream);
int dyf_boundary_conditions_spring_mesh(float* preds, const uint8_t* fixed_mask, const float* base_q, int64_t lead, int32_t batch,
                                        int32_t height, int32_t width, void* stream);

/* Widening row SURVEY.md 8f-4 (second half) -- GPU-resident sliding-window dataset.
 * Replaces: the example tensor that `PhysicalSystemsBenchmarkDataModule.create_dataset_multi_horizon` materialises on the host
 * (src/datamodules/physical_systems_benchmark.py:191-243: `sliding_window_view` over every trajectory, re-arranged to
 * (example, window + horizon, C, H, W) and concatenated) together with the DataLoader collation / host->device copy of each
 * batch.  The trajectories stay in HBM once, back to back; an example is `frames_per_example` consecutive frames.
 * frames: [n_frames, frame_elems] fp32 (device); first_frame_host: HOST array of `batch` start frames (trajectory base +
 * example offset; it travels as a kernel parameter); out: [batch, frames_per_example, frame_elems] fp32 (device).
 * A window reaching outside [0, n_frames) is DYF_ERR_ARG (keeping windows inside ONE trajectory is the caller's index
 * arithmetic, dyffusion_b200/datasets.py).  Also used with frames_per_example = 1 to gather per-trajectory tensors
 * (the static `condition`, masks) by trajectory index. */
int dyf_window_gather(const float* frames, int64_t n_frames, int64_t frame_elems, const int64_t* first_frame_host, int32_t batch,
                      int32_t frames_per_example, float* out, void* stream);

/* Widening row SURVEY.md 8f-1 (optimizer half) -- fused AdamW with global-norm gradient clipping over flat fp32 arrays.
 * Replaces: `torch.optim.AdamW.step` of the optimizer `BaseExperiment._get_optim` builds
 * (src/experiment_types/_base_experiment.py:711-725, src/configs/optimizer/adamw.yaml) preceded by Lightning's
 * `gradient_clip_val` pass (src/configs/trainer/default.yaml:10; torch.nn.utils.clip_grad_norm_ semantics: every gradient
 * scaled by min(1, max_norm / (||g||_2 + 1e-6)) over ALL parameters).  params / grads / exp_avg / exp_avg_sq: device arrays
 * of n floats, 16-byte aligned (the host side keeps all tensors as views of four arenas); `step` counts from 1;
 * `max_grad_norm` <= 0 disables clipping.  The squared norm is reduced on the device in a fixed order and consumed by
 * the update kernel without a host round trip; `dyf_grad_sq_norm` exposes it ([1] double, device) for logging.
 * Arithmetic: fp32, in the operation order of torch's single-tensor AdamW. */
int dyf_adamw_workspace_bytes(int64_t n, size_t* bytes);
int dyf_grad_sq_norm(const float* grads, int64_t n, double* out, void* workspace, size_t workspace_bytes, void* stream);
int dyf_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int64_t step, double max_grad_norm, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Test hook: writes the keep-mask (1/0 bytes) the engine's dropout draws for a tensor of `n_elems` elements
 * (NHWC element order, channel count `channels`) at (seed, stream, site, p).  Lets tests replay engine masks in
 * the oracle. */
int dyf_debug_dropout_mask(uint64_t seed, uint64_t stream, uint32_t site, float p, int64_t n_elems, uint8_t* mask,
                           void* stream_handle);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DYFFUSION_B200_H */
