// On-device ensemble evaluation: CRPS, MSE of the ensemble mean and ensemble variance per sample, without
// the per-horizon device->host copies of the reference (src/experiment_types/forecasting_multi_horizon.py:185-187 feed
// src/utilities/evaluation.py:10-120 with numpy arrays).
//
// CRPS of an N-member ensemble with equal weights (what xskillscore.crps_ensemble -> properscoring integrates):
//   CRPS = 1/N sum_i |x_i - y|  -  1/N^2 sum_{i<j} (x_(j) - x_(i))  =  mean|x - y| - 1/N^2 sum_i (2 rank_i - N + 1) x_i
// with rank_i the position of x_i in the sorted ensemble (ties broken by member index; the sum does not depend on how).
// Ranks are counted, not sorted into place: the members of an element sit in shared memory ([member][thread], conflict
// free) and every thread counts, for each of its members, the smaller ones.  O(N^2) compares per element, no data
// movement, no divergence.
//
// Reductions are two-stage in double precision and in a fixed order (block partials -> per-sample sums), so results are
// bit-reproducible.
#include "engine.hpp"

namespace dyf {
namespace {

constexpr int MT = 256;  // threads per block = elements per block

// grid (blocks_per_sample, samples).  partial: [sample][block][3] = sum crps, sum (mean - y)^2, sum var
__global__ void __launch_bounds__(MT) ensemble_metrics_kernel(const float* __restrict__ preds, const float* __restrict__ tgt,
                                                              int N, long long S, long long D, double* __restrict__ partial) {
  extern __shared__ float s_v[];  // [N][MT]
  __shared__ double s_red[MT / 32][3];
  const long long s = blockIdx.y, i = (long long)blockIdx.x * MT + threadIdx.x;
  const bool active = i < D;
  float crps = 0.f, se = 0.f, var = 0.f;
  if (active) {
    const float y = tgt[s * D + i];
    float sum = 0.f, absdev = 0.f;
    for (int n = 0; n < N; ++n) {
      const float x = preds[((long long)n * S + s) * D + i];
      s_v[n * MT + threadIdx.x] = x;
      sum += x;
      absdev += fabsf(x - y);
    }
    const float mean = sum / (float)N;
    float m2 = 0.f, spread = 0.f;
    for (int n = 0; n < N; ++n) {
      const float x = s_v[n * MT + threadIdx.x];
      int rank = 0;
      for (int j = 0; j < N; ++j) {
        const float z = s_v[j * MT + threadIdx.x];
        rank += (z < x || (z == x && j < n)) ? 1 : 0;
      }
      spread += (float)(2 * rank - N + 1) * x;
      const float d = x - mean;
      m2 += d * d;
    }
    crps = absdev / (float)N - spread / ((float)N * (float)N);
    se = (mean - y) * (mean - y);
    var = m2 / (float)N;  // np.var: population variance (evaluation.py:112)
  }
  double v[3] = {(double)crps, (double)se, (double)var};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 3; ++k) s_red[threadIdx.x >> 5][k] = v[k];
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < MT / 32; ++w) t += s_red[w][threadIdx.x];  // fixed order
    partial[((size_t)s * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
  }
}

__global__ void ensemble_metrics_finish_kernel(const double* __restrict__ partial, int blocks, long long S, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S * 3) return;
  const long long s = t / 3;
  const int k = (int)(t % 3);
  double a = 0.0;
  for (int b = 0; b < blocks; ++b) a += partial[((size_t)s * blocks + b) * 3 + k];  // fixed order
  out[t] = a;
}

// per-member squared error: grid (chunks, N); partial[n][chunk]
__global__ void __launch_bounds__(MT) member_sqerr_kernel(const float* __restrict__ preds, const float* __restrict__ tgt,
                                                          long long SD, long long per_block, double* __restrict__ partial) {
  __shared__ double s_red[MT / 32];
  const long long n = blockIdx.y, b0 = (long long)blockIdx.x * per_block;
  const long long b1 = b0 + per_block < SD ? b0 + per_block : SD;
  double acc = 0.0;
  for (long long e = b0 + threadIdx.x; e < b1; e += MT) {
    const float d = preds[n * SD + e] - tgt[e];
    acc += (double)(d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < MT / 32; ++w) t += s_red[w];
    partial[(size_t)n * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void member_sqerr_finish_kernel(const double* __restrict__ partial, int chunks, int N, double inv_count,
                                           double* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double a = 0.0;
  for (int c = 0; c < chunks; ++c) a += partial[(size_t)n * chunks + c];
  out[n] = a * inv_count;
}

constexpr int MEMBER_CHUNKS = 256;
size_t ws_bytes(int N, long long S, long long D) {
  const size_t blocks = (size_t)((D + MT - 1) / MT);
  return ((size_t)S * blocks * 3 + (size_t)N * MEMBER_CHUNKS) * sizeof(double) + 256;
}

}  // namespace
}  // namespace dyf

using namespace dyf;

extern "C" {

int dyf_ensemble_metrics_workspace_bytes(int32_t n_members, int64_t n_samples, int64_t inner, size_t* bytes) {
  if (!bytes || n_members < 1 || n_samples < 1 || inner < 1) { set_error("bad argument"); return DYF_ERR_ARG; }
  *bytes = ws_bytes(n_members, n_samples, inner);
  return 0;
}

int dyf_ensemble_metrics(const float* preds, const float* targets, int32_t n_members, int64_t n_samples, int64_t inner,
                         double* per_sample, double* per_member_mse, void* workspace, size_t workspace_bytes, void* stream) {
  if (!preds || !targets || !per_sample || !workspace || n_members < 1 || n_samples < 1 || inner < 1) {
    set_error("null or empty argument");
    return DYF_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: dyffusion_b200 has no CPU fallback");
    return DYF_ERR_CUDA;
  }
  if (workspace_bytes < ws_bytes(n_members, n_samples, inner)) { set_error("ensemble metrics: workspace too small"); return DYF_ERR_ARG; }
  const size_t smem = (size_t)n_members * MT * sizeof(float);
  if (smem > 200 * 1024) { set_error("ensemble metrics: at most 200 members"); return DYF_ERR_UNSUPPORTED; }
  if (n_samples > 65535) { set_error("ensemble metrics: at most 65535 samples per call"); return DYF_ERR_UNSUPPORTED; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static size_t configured = 0;
  if (smem > configured) {
    DYF_CUDA_OK(cudaFuncSetAttribute(ensemble_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  double* partial = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  const int blocks = (int)((inner + MT - 1) / MT);
  ProfScope prof(s, KC_ELEMENTWISE);
  ensemble_metrics_kernel<<<dim3(blocks, (unsigned)n_samples), MT, smem, s>>>(preds, targets, n_members, n_samples, inner, partial);
  DYF_LAUNCH_OK("ensemble_metrics_kernel");
  ensemble_metrics_finish_kernel<<<cdiv(n_samples * 3, 256), 256, 0, s>>>(partial, blocks, n_samples, per_sample);
  DYF_LAUNCH_OK("ensemble_metrics_finish_kernel");
  if (per_member_mse) {
    double* mpart = partial + (size_t)n_samples * blocks * 3;
    const long long SD = n_samples * inner;
    const long long per_block = ((SD + MEMBER_CHUNKS - 1) / MEMBER_CHUNKS + MT - 1) / MT * MT;
    const int chunks = (int)((SD + per_block - 1) / per_block);
    member_sqerr_kernel<<<dim3(chunks, n_members), MT, 0, s>>>(preds, targets, SD, per_block, mpart);
    DYF_LAUNCH_OK("member_sqerr_kernel");
    member_sqerr_finish_kernel<<<cdiv(n_members, 128), 128, 0, s>>>(mpart, chunks, n_members, 1.0 / (double)SD, per_member_mse);
    DYF_LAUNCH_OK("member_sqerr_finish_kernel");
  }
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Boundary conditions of the physical-systems benchmark on device (reference:
// src/datamodules/physical_systems_benchmark.py:245-297, a per-sample Python loop of masked writes).
namespace dyf {
namespace {

// Navier-Stokes (:253-276): preds[b, c, h, w] = 0 where fixed_mask[b, c, h, w]; then the parabolic inflow profile on
// channel 0, first grid row: in_velocity * 4 * y * (0.41 - y) / 0.41^2 * (1 - exp(-5 t)), y = vertex_y[b, w].
__global__ void bc_navier_stokes_kernel(float* __restrict__ preds, const uint8_t* __restrict__ mask, const float* __restrict__ vertex_y,
                                        const float* __restrict__ in_velocity, const float* __restrict__ time, int time_stride,
                                        int B, int C, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)C * H * W;
  if (idx >= (long long)B * per) return;
  const int b = (int)(idx / per);
  const long long e = idx - (long long)b * per;
  const int w = (int)(e % W), h = (int)((e / W) % H), c = (int)(e / ((long long)W * H));
  float v = preds[idx];
  if (mask[idx]) v = 0.f;
  if (c == 0 && h == 0) {
    const float y = vertex_y[(size_t)b * W + w];
    const float t = time[(size_t)b * time_stride];
    float p = in_velocity[b] * 4.f;          // same operation order as the reference's fp32 tensor expression
    p = p * y;
    p = p * (0.41f - y);
    p = p / (float)(0.41 * 0.41);
    p = p * (float)(1.0 - exp(-5.0 * (double)t));
    v = p;
  }
  preds[idx] = v;
}

// spring-mesh (:277-287): preds[(e,) b] = where(fixed_mask[b], boundary[b], preds[(e,) b]); boundary = cat(0, base_q)
__global__ void bc_spring_mesh_kernel(float* __restrict__ preds, const uint8_t* __restrict__ mask, const float* __restrict__ base_q,
                                      long long lead, int B, int HW) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = 4LL * HW;
  if (idx >= lead * B * per) return;
  const int b = (int)((idx / per) % B);
  const long long e = idx % per;
  const int c = (int)(e / HW);
  if (mask[(size_t)b * per + e]) preds[idx] = c < 2 ? 0.f : base_q[((size_t)b * 2 + (c - 2)) * HW + (e % HW)];
}

}  // namespace
}  // namespace dyf

extern "C" {

int dyf_boundary_conditions_navier_stokes(float* preds, const uint8_t* fixed_mask, const float* vertex_y, const float* in_velocity,
                                          const float* time, int32_t time_per_sample, int32_t batch, int32_t channels,
                                          int32_t height, int32_t width, void* stream) {
  if (!preds || !fixed_mask || !vertex_y || !in_velocity || !time || batch < 1 || channels < 1 || height < 1 || width < 1) {
    set_error("null or empty argument");
    return DYF_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: dyffusion_b200 has no CPU fallback"); return DYF_ERR_CUDA; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)batch * channels * height * width;
  ProfScope prof(s, KC_ELEMENTWISE);
  bc_navier_stokes_kernel<<<cdiv(total, 256), 256, 0, s>>>(preds, fixed_mask, vertex_y, in_velocity, time, time_per_sample ? 1 : 0,
                                                         batch, channels, height, width);
  DYF_LAUNCH_OK("bc_navier_stokes_kernel");
  return 0;
}

int dyf_boundary_conditions_spring_mesh(float* preds, const uint8_t* fixed_mask, const float* base_q, int64_t lead, int32_t batch,
                                        int32_t height, int32_t width, void* stream) {
  if (!preds || !fixed_mask || !base_q || lead < 1 || batch < 1 || height < 1 || width < 1) { set_error("null or empty argument"); return DYF_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: dyffusion_b200 has no CPU fallback"); return DYF_ERR_CUDA; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = lead * batch * 4LL * height * width;
  ProfScope prof(s, KC_ELEMENTWISE);
  bc_spring_mesh_kernel<<<cdiv(total, 256), 256, 0, s>>>(preds, fixed_mask, base_q, lead, batch, height * width);
  DYF_LAUNCH_OK("bc_spring_mesh_kernel");
  return 0;
}

}  // extern "C"
