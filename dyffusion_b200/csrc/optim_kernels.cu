// Optimizer half of the training step: AdamW over a flat fp32 arena with the global-norm gradient clipping
// folded in.  Reference: `torch.optim.AdamW` built by `BaseExperiment._get_optim` (src/experiment_types/_base_experiment.py:
// 711-725; src/configs/optimizer/adamw.yaml: betas (0.9, 0.99), eps 1e-8) and Lightning's `gradient_clip_val: 1.0`
// (src/configs/trainer/default.yaml:10 = torch.nn.utils.clip_grad_norm_): ~100 small tensors, one clip pass that re-writes
// every gradient, then the foreach update.  Here: all parameters / gradients / moments live in four flat arrays, the
// squared gradient norm is reduced in a fixed order (fp64 partials -> one device scalar, no host sync), and ONE update
// kernel reads the clip coefficient from that scalar and applies it while loading the gradient.
//
// HBM-bound: the norm pass reads 4 B per element, the update reads 16 B (p, g, m, v) and writes 12 B (p, m, v).
// Update arithmetic in fp32 in torch's single-tensor order (torch/optim/adamw.py::_single_tensor_adamw):
//   p *= 1 - lr*wd;  m += (g - m)*(1 - b1);  v = v*b2 + (1 - b2)*g*g;  p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps).
#include "engine.hpp"

namespace dyf {
namespace {

constexpr int OT = 256;

__global__ void __launch_bounds__(OT) grad_sq_partial_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial) {
  __shared__ double s_red[OT / 32];
  double acc = 0.0;
  const long long n4 = n >> 2;
  const float4* __restrict__ g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * OT + threadIdx.x; i < n4; i += (long long)gridDim.x * OT) {
    const float4 v = g4[i];
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; acc += (double)v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < OT / 32; ++w) t += s_red[w];  // fixed order
    partial[blockIdx.x] = t;
  }
}

// One warp; lane l adds partial[l], partial[l + 32], ... in that order, then a fixed butterfly: bit-reproducible, and the loads
// of a lane are independent (a single thread walking the 1 184 partials took 48 us, a quarter of the whole step).
__global__ void __launch_bounds__(32) grad_sq_final_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ out) {
  double t = 0.0;
  for (int b = threadIdx.x; b < blocks; b += 32) t += partial[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) out[0] = t;
}

struct AdamArgs {
  float decay;        // 1 - lr * weight_decay
  float one_m_b1, b2, one_m_b2;
  float step_size;    // lr / (1 - b1^step)
  float bc2_sqrt;     // sqrt(1 - b2^step): torch divides sqrt(v) by it
  float eps;
  float max_norm;     // <= 0: no clipping
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamArgs& a, float clip) {
  g *= clip;
  p *= a.decay;
  m = m + (g - m) * a.one_m_b1;                 // lerp_
  v = v * a.b2 + a.one_m_b2 * g * g;            // mul_ + addcmul_
  // IEEE square root / divisions: the library is built with --use_fast_math, whose approximate forms are not good enough here
  const float denom = __fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt) + a.eps;
  p = p - a.step_size * __fdiv_rn(m, denom);    // addcdiv_
}

__global__ void __launch_bounds__(OT) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, AdamArgs a,
                                                   const double* __restrict__ grad_sq) {
  float clip = 1.f;
  if (a.max_norm > 0.f) {  // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = (float)sqrt(grad_sq[0]);
    clip = fminf(__fdiv_rn(a.max_norm, total + 1e-6f), 1.f);
  }
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = (long long)blockIdx.x * OT + threadIdx.x; i < n4; i += (long long)gridDim.x * OT) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldcs(g4 + i);
    adam1(pp.x, gg.x, mm.x, vv.x, a, clip);
    adam1(pp.y, gg.y, mm.y, vv.y, a, clip);
    adam1(pp.z, gg.z, mm.z, vv.z, a, clip);
    adam1(pp.w, gg.w, mm.w, vv.w, a, clip);
    p4[i] = pp, m4[i] = mm, v4[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float pp = p[i], mm = m[i], vv = v[i];
    adam1(pp, g[i], mm, vv, a, clip);
    p[i] = pp, m[i] = mm, v[i] = vv;
  }
}

int norm_blocks(long long n) {
  long long b = cdiv(n, (long long)OT * 4 * 4);
  return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b));  // <= 148 SMs x 8 resident CTAs
}

}  // namespace
}  // namespace dyf

extern "C" {

int dyf_adamw_workspace_bytes(int64_t n, size_t* bytes) {
  using namespace dyf;
  if (n < 1 || !bytes) { set_error("null or empty argument"); return DYF_ERR_ARG; }
  *bytes = (size_t)(norm_blocks(n) + 1) * sizeof(double);
  return 0;
}

int dyf_grad_sq_norm(const float* grads, int64_t n, double* out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace dyf;
  if (!grads || !out || !workspace || n < 1) { set_error("null or empty argument"); return DYF_ERR_ARG; }
  const int blocks = norm_blocks(n);
  if (workspace_bytes < (size_t)(blocks + 1) * sizeof(double)) { set_error("workspace too small"); return DYF_ERR_ARG; }
  if ((reinterpret_cast<uintptr_t>(grads) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 7)) { set_error("arrays must be 16-byte aligned"); return DYF_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: dyffusion_b200 has no CPU fallback"); return DYF_ERR_CUDA; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  double* partial = reinterpret_cast<double*>(workspace);
  ProfScope prof(s, KC_ELEMENTWISE, 0.0, 4.0 * (double)n);
  grad_sq_partial_kernel<<<blocks, OT, 0, s>>>(grads, n, partial);
  DYF_LAUNCH_OK("grad_sq_partial_kernel");
  grad_sq_final_kernel<<<1, 32, 0, s>>>(partial, blocks, out);
  DYF_LAUNCH_OK("grad_sq_final_kernel");
  return 0;
}

int dyf_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int64_t step, double max_grad_norm, void* workspace,
                   size_t workspace_bytes, void* stream) {
  using namespace dyf;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !workspace || n < 1 || step < 1) { set_error("null or empty argument (step counts from 1)"); return DYF_ERR_ARG; }
  if (!(lr >= 0.0) || !(eps >= 0.0) || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(weight_decay >= 0.0)) {
    set_error("invalid AdamW hyper-parameter");  // torch.optim.AdamW raises ValueError for the same ranges
    return DYF_ERR_ARG;
  }
  const uintptr_t al = reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                       reinterpret_cast<uintptr_t>(exp_avg_sq);
  if (al & 15) { set_error("arrays must be 16-byte aligned"); return DYF_ERR_ARG; }
  const int blocks = norm_blocks(n);
  if (workspace_bytes < (size_t)(blocks + 1) * sizeof(double)) { set_error("workspace too small"); return DYF_ERR_ARG; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  double* partial = reinterpret_cast<double*>(workspace);
  double* total = partial + blocks;
  if (max_grad_norm > 0.0) {
    const int rc = dyf_grad_sq_norm(grads, n, total, workspace, workspace_bytes, stream);
    if (rc != 0) return rc;
  } else {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: dyffusion_b200 has no CPU fallback"); return DYF_ERR_CUDA; }
  }
  AdamArgs a;
  a.decay = (float)(1.0 - lr * weight_decay);
  a.one_m_b1 = (float)(1.0 - beta1);
  a.b2 = (float)beta2;
  a.one_m_b2 = (float)(1.0 - beta2);
  a.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.eps = (float)eps;
  a.max_norm = (float)max_grad_norm;
  long long blk = cdiv(n, (long long)OT * 4);
  if (blk > 148 * 16) blk = 148 * 16;
  ProfScope prof(s, KC_ELEMENTWISE, 0.0, 28.0 * (double)n);
  adamw_kernel<<<(unsigned)blk, OT, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, a, total);
  DYF_LAUNCH_OK("adamw_kernel");
  return 0;
}

}  // extern "C"
