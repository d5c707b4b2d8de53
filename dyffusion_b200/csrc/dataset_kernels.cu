// GPU-resident sliding-window dataset (second half).
//
// The reference materialises every training / evaluation example on the host: for each trajectory a
// `sliding_window_view` over the time axis, re-arranged to (example, window+horizon, C, H, W) and concatenated over
// trajectories (src/datamodules/physical_systems_benchmark.py:191-243), then a DataLoader collates batches of those copies
// and Lightning moves them to the device.  An example is `window + horizon` CONSECUTIVE frames of one trajectory, i.e. one
// contiguous run of the trajectory store -- so here the trajectories stay in HBM once, back to back, and a batch is
// gathered from them by start frame: pure data movement, HBM-bound (algorithmic bytes = 2 x 4 B per gathered element).
//
// Layout: frames [n_frames][frame_elems] fp32 -> out [batch][frames_per_example][frame_elems] fp32.  A CTA row (blockIdx.y)
// is one example; threads stream its run with the widest access both the run start and the output allow
// (16 B when frame_elems % 4 == 0, 8 B when even -- Navier-Stokes frames hold 3*221*42 = 27 846 floats -- else 4 B), four
// independent loads in flight per thread before the stores.
#include "engine.hpp"

namespace dyf {
namespace {

constexpr int GATHER_THREADS = 256;
constexpr int GATHER_UNROLL = 4;
constexpr int GATHER_TABLE = 128;  // examples per launch (start frames travel as a kernel parameter: no device table, no copy)

struct GatherTable {
  long long first_frame[GATHER_TABLE];
};

template <typename V>
__global__ void __launch_bounds__(GATHER_THREADS) window_gather_kernel(const float* __restrict__ frames, float* __restrict__ out,
                                                                      long long run_elems, long long frame_elems,
                                                                      int example0, GatherTable table) {
  constexpr int VE = sizeof(V) / sizeof(float);
  const int e = blockIdx.y;
  const V* __restrict__ src = reinterpret_cast<const V*>(frames + table.first_frame[e] * frame_elems);
  V* __restrict__ dst = reinterpret_cast<V*>(out + (long long)(example0 + e) * run_elems);
  const long long n = run_elems / VE;  // run_elems % VE == 0 by construction of the dispatch below
  const long long stride = (long long)gridDim.x * GATHER_THREADS;
  long long i = (long long)blockIdx.x * GATHER_THREADS + threadIdx.x;
  for (; i + (GATHER_UNROLL - 1) * stride < n; i += GATHER_UNROLL * stride) {
    V v[GATHER_UNROLL];
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) v[u] = __ldcs(src + i + u * stride);  // streamed once: evict-first
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < n; i += stride) dst[i] = __ldcs(src + i);
}

template <typename V>
int launch_gather(const float* frames, float* out, long long run_elems, long long frame_elems, const int64_t* first, int batch,
                  cudaStream_t s) {
  constexpr int VE = sizeof(V) / sizeof(float);
  const long long per_block = (long long)GATHER_THREADS * GATHER_UNROLL * VE;
  long long bx = cdiv(run_elems, per_block);
  if (bx > 4096) bx = 4096;  // grid-stride beyond that; 64 examples x 4096 CTAs is already >> 148 SMs x resident CTAs
  for (int e0 = 0; e0 < batch; e0 += GATHER_TABLE) {
    const int n = batch - e0 < GATHER_TABLE ? batch - e0 : GATHER_TABLE;
    GatherTable t;
    for (int i = 0; i < n; ++i) t.first_frame[i] = first[e0 + i];
    ProfScope prof(s, KC_PACK, 0.0, 8.0 * (double)run_elems * n);
    window_gather_kernel<V><<<dim3((unsigned)bx, (unsigned)n), GATHER_THREADS, 0, s>>>(frames, out, run_elems, frame_elems, e0, t);
    DYF_LAUNCH_OK("window_gather_kernel");
  }
  return 0;
}

}  // namespace
}  // namespace dyf

extern "C" {

int dyf_window_gather(const float* frames, int64_t n_frames, int64_t frame_elems, const int64_t* first_frame_host, int32_t batch,
                      int32_t frames_per_example, float* out, void* stream) {
  using namespace dyf;
  if (!frames || !first_frame_host || !out || n_frames < 1 || frame_elems < 1 || batch < 1 || frames_per_example < 1) {
    set_error("null or empty argument");
    return DYF_ERR_ARG;
  }
  for (int i = 0; i < batch; ++i) {
    const int64_t f = first_frame_host[i];
    if (f < 0 || f + frames_per_example > n_frames) {
      set_error("example " + std::to_string(i) + ": frames [" + std::to_string(f) + ", " + std::to_string(f + frames_per_example) +
                ") lie outside the trajectory store of " + std::to_string(n_frames) + " frames");
      return DYF_ERR_ARG;
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: dyffusion_b200 has no CPU fallback"); return DYF_ERR_CUDA; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long run = (long long)frames_per_example * frame_elems;
  const uintptr_t a = reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(out);
  // every run start is a multiple of frame_elems floats from an aligned base, every output start a multiple of run
  if (frame_elems % 4 == 0 && a % 16 == 0) return launch_gather<float4>(frames, out, run, frame_elems, first_frame_host, batch, s);
  if (frame_elems % 2 == 0 && a % 8 == 0) return launch_gather<float2>(frames, out, run, frame_elems, first_frame_host, batch, s);
  return launch_gather<float>(frames, out, run, frame_elems, first_frame_host, batch, s);
}

}  // extern "C"
