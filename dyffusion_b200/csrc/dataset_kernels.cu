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
// (16 B when frame_elems % 4 == 0; 16-byte stores with 16- or 2 x 8-byte loads when it is even -- Navier-Stokes frames hold
// 3*221*42 = 27 846 floats; else 4 B), 64 bytes of independent loads in flight per thread before the stores.
#include "engine.hpp"

namespace dyf {
namespace {

constexpr int GATHER_THREADS = 256;
constexpr int GATHER_BYTES_IN_FLIGHT = 64;  // per thread, before the first store
constexpr int GATHER_TABLE = 448;  // examples per launch (start frames travel as a kernel parameter -- 3.5 KB of the 4 KB
                                   // parameter space -- so there is no device-side index table and no copy to wait for)

struct GatherTable {
  long long first_frame[GATHER_TABLE];
};

template <typename V>
__global__ void __launch_bounds__(GATHER_THREADS) window_gather_kernel(const float* __restrict__ frames, float* __restrict__ out,
                                                                      long long run_elems, long long frame_elems,
                                                                      int example0, GatherTable table) {
  constexpr int VE = sizeof(V) / sizeof(float);
  constexpr int GATHER_UNROLL = GATHER_BYTES_IN_FLIGHT / sizeof(V) > 8 ? 8 : GATHER_BYTES_IN_FLIGHT / sizeof(V);
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

// Even frame sizes that are not a multiple of 4 floats (Navier-Stokes: 27 846): run starts are only 8-byte aligned, on
// either side, with a parity that changes from example to example.  One 8-byte head unit aligns the OUTPUT to 16 bytes;
// the body then stores 16 bytes per access and loads either 16 bytes (source parity matches) or two 8-byte halves (it does
// not); the branch is uniform over a CTA row.  Measured on B200: 8-byte-only copies reach 80 % of the copy peak, 16-byte ones
// 91 % (profiles/r01_widening.md).
__global__ void __launch_bounds__(GATHER_THREADS) window_gather_even_kernel(const float* __restrict__ frames, float* __restrict__ out,
                                                                           long long run_elems, long long frame_elems,
                                                                           int example0, GatherTable table) {
  constexpr int U = 4;
  const int e = blockIdx.y;
  const long long src_off = table.first_frame[e] * frame_elems;                 // floats, even
  const long long dst_off = (long long)(example0 + e) * run_elems;              // floats, even
  const float2* __restrict__ src = reinterpret_cast<const float2*>(frames + src_off);
  float2* __restrict__ dst = reinterpret_cast<float2*>(out + dst_off);
  long long units = run_elems / 2;                                              // 8-byte units in the run
  const long long tid = (long long)blockIdx.x * GATHER_THREADS + threadIdx.x;
  const long long stride = (long long)gridDim.x * GATHER_THREADS;
  if ((dst_off >> 1) & 1) {                                                     // head: align the output to 16 bytes
    if (tid == 0) dst[0] = __ldcs(src);
    ++src, ++dst, --units;
  }
  const bool src_aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const long long n4 = units / 2;
  float4* __restrict__ d4 = reinterpret_cast<float4*>(dst);
  long long i = tid;
  if (src_aligned) {
    const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
    for (; i + (U - 1) * stride < n4; i += U * stride) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldcs(s4 + i + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u) d4[i + u * stride] = v[u];
    }
    for (; i < n4; i += stride) d4[i] = __ldcs(s4 + i);
  } else {
    for (; i + (U - 1) * stride < n4; i += U * stride) {
      float2 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[u] = __ldcs(src + 2 * (i + u * stride));
        b[u] = __ldcs(src + 2 * (i + u * stride) + 1);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) d4[i + u * stride] = make_float4(a[u].x, a[u].y, b[u].x, b[u].y);
    }
    for (; i < n4; i += stride) {
      const float2 a = __ldcs(src + 2 * i), b = __ldcs(src + 2 * i + 1);
      d4[i] = make_float4(a.x, a.y, b.x, b.y);
    }
  }
  if ((units & 1) && tid == 0) dst[units - 1] = __ldcs(src + units - 1);        // tail unit
}

template <typename V, bool EVEN_MIXED = false>
int launch_gather(const float* frames, float* out, long long run_elems, long long frame_elems, const int64_t* first, int batch,
                  cudaStream_t s) {
  const long long per_block = (long long)GATHER_THREADS * GATHER_BYTES_IN_FLIGHT / sizeof(float);
  long long bx = cdiv(run_elems, per_block);
  if (bx > 4096) bx = 4096;  // grid-stride beyond that; 64 examples x 4096 CTAs is already >> 148 SMs x resident CTAs
  for (int e0 = 0; e0 < batch; e0 += GATHER_TABLE) {
    const int n = batch - e0 < GATHER_TABLE ? batch - e0 : GATHER_TABLE;
    GatherTable t;
    for (int i = 0; i < n; ++i) t.first_frame[i] = first[e0 + i];
    ProfScope prof(s, KC_PACK, 0.0, 8.0 * (double)run_elems * n);
    if (EVEN_MIXED)
      window_gather_even_kernel<<<dim3((unsigned)bx, (unsigned)n), GATHER_THREADS, 0, s>>>(frames, out, run_elems, frame_elems, e0, t);
    else
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
  if (frame_elems % 2 == 0 && a % 16 == 0)
    return launch_gather<float4, true>(frames, out, run, frame_elems, first_frame_host, batch, s);
  if (frame_elems % 2 == 0 && a % 8 == 0) return launch_gather<float2>(frames, out, run, frame_elems, first_frame_host, batch, s);
  return launch_gather<float>(frames, out, run, frame_elems, first_frame_host, batch, s);
}

}  // extern "C"
