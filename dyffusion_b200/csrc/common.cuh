// Shared device/host helpers of the dyffusion_b200 engine (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace dyf {

// ---------------------------------------------------------------- error plumbing (thread-local message, C ABI)
void set_error(const std::string& msg);
const char* get_error();
void count_launch(int n = 1);

// Optional per-launch CUDA-event bracketing (bench.py's roofline measurement): when enabled, every kernel launch
// is timed on its own stream and accumulated per kernel class.
enum KernelClass : int { KC_CONV_MMA = 0, KC_CONV_UMMA, KC_PACK, KC_UPSAMPLE, KC_GROUPNORM, KC_READOUT, KC_TIME,
                         KC_ELEMENTWISE, KC_ATTENTION, KC_CONV_UP, KC_COUNT };
struct ProfScope {
  cudaStream_t s;
  int idx;
  ProfScope(cudaStream_t stream, int klass, double flops = 0.0, double bytes = 0.0);
  ~ProfScope();
};

#define DYF_CUDA_OK(expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      ::dyf::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                            \
      return -2;                                                                                       \
    }                                                                                                  \
  } while (0)

#define DYF_LAUNCH_OK(what)                                                                            \
  do {                                                                                                 \
    cudaError_t _e = cudaGetLastError();                                                               \
    if (_e != cudaSuccess) {                                                                           \
      ::dyf::set_error(std::string("launch of ") + what + " failed: " + cudaGetErrorString(_e));       \
      return -2;                                                                                       \
    }                                                                                                  \
    ::dyf::count_launch();                                                                             \
  } while (0)

// ---------------------------------------------------------------- activations (fp32 math)
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SILU = 3, ACT_GELU = 4 };

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LEAKY: return v > 0.f ? v : 0.2f * v;                       // RELU_LEAK = 0.2 (unet_simple.py:10)
    case ACT_SILU: return v / (1.f + __expf(-v));
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));  // nn.GELU() = exact erf form
    default: return v;
  }
}

// ---------------------------------------------------------------- Philox4x32-7 counter RNG (dropout / noise)
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ b;
      c3 = lo0;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// Dropout of a run of 8 consecutive elements whose first flat element index is `elem0` (a multiple of 8).
// keep-bit j uses 16 random bits; P(keep) = 1 - thresh/65536.  The mask is a pure function of
// (seed, stream, site, element index), independent of the tiling of the kernel that applies it.
struct DropCfg {
  uint64_t seed;
  uint32_t stream_lo, stream_hi_site;  // counter words 2,3
  uint32_t thresh;                     // 0 = dropout off
  float scale;                         // 1/(1-p)
};

__host__ inline DropCfg make_drop(bool on, uint64_t seed, uint64_t stream, uint32_t site, float p) {
  DropCfg d;
  d.seed = seed;
  d.stream_lo = (uint32_t)stream;
  d.stream_hi_site = ((uint32_t)(stream >> 32) & 0xFFFFu) | (site << 16);
  if (!on || p <= 0.f) {
    d.thresh = 0;
    d.scale = 1.f;
  } else {
    double t = (double)p * 65536.0 + 0.5;
    d.thresh = t > 65535.0 ? 65535u : (uint32_t)t;
    d.scale = 1.f / (1.f - p);
  }
  return d;
}

__device__ __forceinline__ uint32_t drop_keep_bits8(const DropCfg& d, uint64_t elem0) {
  Philox ph(d.seed);
  uint64_t g = elem0 >> 3;
  uint4 r = ph((uint32_t)g, (uint32_t)(g >> 32), d.stream_lo, d.stream_hi_site);
  uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t u16 = (w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
    bits |= (u16 >= d.thresh ? 1u : 0u) << j;
  }
  return bits;
}

// ---------------------------------------------------------------- bf16 packing
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace dyf
