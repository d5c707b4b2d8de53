// Shared device/host helpers of the dyffusion_b200 engine (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace dyf {

// ---------------------------------------------------------------- 16-bit storage type of activations and GEMM operands
// fp16 by default: same tensor-core rate as bf16 (tcgen05 kind::f16 / mma.sync take either), but a 10-bit mantissa, i.e. 8x
// less rounding per stored activation / operand -- what keeps 60-530 chained network calls of a sampling trajectory inside
// the stated tolerance.  Conversions saturate (cvt.rn.satfinite), so the narrower exponent range cannot produce infinities;
// accumulation, epilogues, statistics and the sampler state are fp32 either way.  -DDYF_ACT_BF16 builds the bf16 variant.
#ifdef DYF_ACT_BF16
typedef __nv_bfloat16 act_t;
#define DYF_UMMA_FMT 1u                       /* tcgen05 instruction-descriptor a/b format: 1 = bf16 */
#define DYF_MMA_T "bf16"
#define DYF_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define DYF_ACT_NAME "bf16"
#else
typedef __half act_t;
#define DYF_UMMA_FMT 0u                       /* 0 = f16 */
#define DYF_MMA_T "f16"
#define DYF_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define DYF_ACT_NAME "fp16"
#endif

// ---------------------------------------------------------------- error plumbing (thread-local message, C ABI)
void set_error(const std::string& msg);
const char* get_error();
void count_launch(int n = 1);

// Optional per-launch CUDA-event bracketing (bench.py's roofline measurement): when enabled, every kernel launch
// is timed on its own stream and accumulated per kernel class.
enum KernelClass : int { KC_CONV_MMA = 0, KC_CONV_UMMA, KC_PACK, KC_UPSAMPLE, KC_GROUPNORM, KC_READOUT, KC_TIME,
                         KC_ELEMENTWISE, KC_ATTENTION, KC_CONV_UP, KC_CONV_FLAT, KC_COUNT };
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

// ---------------------------------------------------------------- programmatic dependent launch
// Kernels launched through DYF_LAUNCH_PDL may be scheduled while their predecessor on the stream is still draining: the block
// scheduler places their CTAs as SMs free up, and their prologue (barrier init, TMEM allocation, static weight loads)
// overlaps the predecessor's tail.  Rules every such kernel obeys: `pdl_trigger()` first thing (lets ITS successor do the same),
// and EVERY thread executes `pdl_wait()` -- which returns once all predecessor grids have completed and flushed -- before its
// first read or write of memory another kernel touches and before it exits (so completion stays transitive along the stream).
// DYF_DISABLE_PDL=1 launches without the attribute (the wait is then a no-op).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled(int family = 0);  // family: 0 conv_umma, 1 conv_up, 2 GroupNorm, 3 attention (DYF_PDL_MASK, default all)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int family, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled(family) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define DYF_LAUNCH_PDL(family, what, kern, grid, block, smem, stream, ...)                                   \
  do {                                                                                                 \
    cudaError_t _e = ::dyf::launch_pdl(family, kern, grid, block, smem, stream, __VA_ARGS__);                \
    if (_e != cudaSuccess) {                                                                           \
      ::dyf::set_error(std::string("launch of ") + what + " failed: " + cudaGetErrorString(_e));       \
      return -2;                                                                                       \
    }                                                                                                  \
    ::dyf::count_launch();                                                                             \
  } while (0)

// ---------------------------------------------------------------- activations (fp32 math)
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SILU = 3, ACT_GELU = 4 };

// erf to 1.5e-7 absolute (Abramowitz-Stegun 7.1.26): one reciprocal, one exp2 and five FMAs instead of erff's ~40
// instructions -- the GELU epilogue of the spring-mesh layers is issue-bound.  nn.GELU() is the exact erf form; the
// difference (<= 1e-7 |x|) is three orders below the fp16 rounding of the stored activation.
__device__ __forceinline__ float erf_fast(float z) {
  const float az = fabsf(z);
  const float t = __fdividef(1.f, fmaf(0.3275911f, az, 1.f));
  const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  return copysignf(1.f - poly * __expf(-az * az), z);
}

// nn.GELU() (erf form) as v * Phi(v) with the normal CDF's tail written as a power of two: 0.5 erfc(a / sqrt 2) = 2^P(a),
// a = min(|v|, 6), P a degree-7 polynomial (weighted Chebyshev fit of log2 of the tail, evaluated in fp32 Horner form):
// |GELU error| <= 7.4e-8, |Phi error| <= 2.0e-7 over the whole line (tests/micro/gelu_fit.py) -- 7 FMAs + one exp2 instead of
// the 20 instructions of the erf route; the GELU epilogues of the spring-mesh layers are issue-bound.
__device__ __forceinline__ float gelu_fast(float v) {
  const float a = fminf(fabsf(v), 6.f);
  float p = 4.206899575365242e-06f;
  p = fmaf(p, a, -1.211548806168139e-05f);
  p = fmaf(p, a, -0.0005781833897344768f);
  p = fmaf(p, a, 0.007674761116504669f);
  p = fmaf(p, a, -0.05295220762491226f);
  p = fmaf(p, a, -0.4590412676334381f);
  p = fmaf(p, a, -1.1511286497116089f);
  p = fmaf(p, a, -0.999999463558197f);
  const float e = exp2f(p);           // Phi(-|v|)   (-use_fast_math: ex2.approx)
  return v * (v > 0.f ? 1.f - e : e);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LEAKY: return v > 0.f ? v : 0.2f * v;                       // RELU_LEAK = 0.2 (unet_simple.py:10)
    case ACT_SILU: return v / (1.f + __expf(-v));
    case ACT_GELU: return gelu_fast(v);  // nn.GELU() = erf form
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

// Division by a run-time constant as multiply-high + shift (n, d < 2^31; Granlund-Montgomery round-up multiplier).  The
// work-item decoding of the persistent kernels runs in every warp for every item: a 32-bit division by a kernel parameter is
// ~35 instructions, and on short items (single-chunk 1x1 / stem layers) those divisions were most of the issue slots.
struct FastDiv {
  uint32_t d, mul, shr;
  FastDiv() = default;
  explicit FastDiv(uint32_t div) : d(div), mul(0), shr(0) {
    while ((1ull << shr) < div) ++shr;
    mul = (uint32_t)(((((1ull << shr) - div) << 32) / div) + 1);
  }
#ifdef __CUDACC__
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(n, mul) + n) >> shr; }
#endif
};

// Dropout of a run of 8 consecutive elements (first element index a multiple of 8).  keep-bit j uses 16 random bits;
// P(keep) = 1 - thresh/65536.  The mask is a pure function of (seed, logical call, site, GLOBAL row, element inside the
// row): a batch row b of a launch sequence that carries several logical calls of `group_rows` rows each is row
// r = b % group_rows of logical call j = b / group_rows (its own stream id, stream + j), and r is offset by `row_off`, the
// index of the shard's first row in the un-sharded job.  So neither the tiling of a kernel, nor the batching of logical
// calls, nor the number of ranks a job is split over changes which elements are dropped (a row-sharded
// ensemble must not repeat masks across ranks).
struct DropCfg {
  uint64_t seed;
  const uint64_t* seed_ptr;            // when set, the seed is read from device memory (CUDA-graph replays draw new masks)
  uint32_t stream_lo, stream_hi_site;  // counter words 2,3
  uint32_t thresh;                     // 0 = dropout off
  float scale;                         // 1/(1-p)
  uint32_t group_rows;                 // rows per logical call (>= 1)
  uint32_t row_off;                    // global index of batch row 0 of every logical call
};
// what the sampler / a forward passes down to the layers: the RNG coordinates of one launch sequence
struct RngCtx {
  bool on = false;
  uint64_t seed = 0, stream = 0;
  const uint64_t* seed_ptr = nullptr;
  uint32_t group_rows = 1, row_off = 0;
};

__host__ inline DropCfg make_drop(const RngCtx& c, uint32_t site, float p) {
  DropCfg d;
  d.seed = c.seed;
  d.seed_ptr = c.seed_ptr;
  d.stream_lo = (uint32_t)c.stream;
  d.stream_hi_site = ((uint32_t)(c.stream >> 32) & 0xFFFFu) | (site << 16);
  d.group_rows = c.group_rows ? c.group_rows : 1;
  d.row_off = c.row_off;
  if (!c.on || p <= 0.f) {
    d.thresh = 0;
    d.scale = 1.f;
  } else {
    double t = (double)p * 65536.0 + 0.5;
    d.thresh = t > 65535.0 ? 65535u : (uint32_t)t;
    d.scale = 1.f / (1.f - p);
  }
  return d;
}

__device__ __forceinline__ uint64_t rng_seed(uint64_t seed, const uint64_t* seed_ptr) { return seed_ptr ? __ldg(seed_ptr) : seed; }

// RNG coordinates of one batch row: stream word of its logical call and the element offset of its global row
struct DropRow { uint32_t stream; uint64_t base; };
__device__ __forceinline__ DropRow drop_row(const DropCfg& d, int batch_row, uint64_t elems_per_row) {
  const uint32_t j = (uint32_t)batch_row / d.group_rows, r = (uint32_t)batch_row - j * d.group_rows;
  return DropRow{d.stream_lo + j, (uint64_t)(r + d.row_off) * elems_per_row};
}
// keep bits of elements [e, e + 8) of that row (e a multiple of 8, elems_per_row a multiple of 8)
__device__ __forceinline__ uint32_t drop_keep_bits8(const DropCfg& d, const DropRow& dr, uint64_t e) {
  Philox ph(rng_seed(d.seed, d.seed_ptr));
  const uint64_t g = (dr.base + e) >> 3;
  uint4 r = ph((uint32_t)g, (uint32_t)(g >> 32), dr.stream, d.stream_hi_site);
  uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t u16 = (w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
    bits |= (u16 >= d.thresh ? 1u : 0u) << j;
  }
  return bits;
}

// the same masks applied in place: v[j] = keep_j ? v[j] * scale : 0 for elements [e, e + 8) -- each 16-bit field is compared
// and selected directly (no bit vector in between: 4 instead of 7 instructions per element in the issue-bound epilogues)
__device__ __forceinline__ void drop_apply8(const DropCfg& d, const DropRow& dr, uint64_t e, float* v) {
  Philox ph(rng_seed(d.seed, d.seed_ptr));
  const uint64_t g = (dr.base + e) >> 3;
  const uint4 r = ph((uint32_t)g, (uint32_t)(g >> 32), dr.stream, d.stream_hi_site);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = (w[j] & 0xFFFFu) >= d.thresh ? v[2 * j] * d.scale : 0.f;
    v[2 * j + 1] = (w[j] >> 16) >= d.thresh ? v[2 * j + 1] * d.scale : 0.f;
  }
}

// ---------------------------------------------------------------- 16-bit packing
#ifdef DYF_ACT_BF16
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_act2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__host__ __device__ __forceinline__ act_t f2act(float v) { return __float2bfloat16_rn(v); }
__host__ __device__ __forceinline__ float act2f(act_t v) { return __bfloat162float(v); }
#else
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {  // round to nearest, saturate to +-65504
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_act2(uint32_t u) {
  __half2 v = *reinterpret_cast<__half2*>(&u);
  return __half22float2(v);
}
__device__ __forceinline__ act_t f2act(float v) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}
__device__ __forceinline__ float act2f(act_t v) { return __half2float(v); }
#endif
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_act2(u.x), b = unpack_act2(u.y), c = unpack_act2(u.z), d = unpack_act2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]), pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
}

// 256-bit global store (sm_100: STG.E.256) of 16 consecutive 16-bit values to a 32-byte-aligned address: an epilogue thread
// that owns 32+ contiguous bytes writes whole 32-byte sectors per request instead of two half-sector requests
#ifdef __CUDACC__
__device__ __forceinline__ void st_global_256(void* ptr, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
#endif

#ifdef __CUDACC__
// ... and the matching 256-bit read-only load (LDG.E.256.CONSTANT), 32-byte-aligned address
__device__ __forceinline__ void ld_global_nc_256(const void* ptr, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(ptr));
}
#endif

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace dyf
