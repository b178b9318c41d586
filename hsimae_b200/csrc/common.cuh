// Shared device/host helpers for the hsimae_b200 CUDA library (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "hsimae_b200 is written for sm_100a (B200) only"
#endif

namespace hsimae {

// ---------------------------------------------------------------------------
// error reporting across the C ABI: int status + thread-local message
// ---------------------------------------------------------------------------
enum Status : int {
  kOk = 0,
  kInvalidArgument = 1,
  kCudaError = 2,
  kUnsupported = 3,
  kInternal = 4,
};

void set_error(const char* fmt, ...);
const char* last_error();

#define HS_CHECK_CUDA(expr)                                                             \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::hsimae::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                          __FILE__, __LINE__);                                          \
      return ::hsimae::kCudaError;                                                      \
    }                                                                                   \
  } while (0)

// Programmatic dependent launch (see gemm_tc.cu): kernels launched through launch_pdl may begin while the previous
// kernel on the stream drains; they call pdl_wait() before touching global memory and pdl_trigger() when their own
// dependents may be scheduled.  HSIMAE_PDL=0 launches them as ordinary stream-ordered kernels.
// HSIMAE_PDL=0 (environment, default on) or hsimae_set_pdl() at run time (per-kernel timing: overlapped prologues make
// activity-profiler durations of consecutive kernels overlap)
bool pdl_enabled();
void set_pdl(bool on);
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class Kernel, class... Args>
inline cudaError_t launch_pdl(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  const bool pdl = pdl_enabled();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

void count_launch();
long long launch_count();

#define HS_CHECK_LAUNCH(name)                                                           \
  do {                                                                                  \
    ::hsimae::count_launch();                                                           \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      ::hsimae::set_error("launch of %s failed: %s (%s:%d)", name,                      \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                  \
      return ::hsimae::kCudaError;                                                      \
    }                                                                                   \
  } while (0)

#define HS_REQUIRE(cond, ...)                                                           \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      ::hsimae::set_error(__VA_ARGS__);                                                 \
      return ::hsimae::kInvalidArgument;                                                \
    }                                                                                   \
  } while (0)

#define HS_TRY(expr)                                                                    \
  do {                                                                                  \
    int _s = (expr);                                                                    \
    if (_s != 0) return _s;                                                             \
  } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------
// Row-group scaling (stochastic depth): factor index for a token row.
//   mode 0: none (scale==nullptr)
//   mode 1: spatial  -> group (b, t):  b*G + (r%K)/len_l,   G = len_t
//   mode 2: spectral -> group (b, l):  b*G + (r%K)%len_l,   G = len_l
//   mode 3: fusion   -> group b
// ---------------------------------------------------------------------------
struct RowScale {
  const float* scale;  // nullptr => 1.0
  int mode;
  int K;      // tokens per sample
  int len_l;  // inner length
  int G;      // groups per sample
};

__device__ __forceinline__ float row_scale(const RowScale& rs, int r) {
  if (rs.scale == nullptr) return 1.0f;
  int b = r / rs.K;
  int w = r - b * rs.K;
  int g = rs.mode == 1 ? (w / rs.len_l) : (rs.mode == 2 ? (w % rs.len_l) : 0);
  return __ldg(rs.scale + b * rs.G + g);
}

// ---------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

// 16-byte streaming accesses (activations are touched once per kernel)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// x / d and x % d for small runtime divisors in copy loops (x * d < 2^32): one multiply-high instead of the ~20-instruction
// integer division sequence (those sequences were among the hottest SASS of the attention kernels, profiles/r01g_stall_hotspots.md)
struct FastDiv {
  uint32_t inv, d;
  __device__ __forceinline__ explicit FastDiv(int div) : inv((uint32_t)((0x100000000ull + (uint32_t)div - 1) / (uint32_t)div)), d((uint32_t)div) {}
  __device__ __forceinline__ int div(int x) const { return d == 1u ? x : (int)__umulhi((uint32_t)x, inv); }   // inv wraps to 0 for d == 1
  __device__ __forceinline__ void divmod(int x, int& q, int& r) const { q = div(x); r = x - q * (int)d; }
};

__device__ __forceinline__ float silu_f(float a) { return __fdividef(a, 1.0f + __expf(-a)); }

}  // namespace hsimae
