// Shared device/host helpers for libdimb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <string>
#include <atomic>

#include "../../include/dimb200.h"

namespace dimb {

// ---- error plumbing ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern std::atomic<uint64_t> g_launches;

#define DIM_CHECK_CUDA(expr)                                                                         \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::dimb::fail(DIM_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
  } while (0)

#define DIM_REQUIRE(cond, msg)                                                                       \
  do {                                                                                               \
    if (!(cond)) return ::dimb::fail(DIM_EINVAL, std::string(msg) + " [" #cond "]");                 \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define DIM_LAUNCHED()                                                                               \
  do {                                                                                               \
    ::dimb::g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
    cudaError_t _e = cudaGetLastError();                                                             \
    if (_e != cudaSuccess) return ::dimb::fail(DIM_ECUDA, std::string("launch: ") + cudaGetErrorString(_e)); \
  } while (0)

int ensure_device();   // DIM_OK when the current device is sm_100; DIM_ENODEVICE otherwise

// One-time per-DEVICE set-up guard (cudaFuncSetAttribute is a per-device property; a process may hold handles on several GPUs)
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    const bool f = !done[d];
    done[d] = true;
    return f;
  }
};

// ---- optional per-kernel-category timing (dim_profile_*): CUDA events around each launch on the launching stream ----
enum ProfCat {
  CAT_GEMM_TILED = 0, CAT_GEMM_SKINNY, CAT_CONV, CAT_LAYERNORM, CAT_INSTNORM, CAT_ATTN_PREFILL, CAT_ATTN_DECODE,
  CAT_VQ_ARGMIN, CAT_VQ_GATHER, CAT_SAMPLE, CAT_MISC, CAT_GEMM_TC, CAT_GEMM_TC_SKINNY, CAT_DECODE_MK, CAT_COUNT
};
extern bool g_prof_on;
void prof_begin(int cat, cudaStream_t s, double bytes, double flops);
void prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s;
  bool on;
  ProfScope(int cat, cudaStream_t st, double bytes, double flops) : s(st), on(g_prof_on) {
    if (on) prof_begin(cat, s, bytes, flops);
  }
  ~ProfScope() {
    if (on) prof_end(s);
  }
};

// ---- launches with programmatic dependent launch (PDL) ---------------------------------------------------------------
// The decode step is a chain of short dependent kernels.  Launching them with programmaticStreamSerialization lets kernel
// N+1 be scheduled (and run its prologue) while kernel N drains; every such kernel calls pdl_prologue() before it touches
// memory, which blocks until all prerequisite grids have completed and flushed.  Opt-in with DIM_PDL=1: on the B=256 decode
// loop it measured 6 % SLOWER than plain graph edges (profiles/r01_notes.md), so it is off by default.
extern bool g_pdl_on;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- device helpers ---------------------------------------------------------------------------------------------
// First statement of every kernel that may be launched with launch_k(): let the dependents be scheduled, then wait for the
// prerequisite grids (no-op when the launch carried no programmatic dependency).
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

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

__device__ __forceinline__ float act_apply(float x, int act, float slope) {
  switch (act) {
    case DIM_ACT_LEAKY:
      return x > 0.f ? x : x * slope;
    case DIM_ACT_GELU_TANH: {
      // x * 0.5 * (1 + tanh(sqrt(2/pi) * (x + 0.044715 x^3)))  -- utils/base_model_util.py:81-94
      float u = 0.7978845608028654f * (x + 0.044715f * (x * x * x));
      return x * (0.5f * (1.0f + tanhf(u)));
    }
    case DIM_ACT_GELU_ERF:
      return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
    default:
      return x;
  }
}

// Exact bf16 split of fp32 values into `planes` parts (x = h + m + l, round-to-nearest at each step; the remainders are
// exactly representable), written to plane p at dst + p*kp.  4 consecutive elements per call (8-byte stores).
__device__ __forceinline__ void store_planes4(__nv_bfloat16* dst, float4 x, int planes, int kp) {
  float v[4] = {x.x, x.y, x.z, x.w};
  for (int pl = 0; pl < planes; ++pl) {
    unsigned short h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 b = __float2bfloat16_rn(v[j]);
      h[j] = __bfloat16_as_ushort(b);
      v[j] -= __bfloat162float(b);
    }
    uint2 pk;
    pk.x = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
    pk.y = (uint32_t)h[2] | ((uint32_t)h[3] << 16);
    *reinterpret_cast<uint2*>(dst + (size_t)pl * kp) = pk;
  }
}
__device__ __forceinline__ void store_planes1(__nv_bfloat16* dst, float x, int planes, int kp) {
  for (int pl = 0; pl < planes; ++pl) {
    __nv_bfloat16 b = __float2bfloat16_rn(x);
    dst[(size_t)pl * kp] = b;
    x -= __bfloat162float(b);
  }
}

}  // namespace dimb
