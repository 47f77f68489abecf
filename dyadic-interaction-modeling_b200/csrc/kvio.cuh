// K/V cache row I/O (fp32 or bf16 head rows, one 128-bit vector per lane) and cp.async helpers shared by the decode-attention
// kernels (attention.cu) and the persistent decode kernel (decode_mk.cu).
#pragma once
#include "common.cuh"

namespace dimb {

template <bool BF16>
struct KvIo;
template <>
struct KvIo<false> {
  typedef float T;
  static constexpr int EPL = 4;                       // elements per lane
  static __device__ __forceinline__ float dot(const uint4& r, const float* q) {
    return fmaf(q[3], __uint_as_float(r.w), fmaf(q[2], __uint_as_float(r.z), fmaf(q[1], __uint_as_float(r.y), q[0] * __uint_as_float(r.x))));
  }
  static __device__ __forceinline__ void axpy(const uint4& r, float a, float* acc) {
    acc[0] = fmaf(a, __uint_as_float(r.x), acc[0]); acc[1] = fmaf(a, __uint_as_float(r.y), acc[1]);
    acc[2] = fmaf(a, __uint_as_float(r.z), acc[2]); acc[3] = fmaf(a, __uint_as_float(r.w), acc[3]);
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct KvIo<true> {
  typedef __nv_bfloat16 T;
  static constexpr int EPL = 8;
  // a bf16 is the high half of the fp32 with the same value: element 2i = word << 16, element 2i+1 = word & 0xffff0000
  static __device__ __forceinline__ float dot(const uint4& r, const float* q) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      d = fmaf(q[2 * i], __uint_as_float(w[i] << 16), d);
      d = fmaf(q[2 * i + 1], __uint_as_float(w[i] & 0xffff0000u), d);
    }
    return d;
  }
  static __device__ __forceinline__ void axpy(const uint4& r, float a, float* acc) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = fmaf(a, __uint_as_float(w[i] << 16), acc[2 * i]);
      acc[2 * i + 1] = fmaf(a, __uint_as_float(w[i] & 0xffff0000u), acc[2 * i + 1]);
    }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16_hint(uint32_t dst_smem, const void* src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace dimb
