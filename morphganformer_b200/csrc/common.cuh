// common.cuh -- shared helpers for libmgf_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/mgf.h"

namespace mgf {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
unsigned int* overflow_flag();   // device word that fp16-forward stores OR a 1 into when a value left the fp16 range (runtime.cu)
bool fwd_f16();   // true when forward activations / forward GEMM operands are IEEE fp16 (mgf_set_forward_dtype)

#define MGF_FAIL(code, ...) do { ::mgf::set_error(__VA_ARGS__); return (code); } while (0)
#define MGF_CHECK_LAUNCH(name) do { cudaError_t e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) { ::mgf::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_)); return (int)e_; } \
    ::mgf::count_launch(); } while (0)

static inline int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

template <class T> struct Cvt;
template <> struct Cvt<float> {
  __device__ __forceinline__ static float to(float v) { return v; }
  __device__ __forceinline__ static float from(float v) { return v; }
};
template <> struct Cvt<double> {
  __device__ __forceinline__ static double to(double v) { return v; }
  __device__ __forceinline__ static double from(double v) { return v; }
};
template <> struct Cvt<__nv_bfloat16> {
  __device__ __forceinline__ static float to(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static __nv_bfloat16 from(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Cvt<__half> {
  __device__ __forceinline__ static float to(__half v) { return __half2float(v); }
  __device__ __forceinline__ static __half from(float v) { return __float2half_rn(v); }
};
template <class T> struct Acc { typedef float type; };
template <> struct Acc<double> { typedef double type; };

// 16-byte vector of T
template <class T> struct alignas(16) Vec16 { T v[16 / sizeof(T)]; static constexpr int N = 16 / sizeof(T); };

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

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  // bf16 is the upper half of an fp32: one shift and one mask (the library intrinsic compiles to PRMT + shift for the upper element,
  // three instructions per pair -- visible in the issue-bound FIR / epilogue loops)
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// 16-bit storage with a run-time element type: f16 = IEEE half (forward activations/operands when the library is in fp16-forward
// mode: same tensor-core rate, 8x finer rounding than bf16), otherwise bfloat16 (always used for gradients: range).
// fp16 packing saturates to +-65504 in the conversion itself (one F2FP.SATFINITE instead of four clamps + the conversion)
__device__ __forceinline__ uint32_t pack_f16_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));     // first source -> upper half
  return r;
}
// fp16 overflow guard: kernels that store forward activations keep a running max |v| of what they pack (one FMNMX per value) and raise
// the library's device flag if it left the fp16 range (the store itself saturates).  Read with mgf_fp16_overflow_read.
constexpr float F16_MAX = 65504.f;
// running max |v| that PROPAGATES NaN (plain fmaxf drops it, and inf * 0 further down a saturated network is NaN)
__device__ __forceinline__ float ovf_max(float mx, float v) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(mx), "f"(fabsf(v)));
  return r;
}
__device__ __forceinline__ void ovf_commit(unsigned int* flag, float mx) { if (flag && !(mx <= F16_MAX)) atomicOr(flag, 1u); }
__device__ __forceinline__ uint32_t pack16(float a, float b, bool f16) {
  return f16 ? pack_f16_sat(a, b) : pack_bf16(a, b);
}
__device__ __forceinline__ float2 unpack16(uint32_t u, bool f16) {
  if (f16) { __half2 t = *reinterpret_cast<__half2*>(&u); return __half22float2(t); }
  return unpack_bf16(u);
}
__device__ __forceinline__ float load16(const void* p, long long i, bool f16) {
  return f16 ? __half2float(reinterpret_cast<const __half*>(p)[i]) : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

}  // namespace mgf
