// hg_common.cuh -- shared device/host helpers of libhygrid_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <mutex>
#include "../../include/hygrid_b200.h"

namespace hg {

// ---- host side: thread-local error string + launch counter (hg_api.cu) -----------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int finish_launch(const char* what);  // cudaGetLastError() -> return code (+ message), counts one launch

#define HG_REQUIRE(cond, code, ...)                \
  do {                                             \
    if (!(cond)) {                                 \
      ::hg::set_error(__VA_ARGS__);                \
      return (code);                               \
    }                                              \
  } while (0)

// Opt-in dynamic shared memory of one kernel instantiation: the attribute is per (function, device) and
// shared by every host thread (the autograd engine launches the backward kernels from its own thread), so
// the high-water mark is process-wide and only ever raised.
struct SmemReservation {
  std::mutex m;
  size_t bytes[64] = {};
  template <typename K> cudaError_t reserve(K kern, size_t want) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(m);
    size_t& have = bytes[dev & 63];
    if (want <= have) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e == cudaSuccess) have = want; else cudaGetLastError();
    return e;
  }
};

static inline cudaStream_t as_stream(hg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

static inline int dtype_size(int dt) {
  switch (dt) {
    case HG_U8: return 1;
    case HG_I16: case HG_U16: case HG_BF16: return 2;
    case HG_I32: case HG_F32: return 4;
    case HG_I64: case HG_F64: return 8;
    default: return 0;
  }
}

// ---- device side ------------------------------------------------------------------------
// Reference arithmetic is numpy/torch float64 evaluated one operation at a time: the *_rn
// intrinsics are never contracted into FMAs by nvcc, which is what makes HG_MATH_EXACT
// bit-identical to the reference.
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fadd_rn(a, -b); }

// ``astype(int)`` / ``.to(int)``: truncation toward zero.
__device__ __forceinline__ int trunc_i32(double v) { return __double2int_rz(v); }
__device__ __forceinline__ int trunc_i32(float v) { return __float2int_rz(v); }
// ``((v) / 2).astype(int)``: true division then truncation toward zero == C integer division.
__device__ __forceinline__ int trunc_half(int v) { return v / 2; }

template <typename T> __device__ __forceinline__ double to_f64(T v) { return (double)v; }
template <> __device__ __forceinline__ double to_f64<__nv_bfloat16>(__nv_bfloat16 v) { return (double)__bfloat162float(v); }
template <typename T> __device__ __forceinline__ float to_f32(T v) { return (float)v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v) { return (T)v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T> __device__ __forceinline__ T from_f64(double v) { return (T)v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f64<__nv_bfloat16>(double v) { return __float2bfloat16_rn((float)v); }

// streaming (write-once) stores: keep L2 for the gathered source rows
template <typename T> __device__ __forceinline__ void st_stream(T* p, T v) { __stcs(p, v); }
template <> __device__ __forceinline__ void st_stream<__nv_bfloat16>(__nv_bfloat16* p, __nv_bfloat16 v) { *p = v; }
template <> __device__ __forceinline__ void st_stream<uint8_t>(uint8_t* p, uint8_t v) { *p = v; }
template <> __device__ __forceinline__ void st_stream<uint16_t>(uint16_t* p, uint16_t v) { *p = v; }
template <> __device__ __forceinline__ void st_stream<int8_t>(int8_t* p, int8_t v) { *p = v; }

}  // namespace hg
