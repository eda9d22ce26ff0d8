// Shared host/device helpers of libmvregfus_b200 (error state, launch accounting, small device math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mvregfus_b200.h"

namespace mvf {

// ---- thread-local error message + launch counter (C-ABI is re-entrant: fuse_block is called from a
// ---- thread pool in the reference, mv:3797) -------------------------------------------------------
char* error_buffer();              // 512 bytes, thread local
int64_t& launch_counter();         // thread local

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(error_buffer(), 512, fmt, a, b);
  return code;
}

#define MVF_CHECK_ARG(cond, msg)                                                   \
  do {                                                                             \
    if (!(cond)) return mvf::fail(MVF_ERR_INVALID, "%s: %s", __func__, msg);       \
  } while (0)

#define MVF_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess)                                                         \
      return mvf::fail(MVF_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));     \
  } while (0)

#define MVF_LAUNCHED()                                                             \
  do {                                                                             \
    mvf::launch_counter()++;                                                       \
    MVF_CUDA(cudaGetLastError());                                                  \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached per thread)
int sm_count();

// ---- device helpers -----------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<uint16_t>(uint16_t v) { return (float)v; }

// Round like scipy's integer output path: trunc(t + 0.5) for t > 0, clipped (ni_interpolation.c,
// CASE_INTERP_OUT_UINT).  Round-toward-zero add so x.4999999 never reaches the next integer.
__device__ __forceinline__ uint16_t round_half_up_u16(float t) {
  if (!(t > 0.f)) return 0;
  float r = __fadd_rz(t, 0.5f);
  r = fminf(r, 65535.f);
  return (uint16_t)__float2uint_rz(r);
}

// sitk.Clamp to uint16: clamp then truncate toward zero
__device__ __forceinline__ uint16_t clamp_trunc_u16(float t) {
  t = fminf(fmaxf(t, 0.f), 65535.f);
  return (uint16_t)__float2uint_rz(t);
}

// uint16(w * float64(I)) of the reference (mv:4907), exact: the fp32 product rounded toward zero has the
// same integer part as the exact product.
__device__ __forceinline__ uint32_t trunc_product_u16(float w, float v) {
  float p = __fmul_rz(w, v);
  // numpy's float64 -> uint16 cast of out-of-range values wraps modulo 2^16 on x86 (cvttsd2si + truncation)
  return ((uint32_t)__float2int_rz(p)) & 0xFFFFu;
}

}  // namespace mvf
