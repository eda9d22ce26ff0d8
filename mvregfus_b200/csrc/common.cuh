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
// Conversions run on a 16-lane/clk pipe on B200 (profiles/r01/ubench_pipes_b200.txt: F2I/I2F/F2F ~16
// lanes/clk/SM vs 128 for FFMA), so the hot loops use exponent-alignment ("magic number") arithmetic on the
// FMA/ALU pipes instead of cvt instructions wherever the value range allows it.
constexpr float MAGIC23 = 8388608.0f;            // 2^23
constexpr float MAGIC32 = 12582912.0f;           // 1.5 * 2^23
constexpr double MAGIC64 = 6755399441055744.0;   // 1.5 * 2^52

// exact uint16 -> float without I2F
__device__ __forceinline__ float u16_to_float(uint32_t v) { return __uint_as_float(0x4B000000u | v) - MAGIC23; }

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<uint16_t>(uint16_t v) { return u16_to_float(v); }

// integer part of p (0 <= p < 2^23), truncated toward zero, without F2I
__device__ __forceinline__ uint32_t trunc_pos(float p) { return __float_as_uint(__fadd_rz(p, MAGIC23)) & 0x7FFFFFu; }

// Round like scipy's integer output path: trunc(t + 0.5) for t > 0, clipped (ni_interpolation.c,
// CASE_INTERP_OUT_UINT).  Round-toward-zero adds so x.4999999 never reaches the next integer.
__device__ __forceinline__ uint32_t round_half_up_u16(float t) {
  if (!(t > 0.f)) return 0u;
  return min(trunc_pos(__fadd_rz(fminf(t, 70000.f), 0.5f)), 65535u);
}

// sitk.Clamp to uint16: clamp then truncate toward zero
__device__ __forceinline__ uint32_t clamp_trunc_u16(float t) { return trunc_pos(fminf(fmaxf(t, 0.f), 65535.f)); }

// uint16(w * float64(I)) of the reference (mv:4907), exact: the fp32 product rounded toward zero has the
// same integer part as the exact product; the uint16 cast wraps modulo 2^16.  Valid for 0 <= w*I < 2^23.
__device__ __forceinline__ uint32_t trunc_product_u16(float w, float v) { return trunc_pos(__fmul_rz(w, v)) & 0xFFFFu; }

}  // namespace mvf
