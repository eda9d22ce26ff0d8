// Micro-benchmark of the per-SM issue rates that bound the resampling kernels on B200 (fp64 FMA/ADD,
// fp64<->fp32/int conversions, fp32 FMA, shared-memory gathers).  Prints results per clk per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_pipes tools/ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CH = 8;

template <int OP>
__global__ void k(float* out, float seed) {
  float f[CH]; double d[CH]; int n[CH];
  for (int c = 0; c < CH; ++c) { f[c] = seed + c + threadIdx.x * 1e-3f; d[c] = f[c]; n[c] = c; }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (OP == 0) f[c] = fmaf(f[c], 1.0001f, 0.5f);
      if (OP == 1) d[c] = fma(d[c], 1.0001, 0.5);
      if (OP == 2) d[c] = d[c] + 1.5;
      if (OP == 3) { f[c] = (float)d[c]; d[c] += f[c]; }               // F2F.F32.F64 (+ DADD w/ F2F.F64.F32)
      if (OP == 4) { n[c] = __double2int_rd(d[c]); d[c] += 0.25 + (n[c] & 1); }  // F2I.F64 + I2F
      if (OP == 5) { f[c] = floorf(f[c]) * 0.5f + 1.25f; }            // FRND.FLOOR + FFMA
      if (OP == 6) { n[c] = __float2int_rd(f[c]); f[c] = f[c] * 0.999f + (float)(n[c] & 3); }  // F2I.F32+I2F
      if (OP == 7) { d[c] = __dadd_rd(d[c], 4503599627370496.0) - 4503599627370496.0 + 0.3; }  // magic floor
    }
  }
  float s = 0; for (int c = 0; c < CH; ++c) s += f[c] + (float)d[c] + n[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void lds_gather(float* out, int stride) {
  __shared__ float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  int idx = (threadIdx.x * stride) & 8191; float s = 0;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) { s += sm[(idx + c * 37) & 8191]; }
    idx = (idx + 64 + (int)s % 2) & 8191;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name, int ops_per_iter, float* out, int sms, double mhz) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int blocks = sms * 4, threads = 512;
  k<OP><<<blocks, threads>>>(out, 1.f); cudaDeviceSynchronize();
  cudaEventRecord(a); k<OP><<<blocks, threads>>>(out, 1.f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)blocks * threads * ITERS * CH * ops_per_iter;
  printf("%-28s %8.3f ms  %7.2f lane-ops/clk/SM (assuming %.0f MHz)\n", name, ms, ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double mhz = khz / 1000.0;
  float* out; cudaMalloc(&out, sizeof(float) * sms * 4 * 512);
  printf("%s, %d SMs, max clock %.0f MHz\n", p.name, sms, mhz);
  run<0>("FFMA", 1, out, sms, mhz);
  run<1>("DFMA", 1, out, sms, mhz);
  run<2>("DADD", 1, out, sms, mhz);
  run<3>("F2F.f64->f32 + DADD(+cvt)", 1, out, sms, mhz);
  run<4>("F2I.f64 floor + DADD", 1, out, sms, mhz);
  run<5>("FRND.floor f32 + FFMA", 1, out, sms, mhz);
  run<6>("F2I.f32 floor + I2F + FFMA", 1, out, sms, mhz);
  run<7>("magic floor (3 DADD)", 1, out, sms, mhz);
  for (int stride : {1, 2, 17, 32, 33}) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    lds_gather<<<sms * 4, 512>>>(out, stride); cudaDeviceSynchronize();
    cudaEventRecord(a); lds_gather<<<sms * 4, 512>>>(out, stride); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)sms * 4 * 512 * ITERS * CH;
    printf("LDS.32 gather stride %-2d       %8.3f ms  %7.2f lane-loads/clk/SM\n", stride, ms, ops / (ms * 1e-3) / (mhz * 1e6) / sms);
  }
  return 0;
}
