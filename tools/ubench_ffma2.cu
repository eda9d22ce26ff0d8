// Micro-benchmark: packed FP32 (FFMA2, fma.rn.f32x2) issue/throughput on B200 vs scalar FFMA, alone and mixed with
// other pipes.  Prints FMA lane-ops per clk per SM (an FFMA2 counts as 2 per lane).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_ffma2 tools/ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, CH = 16;
struct Taps { float2 k[8]; };

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, float seed, const __grid_constant__ Taps T) {
  __shared__ float2 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(i, -i);
  __syncthreads();
  float2 f[CH]; float g[CH]; int n[CH];
  for (int c = 0; c < CH; ++c) { f[c] = make_float2(seed + c + threadIdx.x * 1e-3f, seed - c); g[c] = seed * c; n[c] = c + threadIdx.x; }
  const float2 a = make_float2(seed * 1.0001f, seed * 0.9999f);
  int idx = threadIdx.x;
  float kr[8]; float2 kp[8];
  for (int c = 0; c < 8; ++c) { kr[c] = sm[c + (threadIdx.x >> 9)].x * 1e-4f; kp[c] = make_float2(kr[c], kr[c] + 1.f); }
  if (OP == 9 || OP == 10 || OP == 11) {
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (OP == 9) f[c] = __ffma2_rn(make_float2(kr[c & 7], kr[c & 7]), f[(c + 1) % CH], f[c]);   // tap scalar broadcast, regs only
        if (OP == 10) f[c] = __ffma2_rn(kp[c & 7], f[(c + 1) % CH], f[c]);                           // tap pair in registers
        if (OP == 11) f[c] = __ffma2_rn(make_float2(kr[c & 7], kr[c & 7]), a, f[c]);                 // in-place accumulate, bcast tap
      }
    }
  }
  for (int i = 0; i < ITERS && OP < 9; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (OP == 0) { g[c] = fmaf(g[c], a.x, a.y); }                                 // FFMA reg,reg,reg
      if (OP == 1) { f[c] = __ffma2_rn(f[c], a, f[(c + 1) % CH]); }                 // FFMA2 reg,reg,reg
      if (OP == 2) { f[c] = __ffma2_rn(T.k[c & 7], a, f[c]); }                      // FFMA2 reg,ureg,reg (acc in place)
      if (OP == 3) { f[c] = __ffma2_rn(T.k[c & 7], make_float2(a.x, a.x), f[c]); }  // broadcast form
      if (OP == 4) { f[c] = __ffma2_rn(T.k[c & 7], a, f[c]); g[c] = fmaf(g[c], a.x, a.y); }   // FFMA2 + FFMA
      if (OP == 5) { f[c] = __ffma2_rn(T.k[c & 7], a, f[c]); n[c] = (n[c] ^ (n[c] >> 3)) + c; }  // FFMA2 + 2 ALU
      if (OP == 6) { f[c] = __ffma2_rn(T.k[c & 7], a, f[c]); n[c] = n[c] * 5 + c; }  // FFMA2 + IMAD
      if (OP == 7) { g[c] = fmaf(g[c], T.k[c & 7].x, a.y); }                        // FFMA reg,ureg,reg
    }
    if (OP == 9 || OP == 10) {   // taps in vector registers (loaded from shared memory once), scalar-broadcast or pair form
      // (handled below: taps hoisted out of the loop)
    }
    if (OP == 8) {   // FFMA2 fed from shared memory: 1 LDS.64 per 4 FFMA2
#pragma unroll
      for (int c = 0; c < CH; c += 4) {
        const float2 v = sm[(idx + c * 33) & 2047];
        f[c] = __ffma2_rn(T.k[0], v, f[c]); f[c + 1] = __ffma2_rn(T.k[1], v, f[c + 1]);
        f[c + 2] = __ffma2_rn(T.k[2], v, f[c + 2]); f[c + 3] = __ffma2_rn(T.k[3], v, f[c + 3]);
      }
      idx += 7;
    }
  }
  float s = 0; for (int c = 0; c < CH; ++c) s += f[c].x + f[c].y + g[c] + n[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name, double fma_per_item, float* out, int sms, double mhz, int threads) {
  Taps T; for (int i = 0; i < 8; ++i) T.k[i] = make_float2(1.f + i * 1e-4f, 1.f - i * 1e-4f);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int blocks = sms * 2;
  k<OP><<<blocks, threads>>>(out, 1.f, T); cudaDeviceSynchronize();
  cudaEventRecord(a); k<OP><<<blocks, threads>>>(out, 1.f, T); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double items = (double)blocks * threads * ITERS * CH;
  printf("%-34s thr/SM %4d %8.3f ms  %7.2f FMA lane-ops/clk/SM  %6.2f warp-items/clk/SM\n", name, threads * 2, ms,
         items * fma_per_item / (ms * 1e-3) / (mhz * 1e6) / sms, items / 32 / (ms * 1e-3) / (mhz * 1e6) / sms);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double mhz = khz / 1000.0;
  float* out; cudaMalloc(&out, sizeof(float) * sms * 2 * 512);
  printf("%s, %d SMs, max clock %.0f MHz\n", p.name, sms, mhz);
  for (int threads : {512, 192}) {
    run<0>("FFMA r,r,r", 1, out, sms, mhz, threads);
    run<7>("FFMA r,ur,r", 1, out, sms, mhz, threads);
    run<1>("FFMA2 r,r,r", 2, out, sms, mhz, threads);
    run<2>("FFMA2 r,ur,r", 2, out, sms, mhz, threads);
    run<3>("FFMA2 r.bcast,ur,r", 2, out, sms, mhz, threads);
    run<4>("FFMA2 + FFMA", 3, out, sms, mhz, threads);
    run<5>("FFMA2 + 2 ALU", 2, out, sms, mhz, threads);
    run<6>("FFMA2 + IMAD", 2, out, sms, mhz, threads);
    run<8>("FFMA2 x4 per LDS.64", 2, out, sms, mhz, threads);
    run<9>("FFMA2 r.bcast(tap),r,r", 2, out, sms, mhz, threads);
    run<10>("FFMA2 r(tap pair),r,r", 2, out, sms, mhz, threads);
    run<11>("FFMA2 r.bcast(tap),r,r acc", 2, out, sms, mhz, threads);
  }
  return 0;
}
