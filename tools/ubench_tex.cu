// Micro-benchmark: point-sampled 3-D texture fetches (tex3D, unnormalised coords, border mode) vs surf3Dread,
// 8 taps per output voxel of a 512^3 float array, rotated access pattern.  nvcc -arch sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void k(cudaTextureObject_t tex, cudaSurfaceObject_t surf, float* out, int n, float c, float s) {
  // output voxel (z,y,x); source coordinate = rotation about y by angle (c,s) around the centre
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), z0 = blockIdx.z * 8;
  float acc = 0.f;
  const float h = 0.5f * (n - 1);
  for (int k = 0; k < 8; ++k) {
    const float zz = z0 + k - h, xx = x - h;
    const float uz = c * zz - s * xx + h, ux = s * zz + c * xx + h, uy = (float)y;
    const float fz = floorf(uz), fy = floorf(uy), fx = floorf(ux);
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float oz = fz + (t >> 2), oy = fy + ((t >> 1) & 1), ox = fx + (t & 1);
      if (MODE == 0) v += tex3D<float>(tex, ox + 0.5f, oy + 0.5f, oz + 0.5f);
      else { float r; surf3Dread(&r, surf, (int)ox * 4, (int)oy, (int)oz, cudaBoundaryModeZero); v += r; }
    }
    acc += v * (uz - fz);
  }
  out[((size_t)(z0)*n + y) * n + x] = acc;
}

int main() {
  const int n = 512;
  cudaArray_t arr;
  cudaChannelFormatDesc d = cudaCreateChannelDesc<float>();
  CK(cudaMalloc3DArray(&arr, &d, make_cudaExtent(n, n, n), cudaArraySurfaceLoadStore));
  float* lin; CK(cudaMalloc(&lin, sizeof(float) * n * n * n)); CK(cudaMemset(lin, 0, sizeof(float) * n * n * n));
  cudaMemcpy3DParms p = {};
  p.srcPtr = make_cudaPitchedPtr(lin, n * 4, n, n); p.dstArray = arr; p.extent = make_cudaExtent(n, n, n); p.kind = cudaMemcpyDeviceToDevice;
  CK(cudaMemcpy3D(&p));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
  td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  cudaSurfaceObject_t surf; CK(cudaCreateSurfaceObject(&surf, &rd));
  float* out; CK(cudaMalloc(&out, sizeof(float) * n * n * n));
  dim3 grid(n / 32, n / 8, n / 8);
  const float angles[3] = {0.f, 0.6f, 1.5708f};
  for (int mode = 0; mode < 2; ++mode)
    for (float a : angles) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, 256>>>(tex, surf, out, n, cosf(a), sinf(a)); else k<1><<<grid, 256>>>(tex, surf, out, n, cosf(a), sinf(a));
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double vv = (double)n * n * n;
      printf("%s angle %.2f rad: %.3f ms, %.1f G voxel-views/s, %.2f taps/clk/SM @1965MHz\n", mode == 0 ? "tex3D point" : "surf3Dread ", a, ms,
             vv / ms / 1e6, vv * 8 / (ms * 1e-3) / 1.965e9 / 148);
    }
  return 0;
}
