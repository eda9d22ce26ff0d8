// cuFFT: 3-D R2C + C2R of a padded RL slab against batched 2-D (z, x) transforms with y as the batch (strided layout),
// the alternative for PSFs of views rotated about y (psf = a(y) (x) B(z, x)).   nvcc -O3 -arch=sm_100a tools/ubench_fft.cu -lcufft
#include <cufft.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { auto r_ = (x); if (r_ != 0) { printf("%s failed: %d\n", #x, (int)r_); return 1; } } while (0)
int main(int argc, char** argv) {
  int Lz = argc > 1 ? atoi(argv[1]) : 168, Ly = argc > 2 ? atoi(argv[2]) : 2100, Lx = argc > 3 ? atoi(argv[3]) : 2100, ny = argc > 4 ? atoi(argv[4]) : 2048;
  size_t nreal = (size_t)Lz * Ly * Lx, ncplx = (size_t)Lz * Ly * (Lx / 2 + 1);
  float* real; cufftComplex* cplx;
  CK(cudaMalloc(&real, nreal * 4)); CK(cudaMalloc(&cplx, ncplx * 8));
  CK(cudaMemset(real, 0, nreal * 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  {
    cufftHandle f, g; size_t w1, w2;
    CK(cufftCreate(&f)); CK(cufftCreate(&g));
    CK(cufftMakePlan3d(f, Lz, Ly, Lx, CUFFT_R2C, &w1)); CK(cufftMakePlan3d(g, Lz, Ly, Lx, CUFFT_C2R, &w2));
    for (int it = 0; it < 2; ++it) { CK(cufftExecR2C(f, real, cplx)); CK(cufftExecC2R(g, cplx, real)); }
    cudaEventRecord(a);
    for (int it = 0; it < 5; ++it) { CK(cufftExecR2C(f, real, cplx)); CK(cufftExecC2R(g, cplx, real)); }
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("3-D  %d x %d x %d: R2C + C2R %.3f ms (work areas %.2f / %.2f GB)\n", Lz, Ly, Lx, ms / 5, w1 / 1e9, w2 / 1e9);
    cufftDestroy(f); cufftDestroy(g);
  }
  {
    // element (y, z, x) at z * (ny * Lx) + y * Lx + x : 2-D transform over (z, x), batch y
    cufftHandle f, g; size_t w1, w2;
    int n[2] = {Lz, Lx};
    int inembed[2] = {Lz, ny * Lx}, onembed[2] = {Lz, ny * (Lx / 2 + 1)};
    CK(cufftCreate(&f)); CK(cufftCreate(&g));
    CK(cufftMakePlanMany(f, 2, n, inembed, 1, Lx, onembed, 1, Lx / 2 + 1, CUFFT_R2C, ny, &w1));
    CK(cufftMakePlanMany(g, 2, n, onembed, 1, Lx / 2 + 1, inembed, 1, Lx, CUFFT_C2R, ny, &w2));
    for (int it = 0; it < 2; ++it) { CK(cufftExecR2C(f, real, cplx)); CK(cufftExecC2R(g, cplx, real)); }
    cudaEventRecord(a);
    for (int it = 0; it < 5; ++it) { CK(cufftExecR2C(f, real, cplx)); CK(cufftExecC2R(g, cplx, real)); }
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("2-D (z, x) x %d rows, strided batch: R2C + C2R %.3f ms (work areas %.2f / %.2f GB)\n", ny, ms / 5, w1 / 1e9, w2 / 1e9);
    cufftDestroy(f); cufftDestroy(g);
  }
  return 0;
}
