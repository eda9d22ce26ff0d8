// Minimal TMA probe: one 3-D box load per CTA into shared memory, copied back to global for checking.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// variant 0: 3-D tensor load, descriptor in the parameter block; 1: descriptor in global memory; 2: 1-D bulk copy (no descriptor)
template <typename T>
__global__ void probe(const __grid_constant__ CUtensorMap map, const CUtensorMap* gmap, const T* src, int variant, T* out, int nbox, int cx, int cy, int cz, int bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (variant == 2) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(smem_u32(sm)), "l"(src), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    } else {
      const CUtensorMap* m = variant == 1 ? gmap : &map;
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
                   "r"(smem_u32(sm)), "l"(m), "r"(smem_u32(&bar)), "r"(cx), "r"(cy), "r"(cz) : "memory");
    }
  }
  __syncthreads();
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  const T* s = reinterpret_cast<const T*>(sm);
  for (int i = threadIdx.x; i < nbox; i += blockDim.x) out[i] = s[i];
}

template <typename T>
int run(int variant, int nz, int ny, int nx, int bz, int by, int bx, int cz, int cy, int cx, CUtensorMapDataType dt) {
  std::vector<T> h((size_t)nz * ny * nx);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (T)(i % 60000);
  T *d, *o;
  cudaMalloc(&d, h.size() * sizeof(T));
  cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  const int nbox = bz * by * bx;
  cudaMalloc(&o, nbox * sizeof(T));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap map;
  cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
  cuuint64_t gstr[2] = {(cuuint64_t)nx * sizeof(T), (cuuint64_t)nx * ny * sizeof(T)};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((enc_t)fn)(&map, dt, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d  tensor %dx%dx%d box %dx%dx%d at (%d,%d,%d) elem %zu\n", (int)r, nz, ny, nx, bz, by, bx, cz, cy, cx, sizeof(T));
  if (r != CUDA_SUCCESS) return 1;
  cudaFuncSetAttribute(probe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, nbox * (int)sizeof(T) + 128);
  CUtensorMap* gmap;
  cudaMalloc(&gmap, sizeof(map));
  cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
  printf("  variant %d\n", variant);
  probe<T><<<1, 128, nbox * sizeof(T) + 128>>>(map, gmap, d, variant, o, nbox, cx, cy, cz, nbox * (int)sizeof(T));
  cudaError_t e = cudaDeviceSynchronize();
  printf("  kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<T> g(nbox);
  cudaMemcpy(g.data(), o, nbox * sizeof(T), cudaMemcpyDeviceToHost);
  long bad = 0;
  if (variant == 2) {
    for (int i = 0; i < nbox; ++i) bad += g[i] != h[i];
    printf("  mismatches (bulk): %ld of %d\n", bad, nbox);
    return bad != 0;
  }
  for (int z = 0; z < bz; ++z)
    for (int y = 0; y < by; ++y)
      for (int x = 0; x < bx; ++x) {
        int gz = cz + z, gy = cy + y, gx = cx + x;
        T want = 0;
        if (gz >= 0 && gz < nz && gy >= 0 && gy < ny && gx >= 0 && gx < nx) want = h[((size_t)gz * ny + gy) * nx + gx];
        if (g[(z * by + y) * bx + x] != want) ++bad;
      }
  printf("  mismatches: %ld of %d\n", bad, nbox);
  return bad != 0;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int shape = argc > 2 ? atoi(argv[2]) : 0;
  int rc = 0;
  if (shape == 0) {
    rc |= run<float>(variant, 40, 36, 48, 18, 11, 36, 3, 5, 7, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    rc |= run<float>(variant, 40, 36, 48, 18, 11, 36, -2, -3, -5, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    rc |= run<float>(variant, 40, 36, 48, 18, 11, 36, 30, 30, 30, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    rc |= run<uint16_t>(variant, 40, 36, 48, 18, 11, 40, 3, 5, 7, CU_TENSOR_MAP_DATA_TYPE_UINT16);
    rc |= run<uint16_t>(variant, 40, 36, 48, 35, 11, 24, -1, 5, 39, CU_TENSOR_MAP_DATA_TYPE_UINT16);
  } else if (shape == 1) {   // inner box = 128 bytes, small box
    rc |= run<float>(variant, 64, 64, 64, 4, 4, 32, 0, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  } else if (shape == 2) {   // 64-element inner box, unaligned start
    rc |= run<float>(variant, 64, 64, 64, 8, 8, 64, 1, 2, 3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  } else {                   // starts aligned to 16 bytes along x, arbitrary elsewhere, partly outside the tensor
    rc |= run<float>(variant, 40, 36, 48, 18, 11, 36, 3, 5, 8, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    rc |= run<float>(variant, 40, 36, 48, 18, 11, 36, -2, -3, -8, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    rc |= run<float>(variant, 40, 36, 48, 18, 11, 36, 30, 30, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    rc |= run<uint16_t>(variant, 40, 36, 48, 18, 11, 40, 3, 5, 8, CU_TENSOR_MAP_DATA_TYPE_UINT16);
    rc |= run<uint16_t>(variant, 40, 36, 48, 35, 11, 24, -1, 5, 40, CU_TENSOR_MAP_DATA_TYPE_UINT16);
  }
  return rc;
}
