// Multi-view RL, general (non-separable) PSF path: circular FFT convolution exactly as the reference does it
// (blur_with_ftpsfind, mvregfus/multiview.py:5654-5673), with cuFFT R2C/C2R and hand-written gather / spectrum
// multiply / fused-epilogue kernels.  The padded length L' per axis is the next 2^a 3^b 5^c 7^d >= n + 2R (the
// reference uses n + K + 1): the padded volume holds ext(p) for p in [0, n-1+R] and the wrap values ext(p - L')
// in its last R entries, zeros in between, so every output voxel sees the same inputs as in the reference.
#include <cufft.h>

#include "common.cuh"

namespace mvf {

struct FftPlan {
  int n[3];    // volume extents
  int L[3];    // padded extents
  cufftHandle r2c, c2r;
  size_t work_bytes;
};

static int good_size(int n) {
  for (int m = n;; ++m) {
    int r = m;
    for (int p : {2, 3, 5, 7})
      while (r % p == 0) r /= p;
    if (r == 1) return m;
  }
}

__device__ __forceinline__ int ext_index_fft(int t, int n, int K) {
  int i = t < 0 ? n - 3 - K - t : (t >= n ? 2 * n - 2 - t : t);
  return max(0, min(i, n - 1));
}

struct GatherArgs {
  const float* in;
  const int* zmap;  // extended z position t + rp -> plane of `in`
  int rp;
  int n[3], L[3], K[3], R[3];
  float* out;       // L[0]*L[1]*L[2]
};

// padded[p] = ext(p) for p <= n-1+R, ext(p-L) for p >= L-R, else 0 (per axis, separable index map)
// One CTA per (pz, py) row segment of FFT_XPT * 256 elements: every thread moves FFT_XPT elements (stride 256, all
// loads issued before the stores), the z / y parts of the index map are computed once per CTA row.
constexpr int FFT_XPT = 4;
__global__ void __launch_bounds__(256) fft_gather_kernel(const __grid_constant__ GatherArgs A) {
  const int py = blockIdx.y, pz = blockIdx.z;
  int srow[2];
  bool zero_row = false;
  const int pr[2] = {pz, py};
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    int t;
    if (pr[d] <= A.n[d] - 1 + A.R[d]) t = pr[d];
    else if (pr[d] >= A.L[d] - A.R[d]) t = pr[d] - A.L[d];
    else { zero_row = true; t = 0; }
    srow[d] = d == 0 ? A.zmap[t + A.rp] : ext_index_fft(t, A.n[d], A.K[d]);
  }
  const float* __restrict__ src = A.in + ((size_t)srow[0] * A.n[1] + srow[1]) * A.n[2];
  float* __restrict__ dst = A.out + ((size_t)pz * A.L[1] + py) * A.L[2];
  float v[FFT_XPT];
#pragma unroll
  for (int k = 0; k < FFT_XPT; ++k) {
    const int px = (blockIdx.x * FFT_XPT + k) * 256 + threadIdx.x;
    v[k] = 0.f;
    if (px < A.L[2] && !zero_row) {
      const bool low = px <= A.n[2] - 1 + A.R[2], high = px >= A.L[2] - A.R[2];
      if (low || high) v[k] = __ldg(src + ext_index_fft(low ? px : px - A.L[2], A.n[2], A.K[2]));
    }
  }
#pragma unroll
  for (int k = 0; k < FFT_XPT; ++k) {
    const int px = (blockIdx.x * FFT_XPT + k) * 256 + threadIdx.x;
    if (px < A.L[2]) dst[px] = v[k];
  }
}

__global__ void __launch_bounds__(256) fft_place_psf_kernel(const float* psf, int kz, int ky, int kx, float* out, int Lz, int Ly, int Lx) {
  const size_t n = (size_t)Lz * Ly * Lx;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int x = (int)(i % Lx), y = (int)((i / Lx) % Ly), z = (int)(i / ((size_t)Lx * Ly));
    out[i] = (z < kz && y < ky && x < kx) ? psf[(z * ky + y) * kx + x] : 0.f;  // np.pad(psf, [0, n+1]) (mv:5649)
  }
}

__global__ void __launch_bounds__(256) fft_multiply_kernel(cufftComplex* a, const cufftComplex* b, size_t n, float scale) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const cufftComplex x = a[i], y = b[i];
    a[i] = make_cuFloatComplex((x.x * y.x - x.y * y.y) * scale, (x.x * y.y + x.y * y.x) * scale);
  }
}

struct FftEpiArgs {
  const float* conv;  // L[0]*L[1]*L[2], circular convolution result
  int n[3], L[3], R[3];
  int mode;
  const void* view; int view_u16;
  const float* weights;
  float* out;
  int first, last;
  float* est; long long est_off;
  double* sum;
};

// out[i] = conv[i + R] cropped (mv:5669), clamped at 0, followed by the same epilogues as the separable kernel
__global__ void __launch_bounds__(256) fft_epilogue_kernel(const __grid_constant__ FftEpiArgs A) {
  const int y = blockIdx.y, z = blockIdx.z;
  double local_sum = 0.0;
  const float* __restrict__ crow = A.conv + ((size_t)(z + A.R[0]) * A.L[1] + (y + A.R[1])) * A.L[2] + A.R[2];
  const size_t row = ((size_t)z * A.n[1] + y) * A.n[2];
  // FFT_XPT elements per thread (stride 256); every operand of all of them is requested before the arithmetic
  float c[FFT_XPT], w[FFT_XPT], I[FFT_XPT], o[FFT_XPT], e[FFT_XPT];
#pragma unroll
  for (int k = 0; k < FFT_XPT; ++k) {
    const int x = (blockIdx.x * FFT_XPT + k) * 256 + threadIdx.x;
    c[k] = w[k] = I[k] = o[k] = e[k] = 0.f;
    if (x < A.n[2]) {
      const size_t idx = row + x;
      c[k] = __ldg(crow + x);
      if (A.mode != 0) w[k] = __ldg(A.weights + idx);
      if (A.mode == 1) I[k] = A.view_u16 ? u16_to_float(__ldg(reinterpret_cast<const uint16_t*>(A.view) + idx))
                                         : __ldg(reinterpret_cast<const float*>(A.view) + idx);
      if (A.mode == 2 && !A.first) o[k] = A.out[idx];
      if (A.mode == 2 && A.last) e[k] = A.est[A.est_off + idx];
    }
  }
#pragma unroll
  for (int k = 0; k < FFT_XPT; ++k) {
    const int x = (blockIdx.x * FFT_XPT + k) * 256 + threadIdx.x;
    if (x >= A.n[2]) continue;
    const size_t idx = row + x;
    const float blur = fmaxf(c[k], 0.f);
    if (A.mode == 0) {
      A.out[idx] = blur;
    } else if (A.mode == 1) {
      const float ratio = __fdiv_rn(I[k], blur + 1e-6f);
      A.out[idx] = w[k] > 1e-5f ? ratio : 0.f;
    } else {
      const float ow = blur * w[k];
      const float corr = A.first ? ow : o[k] + ow;
      if (A.last) {
        const float v = fminf(fmaxf(e[k] * corr, 0.f), 65535.f);
        A.est[A.est_off + idx] = v;
        local_sum += (double)v;
      } else {
        A.out[idx] = corr;
      }
    }
  }
  if (A.mode == 2 && A.last && A.sum) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) local_sum += __shfl_xor_sync(0xffffffffu, local_sum, s);
    if ((threadIdx.x & 31) == 0 && local_sum != 0.0) atomicAdd(A.sum, local_sum);
  }
}

#define MVF_CUFFT(call)                                                                   \
  do {                                                                                    \
    cufftResult r_ = (call);                                                              \
    if (r_ != CUFFT_SUCCESS) {                                                            \
      snprintf(error_buffer(), 512, "%s failed with cufftResult %d", #call, (int)r_);     \
      return MVF_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

}  // namespace mvf

using namespace mvf;

extern "C" int mvf_rl_fft_padded_dims(const int32_t dims[3], const int32_t ksize[3], int32_t padded[3]) {
  MVF_CHECK_ARG(dims && ksize && padded, "NULL argument");
  for (int d = 0; d < 3; ++d) {
    MVF_CHECK_ARG(ksize[d] >= 1 && (ksize[d] & 1), "PSF sizes must be odd");
    padded[d] = good_size(dims[d] + (ksize[d] - 1));
  }
  return MVF_OK;
}

extern "C" int mvf_rl_fft_plan_create(const int32_t dims[3], const int32_t ksize[3], void* workspace_or_null,
                                      size_t* workspace_bytes, void** plan_out) {
  MVF_CHECK_ARG(dims && ksize && plan_out && workspace_bytes, "NULL argument");
  // the closed-form y/x index map (wrap into the far reflected tail) needs n >= K + 3; z goes through the caller's plane map
  for (int d = 1; d < 3; ++d)
    if (dims[d] < ksize[d] + 3)
      return fail(MVF_ERR_UNSUPPORTED, "%s%s", "rl: volume extent smaller than PSF size + 3 (the reference's reflect padding needs it)");
  FftPlan* p = new FftPlan();
  for (int d = 0; d < 3; ++d) {
    p->n[d] = dims[d];
    p->L[d] = good_size(dims[d] + (ksize[d] - 1));
  }
  size_t w1 = 0, w2 = 0;
  MVF_CUFFT(cufftCreate(&p->r2c));
  MVF_CUFFT(cufftCreate(&p->c2r));
  MVF_CUFFT(cufftSetAutoAllocation(p->r2c, 0));
  MVF_CUFFT(cufftSetAutoAllocation(p->c2r, 0));
  MVF_CUFFT(cufftMakePlan3d(p->r2c, p->L[0], p->L[1], p->L[2], CUFFT_R2C, &w1));
  MVF_CUFFT(cufftMakePlan3d(p->c2r, p->L[0], p->L[1], p->L[2], CUFFT_C2R, &w2));
  p->work_bytes = w1 > w2 ? w1 : w2;
  *workspace_bytes = p->work_bytes;
  if (workspace_or_null) {
    MVF_CUFFT(cufftSetWorkArea(p->r2c, workspace_or_null));
    MVF_CUFFT(cufftSetWorkArea(p->c2r, workspace_or_null));
  }
  *plan_out = p;
  return MVF_OK;
}

extern "C" int mvf_rl_fft_plan_set_workspace(void* plan, void* workspace) {
  MVF_CHECK_ARG(plan && workspace, "NULL argument");
  FftPlan* p = static_cast<FftPlan*>(plan);
  MVF_CUFFT(cufftSetWorkArea(p->r2c, workspace));
  MVF_CUFFT(cufftSetWorkArea(p->c2r, workspace));
  return MVF_OK;
}

extern "C" int mvf_rl_fft_plan_destroy(void* plan) {
  if (!plan) return MVF_OK;
  FftPlan* p = static_cast<FftPlan*>(plan);
  cufftDestroy(p->r2c);
  cufftDestroy(p->c2r);
  delete p;
  return MVF_OK;
}

// spectrum_out = rfftn(pad(psf, [0, L - K]))   (psfs_ft, mv:5647-5650)
extern "C" int mvf_rl_fft_psf_spectrum(void* plan, const float* psf, const int32_t ksize[3], float* real_buf,
                                       void* spectrum_out, void* stream) {
  MVF_CHECK_ARG(plan && psf && ksize && real_buf && spectrum_out, "NULL argument");
  FftPlan* p = static_cast<FftPlan*>(plan);
  cudaStream_t st = as_stream(stream);
  fft_place_psf_kernel<<<sm_count() * 8, 256, 0, st>>>(psf, ksize[0], ksize[1], ksize[2], real_buf, p->L[0], p->L[1], p->L[2]);
  MVF_LAUNCHED();
  MVF_CUFFT(cufftSetStream(p->r2c, st));
  MVF_CUFFT(cufftExecR2C(p->r2c, real_buf, static_cast<cufftComplex*>(spectrum_out)));
  return MVF_OK;
}

// One blur + epilogue: mode 0 plain, 1 ratio, 2 correct (arguments as mvf_rl_blur / mvf_rl_ratio / mvf_rl_correct).
extern "C" int mvf_rl_fft_blur(void* plan, const float* in, const int32_t* zmap, int zmap_radius, const int32_t ksize[3],
                               const void* psf_spectrum, float* real_buf, void* cplx_buf, int mode,
                               const void* view, int dtype, const float* weights, float* out, int first, int last,
                               float* est, int64_t est_offset, double* sum_out, void* stream) {
  MVF_CHECK_ARG(plan && in && zmap && ksize && psf_spectrum && real_buf && cplx_buf, "NULL argument");
  MVF_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  FftPlan* p = static_cast<FftPlan*>(plan);
  cudaStream_t st = as_stream(stream);
  GatherArgs G;
  G.in = in; G.zmap = zmap; G.rp = zmap_radius; G.out = real_buf;
  for (int d = 0; d < 3; ++d) { G.n[d] = p->n[d]; G.L[d] = p->L[d]; G.K[d] = ksize[d]; G.R[d] = (ksize[d] - 1) / 2; }
  fft_gather_kernel<<<dim3((p->L[2] + 256 * FFT_XPT - 1) / (256 * FFT_XPT), p->L[1], p->L[0]), 256, 0, st>>>(G);
  MVF_LAUNCHED();
  MVF_CUFFT(cufftSetStream(p->r2c, st));
  MVF_CUFFT(cufftSetStream(p->c2r, st));
  cufftComplex* c = static_cast<cufftComplex*>(cplx_buf);
  MVF_CUFFT(cufftExecR2C(p->r2c, real_buf, c));
  const size_t nc = (size_t)p->L[0] * p->L[1] * (p->L[2] / 2 + 1);
  const float scale = 1.0f / ((float)p->L[0] * (float)p->L[1] * (float)p->L[2]);
  fft_multiply_kernel<<<sm_count() * 8, 256, 0, st>>>(c, static_cast<const cufftComplex*>(psf_spectrum), nc, scale);
  MVF_LAUNCHED();
  MVF_CUFFT(cufftExecC2R(p->c2r, c, real_buf));
  FftEpiArgs E;
  memset(&E, 0, sizeof(E));
  E.conv = real_buf; E.mode = mode; E.view = view; E.view_u16 = dtype == MVF_U16; E.weights = weights; E.out = out;
  E.first = first; E.last = last; E.est = est; E.est_off = est_offset; E.sum = sum_out;
  for (int d = 0; d < 3; ++d) { E.n[d] = p->n[d]; E.L[d] = p->L[d]; E.R[d] = (ksize[d] - 1) / 2; }
  fft_epilogue_kernel<<<dim3((p->n[2] + 256 * FFT_XPT - 1) / (256 * FFT_XPT), p->n[1], p->n[0]), 256, 0, st>>>(E);
  MVF_LAUNCHED();
  return MVF_OK;
}
