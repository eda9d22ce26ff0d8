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
  int L[3];    // padded extents (zx plans: L[1] = n[1], the y pass is a direct convolution)
  cufftHandle r2c, c2r;
  cufftHandle psf2d;   // zx plans: single 2-D R2C for the PSF spectrum
  int zx;
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

// ---- (z, x) FFT + direct y pass: PSFs of views rotated about y -------------------------------------------------------
// Such a PSF is an outer product a(y) (x) B(z, x) (get_psf resamples a separable Gaussian with a rotation that leaves y
// alone, mv:5376-5418; checked numerically per view).  The circular convolution then runs as batched 2-D transforms over
// (z, x) with y as the batch - a third less FFT work, no padding along y - and the y pass is a K-tap shift register in
// the epilogue kernel, which reads the convolved rows anyway.
__global__ void __launch_bounds__(256) fft_multiply_zx_kernel(cufftComplex* a, const cufftComplex* b, int Lxh, int ny, float scale) {
  const int xh = blockIdx.x * 256 + threadIdx.x;
  if (xh >= Lxh) return;
  const size_t i = ((size_t)blockIdx.z * ny + blockIdx.y) * Lxh + xh;
  const cufftComplex x = a[i], y = b[(size_t)blockIdx.z * Lxh + xh];
  a[i] = make_cuFloatComplex((x.x * y.x - x.y * y.y) * scale, (x.x * y.y + x.y * y.x) * scale);
}

constexpr int FFT_KMAX = MVF_RL_KMAX;
struct FftEpiYArgs {
  FftEpiArgs E;
  float ky[FFT_KMAX];   // taps along y, zero padded (centred) to the template size
  int ksz_y;            // original size (index map)
  int ychunk;
};

// thread = one x column of one plane, walks along y: K-deep shift register of partial sums (as rl_z_kernel), then the
// epilogue of fft_epilogue_kernel on every finished row
template <int K>
__global__ void __launch_bounds__(256) fft_epilogue_y_kernel(const __grid_constant__ FftEpiYArgs P) {
  constexpr int R = (K - 1) / 2;
  const FftEpiArgs& A = P.E;
  const int x = blockIdx.x * 256 + threadIdx.x, z = blockIdx.z;
  const bool valid = x < A.n[2];
  const int xc = valid ? x : 0;
  const int yo0 = blockIdx.y * P.ychunk, yo1 = min(A.n[1], yo0 + P.ychunk);
  const float* __restrict__ cplane = A.conv + (size_t)(z + A.R[0]) * A.L[1] * A.L[2] + A.R[2] + xc;   // L[1] == n[1]
  float S[K - 1];
#pragma unroll
  for (int m = 0; m < K - 1; ++m) S[m] = 0.f;
  double local_sum = 0.0;
  float pn = __ldg(cplane + (size_t)ext_index_fft(yo0 - R, A.n[1], P.ksz_y) * A.L[2]);
  for (int t = yo0 - R; t < yo1 + R; ++t) {
    const float p = pn;
    if (t + 1 < yo1 + R) pn = __ldg(cplane + (size_t)ext_index_fft(t + 1, A.n[1], P.ksz_y) * A.L[2]);
    const float e = fmaf(P.ky[0], p, S[0]);
#pragma unroll
    for (int m = 0; m < K - 2; ++m) S[m] = fmaf(P.ky[m + 1], p, S[m + 1]);
    S[K - 2] = P.ky[K - 1] * p;
    const int y = t - R;
    if (y < yo0 || !valid) continue;
    const size_t idx = ((size_t)z * A.n[1] + y) * A.n[2] + x;
    const float blur = fmaxf(e, 0.f);
    if (A.mode == 0) {
      A.out[idx] = blur;
    } else if (A.mode == 1) {
      const float I = A.view_u16 ? u16_to_float(__ldg(reinterpret_cast<const uint16_t*>(A.view) + idx))
                                 : __ldg(reinterpret_cast<const float*>(A.view) + idx);
      const float ratio = __fdiv_rn(I, blur + 1e-6f);
      A.out[idx] = __ldg(A.weights + idx) > 1e-5f ? ratio : 0.f;
    } else {
      const float ow = blur * __ldg(A.weights + idx);
      const float corr = A.first ? ow : A.out[idx] + ow;
      if (A.last) {
        const float v = fminf(fmaxf(A.est[A.est_off + idx] * corr, 0.f), 65535.f);
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

template <int K>
static void launch_epilogue_y(const FftEpiYArgs& P, cudaStream_t st) {
  dim3 grid((P.E.n[2] + 255) / 256, (P.E.n[1] + P.ychunk - 1) / P.ychunk, P.E.n[0]);
  fft_epilogue_y_kernel<K><<<grid, 256, 0, st>>>(P);
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

static int fft_plan_create(const int32_t dims[3], const int32_t ksize[3], void* workspace_or_null, size_t* workspace_bytes,
                           void** plan_out, int zx);

extern "C" int mvf_rl_fft_plan_create(const int32_t dims[3], const int32_t ksize[3], void* workspace_or_null,
                                      size_t* workspace_bytes, void** plan_out) {
  return fft_plan_create(dims, ksize, workspace_or_null, workspace_bytes, plan_out, 0);
}

// zx plan: batched 2-D (z, x) transforms with y as the batch for PSFs a(y) (x) B(z, x); the y pass is direct (<= 43 taps)
extern "C" int mvf_rl_fft_plan_create_zx(const int32_t dims[3], const int32_t ksize[3], void* workspace_or_null,
                                         size_t* workspace_bytes, void** plan_out) {
  MVF_CHECK_ARG(ksize && ksize[1] <= FFT_KMAX, "y taps of a zx plan are limited to 43");
  return fft_plan_create(dims, ksize, workspace_or_null, workspace_bytes, plan_out, 1);
}

static int fft_plan_create(const int32_t dims[3], const int32_t ksize[3], void* workspace_or_null, size_t* workspace_bytes,
                           void** plan_out, int zx) {
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
  p->zx = zx;
  p->psf2d = 0;
  if (zx) p->L[1] = dims[1];
  size_t w1 = 0, w2 = 0;
  MVF_CUFFT(cufftCreate(&p->r2c));
  MVF_CUFFT(cufftCreate(&p->c2r));
  MVF_CUFFT(cufftSetAutoAllocation(p->r2c, 0));
  MVF_CUFFT(cufftSetAutoAllocation(p->c2r, 0));
  if (zx) {
    // element (z, y, x) at (z * ny + y) * Lx + x: a 2-D transform over (z, x) per y = rows of "width" ny * Lx, batches Lx apart
    const int Lxh = p->L[2] / 2 + 1;
    int n2[2] = {p->L[0], p->L[2]};
    int rembed[2] = {p->L[0], p->n[1] * p->L[2]}, cembed[2] = {p->L[0], p->n[1] * Lxh};
    MVF_CUFFT(cufftMakePlanMany(p->r2c, 2, n2, rembed, 1, p->L[2], cembed, 1, Lxh, CUFFT_R2C, p->n[1], &w1));
    MVF_CUFFT(cufftMakePlanMany(p->c2r, 2, n2, cembed, 1, Lxh, rembed, 1, p->L[2], CUFFT_C2R, p->n[1], &w2));
    MVF_CUFFT(cufftPlan2d(&p->psf2d, p->L[0], p->L[2], CUFFT_R2C));
  } else {
    MVF_CUFFT(cufftMakePlan3d(p->r2c, p->L[0], p->L[1], p->L[2], CUFFT_R2C, &w1));
    MVF_CUFFT(cufftMakePlan3d(p->c2r, p->L[0], p->L[1], p->L[2], CUFFT_C2R, &w2));
  }
  p->work_bytes = w1 > w2 ? w1 : w2;
  *workspace_bytes = p->work_bytes;
  if (workspace_or_null) {
    MVF_CUFFT(cufftSetWorkArea(p->r2c, workspace_or_null));
    MVF_CUFFT(cufftSetWorkArea(p->c2r, workspace_or_null));
  }
  *plan_out = p;
  return MVF_OK;
}

extern "C" int mvf_rl_fft_plan_padded(void* plan, int32_t padded[3]) {
  MVF_CHECK_ARG(plan && padded, "NULL argument");
  FftPlan* p = static_cast<FftPlan*>(plan);
  for (int d = 0; d < 3; ++d) padded[d] = p->L[d];
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
  if (p->psf2d) cufftDestroy(p->psf2d);
  delete p;
  return MVF_OK;
}

// spectrum_out = rfftn(pad(psf, [0, L - K]))   (psfs_ft, mv:5647-5650)
extern "C" int mvf_rl_fft_psf_spectrum(void* plan, const float* psf, const int32_t ksize[3], float* real_buf,
                                       void* spectrum_out, void* stream) {
  MVF_CHECK_ARG(plan && psf && ksize && real_buf && spectrum_out, "NULL argument");
  FftPlan* p = static_cast<FftPlan*>(plan);
  cudaStream_t st = as_stream(stream);
  if (p->zx) {   // psf = the (kz, kx) factor B; spectrum (L0, L2/2+1)
    MVF_CHECK_ARG(ksize[1] == 1, "a zx plan takes the (kz, 1, kx) factor of the PSF");
    fft_place_psf_kernel<<<sm_count() * 8, 256, 0, st>>>(psf, ksize[0], 1, ksize[2], real_buf, p->L[0], 1, p->L[2]);
    MVF_LAUNCHED();
    MVF_CUFFT(cufftSetStream(p->psf2d, st));
    MVF_CUFFT(cufftExecR2C(p->psf2d, real_buf, static_cast<cufftComplex*>(spectrum_out)));
    return MVF_OK;
  }
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

// The same blur + epilogue on a zx plan: ksize = full PSF size, host_ky = the ksize[1] taps a(y) (host), spectrum of the
// (z, x) factor from mvf_rl_fft_psf_spectrum.
extern "C" int mvf_rl_fft_blur_zx(void* plan, const float* in, const int32_t* zmap, int zmap_radius, const int32_t ksize[3],
                                  const float* host_ky, const void* psf_spectrum, float* real_buf, void* cplx_buf, int mode,
                                  const void* view, int dtype, const float* weights, float* out, int first, int last,
                                  float* est, int64_t est_offset, double* sum_out, void* stream) {
  MVF_CHECK_ARG(plan && in && zmap && ksize && host_ky && psf_spectrum && real_buf && cplx_buf, "NULL argument");
  MVF_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  FftPlan* p = static_cast<FftPlan*>(plan);
  MVF_CHECK_ARG(p->zx, "plan was not created with mvf_rl_fft_plan_create_zx");
  MVF_CHECK_ARG(ksize[1] >= 1 && ksize[1] <= FFT_KMAX && (ksize[1] & 1), "y taps must be odd and <= 43");
  cudaStream_t st = as_stream(stream);
  GatherArgs G;
  G.in = in; G.zmap = zmap; G.rp = zmap_radius; G.out = real_buf;
  for (int d = 0; d < 3; ++d) { G.n[d] = p->n[d]; G.L[d] = p->L[d]; G.K[d] = ksize[d]; G.R[d] = (ksize[d] - 1) / 2; }
  G.K[1] = 1; G.R[1] = 0;   // rows are copied as they are: no padding along y
  fft_gather_kernel<<<dim3((p->L[2] + 256 * FFT_XPT - 1) / (256 * FFT_XPT), p->L[1], p->L[0]), 256, 0, st>>>(G);
  MVF_LAUNCHED();
  MVF_CUFFT(cufftSetStream(p->r2c, st));
  MVF_CUFFT(cufftSetStream(p->c2r, st));
  cufftComplex* c = static_cast<cufftComplex*>(cplx_buf);
  MVF_CUFFT(cufftExecR2C(p->r2c, real_buf, c));
  const int Lxh = p->L[2] / 2 + 1;
  const float scale = 1.0f / ((float)p->L[0] * (float)p->L[2]);
  fft_multiply_zx_kernel<<<dim3((Lxh + 255) / 256, p->n[1], p->L[0]), 256, 0, st>>>(c, static_cast<const cufftComplex*>(psf_spectrum), Lxh,
                                                                                   p->n[1], scale);
  MVF_LAUNCHED();
  MVF_CUFFT(cufftExecC2R(p->c2r, c, real_buf));
  FftEpiYArgs P;
  memset(&P, 0, sizeof(P));
  FftEpiArgs& E = P.E;
  E.conv = real_buf; E.mode = mode; E.view = view; E.view_u16 = dtype == MVF_U16; E.weights = weights; E.out = out;
  E.first = first; E.last = last; E.est = est; E.est_off = est_offset; E.sum = sum_out;
  for (int d = 0; d < 3; ++d) { E.n[d] = p->n[d]; E.L[d] = p->L[d]; E.R[d] = (ksize[d] - 1) / 2; }
  const int sizes[] = {7, 13, 19, 25, 31, 37, 43};
  int kp = 43;
  for (int sz : sizes)
    if (ksize[1] <= sz) { kp = sz; break; }
  const int pad = (kp - ksize[1]) / 2;
  for (int j = 0; j < ksize[1]; ++j) P.ky[pad + j] = host_ky[j];
  P.ksz_y = ksize[1];
  P.ychunk = p->n[1] < 256 ? p->n[1] : 256;
  switch (kp) {
    case 7: launch_epilogue_y<7>(P, st); break;
    case 13: launch_epilogue_y<13>(P, st); break;
    case 19: launch_epilogue_y<19>(P, st); break;
    case 25: launch_epilogue_y<25>(P, st); break;
    case 31: launch_epilogue_y<31>(P, st); break;
    case 37: launch_epilogue_y<37>(P, st); break;
    default: launch_epilogue_y<43>(P, st); break;
  }
  MVF_LAUNCHED();
  return MVF_OK;
}
