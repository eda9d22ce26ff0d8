/*
 * mvregfus_b200 — C-ABI of the B200-native (sm_100a) view-fusion hot path of MVRegFus.
 *
 * The reference (m-albert/MVRegFus) is pure Python and has no FFI: its "plugin socket" is a set of
 * Python functions in mvregfus/multiview.py that fuse_block() calls by keyword
 * (reference: mvregfus/multiview.py:3817-3916).  The drop-in wrappers in mvregfus_b200/multiview.py
 * keep those names and signatures; this header is what they bind (ctypes).
 * Each entry point cites the reference code whose arithmetic it replaces ("mv" =
 * mvregfus/multiview.py).
 *
 * Conventions
 *   - plain pointers and sizes only; all array pointers are DEVICE pointers unless named host_*;
 *     the caller owns every buffer; nothing is freed behind the caller's back.
 *   - volumes are C-contiguous (nz, ny, nx), x fastest, like the reference's numpy arrays.
 *   - affine maps are index-space maps  u_view = matrix * i_fused + offset  (row-major 3x3, zyx
 *     order, float64), i.e. matrix'/offset' of mv:1486-1496.
 *   - every function returns 0 on success; on failure a negative code, and mvf_last_error()
 *     returns a thread-local message.  There is no CPU fallback.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous
 *     with respect to the host unless stated.
 */
#ifndef MVREGFUS_B200_H
#define MVREGFUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVF_ABI_VERSION 1
#define MVF_MAX_VIEWS 8      /* views fused in one pass (register-resident weights)           */
#define MVF_SIG_N 200        /* samples per axis of the blending field, mv:3566               */
#define MVF_RL_KMAX 43       /* longest PSF factor handled by the RL convolution kernels        */

enum { MVF_U16 = 0, MVF_F32 = 1 };                 /* voxel dtypes                            */
enum { MVF_RULE_SCIPY = 0, MVF_RULE_ITK = 1 };     /* boundary rule of order-1 resampling      */
enum { MVF_W_OFF = 0, MVF_W_BLEND = 1, MVF_W_BLEND_DCT = 2 };

enum {
  MVF_OK = 0,
  MVF_ERR_INVALID = -1,      /* bad argument                                                  */
  MVF_ERR_CUDA = -2,         /* a CUDA runtime call failed (message has the cudaError string) */
  MVF_ERR_UNSUPPORTED = -3   /* valid request this build cannot serve                         */
};

/* One view of a fusion call.  `matrix/offset` map fused index -> view index (scipy path,
 * mv:1494-1496).  `wmatrix/woffset` map fused index -> continuous index of the view's 200^3
 * blending field (ITK path: field origin o_v + s_v, spacing (size_v-2)/200*s_v, mv:3567,3604).
 * `profiles` = 3 x MVF_SIG_N float32 (f_z, f_y, f_x): the field is min(f_z[i], f_y[j], f_x[k])
 * (mv:3558-3598).  `one_lo/one_hi`: per axis the index range [lo,hi) on which f[i]==f[i+1]==1.
 * `cgrid`: optional coarse DCT weight grid of this view (float32, cdims) with the index map
 * fused index -> coarse index given by cscale/coffset (diagonal; mv:4249-4253, 4291-4317). */
typedef struct mvf_view {
  const void* data;
  int32_t dtype;
  int32_t dims[3];
  double matrix[9];
  double offset[3];
  int32_t weight_mode;        /* MVF_W_*; MVF_W_OFF = zero weight (probe early-out, mv:3491-3532) */
  int32_t one_lo[3];
  int32_t one_hi[3];
  double wmatrix[9];
  double woffset[3];
  const float* profiles;
  const float* cgrid;
  int32_t cdims[3];
  double cscale[3];
  double coffset[3];
} mvf_view;

/* ---- library ---------------------------------------------------------------------------- */
int mvf_abi_version(void);
const char* mvf_last_error(void);
/* Number of kernel launches issued by this library on the calling thread since the last reset
 * (bench.py's `gpu_launches`). */
int64_t mvf_launch_count(int reset);

/* ---- K1: order-1 affine resampling of one view -------------------------------------------
 * Replaces scipy.ndimage.affine_transform(order=1, mode='constant') as called by
 * transform_chunk (mv:1546-1616, MVF_RULE_SCIPY: 0 outside [0,n-1]; uint16 = round-half-up) and
 * sitk.Resample(linear)+sitk.Clamp as called by transformStack (mv:6868-6871, MVF_RULE_ITK:
 * valid on [-0.5,n-0.5), edge clamped, integers truncated).  dst has the dtype of src.
 * dst covers the index box [dst_origin, dst_origin+dst_dims) of the fused frame, so a z-slab or
 * a 128^3 chunk can be produced in place. */
int mvf_affine_resample(const void* src, int dtype, const int32_t src_dims[3],
                        const double matrix[9], const double offset[3],
                        void* dst, const int32_t dst_dims[3], const int32_t dst_origin[3],
                        int rule, void* stream);

/* ---- blending weights on a target grid ----------------------------------------------------
 * Replaces get_weights_simple (mv:3469-3633): per view the ITK-linear resampling of the
 * analytic sigmoid-border field, optionally normalised by the sum over views (0 -> 1 guard).
 * weights_out: nviews x dims float32, view-major. */
int mvf_blend_weights(int nviews, const mvf_view* host_views, float* weights_out,
                      const int32_t dims[3], const int32_t origin[3], int normalise, void* stream);

/* ---- weighted additive fusion of already transformed views --------------------------------
 * Replaces fuse_views_weights (mv:4904-4912): f += uint16(w_v * float64(I_v)) per view
 * (truncation per view), clip, uint16.  uint16 views accumulate modulo 2^16 like the reference's
 * uint16 `f`; float32 views accumulate exactly and clip.  weights==NULL -> mean of the views.
 * out_f32 (optional) receives sum_v w_v*I_v before truncation. */
int mvf_fuse_weighted(int nviews, const void* const* host_view_ptrs, int dtype,
                      const float* const* host_weight_ptrs, uint16_t* out, float* out_f32,
                      size_t nvoxels, void* stream);

/* ---- K2: fused single pass: resample + blending(/DCT) weights + normalise + weighted sum ---
 * Replaces transform_stack_dask_numpy -> get_weights_simple (x coarse DCT weights) ->
 * fuse_views_weights for the whole box without materialising transformed views or weights
 * (mv:1458-1519, 3469-3633, 4291-4317, 4434-4439, 4904-4912).  out: uint16 box of out_dims at
 * out_origin of the fused frame.  out_f32 optional as above. */
int mvf_fuse_blend(int nviews, const mvf_view* host_views, uint16_t* out, float* out_f32,
                   const int32_t out_dims[3], const int32_t out_origin[3], void* stream);

/* ---- block semantics of fuse_block on the GPU (mv:3817-3916) ----------------------------------------------
 * fuse_blockwise fuses 128^3 chunks independently and fuse_block drops, per chunk, the views whose transformed
 * voxels are all zero, short-circuits chunks with 0/1 live views to a raw copy, and drops views whose blending
 * weights vanish on the chunk.  mvf_chunk_probe computes the two per-(chunk, view) facts on the device
 * (probe 0: some transformed voxel != 0; probe 1: some raw blending weight > 0; bit v of chunk_flags[chunk],
 * chunks indexed z-major over `nchunks`, flags must be zeroed by the caller); the caller combines them with the
 * per-chunk probe-point early-out (host) into chunk_info words: bits 0-7 = views taking part, bit 8 = copy
 * mode, bits 12-15 = view whose transformed voxels are copied.  mvf_fuse_blend_chunked is mvf_fuse_blend with
 * those per-chunk decisions applied.  Box origins must be multiples of the 8x8x32 tile. */
int mvf_chunk_probe(int nviews, const mvf_view* host_views, int probe, const int32_t dims[3],
                    const int32_t origin[3], uint32_t* chunk_flags, int chunk, const int32_t nchunks[3],
                    void* stream);
int mvf_fuse_blend_chunked(int nviews, const mvf_view* host_views, uint16_t* out, const int32_t out_dims[3],
                           const int32_t out_origin[3], const uint32_t* chunk_info, int chunk,
                           const int32_t nchunks[3], void* stream);

/* ---- K4/K5: multi-view Richardson-Lucy with weights (fuse_LR_with_weights_np, mv:5545-5749) ------
 * A PSF (get_psf, mv:5376-5418) that is an outer product kz (x) ky (x) kx is passed by its factors
 * (true-convolution taps, odd sizes).  blur(x)[i] = sum_j psf[j] * ext(i + R - j) with the
 * reference's boundary handling (reflect above, wrap into the far reflected tail below; closed form
 * of mv:5654-5673), result clamped at 0.  In y and x the index map is applied inside the kernel; in z
 * the caller provides `zmap` (device, length dims[0] + 2*Rp, Rp = (mvf_rl_padded_size(max ksize)-1)/2):
 * zmap[t + Rp] = plane of the input buffer holding extended z position t.  A single-GPU caller fills
 * it with the same closed form; a z-slab caller points it at halo planes received from its
 * neighbours, so the kernels are identical in both cases. */
typedef struct mvf_rl_psf {
  float kz[MVF_RL_KMAX];
  float ky[MVF_RL_KMAX];
  float kx[MVF_RL_KMAX];
  int32_t ksize[3];
  /* A PSF that is not an outer product (oblique view) is passed densely instead: `dense` = device pointer to a
   * Kp^3 float32 array, Kp = mvf_rl_padded_size(max ksize), the PSF centred and zero padded; the factors are
   * then ignored and a direct (non-separable) convolution kernel runs.  NULL = use the factors. */
  const float* dense;
} mvf_rl_psf;

/* kernel template size used for a PSF whose longest factor has k taps (7, 13, ..., 43 in steps of 6; -1 if k > 43) */
int mvf_rl_padded_size(int k);

/* estimate = sum_v w_v * I_v (float32, views added in order; mv:5629); *sum_out += sum(estimate). */
int mvf_rl_init_estimate(int nviews, const void* const* host_view_ptrs, int dtype,
                         const float* const* host_weight_ptrs, float* est, size_t nvoxels,
                         double* sum_out, void* stream);

/* out = max(blur(in), 0)                                                          (mv:5654-5673) */
int mvf_rl_blur(const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                float* out, void* stream);

/* ratio_out = I_v / (max(blur(est),0) + 1e-6) * (w_v > 1e-5)                      (mv:5686-5699) */
int mvf_rl_ratio(const float* est, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                 const void* view, int dtype, const float* weights, float* ratio_out, void* stream);

/* corr (+)= max(blur(ratio),0) * w_v ; on the last view  est = clip(est * corr, 0, 65535) and
 * *sum_out += sum(est)  (mv:5706-5727).  `est` + est_offset addresses output plane 0 (the estimate
 * may live in a halo-padded buffer). */
int mvf_rl_correct(const float* ratio, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                   const float* weights, float* corr, int first, int last, float* est,
                   int64_t est_offset, double* sum_out, void* stream);

/* Streaming formulation of the same blur for separable PSFs (default when rows are 16-byte aligned): mvf_rl_xy runs
 * the x and y passes on every plane of the input buffer (`nplanes` planes of dims[1] x dims[2]; for a z-slab that
 * includes the halo planes), mvf_rl_*_z run the z pass through the plane map on that intermediate and apply the
 * same epilogues as mvf_rl_blur / mvf_rl_ratio / mvf_rl_correct.  Two kernels without shared memory or barriers
 * replace one fused kernel; the intermediate costs one extra float32 round trip through HBM.
 * mvf_rl_ratio_z accepts weights == NULL when the caller has already zeroed the view wherever w <= 1e-5 (I = 0 makes
 * the ratio exactly 0 there, which is all the mask of mv:5636, 5695-5699 does): the pass then reads 4 bytes less per voxel.
 * mvf_rl_stream_supported: 1 if the PSF is separable and dims[2] is a multiple of the vector width (4 columns up
 * to 19 taps, 2 beyond). */
int mvf_rl_stream_supported(const int32_t dims[3], const mvf_rl_psf* psf);
int mvf_rl_xy(const float* in, int32_t nplanes, const int32_t dims[3], const mvf_rl_psf* psf, float* out, void* stream);
int mvf_rl_blur_z(const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf, float* out,
                  void* stream);
int mvf_rl_ratio_z(const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                   const void* view, int dtype, const float* weights, float* ratio_out, void* stream);
int mvf_rl_correct_z(const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                     const float* weights, float* corr, int first, int last, float* est, int64_t est_offset,
                     double* sum_out, void* stream);

/* Second blur with its z pass first (the passes of a separable blur commute), so that the ratio volume never reaches
 * HBM: mvf_rl_zz turns T1 = mvf_rl_xy(estimate) into T2 = blur_z(I / (max(blur_z(T1), 0) + 1e-6)) with two shift
 * registers per column; mvf_rl_xy_correct runs the x/y passes on T2 and applies max(., 0) and the correction epilogue
 * of mvf_rl_correct to every finished row (mv:5686-5723).
 * The two plane maps are HOST arrays: they are piecewise affine and travel in the kernel parameters as (at most 8) runs
 * of constant step (MVF_ERR_UNSUPPORTED beyond that):
 *   zmap1 (host, dims[0] + 4*Rp): extended position s in [-2Rp, dims[0] + 2Rp) -> plane of t1;
 *   rmap  (host, dims[0] + 2*Rp): extended position t in [-Rp, dims[0] + Rp) -> plane of `view` (>= 0), or -1 - k:
 *   the ratio at t is plane k of `side` (positions outside the volume repeat ratios of planes near the top - reflect
 *   above, wrap below - which the caller computes first with mvf_rl_ratio_z on those planes).  `view` must already be
 *   zero wherever the weight is <= 1e-5 (the mask of mv:5695-5699). */
int mvf_rl_zz(const float* t1, const int32_t* host_zmap1, const int32_t* host_rmap, const int32_t dims[3], const mvf_rl_psf* psf,
              const void* view, int dtype, const float* side, float* t2_out, void* stream);
int mvf_rl_xy_correct(const float* t2, int32_t nplanes, const int32_t dims[3], const mvf_rl_psf* psf,
                      const float* weights, float* corr, int first, int last, float* est, int64_t est_offset,
                      double* sum_out, void* stream);

/* ---- RL, general PSFs: cuFFT circular convolution with the reference's padding (mv:5647-5673) ------------
 * For PSFs that are not outer products (oblique views).  The padded extents are the next 2^a3^b5^c7^d >=
 * n + K - 1 per axis (the reference uses n + K + 1; the extra entries only hold zeros, the reflected tail and
 * the wrap-around values are placed so that every output voxel sees the reference's inputs).  The caller owns
 * all buffers: `real_buf` = prod(padded) floats, `cplx_buf` and each PSF spectrum = padded[0]*padded[1]*
 * (padded[2]/2+1) float2, `workspace` = *workspace_bytes bytes (cuFFT work area).  mode: 0 = plain blur into
 * `out`, 1 = ratio, 2 = correct; remaining arguments as for mvf_rl_ratio / mvf_rl_correct. */
int mvf_rl_fft_padded_dims(const int32_t dims[3], const int32_t ksize[3], int32_t padded[3]);
int mvf_rl_fft_plan_create(const int32_t dims[3], const int32_t ksize[3], void* workspace_or_null,
                           size_t* workspace_bytes, void** plan_out);
int mvf_rl_fft_plan_set_workspace(void* plan, void* workspace);
int mvf_rl_fft_plan_destroy(void* plan);
int mvf_rl_fft_psf_spectrum(void* plan, const float* psf, const int32_t ksize[3], float* real_buf,
                            void* spectrum_out, void* stream);
int mvf_rl_fft_blur(void* plan, const float* in, const int32_t* zmap, int zmap_radius, const int32_t ksize[3],
                    const void* psf_spectrum, float* real_buf, void* cplx_buf, int mode,
                    const void* view, int dtype, const float* weights, float* out, int first, int last,
                    float* est, int64_t est_offset, double* sum_out, void* stream);
/* PSFs of views rotated about y factor as a(y) (x) B(z, x): a zx plan runs the circular convolution as batched 2-D (z, x)
 * transforms with y as the batch (a third less FFT work, no padding along y) and the y pass as a direct K-tap
 * convolution inside the epilogue kernel (ksize[1] <= 43).  mvf_rl_fft_psf_spectrum then takes the (kz, 1, kx) factor B,
 * mvf_rl_fft_blur_zx the taps a(y) as a host array; buffers are sized from mvf_rl_fft_plan_padded. */
int mvf_rl_fft_plan_create_zx(const int32_t dims[3], const int32_t ksize[3], void* workspace_or_null,
                              size_t* workspace_bytes, void** plan_out);
int mvf_rl_fft_plan_padded(void* plan, int32_t padded[3]);
int mvf_rl_fft_blur_zx(void* plan, const float* in, const int32_t* zmap, int zmap_radius, const int32_t ksize[3],
                       const float* host_ky, const void* psf_spectrum, float* real_buf, void* cplx_buf, int mode,
                       const void* view, int dtype, const float* weights, float* out, int first, int last,
                       float* est, int64_t est_offset, double* sum_out, void* stream);

/* ---- K3: DCT-entropy quality of blocks of a transformed view (determine_chunk_quality, mv:4496-4561) --
 * For every job (linear index into the nblocks grid) the size^3 block is read with stride `bin` from the
 * transformed view (zero beyond its extents), zero-filled like the reference, DCT-II(ortho)-transformed in
 * float64 and reduced to q = -sum|d| log2(|d|/sum|d|); q_out[job block] receives the raw (un-normalised)
 * value.  `cosm` is the size x size orthonormal DCT-II matrix (device, float64); `scratch` must hold
 * mvf_dct_scratch_bytes(size) bytes.  Which views are inside a block and the normalisation over views are
 * host decisions (blocks_inside, mv:4496-4508, 4548-4556). */
int mvf_dct_scratch_bytes(int size, int* grid_out);
int mvf_dct_entropy(const void* tview, int dtype, const int32_t dims[3], int bin, int size,
                    const int32_t nblocks[3], const int32_t* jobs, int njobs, const double* cosm,
                    double* scratch, double* q_out, void* stream);

/* ---- steps either side of the path: .ims pyramid and raw-view preprocessing ----------------
 * mvf_subsample: out = in[::step[0], ::step[1], ::step[2]] (subsample_data, imaris.py:764-765); out extents are
 * ceil(in_dims / step).  The writer applies it cumulatively level after level (imaris.py:146-148, 321-323).
 * mvf_range_u16 / mvf_hist_u16: the two device passes of np.histogram(data, 256) on uint16 data (imaris.py:155,
 * 324): minmax[0..1] = min, max; `lut` maps every uint16 value to its bin (built on the host with numpy's own
 * binning once the range is known), counts[256] receives the histogram.
 * mvf_background_u16: out = (in - level) * (in > level) in uint16 arithmetic (multiview.py:316).
 * mvf_bin_shrink: sitk.BinShrink (mv_utils.py:288-311): mean over factors[0] x factors[1] x factors[2] (z,y,x)
 * bins, partial bins dropped (out extents floor(in_dims / factors)), uint16 results rounded half up. */
int mvf_subsample(const void* in, int dtype, const int32_t in_dims[3], const int32_t step[3], void* out, void* stream);
int mvf_range_u16(const uint16_t* in, size_t n, uint32_t* minmax, void* stream);
int mvf_hist_u16(const uint16_t* in, size_t n, const uint8_t* lut, uint64_t* counts, void* stream);
int mvf_background_u16(const uint16_t* in, uint16_t* out, size_t n, uint32_t level, void* stream);
int mvf_bin_shrink(const void* in, int dtype, const int32_t in_dims[3], const int32_t factors[3], void* out,
                   void* stream);

/* ---- dual-illumination fusion, plane by plane (illumination_fusion_planewise, mv:391-467) --------------------------
 * pair: (2, nplanes, n, n) uint16, the two light sheets with the sheet's direction along rows (axis -2) - the layout after
 * the reference's swapaxes; square planes as the reference requires.  dev_taps: the 2*radius+1 float64 taps of scipy's
 * gaussian_filter for sigma = n / 50 (device).  scratch: 8 * px + 8 * nplanes * n + 16 * nplanes + 64 bytes, px = nplanes*n*n.
 * out: (nplanes, n, n) uint16. */
int mvf_illum_fuse_planes(const uint16_t* pair, int32_t nplanes, int32_t n, const double* dev_taps, int32_t radius,
                          void* scratch, size_t scratch_bytes, uint16_t* out, void* stream);

/* ---- source brick of a target box: strided host -> device copy ---------------------------------------------
 * view[lo:hi] (zyx, C-contiguous host array of `dims`) -> dense device brick, one cudaMemcpy3DAsync: what a z-slab rank
 * uploads instead of whole views (the per-chunk crop of transform_chunk, mv:1552-1601). */
int mvf_copy_brick_h2d(const void* host, int dtype, const int32_t dims[3], const int32_t lo[3], const int32_t hi[3],
                       void* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVREGFUS_B200_H */
