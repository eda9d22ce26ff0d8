// Steps either side of the fusion path, sm_100a.  All of them are pure streaming
// work bound by HBM bandwidth: one pass, 16-byte accesses where the layout allows, grids sized in multiples of
// the SM count.
//   * output pyramid of the .ims writer: strided subsampling data[::sz, ::sy, ::sx]   (imaris.py:764-765,
//     applied cumulatively level after level, imaris.py:146-148 / 321-323) and the 256-bin histogram of each
//     level (np.histogram(data, 256), imaris.py:155, 324);
//   * raw-view preprocessing: background subtraction (stack - bg) * (stack > bg)       (multiview.py:316) and
//     integer binning sitk.BinShrink                                                    (mv_utils.py:288-311).
#include "common.cuh"

namespace mvf {

// ---- strided subsampling ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) subsample_kernel(const T* __restrict__ in, T* __restrict__ out, int nz, int ny, int nx,
                                                        int oz, int oy, int ox, int sz, int sy, int sx) {
  const size_t total = (size_t)oz * oy * ox;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int x = (int)(i % ox);
    const size_t r = i / ox;
    const int y = (int)(r % oy), z = (int)(r / oy);
    out[i] = in[((size_t)z * sz * ny + (size_t)y * sy) * nx + (size_t)x * sx];
  }
}

// uint16, x step 2, 16-byte aligned rows: a thread turns 8 consecutive source voxels (one LDG.128) into 4 outputs
// (one STG.64), so the even rows of the even planes are streamed with full sectors.
__global__ void __launch_bounds__(256) subsample_x2_u16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int ny, int nx,
                                                               int oz, int oy, int ox, int sz, int sy) {
  const int cpr = nx / 8;                       // chunks per source row (ox == nx / 2 == 4 * cpr)
  const size_t total = (size_t)oz * oy * cpr;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % cpr);
    const size_t r = i / cpr;
    const int y = (int)(r % oy), z = (int)(r / oy);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + ((size_t)z * sz * ny + (size_t)y * sy) * nx) + c);
    const uint2 o = make_uint2((q.x & 0xFFFFu) | (q.y << 16), (q.z & 0xFFFFu) | (q.w << 16));
    reinterpret_cast<uint2*>(out + ((size_t)z * oy + y) * ox)[c] = o;
  }
}

// ---- value range + 256-bin histogram of uint16 data --------------------------------------------------------
// np.histogram(data, 256) bins over [min, max] with float64 edges.  For uint16 data the bin of a voxel is a
// function of its value alone, so the host derives a 65536-entry value -> bin table FROM NUMPY ITSELF (exact by
// construction) once the range is known; the device only counts.  Pass 1: min/max.  Pass 2: table look-up
// (64 KB of shared memory) + per-warp private 256-bin counters, merged with one atomic per bin and CTA.
__global__ void __launch_bounds__(256) range_u16_kernel(const uint16_t* __restrict__ in, size_t n, unsigned* __restrict__ minmax) {
  unsigned lo = 0xFFFFu, hi = 0u;
  const size_t nvec = n / 8;
  const uint4* __restrict__ v = reinterpret_cast<const uint4*>(in);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
    const uint4 q = __ldg(v + i);
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned a = w[j] & 0xFFFFu, b = w[j] >> 16;
      lo = min(lo, min(a, b));
      hi = max(hi, max(a, b));
    }
  }
  if (blockIdx.x == 0)
    for (size_t i = nvec * 8 + threadIdx.x; i < n; i += 256) { lo = min(lo, (unsigned)in[i]); hi = max(hi, (unsigned)in[i]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(minmax, lo); atomicMax(minmax + 1, hi); }
}

constexpr int HREP = 1;   // private counter sets per warp (4 sets at 2 CTAs/SM measured slower: 0.77 vs 0.57 ms on 512x1024x1024)

__global__ void __launch_bounds__(256) hist_u16_kernel(const uint16_t* __restrict__ in, size_t n, const uint8_t* __restrict__ lut,
                                                       unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char hs[];
  uint8_t* slut = hs;                                            // 65536 bytes
  unsigned* sc = reinterpret_cast<unsigned*>(hs + 65536);        // 8 warps x HREP lane groups x 256 counters
  for (int i = threadIdx.x; i < 65536 / 16; i += 256) reinterpret_cast<uint4*>(slut)[i] = __ldg(reinterpret_cast<const uint4*>(lut) + i);
  for (int i = threadIdx.x; i < 8 * HREP * 256; i += 256) sc[i] = 0u;
  __syncthreads();
  unsigned* mine = sc + ((threadIdx.x >> 5) * HREP + (threadIdx.x & (HREP - 1))) * 256;   // lanes of a warp spread over HREP copies
  const size_t nvec = n / 8;
  const uint4* __restrict__ v = reinterpret_cast<const uint4*>(in);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
    const uint4 q = __ldg(v + i);
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
    // neighbouring voxels often share a bin: merge equal consecutive bins before touching shared memory
    unsigned cur = slut[w[0] & 0xFFFFu], run = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const unsigned b = slut[h ? (w[j] >> 16) : (w[j] & 0xFFFFu)];
        if (b == cur) { ++run; } else { atomicAdd(mine + cur, run); cur = b; run = 1; }
      }
    }
    atomicAdd(mine + cur, run);
  }
  if (blockIdx.x == 0)
    for (size_t i = nvec * 8 + threadIdx.x; i < n; i += 256) atomicAdd(mine + slut[in[i]], 1u);
  __syncthreads();
  unsigned s = 0;
#pragma unroll
  for (int w = 0; w < 8 * HREP; ++w) s += sc[w * 256 + threadIdx.x];
  if (s) atomicAdd(counts + threadIdx.x, (unsigned long long)s);
}

// ---- background subtraction (uint16): (stack - bg) * (stack > bg) with numpy's uint16 wrap semantics ------------
__global__ void __launch_bounds__(256) background_u16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, size_t n,
                                                             unsigned bg) {
  const size_t nvec = n / 8;
  const uint4* __restrict__ v = reinterpret_cast<const uint4*>(in);
  uint4* __restrict__ o = reinterpret_cast<uint4*>(out);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
    const uint4 q = __ldg(v + i);
    unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned a = w[j] & 0xFFFFu, b = w[j] >> 16;
      const unsigned ra = a > bg ? a - bg : 0u, rb = b > bg ? b - bg : 0u;
      w[j] = ra | (rb << 16);
    }
    o[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (blockIdx.x == 0)
    for (size_t i = nvec * 8 + threadIdx.x; i < n; i += 256) { const unsigned a = in[i]; out[i] = (uint16_t)(a > bg ? a - bg : 0u); }
}

// ---- BinShrink: mean over fz x fy x fx bins, partial bins dropped; integer output rounded half up -------------
// (itk::BinShrinkImageFilter: accumulate in the real type, divide by the bin volume, RoundIfInteger)
template <typename T>
__global__ void __launch_bounds__(256) bin_shrink_kernel(const T* __restrict__ in, T* __restrict__ out, int ny, int nx, int oz, int oy,
                                                         int ox, int fz, int fy, int fx) {
  const size_t total = (size_t)oz * oy * ox;
  const double num = (double)fz * fy * fx;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int x = (int)(i % ox);
    const size_t r = i / ox;
    const int y = (int)(r % oy), z = (int)(r / oy);
    double s = 0.0;
    for (int a = 0; a < fz; ++a)
      for (int b = 0; b < fy; ++b) {
        const T* __restrict__ row = in + ((size_t)(z * fz + a) * ny + (size_t)(y * fy + b)) * nx + (size_t)x * fx;
        for (int c = 0; c < fx; ++c) s += (double)row[c];
      }
    const double m = s / num;
    if (sizeof(T) == 2) out[i] = (T)(unsigned)floor(m + 0.5);
    else out[i] = (T)m;
  }
}

// uint16, x factor 2, 16-byte aligned rows: exact integer accumulation (a bin holds < 2^16 voxels) and
// floor(sum / num + 0.5) == (2 sum + num) / (2 num) in integers (the quotient never comes within 1/(2 num) of a tie
// without being one), 8 source voxels per LDG.128, 4 outputs per STG.64.
// exact x / d for x < 2^31, d < 2^16 (quotient < 2^16): float estimate, one correction step either way
__device__ __forceinline__ unsigned div_small(unsigned x, unsigned d, float inv) {
  unsigned q = (unsigned)((float)x * inv);
  if (q * d > x) --q;
  if ((q + 1) * d <= x) ++q;
  return q;
}

__global__ void __launch_bounds__(256) bin_shrink_x2_u16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int ny, int nx,
                                                                int oz, int oy, int ox, int fz, int fy) {
  const int cpr = ox / 4;                        // output chunks (4 voxels) per row; the source row has >= 8 * cpr voxels
  const size_t total = (size_t)oz * oy * cpr;
  const unsigned num = (unsigned)(fz * fy * 2);
  const float inv = 1.0f / (float)(2 * num);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % cpr);
    const size_t r = i / cpr;
    const int y = (int)(r % oy), z = (int)(r / oy);
    unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int a = 0; a < fz; ++a)
#pragma unroll 4
      for (int b = 0; b < fy; ++b) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + ((size_t)(z * fz + a) * ny + (size_t)(y * fy + b)) * nx) + c);
        s0 += (q.x & 0xFFFFu) + (q.x >> 16); s1 += (q.y & 0xFFFFu) + (q.y >> 16);
        s2 += (q.z & 0xFFFFu) + (q.z >> 16); s3 += (q.w & 0xFFFFu) + (q.w >> 16);
      }
    const unsigned r0 = div_small(2 * s0 + num, 2 * num, inv), r1 = div_small(2 * s1 + num, 2 * num, inv);
    const unsigned r2 = div_small(2 * s2 + num, 2 * num, inv), r3 = div_small(2 * s3 + num, 2 * num, inv);
    reinterpret_cast<uint2*>(out + ((size_t)z * oy + y) * ox)[c] = make_uint2(r0 | (r1 << 16), r2 | (r3 << 16));
  }
}

static int stream_grid(size_t items) {
  const size_t want = (items + 255) / 256, cap = (size_t)sm_count() * 8;   // 8 CTAs of 256 threads per SM
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace mvf

using namespace mvf;

extern "C" int mvf_subsample(const void* in, int dtype, const int32_t in_dims[3], const int32_t step[3], void* out, void* stream) {
  MVF_CHECK_ARG(in && out && in_dims && step, "NULL argument");
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  for (int d = 0; d < 3; ++d) MVF_CHECK_ARG(in_dims[d] > 0 && step[d] >= 1, "dims must be positive and steps >= 1");
  const int oz = (in_dims[0] + step[0] - 1) / step[0], oy = (in_dims[1] + step[1] - 1) / step[1], ox = (in_dims[2] + step[2] - 1) / step[2];
  const int grid = stream_grid((size_t)oz * oy * ox);
  cudaStream_t st = as_stream(stream);
  const bool al = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (dtype == MVF_U16 && step[2] == 2 && in_dims[2] % 16 == 0 && al)
    subsample_x2_u16_kernel<<<stream_grid((size_t)oz * oy * (in_dims[2] / 8)), 256, 0, st>>>((const uint16_t*)in, (uint16_t*)out, in_dims[1], in_dims[2], oz, oy, ox, step[0], step[1]);
  else if (dtype == MVF_U16)
    subsample_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)in, (uint16_t*)out, in_dims[0], in_dims[1], in_dims[2], oz, oy, ox, step[0], step[1], step[2]);
  else
    subsample_kernel<float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, in_dims[0], in_dims[1], in_dims[2], oz, oy, ox, step[0], step[1], step[2]);
  MVF_LAUNCHED();
  return MVF_OK;
}

extern "C" int mvf_range_u16(const uint16_t* in, size_t n, uint32_t* minmax, void* stream) {
  MVF_CHECK_ARG(in && minmax && n > 0, "NULL argument or empty input");
  MVF_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15u) == 0, "input must be 16-byte aligned");
  const uint32_t init[2] = {0xFFFFu, 0u};
  MVF_CUDA(cudaMemcpyAsync(minmax, init, sizeof(init), cudaMemcpyHostToDevice, as_stream(stream)));
  range_u16_kernel<<<stream_grid(n / 8 + 1), 256, 0, as_stream(stream)>>>(in, n, minmax);
  MVF_LAUNCHED();
  return MVF_OK;
}

extern "C" int mvf_hist_u16(const uint16_t* in, size_t n, const uint8_t* lut, uint64_t* counts, void* stream) {
  MVF_CHECK_ARG(in && lut && counts && n > 0, "NULL argument or empty input");
  MVF_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15u) == 0 && (reinterpret_cast<uintptr_t>(lut) & 15u) == 0, "input and table must be 16-byte aligned");
  const size_t smem = 65536 + 8 * HREP * 256 * sizeof(unsigned);
  MVF_CUDA(cudaFuncSetAttribute(hist_u16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MVF_CUDA(cudaMemsetAsync(counts, 0, 256 * sizeof(uint64_t), as_stream(stream)));
  const size_t want = (n / 8 + 255) / 256 + 1, cap = (size_t)sm_count() * 3;   // 72 KB of shared memory: 3 CTAs per SM
  hist_u16_kernel<<<(int)(want < cap ? want : cap), 256, smem, as_stream(stream)>>>(in, n, lut, reinterpret_cast<unsigned long long*>(counts));
  MVF_LAUNCHED();
  return MVF_OK;
}

extern "C" int mvf_background_u16(const uint16_t* in, uint16_t* out, size_t n, uint32_t level, void* stream) {
  MVF_CHECK_ARG(in && out, "NULL argument");
  MVF_CHECK_ARG(level <= 65535u, "background level must fit uint16");
  MVF_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0, "buffers must be 16-byte aligned");
  if (n == 0) return MVF_OK;
  background_u16_kernel<<<stream_grid(n / 8 + 1), 256, 0, as_stream(stream)>>>(in, out, n, level);
  MVF_LAUNCHED();
  return MVF_OK;
}

extern "C" int mvf_bin_shrink(const void* in, int dtype, const int32_t in_dims[3], const int32_t factors[3], void* out, void* stream) {
  MVF_CHECK_ARG(in && out && in_dims && factors, "NULL argument");
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  int o[3];
  for (int d = 0; d < 3; ++d) {
    MVF_CHECK_ARG(factors[d] >= 1 && in_dims[d] >= factors[d], "bin factors must be >= 1 and <= the extent");
    o[d] = in_dims[d] / factors[d];
  }
  const int grid = stream_grid((size_t)o[0] * o[1] * o[2]);
  cudaStream_t st = as_stream(stream);
  const bool al = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (dtype == MVF_U16 && factors[2] == 2 && in_dims[2] % 16 == 0 && al && factors[0] * factors[1] <= 4096)   // div_small is exact below 2^31: 2 * 65535 * 2 * fz * fy must stay there
    bin_shrink_x2_u16_kernel<<<stream_grid((size_t)o[0] * o[1] * (o[2] / 4)), 256, 0, st>>>((const uint16_t*)in, (uint16_t*)out, in_dims[1], in_dims[2], o[0], o[1], o[2], factors[0], factors[1]);
  else if (dtype == MVF_U16)
    bin_shrink_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)in, (uint16_t*)out, in_dims[1], in_dims[2], o[0], o[1], o[2], factors[0], factors[1], factors[2]);
  else
    bin_shrink_kernel<float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, in_dims[1], in_dims[2], o[0], o[1], o[2], factors[0], factors[1], factors[2]);
  MVF_LAUNCHED();
  return MVF_OK;
}

// ---- strided host -> device copy of a source brick ---------------------------------------------------------------
// The reference crops every view to the bounding box a chunk touches before resampling (mv:1552-1601); a z-slab rank
// does the same with its slab.  One cudaMemcpy3DAsync moves view[lo:hi] out of the caller's array (page-locked: a
// single DMA with a row pitch; pageable: staged by the driver) into a dense device brick.
extern "C" int mvf_copy_brick_h2d(const void* host, int dtype, const int32_t dims[3], const int32_t lo[3],
                                  const int32_t hi[3], void* dst, void* stream) {
  MVF_CHECK_ARG(host && dims && lo && hi && dst, "NULL argument");
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  for (int d = 0; d < 3; ++d) MVF_CHECK_ARG(lo[d] >= 0 && hi[d] > lo[d] && hi[d] <= dims[d], "brick must lie inside the view");
  const size_t esz = dtype == MVF_U16 ? 2 : 4;
  const size_t w = (size_t)(hi[2] - lo[2]), h = (size_t)(hi[1] - lo[1]), dep = (size_t)(hi[0] - lo[0]);
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(host), (size_t)dims[2] * esz, (size_t)dims[2] * esz, (size_t)dims[1]);
  p.srcPos = make_cudaPos((size_t)lo[2] * esz, (size_t)lo[1], (size_t)lo[0]);
  p.dstPtr = make_cudaPitchedPtr(dst, w * esz, w * esz, h);
  p.dstPos = make_cudaPos(0, 0, 0);
  p.extent = make_cudaExtent(w * esz, h, dep);
  p.kind = cudaMemcpyHostToDevice;
  MVF_CUDA(cudaMemcpy3DAsync(&p, as_stream(stream)));
  return MVF_OK;
}

// ---- dual-illumination fusion, plane by plane (illumination_fusion_planewise, mv:391-467) ------------------------------
// pair: (2, P, n, n) uint16 - both light sheets, planes already oriented so that the sheet travels along rows (axis -2).
//   mask = log(I0 + I1 - min(I0, I1 over the plane pair)), rescaled to [0, 1] with the plane's own extremes      (float64)
//   total[x] = sum_y mask ; rows whose index r has total[r] == 0 are set to 1 (the reference selects ROWS with the per-column
//   total - square planes), total == 0 -> n ; right[y][x] = cumsum_y(mask)[y][x] > total[y] / 2 (its transposes broadcast the
//   per-column total along the row index as well)                                                                  (0 / 1)
//   right = gaussian_filter(float32(right), sigma = n / 50 per axis)   (scipy: reflect, radius int(4 sigma + .5), per-axis
//   passes in float64 with a float32 result each; taps from the host, accumulated centre first, then pairs from far to near)
//   out = uint16(I0 * (1 - right) + I1 * right)                                                                  (float32)
namespace mvf {

struct IllumStats { unsigned m, smin, smax, pad; };

__global__ void __launch_bounds__(256) illum_stats_kernel(const uint16_t* __restrict__ pair, size_t plane, size_t pstride2, IllumStats* stats) {
  const uint16_t* p0 = pair + (size_t)blockIdx.x * plane;
  const uint16_t* p1 = p0 + pstride2;
  unsigned m = 0xFFFFFFFFu, smin = 0xFFFFFFFFu, smax = 0u;
  for (size_t i = threadIdx.x; i < plane; i += 256) {
    const unsigned a = p0[i], b = p1[i];
    m = min(m, min(a, b));
    smin = min(smin, a + b);
    smax = max(smax, a + b);
  }
  __shared__ unsigned sm[3][8];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    m = min(m, __shfl_xor_sync(0xffffffffu, m, s));
    smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, s));
    smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, s));
  }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = m; sm[1][threadIdx.x >> 5] = smin; sm[2][threadIdx.x >> 5] = smax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { m = min(m, sm[0][w]); smin = min(smin, sm[1][w]); smax = max(smax, sm[2][w]); }
    stats[blockIdx.x] = IllumStats{m, smin, smax, 0u};
  }
}

__device__ __forceinline__ double illum_mask(unsigned s, unsigned m, double lmin, double inv_den_num) {
  // (log(s - m) - lmin) / (lmax - lmin): the division is done by the caller's denominator, as numpy does it
  return __ddiv_rn(__dsub_rn(log((double)(s - m)), lmin), inv_den_num);
}

// thread = column x of plane p.  PASS 0: total[p][x].  PASS 1: right[p][y][x] (needs every total of the plane).
template <int PASS>
__global__ void __launch_bounds__(128) illum_columns_kernel(const uint16_t* __restrict__ pair, int n, size_t pstride2,
                                                            const IllumStats* __restrict__ stats, double* total, float* right) {
  const int x = blockIdx.x * 128 + threadIdx.x, p = blockIdx.y;
  if (x >= n) return;
  const size_t plane = (size_t)n * n;
  const uint16_t* p0 = pair + (size_t)p * plane + x;
  const uint16_t* p1 = p0 + pstride2;
  const IllumStats st = stats[p];
  const double lmin = log((double)(st.smin - st.m)), lmax = log((double)(st.smax - st.m));
  const double den = __dsub_rn(lmax, lmin);
  double* tot = total + (size_t)p * n;
  if (PASS == 0) {
    double acc = 0.0;
    for (int y = 0; y < n; ++y) acc = __dadd_rn(acc, illum_mask((unsigned)p0[(size_t)y * n] + p1[(size_t)y * n], st.m, lmin, den));
    tot[x] = acc;
  } else {
    // (cumsum.T > total.T / 2).T broadcasts the per-column total along the ROW index: the threshold of row y is total[y] / 2
    float* r = right + (size_t)p * plane + x;
    double cum = 0.0;
    for (int y = 0; y < n; ++y) {
      const double traw = tot[y];
      const double mv = traw == 0.0 ? 1.0 : illum_mask((unsigned)p0[(size_t)y * n] + p1[(size_t)y * n], st.m, lmin, den);
      cum = __dadd_rn(cum, mv);
      r[(size_t)y * n] = cum > __ddiv_rn(traw == 0.0 ? (double)n : traw, 2.0) ? 1.f : 0.f;
    }
  }
}

// one Gaussian pass along AXIS (0: rows / y, 1: columns / x) of every plane; scipy's 'reflect' extension (edge repeated)
template <int AXIS>
__global__ void __launch_bounds__(256) illum_gauss_kernel(const float* __restrict__ in, float* __restrict__ out, int n, size_t total_px,
                                                          const double* __restrict__ w, int r) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total_px) return;
  const int x = (int)(i % n), y = (int)((i / n) % n);
  const int c = AXIS == 0 ? y : x;
  const float* base = in + (i - (AXIS == 0 ? (size_t)y * n : (size_t)x));
  const size_t stride = AXIS == 0 ? (size_t)n : 1;
  auto at = [&](int q) {
    q = q < 0 ? -q - 1 : (q >= n ? 2 * n - 1 - q : q);
    return (double)base[(size_t)q * stride];
  };
  double tmp = __dmul_rn(at(c), w[r]);
  for (int ii = -r; ii < 0; ++ii) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(c + ii), at(c - ii)), w[ii + r]));
  out[i] = (float)tmp;
}

__global__ void __launch_bounds__(256) illum_blend_kernel(const uint16_t* __restrict__ pair, size_t pstride2, const float* __restrict__ right,
                                                          uint16_t* __restrict__ out, size_t total_px) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total_px) return;
  const float rw = right[i];
  const float a = __fmul_rn((float)pair[i], __fsub_rn(1.f, rw)), b = __fmul_rn((float)pair[i + pstride2], rw);
  out[i] = (uint16_t)__fadd_rn(a, b);
}

}  // namespace mvf

extern "C" int mvf_illum_fuse_planes(const uint16_t* pair, int32_t nplanes, int32_t n, const double* dev_taps, int32_t radius,
                                     void* scratch, size_t scratch_bytes, uint16_t* out, void* stream) {
  MVF_CHECK_ARG(pair && dev_taps && scratch && out, "NULL argument");
  MVF_CHECK_ARG(nplanes > 0 && nplanes <= 65535 && n > 1 && radius >= 0 && radius < n, "bad plane count, extent or radius");
  const size_t plane = (size_t)n * n, px = plane * nplanes;
  const size_t need = px * 4 * 2 + (size_t)nplanes * n * 8 + (size_t)nplanes * sizeof(IllumStats) + 64;
  MVF_CHECK_ARG(scratch_bytes >= need, "scratch too small (2 float planes per input plane + n doubles + 16 bytes per plane + 64)");
  cudaStream_t st = as_stream(stream);
  char* sp = static_cast<char*>(scratch);
  float* right = reinterpret_cast<float*>(sp);
  float* tmp = right + px;
  double* total = reinterpret_cast<double*>(tmp + px);
  IllumStats* stats = reinterpret_cast<IllumStats*>(total + (size_t)nplanes * n);
  illum_stats_kernel<<<nplanes, 256, 0, st>>>(pair, plane, px, stats);
  MVF_LAUNCHED();
  dim3 cg((n + 127) / 128, nplanes);
  illum_columns_kernel<0><<<cg, 128, 0, st>>>(pair, n, px, stats, total, right);
  MVF_LAUNCHED();
  illum_columns_kernel<1><<<cg, 128, 0, st>>>(pair, n, px, stats, total, right);
  MVF_LAUNCHED();
  const unsigned gb = (unsigned)((px + 255) / 256);
  illum_gauss_kernel<0><<<gb, 256, 0, st>>>(right, tmp, n, px, dev_taps, radius);
  MVF_LAUNCHED();
  illum_gauss_kernel<1><<<gb, 256, 0, st>>>(tmp, right, n, px, dev_taps, radius);
  MVF_LAUNCHED();
  illum_blend_kernel<<<gb, 256, 0, st>>>(pair, px, right, out, px);
  MVF_LAUNCHED();
  return MVF_OK;
}
