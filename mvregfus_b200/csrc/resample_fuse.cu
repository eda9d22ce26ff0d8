// K1 (affine resampling of one view) and K2 (fused resample + blending/DCT weights + normalise + weighted
// sum) for sm_100a.  Reference arithmetic: mvregfus/multiview.py:1458-1634 (scipy path), 6797-6877 (ITK
// path), 3469-3633 (blending weights), 4291-4317/4434-4439 (DCT weight upsampling), 4904-4912 (fusion).
//
// Work decomposition: one CTA per 8x8x32 (z,y,x) output tile, 256 threads, each thread owns a run of 8
// consecutive x voxels (16-byte uint16 / 2x16-byte float32 stores).  For every view the CTA stages the
// source brick hit by the tile (bounding box of the tile's affine footprint) in shared memory with
// coalesced row loads, then gathers the 8 trilinear taps from shared memory, so the access pattern of a
// rotated view never reaches L1/L2.  Blending weights are analytic (three 200-entry 1-D profiles per view)
// and stay in registers until the sum over views is known; nothing but the views is read from HBM and
// nothing but the fused volume is written.
#include "common.cuh"

namespace mvf {

constexpr int TZ = 8, TY = 8, TX = 32, RX = 8, NT = 256;
static_assert(TZ * TY * (TX / RX) == NT, "tile/thread mapping");

struct KView {
  const void* data;
  const float* prof;
  const float* cgrid;
  double m[9], o[3];
  double wm[9], wo[3];
  double cs[3], co[3];
  int n[3];
  int cd[3];
  int one_lo[3], one_hi[3];
  int wmode;
  int be[3];  // host-computed upper bound of the brick extents (shared-memory sizing)
};

template <int V>
struct FuseArgs {
  KView v[V];
  uint16_t* out;
  float* outf;
  int dims[3];
  int org[3];
};

struct ResampleArgs {
  KView v;
  void* dst;
  int dims[3];
  int org[3];
};

struct Brick {
  int b0[3];  // view index of brick element (0,0,0); may be negative (zero filled)
  int e[3];   // extents
  int px, pz; // shared-memory pitches (elements)
  bool skip;      // brick does not intersect the view
  bool interior;  // every voxel of the tile maps strictly inside [0, n-1]: no per-voxel boundary test
};

// Bounding box, in the index space of an affine map, of the voxels of one tile (fp64).
__device__ __forceinline__ void tile_bbox(const double* m, const double* o, int z0, int y0, int x0, double lo[3], double hi[3]) {
  const int T[3] = {TZ - 1, TY - 1, TX - 1};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double base = fma(m[d * 3 + 2], (double)x0, fma(m[d * 3 + 1], (double)y0, fma(m[d * 3 + 0], (double)z0, o[d])));
    double l = base, h = base;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double ext = m[d * 3 + k] * (double)T[k];
      if (ext < 0) l += ext; else h += ext;
    }
    lo[d] = l - 1e-7 * (1.0 + fabs(l));
    hi[d] = h + 1e-7 * (1.0 + fabs(h));
  }
}

__device__ __forceinline__ Brick tile_brick(const KView& kv, int z0, int y0, int x0) {
  Brick b;
  double lo[3], hi[3];
  tile_bbox(kv.m, kv.o, z0, y0, x0, lo, hi);
  bool skip = false, interior = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    int i0 = __double2int_rd(lo[d]);
    int i1 = __double2int_rd(hi[d]) + 1;
    b.b0[d] = i0;
    int e = i1 - i0 + 1;
    b.e[d] = e < kv.be[d] ? e : kv.be[d];
    if (i1 < 0 || i0 > kv.n[d] - 1) skip = true;
    if (!(lo[d] >= 0.0 && hi[d] <= (double)(kv.n[d] - 1))) interior = false;
  }
  b.px = b.e[2] | 1;
  b.pz = (b.e[1] * b.px) | 1;
  b.skip = skip;
  b.interior = interior;
  return b;
}

// Cooperative, coalesced staging of the brick into shared memory as float32; out-of-volume elements are 0.
template <typename T>
__device__ __forceinline__ void stage_brick(const KView& kv, const Brick& bk, float* sm) {
  const T* __restrict__ src = reinterpret_cast<const T*>(kv.data);
  const unsigned ex = bk.e[2], ey = bk.e[1];
  const unsigned total = bk.e[0] * ey * ex;
  const unsigned mx = 0xFFFFFFFFu / ex + 1u;   // exact quotient for i*ex < 2^32
  const unsigned my = 0xFFFFFFFFu / ey + 1u;
  for (unsigned i = threadIdx.x; i < total; i += NT) {
    unsigned row = __umulhi(i, mx);
    unsigned ix = i - row * ex;
    unsigned iz = __umulhi(row, my);
    unsigned iy = row - iz * ey;
    int gz = bk.b0[0] + (int)iz, gy = bk.b0[1] + (int)iy, gx = bk.b0[2] + (int)ix;
    float v = 0.f;
    if ((unsigned)gz < (unsigned)kv.n[0] && (unsigned)gy < (unsigned)kv.n[1] && (unsigned)gx < (unsigned)kv.n[2])
      v = to_float<T>(__ldg(src + ((size_t)gz * kv.n[1] + gy) * kv.n[2] + gx));
    sm[iz * bk.pz + iy * bk.px + ix] = v;
  }
}

// Nearest grid point r = rint(u), |u - r| as float32 and the sign of (u - r), from one fp64 coordinate, with
// two DADDs and one F2F.  Interpolating from the NEAREST corner towards the far one keeps the lerp weight in
// [0, 0.5], so  near + (far - near) * w  is exact for constant neighbourhoods and relatively accurate for
// non-negative data (no cancellation against a weight close to 1).
__device__ __forceinline__ void split_nearest(double u, int& r, float& w, int& neg) {
  double y = u + MAGIC64;
  r = __double2loint(y);
  double d = u - (y - MAGIC64);
  neg = d < 0.0 ? -1 : 0;  // -1 when the far corner is r-1, 0 when it is r+1 (-0.0 counts as +0)
  w = fabsf((float)d);
}

__device__ __forceinline__ float lerp8(const float* p, int ox, int oy, int oz, float wx, float wy, float wz) {
  float c000 = p[0], c001 = p[ox];
  float c010 = p[oy], c011 = p[oy + ox];
  float c100 = p[oz], c101 = p[oz + ox];
  float c110 = p[oz + oy], c111 = p[oz + oy + ox];
  float x00 = fmaf(c001 - c000, wx, c000);
  float x01 = fmaf(c011 - c010, wx, c010);
  float x10 = fmaf(c101 - c100, wx, c100);
  float x11 = fmaf(c111 - c110, wx, c110);
  float y0 = fmaf(x01 - x00, wy, x00);
  float y1 = fmaf(x11 - x10, wy, x10);
  return fmaf(y1 - y0, wz, y0);
}

// Trilinear value at brick-relative coordinate rel (float64), scipy 'constant' rule: 0 when the view
// coordinate is outside [0, n-1] (ni_interpolation.c NI_GeometricTransform, mode constant).
template <bool INTERIOR>
__device__ __forceinline__ float gather_scipy(const float* sm, const Brick& bk, const KView& kv, const double rel[3]) {
  int r[3], neg[3];
  float w[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) split_nearest(rel[d], r[d], w[d], neg[d]);
  const float* p = sm + r[0] * bk.pz + r[1] * bk.px + r[2];
  const int ox = 1 | neg[2];
  const int oy = (bk.px ^ neg[1]) - neg[1];
  const int oz = (bk.pz ^ neg[0]) - neg[0];
  float val = lerp8(p, ox, oy, oz, w[2], w[1], w[0]);
  if (!INTERIOR) {
    bool inside = true;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      inside = inside && (rel[d] >= -(double)bk.b0[d]) && (rel[d] <= (double)(kv.n[d] - 1 - bk.b0[d]));
    val = inside ? val : 0.f;
  }
  return val;
}

// ITK rule: valid iff -0.5 <= u < n-0.5 per axis, neighbours clamped to the edge, outside -> 0
// (itk::LinearInterpolateImageFunction / ResampleImageFilter).
__device__ __forceinline__ float gather_itk(const float* sm, const Brick& bk, const KView& kv, const double rel[3]) {
  int off[3], base = 0;
  float w[3];
  bool inside = true;
  const int stride[3] = {bk.pz, bk.px, 1};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double u = rel[d] + (double)bk.b0[d];
    inside = inside && (u >= -0.5) && (u < (double)kv.n[d] - 0.5);
    int r, neg;
    split_nearest(u, r, w[d], neg);
    int near = max(0, min(r, kv.n[d] - 1));
    int far = max(0, min(r + (1 | neg), kv.n[d] - 1));
    int ln = max(0, min(near - bk.b0[d], bk.e[d] - 1));
    int lf = max(0, min(far - bk.b0[d], bk.e[d] - 1));
    base += ln * stride[d];
    off[d] = (lf - ln) * stride[d];
  }
  float val = lerp8(sm + base, off[2], off[1], off[0], w[2], w[1], w[0]);
  return inside ? val : 0.f;
}

// ---- blending field ------------------------------------------------------------------------------
enum { WT_ZERO = 0, WT_ONE = 1, WT_RAMP = 2, WT_EDGE = 3 };

// Classify a tile against a view's blending field: every voxel outside the field (weight 0), every voxel
// on the all-ones plateau (weight 1), fully inside the field (fast float32 path) or touching its rim.
__device__ __forceinline__ int classify_tile(const KView& kv, int z0, int y0, int x0) {
  if (kv.wmode == MVF_W_OFF) return WT_ZERO;
  double lo[3], hi[3];
  tile_bbox(kv.wm, kv.wo, z0, y0, x0, lo, hi);
  bool zero = false, one = true, ramp = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    zero = zero || (hi[d] < -0.5) || (lo[d] >= (double)MVF_SIG_N - 0.5);
    one = one && (lo[d] >= (double)kv.one_lo[d]) && (hi[d] < (double)kv.one_hi[d]);
    ramp = ramp && (lo[d] >= 0.0) && (hi[d] < (double)(MVF_SIG_N - 1));
  }
  return zero ? WT_ZERO : (one ? WT_ONE : (ramp ? WT_RAMP : WT_EDGE));
}

__device__ __forceinline__ float field_lerp(const float* f, int i0, int i1, int i2, int j0, int j1, int j2, float t0, float t1, float t2) {
  float a0 = __ldg(f + i0), a1 = __ldg(f + j0);
  float b0 = __ldg(f + MVF_SIG_N + i1), b1 = __ldg(f + MVF_SIG_N + j1);
  float c0 = __ldg(f + 2 * MVF_SIG_N + i2), c1 = __ldg(f + 2 * MVF_SIG_N + j2);
  float m00 = fminf(a0, b0), m01 = fminf(a0, b1), m10 = fminf(a1, b0), m11 = fminf(a1, b1);
  float v000 = fminf(m00, c0), v001 = fminf(m00, c1), v010 = fminf(m01, c0), v011 = fminf(m01, c1);
  float v100 = fminf(m10, c0), v101 = fminf(m10, c1), v110 = fminf(m11, c0), v111 = fminf(m11, c1);
  float x00 = fmaf(v001 - v000, t2, v000);
  float x01 = fmaf(v011 - v010, t2, v010);
  float x10 = fmaf(v101 - v100, t2, v100);
  float x11 = fmaf(v111 - v110, t2, v110);
  float y0 = fmaf(x01 - x00, t1, x00);
  float y1 = fmaf(x11 - x10, t1, x10);
  return fmaf(y1 - y0, t0, y0);
}

// ITK-linear sample of the analytic blending field min(f_z[i], f_y[j], f_x[k]) at continuous field index c
// (exact float64 index arithmetic; used on tiles that touch the rim of the field).
__device__ __forceinline__ float blend_field_exact(const KView& kv, const double c[3]) {
  int i0[3];
  float t[3];
  bool inside = true, ones = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    inside = inside && (c[d] >= -0.5) && (c[d] < (double)MVF_SIG_N - 0.5);
    int i = __double2int_rd(c[d]);
    double fr = c[d] - (double)i;
    if (i < 0) { i = 0; fr = 0.0; }
    if (i >= MVF_SIG_N - 1) { i = MVF_SIG_N - 1; fr = 0.0; }
    i0[d] = i;
    t[d] = (float)fr;
    ones = ones && (i >= kv.one_lo[d]) && (i < kv.one_hi[d]);
  }
  if (!inside) return 0.f;
  if (ones) return 1.f;
  int j0 = min(i0[0] + 1, MVF_SIG_N - 1), j1 = min(i0[1] + 1, MVF_SIG_N - 1), j2 = min(i0[2] + 1, MVF_SIG_N - 1);
  return field_lerp(kv.prof, i0[0], i0[1], i0[2], j0, j1, j2, t[0], t[1], t[2]);
}

// ITK-linear sample of a small float32 grid in global memory (coarse DCT weights, mv:4304-4311).
__device__ __forceinline__ float coarse_sample(const KView& kv, const double c[3]) {
  int i0[3], i1[3];
  float t[3];
  bool inside = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    inside = inside && (c[d] >= -0.5) && (c[d] < (double)kv.cd[d] - 0.5);
    int i = __double2int_rd(c[d]);
    double fr = c[d] - (double)i;
    if (i < 0) { i = 0; fr = 0.0; }
    if (i >= kv.cd[d] - 1) { i = kv.cd[d] - 1; fr = 0.0; }
    i0[d] = i;
    i1[d] = min(i + 1, kv.cd[d] - 1);
    t[d] = (float)fr;
  }
  if (!inside) return 0.f;
  const float* g = kv.cgrid;
  const int sy = kv.cd[2], sz = kv.cd[1] * kv.cd[2];
  float c000 = __ldg(g + i0[0] * sz + i0[1] * sy + i0[2]), c001 = __ldg(g + i0[0] * sz + i0[1] * sy + i1[2]);
  float c010 = __ldg(g + i0[0] * sz + i1[1] * sy + i0[2]), c011 = __ldg(g + i0[0] * sz + i1[1] * sy + i1[2]);
  float c100 = __ldg(g + i1[0] * sz + i0[1] * sy + i0[2]), c101 = __ldg(g + i1[0] * sz + i0[1] * sy + i1[2]);
  float c110 = __ldg(g + i1[0] * sz + i1[1] * sy + i0[2]), c111 = __ldg(g + i1[0] * sz + i1[1] * sy + i1[2]);
  float x00 = fmaf(c001 - c000, t[2], c000);
  float x01 = fmaf(c011 - c010, t[2], c010);
  float x10 = fmaf(c101 - c100, t[2], c100);
  float x11 = fmaf(c111 - c110, t[2], c110);
  float y0 = fmaf(x01 - x00, t[1], x00);
  float y1 = fmaf(x11 - x10, t[1], x10);
  return fmaf(y1 - y0, t[0], y0);
}

__device__ __forceinline__ void affine_base(const double* m, const double* o, int z, int y, int x, double out[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d)
    out[d] = fma(m[d * 3 + 2], (double)x, fma(m[d * 3 + 1], (double)y, fma(m[d * 3 + 0], (double)z, o[d])));
}

// Blending (x coarse DCT) weights of the RX voxels of a thread for all V views, normalised exactly like
// the reference: w_v / sum_v w_v in float32 with a 0 -> 1 guard (mv:3627-3631), and for the DCT mode
// coarse_v * that, normalised again (mv:4317, 4434-4439).  (tz0,ty0,tx0) is the fused-frame index of the
// tile origin, (gz,gy,gx) of the thread's first voxel.
template <int V>
__device__ __forceinline__ void thread_weights(const KView* kvs, int tz0, int ty0, int tx0, int gz, int gy, int gx,
                                               float (&w)[V][RX]) {
  float wsum[RX];
#pragma unroll
  for (int k = 0; k < RX; ++k) wsum[k] = 0.f;
  bool dct = false;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const KView& kv = kvs[v];
    const int cls = classify_tile(kv, tz0, ty0, tx0);  // CTA-uniform
    if (cls == WT_ZERO) {
#pragma unroll
      for (int k = 0; k < RX; ++k) w[v][k] = 0.f;
      continue;
    }
    dct = dct || (kv.wmode == MVF_W_BLEND_DCT);
    if (cls == WT_ONE) {
#pragma unroll
      for (int k = 0; k < RX; ++k) { w[v][k] = 1.f; wsum[k] += 1.f; }
      continue;
    }
    double cb[3];
    affine_base(kv.wm, kv.wo, gz, gy, gx, cb);
    if (cls == WT_RAMP) {
      // float32 increments on a float64 base: |error| < 1e-6 field voxels, i.e. < 3e-7 in the weight
      int ci[3];
      float cf[3], cm[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ci[d] = __double2int_rd(cb[d]);
        cf[d] = (float)(cb[d] - (double)ci[d]);
        cm[d] = (float)kv.wm[d * 3 + 2];
      }
#pragma unroll
      for (int k = 0; k < RX; ++k) {
        int i0[3];
        float t[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float c = fmaf((float)k, cm[d], cf[d]);
          float y = __fadd_rd(c, MAGIC32);
          t[d] = c - (y - MAGIC32);
          i0[d] = ci[d] + (__float_as_int(y) - 0x4B400000);
          i0[d] = max(0, min(i0[d], MVF_SIG_N - 2));  // rounding guard at the bbox limits
        }
        float wv = field_lerp(kv.prof, i0[0], i0[1], i0[2], i0[0] + 1, i0[1] + 1, i0[2] + 1, t[0], t[1], t[2]);
        w[v][k] = wv;
        wsum[k] += wv;
      }
    } else {
#pragma unroll
      for (int k = 0; k < RX; ++k) {
        double c[3] = {fma((double)k, kv.wm[2], cb[0]), fma((double)k, kv.wm[5], cb[1]), fma((double)k, kv.wm[8], cb[2])};
        w[v][k] = blend_field_exact(kv, c);
        wsum[k] += w[v][k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < RX; ++k) {
    float d = wsum[k] == 0.f ? 1.f : wsum[k];
#pragma unroll
    for (int v = 0; v < V; ++v) w[v][k] = __fdiv_rn(w[v][k], d);
  }
  if (dct) {
#pragma unroll
    for (int k = 0; k < RX; ++k) wsum[k] = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const KView& kv = kvs[v];
      if (kv.wmode != MVF_W_BLEND_DCT) continue;
#pragma unroll
      for (int k = 0; k < RX; ++k) {
        float cw = 0.f;
        if (w[v][k] != 0.f) {
          double c[3] = {fma(kv.cs[0], (double)gz, kv.co[0]), fma(kv.cs[1], (double)gy, kv.co[1]),
                         fma(kv.cs[2], (double)(gx + k), kv.co[2])};
          cw = coarse_sample(kv, c);
        }
        w[v][k] = cw * w[v][k];
        wsum[k] += w[v][k];
      }
    }
#pragma unroll
    for (int k = 0; k < RX; ++k) {
      float d = wsum[k] == 0.f ? 1.f : wsum[k];
#pragma unroll
      for (int v = 0; v < V; ++v) w[v][k] = __fdiv_rn(w[v][k], d);
    }
  }
}

__device__ __forceinline__ void thread_voxel(int& lz, int& ly, int& lx) {
  const int t = threadIdx.x;
  lx = (t & 3) * RX;
  ly = (t >> 2) & 7;
  lz = t >> 5;
}

// ------------------------------------------------------------------------------------------------
// K2: fused single pass
// ------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(NT) fuse_blend_kernel(const __grid_constant__ FuseArgs<V> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  const int z0 = blockIdx.z * TZ, y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
  int lz, ly, lx;
  thread_voxel(lz, ly, lx);
  const int z = z0 + lz, y = y0 + ly, x = x0 + lx;              // local (box) index
  const int gz = z + A.org[0], gy = y + A.org[1], gx = x + A.org[2];  // fused-frame index

  float w[V][RX];
  thread_weights<V>(A.v, z0 + A.org[0], y0 + A.org[1], x0 + A.org[2], gz, gy, gx, w);

  uint32_t acc[RX];
  float accf[RX];
#pragma unroll
  for (int k = 0; k < RX; ++k) { acc[k] = 0u; accf[k] = 0.f; }

#pragma unroll
  for (int v = 0; v < V; ++v) {
    const KView& kv = A.v[v];
    const Brick bk = tile_brick(kv, z0 + A.org[0], y0 + A.org[1], x0 + A.org[2]);
    if (bk.skip) continue;  // CTA-uniform
    __syncthreads();
    stage_brick<T>(kv, bk, sm);
    __syncthreads();
    double ub[3];
    affine_base(kv.m, kv.o, gz, gy, gx, ub);
#pragma unroll
    for (int d = 0; d < 3; ++d) ub[d] -= (double)bk.b0[d];
#pragma unroll
    for (int k = 0; k < RX; ++k) {
      double rel[3] = {fma((double)k, kv.m[2], ub[0]), fma((double)k, kv.m[5], ub[1]), fma((double)k, kv.m[8], ub[2])};
      float val = bk.interior ? gather_scipy<true>(sm, bk, kv, rel) : gather_scipy<false>(sm, bk, kv, rel);
      if (sizeof(T) == 2) val = u16_to_float(round_half_up_u16(val));  // the transformed view is uint16 (mv:1626)
      acc[k] += trunc_product_u16(w[v][k], val);
      accf[k] = fmaf(w[v][k], val, accf[k]);
    }
  }

  if (z >= A.dims[0] || y >= A.dims[1]) return;
  const size_t row = ((size_t)z * A.dims[1] + y) * A.dims[2];
  uint32_t r[RX];
#pragma unroll
  for (int k = 0; k < RX; ++k) r[k] = sizeof(T) == 2 ? (acc[k] & 0xFFFFu) : min(acc[k], 65535u);
  if (x + RX <= A.dims[2] && (((row + x) & 7) == 0)) {
    uint4 pk;
    pk.x = r[0] | (r[1] << 16); pk.y = r[2] | (r[3] << 16); pk.z = r[4] | (r[5] << 16); pk.w = r[6] | (r[7] << 16);
    *reinterpret_cast<uint4*>(A.out + row + x) = pk;
  } else {
#pragma unroll
    for (int k = 0; k < RX; ++k) if (x + k < A.dims[2]) A.out[row + x + k] = (uint16_t)r[k];
  }
  if (A.outf) {
#pragma unroll
    for (int k = 0; k < RX; ++k) if (x + k < A.dims[2]) A.outf[row + x + k] = accf[k];
  }
}

// ------------------------------------------------------------------------------------------------
// K1: one view, same dtype out
// ------------------------------------------------------------------------------------------------
template <typename T, int RULE>
__global__ void __launch_bounds__(NT) resample_kernel(const __grid_constant__ ResampleArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  const int z0 = blockIdx.z * TZ, y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
  int lz, ly, lx;
  thread_voxel(lz, ly, lx);
  const int z = z0 + lz, y = y0 + ly, x = x0 + lx;
  const int gz = z + A.org[0], gy = y + A.org[1], gx = x + A.org[2];
  const KView& kv = A.v;
  const Brick bk = tile_brick(kv, z0 + A.org[0], y0 + A.org[1], x0 + A.org[2]);
  T r[RX];
#pragma unroll
  for (int k = 0; k < RX; ++k) r[k] = T(0);
  if (!bk.skip) {
    stage_brick<T>(kv, bk, sm);
    __syncthreads();
    double ub[3];
    affine_base(kv.m, kv.o, gz, gy, gx, ub);
#pragma unroll
    for (int d = 0; d < 3; ++d) ub[d] -= (double)bk.b0[d];
#pragma unroll
    for (int k = 0; k < RX; ++k) {
      double rel[3] = {fma((double)k, kv.m[2], ub[0]), fma((double)k, kv.m[5], ub[1]), fma((double)k, kv.m[8], ub[2])};
      float val;
      if (RULE == MVF_RULE_SCIPY) val = bk.interior ? gather_scipy<true>(sm, bk, kv, rel) : gather_scipy<false>(sm, bk, kv, rel);
      else val = gather_itk(sm, bk, kv, rel);
      if (sizeof(T) == 2) r[k] = (T)(RULE == MVF_RULE_SCIPY ? round_half_up_u16(val) : clamp_trunc_u16(val));
      else r[k] = (T)val;
    }
  }
  if (z >= A.dims[0] || y >= A.dims[1]) return;
  T* dst = reinterpret_cast<T*>(A.dst) + ((size_t)z * A.dims[1] + y) * A.dims[2];
#pragma unroll
  for (int k = 0; k < RX; ++k) if (x + k < A.dims[2]) dst[x + k] = r[k];
}

// ------------------------------------------------------------------------------------------------
// blending weights on a grid (get_weights_simple), V volumes out
// ------------------------------------------------------------------------------------------------
template <int V>
struct WeightArgs {
  KView v[V];
  float* out;
  int dims[3];
  int org[3];
  int normalise;
};

template <int V>
__global__ void __launch_bounds__(NT) blend_weights_kernel(const __grid_constant__ WeightArgs<V> A) {
  const int z0 = blockIdx.z * TZ, y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
  int lz, ly, lx;
  thread_voxel(lz, ly, lx);
  const int z = z0 + lz, y = y0 + ly, x = x0 + lx;
  if (z >= A.dims[0] || y >= A.dims[1]) return;
  const int gz = z + A.org[0], gy = y + A.org[1], gx = x + A.org[2];
  float w[V][RX];
  if (A.normalise) {
    thread_weights<V>(A.v, z0 + A.org[0], y0 + A.org[1], x0 + A.org[2], gz, gy, gx, w);
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const KView& kv = A.v[v];
      double cb[3];
      affine_base(kv.wm, kv.wo, gz, gy, gx, cb);
#pragma unroll
      for (int k = 0; k < RX; ++k) {
        double c[3] = {fma((double)k, kv.wm[2], cb[0]), fma((double)k, kv.wm[5], cb[1]), fma((double)k, kv.wm[8], cb[2])};
        w[v][k] = kv.wmode == MVF_W_OFF ? 0.f : blend_field_exact(kv, c);
      }
    }
  }
  const size_t vol = (size_t)A.dims[0] * A.dims[1] * A.dims[2];
  const size_t row = ((size_t)z * A.dims[1] + y) * A.dims[2];
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int k = 0; k < RX; ++k)
      if (x + k < A.dims[2]) A.out[v * vol + row + x + k] = w[v][k];
}

// ------------------------------------------------------------------------------------------------
// fuse_views_weights on materialised inputs (elementwise)
// ------------------------------------------------------------------------------------------------
struct FuseWeightedArgs {
  const void* views[MVF_MAX_VIEWS];
  const float* weights[MVF_MAX_VIEWS];
  uint16_t* out;
  float* outf;
  size_t n;
  int nviews;
  int has_weights;
};

template <typename T>
__global__ void __launch_bounds__(NT) fuse_weighted_kernel(const __grid_constant__ FuseWeightedArgs A) {
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < A.n; i += (size_t)gridDim.x * NT) {
    uint32_t acc = 0;
    float accf = 0.f;
    if (A.has_weights) {
      for (int v = 0; v < A.nviews; ++v) {
        float val = to_float(reinterpret_cast<const T*>(A.views[v])[i]);
        float w = A.weights[v][i];
        acc += trunc_product_u16(w, val);
        accf = fmaf(w, val, accf);
      }
      acc = sizeof(T) == 2 ? (acc & 0xFFFFu) : min(acc, 65535u);
    } else {  // np.mean(views, 0), clip, truncate (mv:4909-4912): float64 mean for uint16 views, float32 for float32
      if (sizeof(T) == 2) {
        double s = 0.0;
        for (int v = 0; v < A.nviews; ++v) s += (double)to_float(reinterpret_cast<const T*>(A.views[v])[i]);
        s /= (double)A.nviews;
        accf = (float)s;
        acc = (uint32_t)fmin(fmax(s, 0.0), 65535.0);
      } else {
        float s = 0.f;
        for (int v = 0; v < A.nviews; ++v) s = __fadd_rn(s, to_float(reinterpret_cast<const T*>(A.views[v])[i]));
        s = __fdiv_rn(s, (float)A.nviews);
        accf = s;
        acc = (uint32_t)fminf(fmaxf(s, 0.f), 65535.f);
      }
    }
    A.out[i] = (uint16_t)acc;
    if (A.outf) A.outf[i] = accf;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int to_kview(const mvf_view& h, KView& k, bool need_data) {
  if (need_data) {
    if (!h.data) return fail(MVF_ERR_INVALID, "%s%s", "mvf_view.data is NULL");
    if (h.dtype != MVF_U16 && h.dtype != MVF_F32) return fail(MVF_ERR_INVALID, "%s%s", "mvf_view.dtype must be MVF_U16 or MVF_F32");
  }
  k.data = h.data;
  k.prof = h.profiles;
  k.cgrid = h.cgrid;
  for (int i = 0; i < 9; ++i) { k.m[i] = h.matrix[i]; k.wm[i] = h.wmatrix[i]; }
  for (int d = 0; d < 3; ++d) {
    k.o[d] = h.offset[d]; k.wo[d] = h.woffset[d];
    k.cs[d] = h.cscale[d]; k.co[d] = h.coffset[d];
    k.n[d] = h.dims[d]; k.cd[d] = h.cdims[d];
    k.one_lo[d] = h.one_lo[d]; k.one_hi[d] = h.one_hi[d];
    if (need_data && h.dims[d] <= 0) return fail(MVF_ERR_INVALID, "%s%s", "mvf_view.dims must be positive");
  }
  k.wmode = h.weight_mode;
  if (k.wmode != MVF_W_OFF && !h.profiles) return fail(MVF_ERR_INVALID, "%s%s", "mvf_view.profiles is NULL but weight_mode != MVF_W_OFF");
  if (k.wmode == MVF_W_BLEND_DCT && !h.cgrid) return fail(MVF_ERR_INVALID, "%s%s", "mvf_view.cgrid is NULL but weight_mode == MVF_W_BLEND_DCT");
  const int T[3] = {TZ - 1, TY - 1, TX - 1};
  for (int d = 0; d < 3; ++d) {
    double ext = 0;
    for (int l = 0; l < 3; ++l) ext += fabs(h.matrix[d * 3 + l]) * T[l];
    k.be[d] = (int)ceil(ext) + 4;
  }
  return MVF_OK;
}

static size_t brick_bytes(const KView& k) {
  const size_t elem = sizeof(float);
  size_t px = (size_t)k.be[2] | 1;
  size_t pz = ((size_t)k.be[1] * px) | 1;
  return ((size_t)k.be[0] * pz + 16) * elem;
}

static dim3 tile_grid(const int32_t dims[3]) {
  return dim3((dims[2] + TX - 1) / TX, (dims[1] + TY - 1) / TY, (dims[0] + TZ - 1) / TZ);
}

template <typename T, int V>
static int launch_fuse(const FuseArgs<V>& A, size_t smem, cudaStream_t st) {
  auto kern = fuse_blend_kernel<T, V>;
  MVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<tile_grid(A.dims), NT, smem, st>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

template <int V>
static int fuse_blend_v(const mvf_view* hv, uint16_t* out, float* outf, const int32_t dims[3], const int32_t org[3], cudaStream_t st) {
  FuseArgs<V> A;
  size_t smem = 0;
  for (int v = 0; v < V; ++v) {
    int rc = to_kview(hv[v], A.v[v], true);
    if (rc) return rc;
    if (hv[v].dtype != hv[0].dtype) return fail(MVF_ERR_INVALID, "%s%s", "mvf_fuse_blend: all views must share one dtype");
    size_t b = brick_bytes(A.v[v]);
    smem = b > smem ? b : smem;
  }
  if (smem > 220 * 1024) return fail(MVF_ERR_UNSUPPORTED, "%s%s", "mvf_fuse_blend: source brick of one tile exceeds shared memory (extreme down-scaling)");
  A.out = out; A.outf = outf;
  for (int d = 0; d < 3; ++d) { A.dims[d] = dims[d]; A.org[d] = org[d]; }
  return hv[0].dtype == MVF_U16 ? launch_fuse<uint16_t, V>(A, smem, st) : launch_fuse<float, V>(A, smem, st);
}

template <int V>
static int blend_weights_v(const mvf_view* hv, float* out, const int32_t dims[3], const int32_t org[3], int normalise, cudaStream_t st) {
  WeightArgs<V> A;
  for (int v = 0; v < V; ++v) {
    int rc = to_kview(hv[v], A.v[v], false);
    if (rc) return rc;
  }
  A.out = out; A.normalise = normalise;
  for (int d = 0; d < 3; ++d) { A.dims[d] = dims[d]; A.org[d] = org[d]; }
  blend_weights_kernel<V><<<tile_grid(dims), NT, 0, st>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

}  // namespace mvf

using namespace mvf;

#define MVF_DISPATCH_V(n, CALL)                \
  switch (n) {                                 \
    case 1: return CALL(1);                    \
    case 2: return CALL(2);                    \
    case 3: return CALL(3);                    \
    case 4: return CALL(4);                    \
    case 5: return CALL(5);                    \
    case 6: return CALL(6);                    \
    case 7: return CALL(7);                    \
    case 8: return CALL(8);                    \
    default: break;                            \
  }

extern "C" int mvf_fuse_blend(int nviews, const mvf_view* host_views, uint16_t* out, float* out_f32,
                              const int32_t out_dims[3], const int32_t out_origin[3], void* stream) {
  MVF_CHECK_ARG(nviews >= 1 && nviews <= MVF_MAX_VIEWS, "nviews must be in [1, MVF_MAX_VIEWS]");
  MVF_CHECK_ARG(host_views && out && out_dims && out_origin, "NULL argument");
  MVF_CHECK_ARG(out_dims[0] > 0 && out_dims[1] > 0 && out_dims[2] > 0, "out_dims must be positive");
#define CALL(V) fuse_blend_v<V>(host_views, out, out_f32, out_dims, out_origin, as_stream(stream))
  MVF_DISPATCH_V(nviews, CALL)
#undef CALL
  return fail(MVF_ERR_INVALID, "%s%s", "mvf_fuse_blend: bad nviews");
}

extern "C" int mvf_blend_weights(int nviews, const mvf_view* host_views, float* weights_out, const int32_t dims[3],
                                 const int32_t origin[3], int normalise, void* stream) {
  MVF_CHECK_ARG(nviews >= 1 && nviews <= MVF_MAX_VIEWS, "nviews must be in [1, MVF_MAX_VIEWS]");
  MVF_CHECK_ARG(host_views && weights_out && dims && origin, "NULL argument");
  MVF_CHECK_ARG(dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "dims must be positive");
#define CALL(V) blend_weights_v<V>(host_views, weights_out, dims, origin, normalise, as_stream(stream))
  MVF_DISPATCH_V(nviews, CALL)
#undef CALL
  return fail(MVF_ERR_INVALID, "%s%s", "mvf_blend_weights: bad nviews");
}

extern "C" int mvf_affine_resample(const void* src, int dtype, const int32_t src_dims[3], const double matrix[9],
                                   const double offset[3], void* dst, const int32_t dst_dims[3],
                                   const int32_t dst_origin[3], int rule, void* stream) {
  MVF_CHECK_ARG(src && dst && src_dims && matrix && offset && dst_dims && dst_origin, "NULL argument");
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  MVF_CHECK_ARG(rule == MVF_RULE_SCIPY || rule == MVF_RULE_ITK, "rule must be MVF_RULE_SCIPY or MVF_RULE_ITK");
  MVF_CHECK_ARG(dst_dims[0] > 0 && dst_dims[1] > 0 && dst_dims[2] > 0, "dst_dims must be positive");
  mvf_view h;
  memset(&h, 0, sizeof(h));
  h.data = src; h.dtype = dtype;
  for (int d = 0; d < 3; ++d) { h.dims[d] = src_dims[d]; h.offset[d] = offset[d]; }
  for (int i = 0; i < 9; ++i) h.matrix[i] = matrix[i];
  h.weight_mode = MVF_W_OFF;
  ResampleArgs A;
  int rc = to_kview(h, A.v, true);
  if (rc) return rc;
  A.dst = dst;
  for (int d = 0; d < 3; ++d) { A.dims[d] = dst_dims[d]; A.org[d] = dst_origin[d]; }
  size_t smem = brick_bytes(A.v);
  if (smem > 220 * 1024) return fail(MVF_ERR_UNSUPPORTED, "%s%s", "mvf_affine_resample: source brick of one tile exceeds shared memory (extreme down-scaling)");
  cudaStream_t st = as_stream(stream);
#define LAUNCH(T, R)                                                                                   \
  do {                                                                                                 \
    auto kern = resample_kernel<T, R>;                                                                 \
    MVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    kern<<<tile_grid(dst_dims), NT, smem, st>>>(A);                                                    \
    MVF_LAUNCHED();                                                                                    \
  } while (0)
  if (dtype == MVF_U16) { if (rule == MVF_RULE_SCIPY) LAUNCH(uint16_t, MVF_RULE_SCIPY); else LAUNCH(uint16_t, MVF_RULE_ITK); }
  else                  { if (rule == MVF_RULE_SCIPY) LAUNCH(float, MVF_RULE_SCIPY);    else LAUNCH(float, MVF_RULE_ITK); }
#undef LAUNCH
  return MVF_OK;
}

extern "C" int mvf_fuse_weighted(int nviews, const void* const* host_view_ptrs, int dtype,
                                 const float* const* host_weight_ptrs, uint16_t* out, float* out_f32,
                                 size_t nvoxels, void* stream) {
  MVF_CHECK_ARG(nviews >= 1 && nviews <= MVF_MAX_VIEWS, "nviews must be in [1, MVF_MAX_VIEWS]");
  MVF_CHECK_ARG(host_view_ptrs && out, "NULL argument");
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  FuseWeightedArgs A;
  memset(&A, 0, sizeof(A));
  for (int v = 0; v < nviews; ++v) {
    MVF_CHECK_ARG(host_view_ptrs[v], "NULL view pointer");
    A.views[v] = host_view_ptrs[v];
    if (host_weight_ptrs) { MVF_CHECK_ARG(host_weight_ptrs[v], "NULL weight pointer"); A.weights[v] = host_weight_ptrs[v]; }
  }
  A.out = out; A.outf = out_f32; A.n = nvoxels; A.nviews = nviews; A.has_weights = host_weight_ptrs != nullptr;
  if (nvoxels == 0) return MVF_OK;
  size_t blocks = (nvoxels + NT - 1) / NT;
  size_t cap = (size_t)sm_count() * 16;
  int grid = (int)(blocks < cap ? blocks : cap);
  if (dtype == MVF_U16) fuse_weighted_kernel<uint16_t><<<grid, NT, 0, as_stream(stream)>>>(A);
  else fuse_weighted_kernel<float><<<grid, NT, 0, as_stream(stream)>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}
