// K1 (affine resampling of one view) and K2 (fused resample + blending/DCT weights + normalise + weighted
// sum) for sm_100a.  Reference arithmetic: mvregfus/multiview.py:1458-1634 (scipy path), 6797-6877 (ITK
// path), 3469-3633 (blending weights), 4291-4317/4434-4439 (DCT weight upsampling), 4904-4912 (fusion).
//
// Work decomposition: one CTA per 8x8x32 (z,y,x) output tile, 256 threads = 8 warps.  A warp owns one y row
// of the tile, its lanes own 32 consecutive x voxels (coalesced 64-byte stores, conflict-free shared-memory
// gathers for axis-aligned views because all shared-memory pitches are odd), and every thread walks the 8 z
// planes of the tile.  For every view the CTA stages the source brick hit by the tile (bounding box of the
// tile's affine footprint) in shared memory with 16-byte coalesced loads, converted to float32 once, then
// gathers the 8 trilinear taps from shared memory, so the access pattern of a rotated view never reaches
// L1/L2.  Blending weights are analytic (three 200-entry 1-D profiles per view) and wait in shared memory
// until the sum over views is known; nothing but the views is read from HBM and nothing but the fused volume
// is written.  The view loop is a real loop (the first version unrolled it and stalled on instruction fetch:
// profiles/r01/fuse_blend_v2_ncu.md).
#include "common.cuh"

namespace mvf {

constexpr int TZ = 8, TY = 8, TX = 32, NT = 256;
constexpr int TILE_VOX = TZ * TY * TX;
static_assert(TY * TX == NT, "one thread per (y,x) column of the tile");
constexpr int ST = 4;  // tiles per super-tile edge (rasterisation order, see tile_grid)

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
  int vec;    // 16-byte aligned rows: vectorised staging allowed
  int be[3];  // host-computed upper bound of the brick extents (shared-memory sizing)
};

struct KArgs {
  KView v[MVF_MAX_VIEWS];
  int nviews;
  int rule;
  int normalise;
  void* out;      // fused uint16 / resampled view / weights (float32, view-major)
  float* outf;    // optional float32 pre-truncation sum
  int dims[3];
  int org[3];
  // block semantics of fuse_block (mv:3817-3916) on the reference's 128^3 chunk grid of the fused frame
  const uint32_t* chunk_info;  // K2: per chunk, bits 0-7 = views taking part, bit 8 = copy mode, bits 12-15 = view to copy
  uint32_t* chunk_flags;       // probe kernels: per chunk, bit v is OR-ed in when view v has data / weight there
  int flag_bit;                // probe kernels: bit to set (resample probe) 
  int chunk;                   // chunk edge (128)
  int nchunk[3];               // chunk grid of the whole fused frame
};

enum { WT_ZERO = 0, WT_ONE = 1, WT_RAMP = 2, WT_EDGE = 3 };

// index of the 128^3 chunk a tile (fused-frame index of its origin) lies in
__device__ __forceinline__ int chunk_of(const KArgs& A, int gz0, int gy0, int gx0) {
  return ((gz0 / A.chunk) * A.nchunk[1] + gy0 / A.chunk) * A.nchunk[2] + gx0 / A.chunk;
}

// Per-tile, per-view geometry, computed once per CTA (thread v handles view v) and kept in shared memory.
struct TileView {
  int b0[3];   // view index of brick element (0,0,0); may be negative (zero filled)
  int e[3];    // extents
  int px, pz;  // shared-memory pitches (elements), both odd
  int skip;      // brick does not intersect the view
  int interior;  // every voxel of the tile maps inside [0, n-1]: no per-voxel boundary test
  int wclass;    // WT_*
  int wdims;     // WT_RAMP: bit d set when axis d leaves the all-ones plateau somewhere in the tile
  unsigned cpr, total, mc, my;  // staging: 16-byte chunks (or elements) per row, chunk count, magic divisors
};

// Bounding box, in the index space of an affine map, of the voxels of one tile (fp64), slightly inflated.
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

// Classify a tile against a view's blending field: every voxel outside the field (weight 0), every voxel
// on the all-ones plateau (weight 1), fully inside the field (fast float32 path) or touching its rim.
__device__ __forceinline__ int classify_tile(const KView& kv, int z0, int y0, int x0, int& wdims) {
  wdims = 7;
  if (kv.wmode == MVF_W_OFF) return WT_ZERO;
  double lo[3], hi[3];
  tile_bbox(kv.wm, kv.wo, z0, y0, x0, lo, hi);
  bool zero = false, ramp = true;
  int active = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    zero = zero || (hi[d] < -0.5) || (lo[d] >= (double)MVF_SIG_N - 0.5);
    const bool plateau = (lo[d] >= (double)kv.one_lo[d]) && (hi[d] < (double)kv.one_hi[d]);
    active |= plateau ? 0 : (1 << d);
    ramp = ramp && (lo[d] >= 0.0) && (hi[d] < (double)(MVF_SIG_N - 1));
  }
  wdims = active;
  return zero ? WT_ZERO : (active == 0 ? WT_ONE : (ramp ? WT_RAMP : WT_EDGE));
}

template <int ELEM>
__device__ __forceinline__ void make_tile_view(const KView& kv, int z0, int y0, int x0, bool need_data, TileView& t) {
  constexpr int VEC = 16 / ELEM;
  t.skip = 1;
  t.interior = 0;
  if (need_data) {
    double lo[3], hi[3];
    tile_bbox(kv.m, kv.o, z0, y0, x0, lo, hi);
    int skip = 0, interior = 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      int i0 = __double2int_rd(lo[d]);
      int i1 = __double2int_rd(hi[d]) + 1;
      if (i1 < 0 || i0 > kv.n[d] - 1) skip = 1;
      if (!(lo[d] >= 0.0 && hi[d] <= (double)(kv.n[d] - 1))) interior = 0;
      int e = i1 - i0 + 1;
      if (d == 2 && kv.vec) {  // align the brick rows to 16 bytes of the source
        int a0 = i0 & ~(VEC - 1);
        e = (i1 - a0 + VEC) & ~(VEC - 1);
        i0 = a0;
      }
      t.b0[d] = i0;
      t.e[d] = e < kv.be[d] ? e : kv.be[d];
    }
    t.px = t.e[2] | 1;
    t.pz = (t.e[1] * t.px) | 1;
    t.skip = skip;
    t.interior = interior;
    t.cpr = kv.vec ? t.e[2] / VEC : t.e[2];
    t.total = t.e[0] * t.e[1] * t.cpr;
    t.mc = t.cpr > 1 ? 0xFFFFFFFFu / t.cpr + 1u : 0u;
    t.my = t.e[1] > 1 ? 0xFFFFFFFFu / (unsigned)t.e[1] + 1u : 0u;
  }
  t.wclass = classify_tile(kv, z0, y0, x0, t.wdims);
}

// blockIdx.x -> tile origin (box-local voxel index); false when the tile lies outside the box
__device__ __forceinline__ bool tile_origin(const int dims[3], int& z0, int& y0, int& x0) {
  const unsigned snx = (dims[2] + TX * ST - 1) / (TX * ST), sny = (dims[1] + TY * ST - 1) / (TY * ST);
  const unsigned b = blockIdx.x;
  const unsigned inner = b & (ST * ST * ST - 1), sup = b / (ST * ST * ST);
  const unsigned sx = sup % snx, sy = (sup / snx) % sny, sz = sup / (snx * sny);
  x0 = (int)((sx * ST + (inner & 3)) * TX);
  y0 = (int)((sy * ST + ((inner >> 2) & 3)) * TY);
  z0 = (int)((sz * ST + (inner >> 4)) * TZ);
  return z0 < dims[0] && y0 < dims[1] && x0 < dims[2];
}

#ifndef MVF_PF
#define MVF_PF 1  // measured on B200 (C2): PF 1/2/4/6 -> 3.84/3.92/4.03/4.22 ms: registers matter more than prefetch depth
#endif
constexpr int PF = MVF_PF;  // 16-byte chunks a thread keeps in flight while the previous view is being gathered

// chunk index -> (iz, iy, c) of the brick, exact for index * divisor < 2^32
__device__ __forceinline__ void chunk_coords(const TileView& t, unsigned i, unsigned& iz, unsigned& iy, unsigned& c) {
  unsigned row = t.cpr > 1 ? __umulhi(i, t.mc) : i;
  c = i - row * t.cpr;
  iz = t.e[1] > 1 ? __umulhi(row, t.my) : row;
  iy = row - iz * (unsigned)t.e[1];
}

template <typename T>
__device__ __forceinline__ uint4 load_chunk(const KView& kv, const TileView& t, unsigned i) {
  constexpr int VEC = 16 / (int)sizeof(T);
  const T* __restrict__ src = reinterpret_cast<const T*>(kv.data);
  unsigned iz, iy, c;
  chunk_coords(t, i, iz, iy, c);
  int gz = t.b0[0] + (int)iz, gy = t.b0[1] + (int)iy, gx = t.b0[2] + (int)c * VEC;
  uint4 q = make_uint4(0u, 0u, 0u, 0u);
  if ((unsigned)gz < (unsigned)kv.n[0] && (unsigned)gy < (unsigned)kv.n[1] && (unsigned)gx < (unsigned)kv.n[2])
    q = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)gz * kv.n[1] + gy) * kv.n[2] + gx));
  return q;
}

template <typename T>
__device__ __forceinline__ void store_chunk(const TileView& t, unsigned i, const uint4& q, float* sm) {
  constexpr int VEC = 16 / (int)sizeof(T);
  unsigned iz, iy, c;
  chunk_coords(t, i, iz, iy, c);
  float* dst = sm + iz * t.pz + iy * t.px + c * VEC;
  if (sizeof(T) == 4) {
    dst[0] = __uint_as_float(q.x); dst[1] = __uint_as_float(q.y);
    dst[2] = __uint_as_float(q.z); dst[3] = __uint_as_float(q.w);
  } else {
    dst[0] = u16_to_float(q.x & 0xFFFFu); dst[1] = u16_to_float(q.x >> 16);
    dst[2] = u16_to_float(q.y & 0xFFFFu); dst[3] = u16_to_float(q.y >> 16);
    dst[4] = u16_to_float(q.z & 0xFFFFu); dst[5] = u16_to_float(q.z >> 16);
    dst[6] = u16_to_float(q.w & 0xFFFFu); dst[7] = u16_to_float(q.w >> 16);
  }
}

// Issue the loads of the first PF*NT chunks of a brick (vector path); the data stays in registers.
template <typename T>
__device__ __forceinline__ void prefetch_brick(const KView& kv, const TileView& t, uint4 (&q)[PF]) {
#pragma unroll
  for (int u = 0; u < PF; ++u) {
    const unsigned i = threadIdx.x + u * NT;
    q[u] = make_uint4(0u, 0u, 0u, 0u);
    if (kv.vec && !t.skip && i < t.total) q[u] = load_chunk<T>(kv, t, i);
  }
}

// Write a prefetched brick to shared memory as float32 (out-of-volume elements are 0) and fetch whatever
// did not fit the prefetch registers (very oblique views) or could not be prefetched (unaligned views).
template <typename T>
__device__ __forceinline__ void commit_brick(const KView& kv, const TileView& t, const uint4 (&q)[PF], float* sm) {
  if (kv.vec) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const unsigned i = threadIdx.x + u * NT;
      if (i < t.total) store_chunk<T>(t, i, q[u], sm);
    }
    // the rest of the brick: batches of 4 chunks per thread, all loads of a batch issued before its first store
    // (the registers are free here: nothing of the gather is live between the two barriers)
    constexpr int U = 4;
    for (unsigned i0 = threadIdx.x + PF * NT; i0 < t.total; i0 += NT * U) {
      uint4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned i = i0 + u * NT;
        r[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < t.total) r[u] = load_chunk<T>(kv, t, i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned i = i0 + u * NT;
        if (i < t.total) store_chunk<T>(t, i, r[u], sm);
      }
    }
  } else {
    const T* __restrict__ src = reinterpret_cast<const T*>(kv.data);
    for (unsigned i = threadIdx.x; i < t.total; i += NT) {
      unsigned iz, iy, ix;
      chunk_coords(t, i, iz, iy, ix);
      int gz = t.b0[0] + (int)iz, gy = t.b0[1] + (int)iy, gx = t.b0[2] + (int)ix;
      float v = 0.f;
      if ((unsigned)gz < (unsigned)kv.n[0] && (unsigned)gy < (unsigned)kv.n[1] && (unsigned)gx < (unsigned)kv.n[2])
        v = to_float<T>(__ldg(src + ((size_t)gz * kv.n[1] + gy) * kv.n[2] + gx));
      sm[iz * t.pz + iy * t.px + ix] = v;
    }
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
__device__ __forceinline__ float gather_scipy(const float* sm, const TileView& t, const double u[3], const double hi[3]) {
  int r[3], neg[3];
  float w[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) split_nearest(u[d], r[d], w[d], neg[d]);
  const int px = t.px, pz = t.pz;
  const float* p = sm + (r[0] - t.b0[0]) * pz + (r[1] - t.b0[1]) * px + (r[2] - t.b0[2]);
  const int ox = 1 | neg[2];
  const int oy = (px ^ neg[1]) - neg[1];
  const int oz = (pz ^ neg[0]) - neg[0];
  float val = lerp8(p, ox, oy, oz, w[2], w[1], w[0]);
  if (!INTERIOR) {
    bool inside = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) inside = inside && (u[d] >= 0.0) && (u[d] <= hi[d]);
    val = inside ? val : 0.f;
  }
  return val;
}

// ITK rule: valid iff -0.5 <= u < n-0.5 per axis, neighbours clamped to the edge, outside -> 0
// (itk::LinearInterpolateImageFunction / ResampleImageFilter).
__device__ __forceinline__ float gather_itk(const float* sm, const TileView& t, const KView& kv, const double rel[3]) {
  int off[3], base = 0;
  float w[3];
  bool inside = true;
  const int stride[3] = {t.pz, t.px, 1};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double u = rel[d] + (double)t.b0[d];
    inside = inside && (u >= -0.5) && (u < (double)kv.n[d] - 0.5);
    int r, neg;
    split_nearest(u, r, w[d], neg);
    int near = max(0, min(r, kv.n[d] - 1));
    int far = max(0, min(r + (1 | neg), kv.n[d] - 1));
    int ln = max(0, min(near - t.b0[d], t.e[d] - 1));
    int lf = max(0, min(far - t.b0[d], t.e[d] - 1));
    base += ln * stride[d];
    off[d] = (lf - ln) * stride[d];
  }
  float val = lerp8(sm + base, off[2], off[1], off[0], w[2], w[1], w[0]);
  return inside ? val : 0.f;
}

// ---- blending field ------------------------------------------------------------------------------
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
__device__ __noinline__ float blend_field_exact(const KView& kv, double c0, double c1, double c2) {
  const double c[3] = {c0, c1, c2};
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
__device__ __noinline__ float coarse_sample(const KView& kv, double c0, double c1, double c2) {
  const double c[3] = {c0, c1, c2};
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

// field value when only axis D leaves the plateau: min(f_D, 1, 1) = f_D, and the trilinear sample collapses
// to a 1-D lerp (the nested ITK lerps of equal values are exact).
__device__ __forceinline__ float field_lerp_1d(const float* f, int i, float t) {
  float a = __ldg(f + i), b = __ldg(f + i + 1);
  return fmaf(b - a, t, a);
}

// floor and fraction of a float64 field coordinate with two DADDs (round-down magic add) and one F2F
__device__ __forceinline__ void ramp_coord(double c, int& i, float& t) {
  const double y = __dadd_rd(c, MAGIC64);
  i = __double2loint(y);
  t = (float)(c - (y - MAGIC64));
}

// Raw (un-normalised) blending weights of the TZ voxels of a thread (fused index (gz+k, gy, gx)) for one view.
// Field coordinates come from the closed form c_d(z) = wm_d0 * z + (wm_d1 * y + wm_d2 * x + wo_d) in float64, so
// a voxel's weight does not depend on the tile/slab decomposition; the tile class only selects how much of the
// (identical) arithmetic can be skipped.
__device__ __forceinline__ void column_blend(const KView& kv, int cls, int wdims, int gz, int gy, int gx, float (&w)[TZ]) {
  if (cls == WT_ZERO || cls == WT_ONE) {
    const float c = cls == WT_ONE ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < TZ; ++k) w[k] = c;
    return;
  }
  double cb[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) cb[d] = fma(kv.wm[d * 3 + 1], (double)gy, fma(kv.wm[d * 3 + 2], (double)gx, kv.wo[d]));
  double zd = (double)gz;
  if (cls == WT_RAMP) {
    if (wdims == 1 || wdims == 2 || wdims == 4) {  // one active axis (CTA-uniform): 1-D lerp of its profile
      const int d = wdims == 1 ? 0 : (wdims == 2 ? 1 : 2);
      const float* f = kv.prof + d * MVF_SIG_N;
      const double step = kv.wm[d * 3 + 0], c0 = cb[d];
#pragma unroll
      for (int k = 0; k < TZ; ++k) {
        int i;
        float t;
        ramp_coord(fma(step, zd, c0), i, t);
        i = max(0, min(i, MVF_SIG_N - 2));  // rounding guard at the bbox limits
        w[k] = field_lerp_1d(f, i, t);
        zd += 1.0;
      }
      return;
    }
#pragma unroll
    for (int k = 0; k < TZ; ++k) {
      int i0[3];
      float t[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ramp_coord(fma(kv.wm[d * 3 + 0], zd, cb[d]), i0[d], t[d]);
        i0[d] = max(0, min(i0[d], MVF_SIG_N - 2));
      }
      w[k] = field_lerp(kv.prof, i0[0], i0[1], i0[2], i0[0] + 1, i0[1] + 1, i0[2] + 1, t[0], t[1], t[2]);
      zd += 1.0;
    }
  } else {
    // tile touches the rim of the field: ITK inside test [-0.5, N-0.5) on the float64 coordinate, neighbours
    // clamped to the edge
#pragma unroll
    for (int k = 0; k < TZ; ++k) {
      int i0[3], i1[3];
      float t[3];
      bool inside = true;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double c64 = fma(kv.wm[d * 3 + 0], zd, cb[d]);
        inside = inside && (c64 >= -0.5) && (c64 < (double)MVF_SIG_N - 0.5);
        int i;
        float tt;
        ramp_coord(c64, i, tt);
        tt = i < 0 ? 0.f : tt;
        i0[d] = max(0, min(i, MVF_SIG_N - 1));
        i1[d] = min(i0[d] + 1, MVF_SIG_N - 1);
        t[d] = tt;
      }
      float wv = field_lerp(kv.prof, i0[0], i0[1], i0[2], i1[0], i1[1], i1[2], t[0], t[1], t[2]);
      w[k] = inside ? wv : 0.f;
      zd += 1.0;
    }
  }
}

// ITK clamp of one continuous index: base index, upper neighbour, lerp weight, inside flag
__device__ __forceinline__ void itk_axis(double c, int n, int& i0, int& i1, float& t, bool& inside) {
  inside = (c >= -0.5) && (c < (double)n - 0.5);
  int i = __double2int_rd(c);
  double fr = c - (double)i;
  if (i < 0) { i = 0; fr = 0.0; }
  if (i >= n - 1) { i = n - 1; fr = 0.0; }
  i0 = i;
  i1 = min(i + 1, n - 1);
  t = (float)fr;
}

// ITK-linear samples of the coarse DCT weight grid (mv:4304-4311) at the TZ voxels of a column.  The coarse
// index changes by 1/(size*bin) per voxel, so the (y,x) bilinear part is evaluated once per coarse z plane the
// column touches (at most 3) and only the final z lerp runs per voxel - same nesting order (x, y, z) as ITK.
__device__ __forceinline__ void column_coarse(const KView& kv, int gz, int gy, int gx, float (&cw)[TZ]) {
  if (kv.wmode != MVF_W_BLEND_DCT) {
#pragma unroll
    for (int k = 0; k < TZ; ++k) cw[k] = 0.f;
    return;
  }
  int y0, y1, x0, x1;
  float ty, tx;
  bool in_y, in_x;
  itk_axis(fma(kv.cs[1], (double)gy, kv.co[1]), kv.cd[1], y0, y1, ty, in_y);
  itk_axis(fma(kv.cs[2], (double)gx, kv.co[2]), kv.cd[2], x0, x1, tx, in_x);
  const double cz0 = fma(kv.cs[0], (double)gz, kv.co[0]);
  int zb = max(0, min(__double2int_rd(cz0), kv.cd[0] - 1));
  const float* g = kv.cgrid;
  const int sy = kv.cd[2], sz = kv.cd[1] * kv.cd[2];
  float plane[4];  // cs[0] <= 1/4 (DCT blocks are >= 4 voxels, mv:3949): a column of TZ=8 voxels spans <= 3 planes + 1
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* q = g + min(zb + j, kv.cd[0] - 1) * sz;
    float a = __ldg(q + y0 * sy + x0), b = __ldg(q + y0 * sy + x1);
    float c = __ldg(q + y1 * sy + x0), d = __ldg(q + y1 * sy + x1);
    float r0 = fmaf(b - a, tx, a), r1 = fmaf(d - c, tx, c);
    plane[j] = fmaf(r1 - r0, ty, r0);
  }
#pragma unroll
  for (int k = 0; k < TZ; ++k) {
    int z0i, z1i;
    float tz;
    bool in_z;
    itk_axis(fma(kv.cs[0], (double)(gz + k), kv.co[0]), kv.cd[0], z0i, z1i, tz, in_z);
    const int a = z0i - zb, b = z1i - zb;
    const float va = a <= 0 ? plane[0] : (a == 1 ? plane[1] : (a == 2 ? plane[2] : plane[3]));
    const float vb = b <= 0 ? plane[0] : (b == 1 ? plane[1] : (b == 2 ? plane[2] : plane[3]));
    cw[k] = (in_z && in_y && in_x) ? fmaf(vb - va, tz, va) : 0.f;
  }
}

// Weights of all views for the thread's column, normalised exactly like the reference: w_v / sum_v w_v in
// float32 with a 0 -> 1 guard (mv:3627-3631), and for the DCT mode coarse_v * that, normalised again
// (mv:4317, 4434-4439).  Raw values go to wsm[v][k][tid]; the divisor of each voxel is returned in wsum.
__device__ __forceinline__ void column_weights(const KArgs& A, const TileView* tv, float* wsm, int gz, int gy, int gx,
                                               float (&wsum)[TZ]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < TZ; ++k) wsum[k] = 0.f;
  bool dct = false;
  for (int v = 0; v < A.nviews; ++v) {
    const KView& kv = A.v[v];
    float w[TZ];
    column_blend(kv, tv[v].wclass, tv[v].wdims, gz, gy, gx, w);
    dct = dct || (kv.wmode == MVF_W_BLEND_DCT);
#pragma unroll
    for (int k = 0; k < TZ; ++k) {
      wsm[(v * TZ + k) * NT + tid] = w[k];
      wsum[k] += w[k];
    }
  }
#pragma unroll
  for (int k = 0; k < TZ; ++k) wsum[k] = wsum[k] == 0.f ? 1.f : wsum[k];
  if (dct) {
    float wsum2[TZ];
#pragma unroll
    for (int k = 0; k < TZ; ++k) wsum2[k] = 0.f;
    for (int v = 0; v < A.nviews; ++v) {
      const KView& kv = A.v[v];
      float cw[TZ];
      column_coarse(kv, gz, gy, gx, cw);
#pragma unroll
      for (int k = 0; k < TZ; ++k) {
        float wn = __fdiv_rn(wsm[(v * TZ + k) * NT + tid], wsum[k]);
        wn = (wn != 0.f && kv.wmode == MVF_W_BLEND_DCT) ? cw[k] * wn : 0.f * wn;
        wsm[(v * TZ + k) * NT + tid] = wn;
        wsum2[k] += wn;
      }
    }
#pragma unroll
    for (int k = 0; k < TZ; ++k) wsum[k] = wsum2[k] == 0.f ? 1.f : wsum2[k];
  }
}

// Values of one view at the thread's TZ voxels (scipy rule), from the staged brick; each value is handed to
// `sink(k, value)` as soon as it is known so that no per-column array of values stays live.
template <bool INTERIOR, typename Sink>
__device__ __forceinline__ void column_gather(const float* sm, const TileView& t, const KView& kv, int gz, int gy, int gx,
                                              Sink sink) {
  // u_d(z) = m_d0 * z + (m_d1 * y + m_d2 * x + o_d): one float64 FMA per axis and voxel on an exact z, so a voxel's
  // coordinate (hence its value) is the same whatever tile, chunk or z-slab it is computed in.
  double base[3], hi[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    base[d] = fma(kv.m[d * 3 + 1], (double)gy, fma(kv.m[d * 3 + 2], (double)gx, kv.o[d]));
    hi[d] = (double)(kv.n[d] - 1);
  }
  double zd = (double)gz;
#pragma unroll
  for (int k = 0; k < TZ; ++k) {
    const double u[3] = {fma(kv.m[0], zd, base[0]), fma(kv.m[3], zd, base[1]), fma(kv.m[6], zd, base[2])};
    sink(k, gather_scipy<INTERIOR>(sm, t, u, hi));
    zd += 1.0;
  }
}

// ------------------------------------------------------------------------------------------------
// K2: fused single pass
// ------------------------------------------------------------------------------------------------
#ifndef MVF_MINB
#define MVF_MINB 3
#endif
template <typename T, bool WANT_F32>
__global__ void __launch_bounds__(NT, MVF_MINB) fuse_blend_kernel(const __grid_constant__ KArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ TileView tv[MVF_MAX_VIEWS];
  float* wsm = reinterpret_cast<float*>(smem_raw);
  float* sm = wsm + A.nviews * TILE_VOX;
  int z0, y0, x0;
  if (!tile_origin(A.dims, z0, y0, x0)) return;
  const int tid = threadIdx.x;
  const uint32_t cinfo = A.chunk_info ? A.chunk_info[chunk_of(A, z0 + A.org[0], y0 + A.org[1], x0 + A.org[2])] : 0xFFu;
  const bool copy_mode = (cinfo >> 8) & 1u;
  const int copy_view = (int)((cinfo >> 12) & 15u);
  if (tid < A.nviews) {
    make_tile_view<sizeof(T)>(A.v[tid], z0 + A.org[0], y0 + A.org[1], x0 + A.org[2], true, tv[tid]);
    if (!copy_mode && !((cinfo >> tid) & 1u)) tv[tid].wclass = WT_ZERO;  // view dropped by fuse_block on this chunk
    if (copy_mode) {  // 0 or 1 live views: the block is the raw transformed view (mv:3828-3836, 3901-3909)
      tv[tid].wclass = tid == copy_view ? WT_ONE : WT_ZERO;
    }
    if (tv[tid].wclass == WT_ZERO) tv[tid].skip = 1;  // zero weight on the whole tile: the view contributes nothing
  }
  __syncthreads();
  const int y = y0 + (tid >> 5), x = x0 + (tid & 31);                        // local (box) index
  const int gz = z0 + A.org[0], gy = y + A.org[1], gx = x + A.org[2];        // fused-frame index of voxel k=0

  float wsum[TZ];
  column_weights(A, tv, wsm, gz, gy, gx, wsum);

  uint32_t acc[TZ];
  float accf[TZ];
#pragma unroll
  for (int k = 0; k < TZ; ++k) { acc[k] = 0u; accf[k] = 0.f; }

  // Software pipeline over the views: the loads of view v+1's brick are in flight (in registers) while view
  // v is gathered from shared memory.
  uint4 q[PF];
  int v = 0;
  while (v < A.nviews && tv[v].skip) ++v;
  if (v < A.nviews) prefetch_brick<T>(A.v[v], tv[v], q);
  while (v < A.nviews) {
    const TileView& t = tv[v];
    const KView& kv = A.v[v];
    __syncthreads();  // everybody is done reading the previous brick
    commit_brick<T>(kv, t, q, sm);
    __syncthreads();
    int vn = v + 1;
    while (vn < A.nviews && tv[vn].skip) ++vn;
    if (vn < A.nviews) prefetch_brick<T>(A.v[vn], tv[vn], q);
    const float* wv = wsm + v * TZ * NT + tid;
    {
      auto sink = [&](int k, float vv) {
        float wn = __fdiv_rn(wv[k * NT], wsum[k]);
        if (sizeof(T) == 2) vv = u16_to_float(round_half_up_u16(vv));
        acc[k] += trunc_product_u16(wn, vv);
        if (WANT_F32) accf[k] = fmaf(wn, vv, accf[k]);
      };
      if (t.interior) column_gather<true>(sm, t, kv, gz, gy, gx, sink);
      else column_gather<false>(sm, t, kv, gz, gy, gx, sink);
    }
    v = vn;
  }

  if (y >= A.dims[1] || x >= A.dims[2]) return;
  uint16_t* out = reinterpret_cast<uint16_t*>(A.out);
#pragma unroll
  for (int k = 0; k < TZ; ++k) {
    const int z = z0 + k;
    if (z >= A.dims[0]) break;
    const size_t idx = ((size_t)z * A.dims[1] + y) * A.dims[2] + x;
    out[idx] = (uint16_t)(sizeof(T) == 2 ? (acc[k] & 0xFFFFu) : min(acc[k], 65535u));
    if (WANT_F32) A.outf[idx] = accf[k];
  }
}

// ------------------------------------------------------------------------------------------------
// K1: one view, same dtype out
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) resample_kernel(const __grid_constant__ KArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ TileView tv[1];
  float* sm = reinterpret_cast<float*>(smem_raw);
  int z0, y0, x0;
  if (!tile_origin(A.dims, z0, y0, x0)) return;
  const int tid = threadIdx.x;
  const KView& kv = A.v[0];
  if (tid == 0) make_tile_view<sizeof(T)>(kv, z0 + A.org[0], y0 + A.org[1], x0 + A.org[2], true, tv[0]);
  __syncthreads();
  const TileView& t = tv[0];
  const int y = y0 + (tid >> 5), x = x0 + (tid & 31);
  const int gz = z0 + A.org[0], gy = y + A.org[1], gx = x + A.org[2];
  float val[TZ];
#pragma unroll
  for (int k = 0; k < TZ; ++k) val[k] = 0.f;
  if (!t.skip) {
    uint4 q[PF];
    prefetch_brick<T>(kv, t, q);
    commit_brick<T>(kv, t, q, sm);
    __syncthreads();
    if (A.rule == MVF_RULE_SCIPY) {
      auto sink = [&](int k, float vv) { val[k] = vv; };
      if (t.interior) column_gather<true>(sm, t, kv, gz, gy, gx, sink);
      else column_gather<false>(sm, t, kv, gz, gy, gx, sink);
    } else {
      double base[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) base[d] = fma(kv.m[d * 3 + 1], (double)gy, fma(kv.m[d * 3 + 2], (double)gx, kv.o[d]));
      double zd = (double)gz;
#pragma unroll 1
      for (int k = 0; k < TZ; ++k) {
        const double rel[3] = {fma(kv.m[0], zd, base[0]) - (double)t.b0[0], fma(kv.m[3], zd, base[1]) - (double)t.b0[1],
                               fma(kv.m[6], zd, base[2]) - (double)t.b0[2]};
        val[k] = gather_itk(sm, t, kv, rel);
        zd += 1.0;
      }
    }
  }
  const bool in_box = y < A.dims[1] && x < A.dims[2];
  if (A.chunk_flags) {  // probe mode: does the transformed view have a non-zero voxel in this chunk? (mv:3828-3830)
    bool any = false;
    if (in_box) {
#pragma unroll
      for (int k = 0; k < TZ; ++k) {
        if (z0 + k >= A.dims[0]) break;
        any = any || (sizeof(T) == 2 ? (A.rule == MVF_RULE_SCIPY ? round_half_up_u16(val[k]) : clamp_trunc_u16(val[k])) > 0u
                                     : val[k] > 0.f);
      }
    }
    if (__syncthreads_or(any) && tid == 0)
      atomicOr(A.chunk_flags + chunk_of(A, z0 + A.org[0], y0 + A.org[1], x0 + A.org[2]), 1u << A.flag_bit);
    return;
  }
  if (!in_box) return;
  T* dst = reinterpret_cast<T*>(A.out);
#pragma unroll
  for (int k = 0; k < TZ; ++k) {
    const int z = z0 + k;
    if (z >= A.dims[0]) break;
    const size_t idx = ((size_t)z * A.dims[1] + y) * A.dims[2] + x;
    if (sizeof(T) == 2) dst[idx] = (T)(A.rule == MVF_RULE_SCIPY ? round_half_up_u16(val[k]) : clamp_trunc_u16(val[k]));
    else dst[idx] = (T)val[k];
  }
}

// ------------------------------------------------------------------------------------------------
// blending weights on a grid (get_weights_simple), V float32 volumes out
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) blend_weights_kernel(const __grid_constant__ KArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ TileView tv[MVF_MAX_VIEWS];
  float* wsm = reinterpret_cast<float*>(smem_raw);
  int z0, y0, x0;
  if (!tile_origin(A.dims, z0, y0, x0)) return;
  const int tid = threadIdx.x;
  if (tid < A.nviews)
    make_tile_view<4>(A.v[tid], z0 + A.org[0], y0 + A.org[1], x0 + A.org[2], false, tv[tid]);
  __syncthreads();
  const int y = y0 + (tid >> 5), x = x0 + (tid & 31);
  const int gz = z0 + A.org[0], gy = y + A.org[1], gx = x + A.org[2];
  float wsum[TZ];
  column_weights(A, tv, wsm, gz, gy, gx, wsum);
  const bool in_box = y < A.dims[1] && x < A.dims[2];
  if (A.chunk_flags) {  // probe mode: which views have a non-zero blending weight somewhere in this chunk? (mv:3902-3904)
    for (int v = 0; v < A.nviews; ++v) {
      bool any = false;
      if (in_box) {
#pragma unroll
        for (int k = 0; k < TZ; ++k) {
          if (z0 + k >= A.dims[0]) break;
          any = any || wsm[(v * TZ + k) * NT + tid] > 0.f;
        }
      }
      if (__syncthreads_or(any) && tid == 0)
        atomicOr(A.chunk_flags + chunk_of(A, z0 + A.org[0], y0 + A.org[1], x0 + A.org[2]), 1u << v);
    }
    return;
  }
  if (!in_box) return;
  float* out = reinterpret_cast<float*>(A.out);
  const size_t vol = (size_t)A.dims[0] * A.dims[1] * A.dims[2];
  for (int v = 0; v < A.nviews; ++v) {
#pragma unroll
    for (int k = 0; k < TZ; ++k) {
      const int z = z0 + k;
      if (z >= A.dims[0]) break;
      float w = wsm[(v * TZ + k) * NT + tid];
      if (A.normalise) w = __fdiv_rn(w, wsum[k]);
      out[v * vol + ((size_t)z * A.dims[1] + y) * A.dims[2] + x] = w;
    }
  }
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
  const int vecn = h.dtype == MVF_U16 ? 8 : 4;
  k.vec = need_data && ((reinterpret_cast<uintptr_t>(h.data) & 15u) == 0) && (h.dims[2] % vecn == 0);
  const int T[3] = {TZ - 1, TY - 1, TX - 1};
  for (int d = 0; d < 3; ++d) {
    double ext = 0;
    for (int l = 0; l < 3; ++l) ext += fabs(h.matrix[d * 3 + l]) * T[l];
    k.be[d] = (int)ceil(ext) + 4 + (d == 2 && k.vec ? 2 * vecn : 0);
  }
  return MVF_OK;
}

static size_t brick_bytes(const KView& k) {
  size_t px = (size_t)k.be[2] | 1;
  size_t pz = ((size_t)k.be[1] * px) | 1;
  return ((size_t)k.be[0] * pz + 16) * sizeof(float);
}

// Tiles are rasterised in 4x4x4 super tiles (32x32x128 voxels) so that tiles sharing brick halos run close
// together in time and the halo re-reads hit L2 instead of HBM.
static dim3 tile_grid(const int32_t dims[3]) {
  unsigned nx = (dims[2] + TX * ST - 1) / (TX * ST), ny = (dims[1] + TY * ST - 1) / (TY * ST), nz = (dims[0] + TZ * ST - 1) / (TZ * ST);
  return dim3(nx * ny * nz * ST * ST * ST);
}

static int fill_args(KArgs& A, int nviews, const mvf_view* hv, bool need_data, void* out, float* outf,
                     const int32_t dims[3], const int32_t org[3], size_t* brick) {
  memset(&A, 0, sizeof(A));
  size_t smem = 0;
  for (int v = 0; v < nviews; ++v) {
    int rc = to_kview(hv[v], A.v[v], need_data);
    if (rc) return rc;
    if (need_data && hv[v].dtype != hv[0].dtype) return fail(MVF_ERR_INVALID, "%s%s", "all views of one call must share one dtype");
    if (need_data) {
      size_t b = brick_bytes(A.v[v]);
      smem = b > smem ? b : smem;
    }
  }
  A.nviews = nviews;
  A.out = out;
  A.outf = outf;
  for (int d = 0; d < 3; ++d) { A.dims[d] = dims[d]; A.org[d] = org[d]; }
  *brick = smem;
  return MVF_OK;
}

template <typename K>
static int launch_tiles(K kern, const KArgs& A, size_t smem, cudaStream_t st) {
  if (smem > 220 * 1024)
    return fail(MVF_ERR_UNSUPPORTED, "%s%s", "source brick of one output tile exceeds shared memory (extreme down-scaling)");
  MVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<tile_grid(A.dims), NT, smem, st>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

}  // namespace mvf

#include <stdlib.h>
#include "fuse_axis.cuh"

using namespace mvf;

extern "C" int mvf_fuse_blend(int nviews, const mvf_view* host_views, uint16_t* out, float* out_f32,
                              const int32_t out_dims[3], const int32_t out_origin[3], void* stream) {
  MVF_CHECK_ARG(nviews >= 1 && nviews <= MVF_MAX_VIEWS, "nviews must be in [1, MVF_MAX_VIEWS]");
  MVF_CHECK_ARG(host_views && out && out_dims && out_origin, "NULL argument");
  MVF_CHECK_ARG(out_dims[0] > 0 && out_dims[1] > 0 && out_dims[2] > 0, "out_dims must be positive");
  KArgs A;
  size_t brick;
  int rc = fill_args(A, nviews, host_views, true, out, out_f32, out_dims, out_origin, &brick);
  if (rc) return rc;
  const size_t smem = brick + (size_t)nviews * TILE_VOX * sizeof(float);
  cudaStream_t st = as_stream(stream);
  // axis-aligned views (signed permutation x diagonal index maps): TMA-staged, table-driven kernel
  int taken = 0;
  rc = try_fuse_axis(nviews, host_views, A, st, &taken);
  if (rc || taken) return rc;
  if (host_views[0].dtype == MVF_U16)
    return out_f32 ? launch_tiles(fuse_blend_kernel<uint16_t, true>, A, smem, st) : launch_tiles(fuse_blend_kernel<uint16_t, false>, A, smem, st);
  return out_f32 ? launch_tiles(fuse_blend_kernel<float, true>, A, smem, st) : launch_tiles(fuse_blend_kernel<float, false>, A, smem, st);
}

static int set_chunks(KArgs& A, const uint32_t* info, uint32_t* flags, int chunk, const int32_t nchunk[3], const int32_t org[3]) {
  if (!info && !flags) return MVF_OK;
  if (chunk <= 0 || chunk % TX != 0 || !nchunk) return fail(MVF_ERR_INVALID, "%s%s", "chunk edge must be a positive multiple of 32");
  if (org[0] % TZ || org[1] % TY || org[2] % TX)
    return fail(MVF_ERR_INVALID, "%s%s", "with chunk tables the box origin must be a multiple of the 8x8x32 tile");
  A.chunk_info = info; A.chunk_flags = flags; A.chunk = chunk;
  for (int d = 0; d < 3; ++d) A.nchunk[d] = nchunk[d];
  return MVF_OK;
}

extern "C" int mvf_fuse_blend_chunked(int nviews, const mvf_view* host_views, uint16_t* out, const int32_t out_dims[3],
                                      const int32_t out_origin[3], const uint32_t* chunk_info, int chunk,
                                      const int32_t nchunks[3], void* stream) {
  MVF_CHECK_ARG(nviews >= 1 && nviews <= MVF_MAX_VIEWS, "nviews must be in [1, MVF_MAX_VIEWS]");
  MVF_CHECK_ARG(host_views && out && out_dims && out_origin && chunk_info && nchunks, "NULL argument");
  MVF_CHECK_ARG(out_dims[0] > 0 && out_dims[1] > 0 && out_dims[2] > 0, "out_dims must be positive");
  KArgs A;
  size_t brick;
  int rc = fill_args(A, nviews, host_views, true, out, nullptr, out_dims, out_origin, &brick);
  if (rc) return rc;
  rc = set_chunks(A, chunk_info, nullptr, chunk, nchunks, out_origin);
  if (rc) return rc;
  const size_t smem = brick + (size_t)nviews * TILE_VOX * sizeof(float);
  cudaStream_t st = as_stream(stream);
  return host_views[0].dtype == MVF_U16 ? launch_tiles(fuse_blend_kernel<uint16_t, false>, A, smem, st)
                                        : launch_tiles(fuse_blend_kernel<float, false>, A, smem, st);
}

// chunk_flags[chunk] |= 1 << v for every view v whose transformed (scipy-rule) voxels are not all zero on the chunk
// (probe == 0) or whose raw blending weight is positive somewhere on the chunk (probe == 1).
extern "C" int mvf_chunk_probe(int nviews, const mvf_view* host_views, int probe, const int32_t dims[3],
                               const int32_t origin[3], uint32_t* chunk_flags, int chunk, const int32_t nchunks[3],
                               void* stream) {
  MVF_CHECK_ARG(nviews >= 1 && nviews <= MVF_MAX_VIEWS, "nviews must be in [1, MVF_MAX_VIEWS]");
  MVF_CHECK_ARG(host_views && dims && origin && chunk_flags && nchunks, "NULL argument");
  MVF_CHECK_ARG(probe == 0 || probe == 1, "probe must be 0 (data) or 1 (weights)");
  cudaStream_t st = as_stream(stream);
  if (probe == 1) {
    KArgs A;
    size_t brick;
    int rc = fill_args(A, nviews, host_views, false, nullptr, nullptr, dims, origin, &brick);
    if (rc) return rc;
    rc = set_chunks(A, nullptr, chunk_flags, chunk, nchunks, origin);
    if (rc) return rc;
    return launch_tiles(blend_weights_kernel, A, (size_t)nviews * TILE_VOX * sizeof(float), st);
  }
  for (int v = 0; v < nviews; ++v) {
    KArgs A;
    size_t brick;
    int rc = fill_args(A, 1, host_views + v, true, nullptr, nullptr, dims, origin, &brick);
    if (rc) return rc;
    rc = set_chunks(A, nullptr, chunk_flags, chunk, nchunks, origin);
    if (rc) return rc;
    A.rule = MVF_RULE_SCIPY;
    A.flag_bit = v;
    rc = host_views[v].dtype == MVF_U16 ? launch_tiles(resample_kernel<uint16_t>, A, brick, st)
                                        : launch_tiles(resample_kernel<float>, A, brick, st);
    if (rc) return rc;
  }
  return MVF_OK;
}

extern "C" int mvf_blend_weights(int nviews, const mvf_view* host_views, float* weights_out, const int32_t dims[3],
                                 const int32_t origin[3], int normalise, void* stream) {
  MVF_CHECK_ARG(nviews >= 1 && nviews <= MVF_MAX_VIEWS, "nviews must be in [1, MVF_MAX_VIEWS]");
  MVF_CHECK_ARG(host_views && weights_out && dims && origin, "NULL argument");
  MVF_CHECK_ARG(dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "dims must be positive");
  KArgs A;
  size_t brick;
  int rc = fill_args(A, nviews, host_views, false, weights_out, nullptr, dims, origin, &brick);
  if (rc) return rc;
  A.normalise = normalise;
  return launch_tiles(blend_weights_kernel, A, (size_t)nviews * TILE_VOX * sizeof(float), as_stream(stream));
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
  KArgs A;
  size_t brick;
  int rc = fill_args(A, 1, &h, true, dst, nullptr, dst_dims, dst_origin, &brick);
  if (rc) return rc;
  A.rule = rule;
  return dtype == MVF_U16 ? launch_tiles(resample_kernel<uint16_t>, A, brick, as_stream(stream))
                          : launch_tiles(resample_kernel<float>, A, brick, as_stream(stream));
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
