// K2, bandwidth-shaped variant for views whose index map  u_view = matrix' * i_fused + offset'  is a signed axis
// permutation times a diagonal (the reference's default acquisition geometry: stage rotations of 0/90/180/270 deg,
// mvregfus/mv_graph.py:66, with per-axis scale): every view coordinate then depends on ONE fused index, so
// everything the per-voxel arithmetic of resample_fuse.cu derives from float64 coordinates (floor index, lerp
// weights, scipy's inside test, the ITK sample of the blending field, mv:1486-1496, 1606-1610, 3558-3611) becomes
// three small per-axis tables per view and tile.  Same results as the general kernel to float32 rounding.
//
//   * CTA = 16x8x32 (z,y,x) output tile, 256 threads: warp <-> y row, lane <-> x, a thread walks 16 z voxels.
//   * The source brick of every view is a dense box fetched by TMA (cp.async.bulk.tensor.3d, one CUtensorMap per
//     view, out-of-volume elements zero filled by the hardware) into one of two shared-memory buffers; the brick
//     of view v+2 is requested as soon as view v has been gathered, completion is tracked with mbarriers.  No LSU
//     instruction, address arithmetic or bounds test is spent on staging.
//   * Per z step a thread needs the bilinear value of two neighbouring planes of the brick along the view axis its
//     z walk follows; the plane index is uniform over the tile, so the planes are carried from step to step
//     (scale ~1: one new plane = 4 LDS + 6 FP per voxel-view; 2x up-sampling: every other step loads nothing).
//   * Lerps are  (1-t)*a + t*b  with both weights taken from the float64 fraction: no cancellation for the
//     non-negative data of the path, exact on constant neighbourhoods, zero when the voxel is outside the view
//     (both weights are 0 there, which is scipy's mode='constant', cval=0).
//   * Blending weights: per view and tile one of {0, 1, 1-D ramp along z / y / x, general}; tiles on which every
//     view is 0 or 1 (the interior) skip the per-voxel normalisation altogether.
#pragma once
#include <cuda.h>

namespace mvf {

constexpr int AZ = 16, AY = 8, AX = 32, ANT = 256, ATAB = AZ + AY + AX;
static_assert(AY * AX == ANT, "one thread per (y,x) column of the tile");
constexpr int AST = 4;   // tiles per super-tile edge (rasterisation order: neighbours in space run close in time)

struct TmaMaps { CUtensorMap m[MVF_MAX_VIEWS]; };
struct __align__(16) GTab { int off; float t, omt; int mode; };   // gather table entry of one fused index
struct __align__(16) WTab { float v0, v1, t, w1; };               // blending-field entry: profile pair, fraction (-1 = outside), 1-D weight

// Everything the kernel needs per view, per FUSED axis a (0 = z, 1 = y, 2 = x); kept small so that the whole
// parameter block (tensor maps included) stays below 4 KB.
struct AxView {
  double sc[3], of[3];     // view coordinate along view axis vax[a]:  u = sc[a] * g_a + of[a]   (matrix'/offset', mv:1494-1496)
  double wsc[3], wof[3];   // blending-field coordinate along field axis vax[a]                   (mv:3567, 3604)
  const float* prof;       // 3 x MVF_SIG_N border profiles, field axis order
  int vax[3];    // vax[a] = view axis driven by fused axis a
  int n[3];      // n[a] = view extent along view axis vax[a]
  int box[3];    // brick extents in view axis order (z, y, x): the TMA box
  int str[3];    // str[a] = element stride, inside the brick, of view axis vax[a]
  int one_lo[3], one_hi[3];   // plateau of profile vax[a]
  int bytes;     // box bytes (the mbarrier's expected transaction count)
  int wmode;
  int xalign;    // elements per 16 bytes
  int dsz;       // signed stride of one step of the z walk: str[0] * sign(sc[0])
};
// weight class of a view on a tile: which fused axes leave the all-ones plateau of the blending field there
enum { AW_ZERO = 0, AW_ONE = 1, AW_Z = 2, AW_Y = 3, AW_X = 4, AW_3D = 5, AW_ZT = 6 /* z and one of y / x */, AW_YX = 7 };

struct __align__(16) AxTile {
  int b0[3];   // view index of brick element (0,0,0); b0[a] belongs to view axis vax[a]
  int skip;    // the tile sees no voxel of the view (or its weight is 0 everywhere): nothing to gather
  int wcls;    // AW_*
  int wactive; // bit a set: fused axis a leaves the plateau somewhere in the tile
  int corr;    // element offset of brick element (0,0,0) in the absolute offsets of the tables
  int regular; // the z walk never jumps over a plane inside this tile
  int unit;    // ... and moves on by exactly one plane at every step
  int pad[3];
};
static_assert(sizeof(AxTile) == 48, "tile records move as three 16-byte vectors");

struct AxArgs {
  AxView v[MVF_MAX_VIEWS];
  void* out;
  float* outf;
  int dims[3], org[3];
  int nviews;
  int brick_stride;   // bytes between the two brick buffers (multiple of 128)
  AxTile* tiles;      // per (tile, view) geometry and weight class (ax_tiles_kernel), tile-major
  GTab* gtab;         // per-axis tables of the box, view-major then z | y | x (ax_tables_kernel)
  WTab* wtab;
};


// ---- TMA / mbarrier primitives (sm_90+ PTX) ---------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MVF_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MVF_DONE_%=;\n\t"
      "bra MVF_WAIT_%=;\n\t"
      "MVF_DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int cx, int cy, int cz) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
      "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(cx), "r"(cy), "r"(cz) : "memory");
}

template <typename T> __device__ __forceinline__ float tap(const T* p, int off);
template <> __device__ __forceinline__ float tap<float>(const float* p, int off) { return p[off]; }
template <> __device__ __forceinline__ float tap<uint16_t>(const uint16_t* p, int off) { return u16_to_float(p[off]); }

// 1 / s to within an ulp for s in [~1e-3, 8] (sums of blending weights): hardware approximation + one Newton step,
// without the special-case path of the IEEE reciprocal
__device__ __forceinline__ float rcp_newton(float s) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
  return fmaf(y, fmaf(-s, y, 1.f), y);
}

// RN(w / s) from rs = RN(1 / s): one product, one exact residual, one correction (Markstein)
__device__ __forceinline__ float div_by(float w, float s, float rs) {
  const float q = w * rs;
  return fmaf(fmaf(-q, s, w), rs, q);
}

__device__ __forceinline__ bool ax_tile_origin(const int dims[3], unsigned b, int& z0, int& y0, int& x0) {
  const unsigned snx = (dims[2] + AX * AST - 1) / (AX * AST), sny = (dims[1] + AY * AST - 1) / (AY * AST);
  const unsigned inner = b & (AST * AST * AST - 1), sup = b / (AST * AST * AST);
  const unsigned sx = sup % snx, sy = (sup / snx) % sny, sz = sup / (snx * sny);
  x0 = (int)((sx * AST + (inner & 3)) * AX);
  y0 = (int)((sy * AST + ((inner >> 2) & 3)) * AY);
  z0 = (int)((sz * AST + (inner >> 4)) * AZ);
  return z0 < dims[0] && y0 < dims[1] && x0 < dims[2];
}

// ---- per-axis tables of the whole box, filled once per launch -----------------------------------------------
// For every view, fused axis a and index i of the box one 16-byte gather entry and one 16-byte blending-field entry:
// all float64 coordinate work of a launch happens here ((nz + ny + nx) * V threads); the tiles of the main kernel
// only copy their slices.
//   gather entry, y/x axes: j = floor(u), t / omt = weights of elements j+1 / j (both 0 when u is outside [0, n-1]:
//     scipy mode='constant', cval=0).  z axis (the walk): kept in walk direction - j = LEADING plane (the one a step
//     newly needs: j+1 when the view index grows with z, j when it falls), t = its weight, omt = weight of the trailing
//     plane; flags bits 0-1: 0 = same plane pair as index i-1, 1 = moved on by one plane, 3 = anything else.
//   flags bit 2 / 3: u below 0 / above n-1;  bits 8-15: blending-field floor index;  bit 16 / 17: field coordinate
//     below -0.5 / at or above N-0.5 (ITK: outside).
__global__ void __launch_bounds__(256) ax_tables_kernel(const __grid_constant__ AxArgs X) {
  const int ntot = X.dims[0] + X.dims[1] + X.dims[2];
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= X.nviews * ntot) return;
  const int v = e / ntot, r = e - v * ntot;
  const int a = r < X.dims[0] ? 0 : (r < X.dims[0] + X.dims[1] ? 1 : 2);
  const int i = r - (a == 0 ? 0 : (a == 1 ? X.dims[0] : X.dims[0] + X.dims[1]));
  const AxView& av = X.v[v];
  const int g = X.org[a] + i;
  auto floor_of = [&](int gg, double& u) {
    u = fma(av.sc[a], (double)gg, av.of[a]);
    return __double2int_rd(fmin(fmax(u, -1.0e9), 1.0e9));
  };
  double u;
  const int j = floor_of(g, u);
  const double fr = u - (double)j;
  const bool below = u < 0.0, above = u > (double)(av.n[a] - 1);
  const float t = (below || above) ? 0.f : (float)fr, omt = (below || above) ? 0.f : (float)(1.0 - fr);
  GTab ge;
  int flags = (below ? 4 : 0) | (above ? 8 : 0);
  if (a != 0) {
    ge.off = j; ge.t = t; ge.omt = omt;
  } else {
    const bool fwd = av.sc[0] > 0.0;
    ge.off = fwd ? j + 1 : j;
    ge.t = fwd ? t : omt;
    ge.omt = fwd ? omt : t;
    if (i > 0) {
      double up;
      const int dl = (j - floor_of(g - 1, up)) * (fwd ? 1 : -1);
      flags |= dl == 0 ? 0 : (dl == 1 ? 1 : 3);
    }
  }
  ge.off *= av.str[a];   // element offset inside the view's brick layout (the tile subtracts its brick start)
  WTab we = {1.f, 1.f, 0.f, 1.f};
  if (av.wmode != MVF_W_OFF) {
    const int d = av.vax[a];
    const double c = fma(av.wsc[a], (double)g, av.wof[a]);
    const bool fbelow = c < -0.5, fabove = !(c < (double)MVF_SIG_N - 0.5);   // ITK: valid on [-0.5, N-0.5), neighbours clamped
    int k = __double2int_rd(fmin(fmax(c, -1.0e6), 1.0e6));
    double cf = c - (double)k;
    if (k < 0) { k = 0; cf = 0.0; }
    if (k >= MVF_SIG_N - 1) { k = MVF_SIG_N - 1; cf = 0.0; }
    const int k1 = min(k + 1, MVF_SIG_N - 1);
    const float* f = av.prof + d * MVF_SIG_N;
    we.v0 = __ldg(f + k);
    we.v1 = __ldg(f + k1);
    const float tt = (float)cf;
    const bool inside = !fbelow && !fabove;
    we.t = inside ? tt : -1.f;
    we.w1 = inside ? fmaf(we.v1 - we.v0, tt, we.v0) : 0.f;
    flags |= (k << 8) | (fbelow ? 1 << 16 : 0) | (fabove ? 1 << 17 : 0);
  }
  ge.mode = flags;
  X.gtab[e] = ge;
  X.wtab[e] = we;
}

// Per view and tile, from the table entries at the two ends of the tile along every axis (index maps are monotone):
// where the brick starts, whether the tile sees the view at all, the class of its blending weights.
__device__ __forceinline__ void ax_make_tile(const AxArgs& X, int v, const int loc0[3], AxTile& t) {
  const AxView& av = X.v[v];
  const int EXT[3] = {AZ, AY, AX};
  const int ntot = X.dims[0] + X.dims[1] + X.dims[2];
  bool skip = false, zero = av.wmode == MVF_W_OFF;
  int active = 0, corr = 0, pre = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const GTab* g = X.gtab + (size_t)v * ntot + pre;
    const GTab ea = g[loc0[a]], eb = g[min(loc0[a] + EXT[a] - 1, X.dims[a] - 1)];
    pre += X.dims[a];
    const int lead = (a == 0 && av.dsz > 0) ? av.str[0] : 0;   // z entries hold the leading plane
    int b = min(ea.off - lead, eb.off - lead);                  // element offset of the lowest floor index of the tile
    if (av.vax[a] == 2) b &= ~(av.xalign - 1);                  // TMA: the box starts on a 16-byte boundary of the view's rows
    t.b0[a] = b / av.str[a];                                    // exact: offsets are multiples of the stride
    corr += b;
    skip = skip || ((ea.mode & eb.mode & 4) != 0) || ((ea.mode & eb.mode & 8) != 0);
    if (av.wmode != MVF_W_OFF) {
      zero = zero || ((ea.mode & eb.mode & (1 << 16)) != 0) || ((ea.mode & eb.mode & (1 << 17)) != 0);
      const int ia = (ea.mode >> 8) & 255, ib = (eb.mode >> 8) & 255;
      const bool inside = ((ea.mode | eb.mode) & (3 << 16)) == 0;
      const bool plateau = inside && min(ia, ib) >= av.one_lo[a] && max(ia, ib) < av.one_hi[a];
      active |= plateau ? 0 : (1 << a);
    }
  }
  t.corr = corr;
  // bit 0 = z, 1 = y, 2 = x active
  t.wcls = zero ? AW_ZERO : (active == 0 ? AW_ONE : (active == 1 ? AW_Z : (active == 2 ? AW_Y : (active == 4 ? AW_X :
           (active == 6 ? AW_YX : (active == 7 ? AW_3D : AW_ZT))))));
  t.wactive = active;
  t.skip = skip || zero;
}

// One thread per (tile, view): everything a tile needs to know about a view before it can request its brick, so that
// the fused kernel's first action per view is one 48-byte load instead of table look-ups and integer logic.
__global__ void __launch_bounds__(128) ax_tiles_kernel(const __grid_constant__ AxArgs X, unsigned ntiles) {
  const unsigned e = blockIdx.x * 128 + threadIdx.x;
  if (e >= ntiles * (unsigned)X.nviews) return;
  const unsigned tile = e / X.nviews;
  const int v = (int)(e - tile * X.nviews);
  int z0, y0, x0;
  AxTile t;
  if (!ax_tile_origin(X.dims, tile, z0, y0, x0)) return;
  const int loc0[3] = {z0, y0, x0};
  ax_make_tile(X, v, loc0, t);
  const int ntot = X.dims[0] + X.dims[1] + X.dims[2];
  bool regular = true, unit = true;   // z walk of this view inside the tile (step k = 1 .. AZ-1)
  for (int k = 1; k < AZ; ++k) {
    const int code = z0 + k < X.dims[0] ? (X.gtab[(size_t)v * ntot + z0 + k].mode & 3) : 0;
    regular = regular && code != 3;
    unit = unit && code == 1;
  }
  t.regular = regular; t.unit = unit;
  t.pad[0] = t.pad[1] = t.pad[2] = 0;
  X.tiles[e] = t;
}

// ITK-linear sample of min(f_z, f_y, f_x) from the three per-axis entries (lerp x, then y, then z)
__device__ __forceinline__ float ax_weight_3d(const WTab& ez, const WTab& ey, const WTab& ex) {
  const float m00 = fminf(ez.v0, ey.v0), m01 = fminf(ez.v0, ey.v1), m10 = fminf(ez.v1, ey.v0), m11 = fminf(ez.v1, ey.v1);
  const float v000 = fminf(m00, ex.v0), v001 = fminf(m00, ex.v1), v010 = fminf(m01, ex.v0), v011 = fminf(m01, ex.v1);
  const float v100 = fminf(m10, ex.v0), v101 = fminf(m10, ex.v1), v110 = fminf(m11, ex.v0), v111 = fminf(m11, ex.v1);
  const float x00 = fmaf(v001 - v000, ex.t, v000);
  const float x01 = fmaf(v011 - v010, ex.t, v010);
  const float x10 = fmaf(v101 - v100, ex.t, v100);
  const float x11 = fmaf(v111 - v110, ex.t, v110);
  const float y0 = fmaf(x01 - x00, ey.t, x00);
  const float y1 = fmaf(x11 - x10, ey.t, x10);
  const float w = fmaf(y1 - y0, ez.t, y0);
  return (ez.t >= 0.f && ey.t >= 0.f && ex.t >= 0.f) ? w : 0.f;
}


// ---- one view's values at the thread's AZ voxels ----------------------------------------------------------------
// Plane walk: per step the bilinear value of the leading plane is taken from the brick (4 LDS + 6 FP) and the pair
// (trailing, leading) is carried along.  REGULAR walks (never more than one plane per step) are straight-line code: the
// new plane is always loaded and the pair moves by predicated selects, so the loads of step k+1 overlap the arithmetic
// of step k.
template <bool REGULAR, bool UNIT, typename Plane>
__device__ __forceinline__ void ax_walk_planes(const GTab* __restrict__ gz_tab, Plane plane, int dsz, float (&val)[AZ]) {
  float Pl = 0.f, Pd = plane(gz_tab[0].off - dsz);   // trailing plane of the first step
#pragma unroll
  for (int k = 0; k < AZ; ++k) {
    const GTab gz = gz_tab[k];   // uniform over the CTA: one broadcast LDS.128
    if (UNIT) {            // every step moves on by exactly one plane: no selects
      Pl = Pd;
      Pd = plane(gz.off);
    } else if (REGULAR) {
      const float np = plane(gz.off);
      const bool adv = gz.mode != 0;
      Pl = adv ? Pd : Pl;
      Pd = adv ? np : Pd;
    } else {
      if (gz.mode == 3) { Pl = plane(gz.off - dsz); Pd = plane(gz.off); }
      else if (gz.mode == 1) { Pl = Pd; Pd = plane(gz.off); }
    }
    val[k] = fmaf(gz.t, Pd, gz.omt * Pl);
  }
}

// Window walk, for float32 views whose z walk runs along the view's contiguous x axis and moves by exactly one element
// per step (stage rotations of 90 / 270 degrees at ~unit scale).  The lanes of a warp then sit on different view z
// planes: a 32-bit shared-memory load from a dense TMA brick is at best 4-way bank conflicted (all pitches are
// multiples of 16 bytes).  16-byte loads are conflict free when pitch/16 is odd (axis_view guarantees it), and the 17
// elements a thread needs along x are contiguous: 5 LDS.128 per corner row, bilinear reduction on the whole window,
// then the step values are lerps of neighbouring window elements.  phi = position of the first element inside its
// 16-byte chunk (uniform over the tile).
template <bool FWD>
__device__ __forceinline__ void ax_walk_window(const GTab* __restrict__ gz_tab, const float* __restrict__ p00, const float* __restrict__ p01,
                                               const float* __restrict__ p10, const float* __restrict__ p11, float tX, float oX,
                                               float tY, float oY, float (&val)[AZ]) {
  const int xmin = FWD ? gz_tab[0].off - 1 : gz_tab[AZ - 1].off;   // lowest element of the walk (trailing plane of step 0 / leading plane of the last step)
  const int c0 = xmin & ~3, phi = xmin & 3;
  float Q[20];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(p00 + c0 + 4 * c), bb = *reinterpret_cast<const float4*>(p01 + c0 + 4 * c);
    const float4 cc = *reinterpret_cast<const float4*>(p10 + c0 + 4 * c), d = *reinterpret_cast<const float4*>(p11 + c0 + 4 * c);
    Q[4 * c + 0] = fmaf(tY, fmaf(tX, d.x, oX * cc.x), oY * fmaf(tX, bb.x, oX * a.x));
    Q[4 * c + 1] = fmaf(tY, fmaf(tX, d.y, oX * cc.y), oY * fmaf(tX, bb.y, oX * a.y));
    Q[4 * c + 2] = fmaf(tY, fmaf(tX, d.z, oX * cc.z), oY * fmaf(tX, bb.z, oX * a.z));
    Q[4 * c + 3] = fmaf(tY, fmaf(tX, d.w, oX * cc.w), oY * fmaf(tX, bb.w, oX * a.w));
  }
  // B[i] = Q[i + phi], i = 0..16, by two rounds of uniform selects
  const bool s1 = (phi & 1) != 0, s2 = (phi & 2) != 0;
  float A[19];
#pragma unroll
  for (int i = 0; i < 19; ++i) A[i] = s1 ? Q[i + 1] : Q[i];
  float B[17];
#pragma unroll
  for (int i = 0; i < 17; ++i) B[i] = s2 ? A[i + 2] : A[i];
#pragma unroll
  for (int k = 0; k < AZ; ++k) {
    const GTab gz = gz_tab[k];   // .t / .omt: weight of the leading / trailing plane
    const float lead = FWD ? B[k + 1] : B[AZ - 1 - k], trail = FWD ? B[k] : B[AZ - k];
    val[k] = fmaf(gz.t, lead, gz.omt * trail);
  }
}

// ---- weighting and accumulation of one view's values -----------------------------------------------------------
// Two active axes: the third profile is 1 on the whole tile, min(f_a, f_b, 1) = min(f_a, f_b), and the lerp along the
// plateau axis is a lerp of equal values (exact) - same bits as ax_weight_3d at a third of the instructions.
// `eo` = entry of the outer (slower) axis, `ei` = entry of the inner one, in z, y, x nesting order.
__device__ __forceinline__ float ax_weight_2d(const WTab& eo, const WTab& ei) {
  const float v00 = fminf(eo.v0, ei.v0), v01 = fminf(eo.v0, ei.v1), v10 = fminf(eo.v1, ei.v0), v11 = fminf(eo.v1, ei.v1);
  const float i0 = fmaf(v01 - v00, ei.t, v00);
  const float i1 = fmaf(v11 - v10, ei.t, v10);
  const float w = fmaf(i1 - i0, eo.t, i0);
  return (eo.t >= 0.f && ei.t >= 0.f) ? w : 0.f;
}

enum { WC_NORMALISED = 0, WC_THREAD = 1, WC_ZTABLE = 2, WC_FIELD = 3, WC_ZT = 4 };

// WC selects where the blending weight comes from (one uniform branch per view instead of one per voxel).
template <typename T, int WC, bool WANT_F32>
__device__ __forceinline__ void ax_accumulate(const float (&val)[AZ], const WTab* __restrict__ wz_tab, float wconst, const WTab& ey,
                                              const WTab& ex, const float (&S)[AZ], const float (&rS)[AZ], uint32_t (&acc)[AZ],
                                              float (&accf)[AZ]) {
#pragma unroll
  for (int k = 0; k < AZ; ++k) {
    float wn;
    if (WC == WC_NORMALISED) wn = wconst;
    else if (WC == WC_THREAD) wn = div_by(wconst, S[k], rS[k]);
    else if (WC == WC_ZTABLE) wn = div_by(wz_tab[k].w1, S[k], rS[k]);
    else if (WC == WC_ZT) wn = div_by(ax_weight_2d(wz_tab[k], ey), S[k], rS[k]);   // `ey` = the active one of the thread's two entries
    else wn = div_by(ax_weight_3d(wz_tab[k], ey, ex), S[k], rS[k]);
    float vv = val[k];
    if (sizeof(T) == 2) vv = u16_to_float(round_half_up_u16(vv));
    acc[k] += trunc_product_u16(wn, vv);
    if (WANT_F32) accf[k] = fmaf(wn, vv, accf[k]);
  }
}

#ifndef MVF_AX_MINB
#define MVF_AX_MINB 2
#endif

template <typename T, bool WANT_F32>
__global__ void __launch_bounds__(ANT, MVF_AX_MINB)
fuse_axis_kernel(const __grid_constant__ TmaMaps TM, const __grid_constant__ AxArgs X) {
  extern __shared__ __align__(128) unsigned char ax_smem[];
  __shared__ GTab gt[MVF_MAX_VIEWS][ATAB];
  __shared__ WTab wt[MVF_MAX_VIEWS][ATAB];
  __shared__ AxTile tv[MVF_MAX_VIEWS];
  __shared__ __align__(8) uint64_t mbar[2];    // "brick b has landed" (TMA transaction count)
  __shared__ __align__(8) uint64_t mfree[2];   // "every warp is done reading brick b" (one arrival per warp)
  __shared__ int live[MVF_MAX_VIEWS], nlive_s, nweight_s, uniform_s;

  int z0, y0, x0;
  if (!ax_tile_origin(X.dims, blockIdx.x, z0, y0, x0)) return;
  const int tid = threadIdx.x;
  // TMA destinations must be 128-byte aligned whatever the static shared memory adds up to
  unsigned char* const bricks = ax_smem + ((128u - (smem_u32(ax_smem) & 127u)) & 127u);
  // TMA coordinates of a view's brick: (x, y, z) of the view, from the per-fused-axis starts
  auto issue = [&](int v, int b) {
    const AxView& av = X.v[v];
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) c[av.vax[a]] = tv[v].b0[a];   // view (z, y, x) start of the brick
    mbar_expect_tx(&mbar[b], (unsigned)av.bytes);
    tma_load_3d(bricks + b * X.brick_stride, &TM.m[v], &mbar[b], c[2], c[1], c[0]);
  };

  // Warp 7: tile geometry of every view, then the first two brick requests go out at once.  Warps 0-6 meanwhile copy
  // the tile's slices of the per-axis tables into shared memory (integer work only).
  const int loc0[3] = {z0, y0, x0};
  if (tid >= ANT - 32) {
    const int lane = tid - (ANT - 32);
    if (lane == 0) {   // independent of the records: overlaps their load
      mbar_init(&mbar[0], 1);
      mbar_init(&mbar[1], 1);
      mbar_init(&mfree[0], ANT / 32);
      mbar_init(&mfree[1], ANT / 32);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (lane < X.nviews) tv[lane] = X.tiles[(size_t)blockIdx.x * X.nviews + lane];
    __syncwarp();
    if (lane == 0) {
      int n = 0, nw = 0, uni = 1;
      for (int v = 0; v < X.nviews; ++v) {
        if (!tv[v].skip) live[n++] = v;
        if (tv[v].wcls != AW_ZERO) ++nw;
        if (tv[v].wcls > AW_ONE) uni = 0;
      }
      nlive_s = n; nweight_s = nw; uniform_s = uni;
      for (int b = 0; b < 2 && b < n; ++b) issue(live[b], b);
    }
  } else {
    const int ntot = X.dims[0] + X.dims[1] + X.dims[2];
    for (int e = tid; e < X.nviews * ATAB; e += ANT - 32) {
      const int v = e / ATAB, i = e - v * ATAB;
      const int a = i < AZ ? 0 : (i < AZ + AY ? 1 : 2);
      const int idx = i - (a == 0 ? 0 : (a == 1 ? AZ : AZ + AY));
      const int li = loc0[a] + idx, lc = min(li, X.dims[a] - 1);
      const size_t src = (size_t)v * ntot + (a == 0 ? 0 : (a == 1 ? X.dims[0] : X.dims[0] + X.dims[1])) + lc;
      GTab ge = X.gtab[src];
      if (a == 0) {
        ge.mode = idx == 0 ? 1 : (li > lc ? 0 : (ge.mode & 3));   // the first step finds its trailing plane in the prologue
      }
      gt[v][i] = ge;
      if (X.v[v].wmode != MVF_W_OFF) wt[v][i] = X.wtab[src];
    }
  }
  __syncthreads();

  const int ty = tid >> 5, tx = tid & 31;
  const int nlive = nlive_s;
  const bool uniform = uniform_s != 0;
  const float wn_uniform = __fdiv_rn(1.f, (float)max(nweight_s, 1));   // every weighted view has w = 1 here

  // ---- pass 1 over the weights (only tiles that touch a ramp): S = sum_v w_v, 0 -> 1 (mv:3627-3631) ---------
  float S[AZ], rS[AZ];
#pragma unroll
  for (int k = 0; k < AZ; ++k) { S[k] = 0.f; rS[k] = 0.f; }
  if (!uniform) {
    for (int v = 0; v < X.nviews; ++v) {
      const int cls = tv[v].wcls;
      if (cls == AW_ZERO) continue;
      if (cls == AW_Z) {
#pragma unroll
        for (int k = 0; k < AZ; ++k) S[k] += wt[v][k].w1;
      } else if (cls == AW_3D) {
        const WTab ey = wt[v][AZ + ty], ex = wt[v][AZ + AY + tx];
#pragma unroll
        for (int k = 0; k < AZ; ++k) S[k] += ax_weight_3d(wt[v][k], ey, ex);
      } else if (cls == AW_ZT) {
        const WTab et = (tv[v].wactive & 2) ? wt[v][AZ + ty] : wt[v][AZ + AY + tx];
#pragma unroll
        for (int k = 0; k < AZ; ++k) S[k] += ax_weight_2d(wt[v][k], et);
      } else {
        const float w = cls == AW_ONE ? 1.f : (cls == AW_Y ? wt[v][AZ + ty].w1 : (cls == AW_X ? wt[v][AZ + AY + tx].w1
                                                 : ax_weight_2d(wt[v][AZ + ty], wt[v][AZ + AY + tx])));
#pragma unroll
        for (int k = 0; k < AZ; ++k) S[k] += w;
      }
    }
#pragma unroll
    for (int k = 0; k < AZ; ++k) {
      S[k] = S[k] == 0.f ? 1.f : S[k];
      rS[k] = rcp_newton(S[k]);
    }
  }

  uint32_t acc[AZ];
  float accf[AZ];
#pragma unroll
  for (int k = 0; k < AZ; ++k) { acc[k] = 0u; accf[k] = 0.f; }

  for (int li = 0; li < nlive; ++li) {
    const int v = live[li], b = li & 1;
    const AxView& av = X.v[v];
    const T* __restrict__ br = reinterpret_cast<const T*>(bricks + b * X.brick_stride);
    const GTab gy = gt[v][AZ + ty], gx = gt[v][AZ + AY + tx];
    const int sY = av.str[1], sX = av.str[2];
    const T* __restrict__ p00 = br + (gy.off + gx.off - tv[v].corr);   // table offsets are absolute: subtract the brick start
    const T* __restrict__ p01 = p00 + sX;
    const T* __restrict__ p10 = p00 + sY;
    const T* __restrict__ p11 = p10 + sX;
    const float tY = gy.t, oY = gy.omt, tX = gx.t, oX = gx.omt;
    const int cls = tv[v].wcls;
    const bool regular = tv[v].regular != 0;
    const int dsz = av.dsz;
    const WTab ey = wt[v][AZ + ty], ex = wt[v][AZ + AY + tx];
    const float wthread = cls == AW_Y ? ey.w1 : (cls == AW_X ? ex.w1 : (cls == AW_YX ? ax_weight_2d(ey, ex) : 1.f));
    const WTab et = (tv[v].wactive & 2) ? ey : ex;   // AW_ZT: the thread's entry along the active one of y / x
    auto plane = [&](int off) {
      const float a = tap<T>(p00, off), bb = tap<T>(p01, off), c = tap<T>(p10, off), d = tap<T>(p11, off);
      const float r0 = fmaf(tX, bb, oX * a), r1 = fmaf(tX, d, oX * c);
      return fmaf(tY, r1, oY * r0);
    };
    mbar_wait(&mbar[b], (unsigned)((li >> 1) & 1));
    float val[AZ];
    if (sizeof(T) == 4 && sX != 1 && av.str[0] == 1 && tv[v].unit) {   // z walk along the view's x: conflict-free 16-byte window loads
      const float *q00 = reinterpret_cast<const float*>(p00), *q01 = reinterpret_cast<const float*>(p01),
                  *q10 = reinterpret_cast<const float*>(p10), *q11 = reinterpret_cast<const float*>(p11);
      if (dsz > 0) ax_walk_window<true>(gt[v], q00, q01, q10, q11, tX, oX, tY, oY, val);
      else ax_walk_window<false>(gt[v], q00, q01, q10, q11, tX, oX, tY, oY, val);
    } else if (tv[v].unit) {
      ax_walk_planes<true, true>(gt[v], plane, dsz, val);
    } else if (regular) {
      ax_walk_planes<true, false>(gt[v], plane, dsz, val);
    } else {
      ax_walk_planes<false, false>(gt[v], plane, dsz, val);
    }
    if (uniform) ax_accumulate<T, WC_NORMALISED, WANT_F32>(val, wt[v], wn_uniform, ey, ex, S, rS, acc, accf);
    else if (cls == AW_Z) ax_accumulate<T, WC_ZTABLE, WANT_F32>(val, wt[v], 0.f, ey, ex, S, rS, acc, accf);
    else if (cls == AW_3D) ax_accumulate<T, WC_FIELD, WANT_F32>(val, wt[v], 0.f, ey, ex, S, rS, acc, accf);
    else if (cls == AW_ZT) ax_accumulate<T, WC_ZT, WANT_F32>(val, wt[v], 0.f, et, ex, S, rS, acc, accf);
    else ax_accumulate<T, WC_THREAD, WANT_F32>(val, wt[v], wthread, ey, ex, S, rS, acc, accf);
    // The buffer goes back to the TMA once every warp has left it; only the issuing thread waits for that, the other
    // warps move on to the next view (whose brick sits in the other buffer) without a CTA-wide barrier.
    if (li + 2 < nlive) {
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&mfree[b]);
      if (tid == 0) {
        mbar_wait(&mfree[b], (unsigned)((li >> 1) & 1));
        issue(live[li + 2], b);
      }
    }
  }

  const int y = y0 + ty, x = x0 + tx;
  if (y >= X.dims[1] || x >= X.dims[2]) return;
  uint16_t* out = reinterpret_cast<uint16_t*>(X.out) + ((size_t)z0 * X.dims[1] + y) * X.dims[2] + x;
  float* outf = WANT_F32 ? X.outf + ((size_t)z0 * X.dims[1] + y) * X.dims[2] + x : nullptr;
  const size_t pl = (size_t)X.dims[1] * X.dims[2];
#pragma unroll
  for (int k = 0; k < AZ; ++k) {
    if (z0 + k >= X.dims[0]) break;
    out[k * pl] = (uint16_t)(sizeof(T) == 2 ? (acc[k] & 0xFFFFu) : min(acc[k], 65535u));
    if (WANT_F32) outf[k * pl] = accf[k];
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tensor_map_encode_fn tensor_map_encoder() {
  static tensor_map_encode_fn fn = []() -> tensor_map_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<tensor_map_encode_fn>(p);
  }();
  return fn;
}

// Is view `h` (index map matrix/offset, blending map wmatrix/woffset) a signed permutation times a diagonal on the
// box [org, org+dims)?  Off-axis terms that move a coordinate by less than 1e-9 voxels over the box (cos(pi/2) =
// 6e-17 of a float64 rotation matrix) count as zero.  Fills vax/box/str/bytes.
static bool axis_view(const mvf_view& h, const int32_t dims[3], const int32_t org[3], AxView& av) {
  const int es = h.dtype == MVF_U16 ? 2 : 4, vec = 16 / es;
  if ((reinterpret_cast<uintptr_t>(h.data) & 15u) || (h.dims[2] % vec) != 0) return false;   // TMA: 16-byte base and strides
  const int T[3] = {AZ, AY, AX};
  int used = 0;
  for (int d = 0; d < 3; ++d) {
    int a = 0;
    for (int l = 1; l < 3; ++l)
      if (fabs(h.matrix[d * 3 + l]) > fabs(h.matrix[d * 3 + a])) a = l;
    if (used & (1 << a)) return false;
    used |= 1 << a;
    for (int l = 0; l < 3; ++l) {
      if (l == a) continue;
      const double reach = (double)(fabs((double)org[l]) + dims[l] + T[l]);
      if (fabs(h.matrix[d * 3 + l]) * reach > 1e-9) return false;
      if (h.weight_mode != MVF_W_OFF && fabs(h.wmatrix[d * 3 + l]) * reach > 1e-9) return false;
    }
    const double sc = fabs(h.matrix[d * 3 + a]);
    if (!(sc > 0.0) || !(sc < 64.0)) return false;
    av.vax[a] = d;
    av.box[d] = (int)floor(sc * (T[a] - 1)) + 3;
    av.sc[a] = h.matrix[d * 3 + a];
    av.of[a] = h.offset[d];
    av.wsc[a] = h.wmatrix[d * 3 + a];
    av.wof[a] = h.woffset[d];
    av.n[a] = h.dims[d];
    av.one_lo[a] = h.one_lo[d];
    av.one_hi[a] = h.one_hi[d];
  }
  av.prof = h.profiles;
  av.wmode = h.weight_mode;
  // dense boxes: rows of whole 16-byte chunks; (chunks per row) and (rows per plane) odd, so that a warp walking
  // along view y or z meets 8 different bank groups instead of 1 or 2
  av.xalign = vec;
  av.box[2] = (av.box[2] + (vec - 1) + vec - 1) / vec;   // + (vec - 1): the box start is aligned down to 16 bytes
  av.box[2] = (av.box[2] | 1) * vec;
  av.box[1] |= 1;
  if (av.box[0] > 256 || av.box[1] > 256 || av.box[2] > 256) return false;
  const int str_of[3] = {av.box[1] * av.box[2], av.box[2], 1};
  for (int a = 0; a < 3; ++a) av.str[a] = str_of[av.vax[a]];
  av.dsz = av.sc[0] > 0 ? av.str[0] : -av.str[0];
  av.bytes = av.box[0] * av.box[1] * av.box[2] * es;
  return av.bytes <= 48 * 1024;
}

static int encode_view_map(const mvf_view& h, const AxView& av, CUtensorMap* map) {
  tensor_map_encode_fn enc = tensor_map_encoder();
  if (!enc) return fail(MVF_ERR_CUDA, "%s%s", "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t es = h.dtype == MVF_U16 ? 2 : 4;
  const cuuint64_t gdim[3] = {(cuuint64_t)h.dims[2], (cuuint64_t)h.dims[1], (cuuint64_t)h.dims[0]};
  const cuuint64_t gstr[2] = {(cuuint64_t)h.dims[2] * es, (cuuint64_t)h.dims[2] * h.dims[1] * es};
  const cuuint32_t box[3] = {(cuuint32_t)av.box[2], (cuuint32_t)av.box[1], (cuuint32_t)av.box[0]};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, h.dtype == MVF_U16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(h.data), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MVF_ERR_CUDA, "%s%s", "cuTensorMapEncodeTiled failed");
  return MVF_OK;
}

static dim3 ax_tile_grid(const int32_t dims[3]) {
  unsigned nx = (dims[2] + AX * AST - 1) / (AX * AST), ny = (dims[1] + AY * AST - 1) / (AY * AST), nz = (dims[0] + AZ * AST - 1) / (AZ * AST);
  return dim3(nx * ny * nz * AST * AST * AST);
}

// Launches the axis-aligned kernel when every view qualifies; *taken = 0 (and nothing launched) otherwise.
static int try_fuse_axis(int nviews, const mvf_view* hv, const KArgs& A, cudaStream_t st, int* taken) {
  *taken = 0;
  const char* sel = getenv("MVF_FUSE_KERNEL");   // "general": always the per-voxel kernel (A/B runs, tests)
  if (sel && !strcmp(sel, "general")) return MVF_OK;
  AxArgs X;
  memset(&X, 0, sizeof(X));
  static_assert(sizeof(AxArgs) + sizeof(TmaMaps) <= 4096, "kernel parameters (tensor maps included) must stay below 4 KB");
  X.out = A.out; X.outf = A.outf; X.nviews = nviews;
  for (int d = 0; d < 3; ++d) { X.dims[d] = A.dims[d]; X.org[d] = A.org[d]; }
  int maxb = 0;
  for (int v = 0; v < nviews; ++v) {
    if (hv[v].weight_mode == MVF_W_BLEND_DCT) return MVF_OK;
    if (!axis_view(hv[v], A.dims, A.org, X.v[v])) return MVF_OK;
    maxb = X.v[v].bytes > maxb ? X.v[v].bytes : maxb;
  }
  X.brick_stride = (maxb + 127) & ~127;
  TmaMaps TM;
  for (int v = 0; v < nviews; ++v) {
    int rc = encode_view_map(hv[v], X.v[v], &TM.m[v]);
    if (rc) return rc;
  }
  const size_t smem = 2 * (size_t)X.brick_stride + 128;
  // per-axis tables: stream-ordered scratch, filled by one small kernel in front of the fused kernel
  const size_t nent = (size_t)nviews * ((size_t)A.dims[0] + A.dims[1] + A.dims[2]);
  static thread_local int pool_ready_dev = -1;   // keep the stream-ordered pool's memory across synchronisations
  int cur_dev = 0;
  MVF_CUDA(cudaGetDevice(&cur_dev));
  if (pool_ready_dev != cur_dev) {
    cudaMemPool_t pool;
    unsigned long long keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, cur_dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pool_ready_dev = cur_dev;
  }
  void* scratch = nullptr;
  const unsigned ntiles = ax_tile_grid(A.dims).x;
  MVF_CUDA(cudaMallocAsync(&scratch, nent * (sizeof(GTab) + sizeof(WTab)) + (size_t)ntiles * nviews * sizeof(AxTile), st));
  X.gtab = reinterpret_cast<GTab*>(scratch);
  X.wtab = reinterpret_cast<WTab*>(X.gtab + nent);
  X.tiles = reinterpret_cast<AxTile*>(X.wtab + nent);
  ax_tables_kernel<<<(unsigned)((nent + 255) / 256), 256, 0, st>>>(X);
  ax_tiles_kernel<<<(ntiles * nviews + 127) / 128, 128, 0, st>>>(X, ntiles);
  mvf::launch_counter() += 2;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      kern<<<ax_tile_grid(A.dims), ANT, smem, st>>>(TM, X);
      mvf::launch_counter()++;
      e = cudaGetLastError();
    }
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return fail(MVF_ERR_CUDA, "%s: %s", "fuse_axis_kernel", cudaGetErrorString(e));
    return MVF_OK;
  };
  *taken = 1;
  const bool f32 = A.outf != nullptr;
  if (hv[0].dtype == MVF_U16) return f32 ? launch(fuse_axis_kernel<uint16_t, true>) : launch(fuse_axis_kernel<uint16_t, false>);
  return f32 ? launch(fuse_axis_kernel<float, true>) : launch(fuse_axis_kernel<float, false>);
}

}  // namespace mvf
