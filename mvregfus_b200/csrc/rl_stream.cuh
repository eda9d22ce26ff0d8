// K4/K5, streaming formulation of the separable RL blur: two kernels without shared memory and without barriers.
//
//   rl_xy_kernel : tmp[p] = blur_y(blur_x(in[p])) for every plane p of the input buffer.  A thread owns CW
//                  consecutive x columns and walks along y: per row it loads the x window of its columns straight
//                  from global memory (aligned LDG.128/64; neighbouring threads share the lines through L1, the
//                  thread's own chunk is prefetched into L2 a few rows ahead), does the x pass in registers and
//                  pushes the CW results into a K-deep shift register of partial sums along y.
//   rl_z_kernel  : out = epilogue(blur_z(tmp)).  A thread owns CW consecutive voxels of a plane and walks along z
//                  through the caller's plane map with the same shift register; the RL pointwise algebra (ratio /
//                  correction / update, identical to rl_blur_kernel) runs on 16-byte vectors.
//
// Per voxel and pass: one coalesced vector load, K FFMA with the tap in a uniform register (full rate for scalar
// FFMA), one vector store - ~90 % of the issued instructions are FFMA, against 36 % in the fused kernel, whose x/y
// passes pay for shared-memory windows, halos (1.56x redundant x pass) and two barriers per plane.  The price is one
// extra round trip of a float32 volume through HBM (22 B instead of 14 B per voxel and blur), which the kernels
// can afford: the fused kernel is issue-bound at a quarter of the HBM bandwidth.
// Accumulation order per output is fixed (x taps ascending, y and z contributions in ascending row / plane
// order), so z-slabs and chunks reproduce the whole volume bit for bit, as in the fused kernel.
#pragma once
#include <type_traits>

namespace mvf {

template <int CW> struct VecOf;
template <> struct VecOf<4> { using type = float4; };
template <> struct VecOf<2> { using type = float2; };

template <int CW> __device__ __forceinline__ void vec_load(const float* p, float (&v)[CW]) {
  using V = typename VecOf<CW>::type;
  const V q = __ldg(reinterpret_cast<const V*>(p));
  if constexpr (CW == 4) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; } else { v[0] = q.x; v[1] = q.y; }
}
template <int CW> __device__ __forceinline__ void vec_store(float* p, const float (&v)[CW]) {
  using V = typename VecOf<CW>::type;
  if constexpr (CW == 4) *reinterpret_cast<V*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else *reinterpret_cast<V*>(p) = make_float2(v[0], v[1]);
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

#ifndef MVF_RLS_NT
#define MVF_RLS_NT 256
#endif
#ifndef MVF_RLS_MINB
#define MVF_RLS_MINB 2
#endif
constexpr int RLS_NT = MVF_RLS_NT;   // threads per CTA of the streaming kernels, MVF_RLS_MINB CTAs per SM
#ifndef MVF_RLS_AHEAD
#define MVF_RLS_AHEAD 6
#endif
#ifndef MVF_RLS_NEAR
#define MVF_RLS_NEAR 0
#endif
constexpr int RLS_AHEAD = MVF_RLS_AHEAD;   // rows / planes the L2 prefetch runs ahead of the loads
constexpr int RLS_NEAR = MVF_RLS_NEAR;     // rows / planes the L1 prefetch runs ahead (0: off)

// ---- x and y passes of every plane -------------------------------------------------------------------------
// MODE == RL_PLAIN: tmp = blur_y(blur_x(in)).  MODE == RL_CORRECT: the x/y passes are the LAST passes of a blur whose z
// pass has already run (rl_zz_kernel), so the correction epilogue is applied to every finished row on the spot:
//   C (+)= max(row, 0) * w ; last view: est = clip(est * C, 0, 65535), sum(est) accumulated.
template <int K, int CW, int MODE>
// (43 taps: window + shift register exceed 128 registers, one CTA per SM)
__global__ void __launch_bounds__(RLS_NT, K > 37 ? 1 : MVF_RLS_MINB) rl_xy_kernel(const __grid_constant__ RLArgs A) {
  constexpr int R = (K - 1) / 2;
  constexpr int LEAD = (R + CW - 1) / CW * CW;       // x halo in whole chunks
  constexpr int NCH = (CW + 2 * LEAD) / CW;          // chunks of the x window
  const int x0raw = (blockIdx.x * RLS_NT + threadIdx.x) * CW;
  if (MODE == RL_PLAIN && x0raw >= A.nx) return;
  const int x0 = x0raw < A.nx ? x0raw : 0;           // (epilogue mode: every lane stays for the block sum)
  const int yo0 = blockIdx.y * A.zchunk, yo1 = min(A.ny, yo0 + A.zchunk);   // zchunk doubles as the y chunk here
  const size_t plane = (size_t)A.ny * A.nx;
  const float* __restrict__ src = A.in + (size_t)blockIdx.z * plane;
  float* __restrict__ dst = A.out + (size_t)blockIdx.z * plane;
  const float* __restrict__ wts = MODE == RL_CORRECT ? A.weights + (size_t)blockIdx.z * plane : nullptr;
  float* __restrict__ est = (MODE == RL_CORRECT && A.est) ? A.est + A.est_off + (size_t)blockIdx.z * plane : nullptr;
  double local_sum = 0.0;
  // Threads whose window would leave [0, nx) (the first / last LEAD columns of a row) shift it back inside: their
  // results are placeholders that rl_xy_edge_kernel overwrites afterwards with the index-mapped values.  Every
  // thread of this kernel therefore runs the same branch-free vector code.
  const int xw = min(max(x0, LEAD), A.nx - CW - LEAD);
  // epilogue mode: the border threads must not touch C / est with their placeholder rows
  const bool own = MODE != RL_CORRECT || (x0raw < A.nx && x0 == xw);
  float eW[CW], eC[CW], eI[CW];
  float S[CW][K - 1];
#pragma unroll
  for (int i = 0; i < CW; ++i)
#pragma unroll
    for (int m = 0; m < K - 1; ++m) S[i][m] = 0.f;

  // Software pipeline: the window of row t+1 is requested right after the x pass of row t has consumed the window
  // registers, so the loads have the whole y pass (CW * K FFMA) to land.
  float win[NCH * CW];
  // Row pointers advance by nx inside the volume; only rows that go through the index map (t < 0 or t >= ny, i.e.
  // the ramps at the volume borders) pay for ext_index.
  auto row_ptr = [&](int t) { return src + (size_t)ext_index(t, A.ny, A.ksz_y) * A.nx; };
  auto load_window = [&](const float* __restrict__ rowp) {   // rowp: first column of the thread's window
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float v[CW];
      vec_load<CW>(rowp + j * CW, v);
#pragma unroll
      for (int i = 0; i < CW; ++i) win[j * CW + i] = v[i];
    }
  };
  // per-thread pointers (source offsets include the thread's columns): row t, row t + RLS_AHEAD; output row t - R as an offset
  const float* __restrict__ cur = row_ptr(yo0 - R) + (xw - LEAD);
  const float* __restrict__ far = row_ptr(yo0 - R + RLS_AHEAD) + x0;
  ptrdiff_t ooff = (ptrdiff_t)(yo0 - 2 * R) * A.nx + x0;   // offset of output row t - R in dst; first stored at t = yo0 + R
  load_window(cur);
  // One step = x pass of row t, request of row t+1, y pass, store of output row t - R.  INTERIOR steps know that the
  // rows t+1 and t+RLS_AHEAD+1 lie inside the volume: all three pointers just advance by nx (no index map, no
  // multiplies); the few steps at the chunk / volume borders take the general path.
  auto step = [&](int t, auto interior_tag) {
    constexpr bool INTERIOR = decltype(interior_tag)::value;
    if (RLS_AHEAD > 0 && (INTERIOR || t + RLS_AHEAD < yo1 + R)) prefetch_l2(far);
    if (MODE == RL_CORRECT && own && (INTERIOR || t - R >= yo0)) {
      // operands of the row this step finishes: pulled into L1 now (no registers while the passes run), loaded after
      // the y pass; the rows RLS_AHEAD steps ahead go to L2
      prefetch_l1(wts + ooff);
      if (!A.first) prefetch_l1(dst + ooff);
      if (A.last) prefetch_l1(est + ooff);
      if (RLS_AHEAD > 0 && (INTERIOR || t - R + RLS_AHEAD < yo1)) {
        const ptrdiff_t oa = ooff + (ptrdiff_t)RLS_AHEAD * A.nx;
        prefetch_l2(wts + oa);
        if (!A.first) prefetch_l2(dst + oa);
        if (A.last) prefetch_l2(est + oa);
      }
    }
    float o[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {   // out[x] = sum_j k[j] in[x + R - j]
      const float kj = A.kx[j];
#pragma unroll
      for (int i = 0; i < CW; ++i) o[i] = fmaf(kj, win[LEAD + i + R - j], o[i]);
    }
    if (INTERIOR) {
      cur += A.nx;
      far += A.nx;
      load_window(cur);
    } else {
      cur = row_ptr(t + 1) + (xw - LEAD);
      far = row_ptr(t + RLS_AHEAD + 1) + x0;
      if (t + 1 < yo1 + R) load_window(cur);
    }
    float e[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      e[i] = fmaf(A.ky[0], o[i], S[i][0]);
#pragma unroll
      for (int m = 0; m < K - 2; ++m) S[i][m] = fmaf(A.ky[m + 1], o[i], S[i][m + 1]);
      S[i][K - 2] = A.ky[K - 1] * o[i];
    }
    if (MODE == RL_CORRECT) {
      if (own && (INTERIOR || t - R >= yo0)) {
        vec_load<CW>(wts + ooff, eW);
        if (!A.first) vec_load<CW>(dst + ooff, eC);
        if (A.last) vec_load<CW>(est + ooff, eI);
        float r[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const float ow = fmaxf(e[i], 0.f) * eW[i];
          const float corr = A.first ? ow : eC[i] + ow;
          if (A.last) {
            r[i] = fminf(fmaxf(eI[i] * corr, 0.f), 65535.f);
            local_sum += (double)r[i];
          } else {
            r[i] = corr;
          }
        }
        vec_store<CW>((A.last ? est : dst) + ooff, r);
      }
    } else if (INTERIOR || t - R >= yo0) {
      vec_store<CW>(dst + ooff, e);
    }
    ooff += A.nx;
  };
  // interior steps: the output row exists (t - R >= yo0), rows t+1 and t+RLS_AHEAD+1 are real rows, t+1 is needed
  const int t_lo = max(yo0 + R, 0), t_hi = min(yo1 + R - 1, A.ny - RLS_AHEAD - 1);
  int t = yo0 - R;
  for (; t < min(t_lo, yo1 + R); ++t) step(t, std::false_type{});
  for (; t < t_hi; ++t) step(t, std::true_type{});
  for (; t < yo1 + R; ++t) step(t, std::false_type{});
  if (MODE == RL_CORRECT && A.last && A.sum) {
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) local_sum += __shfl_xor_sync(0xffffffffu, local_sum, sh);
    if ((threadIdx.x & 31) == 0) atomicAdd(A.sum, local_sum);
  }
}

// The LEAD columns next to each x border of every row (the main kernel leaves placeholders / skips them there).  One CTA
// per (side, 64-row chunk, plane): the index-mapped source strip of the whole chunk is staged in shared memory with all
// loads in flight at once, the x pass of every (row, column) pair and then the y pass + epilogue of every output run from
// there.  Per output the taps are accumulated in the order of the main kernel's shift registers (x ascending from 0; y
// from the oldest row), independent of the chunking.
constexpr int RLE_NT = 128, RLE_ROWS = 64;
template <int K, int CW, int MODE>
__global__ void __launch_bounds__(RLE_NT) rl_xy_edge_kernel(const __grid_constant__ RLArgs A) {
  constexpr int R = (K - 1) / 2;
  constexpr int LEAD = (R + CW - 1) / CW * CW;
  constexpr int W = LEAD + 2 * R;               // strip width: the x windows of LEAD columns
  extern __shared__ float edge_smem[];
  const int yo0 = blockIdx.y * RLE_ROWS, yo1 = min(A.ny, yo0 + RLE_ROWS);
  const int nrows = yo1 - yo0 + 2 * R;          // source rows yo0 - R .. yo1 + R - 1
  float* __restrict__ strip = edge_smem;        // [nrows][W]
  float* __restrict__ xres = edge_smem + (RLE_ROWS + 2 * R) * W;   // [nrows][LEAD]
  const int xs = blockIdx.x == 0 ? 0 : A.nx - LEAD;   // first output column of this side
  const size_t plane = (size_t)A.ny * A.nx;
  const float* __restrict__ src = A.in + (size_t)blockIdx.z * plane;
  float* __restrict__ dst = A.out + (size_t)blockIdx.z * plane;
  const float* __restrict__ wts = MODE == RL_CORRECT ? A.weights + (size_t)blockIdx.z * plane : nullptr;
  float* __restrict__ est = (MODE == RL_CORRECT && A.est) ? A.est + A.est_off + (size_t)blockIdx.z * plane : nullptr;
  // Thread <-> (column, row group): the column part of every index is computed once per thread, rows advance by a
  // constant - no division and one index map per element (the kernel was bound by its integer arithmetic).
  // Staging: fixed trip counts, every load of a thread issued before the first store (one memory round trip per CTA).
  constexpr int WP = W <= 32 ? 32 : 64, RPW = RLE_NT / WP;          // threads per strip row, rows per pass
  constexpr int NLOAD = (RLE_ROWS + 2 * R + RPW - 1) / RPW;
  {
    const int i = threadIdx.x % WP, r0 = threadIdx.x / WP;
    const int colx = ext_index(xs + i - R, A.nx, A.ksz_x);
    float v[NLOAD];
#pragma unroll
    for (int k = 0; k < NLOAD; ++k) {
      const int r = r0 + k * RPW;
      v[k] = 0.f;
      if (i < W && r < nrows) v[k] = __ldg(src + (size_t)ext_index(yo0 - R + r, A.ny, A.ksz_y) * A.nx + colx);
    }
#pragma unroll
    for (int k = 0; k < NLOAD; ++k) {
      const int r = r0 + k * RPW;
      if (i < W && r < nrows) strip[r * W + i] = v[k];
    }
  }
  __syncthreads();
  constexpr int LP = LEAD <= 16 ? 16 : 32, RPL = RLE_NT / LP;       // threads per output row, rows per pass
  const int c = threadIdx.x % LP, rl0 = threadIdx.x / LP;
  if (c < LEAD) {
    for (int r = rl0; r < nrows; r += RPL) {   // out[x] = sum_j k[j] in[x + R - j]
      const float* __restrict__ w = strip + r * W + c + 2 * R;
      float o = 0.f;
#pragma unroll
      for (int j = 0; j < K; ++j) o = fmaf(A.kx[j], w[-j], o);
      xres[r * LEAD + c] = o;
    }
  }
  __syncthreads();
  double local_sum = 0.0;
  constexpr int NOUT = (RLE_ROWS + RPL - 1) / RPL;
  const int nout = yo1 - yo0;
  float eW[NOUT], eC[NOUT], eI[NOUT];
  if (MODE == RL_CORRECT) {   // epilogue operands of all outputs of the thread in flight before the y passes
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      const int r = rl0 + k * RPL;
      eW[k] = eC[k] = eI[k] = 0.f;
      if (c < LEAD && r < nout) {
        const size_t o1 = (size_t)(yo0 + r) * A.nx + xs + c;
        eW[k] = __ldg(wts + o1);
        if (!A.first) eC[k] = dst[o1];
        if (A.last) eI[k] = est[o1];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
    const int r = rl0 + k * RPL;                 // output row yo0 + r: x-pass rows r .. r + 2R of the strip
    if (c >= LEAD || r >= nout) continue;
    const size_t o1 = (size_t)(yo0 + r) * A.nx + xs + c;
    const float* __restrict__ col = xres + r * LEAD + c;
    float acc = A.ky[K - 1] * col[0];            // oldest row first, as the shift register does
#pragma unroll
    for (int j = K - 2; j >= 0; --j) acc = fmaf(A.ky[j], col[(K - 1 - j) * LEAD], acc);
    if (MODE == RL_CORRECT) {
      const float ow = fmaxf(acc, 0.f) * eW[k];
      const float corr = A.first ? ow : eC[k] + ow;
      if (A.last) {
        const float rr = fminf(fmaxf(eI[k] * corr, 0.f), 65535.f);
        est[o1] = rr;
        local_sum += (double)rr;
      } else {
        dst[o1] = corr;
      }
    } else {
      dst[o1] = acc;
    }
  }
  if (MODE == RL_CORRECT && A.last && A.sum) {
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) local_sum += __shfl_xor_sync(0xffffffffu, local_sum, sh);
    if ((threadIdx.x & 31) == 0) atomicAdd(A.sum, local_sum);
  }
}

// ---- z pass + RL epilogue ---------------------------------------------------------------------------------------
template <int K, int MODE, int CW>
__global__ void __launch_bounds__(RLS_NT, MVF_RLS_MINB) rl_z_kernel(const __grid_constant__ RLArgs A) {
  constexpr int R = (K - 1) / 2;
  const size_t plane = (size_t)A.ny * A.nx;
  // Threads beyond the plane (last CTA) stay alive - the block sum below shuffles across the full warp - but work on
  // voxel 0 of the plane and neither store nor count.
  const size_t i0raw = ((size_t)blockIdx.x * RLS_NT + threadIdx.x) * CW;   // first voxel of the thread inside a plane
  const bool active = i0raw < plane;
  const size_t i0 = active ? i0raw : 0;
  const int zo0 = blockIdx.y * A.zchunk, zo1 = min(A.nz, zo0 + A.zchunk);
  const int* __restrict__ zmap = A.zmap + R;    // zmap[t], t in [-R, nz + R)
  float S[CW][K - 1];
#pragma unroll
  for (int i = 0; i < CW; ++i)
#pragma unroll
    for (int m = 0; m < K - 1; ++m) S[i][m] = 0.f;
  const float* __restrict__ wts = A.weights;
  float* __restrict__ outp = A.out;
  float* __restrict__ est = A.est ? A.est + A.est_off : nullptr;
  double local_sum = 0.0;

  // Software pipeline: the input plane of step t+1 is requested before the multiply-adds of step t (one extra
  // vector of registers), the epilogue operands of step t+1 right after the epilogue of step t has consumed theirs.
  float p[CW], pn[CW];
  float eW[CW], eC[CW], eI[CW];
  unsigned eRaw[2] = {0u, 0u};   // uint16 view voxels stay packed until they are used
  // ratio mode with weights == NULL: the caller has zeroed the view wherever w <= 1e-5 (I = 0 gives R = 0 exactly),
  // so the mask costs no 4-byte read per voxel
  const bool have_w = wts != nullptr;
#pragma unroll
  for (int i = 0; i < CW; ++i) eW[i] = 1.f;
  // Plane offsets are running 64-bit sums (one add per step) or one 32 x 32 -> 64 bit multiply of a plane index, never a
  // 64 x 64 bit product: address arithmetic was a quarter of this kernel's instructions.
  const unsigned plane32 = (unsigned)plane;
  auto in_plane = [&](int zm) { return A.in + ((size_t)((unsigned long long)(unsigned)zm * plane32) + i0); };
  auto load_operands = [&](size_t o) {
    if (have_w) vec_load<CW>(wts + o, eW);
    if (MODE == RL_RATIO) {
      if (A.view_u16) {
        const uint16_t* vp = reinterpret_cast<const uint16_t*>(A.view) + o;
        if constexpr (CW == 4) {
          const uint2 q = __ldg(reinterpret_cast<const uint2*>(vp));
          eRaw[0] = q.x; eRaw[1] = q.y;
        } else {
          eRaw[0] = __ldg(reinterpret_cast<const unsigned*>(vp));
        }
      } else {
        vec_load<CW>(reinterpret_cast<const float*>(A.view) + o, eI);
      }
    }
    if (MODE == RL_CORRECT) {
      if (!A.first) vec_load<CW>(outp + o, eC);
      if (A.last) vec_load<CW>(est + o, eI);
    }
  };
  vec_load<CW>(in_plane(__ldg(zmap + zo0 - R)), p);
  if (MODE != RL_PLAIN && R == 0) load_operands((size_t)zo0 * plane + i0);
  size_t o = (size_t)((long long)(zo0 - 2 * R) * (long long)plane) + i0;   // offset of output plane z = t - R (meaningful once z >= zo0)
  // step t consumes `cur` (plane t) and requests plane t+1 into `nxt`; the two buffers swap by NAME (loop unrolled twice):
  // a register move out of a buffer whose load is still in flight would wait for the load
  auto step = [&](int t, float (&cur)[CW], float (&nxt)[CW]) {
    const int z = t - R;
    const bool emit = z >= zo0;
    if (RLS_AHEAD > 0 && t + RLS_AHEAD < zo1 + R) {
      prefetch_l2(in_plane(__ldg(zmap + t + RLS_AHEAD)));
      if (MODE != RL_PLAIN && z + RLS_AHEAD >= zo0) {   // epilogue operands of the plane completed RLS_AHEAD steps from now
        const size_t oa = o + (size_t)RLS_AHEAD * plane;
        if (have_w) prefetch_l2(wts + oa);
        if (MODE == RL_RATIO) prefetch_l2(reinterpret_cast<const unsigned char*>(A.view) + oa * (A.view_u16 ? 2 : 4));
        if (MODE == RL_CORRECT && !A.first) prefetch_l2(outp + oa);
        if (MODE == RL_CORRECT && A.last) prefetch_l2(est + oa);
      }
    }
    if (t + 1 < zo1 + R) vec_load<CW>(in_plane(__ldg(zmap + t + 1)), nxt);
    float e[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      e[i] = fmaf(A.kz[0], cur[i], S[i][0]);
#pragma unroll
      for (int m = 0; m < K - 2; ++m) S[i][m] = fmaf(A.kz[m + 1], cur[i], S[i][m + 1]);
      S[i][K - 2] = A.kz[K - 1] * cur[i];
    }
    if (emit) {
      if (MODE == RL_RATIO && A.view_u16) {
        if constexpr (CW == 4) {
          eI[0] = u16_to_float(eRaw[0] & 0xFFFFu); eI[1] = u16_to_float(eRaw[0] >> 16);
          eI[2] = u16_to_float(eRaw[1] & 0xFFFFu); eI[3] = u16_to_float(eRaw[1] >> 16);
        } else {
          eI[0] = u16_to_float(eRaw[0] & 0xFFFFu); eI[1] = u16_to_float(eRaw[0] >> 16);
        }
      }
      float r[CW];
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        const float blur = fmaxf(e[i], 0.f);   // img_ft[img_ft<0] = 0 (mv:5671)
        if (MODE == RL_PLAIN) {
          r[i] = blur;
        } else if (MODE == RL_RATIO) {
          const float ratio = eI[i] * rcp_approx(blur + 1e-6f);   // divisor >= 1e-6: no denormal / huge-divisor guards needed
          r[i] = eW[i] > 1e-5f ? ratio : 0.f;
        } else {
          const float ow = blur * eW[i];
          const float corr = A.first ? ow : eC[i] + ow;
          if (A.last) {
            r[i] = fminf(fmaxf(eI[i] * corr, 0.f), 65535.f);
            if (active) local_sum += (double)r[i];
          } else {
            r[i] = corr;
          }
        }
      }
      if (active) vec_store<CW>((MODE == RL_CORRECT && A.last) ? est + o : outp + o, r);
    }
    // operands of the plane the next step completes
    if (MODE != RL_PLAIN && z + 1 >= zo0 && z + 1 < zo1) load_operands(o + plane);
    o += plane;
  };
  int t = zo0 - R;
  for (; t + 1 < zo1 + R; t += 2) {
    step(t, p, pn);
    step(t + 1, pn, p);
  }
  if (t < zo1 + R) step(t, p, pn);
  if (MODE == RL_CORRECT && A.last && A.sum) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) local_sum += __shfl_xor_sync(0xffffffffu, local_sum, s);
    if ((threadIdx.x & 31) == 0) atomicAdd(A.sum, local_sum);
  }
}

// ---- z pass of the first blur + ratio + z pass of the second blur --------------------------------------------------
// T2 = blur_z(R),  R = I / (max(blur_z(T1), 0) + 1e-6) * mask,  T1 = blur_y(blur_x(estimate)): the ratio volume never
// reaches HBM.  The passes of a separable blur commute, so the second blur runs z first and finishes with the x/y passes
// (rl_xy_kernel<RL_CORRECT>), which apply max(., 0) and the correction epilogue.  A thread owns CW voxels of a plane and
// walks along z with two shift registers: step s pushes T1[zmap1[s]] into the first (blur_z(T1) at t = s - R completes),
// turns it into the ratio at t and pushes that into the second (T2 at z = s - 2R completes).
//   zmap1[s], s in [-2R, nz + 2R): plane of T1 (the reference's boundary map / halo planes, as for rl_z_kernel)
//   rmap[t],  t in [-R, nz + R)  : >= 0: plane of the view buffer that holds I at extended position t;
//                                  <  0: the ratio at t is plane -1 - rmap[t] of `side` - extended positions outside
//                                  the volume repeat ratios of planes near the top (reflect above, wrap below), which
//                                  the caller has computed beforehand with rl_z_kernel<RL_RATIO> on those few planes.
// The two plane maps are piecewise affine (runs of constant step: volume interior, reflected / wrapped ramps, clamped
// tails), so they travel in the kernel parameters as a handful of segments and are evaluated by cursors - an add and a
// compare per step on warp-uniform values - instead of dependent loads that would sit in front of every address.
constexpr int ZZ_MAXSEG = 8;
struct MapSeg { int start, base, step; };
struct MapCursor { int v, step, next, idx; };
struct RLZZArgs {
  const float* in;
  MapSeg seg1[ZZ_MAXSEG], segr[ZZ_MAXSEG];   // zmap1 over s in [-2R, nz + 2R), rmap over t in [-R, nz + R)
  int nseg1, nsegr;
  float kz[RL_KMAX];
  int nz, ny, nx, zchunk;
  const void* view;   // masked view (zero where w <= 1e-5), uint16 or float32
  int view_u16;
  const float* side;
  float* out;
};

__device__ __forceinline__ void cursor_seek(MapCursor& c, const MapSeg* segs, int nseg, int pos) {
  int i = 0;
  while (i + 1 < nseg && segs[i + 1].start <= pos) ++i;
  c.idx = i;
  c.step = segs[i].step;
  c.v = segs[i].base + segs[i].step * (pos - segs[i].start);
  c.next = i + 1 < nseg ? segs[i + 1].start : 0x7fffffff;
}
// value at position `pos` (the cursor stood at pos - 1)
__device__ __forceinline__ void cursor_step(MapCursor& c, const MapSeg* segs, int nseg, int pos) {
  if (pos == c.next) {
    const int i = ++c.idx;
    c.v = segs[i].base;
    c.step = segs[i].step;
    c.next = i + 1 < nseg ? segs[i + 1].start : 0x7fffffff;
  } else {
    c.v += c.step;
  }
}

// Operands of one ratio step: the view voxels (uint16 stays packed until used) or, outside the volume, the ratio itself
template <int CW> struct ZZOperands {
  float f[CW];
  unsigned raw[2];
  int rm;
};

template <int K, int CW, int MINB, bool U16>
__global__ void __launch_bounds__(RLS_NT, MINB) rl_zz_kernel(const __grid_constant__ RLZZArgs A) {
  constexpr int R = (K - 1) / 2;
  const size_t plane = (size_t)A.ny * A.nx;
  const size_t i0 = ((size_t)blockIdx.x * RLS_NT + threadIdx.x) * CW;
  if (i0 >= plane) return;
  const int zo0 = blockIdx.y * A.zchunk, zo1 = min(A.nz, zo0 + A.zchunk);
  float S1[CW][K - 1], S2[CW][K - 1];
#pragma unroll
  for (int i = 0; i < CW; ++i)
#pragma unroll
    for (int m = 0; m < K - 1; ++m) { S1[i][m] = 0.f; S2[i][m] = 0.f; }
  const unsigned plane32 = (unsigned)plane;
  auto plane_off = [&](int p) { return (size_t)((unsigned long long)(unsigned)p * plane32) + i0; };
  auto load_view = [&](int rm, ZZOperands<CW>& E) {
    if constexpr (U16) {
      const uint16_t* vp = reinterpret_cast<const uint16_t*>(A.view) + plane_off(rm);
      if constexpr (CW == 4) {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(vp));
        E.raw[0] = q.x; E.raw[1] = q.y;
      } else {
        E.raw[0] = __ldg(reinterpret_cast<const unsigned*>(vp));
      }
    } else {
      vec_load<CW>(reinterpret_cast<const float*>(A.view) + plane_off(rm), E.f);
    }
  };
  auto load_operands = [&](int rm, ZZOperands<CW>& E) {
    E.rm = rm;
    if (rm < 0) vec_load<CW>(A.side + plane_off(-1 - rm), E.f);
    else load_view(rm, E);
  };
  auto prefetch_view = [&](int r) {
    prefetch_l2(reinterpret_cast<const unsigned char*>(A.view) + plane_off(r) * (U16 ? 2 : 4));
  };
  auto ratio_of = [&](const ZZOperands<CW>& E, const float (&e1)[CW], float (&q)[CW]) {
    float I[CW];
    if constexpr (U16) {
      I[0] = u16_to_float(E.raw[0] & 0xFFFFu); I[1] = u16_to_float(E.raw[0] >> 16);
      if constexpr (CW == 4) { I[2] = u16_to_float(E.raw[1] & 0xFFFFu); I[3] = u16_to_float(E.raw[1] >> 16); }
    } else {
#pragma unroll
      for (int i = 0; i < CW; ++i) I[i] = E.f[i];
    }
#pragma unroll
    for (int i = 0; i < CW; ++i) q[i] = I[i] * rcp_approx(fmaxf(e1[i], 0.f) + 1e-6f);
  };
  auto push = [&](float (&S)[CW][K - 1], const float (&in)[CW], float (&out)[CW]) {
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      out[i] = fmaf(A.kz[0], in[i], S[i][0]);
#pragma unroll
      for (int m = 0; m < K - 2; ++m) S[i][m] = fmaf(A.kz[m + 1], in[i], S[i][m + 1]);
      S[i][K - 2] = A.kz[K - 1] * in[i];
    }
  };
  // Two-deep software pipeline: with two shift registers a thread holds few columns and an SM few warps, so one step
  // of the other warps does not cover a load.  The plane of step s+2 and the operands of the ratio at t+2 are requested
  // at the top of step s (L2 prefetches run RLS_AHEAD steps ahead of that).  The three buffer sets rotate by NAME (steps
  // come in triples): a register move out of a buffer whose load is still in flight would wait for it.
  // A triple whose steps all lie in the steady state (no ramp, every plane-map cursor inside a run, ratios computed
  // rather than read from `side`) runs without a single condition - with two warps per scheduler every branch costs
  // issue time that nothing hides; the other triples (both ends of a chunk, run boundaries) take the general step.
  const int s0 = zo0 - 2 * R, s1 = zo1 + 2 * R;
  float pa[CW], pb[CW], pc[CW];
  ZZOperands<CW> Ea, Eb, Ec;
  Ea.rm = Eb.rm = Ec.rm = 0;
  Ea.raw[0] = Ea.raw[1] = Eb.raw[0] = Eb.raw[1] = Ec.raw[0] = Ec.raw[1] = 0u;
#pragma unroll
  for (int i = 0; i < CW; ++i) { Ea.f[i] = Eb.f[i] = Ec.f[i] = 0.f; pc[i] = 0.f; }
  const int tlo = zo0 - R;                        // first ratio position of this chunk
  // cursors: zmap1 at s + 2 / s + 2 + RLS_AHEAD, rmap at t + 2 / t + 2 + RLS_AHEAD (positions beyond the maps are never used)
  MapCursor cz, czp, cr, crp;
  cursor_seek(cz, A.seg1, A.nseg1, s0);
  vec_load<CW>(A.in + plane_off(cz.v), pa);
  cursor_step(cz, A.seg1, A.nseg1, s0 + 1);
  vec_load<CW>(A.in + plane_off(cz.v), pb);
  cursor_step(cz, A.seg1, A.nseg1, s0 + 2);
  cursor_seek(czp, A.seg1, A.nseg1, s0 + 2 + RLS_AHEAD);
  cursor_seek(cr, A.segr, A.nsegr, s0 - R + 2);
  cursor_seek(crp, A.segr, A.nsegr, s0 - R + 2 + RLS_AHEAD);
  ptrdiff_t oo = (ptrdiff_t)(zo0 - 4 * R) * (ptrdiff_t)plane + (ptrdiff_t)i0;   // offset of output plane z = s - 2R (used once z >= zo0)
  // step s: `cur` / `Ecur` hold plane s and the operands of the ratio at t = s - R; `nx2` / `Enx2` receive those of s + 2
  auto step = [&](int s, float (&cur)[CW], float (&nx2)[CW], ZZOperands<CW>& Ecur, ZZOperands<CW>& Enx2, auto steady_tag) {
    constexpr bool STEADY = decltype(steady_tag)::value;
    float e1[CW], q[CW], e2[CW];
    if constexpr (STEADY) {
      vec_load<CW>(A.in + plane_off(cz.v), nx2);
      load_view(cr.v, Enx2);
      if (RLS_AHEAD > 0) {
        prefetch_l2(A.in + plane_off(czp.v));
        prefetch_view(crp.v);
      }
      cz.v += cz.step; czp.v += czp.step; cr.v += cr.step; crp.v += crp.step;
      push(S1, cur, e1);
      ratio_of(Ecur, e1, q);
      push(S2, q, e2);
      vec_store<CW>(A.out + oo, e2);
    } else {
      if (s >= s1) return;
      const int t = s - R, z = s - 2 * R;
      if (s + 2 < s1) vec_load<CW>(A.in + plane_off(cz.v), nx2);
      if (t + 2 >= tlo && s + 2 < s1) load_operands(cr.v, Enx2);
      if (RLS_AHEAD > 0 && s + 2 + RLS_AHEAD < s1) {
        prefetch_l2(A.in + plane_off(czp.v));
        if (t + 2 + RLS_AHEAD >= tlo) {
          if (crp.v < 0) prefetch_l2(A.side + plane_off(-1 - crp.v));
          else prefetch_view(crp.v);
        }
      }
      cursor_step(cz, A.seg1, A.nseg1, s + 3);
      cursor_step(czp, A.seg1, A.nseg1, s + 3 + RLS_AHEAD);
      cursor_step(cr, A.segr, A.nsegr, t + 3);
      cursor_step(crp, A.segr, A.nsegr, t + 3 + RLS_AHEAD);
      push(S1, cur, e1);
      if (t >= tlo) {   // the ratio at extended position t feeds the second blur
        if (Ecur.rm < 0) {
#pragma unroll
          for (int i = 0; i < CW; ++i) q[i] = Ecur.f[i];
        } else {
          ratio_of(Ecur, e1, q);
        }
        push(S2, q, e2);
        if (z >= zo0) vec_store<CW>(A.out + oo, e2);
      }
    }
    oo += (ptrdiff_t)plane;
  };
  for (int s = s0; s < s1; s += 3) {
    // steady triple: outputs exist (z >= zo0 from the first step on), the loads / prefetches of its last step stay inside
    // the chunk, no cursor reaches a run boundary while it advances three times, and no ratio comes from `side`
    const int t = s - R;
    const bool steady = s >= zo0 + 2 * R && s + 4 + RLS_AHEAD < s1 &&
                        cz.next > s + 5 && czp.next > s + 5 + RLS_AHEAD && cr.next > t + 5 && crp.next > t + 5 + RLS_AHEAD &&
                        Ea.rm >= 0 && Eb.rm >= 0 && Ec.rm >= 0 && cr.v >= 0 && cr.step >= 0 && crp.v >= 0 && crp.step >= 0;
    if (steady) {
      step(s, pa, pc, Ea, Ec, std::true_type{});
      step(s + 1, pb, pa, Eb, Ea, std::true_type{});
      step(s + 2, pc, pb, Ec, Eb, std::true_type{});
    } else {
      step(s, pa, pc, Ea, Ec, std::false_type{});
      step(s + 1, pb, pa, Eb, Ea, std::false_type{});
      step(s + 2, pc, pb, Ec, Eb, std::false_type{});
    }
  }
}

// the vector paths need 16-byte (CW = 4) / 8-byte (CW = 2) aligned rows
static bool rl_stream_ok(const RLArgs& A, int cw, const void* extra0 = nullptr, const void* extra1 = nullptr,
                         const void* extra2 = nullptr) {
  const uintptr_t m = (uintptr_t)cw * 4 - 1;
  uintptr_t bits = reinterpret_cast<uintptr_t>(A.in) | reinterpret_cast<uintptr_t>(A.out) | reinterpret_cast<uintptr_t>(extra0) |
                   reinterpret_cast<uintptr_t>(extra1) | reinterpret_cast<uintptr_t>(extra2);
  return (A.nx % cw) == 0 && (bits & m) == 0;
}

static int rl_chunk(int n, int cols_ctas, int per_sm) {
  // ~6 waves of CTAs, at least 128 rows/planes per chunk so that the 2R ramp stays below ~15 % for 19 taps
  const int chunks = (6 * per_sm * sm_count() + cols_ctas - 1) / (cols_ctas > 0 ? cols_ctas : 1);
  int c = (n + (chunks > 0 ? chunks : 1) - 1) / (chunks > 0 ? chunks : 1);
  if (c < 128) c = 128;
  if (c > n) c = n;
  return c;
}

template <int K, int CW, int MODE>
static int launch_rl_xy_k(RLArgs A, int nplanes, cudaStream_t st) {
  const int xb = (A.nx / CW + RLS_NT - 1) / RLS_NT;
  A.zchunk = rl_chunk(A.ny, xb * nplanes, 2);
  dim3 grid(xb, (A.ny + A.zchunk - 1) / A.zchunk, nplanes);
  rl_xy_kernel<K, CW, MODE><<<grid, RLS_NT, 0, st>>>(A);
  MVF_LAUNCHED();
  constexpr int ER = (K - 1) / 2, ELEAD = (ER + CW - 1) / CW * CW;
  dim3 egrid(2, (A.ny + RLE_ROWS - 1) / RLE_ROWS, nplanes);
  const size_t esmem = (size_t)(RLE_ROWS + 2 * ER) * (2 * ELEAD + 2 * ER) * sizeof(float);
  rl_xy_edge_kernel<K, CW, MODE><<<egrid, RLE_NT, esmem, st>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

template <int K, int CW, int MINB>
static int launch_rl_zz_k(RLZZArgs A, cudaStream_t st) {
  // (view dtype as a template parameter: no dtype branch in the steady loop)
  const size_t plane = (size_t)A.ny * A.nx;
  const int pb = (int)((plane / CW + RLS_NT - 1) / RLS_NT);
  // both ramps (4R planes) are walked by every chunk: chunks of at least 256 planes
  const int chunks = (4 * MINB * sm_count() + pb - 1) / pb;
  int c = (A.nz + (chunks > 0 ? chunks : 1) - 1) / (chunks > 0 ? chunks : 1);
  if (c < 256) c = 256;
  if (c > A.nz) c = A.nz;
  A.zchunk = c;
  dim3 grid(pb, (A.nz + A.zchunk - 1) / A.zchunk);
  if (A.view_u16) rl_zz_kernel<K, CW, MINB, true><<<grid, RLS_NT, 0, st>>>(A);
  else rl_zz_kernel<K, CW, MINB, false><<<grid, RLS_NT, 0, st>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

template <int K, int MODE, int CW>
static int launch_rl_z_k(RLArgs A, cudaStream_t st) {
  const size_t plane = (size_t)A.ny * A.nx;
  const int pb = (int)((plane / CW + RLS_NT - 1) / RLS_NT);
  A.zchunk = rl_chunk(A.nz, pb, 2);
  dim3 grid(pb, (A.nz + A.zchunk - 1) / A.zchunk);
  rl_z_kernel<K, MODE, CW><<<grid, RLS_NT, 0, st>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

}  // namespace mvf
