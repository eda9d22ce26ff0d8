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
template <int K, int CW>
// (43 taps: window + shift register exceed 128 registers, one CTA per SM)
__global__ void __launch_bounds__(RLS_NT, K > 37 ? 1 : MVF_RLS_MINB) rl_xy_kernel(const __grid_constant__ RLArgs A) {
  constexpr int R = (K - 1) / 2;
  constexpr int LEAD = (R + CW - 1) / CW * CW;       // x halo in whole chunks
  constexpr int NCH = (CW + 2 * LEAD) / CW;          // chunks of the x window
  const int x0 = (blockIdx.x * RLS_NT + threadIdx.x) * CW;
  if (x0 >= A.nx) return;
  const int yo0 = blockIdx.y * A.zchunk, yo1 = min(A.ny, yo0 + A.zchunk);   // zchunk doubles as the y chunk here
  const size_t plane = (size_t)A.ny * A.nx;
  const float* __restrict__ src = A.in + (size_t)blockIdx.z * plane;
  float* __restrict__ dst = A.out + (size_t)blockIdx.z * plane;
  // Threads whose window would leave [0, nx) (the first / last LEAD columns of a row) shift it back inside: their
  // results are placeholders that rl_xy_edge_kernel overwrites afterwards with the index-mapped values.  Every
  // thread of this kernel therefore runs the same branch-free vector code.
  const int xw = min(max(x0, LEAD), A.nx - CW - LEAD);
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
    if (INTERIOR || t - R >= yo0) vec_store<CW>(dst + ooff, e);
    ooff += A.nx;
  };
  // interior steps: the output row exists (t - R >= yo0), rows t+1 and t+RLS_AHEAD+1 are real rows, t+1 is needed
  const int t_lo = max(yo0 + R, 0), t_hi = min(yo1 + R - 1, A.ny - RLS_AHEAD - 1);
  int t = yo0 - R;
  for (; t < min(t_lo, yo1 + R); ++t) step(t, std::false_type{});
  for (; t < t_hi; ++t) step(t, std::true_type{});
  for (; t < yo1 + R; ++t) step(t, std::false_type{});
}

// The 2 * LEAD columns next to the x borders of every row: thread = column, index-mapped scalar window (the K source
// columns of a thread are loop invariant and live in registers), same y shift register.  ~5 % of the main kernel.
template <int K, int CW>
__global__ void __launch_bounds__(64) rl_xy_edge_kernel(const __grid_constant__ RLArgs A) {
  constexpr int R = (K - 1) / 2;
  constexpr int LEAD = (R + CW - 1) / CW * CW;
  const int lane = threadIdx.x;   // one small CTA (32 or 64 threads >= 2 * LEAD columns) per (plane, y chunk): many resident CTAs
  const int yo0 = blockIdx.y * A.zchunk, yo1 = min(A.ny, yo0 + A.zchunk);
  if (yo0 >= A.ny || lane >= 2 * LEAD) return;
  const int x = lane < LEAD ? lane : A.nx - 2 * LEAD + lane;
  if (x < 0 || x >= A.nx) return;
  const size_t plane = (size_t)A.ny * A.nx;
  const float* __restrict__ src = A.in + (size_t)blockIdx.z * plane;
  float* __restrict__ dst = A.out + (size_t)blockIdx.z * plane;
  int col[K];
#pragma unroll
  for (int j = 0; j < K; ++j) col[j] = ext_index(x + R - j, A.nx, A.ksz_x);
  float S[K - 1];
#pragma unroll
  for (int m = 0; m < K - 1; ++m) S[m] = 0.f;
  float w[K];   // window of the current row; the next row's is requested as soon as the x pass has consumed it
  {
    const float* __restrict__ row = src + (size_t)ext_index(yo0 - R, A.ny, A.ksz_y) * A.nx;
#pragma unroll
    for (int j = 0; j < K; ++j) w[j] = __ldg(row + col[j]);
  }
  for (int t = yo0 - R; t < yo1 + R; ++t) {
    if (t + 8 < yo1 + R) prefetch_l2(src + (size_t)ext_index(t + 8, A.ny, A.ksz_y) * A.nx + x);   // the row's two border lines
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) o = fmaf(A.kx[j], w[j], o);   // same tap order as the main kernel
    if (t + 1 < yo1 + R) {
      const float* __restrict__ row = src + (size_t)ext_index(t + 1, A.ny, A.ksz_y) * A.nx;
#pragma unroll
      for (int j = 0; j < K; ++j) w[j] = __ldg(row + col[j]);
    }
    const float e = fmaf(A.ky[0], o, S[0]);
#pragma unroll
    for (int m = 0; m < K - 2; ++m) S[m] = fmaf(A.ky[m + 1], o, S[m + 1]);
    S[K - 2] = A.ky[K - 1] * o;
    const int y = t - R;
    if (y >= yo0) dst[(size_t)y * A.nx + x] = e;
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
  vec_load<CW>(in_plane(__ldg(zmap + zo0 - R)), pn);
  if (MODE != RL_PLAIN && R == 0) load_operands((size_t)zo0 * plane + i0);
  size_t o = (size_t)((long long)(zo0 - 2 * R) * (long long)plane) + i0;   // offset of output plane z = t - R (meaningful once z >= zo0)
  for (int t = zo0 - R; t < zo1 + R; ++t, o += plane) {
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
#pragma unroll
    for (int i = 0; i < CW; ++i) p[i] = pn[i];
    if (t + 1 < zo1 + R) vec_load<CW>(in_plane(__ldg(zmap + t + 1)), pn);
    float e[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      e[i] = fmaf(A.kz[0], p[i], S[i][0]);
#pragma unroll
      for (int m = 0; m < K - 2; ++m) S[i][m] = fmaf(A.kz[m + 1], p[i], S[i][m + 1]);
      S[i][K - 2] = A.kz[K - 1] * p[i];
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
  }
  if (MODE == RL_CORRECT && A.last && A.sum) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) local_sum += __shfl_xor_sync(0xffffffffu, local_sum, s);
    if ((threadIdx.x & 31) == 0) atomicAdd(A.sum, local_sum);
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

template <int K, int CW>
static int launch_rl_xy_k(RLArgs A, int nplanes, cudaStream_t st) {
  const int xb = (A.nx / CW + RLS_NT - 1) / RLS_NT;
  A.zchunk = rl_chunk(A.ny, xb * nplanes, 2);
  dim3 grid(xb, (A.ny + A.zchunk - 1) / A.zchunk, nplanes);
  rl_xy_kernel<K, CW><<<grid, RLS_NT, 0, st>>>(A);
  MVF_LAUNCHED();
  A.zchunk = A.ny < 64 ? A.ny : 64;   // short y chunks: the edge columns are few, parallelism comes from the chunks
  constexpr int ELEAD = ((K - 1) / 2 + CW - 1) / CW * CW;
  dim3 egrid(1, (A.ny + A.zchunk - 1) / A.zchunk, nplanes);
  rl_xy_edge_kernel<K, CW><<<egrid, 2 * ELEAD <= 32 ? 32 : 64, 0, st>>>(A);
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
