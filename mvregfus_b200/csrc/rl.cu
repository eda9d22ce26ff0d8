// K4/K5: multi-view Richardson-Lucy deconvolution (reference: mvregfus/multiview.py:5545-5749,
// fuse_LR_with_weights_np) for separable PSFs, sm_100a.
//
// One kernel computes a full 3-D blur-in-view  out[i] = sum_j psf[j] * ext(i + R - j)  (the closed form of the
// reference's reflect-padded circular FFT convolution) as three 1-D passes without any
// intermediate volume in HBM: a CTA owns a 32x32 (y,x) column of the volume and streams along z; per input
// plane the x pass and the y pass run through shared memory with 4-wide register windows, and the z pass is a
// K-deep shift register of partial sums held in registers (4 columns x K accumulators per thread).  The RL
// pointwise algebra is fused into the epilogue:
//   ratio  : R_v = I_v / (max(blur(est),0) + 1e-6) * (w_v > 1e-5)                         (mv:5686-5699)
//   correct: C (+)= max(blur(R_v),0) * w_v ; last view: est = clip(est * C, 0, 65535)      (mv:5706-5723)
// so one blur moves 4 B in + its epilogue operands instead of three read+write passes.  The same PSF is used
// un-flipped for the "transpose" (mv:5710).  Boundary handling in y/x is the closed-form index map; in z it is a
// caller-provided plane map so that a z-slab with NCCL-exchanged halo planes uses the very same kernel.
#include "common.cuh"
#include <stdlib.h>

namespace mvf {

#ifndef MVF_RL_XS
#define MVF_RL_XS 4
#endif
constexpr int RL_TX = 32, RL_TY = 32;
// threads own CPT consecutive y rows of one x column: CPT = 4 (256 threads) for PSFs up to 19 taps, CPT = 2
// (512 threads) beyond, so that the CPT x K shift register of partial sums still fits the register file
constexpr int RL_KMAX = MVF_RL_KMAX;

enum { RL_PLAIN = 0, RL_RATIO = 1, RL_CORRECT = 2 };

struct RLArgs {
  const float* in;      // input buffer (planes of ny*nx floats)
  const int* zmap;      // extended z position (t + Rz) -> plane of `in`; length nz + 2*Rz
  float kz[RL_KMAX], ky[RL_KMAX], kx[RL_KMAX];  // taps, zero padded (centred) to the template size
  int ksz_y, ksz_x;     // ORIGINAL kernel sizes in y/x (the wrap quirk of the index map depends on them)
  int nz, ny, nx;       // output extents
  int zchunk;           // output planes per CTA
  // epilogue operands
  const void* view;     // RATIO: I_v (uint16 or float32)
  int view_u16;
  const float* weights; // RATIO: mask = w > 1e-5 ; CORRECT: multiplier
  float* out;           // PLAIN: blur ; RATIO: R_v ; CORRECT: correction factor C
  int first, last;      // CORRECT: first view initialises C, last view applies the update
  float* est;           // CORRECT+last: estimate planes (may live in a halo-padded buffer)
  long long est_off;    // element offset of output plane 0 inside `est`
  double* sum;          // CORRECT+last: sum(est) accumulated here (convergence test, mv:5726)
  const float* dense;   // non-separable PSF: Kp^3 taps (device), centred and zero padded
  int ksz_z;            // original z size (direct kernel: index map offsets)
};

// index map of the reference's boundary handling: reflect (no edge repeat) above, wrap into
// the far reflected tail below.  Clamped so that positions never used by a valid output stay in range.
__device__ __forceinline__ int ext_index(int t, int n, int K) {
  int i = t < 0 ? n - 3 - K - t : (t >= n ? 2 * n - 2 - t : t);
  return max(0, min(i, n - 1));
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int K, int MODE, int RL_CPT>
__global__ void __launch_bounds__(32 * (RL_TY / RL_CPT), RL_CPT == 4 ? 2 : 1) rl_blur_kernel(const __grid_constant__ RLArgs A) {
  constexpr int RL_NT = 32 * (RL_TY / RL_CPT);
  constexpr int R = (K - 1) / 2;
  constexpr int LEAD = (R + 3) & ~3;            // x halo rounded up to whole 16-byte chunks
  constexpr int AY = RL_TY + 2 * R, AX = RL_TX + 2 * LEAD;
  constexpr int XS = MVF_RL_XS;                 // outputs per x-pass work item (4 or 8)
  // multiple of 4 (rows 16-byte aligned for cp.async / LDS.128); with 8-wide strips two tile rows share a quarter
  // warp, so the pitch is made = 4 (mod 8) words to keep their LDS.128 on disjoint banks
  constexpr int PA = (XS == 8 && (AX % 8) == 0) ? AX + 4 : AX;
  constexpr int PB = RL_TX + 1;
  constexpr int CPR = AX / 4;                   // 16-byte chunks per tile row
  constexpr int NV = (AY * CPR + RL_NT - 1) / RL_NT;   // vector copies per thread and plane
  constexpr int NS = (AY * AX + RL_NT - 1) / RL_NT;    // scalar copies (tiles touching the x border / unaligned rows)
  constexpr int WIN = XS + 2 * LEAD;            // x-pass register window (whole chunks)
  // dynamic shared memory: two plane tiles (plane t+1 streams in while plane t is filtered), the x-pass output,
  // the ext-mapped row offsets / columns of the tile and the plane map of this CTA's z range
  extern __shared__ __align__(16) unsigned char rl_smem[];
  float (*sA)[AY * PA] = reinterpret_cast<float (*)[AY * PA]>(rl_smem);
  float* sB = reinterpret_cast<float*>(rl_smem) + 2 * AY * PA;
  int* sRowOff = reinterpret_cast<int*>(sB + AY * PB);
  int* sCol = sRowOff + AY;
  int* sZ = sCol + AX;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * RL_TX, y0 = blockIdx.y * RL_TY;
  const int zo0 = blockIdx.z * A.zchunk, zo1 = min(A.nz, zo0 + A.zchunk);
  const int t_begin = zo0 - R, t_end = zo1 + R;
  if (tid < AY) sRowOff[tid] = ext_index(y0 - R + tid, A.ny, A.ksz_y) * A.nx;
  if (tid >= 64 && tid < 64 + AX) sCol[tid - 64] = ext_index(x0 - LEAD + (tid - 64), A.nx, A.ksz_x);
  for (int i = tid; i < t_end - t_begin; i += RL_NT) sZ[i] = A.zmap[t_begin + R + i];
  // vector path: every tile column is a real, contiguous, 16-byte aligned source column
  const bool vec = (x0 - LEAD >= 0) && (x0 + RL_TX + LEAD <= A.nx) && ((A.nx & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(A.in) & 15u) == 0);

  float S[RL_CPT][K];
#pragma unroll
  for (int c = 0; c < RL_CPT; ++c)
#pragma unroll
    for (int j = 0; j < K; ++j) S[c][j] = 0.f;

  const int x = x0 + lane;
  const size_t plane = (size_t)A.ny * A.nx;
  double local_sum = 0.0;
  __syncthreads();

  // per-thread copy descriptors, computed once: offset inside a source plane and inside a tile buffer
  int goff[NV], soff[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int i = tid + u * RL_NT;
    const int r = i / CPR, q = i - r * CPR;
    const bool ok = vec && i < AY * CPR;
    goff[u] = ok ? sRowOff[r] + (x0 - LEAD) + q * 4 : -1;
    soff[u] = r * PA + q * 4;
  }
  auto issue_plane = [&](int t, int buf) {
    const float* __restrict__ src = A.in + (size_t)sZ[t - t_begin] * plane;
    if (vec) {
#pragma unroll
      for (int u = 0; u < NV; ++u)
        if (goff[u] >= 0) cp_async16(&sA[buf][soff[u]], src + goff[u]);
    } else {
#pragma unroll 2
      for (int u = 0; u < NS; ++u) {
        const int i = tid + u * RL_NT;
        const int r = i / AX, c = i - r * AX;
        if (i < AY * AX) cp_async4(&sA[buf][r * PA + c], src + sRowOff[r] + sCol[c]);
      }
    }
    cp_async_commit();
  };

  // epilogue addressing: offset of the thread's 4 columns inside a plane (-1: outside the volume)
  int coff[RL_CPT];
#pragma unroll
  for (int c = 0; c < RL_CPT; ++c) {
    const int y = y0 + warp * RL_CPT + c;
    coff[c] = (y < A.ny && x < A.nx) ? y * A.nx + x : -1;
  }
  const float* __restrict__ wplane = A.weights ? A.weights + (size_t)zo0 * plane : nullptr;
  float* __restrict__ oplane = A.out ? A.out + (size_t)zo0 * plane : nullptr;
  float* __restrict__ eplane = A.est ? A.est + A.est_off + (size_t)zo0 * plane : nullptr;
  const unsigned char* vplane = A.view ? reinterpret_cast<const unsigned char*>(A.view) + (size_t)zo0 * plane * (A.view_u16 ? 2 : 4) : nullptr;
  const size_t vstep = plane * (A.view_u16 ? 2 : 4);

  issue_plane(t_begin, 0);
  for (int t = t_begin; t < t_end; ++t) {
    const int buf = (t - t_begin) & 1;
    if (t + 1 < t_end) { issue_plane(t + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    // epilogue operands of the plane completed by this input plane: fetched now, consumed after the filters
    const int z = t - R;
    const bool emit = z >= zo0;
    float eI[RL_CPT], eW[RL_CPT], eC[RL_CPT];
    unsigned eRaw[RL_CPT];  // uint16 view voxels stay raw until the epilogue: converting here would wait for the load
    if (MODE != RL_PLAIN && emit) {
#pragma unroll
      for (int c = 0; c < RL_CPT; ++c) {
        eI[c] = eW[c] = eC[c] = 0.f;
        eRaw[c] = 0u;
        if (coff[c] >= 0) {
          eW[c] = __ldg(wplane + coff[c]);
          if (MODE == RL_RATIO) {
            if (A.view_u16) eRaw[c] = __ldg(reinterpret_cast<const uint16_t*>(vplane) + coff[c]);
            else eI[c] = __ldg(reinterpret_cast<const float*>(vplane) + coff[c]);
          }
          if (MODE == RL_CORRECT) {
            if (!A.first) eC[c] = oplane[coff[c]];
            if (A.last) eI[c] = eplane[coff[c]];
          }
        }
      }
    }
    __syncthreads();
    // ---- x pass: strips of 4 outputs; the window is read as whole 16-byte chunks (conflict free: the 8
    // lanes of a quarter warp read 128 contiguous bytes of one tile row)
    const float* tile = sA[buf];
    for (int item = tid; item < AY * (RL_TX / XS); item += RL_NT) {
      const int r = item / (RL_TX / XS), s = item % (RL_TX / XS);
      const float4* a = reinterpret_cast<const float4*>(tile + r * PA + s * XS);
      float win[WIN];
#pragma unroll
      for (int j = 0; j < WIN / 4; ++j) {
        const float4 v = a[j];
        win[4 * j + 0] = v.x; win[4 * j + 1] = v.y; win[4 * j + 2] = v.z; win[4 * j + 3] = v.w;
      }
      float o[XS];
#pragma unroll
      for (int i = 0; i < XS; ++i) o[i] = 0.f;
#pragma unroll
      for (int j = 0; j < K; ++j) {  // out[x] = sum_j k[j] * in[x + R - j]; tile column of x0+XS*s+i is LEAD+XS*s+i
        const float kj = A.kx[j];
#pragma unroll
        for (int i = 0; i < XS; ++i) o[i] = fmaf(kj, win[LEAD + i + R - j], o[i]);
      }
      float* b = sB + r * PB + s * XS;
#pragma unroll
      for (int i = 0; i < XS; ++i) b[i] = o[i];
    }
    __syncthreads();
    // ---- y pass: thread (lane = x, warp = strip of 4 y), then the z shift register
    {
      const float* b = sB + (warp * RL_CPT) * PB + lane;
      float win[RL_CPT + 2 * R];
#pragma unroll
      for (int j = 0; j < RL_CPT + 2 * R; ++j) win[j] = b[j * PB];
      float p[RL_CPT];
#pragma unroll
      for (int c = 0; c < RL_CPT; ++c) p[c] = 0.f;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float kj = A.ky[j];
#pragma unroll
        for (int c = 0; c < RL_CPT; ++c) p[c] = fmaf(kj, win[c + 2 * R - j], p[c]);
      }
#pragma unroll
      for (int c = 0; c < RL_CPT; ++c) {
        const float e = fmaf(A.kz[0], p[c], S[c][0]);
#pragma unroll
        for (int m = 0; m < K - 1; ++m) S[c][m] = fmaf(A.kz[m + 1], p[c], S[c][m + 1]);
        S[c][K - 1] = 0.f;
        if (emit && coff[c] >= 0) {
          const float blur = fmaxf(e, 0.f);  // img_ft[img_ft<0] = 0 (mv:5671)
          if (MODE == RL_PLAIN) {
            oplane[coff[c]] = blur;
          } else if (MODE == RL_RATIO) {
            const float I = A.view_u16 ? u16_to_float(eRaw[c]) : eI[c];
            // 2-ulp fast division: the RL tolerance (1e-3 after N iterations) is five orders of magnitude looser
            const float ratio = __fdividef(I, blur + 1e-6f);
            oplane[coff[c]] = eW[c] > 1e-5f ? ratio : 0.f;
          } else {
            const float o = blur * eW[c];
            const float corr = A.first ? o : eC[c] + o;
            if (A.last) {
              const float v = fminf(fmaxf(eI[c] * corr, 0.f), 65535.f);
              eplane[coff[c]] = v;
              local_sum += (double)v;
            } else {
              oplane[coff[c]] = corr;
            }
          }
        }
      }
    }
    if (emit) {  // advance the epilogue planes
      if (wplane) wplane += plane;
      if (oplane) oplane += plane;
      if (eplane) eplane += plane;
      if (vplane) vplane += vstep;
    }
    // the y pass reads sB; sB is rewritten only after the next barrier, sA[buf] only by the plane after next
  }
  if (MODE == RL_CORRECT && A.last && A.sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_sum += __shfl_xor_sync(0xffffffffu, local_sum, o);
    if (lane == 0) atomicAdd(A.sum, local_sum);
  }
}

// Direct (non-separable) convolution for PSFs of oblique views: every thread produces 4 consecutive x outputs,
// walking the Kp^2 (z,y) tap rows with a (Kp+3)-wide register window per row.  O(K^3) per voxel: the fallback
// for PSFs that are not outer products (the separable kernel above is the fast path).
template <int K, int MODE>
__global__ void __launch_bounds__(256) rl_direct_kernel(const __grid_constant__ RLArgs A) {
  constexpr int R = (K - 1) / 2;
  const int xq = blockIdx.x * 32 + (threadIdx.x & 31);   // group of 4 x outputs
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int z = blockIdx.z;
  const int x0 = xq * 4;
  if (y >= A.ny || x0 >= A.nx) return;
  const size_t plane = (size_t)A.ny * A.nx;
  int cols[K + 3];
#pragma unroll
  for (int j = 0; j < K + 3; ++j) cols[j] = ext_index(x0 - R + j, A.nx, A.ksz_x);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int jz = 0; jz < K; ++jz) {
    const float* __restrict__ src = A.in + (size_t)A.zmap[z + R - jz + R] * plane;  // zmap[t + Rp], t = z + R - jz
    for (int jy = 0; jy < K; ++jy) {
      const float* __restrict__ row = src + (size_t)ext_index(y + R - jy, A.ny, A.ksz_y) * A.nx;
      const float* __restrict__ taps = A.dense + (jz * K + jy) * K;
      float win[K + 3];
#pragma unroll
      for (int j = 0; j < K + 3; ++j) win[j] = __ldg(row + cols[j]);
#pragma unroll
      for (int jx = 0; jx < K; ++jx) {
        const float kj = __ldg(taps + jx);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(kj, win[i + 2 * R - jx], acc[i]);
      }
    }
  }
  double local_sum = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x0 + i;
    if (x >= A.nx) break;
    const size_t idx = (size_t)z * plane + (size_t)y * A.nx + x;
    const float blur = fmaxf(acc[i], 0.f);
    if (MODE == RL_PLAIN) {
      A.out[idx] = blur;
    } else if (MODE == RL_RATIO) {
      const float I = A.view_u16 ? u16_to_float(reinterpret_cast<const uint16_t*>(A.view)[idx])
                                 : reinterpret_cast<const float*>(A.view)[idx];
      const float ratio = __fdiv_rn(I, blur + 1e-6f);
      A.out[idx] = A.weights[idx] > 1e-5f ? ratio : 0.f;
    } else {
      const float o = blur * A.weights[idx];
      const float corr = A.first ? o : A.out[idx] + o;
      if (A.last) {
        const float v = fminf(fmaxf(A.est[A.est_off + idx] * corr, 0.f), 65535.f);
        A.est[A.est_off + idx] = v;
        local_sum += (double)v;
      } else {
        A.out[idx] = corr;
      }
    }
  }
  if (MODE == RL_CORRECT && A.last && A.sum) atomicAdd(A.sum, local_sum);
}

// estimate = sum_v w_v * I_v in float32, views added in order (np.sum(weights * data, 0), mv:5629)
struct RLInitArgs {
  const void* views[MVF_MAX_VIEWS];
  const float* weights[MVF_MAX_VIEWS];
  float* est;
  double* sum;
  size_t n;
  int nviews, u16;
};

__global__ void __launch_bounds__(256) rl_init_kernel(const __grid_constant__ RLInitArgs A) {
  double local = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < A.n; i += (size_t)gridDim.x * 256) {
    float s = 0.f;
    for (int v = 0; v < A.nviews; ++v) {
      const float I = A.u16 ? u16_to_float(reinterpret_cast<const uint16_t*>(A.views[v])[i])
                            : reinterpret_cast<const float*>(A.views[v])[i];
      const float p = __fmul_rn(A.weights[v][i], I);
      s = v == 0 ? p : __fadd_rn(s, p);
    }
    A.est[i] = s;
    local += (double)s;
  }
  if (A.sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(A.sum, local);
  }
}

static int padded_size(int k) {
  const int sizes[] = {7, 13, 19, 25, 31, 37, 43};
  for (int s : sizes)
    if (k <= s) return s;
  return -1;
}

template <int MODE>
static int launch_rl_direct(int kp, const RLArgs& A, cudaStream_t st) {
  dim3 grid((A.nx + 127) / 128, (A.ny + 7) / 8, A.nz);
  switch (kp) {
    case 7: rl_direct_kernel<7, MODE><<<grid, 256, 0, st>>>(A); break;
    case 13: rl_direct_kernel<13, MODE><<<grid, 256, 0, st>>>(A); break;
    case 19: rl_direct_kernel<19, MODE><<<grid, 256, 0, st>>>(A); break;
    case 25: rl_direct_kernel<25, MODE><<<grid, 256, 0, st>>>(A); break;
    case 31: rl_direct_kernel<31, MODE><<<grid, 256, 0, st>>>(A); break;
    case 37: rl_direct_kernel<37, MODE><<<grid, 256, 0, st>>>(A); break;
    case 43: rl_direct_kernel<43, MODE><<<grid, 256, 0, st>>>(A); break;
    default: return fail(MVF_ERR_UNSUPPORTED, "%s%s", "PSF longer than 43 taps");
  }
  MVF_LAUNCHED();
  return MVF_OK;
}

template <int K, int MODE, int CPT>
static int launch_rl_k(const RLArgs& A, cudaStream_t st) {
  constexpr int R = (K - 1) / 2, LEAD = (R + 3) & ~3;
  constexpr int AY = RL_TY + 2 * R, AX = RL_TX + 2 * LEAD;
  dim3 grid((A.nx + RL_TX - 1) / RL_TX, (A.ny + RL_TY - 1) / RL_TY, (A.nz + A.zchunk - 1) / A.zchunk);
  constexpr int PA = (MVF_RL_XS == 8 && (AX % 8) == 0) ? AX + 4 : AX;
  const size_t smem = (size_t)(2 * AY * PA + AY * (RL_TX + 1)) * sizeof(float) + (size_t)(AY + AX + A.zchunk + K - 1) * sizeof(int);
  auto kern = rl_blur_kernel<K, MODE, CPT>;
  MVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 32 * (RL_TY / CPT), smem, st>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

}  // namespace mvf
#include "rl_stream.cuh"
namespace mvf {

template <int MODE>
static int launch_rl(int kp, const RLArgs& A, cudaStream_t st) {
  if (A.dense) return launch_rl_direct<MODE>(kp, A, st);
  switch (kp) {
    case 7: return launch_rl_k<7, MODE, 4>(A, st);
    case 13: return launch_rl_k<13, MODE, 4>(A, st);
    case 19: return launch_rl_k<19, MODE, 4>(A, st);
    case 25: return launch_rl_k<25, MODE, 2>(A, st);
    case 31: return launch_rl_k<31, MODE, 2>(A, st);
    case 37: return launch_rl_k<37, MODE, 2>(A, st);
    case 43: return launch_rl_k<43, MODE, 2>(A, st);
    default: return fail(MVF_ERR_UNSUPPORTED, "%s%s", "separable PSF longer than 43 taps");
  }
}

// streaming two-kernel formulation (rl_stream.cuh): CW = 4 columns per thread up to 19 taps, 2 beyond
template <int MODE>
static int launch_rl_xy(int kp, const RLArgs& A, int nplanes, cudaStream_t st) {
  switch (kp) {
    case 7: return launch_rl_xy_k<7, 4, MODE>(A, nplanes, st);
    case 13: return launch_rl_xy_k<13, 4, MODE>(A, nplanes, st);
    case 19: return launch_rl_xy_k<19, 4, MODE>(A, nplanes, st);
    case 25: return launch_rl_xy_k<25, 2, MODE>(A, nplanes, st);
    case 31: return launch_rl_xy_k<31, 2, MODE>(A, nplanes, st);
    case 37: return launch_rl_xy_k<37, 2, MODE>(A, nplanes, st);
    case 43: return launch_rl_xy_k<43, 2, MODE>(A, nplanes, st);
    default: return fail(MVF_ERR_UNSUPPORTED, "%s%s", "separable PSF longer than 43 taps");
  }
}

// z pass + ratio + z pass (rl_zz_kernel): two shift registers per column.  Up to 19 taps: 4 columns per thread and one
// 256-thread CTA per SM (MVF_RL_ZZ_CW=2: 2 columns, two CTAs); beyond: 2 columns per thread, one CTA per SM.
static int zz_cw19() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MVF_RL_ZZ_CW");
    v = (e && atoi(e) == 2) ? 2 : 4;
  }
  return v;
}
static int launch_rl_zz(int kp, const RLZZArgs& A, cudaStream_t st) {
  const bool wide = zz_cw19() == 4 && A.nx % 4 == 0;
  switch (kp) {
    case 7: return wide ? launch_rl_zz_k<7, 4, 2>(A, st) : launch_rl_zz_k<7, 2, 2>(A, st);
    case 13: return wide ? launch_rl_zz_k<13, 4, 1>(A, st) : launch_rl_zz_k<13, 2, 2>(A, st);
    case 19: return wide ? launch_rl_zz_k<19, 4, 1>(A, st) : launch_rl_zz_k<19, 2, 2>(A, st);
    case 25: return launch_rl_zz_k<25, 2, 1>(A, st);
    case 31: return launch_rl_zz_k<31, 2, 1>(A, st);
    case 37: return launch_rl_zz_k<37, 2, 1>(A, st);
    case 43: return launch_rl_zz_k<43, 2, 1>(A, st);
    default: return fail(MVF_ERR_UNSUPPORTED, "%s%s", "separable PSF longer than 43 taps");
  }
}

template <int MODE>
static int launch_rl_z(int kp, const RLArgs& A, cudaStream_t st) {
  switch (kp) {
    case 7: return launch_rl_z_k<7, MODE, 4>(A, st);
    case 13: return launch_rl_z_k<13, MODE, 4>(A, st);
    case 19: return launch_rl_z_k<19, MODE, 4>(A, st);
    case 25: return launch_rl_z_k<25, MODE, 2>(A, st);
    case 31: return launch_rl_z_k<31, MODE, 2>(A, st);
    case 37: return launch_rl_z_k<37, MODE, 2>(A, st);
    case 43: return launch_rl_z_k<43, MODE, 2>(A, st);
    default: return fail(MVF_ERR_UNSUPPORTED, "%s%s", "separable PSF longer than 43 taps");
  }
}

static int fill_rl(RLArgs& A, const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf, int* kp) {
  memset(&A, 0, sizeof(A));
  if (!in || !zmap || !dims || !psf) return fail(MVF_ERR_INVALID, "%s%s", "rl: NULL argument");
  int kmax = 0;
  for (int d = 0; d < 3; ++d) {
    if (psf->ksize[d] < 1 || psf->ksize[d] > RL_KMAX || !(psf->ksize[d] & 1))
      return fail(MVF_ERR_INVALID, "%s%s", "rl: PSF sizes must be odd and <= 43");
    if (d > 0 && dims[d] < psf->ksize[d] + 3)   // (z runs through the caller's plane map: slabs / plane ranges may be thinner)
      return fail(MVF_ERR_UNSUPPORTED, "%s%s", "rl: volume extent smaller than PSF size + 3 (the reference's reflect padding needs it)");
    kmax = psf->ksize[d] > kmax ? psf->ksize[d] : kmax;
  }
  *kp = padded_size(kmax);
  if (*kp < 0) return fail(MVF_ERR_UNSUPPORTED, "%s%s", "rl: PSF too large");
  const float* taps[3] = {psf->kz, psf->ky, psf->kx};
  float* dst[3] = {A.kz, A.ky, A.kx};
  for (int d = 0; d < 3; ++d) {
    const int pad = (*kp - psf->ksize[d]) / 2;
    for (int j = 0; j < psf->ksize[d]; ++j) dst[d][pad + j] = taps[d][j];
  }
  A.in = in; A.zmap = zmap;
  A.dense = psf->dense;
  A.ksz_z = psf->ksize[0]; A.ksz_y = psf->ksize[1]; A.ksz_x = psf->ksize[2];
  A.nz = dims[0]; A.ny = dims[1]; A.nx = dims[2];
  // z chunk: enough CTAs to fill the GPU twice, at least 32 planes so the 2R-plane ramp-up stays small
  const int cols = ((A.nx + RL_TX - 1) / RL_TX) * ((A.ny + RL_TY - 1) / RL_TY);
  int chunks = (6 * 2 * sm_count() + cols - 1) / cols;  // ~6 waves of 2 CTAs/SM: tail < 10 %
  int zc = (A.nz + chunks - 1) / chunks;
  if (zc < 64) zc = 64;                                // keeps the 2R-plane ramp-up of every chunk small
  if (zc > 2048) zc = 2048;
  if (zc > A.nz) zc = A.nz;
  A.zchunk = zc;
  return MVF_OK;
}

}  // namespace mvf

using namespace mvf;

extern "C" int mvf_rl_padded_size(int k) { return padded_size(k); }

extern "C" int mvf_rl_blur(const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                           float* out, void* stream) {
  RLArgs A;
  int kp;
  int rc = fill_rl(A, in, zmap, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(out, "NULL out");
  A.out = out;
  return launch_rl<RL_PLAIN>(kp, A, as_stream(stream));
}

extern "C" int mvf_rl_ratio(const float* est, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                            const void* view, int dtype, const float* weights, float* ratio_out, void* stream) {
  RLArgs A;
  int kp;
  int rc = fill_rl(A, est, zmap, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(view && weights && ratio_out, "NULL argument");
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  A.view = view; A.view_u16 = dtype == MVF_U16; A.weights = weights; A.out = ratio_out;
  return launch_rl<RL_RATIO>(kp, A, as_stream(stream));
}

extern "C" int mvf_rl_correct(const float* ratio, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                              const float* weights, float* corr, int first, int last, float* est,
                              int64_t est_offset, double* sum_out, void* stream) {
  RLArgs A;
  int kp;
  int rc = fill_rl(A, ratio, zmap, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(weights && corr, "NULL argument");
  MVF_CHECK_ARG(!last || est, "est must be given for the last view");
  A.weights = weights; A.out = corr; A.first = first; A.last = last; A.est = est; A.est_off = est_offset; A.sum = sum_out;
  return launch_rl<RL_CORRECT>(kp, A, as_stream(stream));
}

extern "C" int mvf_rl_init_estimate(int nviews, const void* const* host_view_ptrs, int dtype,
                                    const float* const* host_weight_ptrs, float* est, size_t nvoxels,
                                    double* sum_out, void* stream) {
  MVF_CHECK_ARG(nviews >= 1 && nviews <= MVF_MAX_VIEWS, "nviews must be in [1, MVF_MAX_VIEWS]");
  MVF_CHECK_ARG(host_view_ptrs && host_weight_ptrs && est, "NULL argument");
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  RLInitArgs A;
  memset(&A, 0, sizeof(A));
  for (int v = 0; v < nviews; ++v) {
    MVF_CHECK_ARG(host_view_ptrs[v] && host_weight_ptrs[v], "NULL view/weight pointer");
    A.views[v] = host_view_ptrs[v];
    A.weights[v] = host_weight_ptrs[v];
  }
  A.est = est; A.sum = sum_out; A.n = nvoxels; A.nviews = nviews; A.u16 = dtype == MVF_U16;
  if (nvoxels == 0) return MVF_OK;
  size_t blocks = (nvoxels + 255) / 256, cap = (size_t)sm_count() * 16;
  rl_init_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

// ---- streaming formulation: x/y passes of all planes, then z pass + epilogue ---------------------------------
extern "C" int mvf_rl_stream_supported(const int32_t dims[3], const mvf_rl_psf* psf) {
  if (!dims || !psf || psf->dense) return 0;
  int kmax = 0;
  for (int d = 0; d < 3; ++d) kmax = psf->ksize[d] > kmax ? psf->ksize[d] : kmax;
  const int kp = padded_size(kmax);
  if (kp < 0) return 0;
  const int cw = kp <= 19 ? 4 : 2, lead = ((kp - 1) / 2 + cw - 1) / cw * cw;
  return dims[2] % cw == 0 && dims[2] >= 2 * lead + cw;
}

extern "C" int mvf_rl_xy(const float* in, int32_t nplanes, const int32_t dims[3], const mvf_rl_psf* psf, float* out,
                         void* stream) {
  RLArgs A;
  int kp;
  const int32_t zm_dummy = 0;
  int rc = fill_rl(A, in, &zm_dummy, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(out && nplanes > 0 && nplanes <= 65535, "NULL out or bad plane count");
  MVF_CHECK_ARG(!psf->dense, "the streaming kernels need a separable PSF");
  A.out = out;
  MVF_CHECK_ARG(rl_stream_ok(A, kp <= 19 ? 4 : 2), "rows must be 16-byte aligned (nx % 4 == 0; nx % 2 beyond 19 taps)");
  return launch_rl_xy<RL_PLAIN>(kp, A, nplanes, as_stream(stream));
}

extern "C" int mvf_rl_blur_z(const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                             float* out, void* stream) {
  RLArgs A;
  int kp;
  int rc = fill_rl(A, in, zmap, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(out, "NULL out");
  A.out = out;
  MVF_CHECK_ARG(rl_stream_ok(A, kp <= 19 ? 4 : 2), "buffers must be 16-byte aligned and nx a multiple of the vector width");
  return launch_rl_z<RL_PLAIN>(kp, A, as_stream(stream));
}

extern "C" int mvf_rl_ratio_z(const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                              const void* view, int dtype, const float* weights, float* ratio_out, void* stream) {
  RLArgs A;
  int kp;
  int rc = fill_rl(A, in, zmap, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(view && ratio_out, "NULL argument");   // weights may be NULL: view already zeroed where w <= 1e-5
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  A.view = view; A.view_u16 = dtype == MVF_U16; A.weights = weights; A.out = ratio_out;
  // a uint16 view moves as 2 * CW bytes per thread: checking it with a doubled address applies the halved requirement
  const void* view_chk = dtype == MVF_U16 ? reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(view) * 2) : view;
  MVF_CHECK_ARG(rl_stream_ok(A, kp <= 19 ? 4 : 2, view_chk, weights), "buffers must be 16-byte aligned and nx a multiple of the vector width");
  return launch_rl_z<RL_RATIO>(kp, A, as_stream(stream));
}

extern "C" int mvf_rl_correct_z(const float* in, const int32_t* zmap, const int32_t dims[3], const mvf_rl_psf* psf,
                                const float* weights, float* corr, int first, int last, float* est,
                                int64_t est_offset, double* sum_out, void* stream) {
  RLArgs A;
  int kp;
  int rc = fill_rl(A, in, zmap, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(weights && corr, "NULL argument");
  MVF_CHECK_ARG(!last || est, "est must be given for the last view");
  A.weights = weights; A.out = corr; A.first = first; A.last = last; A.est = est; A.est_off = est_offset; A.sum = sum_out;
  MVF_CHECK_ARG(rl_stream_ok(A, kp <= 19 ? 4 : 2, weights, est ? est + est_offset : nullptr),
                "buffers must be 16-byte aligned and nx a multiple of the vector width");
  return launch_rl_z<RL_CORRECT>(kp, A, as_stream(stream));
}

// ---- second blur with the z pass first: T1 -> (ratio) -> T2, then x/y passes + correction epilogue -----------------
// runs of constant step of a host plane map (first position `first`): false if it needs more than ZZ_MAXSEG segments
static bool compress_map(const int32_t* map, int n, int first, MapSeg* segs, int* nseg) {
  int k = 0;
  for (int i = 0; i < n;) {
    if (k == ZZ_MAXSEG) return false;
    int j = i + 1;
    const int step = j < n ? map[j] - map[i] : 0;
    while (j < n && map[j] - map[j - 1] == step) ++j;
    segs[k].start = first + i; segs[k].base = map[i]; segs[k].step = step;
    ++k;
    i = j;
  }
  *nseg = k;
  return true;
}

extern "C" int mvf_rl_zz(const float* t1, const int32_t* host_zmap1, const int32_t* host_rmap, const int32_t dims[3],
                         const mvf_rl_psf* psf, const void* view, int dtype, const float* side, float* t2_out,
                         void* stream) {
  RLArgs B;
  int kp;
  const int32_t zm_dummy = 0;
  int rc = fill_rl(B, t1, &zm_dummy, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(host_zmap1 && host_rmap && view && t2_out, "NULL argument");
  MVF_CHECK_ARG(dtype == MVF_U16 || dtype == MVF_F32, "dtype must be MVF_U16 or MVF_F32");
  MVF_CHECK_ARG(!psf->dense, "the streaming kernels need a separable PSF");
  B.out = t2_out;
  const void* view_chk = dtype == MVF_U16 ? reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(view) * 2) : view;
  MVF_CHECK_ARG(rl_stream_ok(B, kp <= 19 ? 4 : 2, view_chk, side), "buffers must be 16-byte aligned and nx a multiple of the vector width");
  RLZZArgs A;
  memset(&A, 0, sizeof(A));
  const int rp = (kp - 1) / 2;
  if (!compress_map(host_zmap1, dims[0] + 4 * rp, -2 * rp, A.seg1, &A.nseg1) ||
      !compress_map(host_rmap, dims[0] + 2 * rp, -rp, A.segr, &A.nsegr))
    return fail(MVF_ERR_UNSUPPORTED, "%s%s", "rl_zz: plane map with more than 8 affine runs");
  A.in = t1;
  memcpy(A.kz, B.kz, sizeof(A.kz));
  A.nz = dims[0]; A.ny = dims[1]; A.nx = dims[2];
  A.view = view; A.view_u16 = dtype == MVF_U16; A.side = side; A.out = t2_out;
  return launch_rl_zz(kp, A, as_stream(stream));
}

extern "C" int mvf_rl_xy_correct(const float* t2, int32_t nplanes, const int32_t dims[3], const mvf_rl_psf* psf,
                                 const float* weights, float* corr, int first, int last, float* est,
                                 int64_t est_offset, double* sum_out, void* stream) {
  RLArgs A;
  int kp;
  const int32_t zm_dummy = 0;
  int rc = fill_rl(A, t2, &zm_dummy, dims, psf, &kp);
  if (rc) return rc;
  MVF_CHECK_ARG(weights && corr && nplanes > 0 && nplanes <= 65535, "NULL argument or bad plane count");
  MVF_CHECK_ARG(!last || est, "est must be given for the last view");
  MVF_CHECK_ARG(!psf->dense, "the streaming kernels need a separable PSF");
  A.weights = weights; A.out = corr; A.first = first; A.last = last; A.est = est; A.est_off = est_offset; A.sum = sum_out;
  MVF_CHECK_ARG(rl_stream_ok(A, kp <= 19 ? 4 : 2, weights, est ? est + est_offset : nullptr),
                "buffers must be 16-byte aligned and nx a multiple of the vector width");
  return launch_rl_xy<RL_CORRECT>(kp, A, nplanes, as_stream(stream));
}
