// K4/K5, second generation: the separable RL blur with TWO z planes per step on packed FP32 (FFMA2, sm_100a).
//
// Same mathematics, index maps and epilogues as rl_blur_kernel (rl.cu); what changes is the data path:
//   * a step consumes the input planes (t, t+1) together.  The tile in shared memory is plane-interleaved
//     (float2 = {plane t, plane t+1} per (y,x)), so the x pass and the y pass work on naturally aligned register
//     pairs and every multiply-add is one FFMA2 (fma.rn.f32x2) with the tap broadcast from a uniform register
//     pair: half the issue slots per FMA of the scalar kernel.
//   * the z pass keeps K+1 partial sums per column as (K+1)/2 register pairs over ADJACENT OUTPUT PLANES; the two
//     new planes update each pair with two FFMA2 (tap pairs {kz[m],kz[m+1]} / {kz[m-1],kz[m]}), two outputs
//     leave per step and the shift by two is a pair-aligned register renaming.  Every output still receives its
//     K contributions in ascending plane order, so z-slabs / chunks reproduce the whole volume bit for bit.
//   * quarter warps of the x pass read 8 different tile rows (pitch = 1 mod 8 sixteen-byte units): LDS.128
//     without bank conflicts; x strips of XS outputs, y strips of CPT rows per thread.
//   * the next plane pair is fetched with LDG.128 into registers while the x pass runs and interleaved into the
//     tile after it (cp.async cannot interleave); two barriers per two planes.
#pragma once

namespace mvf {

// FFMA2 runs at full rate (2 clk per warp instruction and SM sub-partition) only with all operands in vector
// registers; with a uniform-register / constant-bank operand it takes 4 clk (tools/ubench_ffma2.cu,
// profiles/r01/ubench_ffma2_b200.txt).  The taps therefore live in shared memory and are pulled into registers
// (LDS.128, broadcast) at the start of each pass; the scalar-broadcast operand form (R.F32) needs one register per tap.
__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }

template <int K, int MODE, int TY, int CPT, int XS>
__global__ void __launch_bounds__(32 * (TY / CPT), (32 * (TY / CPT) <= 256) ? 2 : 1)
rl_pair_kernel(const __grid_constant__ RLArgs A) {
  constexpr int NT = 32 * (TY / CPT), NW = NT / 32;
  constexpr int R = (K - 1) / 2;
  constexpr int KT = (K + 3) & ~3;               // taps per axis in shared memory (whole LDS.128)
  static_assert(CPT % 2 == 0, "the z pass pairs adjacent rows");
  constexpr int LEAD = (R + 3) & ~3;
  constexpr int AY = TY + 2 * R, AX = RL_TX + 2 * LEAD;
  constexpr int UNITS = (AX / 2) | 1;             // 16-byte units per tile row, odd: 8 rows hit 8 different banks groups
  constexpr int PI = 2 * UNITS;                  // tile row pitch in float2
  constexpr int PB = RL_TX + 1;                  // x-pass output pitch in float2
  constexpr int CPR = AX / 4;                    // 4-column chunks per tile row
  constexpr int NV = (AY * CPR + NT - 1) / NT;   // chunks per thread
  constexpr int WIN = XS + 2 * LEAD;             // x-pass window (columns)
  constexpr int RPW = (AY + NW - 1) / NW;        // tile rows per warp in the x pass (quarter warp = one strip)
  static_assert(XS == 8 && RPW <= 8, "x pass: 4 strips of 8 outputs, at most 8 rows per warp");
  static_assert(TY % CPT == 0, "TY/CPT");

  extern __shared__ __align__(16) unsigned char rl_smem[];
  float2* sI = reinterpret_cast<float2*>(rl_smem);           // 2 x (AY x PI): tile of step t lives in buffer t & 1
  float2* sB = sI + 2 * AY * PI;                             // 2 x (AY x PB): x-pass output of step t in buffer t & 1
  float* sTap = reinterpret_cast<float*>(sB + 2 * AY * PB);  // kx | ky | kz, KT floats each (16-byte aligned: AY even)
  int* sCol = reinterpret_cast<int*>(sTap + 3 * KT);         // 16-byte aligned (AX multiple of 8)
  int* sRowOff = sCol + AX;
  int* sZ = sRowOff + AY;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * RL_TX, y0 = blockIdx.y * TY;
  const int zo0 = blockIdx.z * A.zchunk, zo1 = min(A.nz, zo0 + A.zchunk);
  const int t_begin = zo0 - R;
  const int nplanes = zo1 - zo0 + 2 * R;          // input planes of this chunk
  const int nsteps = (nplanes + 1) >> 1;
  for (int i = tid; i < 3 * KT; i += NT) {
    const int d = i / KT, j = i - d * KT;
    sTap[i] = j < K ? (d == 0 ? A.kx[j] : d == 1 ? A.ky[j] : A.kz[j]) : 0.f;
  }
  for (int i = tid; i < AY; i += NT) sRowOff[i] = ext_index(y0 - R + i, A.ny, A.ksz_y) * A.nx;
  for (int i = tid; i < AX; i += NT) sCol[i] = ext_index(x0 - LEAD + i, A.nx, A.ksz_x);
  for (int i = tid; i < 2 * nsteps; i += NT) sZ[i] = A.zmap[t_begin + R + min(i, nplanes - 1)];
  // chunk = 4 consecutive tile columns of one row.  Interior tiles: every chunk is a real, contiguous, 16-byte
  // aligned piece of a source row (two LDG.128 per plane pair).  Tiles touching the x border (or unaligned rows):
  // chunks whose columns go through the index map are fetched as 4 + 4 scalars into the same registers.
  const bool aligned = ((A.nx & 3) == 0) && ((reinterpret_cast<uintptr_t>(A.in) & 15u) == 0);
  const bool interior = aligned && (x0 - LEAD >= 0) && (x0 + RL_TX + LEAD <= A.nx);
  const size_t plane = (size_t)A.ny * A.nx;
  __syncthreads();

  // per-thread chunk descriptors; threads beyond the last chunk repeat it (same data to the same place)
  int goff[NV], soff[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int i = min(tid + u * NT, AY * CPR - 1);
    const int r = i / CPR, q = i - r * CPR;
    const int xs = x0 - LEAD + q * 4;
    const bool direct = aligned && xs >= 0 && xs + 4 <= A.nx;
    goff[u] = direct ? sRowOff[r] + xs : -1 - i;   // < 0: mapped chunk, index encoded
    soff[u] = r * PI + q * 4;
  }
  // Chunks [0, NE) of a thread ("early") are fetched at the top of an iteration and parked after the x pass; the
  // remaining ones ("late", only the first warps have one) reuse the same registers: fetched after the parking,
  // they land during the y/z passes and are parked at the end of the iteration.  Peak register pressure (x pass)
  // thus holds NE chunk pairs instead of NV.
  constexpr int NE = NV < 2 ? NV : 2, NL = NV - NE;
  static_assert(NL <= NE, "late chunks reuse the early registers");
  const bool late = NL > 0 && warp * 32 + NE * NT < AY * CPR;   // warp-uniform
  float4 pa[NE], pb[NE];
  auto fetch = [&](int step, int u0, int n) {
    const float* __restrict__ s0 = A.in + (size_t)sZ[2 * step] * plane;
    const float* __restrict__ s1 = A.in + (size_t)sZ[2 * step + 1] * plane;
    if (interior) {
#pragma unroll
      for (int u = 0; u < NE; ++u)
        if (u < n) {
          pa[u] = __ldg(reinterpret_cast<const float4*>(s0 + goff[u0 + u]));
          pb[u] = __ldg(reinterpret_cast<const float4*>(s1 + goff[u0 + u]));
        }
    } else {
#pragma unroll
      for (int u = 0; u < NE; ++u)
        if (u < n) {
          const int g = goff[u0 + u];
          if (g >= 0) {
            pa[u] = __ldg(reinterpret_cast<const float4*>(s0 + g));
            pb[u] = __ldg(reinterpret_cast<const float4*>(s1 + g));
          } else {
            const int i = -1 - g;
            const int r = i / CPR, q = i - r * CPR;
            const int ro = sRowOff[r];
            const int4 cc = *reinterpret_cast<const int4*>(sCol + 4 * q);
            pa[u] = make_float4(__ldg(s0 + ro + cc.x), __ldg(s0 + ro + cc.y), __ldg(s0 + ro + cc.z), __ldg(s0 + ro + cc.w));
            pb[u] = make_float4(__ldg(s1 + ro + cc.x), __ldg(s1 + ro + cc.y), __ldg(s1 + ro + cc.z), __ldg(s1 + ro + cc.w));
          }
        }
    }
  };
  auto stash = [&](int buf, int u0, int n) {
#pragma unroll
    for (int u = 0; u < NE; ++u)
      if (u < n) {
        float4* d = reinterpret_cast<float4*>(sI + buf * (AY * PI) + soff[u0 + u]);
        d[0] = make_float4(pa[u].x, pb[u].x, pa[u].y, pb[u].y);
        d[1] = make_float4(pa[u].z, pb[u].z, pa[u].w, pb[u].w);
      }
  };

  // z pass state: K-1 partial sums per column, two adjacent rows (c, c+1) packed per register pair
  float2 S[CPT / 2][K - 1];
#pragma unroll
  for (int c = 0; c < CPT / 2; ++c)
#pragma unroll
    for (int j = 0; j < K - 1; ++j) S[c][j] = make_float2(0.f, 0.f);
  auto load_taps = [&](float (&k)[KT], int axis) {
#pragma unroll
    for (int u = 0; u < KT / 4; ++u) {
      const float4 t = reinterpret_cast<const float4*>(sTap + axis * KT)[u];
      k[4 * u] = t.x; k[4 * u + 1] = t.y; k[4 * u + 2] = t.z; k[4 * u + 3] = t.w;
    }
  };

  const int x = x0 + lane;
  int coff[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int y = y0 + warp * CPT + c;
    coff[c] = (y < A.ny && x < A.nx) ? y * A.nx + x : -1;
  }
  // epilogue addressing: CTA-uniform base pointers of the chunk (uniform registers) + a 32-bit voxel offset
  // (the launcher keeps zchunk * plane below 2^31)
  const size_t zbase = (size_t)zo0 * plane;
  const float* __restrict__ wbase = A.weights ? A.weights + zbase : nullptr;
  float* __restrict__ obase = A.out ? A.out + zbase : nullptr;
  float* __restrict__ ebase = A.est ? A.est + A.est_off + zbase : nullptr;
  const uint16_t* __restrict__ v16 = reinterpret_cast<const uint16_t*>(A.view) + zbase;
  const unsigned* __restrict__ v32 = reinterpret_cast<const unsigned*>(A.view) + zbase;
  const unsigned uplane = (unsigned)plane;
  double local_sum = 0.0;

  fetch(0, 0, NE);
  stash(0, 0, NE);
  if (late) { fetch(0, NE, NL); stash(0, NE, NL); }

  // Software pipeline, ONE barrier per step: in iteration `step` every warp first runs its share of the x pass of
  // step+1 (tile buffer (step+1)&1 -> sB buffer (step+1)&1), then the y/z passes of `step` from sB buffer step&1,
  // and finally parks the tile of step+2 (fetched into registers at the top) in tile buffer step&1.  Warps drift
  // between the shared-memory-heavy and the FMA-heavy phases instead of meeting at a barrier in between.
  for (int step = -1; step < nsteps; ++step) {
    __syncthreads();
    const bool park = step + 2 < nsteps;
    if (park) fetch(step + 2, 0, NE);
    if (step + 1 < nsteps) {
      // ---- x pass of step+1: quarter warp q = strip of 8 outputs, lanes of a quarter = RPW tile rows
      const int r = warp * RPW + (lane & 7);
      const int q = lane >> 3;
      if ((lane & 7) < RPW && r < AY) {
        const float4* a = reinterpret_cast<const float4*>(sI + ((step + 1) & 1) * (AY * PI) + r * PI + XS * q);
        float kx[KT];
        load_taps(kx, 0);
        float2 o[XS];
#pragma unroll
        for (int i = 0; i < XS; ++i) o[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < WIN / 2; ++u) {
          if (2 * u + 1 < LEAD - R || 2 * u > LEAD + R + XS - 1) continue;   // columns no output uses
          const float4 v = a[u];
          const float2 c0 = make_float2(v.x, v.y), c1 = make_float2(v.z, v.w);
#pragma unroll
          for (int i = 0; i < XS; ++i) {   // out[x] = sum_j k[j] in[x + R - j]; output i sits at window column LEAD + i
            const int j0 = LEAD + i + R - 2 * u, j1 = j0 - 1;
            if (j0 >= 0 && j0 < K) o[i] = __ffma2_rn(bcast2(kx[j0]), c0, o[i]);
            if (j1 >= 0 && j1 < K) o[i] = __ffma2_rn(bcast2(kx[j1]), c1, o[i]);
          }
        }
        float2* b = sB + ((step + 1) & 1) * (AY * PB) + r * PB + XS * q;
#pragma unroll
        for (int i = 0; i < XS; ++i) b[i] = o[i];
      }
    }
    // tile buffer step&1 was last read by the x pass of the previous iteration: park the fetched tile now, so its
    // registers are free during the y/z passes
    if (park) {
      stash(step & 1, 0, NE);
      if (late) fetch(step + 2, NE, NL);
    }
    if (step < 0) {
      if (park && late) stash(1, NE, NL);
      continue;
    }   // prologue iteration: x pass of step 0, tile of step 1
    // ---- epilogue operands of the two planes this step completes
    const int z = zo0 - 2 * R + 2 * step;          // first completed output plane
    const bool emit = z >= zo0;
    const bool second = z + 1 < zo1;
    const unsigned eo = (unsigned)(z - zo0) * uplane;   // offset of output plane z inside the chunk
    float eW[CPT][2], eC[CPT][2];
    unsigned eV[CPT][2];   // RATIO: raw view voxel (uint16 value or float32 bits; converted in the epilogue); CORRECT+last: estimate bits
    if (MODE != RL_PLAIN && emit) {
#pragma unroll
      for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          eW[c][h] = eC[c][h] = 0.f;
          eV[c][h] = 0u;
          if (coff[c] >= 0 && (h == 0 || second)) {
            const unsigned o = eo + (unsigned)coff[c] + (h ? uplane : 0u);
            eW[c][h] = __ldg(wbase + o);
            if (MODE == RL_RATIO) {
              if (A.view_u16) eV[c][h] = __ldg(v16 + o);
              else eV[c][h] = __ldg(v32 + o);
            }
            if (MODE == RL_CORRECT) {
              if (!A.first) eC[c][h] = obase[o];
              if (A.last) eV[c][h] = __float_as_uint(ebase[o]);
            }
          }
        }
    }
    // ---- y pass (lane = x, warp = strip of CPT rows), then the z pass on register pairs
    {
      const float2* b = sB + (step & 1) * (AY * PB) + (warp * CPT) * PB + lane;
      float2 p[CPT];   // {plane z, plane z+1} of the thread's CPT rows after the y pass
      {
        float ky[KT];
        load_taps(ky, 1);
#pragma unroll
        for (int c = 0; c < CPT; ++c) p[c] = make_float2(0.f, 0.f);
#pragma unroll
        for (int jj = 0; jj < CPT + 2 * R; ++jj) {
          const float2 v = b[jj * PB];
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const int j = c + 2 * R - jj;
            if (j >= 0 && j < K) p[c] = __ffma2_rn(bcast2(ky[j]), v, p[c]);
          }
        }
      }
      float kz[KT];
      load_taps(kz, 2);
#pragma unroll
      for (int cp = 0; cp < CPT / 2; ++cp) {
        // repack to row pairs per plane, then one shift-register update per plane (ascending plane order)
        const float2 q0 = make_float2(p[2 * cp].x, p[2 * cp + 1].x), q1 = make_float2(p[2 * cp].y, p[2 * cp + 1].y);
        float2 e[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 qv = h ? q1 : q0;
          e[h] = __ffma2_rn(bcast2(kz[0]), qv, S[cp][0]);
#pragma unroll
          for (int m = 0; m < K - 2; ++m) S[cp][m] = __ffma2_rn(bcast2(kz[m + 1]), qv, S[cp][m + 1]);
          S[cp][K - 2] = __fmul2_rn(bcast2(kz[K - 1]), qv);
        }
        if (emit) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * cp + cc;
            if (coff[c] < 0) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 1 && !second) break;
              const unsigned o = eo + (unsigned)coff[c] + (h ? uplane : 0u);
              const float blur = fmaxf(cc ? e[h].y : e[h].x, 0.f);  // img_ft[img_ft<0] = 0 (mv:5671)
              if (MODE == RL_PLAIN) {
                obase[o] = blur;
              } else if (MODE == RL_RATIO) {
                const float I = A.view_u16 ? u16_to_float(eV[c][h]) : __uint_as_float(eV[c][h]);
                const float ratio = __fdividef(I, blur + 1e-6f);
                obase[o] = eW[c][h] > 1e-5f ? ratio : 0.f;
              } else {
                const float ow = blur * eW[c][h];
                const float corr = A.first ? ow : eC[c][h] + ow;
                if (A.last) {
                  const float v = fminf(fmaxf(__uint_as_float(eV[c][h]) * corr, 0.f), 65535.f);
                  ebase[o] = v;
                  local_sum += (double)v;
                } else {
                  obase[o] = corr;
                }
              }
            }
          }
        }
      }
    }
    if (park && late) stash(step & 1, NE, NL);
  }
  if (MODE == RL_CORRECT && A.last && A.sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_sum += __shfl_xor_sync(0xffffffffu, local_sum, o);
    if (lane == 0) atomicAdd(A.sum, local_sum);
  }
}

template <int K, int MODE, int TY, int CPT, int XS>
static int launch_rl_pair_cfg(const RLArgs& A0, cudaStream_t st) {
  constexpr int R = (K - 1) / 2, LEAD = (R + 3) & ~3, KT = (K + 3) & ~3;
  constexpr int AY = TY + 2 * R, AX = RL_TX + 2 * LEAD;
  constexpr int UNITS = (AX / 2) | 1, PI = 2 * UNITS, PB = RL_TX + 1;
  RLArgs A = A0;
  // z chunks: even (two planes per step), ~6 waves of CTAs, at least 64 planes so the 2R-plane ramp stays small
  constexpr int NT = 32 * (TY / CPT);
  const int per_sm = NT <= 256 ? 2 : 1;
  const int cols = ((A.nx + RL_TX - 1) / RL_TX) * ((A.ny + TY - 1) / TY);
  const int chunks = (6 * per_sm * sm_count() + cols - 1) / cols;
  int zc = (A.nz + chunks - 1) / chunks;
  if (zc < 64) zc = 64;
  zc = (zc + 1) & ~1;
  const long long plane = (long long)A.ny * A.nx;   // 32-bit voxel offsets inside a chunk
  while ((long long)(zc + 2) * plane >= (1LL << 31) && zc > 2) zc = ((zc / 2) + 1) & ~1;
  if (zc > A.nz) zc = A.nz;
  A.zchunk = zc;
  dim3 grid((A.nx + RL_TX - 1) / RL_TX, (A.ny + TY - 1) / TY, (A.nz + zc - 1) / zc);
  const size_t smem = (size_t)(2 * AY * PI + 2 * AY * PB) * sizeof(float2) + (size_t)(3 * KT + AY + AX + zc + 2 * R + 2) * sizeof(int);
  auto kern = rl_pair_kernel<K, MODE, TY, CPT, XS>;
  MVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NT, smem, st>>>(A);
  MVF_LAUNCHED();
  return MVF_OK;
}

}  // namespace mvf
