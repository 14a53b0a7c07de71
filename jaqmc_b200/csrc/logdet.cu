// Envelope product, multi-determinant slogdet with its forward-Laplacian rule, and the log-sum-exp
// combination that yields log|psi|, sign, grad log psi, laplacian and the kinetic energy.
//
// Reference semantics:
//   * envelope  E_ik = sum_I pi_kI exp(-|sigma_kI| r_iI)          wavefunction/output/envelope.py:98-140
//     (isotropic: exponent sigma * -r; diagonal: exp(-||sigma_kI (.) r_iI||), sigma (n_orb, A, 3, D), :143-163)
//   * orbitals * envelope (product rule, Local1 x dense)           app/molecule/wavefunction/ferminet.py:90-93,
//                                                                  laplacian/primitives/core.py:450-501
//   * slogdet rule ld_J[k] = tr(A^-1 dA_k), ld_L = tr(A^-1 A_L) - sum_k tr((A^-1 dA_k)^2)
//                                                                  laplacian/primitives/slogdet.py:46-72
//   * log-sum-exp over determinants with max shift                 wavefunction/output/logdet.py:65-79
//   * E_kin = -1/2 (lap + |grad|^2)                                estimator/kinetic/_common.py:61-73
#include <cstdlib>

#include "aug.cuh"

// q / d and q % d for block-uniform runtime d without the ~20-instruction integer division sequence
// (inv = 1.0f / d; exact after the one-step correction for q < 2^23).
__device__ __forceinline__ void jq_divmod(int q, int d, float inv, int* quo, int* rem) {
  int a = (int)((float)q * inv);
  int r = q - a * d;
  if (r >= d) {
    r -= d;
    ++a;
  } else if (r < 0) {
    r += d;
    --a;
  }
  *quo = a;
  *rem = r;
}

// ------------------------------------------------------------------------------------------------
// orb[w][j][c][d*n+i] *= envelope(j, i, d)   (in place, product rule).  One item per (w, j, d, i).
// ------------------------------------------------------------------------------------------------
__global__ void k_orb_envelope(float* __restrict__ orb, const float* __restrict__ el, const float* __restrict__ atoms,
                               JqEnvelopeArgs env, long long items, JqSpins sp, int A, int D, int track, int use_smem) {
  const int n = sp.n();
  const int DN = D * n;
  const int C = track ? 3 * n + 2 : 1;
  // Optional staging of the envelope parameters [n_orb][A][D] in shared memory, transposed to [channel][atom][col]
  // (col = d * n + i is the fastest index of consecutive items: the direct reads are A * D floats apart per lane).
  JQ_DYN_SMEM(float, sm);
  const int nch_p = (env.pi[1] != nullptr) ? 2 : 1;
  if (use_smem) {
    for (int q = threadIdx.x; q < nch_p * A * DN; q += blockDim.x) {
      const int ch = q / (A * DN), r = q - ch * (A * DN);
      const int I = r / DN, col = r - I * DN;
      const int d = col / n, i = col - d * n;
      float sv = env.sigma[ch][(i * A + I) * D + d];
      if (env.type == 1) sv = fabsf(sv);
      sm[q] = sv;
      sm[nch_p * A * DN + q] = env.pi[ch][(i * A + I) * D + d];
    }
    __syncthreads();
  }
  const bool small = items < 0x7fffffffLL;   // 32-bit index arithmetic (the 64-bit divisions dominate the value-only pass)
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    int col;
    long long g;  // (w, j)
    int j;
    if (small) {
      const unsigned u = (unsigned)it, gq = u / (unsigned)DN;
      col = (int)(u - gq * (unsigned)DN);
      g = gq;
      j = (int)(gq % (unsigned)n);
    } else {
      col = (int)(it % DN);
      g = it / DN;
      j = (int)(g % n);
    }
    int d = col / n, i = col % n;
    int ch = (env.pi[1] != nullptr) ? sp.chan_of(j) : 0;
    const float* pi = env.pi[ch];
    const float* sg = env.sigma[ch];
    const float* e = el + g * 3;
    float ex = 0.f, ej[3] = {0.f, 0.f, 0.f}, elap = 0.f;
    for (int I = 0; I < A; ++I) {
      float dx = e[0] - atoms[I * 3], dy = e[1] - atoms[I * 3 + 1], dz = e[2] - atoms[I * 3 + 2];
      if (env.type == 3) {
        // diagonal envelope: rho = ||sigma (.) r_vec||, t = pi exp(-rho);
        //   dt/dx_a = -t sigma_a^2 x_a / rho,   lap t = t (sum_a sigma_a^4 x_a^2 (1/rho^2 + 1/rho^3) - sum_a sigma_a^2 / rho)
        const float* s3 = sg + ((long long)(i * A + I) * 3) * D + d;
        const float s0 = s3[0] * s3[0], s1 = s3[D] * s3[D], s2 = s3[2 * D] * s3[2 * D];
        const float q0 = s0 * dx, q1 = s1 * dy, q2 = s2 * dz;          // sigma_a^2 x_a
        const float rho = sqrtf(q0 * dx + q1 * dy + q2 * dz);
        const float t = pi[(i * A + I) * D + d] * expf(-rho);
        ex += t;
        if (track) {
          const float rinv = 1.0f / rho;
          const float c1 = -t * rinv;
          ej[0] += c1 * q0;
          ej[1] += c1 * q1;
          ej[2] += c1 * q2;
          const float qq = q0 * q0 + q1 * q1 + q2 * q2;
          elap += t * (qq * rinv * rinv * (1.0f + rinv) - (s0 + s1 + s2) * rinv);
        }
        continue;
      }
      float r = sqrtf(dx * dx + dy * dy + dz * dz);
      float s, pv;
      if (use_smem) {
        s = sm[(ch * A + I) * DN + col];
        pv = sm[(nch_p + ch) * A * DN + I * DN + col];
      } else {
        s = sg[(i * A + I) * D + d];
        if (env.type == 1) s = fabsf(s);
        pv = pi[(i * A + I) * D + d];
      }
      float t = pv * expf(-s * r);
      ex += t;
      if (track) {
        float rinv = 1.0f / r;
        float c1 = -s * t * rinv;
        ej[0] += c1 * dx;
        ej[1] += c1 * dy;
        ej[2] += c1 * dz;
        elap += t * (s * s - 2.0f * s * rinv);
      }
    }
    float* o = orb + g * (long long)C * DN + col;
    float ox = o[0];
    if (track) {
      float cross = 0.f;
      for (int a = 0; a < 3; ++a) cross = fmaf(o[(long long)(1 + 3 * j + a) * DN], ej[a], cross);
      float ol = o[(long long)(C - 1) * DN];
      o[(long long)(C - 1) * DN] = ol * ex + ox * elap + 2.0f * cross;
      for (int c = 1; c < C - 1; ++c) {
        float v = o[(long long)c * DN] * ex;
        int k = c - 1;
        if (k / 3 == j) v = fmaf(ox, ej[k % 3], v);
        o[(long long)c * DN] = v;
      }
    }
    o[0] = ox * ex;
  }
}

#ifndef JAQMC_HOST_EMU
// Value-only envelope product (sampling path): a thread owns one column (determinant d, orbital i) with its pi / sigma
// in registers and walks groups (walker, electron); no per-item index arithmetic.  Up to ENVV_A atoms.
constexpr int ENVV_A = 8;
constexpr int ENVV_GP = 32;   // groups per block
__global__ void __launch_bounds__(256) k_orb_envelope_value(float* __restrict__ orb, const float* __restrict__ el,
                                                           const float* __restrict__ atoms, JqEnvelopeArgs env,
                                                           long long G, JqSpins sp, int A, int D) {
  const int n = sp.n(), DN = D * n;
  const long long g0 = (long long)blockIdx.x * ENVV_GP;
  const long long g1 = (g0 + ENVV_GP < G) ? g0 + ENVV_GP : G;
  const bool two = env.pi[1] != nullptr;
  float ax[ENVV_A], ay[ENVV_A], az[ENVV_A];
#pragma unroll
  for (int I = 0; I < ENVV_A; ++I) {
    ax[I] = (I < A) ? atoms[3 * I] : 0.f;
    ay[I] = (I < A) ? atoms[3 * I + 1] : 0.f;
    az[I] = (I < A) ? atoms[3 * I + 2] : 0.f;
  }
  for (int col = threadIdx.x; col < DN; col += blockDim.x) {
    const int d = col / n, i = col - d * n;
    float pv[2][ENVV_A], sv[2][ENVV_A];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
      for (int I = 0; I < ENVV_A; ++I) {
        const bool on = I < A && (ch == 0 || two);
        float s = on ? env.sigma[ch][(i * A + I) * D + d] : 0.f;
        if (env.type == 1) s = fabsf(s);
        sv[ch][I] = s;
        pv[ch][I] = on ? env.pi[ch][(i * A + I) * D + d] : 0.f;
      }
    // groups in batches of 8: their orbital values and electron positions are requested together
    for (long long gb = g0; gb < g1; gb += 8) {
      float ov[8], px[8], py[8], pz[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long g = (gb + u < g1) ? gb + u : g1 - 1;
        ov[u] = orb[g * DN + col];
        px[u] = el[g * 3];
        py[u] = el[g * 3 + 1];
        pz[u] = el[g * 3 + 2];
      }
      int j = (int)(gb % n);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool c1 = two && sp.chan_of(j) == 1;
        float e = 0.f;
#pragma unroll
        for (int I = 0; I < ENVV_A; ++I)
          if (I < A) {
            const float dx = px[u] - ax[I], dy = py[u] - ay[I], dz = pz[u] - az[I];
            const float r = sqrtf(dx * dx + dy * dy + dz * dz);
            e += (c1 ? pv[1][I] : pv[0][I]) * expf(-(c1 ? sv[1][I] : sv[0][I]) * r);
          }
        if (gb + u < g1) orb[(gb + u) * DN + col] = ov[u] * e;
        if (++j == n) j = 0;
      }
    }
  }
}
#endif

int jq_launch_orb_envelope(float* orb, const float* electrons, const float* atoms, const JqEnvelopeArgs& env,
                           int W, JqSpins sp, int A, int D, int track, cudaStream_t st) {
  if (env.type == 2) return JQ_OK;  // null envelope: ones
  long long items = (long long)W * sp.n() * D * sp.n();
  if (items <= 0) return JQ_OK;
#ifndef JAQMC_HOST_EMU
  if (!track && A <= ENVV_A && env.type != 3) {
    const long long G = (long long)W * sp.n();
    jq_prof_work(0.0, 8.0 * (double)items);
    JQ_LAUNCH(k_orb_envelope_value, dim3((unsigned)jq_cdiv(G, ENVV_GP)), dim3(256), 0, st, orb, electrons, atoms, env, G, sp,
              A, D);
    JQ_CHECK_LAUNCH();
    return JQ_OK;
  }
#endif
  int grid = jq_cdiv(items, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  jq_prof_work(0.0, 8.0 * (double)items * (track ? 3 * sp.n() + 2 : 1));
  const int nch_p = env.pi[1] ? 2 : 1;
  const size_t smem = sizeof(float) * 2 * (size_t)nch_p * A * D * sp.n();
  const int use_smem = smem <= 40 * 1024 && env.type != 3;
  JQ_LAUNCH(k_orb_envelope, dim3(grid), dim3(256), use_smem ? smem : 0, st, orb, electrons, atoms, env, items, sp, A, D,
            track, use_smem);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

#define LD_NP 16
#ifndef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// slogdet + forward-Laplacian rule for n <= 16 on the tensor cores (r2): one WARP per determinant.
//   M_c = A^-1 dA_c for every component c, ld_J[c] = tr M_c, ld_L = tr(A^-1 A_L) - sum_k tr(M_k^2)   (slogdet.py:46-72)
// Phase 1: Gauss-Jordan inversion in registers (k_logdet_small's, run redundantly by both half-warps); A^-1 goes to a
// zero-padded 16 x 16 shared tile and from there into mma.sync fragments that stay in registers for all 3n+1 slabs.
// Phase 2, per slab: the thread loads the 8 entries dA[j][i2], j = t + 4s, i2 = g + 8r (g = lane / 4, t = lane % 4) --
// which are at once the B fragments of  M = A^-1 dA  and the A fragments of  M^T = dA^T A^-T  (m16n8k8: the A fragment of
// X^T and the B fragment of X hold the same elements; likewise the A^-1 registers serve as A fragment of the first and
// B fragment of the second product).  Both products land in the same accumulator layout, so
//   tr M^2 = sum_ab M[a][b] M^T[a][b]   and   tr M
// are thread-local sums followed by one warp reduction: no shared-memory traffic in the loop (the FP32 kernel it
// replaces spent its time re-reading A^-1 from shared memory: 512 B per LDS.128 warp instruction).
// Arithmetic: 3xTF32 (x = hi + lo; lo*hi + hi*lo + hi*hi, FP32 accumulation), i.e. FP32-faithful to ~2^-21.
// ------------------------------------------------------------------------------------------------
#define LDM_KMAX 50   // slabs: 3 * 16 + 1
__device__ __forceinline__ unsigned jq_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void jq_mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                            unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Index maps of the fragments.  The products are invariant under a common permutation pi of the matrix index (rows and
// columns of M together): pi(g + 8r) = 2g + r puts the two columns a thread owns next to each other in memory, so a
// slab costs four 8-byte loads per thread (n even; odd n takes 4-byte loads).
template <bool VEC2>
__global__ void __launch_bounds__(256, 2) k_logdet_mma(const float* __restrict__ orb, int n, int D, int C, long long MT,
                                                        float* __restrict__ det_sign, float* __restrict__ det_logabs,
                                                        float* __restrict__ det_grad, float* __restrict__ det_lap) {
  __shared__ float tile_all[8][16 * 17];
  __shared__ float diag_all[8][LDM_KMAX][8];   // per-slab partial traces of the 8 diagonal-holding lanes
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long m = (long long)blockIdx.x * 8 + warp;   // matrix = (walker, determinant)
  if (m >= MT) return;
  const long long w = m / D;
  const int d = (int)(m - w * D);
  const int DN = D * n;
  const int K = C - 2, KT = C - 1;
  const float* ow = orb + (w * n) * (long long)C * DN + d * n;  // (j, c, i) at ow[(j*C + c)*DN + i]
  float* tile = tile_all[warp];
  for (int q = lane; q < 16 * 17; q += 32) tile[q] = 0.f;
  __syncwarp();
  // ---- phase 1: [A | I] -> [I | A^-1], lane hl = row hl, both half-warps on the same matrix
  {
    const int hl = lane & 15;
    bool used = hl >= n;
    float a[LD_NP], bi[LD_NP];
#pragma unroll
    for (int c = 0; c < LD_NP; ++c) {
      a[c] = (!used && c < n) ? ow[(long long)hl * C * DN + c] : 0.f;
      bi[c] = (c == hl) ? 1.0f : 0.f;
    }
    int step_of = 0;
    float sg = 1.0f;
    double mant = 1.0;
    int expo = 0;
#pragma unroll
    for (int p = 0; p < LD_NP; ++p) {
      if (p < n) {
        unsigned key = used ? 0u : __float_as_uint(fabsf(a[p]));
        unsigned mx = key;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(full, mx, o));
        const unsigned cand = __ballot_sync(full, !used && key == mx) & 0xffffu;
        const int pl = cand ? __ffs(cand) - 1 : 0;   // first unused row holding the largest magnitude
        const float pv = __shfl_sync(full, a[p], pl, 16);
        if (pv < 0.f) sg = -sg;
        if (pv == 0.f) sg = 0.f;
        {
          int e;
          mant *= (double)frexpf(fabsf(pv), &e);
          expo += e;
        }
        const float pinv = 1.0f / pv;
        const bool is_p = (hl == pl);
        if (is_p && !used) {
          used = true;
          step_of = p;
        }
        const float f = is_p ? 0.f : a[p];   // column p before it is overwritten
#pragma unroll
        for (int c = 0; c < LD_NP; ++c)
          if (c < n) {
            const float ap = __shfl_sync(full, a[c], pl, 16) * pinv;
            const float bp = __shfl_sync(full, bi[c], pl, 16) * pinv;
            if (is_p) {
              a[c] = ap;
              bi[c] = bp;
            } else {
              a[c] = fmaf(-f, ap, a[c]);
              bi[c] = fmaf(-f, bp, bi[c]);
            }
          }
      }
    }
    int invc = 0;
    for (int i = 0; i < n; ++i) {
      const int si = __shfl_sync(full, step_of, i, 16);
      if (i < hl && hl < n && si > step_of) ++invc;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) invc += __shfl_xor_sync(full, invc, o);
    if (lane < n) {
#pragma unroll
      for (int c = 0; c < LD_NP; ++c)
        if (c < n) tile[step_of * 17 + c] = bi[c];
    }
    if (lane == 0) {
      det_sign[m] = (invc & 1) ? -sg : sg;
      det_logabs[m] = (float)(log(mant) + (double)expo * 0.69314718055994530942);
    }
  }
  __syncwarp();
  // ---- A^-1 fragments (hi / lo), kept for every slab: ai[r][s] = A^-1[pi(g + 8r)][t + 4s]
  const int g = lane >> 2, t = lane & 3;
  const int i0 = 2 * g;   // pi(g) = 2g, pi(g + 8) = 2g + 1
  unsigned aih[2][4], ail[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) {
      const float v = tile[(i0 + r) * 17 + t + 4 * s2];
      aih[r][s2] = jq_tf32(v);
      ail[r][s2] = jq_tf32(v - __uint_as_float(aih[r][s2]));   // rounded, not left to the tensor core's truncation (measured: 5x the error)
    }
  // ---- phase 2
  bool okj[4];
#pragma unroll
  for (int s2 = 0; s2 < 4; ++s2) okj[s2] = (t + 4 * s2) < n;
  const bool ok0 = i0 < n, ok1 = i0 + 1 < n;
  const float* rowp[4];   // slab 0 of row j = t + 4s, column i0
#pragma unroll
  for (int s2 = 0; s2 < 4; ++s2) rowp[s2] = ow + ((long long)(okj[s2] ? t + 4 * s2 : 0) * C + 1) * DN + (ok0 ? i0 : 0);
  auto fetch = [&](float (&dst)[2][4]) {
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) {
      if (VEC2) {
        float2 v = make_float2(0.f, 0.f);
        if (okj[s2] && ok0) v = *reinterpret_cast<const float2*>(rowp[s2]);   // n even: i0 + 1 < n whenever i0 < n
        dst[0][s2] = v.x;
        dst[1][s2] = v.y;
      } else {
        dst[0][s2] = (okj[s2] && ok0) ? rowp[s2][0] : 0.f;
        dst[1][s2] = (okj[s2] && ok1) ? rowp[s2][1] : 0.f;
      }
      rowp[s2] += DN;
    }
  };
  float nxt[2][4];
  fetch(nxt);
  // lanes holding diagonal entries of the accumulator tiles: fragment row g has fragment columns 2t, 2t + 1
  const int e = g - 2 * t;
  const bool has_diag = (e == 0 || e == 1);
  float* dsm = diag_all[warp][0] + (2 * t + (e & 1));   // slot of this lane among the 8 diagonal holders
  float t2acc = 0.f;
  for (int kk = 0; kk < KT; ++kk) {
    unsigned dh[2][4], dl[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {
        const float v = nxt[r][s2];
        dh[r][s2] = jq_tf32(v);
        dl[r][s2] = jq_tf32(v - __uint_as_float(dh[r][s2]));
      }
    if (kk + 1 < KT) fetch(nxt);
    float c1[2][4], c2[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q) c1[u][q] = c2[u][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        // M = A^-1 dA : A fragment from A^-1 (rows g, g+8; k = 8ks + t, +4), B fragment from dA (k = 8ks + t, +4; col 8u + g)
        jq_mma_tf32(c1[u], ail[0][2 * ks], ail[1][2 * ks], ail[0][2 * ks + 1], ail[1][2 * ks + 1], dh[u][2 * ks], dh[u][2 * ks + 1]);
        jq_mma_tf32(c1[u], aih[0][2 * ks], aih[1][2 * ks], aih[0][2 * ks + 1], aih[1][2 * ks + 1], dl[u][2 * ks], dl[u][2 * ks + 1]);
        jq_mma_tf32(c1[u], aih[0][2 * ks], aih[1][2 * ks], aih[0][2 * ks + 1], aih[1][2 * ks + 1], dh[u][2 * ks], dh[u][2 * ks + 1]);
        // M^T = dA^T A^-T : A fragment from dA (rows i2 = g, g+8; k = 8ks + t, +4), B fragment from A^-1 (col i = 8u + g)
        jq_mma_tf32(c2[u], dl[0][2 * ks], dl[1][2 * ks], dl[0][2 * ks + 1], dl[1][2 * ks + 1], aih[u][2 * ks], aih[u][2 * ks + 1]);
        jq_mma_tf32(c2[u], dh[0][2 * ks], dh[1][2 * ks], dh[0][2 * ks + 1], dh[1][2 * ks + 1], ail[u][2 * ks], ail[u][2 * ks + 1]);
        jq_mma_tf32(c2[u], dh[0][2 * ks], dh[1][2 * ks], dh[0][2 * ks + 1], dh[1][2 * ks + 1], aih[u][2 * ks], aih[u][2 * ks + 1]);
      }
    }
    // tr M^2 = sum_ab M[a][b] M^T[a][b]: thread-local partial, reduced over the warp once after the loop
    if (kk < K) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) t2acc = fmaf(c1[u][q], c2[u][q], t2acc);
    }
    // tr M: fragment rows g hold columns 2t, 2t+1 of tile u = 0; rows g + 8 hold columns 8 + 2t, 8 + 2t + 1 of tile u = 1
    if (has_diag) dsm[kk * 8] = (e == 0) ? c1[0][0] + c1[1][2] : c1[0][1] + c1[1][3];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t2acc += __shfl_xor_sync(full, t2acc, o);
  __syncwarp();
  float trl = 0.f;
  for (int kk = lane; kk < KT; kk += 32) {
    const float* p8 = diag_all[warp][kk];
    const float v = ((p8[0] + p8[1]) + (p8[2] + p8[3])) + ((p8[4] + p8[5]) + (p8[6] + p8[7]));
    if (kk < K) det_grad[m * (long long)K + kk] = v;
    else trl = v;
  }
  trl = __shfl_sync(full, trl, K % 32);   // the Laplacian row is slab K
  if (lane == 0) det_lap[m] = trl - t2acc;
}
#endif

#ifndef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// slogdet + forward-Laplacian rule for n <= 16 (r2): one HALF-warp per determinant, no block-level synchronisation.
// Phase 1: Gauss-Jordan inversion in registers (lane r = row r of [A | I]); rows are never moved: at step p the pivot is
// the largest |a[r][p]| among the rows not used yet, its row is scaled and eliminated from all others (same pivots and
// fmaf sequence as the shared-memory elimination of k_logdet).  Row p of A^-1 is the right half of the row that served
// as pivot p; it goes into a row-padded shared copy.  sign = parity(row -> step) * prod sign(pivot),
// log|det| = log prod |pivot| in double.   Phase 2: the traces, see inside.
// Shared: invp[DB][n*16+4] | scr[DB][MS]
// ------------------------------------------------------------------------------------------------
// ST (r2, late): the derivative slabs of the warp's two determinants (n rows of 2n contiguous floats per slab) are
// staged by the warp itself with 16-byte cp.async copies into a private double buffer, NS slabs (one pass) ahead, instead
// of 14 dependent 4-byte global loads per lane and slab: no global-load latency on the critical path at 4 warps per
// scheduler.  Needs n and D even (16-byte source alignment); same FMA order, results bit-identical to ST = false.
// Shared (ST): ... | ring[8 warps][2][NS][n][2n]
template <int NS, bool ST>
__global__ void __launch_bounds__(256, 2) k_logdet_small(const float* __restrict__ orb, int n, int D, int C, int DB,
                                                          float* __restrict__ det_sign, float* __restrict__ det_logabs,
                                                          float* __restrict__ det_grad, float* __restrict__ det_lap) {
  JQ_DYN_SMEM(float, sm);
  const int nn = n * n;
  const int K = C - 2, KT = C - 1;
  const int ngrp = (D + DB - 1) / DB;
  const long long w = blockIdx.x / ngrp;
  const int d0 = (int)(blockIdx.x % ngrp) * DB;
  const int db = (D - d0 < DB) ? D - d0 : DB;
  const int DN = D * n;
  const int tid = threadIdx.x;
  const float* ow = orb + (w * n) * (long long)C * DN + d0 * n;  // (j, c, d, i) at ow[(j*C + c)*DN + d*n + i]
  float* Jc = sm;
  {
    const unsigned full = 0xffffffffu;
    const int hl = tid & 15, half = (tid >> 4) & 1, warp = tid >> 5;
    const int IS = n * LD_NP + 4;
    float* invp_w = Jc;
    for (int dbase = 2 * warp; dbase < db; dbase += 16) {
      const int d = dbase + half;
      const bool on = d < db;
      bool used = !on || hl >= n;
      // ---- ST: per-warp staging of the derivative slabs of determinants (dbase, dbase + 1)
      const int lane = tid & 31;
      const int n2 = 2 * n;
      const int nch = n * (n >> 1);   // 16-byte chunks per slab: n rows of 2n floats
      int soff[4];                    // source offset (floats) of this lane's chunks q = lane + 32 i within a slab
      uint32_t ring_w = 0;            // shared address of this warp's ring [2][NS][n][2n]
      const float* ring_f = nullptr;
      const float* owd = ow + dbase * n;
      if (ST) {
        const int MSr = nn + ((n - nn) % 32 + 32) % 32;
        float* rf = Jc + (size_t)DB * IS + (size_t)DB * NS * MSr + (size_t)warp * (2 * NS * n * n2);
        ring_f = rf;
        ring_w = (uint32_t)__cvta_generic_to_shared(rf);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = lane + 32 * i;
          const int j = q / (n >> 1), part = q - j * (n >> 1);
          soff[i] = j * C * DN + part * 4;
        }
      }
      auto stage_pass = [&](int it) {   // slabs it * NS ... of this warp's determinant pair -> buffer it & 1
#pragma unroll
        for (int s2i = 0; s2i < NS; ++s2i) {
          const int kq = it * NS + s2i;
          if (kq < KT) {
            const float* src = owd + (long long)(1 + kq) * DN;
            const uint32_t dst = ring_w + 4u * (unsigned)((((it & 1) * NS + s2i) * n) * n2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int q = lane + 32 * i;
              if (q < nch)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (unsigned)q), "l"(src + soff[i]) : "memory");
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      if (ST) stage_pass(0);   // in flight during the inversion
      float a[LD_NP], bi[LD_NP];
#pragma unroll
      for (int c = 0; c < LD_NP; ++c) {
        a[c] = (!used && c < n) ? ow[(long long)hl * C * DN + d * n + c] : 0.f;
        bi[c] = (c == hl) ? 1.0f : 0.f;
      }
      int step_of = 0;
      float sg = 1.0f;
      double mant = 1.0;
      int expo = 0;
#pragma unroll
      for (int p = 0; p < LD_NP; ++p) {
        if (p < n) {
          unsigned key = used ? 0u : __float_as_uint(fabsf(a[p]));
          unsigned mx = key;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(full, mx, o));
          const unsigned cand = (__ballot_sync(full, !used && key == mx) >> (16 * half)) & 0xffffu;
          const int pl = cand ? __ffs(cand) - 1 : 0;   // first unused row holding the largest magnitude
          const float pv = __shfl_sync(full, a[p], pl, 16);
          if (pv < 0.f) sg = -sg;
          if (pv == 0.f) sg = 0.f;
          {
            int e;
            mant *= (double)frexpf(fabsf(pv), &e);
            expo += e;
          }
          const float pinv = 1.0f / pv;
          const bool is_p = (hl == pl);
          if (is_p && !used) {
            used = true;
            step_of = p;
          }
          const float f = is_p ? 0.f : a[p];   // column p before it is overwritten
#pragma unroll
          for (int c = 0; c < LD_NP; ++c)
            if (c < n) {
              // pivot row, scaled (its entry in column p becomes 1)
              const float ap = __shfl_sync(full, a[c], pl, 16) * pinv;
              const float bp = __shfl_sync(full, bi[c], pl, 16) * pinv;
              if (is_p) {
                a[c] = ap;
                bi[c] = bp;
              } else {
                a[c] = fmaf(-f, ap, a[c]);
                bi[c] = fmaf(-f, bp, bi[c]);
              }
            }
        }
      }
      // permutation parity: inversions of row -> step, counted per lane and summed over the half-warp
      int invc = 0;
      for (int i = 0; i < n; ++i) {
        const int si = __shfl_sync(full, step_of, i, 16);
        if (i < hl && hl < n && si > step_of) ++invc;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) invc += __shfl_xor_sync(full, invc, o);
      if (on && hl < n) {
        float* dst = invp_w + (size_t)d * IS + step_of * LD_NP;
#pragma unroll
        for (int c = 0; c < LD_NP; ++c) dst[c] = (c < n) ? bi[c] : 0.f;
      }
      if (on && hl == 0) {
        det_sign[w * D + d0 + d] = (invc & 1) ? -sg : sg;
        det_logabs[w * D + d0 + d] = (float)(log(mant) + (double)expo * 0.69314718055994530942);
      }
      // ---- traces of this determinant, by the same half-warp and without any block-level synchronisation (r2):
      // lane i2 holds column i2 of dA_c in registers and forms column i2 of M = A^-1 dA_c (the inverse is read as
      // float4 broadcasts from the row-padded copy this half-warp has just written); tr M is the sum of the lanes'
      // diagonal entries, tr M^2 = sum_{i,i2} M[i][i2] M[i2][i] needs the transposed element, exchanged through a
      // per-determinant scratch tile.  NS slabs share one pass over A^-1 (each float4 of the inverse feeds NS columns).
      // Measured r2 (ncu, N2 4096 walkers x 16 determinants): NS = 1 with a register prefetch of the next column 1.88 ms,
      // NS = 2 1.38 ms (default), NS = 3 1.41 ms, NS = 4 1.63 ms (spills at the 128-register cap); dropping the kept
      // column of M in favour of re-reading the scratch tile, or prefetching with NS = 2, both lose (1.45 - 1.55 ms).
      __syncwarp();
      {
        // an idle half-warp (odd determinant count) addresses the slot of its own warp's other half: written before the
        // __syncwarp above, so it never reads a tile another warp is still writing (compute-sanitizer racecheck, r2)
        const int ds = on ? d : dbase;
        const float* ib = invp_w + (size_t)ds * IS;
        float* scr = invp_w + (size_t)DB * IS + (size_t)ds * NS * (nn + ((n - nn) % 32 + 32) % 32);
        const bool act = on && hl < n;
        const float* ocol = ow + ds * n + (hl < n ? hl : 0);   // column (d, i2 = hl); + (1 + kk) * DN + j * C * DN
        float t2acc = 0.f, trl = 0.f;
        float* gout = det_grad + (w * D + d0 + ds) * (long long)K;
        const int MS = nn + ((n - nn) % 32 + 32) % 32;
        // NS slabs per pass over A^-1: each float4 of the inverse read from shared memory feeds NS columns.
        for (int kk = 0; kk < KT; kk += NS) {
          float col[NS][LD_NP];
          if (ST) {
            const int it = kk / NS;
            stage_pass(it + 1);   // the other buffer: its readers finished before the __syncwarp that ended pass it - 1
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            const float* stg = ring_f + (size_t)((it & 1) * NS) * n * n2 + half * n + (hl < n ? hl : 0);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              const bool sv = act && kk + s < KT;
#pragma unroll
              for (int j = 0; j < LD_NP; ++j) col[s][j] = (sv && j < n) ? stg[(s * n + j) * n2] : 0.f;
            }
          } else {
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            const bool sv = act && kk + s < KT;
            const float* oc = ocol + (long long)(1 + kk + s) * DN;
#pragma unroll
            for (int j = 0; j < LD_NP; ++j) col[s][j] = (sv && j < n) ? oc[(long long)j * C * DN] : 0.f;
          }
          }
          float m[NS][LD_NP];
          float dg[NS];
#pragma unroll
          for (int s = 0; s < NS; ++s) dg[s] = 0.f;
#pragma unroll
          for (int i = 0; i < LD_NP; ++i) {
            float acc[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[s] = 0.f;
            if (i < n) {
              const float4* r4 = reinterpret_cast<const float4*>(ib + i * LD_NP);
#pragma unroll
              for (int j4 = 0; j4 < LD_NP / 4; ++j4) {
                const float4 v = r4[j4];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                  acc[s] = fmaf(v.x, col[s][4 * j4 + 0], acc[s]);
                  acc[s] = fmaf(v.y, col[s][4 * j4 + 1], acc[s]);
                  acc[s] = fmaf(v.z, col[s][4 * j4 + 2], acc[s]);
                  acc[s] = fmaf(v.w, col[s][4 * j4 + 3], acc[s]);
                }
              }
              if (act) {
#pragma unroll
                for (int s = 0; s < NS; ++s) scr[s * MS + i * n + hl] = acc[s];
              }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              m[s][i] = acc[s];
              if (i == hl) dg[s] = acc[s];
            }
          }
          __syncwarp();
          float s2[NS];
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            s2[s] = 0.f;
#pragma unroll
            for (int i = 0; i < LD_NP; ++i)
              if (i < n && act) s2[s] = fmaf(m[s][i], scr[s * MS + hl * n + i], s2[s]);
          }
          __syncwarp();
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            float dgs = act ? dg[s] : 0.f, s2s = s2[s];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
              dgs += __shfl_xor_sync(full, dgs, o);
              s2s += __shfl_xor_sync(full, s2s, o);
            }
            if (kk + s < K) {
              if (on && hl == 0) gout[kk + s] = dgs;
              t2acc += s2s;
            } else if (kk + s == K) {
              trl = dgs;
            }
          }
        }
        if (on && hl == 0) det_lap[w * D + d0 + d] = trl - t2acc;
        if (ST) {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
        }
      }
    }
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// slogdet + forward-Laplacian rule.  One block per (walker, group of DB determinants): a walker's orbital slab
// [n][C][D*n] is read in rows of DB*n contiguous floats, and all DB matrices go through every phase together.
// In-place Gauss-Jordan with partial (row) pivoting; sign = prod sign(pivot) * (-1)^swaps,
// log|det| = sum log|pivot| accumulated in double.  Then, KC derivative slabs at a time,
//   M = A^-1 dA_c,  ld_J[c] = tr M,  and  ld_L = tr(A^-1 A_L) - sum_k tr(M_k^2)      (primitives/slogdet.py:46-72).
// Shared: inv[DB][nn] | colp[DB][n] | piv[DB][n] | pivinv[DB] sgn[DB] trL[DB] t2[DB] | logabs[DB] (double) |
//         Jc[KC][DB][nn] | Mc[KC][DB][nn] | p1[KC][DB][n] | p2[KC][DB][n]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) k_logdet(const float* __restrict__ orb, int n, int D, int C, int DB, int KC,
                         float* __restrict__ det_sign, float* __restrict__ det_logabs, float* __restrict__ det_grad,
                         float* __restrict__ det_lap) {
  JQ_DYN_SMEM(float, sm);
  const int nn = n * n;
  const int K = C - 2;                 // Jacobian columns (C == 1: value only)
  const int KT = (C > 1) ? C - 1 : 0;  // J columns + the Laplacian row
  double* logabs = reinterpret_cast<double*>(sm);  // first, for 8-byte alignment
  float* inv = reinterpret_cast<float*>(logabs + DB);
  float* colp = inv + (size_t)DB * nn;
  int* piv = reinterpret_cast<int*>(colp + DB * n);
  float* pivinv = reinterpret_cast<float*>(piv + DB * n);
  float* sgn = pivinv + DB;
  float* trL = sgn + DB;
  float* t2 = trL + DB;
  float* Jc = sm + (((t2 + DB) - sm + 3) & ~3);  // 16-byte aligned (float4 reads of the padded inverses)
  const int NP4 = (n + 3) & ~3;        // row stride of the slab / transposed-inverse tiles: float4 reads (r2)
  float* Mc = Jc + (size_t)KC * DB * n * NP4;
  float* p1 = Mc + (size_t)KC * DB * nn;
  float* p2 = p1 + KC * DB * n;
  const int ngrp = (D + DB - 1) / DB;
  const long long w = blockIdx.x / ngrp;
  const int d0 = (int)(blockIdx.x % ngrp) * DB;
  const int db = (D - d0 < DB) ? D - d0 : DB;  // determinants in this block
  const int DN = D * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool need_inv = (C > 1);
  const int TX = (nt >= 64) ? 8 : 1, TY = nt / TX;
  const int tx = tid % TX, ty = tid / TX;
  const float inv_n = 1.0f / (float)n, inv_nn = 1.0f / (float)nn, inv_db = 1.0f / (float)db,
              inv_dbn = 1.0f / (float)(db * n);
  const float* ow = orb + (w * n) * (long long)C * DN + d0 * n;  // (j, c, d, i) at ow[(j*C + c)*DN + d*n + i]

  // value slab: inv[d][j][i] = A_d[j][i]; consecutive items read db*n contiguous floats
  for (int q = tid; q < n * db * n; q += nt) {
    int j, r, d, i;
    jq_divmod(q, db * n, inv_dbn, &j, &r);
    jq_divmod(r, n, inv_n, &d, &i);
    inv[d * nn + j * n + i] = ow[(long long)j * C * DN + r];
  }
  for (int d = tid; d < db; d += nt) {
    logabs[d] = 0.0;
    sgn[d] = 1.0f;
    t2[d] = 0.f;
    trL[d] = 0.f;
  }
  __syncthreads();
  const bool fast_inv = false;   // n <= 16 with 256 threads runs k_logdet_small on the device (r2)
  if (!fast_inv) {
  for (int p = 0; p < n; ++p) {
    for (int d = tid; d < db; d += nt) {
      const float* a = inv + d * nn;
      int r = p;
      float best = fabsf(a[p * n + p]);
      for (int q = p + 1; q < n; ++q) {
        float v = fabsf(a[q * n + p]);
        if (v > best) { best = v; r = q; }
      }
      piv[d * n + p] = r;
      float pv = a[r * n + p];
      float s = sgn[d];
      if (r != p) s = -s;
      if (pv < 0.f) s = -s;
      if (pv == 0.f) s = 0.f;
      sgn[d] = s;
      logabs[d] += log((double)fabsf(pv));
      pivinv[d] = 1.0f / pv;
    }
    __syncthreads();
    for (int q = tid; q < db * n; q += nt) {
      int d = q / n, c = q % n;
      int r = piv[d * n + p];
      float* a = inv + d * nn;
      if (r != p) {
        float t = a[p * n + c];
        a[p * n + c] = a[r * n + c];
        a[r * n + c] = t;
      }
    }
    __syncthreads();
    if (need_inv) {
      for (int q = tid; q < db * n; q += nt) {
        int d = q / n, i = q % n;
        colp[q] = inv[d * nn + i * n + p];  // column p before it is overwritten
      }
      __syncthreads();
      for (int q = tid; q < db * n; q += nt) {
        int d = q / n, c = q % n;
        float* a = inv + d * nn;
        a[p * n + c] = ((c == p) ? 1.0f : a[p * n + c]) * pivinv[d];
      }
      __syncthreads();
      for (int q = tid; q < db * nn; q += nt) {
        int d, rem, i, c;
        jq_divmod(q, nn, inv_nn, &d, &rem);
        jq_divmod(rem, n, inv_n, &i, &c);
        if (i == p) continue;
        float* a = inv + d * nn;
        float base = (c == p) ? 0.f : a[rem];
        a[rem] = fmaf(-colp[d * n + i], a[p * n + c], base);
      }
      __syncthreads();
    } else {
      // value only: eliminate below the pivot (LU), no inverse needed
      for (int q = tid; q < db * n; q += nt) {
        int d = q / n, i = q % n;
        colp[q] = inv[d * nn + i * n + p] * pivinv[d];
      }
      __syncthreads();
      for (int q = tid; q < db * nn; q += nt) {
        int d, rem, i, c;
        jq_divmod(q, nn, inv_nn, &d, &rem);
        jq_divmod(rem, n, inv_n, &i, &c);
        if (i <= p || c <= p) continue;
        float* a = inv + d * nn;
        a[rem] = fmaf(-colp[d * n + i], a[p * n + c], a[rem]);
      }
      __syncthreads();
    }
  }
  for (int d = tid; d < db; d += nt) {
    det_sign[w * D + d0 + d] = sgn[d];
    det_logabs[w * D + d0 + d] = (float)logabs[d];
  }
  if (!need_inv) return;
  // undo the row swaps as column swaps, in reverse order
  for (int p = n - 1; p >= 0; --p) {
    for (int q = tid; q < db * n; q += nt) {
      int d = q / n, i = q % n;
      int r = piv[d * n + p];
      if (r != p) {
        float* a = inv + d * nn;
        float t = a[i * n + p];
        a[i * n + p] = a[i * n + r];
        a[i * n + r] = t;
      }
    }
    __syncthreads();
  }
  }  // !fast_inv
  if (n <= LD_NP) {
    // ---- small matrices (n <= 16): one item per (determinant, column i2) and derivative slab.  The item holds
    // column i2 of dA_c in registers and forms column i2 of M = A^-1 dA_c with the inverse read as float4 broadcasts
    // from a row-padded copy: ~1.3 instructions per multiply-add instead of 3-4 for the shared-memory tiled product.
    // per-determinant strides padded so that the 2-3 determinants a warp touches fall into different banks:
    // IS = n*16 + 4 (float4 rows shifted by one 16-byte bank group), MS = nn rounded up to n modulo 32
    const int IS = n * LD_NP + 4;
    const int MS = nn + ((n - nn) % 32 + 32) % 32;
    float* invp = Jc;                       // [DB][IS]  (the slab area is carved differently on this path)
    float* Ms = invp + (size_t)DB * IS;     // [DB][MS]
    p1 = Ms + (size_t)DB * MS;
    p2 = p1 + DB * n;
    if (!fast_inv) {
      for (int q = tid; q < db * n * LD_NP; q += nt) {
        int d = q / (n * LD_NP), r = q % (n * LD_NP);
        int i = r / LD_NP, j = r % LD_NP;
        invp[d * IS + r] = (j < n) ? inv[d * nn + i * n + j] : 0.f;
      }
      __syncthreads();
    }
    // When every item has its own thread (db * n <= blockDim, the common case) the next slab's column is fetched
    // into registers before the current one is consumed, so the global-load latency overlaps the products.
    const bool one_item = (db * n <= nt);
    float nxt[LD_NP];
    if (one_item && tid < db * n) {
#pragma unroll
      for (int j = 0; j < LD_NP; ++j) nxt[j] = (j < n) ? ow[(long long)DN + (long long)j * C * DN + tid] : 0.f;
    }
    for (int kk = 0; kk < KT; ++kk) {
      const float* oc = ow + (long long)(1 + kk) * DN;  // slab: (j, d, i) at oc[j*C*DN + d*n + i]
      for (int q = tid; q < db * n; q += nt) {
        int d, i2;
        jq_divmod(q, n, inv_n, &d, &i2);
        float col[LD_NP];
        if (one_item) {
#pragma unroll
          for (int j = 0; j < LD_NP; ++j) col[j] = nxt[j];
          if (kk + 1 < KT) {
#pragma unroll
            for (int j = 0; j < LD_NP; ++j) nxt[j] = (j < n) ? oc[(long long)DN + (long long)j * C * DN + q] : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < LD_NP; ++j) col[j] = (j < n) ? oc[(long long)j * C * DN + q] : 0.f;
        }
        const float* ib = invp + (size_t)d * IS;
        float* mo = Ms + d * MS + i2;
        for (int i = 0; i < n; ++i) {
          const float4* r4 = reinterpret_cast<const float4*>(ib + i * LD_NP);
          float acc = 0.f;
#pragma unroll
          for (int j4 = 0; j4 < LD_NP / 4; ++j4) {
            float4 v = r4[j4];
            acc = fmaf(v.x, col[4 * j4 + 0], acc);
            acc = fmaf(v.y, col[4 * j4 + 1], acc);
            acc = fmaf(v.z, col[4 * j4 + 2], acc);
            acc = fmaf(v.w, col[4 * j4 + 3], acc);
          }
          mo[i * n] = acc;
        }
      }
      __syncthreads();
      for (int q = tid; q < db * n; q += nt) {
        int d, i2;
        jq_divmod(q, n, inv_n, &d, &i2);
        const float* m = Ms + d * MS;
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc = fmaf(m[i * n + i2], m[i2 * n + i], acc);
        p1[q] = m[i2 * n + i2];
        p2[q] = acc;
      }
      __syncthreads();
      for (int d = tid; d < db; d += nt) {
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < n; ++i) {
          s1 += p1[d * n + i];
          s2 += p2[d * n + i];
        }
        if (kk < K) {
          det_grad[(w * D + d0 + d) * K + kk] = s1;
          t2[d] += s2;
        } else {
          trL[d] = s1;
        }
      }
    }
    __syncthreads();
    for (int d = tid; d < db; d += nt) det_lap[w * D + d0 + d] = trL[d] - t2[d];
    return;
  }
  // traces, KC derivative slabs at a time.  slab kk <-> component c = 1 + k0 + kk (k0 + kk == K is the Laplacian row)
  const int nb = (n + 3) / 4, tiles = nb * nb;
  float* invT = sm + (((p2 + KC * DB * n) - sm + 3) & ~3);   // [DB][n][NP4]: invT[d][j][i] = inv[d][i][j], 16-byte aligned
  for (int q = tid; q < db * nn; q += nt) {
    int d, rem, i, j;
    jq_divmod(q, nn, inv_nn, &d, &rem);
    jq_divmod(rem, n, inv_n, &i, &j);
    invT[(d * n + j) * NP4 + i] = inv[q];
  }
  __syncthreads();
  for (int k0 = 0; k0 < KT; k0 += KC) {
    const int kc = (KT - k0 < KC) ? KT - k0 : KC;
    // Slab staging.  r2 profile (benzene): 40 % of the kernel's stall samples sat on the shared store of the plain
    // load -> store loop, each element paying the global latency on its own (0.3 TB/s).  Device build: a warp takes
    // (slab, row j) pairs -- db * n contiguous floats each -- and issues 4-byte cp.async copies, all of a thread's
    // elements in flight together; one wait per batch.
#ifndef JAQMC_HOST_EMU
    {
      const int warp_ = tid >> 5, lane_ = tid & 31, nw_ = nt >> 5;
      const unsigned jc0 = (unsigned)__cvta_generic_to_shared(Jc);
      for (int pr = warp_; pr < kc * n; pr += nw_) {
        const int kk = pr / n, j = pr - kk * n;
        const float* src = ow + ((long long)j * C + (1 + k0 + kk)) * DN;
        for (int r = lane_; r < db * n; r += 32) {
          const int d = r / n, i = r - d * n;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(jc0 + 4u * (unsigned)(((kk * DB + d) * n + j) * NP4 + i)),
                       "l"(src + r));
        }
      }
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
#else
    for (int q = tid; q < kc * n * db * n; q += nt) {
      int t, r, kk, j, d, i;
      jq_divmod(q, db * n, inv_dbn, &t, &r);
      jq_divmod(t, n, inv_n, &kk, &j);
      jq_divmod(r, n, inv_n, &d, &i);
      Jc[((kk * DB + d) * n + j) * NP4 + i] = ow[((long long)j * C + (1 + k0 + kk)) * DN + r];
    }
#endif
    __syncthreads();
    // M = inv . J in 4 x 4 register tiles: per contraction index j a tile reads one float4 of the j-major (transposed)
    // inverse and one of the slab -- rows padded to a multiple of 4 so that both are single 16-byte shared loads (r2:
    // eight scalar loads per 16 multiply-adds before; the kernel is bound by the shared-memory pipe).  The padding
    // columns are never initialised: they only reach accumulators that are not stored.
    for (int q = tid; q < kc * db * tiles; q += nt) {
      const int t = q % tiles, dk = q / tiles;
      const int d = dk % db, kk = dk / db;
      const int i0 = 4 * (t / nb), c0 = 4 * (t % nb);
      const float* tp = invT + (size_t)d * n * NP4 + i0;
      const float* jb = Jc + (size_t)(kk * DB + d) * n * NP4 + c0;
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
      // (r2: four contraction steps per trip with all eight loads hoisted was tried after the profile showed the first
      // FMA behind the loads as the top stall site: 15.3 -> 15.7 ms, i.e. slower; reverted)
#pragma unroll 2
      for (int j = 0; j < n; ++j) {
        const float4 x4 = *reinterpret_cast<const float4*>(tp + j * NP4);
        const float4 y4 = *reinterpret_cast<const float4*>(jb + j * NP4);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xv[a], yv[b], acc[a][b]);
      }
      float* mo = Mc + (kk * DB + d) * nn;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (i0 + a < n && c0 + b < n) mo[(i0 + a) * n + c0 + b] = acc[a][b];
    }
    __syncthreads();
    for (int q = tid; q < kc * db * n; q += nt) {
      int t, i, kk, d;
      jq_divmod(q, n, inv_n, &t, &i);
      jq_divmod(t, db, inv_db, &kk, &d);
      const float* m = Mc + (kk * DB + d) * nn;
      float acc = 0.f;
      for (int i2 = 0; i2 < n; ++i2) acc = fmaf(m[i * n + i2], m[i2 * n + i], acc);
      p1[(kk * DB + d) * n + i] = m[i * n + i];
      p2[(kk * DB + d) * n + i] = acc;
    }
    __syncthreads();
    for (int q = tid; q < kc * db; q += nt) {
      int d = q % db, kk = q / db;
      float s1 = 0.f, s2 = 0.f;
      for (int i = 0; i < n; ++i) {
        s1 += p1[(kk * DB + d) * n + i];
        s2 += p2[(kk * DB + d) * n + i];
      }
      int k = k0 + kk;
      if (k < K) {
        det_grad[(w * D + d0 + d) * K + k] = s1;
        p2[(kk * DB + d) * n] = s2;  // parked for the serial accumulation below
      } else {
        trL[d] = s1;
      }
    }
    __syncthreads();
    for (int d = tid; d < db; d += nt) {
      float acc = t2[d];
      for (int kk = 0; kk < kc; ++kk)
        if (k0 + kk < K) acc += p2[(kk * DB + d) * n];
      t2[d] = acc;
    }
    __syncthreads();
  }
  for (int d = tid; d < db; d += nt) det_lap[w * D + d0 + d] = trL[d] - t2[d];
}

#ifndef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// Value-only slogdet (the MH sampling path: 11 of the 12 forward passes of a VMC iteration): one WARP per matrix.
// Lane r keeps row r in registers; LU with partial pivoting without moving rows: at step p the pivot is the largest
// |a[r][p]| among the rows not used yet (warp max + ballot), its row is broadcast by shuffles and eliminated from the
// remaining rows -- the same pivots and the same fmaf sequence as the row-swapping elimination of k_logdet (and of
// LAPACK getrf behind jnp.linalg.slogdet, wavefunction/output/logdet.py:65), so the two kernels agree on sign and
// log|det|.  sign = parity(step -> row permutation) * prod sign(pivot); log|det| = log prod |pivot| in double.
// ------------------------------------------------------------------------------------------------
template <int NMAX>
__global__ void __launch_bounds__(256) k_logdet_value_warp(const float* __restrict__ orb, const float* __restrict__ el,
                                                          const float* __restrict__ atoms, JqEnvelopeArgs env, int has_env,
                                                          JqSpins sp, int A, int n, int D, long long M,
                                                          float* __restrict__ det_sign, float* __restrict__ det_logabs) {
  // r2: NMAX lanes per matrix (32 / NMAX matrices per warp, sub-group shuffles), and the isotropic envelope
  // (output/envelope.py:131-135) applied to the row as it is loaded when has_env -- k_orb_envelope_value's arithmetic,
  // without its pass over the orbital buffer (the two kernels were 20 % of a sampling forward pass).
  constexpr int LPM = NMAX, MPW = 32 / NMAX;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPM, hl = lane - sub * LPM;
  // all collectives below use the FULL mask with a width: sub-group masks (tried first) compile to MATCH / branch /
  // BSYNC sequences that made the LU 13 % slower than the one-matrix-per-warp kernel it replaced (ncu, r2)
  const unsigned full = 0xffffffffu;
  const unsigned smask = (LPM == 32) ? 0xffffffffu : ((1u << LPM) - 1u);
  const long long m0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * MPW + sub;   // (walker, determinant)
  const bool on = m0 < M;
  const long long m = on ? m0 : M - 1;
  const long long w = m / D;
  const int d = (int)(m - w * D);
  const int DN = D * n;
  float a[NMAX];
  bool used = hl >= n || !on;
  const int j = hl < n ? hl : 0;   // electron = row
  {
    const float* row = orb + (w * n + j) * (long long)DN + d * n;
#pragma unroll
    for (int c = 0; c < NMAX; ++c) a[c] = (c < n) ? row[c] : 0.f;
  }
  if (has_env) {
    const float* e3 = el + (w * n + j) * 3;
    const float px = e3[0], py = e3[1], pz = e3[2];
    const int ch = (env.pi[1] != nullptr && sp.chan_of(j) == 1) ? 1 : 0;
    const float* pi = env.pi[ch];
    const float* sg = env.sigma[ch];
    float r[ENVV_A];
#pragma unroll
    for (int I = 0; I < ENVV_A; ++I) {
      const float dx = px - ((I < A) ? atoms[3 * I] : 0.f), dy = py - ((I < A) ? atoms[3 * I + 1] : 0.f),
                  dz = pz - ((I < A) ? atoms[3 * I + 2] : 0.f);
      r[I] = sqrtf(dx * dx + dy * dy + dz * dz);
    }
#pragma unroll
    for (int c = 0; c < NMAX; ++c)
      if (c < n) {
        float e = 0.f;
#pragma unroll
        for (int I = 0; I < ENVV_A; ++I)
          if (I < A) {
            float sv = sg[(c * A + I) * D + d];
            if (env.type == 1) sv = fabsf(sv);
            e += pi[(c * A + I) * D + d] * expf(-sv * r[I]);
          }
        a[c] *= e;
      }
  }
  int step_of = 0;
  float sgn = 1.0f;
  // |det| = prod |pivot| kept as (mantissa product in double, binary exponent sum): one double log per matrix
  double mant = 1.0;
  int expo = 0;
#pragma unroll
  for (int p = 0; p < NMAX; ++p) {
    if (p < n) {
      const unsigned key = used ? 0u : __float_as_uint(fabsf(a[p]));
      unsigned mx = key;
#pragma unroll
      for (int o = LPM / 2; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(full, mx, o, LPM));
      const unsigned cand = (__ballot_sync(full, !used && key == mx) >> (sub * LPM)) & smask;
      const int pl = cand ? __ffs(cand) - 1 : 0;    // first row holding the largest magnitude
      const float pv = __shfl_sync(full, a[p], pl, LPM);
      if (pv < 0.f) sgn = -sgn;
      if (pv == 0.f) sgn = 0.f;
      {
        int e;
        mant *= (double)frexpf(fabsf(pv), &e);   // zero pivot: mantissa 0 -> log 0 = -inf, like slogdet
        expo += e;
      }
      const float pinv = 1.0f / pv;
      if (hl == pl && !used) {
        used = true;
        step_of = p;
      }
      const float f = used ? 0.f : a[p] * pinv;
#pragma unroll
      for (int c = p + 1; c < NMAX; ++c)
        if (c < n) {
          const float pc = __shfl_sync(full, a[c], pl, LPM);
          a[c] = fmaf(-f, pc, a[c]);
        }
    }
  }
  // parity of the permutation row -> step: inversions counted per lane, summed over the sub-group
  int inv_count = 0;
  for (int i = 0; i < n; ++i) {
    const int si = __shfl_sync(full, step_of, i, LPM);
    if (i < hl && hl < n && si > step_of) ++inv_count;
  }
  int total = inv_count;
#pragma unroll
  for (int o = LPM / 2; o > 0; o >>= 1) total += __shfl_xor_sync(full, total, o, LPM);
  if (on && hl == 0) {
    det_sign[m] = (total & 1) ? -sgn : sgn;
    det_logabs[m] = (float)(log(mant) + (double)expo * 0.69314718055994530942);
  }
}
#endif

#ifndef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// slogdet + forward-Laplacian rule for TINY matrices (n <= 4: Li, LiH ...; r2): one THREAD per determinant, everything
// in registers.  The half-warp kernel above keeps 16 lanes per determinant and ran at 252 us for the 65536 3 x 3
// matrices of the Li configuration (3 of 16 lanes busy).  Gauss-Jordan with partial pivoting (row swaps), sign =
// (-1)^swaps prod sign(pivot), log|det| = log prod |pivot| in double; then per derivative slab M = A^-1 dA_c,
// tr M and tr M^2 = sum_ij M_ij M_ji.  The 16 determinants of a walker are adjacent threads: a row (j, c) of the slab
// is D * n contiguous floats.
// ------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) k_logdet_tiny(const float* __restrict__ orb, int D, int C, long long M,
                                                     float* __restrict__ det_sign, float* __restrict__ det_logabs,
                                                     float* __restrict__ det_grad, float* __restrict__ det_lap) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (walker, determinant)
  if (m >= M) return;
  const long long w = m / D;
  const int d = (int)(m - w * D);
  const int DN = D * N, K = C - 2, KT = C - 1;
  const float* ow = orb + (w * N) * (long long)C * DN + d * N;   // (j, c, i) at ow[(j * C + c) * DN + i]
  float a[N][N], b[N][N];
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      a[j][i] = ow[(long long)j * C * DN + i];
      b[j][i] = (i == j) ? 1.0f : 0.f;
    }
  float sg = 1.0f;
  double mant = 1.0;
  int expo = 0;
#pragma unroll
  for (int p = 0; p < N; ++p) {
    // pivot: largest |a[r][p]|, r >= p (first one on ties); rows are swapped with predicated moves (N is tiny)
    int r = p;
    float best = fabsf(a[p][p]);
#pragma unroll
    for (int q = p + 1; q < N; ++q)
      if (fabsf(a[q][p]) > best) {
        best = fabsf(a[q][p]);
        r = q;
      }
#pragma unroll
    for (int q = p + 1; q < N; ++q)
      if (q == r) {
#pragma unroll
        for (int c = 0; c < N; ++c) {
          float t = a[p][c]; a[p][c] = a[q][c]; a[q][c] = t;
          t = b[p][c]; b[p][c] = b[q][c]; b[q][c] = t;
        }
        sg = -sg;
      }
    const float pv = a[p][p];
    if (pv < 0.f) sg = -sg;
    if (pv == 0.f) sg = 0.f;
    {
      int e;
      mant *= (double)frexpf(fabsf(pv), &e);
      expo += e;
    }
    const float pinv = 1.0f / pv;
#pragma unroll
    for (int c = 0; c < N; ++c) {
      a[p][c] *= pinv;
      b[p][c] *= pinv;
    }
#pragma unroll
    for (int q = 0; q < N; ++q)
      if (q != p) {
        const float f = a[q][p];
#pragma unroll
        for (int c = 0; c < N; ++c) {
          a[q][c] = fmaf(-f, a[p][c], a[q][c]);
          b[q][c] = fmaf(-f, b[p][c], b[q][c]);
        }
      }
  }
  det_sign[m] = sg;
  det_logabs[m] = (float)(log(mant) + (double)expo * 0.69314718055994530942);
  // b = A^-1 (A[j][i]: row = electron j, column = orbital i): b[i][j]
  float t2 = 0.f, trl = 0.f;
  float* gout = det_grad + m * (long long)K;
  for (int kk = 0; kk < KT; ++kk) {
    float da[N][N], mm[N][N];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = 0; i < N; ++i) da[j][i] = ow[((long long)j * C + 1 + kk) * DN + i];
    float tr = 0.f, s2 = 0.f;
#pragma unroll
    for (int x = 0; x < N; ++x)
#pragma unroll
      for (int y = 0; y < N; ++y) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) acc = fmaf(b[x][j], da[j][y], acc);
        mm[x][y] = acc;
        if (x == y) tr += acc;
      }
#pragma unroll
    for (int x = 0; x < N; ++x)
#pragma unroll
      for (int y = 0; y < N; ++y) s2 = fmaf(mm[x][y], mm[y][x], s2);
    if (kk < K) {
      gout[kk] = tr;
      t2 += s2;
    } else {
      trl = tr;
    }
  }
  det_lap[m] = trl - t2;
}
#endif

// Value-only slogdet of orb [W][n][D*n] (n <= 32), optionally times the isotropic envelope of (electrons, atoms) first.
bool jq_logdet_value_env_eligible(int n, int A, int env_type) {
#ifdef JAQMC_HOST_EMU
  return false;
#else
  static const bool off = getenv("JAQMC_B200_UNFUSED_ENV_LOGDET") != nullptr;   // A/B switch
  return !off && n <= 32 && A <= ENVV_A && (env_type == 0 || env_type == 1);
#endif
}

int jq_launch_logdet_value_env(const float* orb, const float* electrons, const float* atoms, const JqEnvelopeArgs& env,
                               int has_env, int W, JqSpins sp, int A, int D, float* det_sign, float* det_logabs,
                               cudaStream_t st) {
#ifdef JAQMC_HOST_EMU
  return JQ_ERR_UNSUPPORTED;
#else
  const int n = sp.n();
  const long long M = (long long)W * D;
  if (M <= 0) return JQ_OK;
  JQ_REQUIRE(n <= 32, JQ_ERR_UNSUPPORTED, "logdet (value): n=%d", n);
  const int lpm = n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32;
  const dim3 grid((unsigned)jq_cdiv(M, 8 * (32 / lpm))), block(256);
  jq_prof_work((double)M * 0.67 * n * n * n, 4.0 * (double)M * n * n);
  if (lpm == 4) JQ_LAUNCH(k_logdet_value_warp<4>, grid, block, 0, st, orb, electrons, atoms, env, has_env, sp, A, n, D, M, det_sign, det_logabs);
  else if (lpm == 8) JQ_LAUNCH(k_logdet_value_warp<8>, grid, block, 0, st, orb, electrons, atoms, env, has_env, sp, A, n, D, M, det_sign, det_logabs);
  else if (lpm == 16) JQ_LAUNCH(k_logdet_value_warp<16>, grid, block, 0, st, orb, electrons, atoms, env, has_env, sp, A, n, D, M, det_sign, det_logabs);
  else JQ_LAUNCH(k_logdet_value_warp<32>, grid, block, 0, st, orb, electrons, atoms, env, has_env, sp, A, n, D, M, det_sign, det_logabs);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
#endif
}

int jq_launch_logdet(const float* orb, int W, int n, int D, int track, float* det_sign, float* det_logabs,
                     float* det_grad, float* det_lap, cudaStream_t st) {
  if ((long long)W * D <= 0) return JQ_OK;
#ifndef JAQMC_HOST_EMU
  if (!track && n <= 32) {
    JqEnvelopeArgs no_env;
    memset(&no_env, 0, sizeof(no_env));
    return jq_launch_logdet_value_env(orb, nullptr, nullptr, no_env, 0, W, JqSpins{n, 0}, 0, D, det_sign, det_logabs, st);
  }
#endif
  const int C = track ? 3 * n + 2 : 1;
  const int KT = C > 1 ? C - 1 : 0;
  const size_t nn = (size_t)n * n;
#ifndef JAQMC_HOST_EMU
  // Opt-in: traces on the tensor cores (mma.sync 3xTF32), 20 % faster than the FP32 kernel but measured 3-10x its error
  // on ill-conditioned determinants (22 mantissa bits per operand, truncating accumulation): LapNet-N2 parity no
  // longer meets the north-star tolerance with it, so the FP32 kernel stays the default.
  static const bool tc_traces = getenv("JAQMC_B200_LOGDET_TC") != nullptr;
  if (track && n <= LD_NP && tc_traces) {
    // warp per determinant, traces on the tensor cores (mma.sync 3xTF32)
    const long long MT = (long long)W * D;
    jq_prof_work((double)MT * (2.0 * n * n * n * ((C - 1) * 2 + 1)), 4.0 * (double)MT * C * n * n);
    if (n % 2 == 0 && (reinterpret_cast<uintptr_t>(orb) & 7) == 0)
      JQ_LAUNCH(k_logdet_mma<true>, dim3((unsigned)jq_cdiv(MT, 8)), dim3(256), 0, st, orb, n, D, C, MT, det_sign,
                det_logabs, det_grad, det_lap);
    else
      JQ_LAUNCH(k_logdet_mma<false>, dim3((unsigned)jq_cdiv(MT, 8)), dim3(256), 0, st, orb, n, D, C, MT, det_sign,
                det_logabs, det_grad, det_lap);
    JQ_CHECK_LAUNCH();
    return JQ_OK;
  }
  static const bool no_tiny = getenv("JAQMC_B200_LOGDET_NO_TINY") != nullptr;   // A/B switch
  if (track && n <= 4 && !no_tiny) {
    // thread per determinant (n <= 4)
    const long long MT = (long long)W * D;
    jq_prof_work((double)MT * (2.0 * n * n * n * ((C - 1) * 2 + 1)), 4.0 * (double)MT * C * n * n);
    const dim3 grid((unsigned)jq_cdiv(MT, 128)), block(128);
    if (n == 1) JQ_LAUNCH(k_logdet_tiny<1>, grid, block, 0, st, orb, D, C, MT, det_sign, det_logabs, det_grad, det_lap);
    else if (n == 2) JQ_LAUNCH(k_logdet_tiny<2>, grid, block, 0, st, orb, D, C, MT, det_sign, det_logabs, det_grad, det_lap);
    else if (n == 3) JQ_LAUNCH(k_logdet_tiny<3>, grid, block, 0, st, orb, D, C, MT, det_sign, det_logabs, det_grad, det_lap);
    else JQ_LAUNCH(k_logdet_tiny<4>, grid, block, 0, st, orb, D, C, MT, det_sign, det_logabs, det_grad, det_lap);
    JQ_CHECK_LAUNCH();
    return JQ_OK;
  }
  if (track && n <= LD_NP) {
    // half-warp per determinant: 16 determinants per block; NS derivative slabs per pass over A^-1
    const int DB = D < 16 ? D : 16;
    const size_t MS = nn + (size_t)((((int)n - (int)nn) % 32 + 32) % 32);
    static const int ns_env = getenv("JAQMC_B200_LOGDET_SLABS") ? atoi(getenv("JAQMC_B200_LOGDET_SLABS")) : 2;
    const int NS = ns_env == 1 ? 1 : (ns_env == 3 ? 3 : 2);
    // staged slabs (cp.async, per-warp double buffer): n and D even, 16-byte aligned orbital buffer, default NS
    static const bool sync_loads = getenv("JAQMC_B200_LOGDET_SYNC_LOADS") != nullptr;   // A/B switch
    const size_t smem_sync = sizeof(float) * ((size_t)DB * (n * LD_NP + 4) + (size_t)DB * MS * NS) + 64;
    const size_t ring_bytes = sizeof(float) * (size_t)8 * 2 * NS * n * 2 * n;   // 8 warps x double buffer x NS slabs x [n][2n]
    // (two blocks per SM need <= 112 KB each: 16 electrons x 16 determinants would take 117 KB and keeps the synchronous loads)
    const bool staged = NS == 2 && !sync_loads && n >= 6 && (n % 2) == 0 && (D % 2) == 0 &&
                        (reinterpret_cast<uintptr_t>(orb) & 15) == 0 && smem_sync + ring_bytes <= 112 * 1024;
    const size_t smem = smem_sync + (staged ? ring_bytes : 0);
    void (*kern)(const float*, int, int, int, int, float*, float*, float*, float*) =
        staged ? k_logdet_small<2, true>
               : NS == 1 ? k_logdet_small<1, false> : NS == 3 ? k_logdet_small<3, false> : k_logdet_small<2, false>;   // for the attribute call only
    if (smem > 48 * 1024) {
      static JqPerDeviceFlag attr_set[5];
      const int dev = jq_current_device();
      const int slot = staged ? 4 : NS;
      if (!attr_set[slot].done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "logdet: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set[slot].done[dev] = true;
      }
    }
    const long long blocks = (long long)W * ((D + DB - 1) / DB);
    jq_prof_work((double)W * D * (2.0 * n * n * n * ((C - 1) * 2 + 1)), 4.0 * (double)W * D * C * n * n);
    // launched through named pointers: JQ_LAUNCH labels the profile entry with its first argument
    auto k_logdet_small_staged = k_logdet_small<2, true>;
    auto k_logdet_small_ns1 = k_logdet_small<1, false>;
    auto k_logdet_small_ns2 = k_logdet_small<2, false>;
    auto k_logdet_small_ns3 = k_logdet_small<3, false>;
    if (staged) JQ_LAUNCH(k_logdet_small_staged, dim3((unsigned)blocks), dim3(256), smem, st, orb, n, D, C, DB, det_sign, det_logabs, det_grad, det_lap);
    else if (NS == 1) JQ_LAUNCH(k_logdet_small_ns1, dim3((unsigned)blocks), dim3(256), smem, st, orb, n, D, C, DB, det_sign, det_logabs, det_grad, det_lap);
    else if (NS == 3) JQ_LAUNCH(k_logdet_small_ns3, dim3((unsigned)blocks), dim3(256), smem, st, orb, n, D, C, DB, det_sign, det_logabs, det_grad, det_lap);
    else JQ_LAUNCH(k_logdet_small_ns2, dim3((unsigned)blocks), dim3(256), smem, st, orb, n, D, C, DB, det_sign, det_logabs, det_grad, det_lap);
    JQ_CHECK_LAUNCH();
    return JQ_OK;
  }
#endif
  // DB matrices per block: about 16 KB of inverses; KC slabs so that the block stays under ~100 KB
  int DB = (int)(16384 / (nn * 4));
  if (DB < 1) DB = 1;
  if (DB > D) DB = D;
  int KC = 1;
  auto smem_for = [&](int db, int kc) {
    if (track && kc == 0)  // small path: inv | colp piv scalars | padded inverses | one product slab | p1 p2
      return (size_t)8 * db + sizeof(float) * ((size_t)db * nn + 2 * (size_t)db * n + 4 * (size_t)db +
                                              (size_t)db * (n * LD_NP + 4) + (size_t)db * (nn + 32) + 2 * (size_t)db * n) + 32;
    const size_t np4 = (size_t)((n + 3) & ~3);   // padded rows of the slab and transposed-inverse tiles
    return (size_t)8 * db + sizeof(float) * ((size_t)db * nn + 2 * (size_t)db * n + 4 * (size_t)db +
                                            (track ? (size_t)kc * db * (n * np4 + nn) + (size_t)kc * db * n * 2 +
                                                         (size_t)db * n * np4 + 8 : 0)) + 32;
  };
  if (track && n <= LD_NP) {
    // small-matrix path: all determinants of a walker in one block when they fit; the slab area [2*KC][DB][nn] must
    // hold the row-padded inverses [DB][n][16] and one product slab [DB][nn]
    DB = D;
    KC = 0;  // marks the small path for the size computation below
    while (DB > 1 && smem_for(DB, 0) > 64 * 1024) DB = (DB + 1) / 2;
  } else if (track) {
    while (KC < KT && smem_for(DB, KC + 1) <= 100 * 1024) ++KC;
    while (DB > 1 && smem_for(DB, KC) > 200 * 1024) --DB;
  }
  size_t smem = smem_for(DB, KC);
  if (KC == 0) KC = 1;
  JQ_REQUIRE(smem <= 200 * 1024, JQ_ERR_UNSUPPORTED, "logdet: %d electrons need %zu bytes of shared memory", n, smem);
#ifndef JAQMC_HOST_EMU
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_logdet, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "logdet: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
#endif
  const long long blocks = (long long)W * ((D + DB - 1) / DB);
  jq_prof_work((double)W * D * (2.0 * n * n * n * (track ? (C - 1) * 2 + 1 : 0.34)), 4.0 * (double)W * D * C * n * n);
  JQ_LAUNCH(k_logdet, dim3((unsigned)blocks), dim3(256), smem, st, orb, n, D, C, DB, KC, det_sign, det_logabs,
            det_grad, det_lap);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// log-sum-exp over determinants (+ optional additive Jastrow term `extra` [W][C]).  One item per walker.
// ------------------------------------------------------------------------------------------------
__global__ void k_logdet_combine(const float* __restrict__ det_sign, const float* __restrict__ det_logabs,
                                 const float* __restrict__ det_grad, const float* __restrict__ det_lap, int W, int n,
                                 int D, int track, const float* __restrict__ extra, float* __restrict__ logpsi,
                                 float* __restrict__ sign, float* __restrict__ grad, float* __restrict__ lap,
                                 float* __restrict__ e_kin) {
  const int K = 3 * n, C = K + 2;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W;
       w += (long long)gridDim.x * blockDim.x) {
    const float* s = det_sign + w * D;
    const float* l = det_logabs + w * D;
    float lmax = l[0];
    for (int d = 1; d < D; ++d) lmax = fmaxf(lmax, l[d]);
    float S = 0.f;
    for (int d = 0; d < D; ++d) S += s[d] * expf(l[d] - lmax);
    float sg = (S > 0.f) ? 1.f : ((S < 0.f) ? -1.f : 0.f);
    float lp = logf(fabsf(S)) + lmax;
    const float* ex = extra ? extra + w * (track ? C : 1) : nullptr;
    if (ex) lp += ex[0];
    logpsi[w] = lp;
    sign[w] = sg;
    if (!track) continue;
    const float* g = det_grad + w * D * K;
    float Sinv = 1.0f / S;
    float acc_l = 0.f;
    for (int d = 0; d < D; ++d) {
      float wd = s[d] * expf(l[d] - lmax) * Sinv;
      float g2 = 0.f;
      for (int k = 0; k < K; ++k) g2 = fmaf(g[d * K + k], g[d * K + k], g2);
      acc_l = fmaf(wd, det_lap[w * D + d] + g2, acc_l);
    }
    float gg_det = 0.f, gg = 0.f;
    for (int k = 0; k < K; ++k) {
      float gk = 0.f;
      for (int d = 0; d < D; ++d) gk = fmaf(s[d] * expf(l[d] - lmax) * Sinv, g[d * K + k], gk);
      gg_det = fmaf(gk, gk, gg_det);
      if (ex) gk += ex[1 + k];
      grad[w * K + k] = gk;
      gg = fmaf(gk, gk, gg);
    }
    float lp_l = acc_l - gg_det;
    if (ex) lp_l += ex[C - 1];
    lap[w] = lp_l;
    e_kin[w] = -0.5f * (lp_l + gg);
  }
}

#ifndef JAQMC_HOST_EMU
// Same combination with one WARP per walker (k_logdet_combine gives a walker to one thread: 16 x 42 serial loads, which
// is the whole cost of the kernel once a GPU holds only a few hundred walkers).  Lane d holds determinant d's weight;
// the gradient components are spread over the lanes.  D <= 32.
__global__ void __launch_bounds__(256) k_logdet_combine_warp(const float* __restrict__ det_sign,
                                                            const float* __restrict__ det_logabs,
                                                            const float* __restrict__ det_grad,
                                                            const float* __restrict__ det_lap, int W, int n, int D, int track,
                                                            const float* __restrict__ extra, float* __restrict__ logpsi,
                                                            float* __restrict__ sign, float* __restrict__ grad,
                                                            float* __restrict__ lap, float* __restrict__ e_kin) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int K = 3 * n, C = K + 2;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < W;
       w += (long long)gridDim.x * (blockDim.x >> 5)) {
    const bool on = lane < D;
    const float sd = on ? det_sign[w * D + lane] : 0.f;
    const float ld = on ? det_logabs[w * D + lane] : -INFINITY;
    float lmax = ld;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(full, lmax, o));
    // the sums over determinants run in determinant order on every lane (same order as the scalar kernel)
    const float term = on ? sd * expf(ld - lmax) : 0.f;
    float S = 0.f;
    for (int d = 0; d < D; ++d) S += __shfl_sync(full, term, d);
    const float sg = (S > 0.f) ? 1.f : ((S < 0.f) ? -1.f : 0.f);
    float lp = logf(fabsf(S)) + lmax;
    const float* ex = extra ? extra + w * (track ? C : 1) : nullptr;
    if (ex) lp += ex[0];
    if (lane == 0) {
      logpsi[w] = lp;
      sign[w] = sg;
    }
    if (!track) continue;
    const float* g = det_grad + w * D * K;
    const float wd = term * (1.0f / S);   // lane d: weight of determinant d
    // per-determinant |grad|^2, lane d keeps determinant d's
    float g2_mine = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {   // (unrolled: the loads of four determinants in flight; the kernel is pure latency)
      float part = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float v = g[d * K + k];
        part = fmaf(v, v, part);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(full, part, o);
      if (lane == d) g2_mine = part;
    }
    float acc_term = on ? wd * (det_lap[w * D + lane] + g2_mine) : 0.f;
    float acc_l = 0.f;
    for (int d = 0; d < D; ++d) acc_l += __shfl_sync(full, acc_term, d);
    float gg_det = 0.f, gg = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      float gk = 0.f;
#pragma unroll 4
      for (int d = 0; d < D; ++d) {
        const float wdd = __shfl_sync(full, wd, d);
        if (k < K) gk = fmaf(wdd, g[d * K + k], gk);
      }
      if (k < K) {
        gg_det = fmaf(gk, gk, gg_det);
        if (ex) gk += ex[1 + k];
        grad[w * K + k] = gk;
        gg = fmaf(gk, gk, gg);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gg_det += __shfl_xor_sync(full, gg_det, o);
      gg += __shfl_xor_sync(full, gg, o);
    }
    if (lane == 0) {
      float lp_l = acc_l - gg_det;
      if (ex) lp_l += ex[C - 1];
      lap[w] = lp_l;
      e_kin[w] = -0.5f * (lp_l + gg);
    }
  }
}
#endif

int jq_launch_logdet_combine(const float* det_sign, const float* det_logabs, const float* det_grad,
                             const float* det_lap, int W, int n, int D, int track, const float* extra_logpsi,
                             float* logpsi, float* sign, float* grad, float* lap, float* e_kin, cudaStream_t st) {
  if (W <= 0) return JQ_OK;
#ifndef JAQMC_HOST_EMU
  if (D <= 32) {
    JQ_LAUNCH(k_logdet_combine_warp, dim3(jq_cdiv(W, 8)), dim3(256), 0, st, det_sign, det_logabs, det_grad, det_lap, W, n,
              D, track, extra_logpsi, logpsi, sign, grad, lap, e_kin);
    JQ_CHECK_LAUNCH();
    return JQ_OK;
  }
#endif
  JQ_LAUNCH(k_logdet_combine, dim3(jq_cdiv(W, 64)), dim3(64), 0, st, det_sign, det_logabs, det_grad, det_lap, W, n, D,
            track, extra_logpsi, logpsi, sign, grad, lap, e_kin);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}
