// Envelope product, multi-determinant slogdet with its forward-Laplacian rule, and the log-sum-exp
// combination that yields log|psi|, sign, grad log psi, laplacian and the kinetic energy.
//
// Reference semantics:
//   * envelope  E_ik = sum_I pi_kI exp(-|sigma_kI| r_iI)          wavefunction/output/envelope.py:98-140
//   * orbitals * envelope (product rule, Local1 x dense)           app/molecule/wavefunction/ferminet.py:90-93,
//                                                                  laplacian/primitives/core.py:450-501
//   * slogdet rule ld_J[k] = tr(A^-1 dA_k), ld_L = tr(A^-1 A_L) - sum_k tr((A^-1 dA_k)^2)
//                                                                  laplacian/primitives/slogdet.py:46-72
//   * log-sum-exp over determinants with max shift                 wavefunction/output/logdet.py:65-79
//   * E_kin = -1/2 (lap + |grad|^2)                                estimator/kinetic/_common.py:61-73
#include "aug.cuh"

// ------------------------------------------------------------------------------------------------
// orb[w][j][c][d*n+i] *= envelope(j, i, d)   (in place, product rule).  One item per (w, j, d, i).
// ------------------------------------------------------------------------------------------------
__global__ void k_orb_envelope(float* __restrict__ orb, const float* __restrict__ el, const float* __restrict__ atoms,
                               JqEnvelopeArgs env, long long items, JqSpins sp, int A, int D, int track) {
  const int n = sp.n();
  const int DN = D * n;
  const int C = track ? 3 * n + 2 : 1;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    int col = (int)(it % DN);
    long long g = it / DN;  // (w, j)
    int j = (int)(g % n);
    int d = col / n, i = col % n;
    int ch = (env.pi[1] != nullptr) ? sp.chan_of(j) : 0;
    const float* pi = env.pi[ch];
    const float* sg = env.sigma[ch];
    const float* e = el + g * 3;
    float ex = 0.f, ej[3] = {0.f, 0.f, 0.f}, elap = 0.f;
    for (int I = 0; I < A; ++I) {
      float dx = e[0] - atoms[I * 3], dy = e[1] - atoms[I * 3 + 1], dz = e[2] - atoms[I * 3 + 2];
      float r = sqrtf(dx * dx + dy * dy + dz * dz);
      float s = sg[(i * A + I) * D + d];
      if (env.type == 1) s = fabsf(s);
      float t = pi[(i * A + I) * D + d] * expf(-s * r);
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

int jq_launch_orb_envelope(float* orb, const float* electrons, const float* atoms, const JqEnvelopeArgs& env,
                           int W, JqSpins sp, int A, int D, int track, cudaStream_t st) {
  if (env.type == 2) return JQ_OK;  // null envelope: ones
  long long items = (long long)W * sp.n() * D * sp.n();
  if (items <= 0) return JQ_OK;
  int grid = jq_cdiv(items, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  jq_prof_work(0.0, 8.0 * (double)items * (track ? 3 * sp.n() + 2 : 1));
  JQ_LAUNCH(k_orb_envelope, dim3(grid), dim3(256), 0, st, orb, electrons, atoms, env, items, sp, A, D, track);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// slogdet + forward-Laplacian rule.  One block per (walker, determinant).
// Shared memory: inv[n*n] | colp[n] | piv[n] | scal[8] | t2[KT] | Jc[KC*n*n] | Mc[KC*n*n] | p1[KC*n] | p2[KC*n]
// In-place Gauss-Jordan with partial (row) pivoting; sign = prod sign(pivot) * (-1)^swaps,
// log|det| = sum log|pivot| (accumulated in double by the single pivot-search item).
// ------------------------------------------------------------------------------------------------
__global__ void k_logdet(const float* __restrict__ orb, int n, int D, int C, int KC, float* __restrict__ det_sign,
                         float* __restrict__ det_logabs, float* __restrict__ det_grad, float* __restrict__ det_lap) {
  JQ_DYN_SMEM(float, sm);
  const int nn = n * n;
  const int K = C - 2;           // Jacobian columns (C == 1: value only)
  const int KT = (C > 1) ? C - 1 : 0;  // J columns + the Laplacian row
  float* inv = sm;
  float* colp = inv + nn;
  int* piv = reinterpret_cast<int*>(colp + n);
  float* scal = reinterpret_cast<float*>(piv + n);  // [0] pivinv, [1] sign, [2] logabs, [3] trL
  float* t2 = scal + 8;
  float* Jc = t2 + (KT > 0 ? KT : 1);
  float* Mc = Jc + (long long)KC * nn;
  float* p1 = Mc + (long long)KC * nn;
  float* p2 = p1 + KC * n;
  const long long b = blockIdx.x;
  const long long w = b / D;
  const int d = (int)(b % D);
  const int DN = D * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool need_inv = (C > 1);

  // load A[j][i] (value rows)
  for (int q = tid; q < nn; q += nt) {
    int j = q / n, i = q % n;
    inv[q] = orb[((w * n + j) * (long long)C) * DN + d * n + i];
  }
  __syncthreads();
  double logabs = 0.0;  // only meaningful in the item that runs the pivot search (tid 0)
  float sgn = 1.0f;
  for (int p = 0; p < n; ++p) {
    if (tid == 0) {
      int r = p;
      float best = fabsf(inv[p * n + p]);
      for (int q = p + 1; q < n; ++q) {
        float v = fabsf(inv[q * n + p]);
        if (v > best) { best = v; r = q; }
      }
      piv[p] = r;
      float pv = inv[r * n + p];
      if (r != p) sgn = -sgn;
      if (pv < 0.f) sgn = -sgn;
      if (pv == 0.f) sgn = 0.f;
      logabs += log((double)fabsf(pv));
      scal[0] = 1.0f / pv;
    }
    __syncthreads();
    {
      int r = piv[p];
      if (r != p)
        for (int c = tid; c < n; c += nt) {
          float t = inv[p * n + c];
          inv[p * n + c] = inv[r * n + c];
          inv[r * n + c] = t;
        }
    }
    __syncthreads();
    float pivinv = scal[0];
    if (need_inv) {
      for (int q = tid; q < n; q += nt) colp[q] = inv[q * n + p];  // column p before it is overwritten
      __syncthreads();
      for (int c = tid; c < n; c += nt) inv[p * n + c] = ((c == p) ? 1.0f : inv[p * n + c]) * pivinv;
      __syncthreads();
      for (int q = tid; q < nn; q += nt) {
        int i = q / n, c = q % n;
        if (i == p) continue;
        float base = (c == p) ? 0.f : inv[q];
        inv[q] = fmaf(-colp[i], inv[p * n + c], base);
      }
      __syncthreads();
    } else {
      // value only: eliminate below the pivot (LU), no inverse needed
      for (int q = tid; q < n; q += nt) colp[q] = inv[q * n + p] * pivinv;
      __syncthreads();
      for (int q = tid; q < nn; q += nt) {
        int i = q / n, c = q % n;
        if (i <= p || c <= p) continue;
        inv[q] = fmaf(-colp[i], inv[p * n + c], inv[q]);
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    det_sign[b] = sgn;
    det_logabs[b] = (float)logabs;
  }
  if (!need_inv) return;
  // undo the row swaps as column swaps, in reverse order
  for (int p = n - 1; p >= 0; --p) {
    int r = piv[p];
    if (r != p) {
      for (int q = tid; q < n; q += nt) {
        float t = inv[q * n + p];
        inv[q * n + p] = inv[q * n + r];
        inv[q * n + r] = t;
      }
    }
    __syncthreads();
  }
  // traces, KC derivative slabs at a time.  slab kk <-> component c = 1 + kk (kk == K is the Laplacian row)
  for (int k0 = 0; k0 < KT; k0 += KC) {
    int kc = (KT - k0 < KC) ? KT - k0 : KC;
    for (int q = tid; q < kc * nn; q += nt) {
      int kk = q / nn, rem = q % nn;
      int j = rem / n, i = rem % n;
      Jc[q] = orb[((w * n + j) * (long long)C + (1 + k0 + kk)) * DN + d * n + i];
    }
    __syncthreads();
    for (int q = tid; q < kc * nn; q += nt) {
      int kk = q / nn, rem = q % nn;
      int i = rem / n, i2 = rem % n;
      const float* jp = Jc + kk * nn + i2;
      const float* ip = inv + i * n;
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(ip[j], jp[j * n], acc);
      Mc[q] = acc;
    }
    __syncthreads();
    for (int q = tid; q < kc * n; q += nt) {
      int kk = q / n, i = q % n;
      const float* m = Mc + kk * nn;
      float acc = 0.f;
      for (int i2 = 0; i2 < n; ++i2) acc = fmaf(m[i * n + i2], m[i2 * n + i], acc);
      p1[q] = m[i * n + i];
      p2[q] = acc;
    }
    __syncthreads();
    for (int kk = tid; kk < kc; kk += nt) {
      float s1 = 0.f, s2 = 0.f;
      for (int i = 0; i < n; ++i) {
        s1 += p1[kk * n + i];
        s2 += p2[kk * n + i];
      }
      int k = k0 + kk;
      if (k < K) {
        det_grad[b * K + k] = s1;
        t2[k] = s2;
      } else {
        scal[3] = s1;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += t2[k];
    det_lap[b] = scal[3] - s;
  }
}

static int logdet_kc(int n, int C) {
  int KT = C > 1 ? C - 1 : 0;
  if (KT == 0) return 1;
  int kc = (64 * 1024 / 4 / 2) / (n * n);
  if (kc < 1) kc = 1;
  if (kc > KT) kc = KT;
  return kc;
}

int jq_launch_logdet(const float* orb, int W, int n, int D, int track, float* det_sign, float* det_logabs,
                     float* det_grad, float* det_lap, cudaStream_t st) {
  long long blocks = (long long)W * D;
  if (blocks <= 0) return JQ_OK;
  int C = track ? 3 * n + 2 : 1;
  int KC = logdet_kc(n, C);
  int KT = C > 1 ? C - 1 : 0;
  size_t smem = sizeof(float) * ((size_t)n * n + n + n + 8 + (KT > 0 ? KT : 1) +
                                 (track ? (size_t)KC * n * n * 2 + (size_t)KC * n * 2 : 0));
  JQ_REQUIRE(smem <= 200 * 1024, JQ_ERR_UNSUPPORTED, "logdet: %d electrons need %zu bytes of shared memory", n, smem);
#ifndef JAQMC_HOST_EMU
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_logdet, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "logdet: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
#endif
  int threads = track ? 256 : 64;
  jq_prof_work((double)blocks * (2.0 * n * n * n * (track ? (C - 1) * 2 + 1 : 0.34)), 4.0 * (double)blocks * C * n * n);
  JQ_LAUNCH(k_logdet, dim3((unsigned)blocks), dim3(threads), smem, st, orb, n, D, C, KC, det_sign, det_logabs,
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

int jq_launch_logdet_combine(const float* det_sign, const float* det_logabs, const float* det_grad,
                             const float* det_lap, int W, int n, int D, int track, const float* extra_logpsi,
                             float* logpsi, float* sign, float* grad, float* lap, float* e_kin, cudaStream_t st) {
  if (W <= 0) return JQ_OK;
  JQ_LAUNCH(k_logdet_combine, dim3(jq_cdiv(W, 64)), dim3(64), 0, st, det_sign, det_logabs, det_grad, det_lap, W, n, D,
            track, extra_logpsi, logpsi, sign, grad, lap, e_kin);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}
