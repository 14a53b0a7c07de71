// Periodic (solid-state) FermiNet: complex orbitals, Bloch phases, complex multi-determinant LogDet with its
// forward-Laplacian rule, and the pipeline that composes them.
//
// Reference semantics: app/solid/wavefunction.py:91-147 (SolidWavefunction: features -> FermiLayers -> real + i imag
// orbital projections -> x envelope(r_ae) -> x exp(i k.r) -> LogDet), output/logdet.py:53-79 (complex branch: max over
// the real parts, complex log-sum-exp), laplacian/primitives/slogdet.py:46-72 (complex rule: ld = log|det| + i arg det,
// ld_J = tr(A^-1 dA), ld_L = tr(A^-1 A_L) - sum_k tr((A^-1 dA_k)^2), all complex), estimator/kinetic/_common.py:61-73
// (E_kin = -1/2 (lap + sum_k J_k^2) with the complex square).
// Complex tensors are stored as two float planes (re, im) with the layouts of their real counterparts.
#include "wf.cuh"

// ------------------------------------------------------------------------------------------------
// (orb_r + i orb_i)[w][j][c][d*n+i] *= envelope(j, i, d) * exp(i k_i . r_j)     in place, product rule.
// envelope E = sum_I pi exp(-|sigma| sd_I) on the periodic distance sd = r_ae (Local1, rows {x, 3 J, L} per electron).
// F = E P,  dF_a = (dE_a + i k_a E) P,  lap F = (lap E - |k|^2 E + 2 i k . dE) P.      One item per (w, j, d, i).
// ------------------------------------------------------------------------------------------------
__global__ void k_solid_orb_factor(float* __restrict__ orb_r, float* __restrict__ orb_i, const float* __restrict__ el,
                                   const float* __restrict__ r_ae, const float* __restrict__ klist, JqEnvelopeArgs env,
                                   long long items, JqSpins sp, int A, int D, int track) {
  const int n = sp.n();
  const int DN = D * n;
  const int C = track ? 3 * n + 2 : 1, C1 = track ? 5 : 1;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(it % DN);
    const long long g = it / DN;  // (w, j)
    const int j = (int)(g % n);
    const int d = col / n, i = col % n;
    const int ch = (env.pi[1] != nullptr) ? sp.chan_of(j) : 0;
    const float* pi = env.pi[ch];
    const float* sg = env.sigma[ch];
    const float* ra = r_ae + g * C1 * A;  // component c of atom I at ra[c*A + I]
    float E = (env.type == 2) ? 1.f : 0.f, dE[3] = {0.f, 0.f, 0.f}, lE = 0.f;
    if (env.type != 2)
      for (int I = 0; I < A; ++I) {
        float s = sg[(i * A + I) * D + d];
        if (env.type == 1) s = fabsf(s);
        const float t = pi[(i * A + I) * D + d] * expf(-s * ra[I]);
        E += t;
        if (track) {
          float g2 = 0.f;
          for (int a = 0; a < 3; ++a) {
            const float da = ra[(1 + a) * A + I];
            dE[a] = fmaf(-s * t, da, dE[a]);
            g2 = fmaf(da, da, g2);
          }
          lE += t * (s * s * g2 - s * ra[4 * A + I]);
        }
      }
    const float* e = el + g * 3;
    const float k0 = klist[3 * i], k1 = klist[3 * i + 1], k2 = klist[3 * i + 2];
    float ps, pc;
    sincosf_(k0 * e[0] + k1 * e[1] + k2 * e[2], &ps, &pc);
    // F = (fr, fi);  dF_a = (gr[a], gi[a]);  lap F = (lr, li)
    const float fr = E * pc, fi = E * ps;
    const float kk[3] = {k0, k1, k2};
    float gr[3], gi[3], lr = 0.f, li = 0.f;
    if (track) {
      float kdE = 0.f;
      for (int a = 0; a < 3; ++a) {
        // (dE + i k E)(pc + i ps)
        gr[a] = dE[a] * pc - kk[a] * E * ps;
        gi[a] = dE[a] * ps + kk[a] * E * pc;
        kdE = fmaf(kk[a], dE[a], kdE);
      }
      const float ar = lE - (k0 * k0 + k1 * k1 + k2 * k2) * E, ai = 2.f * kdE;
      lr = ar * pc - ai * ps;
      li = ar * ps + ai * pc;
    }
    float* pr = orb_r + g * (long long)C * DN + col;
    float* pq = orb_i + g * (long long)C * DN + col;
    const float o0r = pr[0], o0i = pq[0];
    if (track) {
      float cr = 0.f, ci = 0.f;  // sum_a O_{own a} dF_a
      for (int a = 0; a < 3; ++a) {
        const float ur = pr[(long long)(1 + 3 * j + a) * DN], ui = pq[(long long)(1 + 3 * j + a) * DN];
        cr += ur * gr[a] - ui * gi[a];
        ci += ur * gi[a] + ui * gr[a];
      }
      const float olr = pr[(long long)(C - 1) * DN], oli = pq[(long long)(C - 1) * DN];
      pr[(long long)(C - 1) * DN] = olr * fr - oli * fi + o0r * lr - o0i * li + 2.f * cr;
      pq[(long long)(C - 1) * DN] = olr * fi + oli * fr + o0r * li + o0i * lr + 2.f * ci;
      // Jacobian rows in batches of four: the eight loads of a batch are issued before any of its stores (r2 profile:
      // 74 % of the stall samples waited on one load pair at a time; the kernel streamed at 3.8 TB/s)
      int c = 1;
      for (; c + 4 <= C - 1; c += 4) {
        float ur[4], ui[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ur[u] = pr[(long long)(c + u) * DN];
          ui[u] = pq[(long long)(c + u) * DN];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float vr = ur[u] * fr - ui[u] * fi, vi = ur[u] * fi + ui[u] * fr;
          const int k = c + u - 1;
          if (k / 3 == j) {
            vr += o0r * gr[k % 3] - o0i * gi[k % 3];
            vi += o0r * gi[k % 3] + o0i * gr[k % 3];
          }
          pr[(long long)(c + u) * DN] = vr;
          pq[(long long)(c + u) * DN] = vi;
        }
      }
      for (; c < C - 1; ++c) {
        const float ur = pr[(long long)c * DN], ui = pq[(long long)c * DN];
        float vr = ur * fr - ui * fi, vi = ur * fi + ui * fr;
        const int k = c - 1;
        if (k / 3 == j) {
          vr += o0r * gr[k % 3] - o0i * gi[k % 3];
          vi += o0r * gi[k % 3] + o0i * gr[k % 3];
        }
        pr[(long long)c * DN] = vr;
        pq[(long long)c * DN] = vi;
      }
    }
    pr[0] = o0r * fr - o0i * fi;
    pq[0] = o0r * fi + o0i * fr;
  }
}

// ------------------------------------------------------------------------------------------------
// complex slogdet + forward-Laplacian rule.  One block per (walker, group of DB determinants); same scheme as the
// real kernel (logdet.cu) in complex arithmetic: Gauss-Jordan with partial pivoting on |z|,
//   log det = sum log|p| + i (sum arg p + pi * swaps),  M = A^-1 dA_c,  ld_J[c] = tr M,  ld_L = tr(A^-1 A_L) - sum_k tr(M_k^2).
// Shared (floats): logabs[DB] arg[DB] (double) | inv_r inv_i [DB][nn] | colp_r colp_i [DB][n] | piv[DB][n] |
//                  (J_r J_i invT_r invT_i [DB][n][NP4], then M_r M_i [DB][MT] tile-major, first) pv_r pv_i [DB] | trL_r trL_i t2_r t2_i [DB] |
//                  p1r p1i p2r p2i [DB][max(n, tiles)]
// ------------------------------------------------------------------------------------------------
__global__ void k_logdet_c(const float* __restrict__ orb_r, const float* __restrict__ orb_i, int n, int D, int C, int DB,
                           float* __restrict__ det_ld, float* __restrict__ det_grad, float* __restrict__ det_lap,
                           int plain_staging) {
  JQ_DYN_SMEM(float, sm);
  const int nn = n * n;
  const int K = C - 2;
  const int KT = (C > 1) ? C - 1 : 0;
  double* logabs = reinterpret_cast<double*>(sm);
  double* argsum = logabs + DB;
  // the four arrays read as float4 come first (16-byte aligned: 16 DB bytes of doubles before them), rows padded to NP4
  const int NP4 = (n + 3) & ~3;
  float* J_r = reinterpret_cast<float*>(argsum + DB);
  float* J_i = J_r + (size_t)DB * n * NP4;
  float* invT_r = J_i + (size_t)DB * n * NP4;
  float* invT_i = invT_r + (size_t)DB * n * NP4;
  // M = A^-1 dA_c is kept TILE-major (r2, late): the 4 x 4 tile (ti, tj) of a determinant is 16 contiguous floats at
  // ti * MS1 + tj * 20 -- a thread stores / re-reads its tile and reads the transposed tile (tj, ti) as float4s, and with
  // strides 20 and MS1 = 20 nb + 4 both walks are bank-conflict free for nb = 8.  The row-major M cost 4-way conflicts
  // on its 32 scalar stores and 4- / 8-way conflicts on the 64 scalar loads of the trace phase: half of the kernel's
  // shared-memory wavefronts (ncu, LiH 2x2x2: 1000 wavefronts per warp and slab, 436 of them in the product loop).
  const int nbt = (n + 3) >> 2;
  const int MS1 = 20 * nbt + 4, MT = nbt * MS1;
  float* M_r = invT_i + (size_t)DB * n * NP4;
  float* M_i = M_r + (size_t)DB * MT;
  float* inv_r = M_i + (size_t)DB * MT;
  float* inv_i = inv_r + (size_t)DB * nn;
  float* colp_r = inv_i + (size_t)DB * nn;
  float* colp_i = colp_r + DB * n;
  int* piv = reinterpret_cast<int*>(colp_i + DB * n);
  float* pv_r = reinterpret_cast<float*>(piv + DB * n);
  float* pv_i = pv_r + DB;
  float* trL_r = pv_i + DB;
  float* trL_i = trL_r + DB;
  float* t2_r = trL_i + DB;
  float* t2_i = t2_r + DB;
  const int np_ = ((n + 3) / 4) * ((n + 3) / 4) > n ? ((n + 3) / 4) * ((n + 3) / 4) : n;   // per-tile partials
  float* p1r = t2_i + DB;
  float* p1i = p1r + DB * np_;
  float* p2r = p1i + DB * np_;
  float* p2i = p2r + DB * np_;
  const int ngrp = (D + DB - 1) / DB;
  const long long w = blockIdx.x / ngrp;
  const int d0 = (int)(blockIdx.x % ngrp) * DB;
  const int db = (D - d0 < DB) ? D - d0 : DB;
  const int DN = D * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool need_inv = (C > 1);
  const long long base = (w * n) * (long long)C * DN + d0 * n;  // (j, c, d, i) at base + (j*C + c)*DN + d*n + i

  for (int q = tid; q < n * db * n; q += nt) {
    const int j = q / (db * n), r = q % (db * n);
    const int d = r / n, i = r % n;
    inv_r[d * nn + j * n + i] = orb_r[base + (long long)j * C * DN + r];
    inv_i[d * nn + j * n + i] = orb_i[base + (long long)j * C * DN + r];
  }
  for (int d = tid; d < db; d += nt) {
    logabs[d] = 0.0;
    argsum[d] = 0.0;
    t2_r[d] = t2_i[d] = trL_r[d] = trL_i[d] = 0.f;
  }
  __syncthreads();
  for (int p = 0; p < n; ++p) {
    for (int d = tid; d < db; d += nt) {
      const float* ar = inv_r + d * nn;
      const float* ai = inv_i + d * nn;
      int r = p;
      float best = ar[p * n + p] * ar[p * n + p] + ai[p * n + p] * ai[p * n + p];
      for (int q = p + 1; q < n; ++q) {
        float v = ar[q * n + p] * ar[q * n + p] + ai[q * n + p] * ai[q * n + p];
        if (v > best) { best = v; r = q; }
      }
      piv[d * n + p] = r;
      const float zr = ar[r * n + p], zi = ai[r * n + p];
      logabs[d] += 0.5 * log((double)best);
      argsum[d] += atan2((double)zi, (double)zr) + ((r != p) ? 3.14159265358979323846 : 0.0);
      const float inv2 = 1.0f / best;  // 1/z = conj(z)/|z|^2
      pv_r[d] = zr * inv2;
      pv_i[d] = -zi * inv2;
    }
    __syncthreads();
    for (int q = tid; q < db * n; q += nt) {
      const int d = q / n, c = q % n;
      const int r = piv[d * n + p];
      if (r != p) {
        float* ar = inv_r + d * nn;
        float* ai = inv_i + d * nn;
        float t = ar[p * n + c];
        ar[p * n + c] = ar[r * n + c];
        ar[r * n + c] = t;
        t = ai[p * n + c];
        ai[p * n + c] = ai[r * n + c];
        ai[r * n + c] = t;
      }
    }
    __syncthreads();
    if (need_inv) {
      for (int q = tid; q < db * n; q += nt) {
        const int d = q / n, i = q % n;
        colp_r[q] = inv_r[d * nn + i * n + p];
        colp_i[q] = inv_i[d * nn + i * n + p];
      }
      __syncthreads();
      for (int q = tid; q < db * n; q += nt) {
        const int d = q / n, c = q % n;
        float* ar = inv_r + d * nn + p * n + c;
        float* ai = inv_i + d * nn + p * n + c;
        const float xr = (c == p) ? 1.0f : *ar, xi = (c == p) ? 0.0f : *ai;
        *ar = xr * pv_r[d] - xi * pv_i[d];
        *ai = xr * pv_i[d] + xi * pv_r[d];
      }
      __syncthreads();
      for (int q = tid; q < db * nn; q += nt) {
        const int d = q / nn, rem = q % nn;
        const int i = rem / n, c = rem % n;
        if (i == p) continue;
        float* ar = inv_r + d * nn;
        float* ai = inv_i + d * nn;
        const float br = (c == p) ? 0.f : ar[rem], bi = (c == p) ? 0.f : ai[rem];
        const float cr = colp_r[d * n + i], ci = colp_i[d * n + i];
        const float rr = ar[p * n + c], ri = ai[p * n + c];
        ar[rem] = br - (cr * rr - ci * ri);
        ai[rem] = bi - (cr * ri + ci * rr);
      }
      __syncthreads();
    } else {
      for (int q = tid; q < db * n; q += nt) {
        const int d = q / n, i = q % n;
        const float xr = inv_r[d * nn + i * n + p], xi = inv_i[d * nn + i * n + p];
        colp_r[q] = xr * pv_r[d] - xi * pv_i[d];
        colp_i[q] = xr * pv_i[d] + xi * pv_r[d];
      }
      __syncthreads();
      for (int q = tid; q < db * nn; q += nt) {
        const int d = q / nn, rem = q % nn;
        const int i = rem / n, c = rem % n;
        if (i <= p || c <= p) continue;
        float* ar = inv_r + d * nn;
        float* ai = inv_i + d * nn;
        const float cr = colp_r[d * n + i], ci = colp_i[d * n + i];
        const float rr = ar[p * n + c], ri = ai[p * n + c];
        ar[rem] -= cr * rr - ci * ri;
        ai[rem] -= cr * ri + ci * rr;
      }
      __syncthreads();
    }
  }
  for (int d = tid; d < db; d += nt) {
    det_ld[(w * D + d0 + d) * 2] = (float)logabs[d];
    double a = fmod(argsum[d], 6.283185307179586476925286766559);
    if (a > 3.14159265358979323846) a -= 6.283185307179586476925286766559;
    if (a <= -3.14159265358979323846) a += 6.283185307179586476925286766559;
    det_ld[(w * D + d0 + d) * 2 + 1] = (float)a;
  }
  if (!need_inv) return;
  for (int p = n - 1; p >= 0; --p) {
    for (int q = tid; q < db * n; q += nt) {
      const int d = q / n, i = q % n;
      const int r = piv[d * n + p];
      if (r != p) {
        float* ar = inv_r + d * nn;
        float* ai = inv_i + d * nn;
        float t = ar[i * n + p];
        ar[i * n + p] = ar[i * n + r];
        ar[i * n + r] = t;
        t = ai[i * n + p];
        ai[i * n + p] = ai[i * n + r];
        ai[i * n + r] = t;
      }
    }
    __syncthreads();
  }
  // ---- traces (r2).  M = A^-1 dA_c in 4 x 4 complex register tiles: per contraction index j a tile reads 4 entries of a
  // j-major (transposed) copy of the inverse and 4 of the slab -- 16 shared-memory words for 16 complex multiply-adds
  // (64 FMAs), against 4 words per complex multiply-add for one output per thread (the kernel was bound by the
  // shared-memory pipe at 10 TFLOP/s).  tr M and tr M^2 = sum M[i][i2] M[i2][i] are per-tile partials summed in a
  // fixed order.
  const int nb = (n + 3) / 4, tiles = nb * nb;
  for (int q = tid; q < db * nn; q += nt) {
    const int d = q / nn, rem = q % nn;
    const int i = rem / n, j = rem % n;
    invT_r[(d * n + j) * NP4 + i] = inv_r[q];
    invT_i[(d * n + j) * NP4 + i] = inv_i[q];
  }
  __syncthreads();
  // ---- slab staging (r2, late).  mode 0: 16-byte cp.async chunks (4 consecutive orbitals of one (j, d) row) whose source /
  // destination offsets are computed ONCE per thread, double-buffered one slab ahead -- the second buffer is the inv_r /
  // inv_i region, dead after the transposed copy above (needs n % 4 == 0, i.e. NP4 == n, and 16-byte aligned rows).
  // Before: 4-byte copies with three integer divisions per element (~2.2 k of the ~4.8 k instructions a thread issued per
  // slab) and a wait for the slab's own global latency in every iteration.  mode 2 keeps that path, mode 1 plain loads.
#ifndef JAQMC_HOST_EMU
  constexpr int LC_MAXCH = 8;
  const int nq = n >> 2;
  const int nch = n * db * nq;   // 16-byte chunks per slab and array
  const bool vec = plain_staging == 0 && (n & 3) == 0 && (DN & 3) == 0 && nch <= LC_MAXCH * nt &&
                   ((reinterpret_cast<uintptr_t>(orb_r) | reinterpret_cast<uintptr_t>(orb_i)) & 15) == 0;
  int so[LC_MAXCH], dof[LC_MAXCH];
  unsigned jb0r = 0, jb0i = 0, jb1r = 0, jb1i = 0;
  if (vec) {
#pragma unroll
    for (int s = 0; s < LC_MAXCH; ++s) {
      const int q = tid + s * nt;
      const int j = q / (db * nq), r4 = q - j * (db * nq);
      const int d = r4 / nq, i4 = r4 - d * nq;
      so[s] = j * C * DN + d * n + 4 * i4;
      dof[s] = 4 * ((d * n + j) * NP4 + 4 * i4);
    }
    jb0r = (unsigned)__cvta_generic_to_shared(J_r);
    jb0i = (unsigned)__cvta_generic_to_shared(J_i);
    jb1r = (unsigned)__cvta_generic_to_shared(inv_r);
    jb1i = (unsigned)__cvta_generic_to_shared(inv_i);
  }
  auto stage_vec = [&](int slab) {
    if (slab < KT) {
      const float* sr_ = orb_r + base + (long long)(1 + slab) * DN;
      const float* si_ = orb_i + base + (long long)(1 + slab) * DN;
      const unsigned br = (slab & 1) ? jb1r : jb0r, bi = (slab & 1) ? jb1i : jb0i;
#pragma unroll
      for (int s = 0; s < LC_MAXCH; ++s)
        if (tid + s * nt < nch) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(br + (unsigned)dof[s]), "l"(sr_ + so[s]) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bi + (unsigned)dof[s]), "l"(si_ + so[s]) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (vec) stage_vec(0);
#else
  const bool vec = false;
#endif
  for (int kk = 0; kk < KT; ++kk) {
    const float* Jc_r = J_r;
    const float* Jc_i = J_i;
#ifndef JAQMC_HOST_EMU
    if (vec) {
      stage_vec(kk + 1);   // the other buffer: last read before the barriers that ended slab kk - 1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      if (kk & 1) {
        Jc_r = inv_r;
        Jc_i = inv_i;
      }
    } else if (plain_staging != 1) {
      const unsigned jr0 = (unsigned)__cvta_generic_to_shared(J_r), ji0 = (unsigned)__cvta_generic_to_shared(J_i);
      for (int q = tid; q < n * db * n; q += nt) {
        const int j = q / (db * n), r = q % (db * n);
        const int d = r / n, i = r % n;
        const long long src = base + ((long long)j * C + (1 + kk)) * DN + r;
        const unsigned off = 4u * (unsigned)((d * n + j) * NP4 + i);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(jr0 + off), "l"(orb_r + src));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ji0 + off), "l"(orb_i + src));
      }
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else
#endif
    for (int q = tid; q < n * db * n; q += nt) {
      const int j = q / (db * n), r = q % (db * n);
      const int d = r / n, i = r % n;
      const long long src = base + ((long long)j * C + (1 + kk)) * DN + r;
      J_r[(d * n + j) * NP4 + i] = orb_r[src];
      J_i[(d * n + j) * NP4 + i] = orb_i[src];
    }
    __syncthreads();
    for (int q = tid; q < db * tiles; q += nt) {
      const int d = q / tiles, t = q % tiles;
      const int i0 = 4 * (t / nb), c0 = 4 * (t % nb);
      const float* tr = invT_r + (size_t)d * n * NP4 + i0;
      const float* ti = invT_i + (size_t)d * n * NP4 + i0;
      const float* jr = Jc_r + (size_t)d * n * NP4 + c0;
      const float* ji = Jc_i + (size_t)d * n * NP4 + c0;
      float ar[4][4], ai[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) ar[a][b] = ai[a][b] = 0.f;
      // one float4 per operand and contraction index (rows padded to a multiple of 4; the padding columns only reach
      // accumulators that are not stored): 4 shared loads per 64 FMAs instead of 16 predicated scalar ones
#pragma unroll 2
      for (int j = 0; j < n; ++j) {
        const float4 xr4 = *reinterpret_cast<const float4*>(tr + j * NP4), xi4 = *reinterpret_cast<const float4*>(ti + j * NP4);
        const float4 yr4 = *reinterpret_cast<const float4*>(jr + j * NP4), yi4 = *reinterpret_cast<const float4*>(ji + j * NP4);
        const float xr[4] = {xr4.x, xr4.y, xr4.z, xr4.w}, xi[4] = {xi4.x, xi4.y, xi4.z, xi4.w};
        const float yr[4] = {yr4.x, yr4.y, yr4.z, yr4.w}, yi[4] = {yi4.x, yi4.y, yi4.z, yi4.w};
        // packed FMAs over column pairs (b, b + 1): same operations and order per accumulator as the scalar form
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float nxi = -xi[a];
#pragma unroll
          for (int b = 0; b < 4; b += 2) {
            jq_fma2(ar[a][b], ar[a][b + 1], xr[a], yr[b], yr[b + 1]);
            jq_fma2(ar[a][b], ar[a][b + 1], nxi, yi[b], yi[b + 1]);
            jq_fma2(ai[a][b], ai[a][b + 1], xr[a], yi[b], yi[b + 1]);
            jq_fma2(ai[a][b], ai[a][b + 1], xi[a], yr[b], yr[b + 1]);
          }
        }
      }
      float* mtr = M_r + (size_t)d * MT + (t / nb) * MS1 + (t % nb) * 20;
      float* mti = M_i + (size_t)d * MT + (t / nb) * MS1 + (t % nb) * 20;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float4 vr, vi;
        vr.x = ar[a][0]; vr.y = ar[a][1]; vr.z = ar[a][2]; vr.w = ar[a][3];
        vi.x = ai[a][0]; vi.y = ai[a][1]; vi.z = ai[a][2]; vi.w = ai[a][3];
        *reinterpret_cast<float4*>(mtr + 4 * a) = vr;
        *reinterpret_cast<float4*>(mti + 4 * a) = vi;
      }
    }
    __syncthreads();
    for (int q = tid; q < db * tiles; q += nt) {
      const int d = q / tiles, t = q % tiles;
      const int i0 = 4 * (t / nb), c0 = 4 * (t % nb);
      const int ti = t / nb, tj = t % nb;
      const float* own_r = M_r + (size_t)d * MT + ti * MS1 + tj * 20;
      const float* own_i = M_i + (size_t)d * MT + ti * MS1 + tj * 20;
      const float* tp_r = M_r + (size_t)d * MT + tj * MS1 + ti * 20;   // the transposed tile (tj, ti)
      const float* tp_i = M_i + (size_t)d * MT + tj * MS1 + ti * 20;
      float xr_[4][4], xi_[4][4], yr_[4][4], yi_[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float4 v0 = *reinterpret_cast<const float4*>(own_r + 4 * a), v1 = *reinterpret_cast<const float4*>(own_i + 4 * a);
        const float4 v2 = *reinterpret_cast<const float4*>(tp_r + 4 * a), v3 = *reinterpret_cast<const float4*>(tp_i + 4 * a);
        xr_[a][0] = v0.x; xr_[a][1] = v0.y; xr_[a][2] = v0.z; xr_[a][3] = v0.w;
        xi_[a][0] = v1.x; xi_[a][1] = v1.y; xi_[a][2] = v1.z; xi_[a][3] = v1.w;
        yr_[a][0] = v2.x; yr_[a][1] = v2.y; yr_[a][2] = v2.z; yr_[a][3] = v2.w;
        yi_[a][0] = v3.x; yi_[a][1] = v3.y; yi_[a][2] = v3.z; yi_[a][3] = v3.w;
      }
      float s1r = 0.f, s1i = 0.f, s2r = 0.f, s2i = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = i0 + a, i2 = c0 + b;
          if (i >= n || i2 >= n) continue;
          const float xr = xr_[a][b], xi = xi_[a][b], yr = yr_[b][a], yi = yi_[b][a];   // M[i][i2], M[i2][i]
          s2r += xr * yr - xi * yi;
          s2i += xr * yi + xi * yr;
          if (i == i2) {
            s1r += xr;
            s1i += xi;
          }
        }
      p1r[q] = s1r;
      p1i[q] = s1i;
      p2r[q] = s2r;
      p2i[q] = s2i;
    }
    __syncthreads();
#ifndef JAQMC_HOST_EMU
    // one warp per determinant sums the tile partials (lanes stride over the tiles, then a shuffle tree: fixed order);
    // only lane 0 of that warp ever touches t2 / trL of the determinant, and the partial arrays are rewritten two
    // barriers into the next slab, so no barrier is needed after this phase.  (Before: db threads summed tiles x 4
    // partials serially while the block waited.)
    for (int d = tid >> 5; d < db; d += nt >> 5) {
      const int lane = tid & 31;
      float s1r = 0.f, s1i = 0.f, s2r = 0.f, s2i = 0.f;
      for (int t = lane; t < tiles; t += 32) {
        s1r += p1r[d * tiles + t];
        s1i += p1i[d * tiles + t];
        s2r += p2r[d * tiles + t];
        s2i += p2i[d * tiles + t];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1r += __shfl_xor_sync(0xffffffffu, s1r, o);
        s1i += __shfl_xor_sync(0xffffffffu, s1i, o);
        s2r += __shfl_xor_sync(0xffffffffu, s2r, o);
        s2i += __shfl_xor_sync(0xffffffffu, s2i, o);
      }
      if (lane == 0) {
        if (kk < K) {
          det_grad[((w * D + d0 + d) * K + kk) * 2] = s1r;
          det_grad[((w * D + d0 + d) * K + kk) * 2 + 1] = s1i;
          t2_r[d] += s2r;
          t2_i[d] += s2i;
        } else {
          trL_r[d] = s1r;
          trL_i[d] = s1i;
        }
      }
    }
#else
    for (int d = tid; d < db; d += nt) {
      float s1r = 0.f, s1i = 0.f, s2r = 0.f, s2i = 0.f;
      for (int t = 0; t < tiles; ++t) {
        s1r += p1r[d * tiles + t];
        s1i += p1i[d * tiles + t];
        s2r += p2r[d * tiles + t];
        s2i += p2i[d * tiles + t];
      }
      if (kk < K) {
        det_grad[((w * D + d0 + d) * K + kk) * 2] = s1r;
        det_grad[((w * D + d0 + d) * K + kk) * 2 + 1] = s1i;
        t2_r[d] += s2r;
        t2_i[d] += s2i;
      } else {
        trL_r[d] = s1r;
        trL_i[d] = s1i;
      }
    }
    __syncthreads();
#endif
  }
#ifndef JAQMC_HOST_EMU
  if (vec) asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  __syncthreads();
  for (int d = tid; d < db; d += nt) {
    det_lap[(w * D + d0 + d) * 2] = trL_r[d] - t2_r[d];
    det_lap[(w * D + d0 + d) * 2 + 1] = trL_i[d] - t2_i[d];
  }
}

// complex log-sum-exp over determinants -> log psi (re, im), grad (3n complex), lap (complex), E_kin (complex)
__global__ void k_logdet_combine_c(const float* __restrict__ det_ld, const float* __restrict__ det_grad,
                                   const float* __restrict__ det_lap, int W, int n, int D, int track,
                                   float* __restrict__ logpsi_re, float* __restrict__ logpsi_im, float* __restrict__ grad,
                                   float* __restrict__ lap, float* __restrict__ e_kin) {
  const int K = 3 * n;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W;
       w += (long long)gridDim.x * blockDim.x) {
    const float* ld = det_ld + w * D * 2;
    float lmax = ld[0];
    for (int d = 1; d < D; ++d) lmax = fmaxf(lmax, ld[2 * d]);
    float sr = 0.f, si = 0.f;
    for (int d = 0; d < D; ++d) {
      const float m = expf(ld[2 * d] - lmax);
      float sn, cs;
      sincosf_(ld[2 * d + 1], &sn, &cs);
      sr = fmaf(m, cs, sr);
      si = fmaf(m, sn, si);
    }
    const float s2 = sr * sr + si * si;
    logpsi_re[w] = 0.5f * logf(s2) + lmax;
    logpsi_im[w] = atan2f(si, sr);
    if (!track) continue;
    const float* g = det_grad + w * D * K * 2;
    const float inv2 = 1.0f / s2;
    float alr = 0.f, ali = 0.f;  // sum_d w_d (lap_d + sum_k g_dk^2)
    for (int d = 0; d < D; ++d) {
      const float m = expf(ld[2 * d] - lmax);
      float sn, cs;
      sincosf_(ld[2 * d + 1], &sn, &cs);
      const float er = m * cs, ei = m * sn;  // exp(ld_d - lmax)
      const float wr = (er * sr + ei * si) * inv2, wi = (ei * sr - er * si) * inv2;  // / s
      float qr = det_lap[(w * D + d) * 2], qi = det_lap[(w * D + d) * 2 + 1];
      for (int k = 0; k < K; ++k) {
        const float xr = g[(d * K + k) * 2], xi = g[(d * K + k) * 2 + 1];
        qr += xr * xr - xi * xi;
        qi += 2.f * xr * xi;
      }
      alr += wr * qr - wi * qi;
      ali += wr * qi + wi * qr;
    }
    float ggr = 0.f, ggi = 0.f;
    for (int k = 0; k < K; ++k) {
      float gr = 0.f, gi = 0.f;
      for (int d = 0; d < D; ++d) {
        const float m = expf(ld[2 * d] - lmax);
        float sn, cs;
        sincosf_(ld[2 * d + 1], &sn, &cs);
        const float er = m * cs, ei = m * sn;
        const float wr = (er * sr + ei * si) * inv2, wi = (ei * sr - er * si) * inv2;
        const float xr = g[(d * K + k) * 2], xi = g[(d * K + k) * 2 + 1];
        gr += wr * xr - wi * xi;
        gi += wr * xi + wi * xr;
      }
      grad[(w * K + k) * 2] = gr;
      grad[(w * K + k) * 2 + 1] = gi;
      ggr += gr * gr - gi * gi;
      ggi += 2.f * gr * gi;
    }
    const float lr = alr - ggr, li = ali - ggi;
    lap[2 * w] = lr;
    lap[2 * w + 1] = li;
    e_kin[2 * w] = -0.5f * (lr + ggr);
    e_kin[2 * w + 1] = -0.5f * (li + ggi);
  }
}

#ifndef JAQMC_HOST_EMU
// The same reduction with one WARP per walker (D <= 32; r2): lane d holds determinant d's weight
// w_d = exp(ld_d - lmax) / sum, the lanes stride over the 3n derivative components and make ONE pass over det_grad
// (coalesced float2 reads), accumulating grad_k = sum_d w_d g_dk and sum_d w_d g_dk^2 together.  The thread-per-walker
// kernel above took 0.80 ms for a 512-walker shard of LiH 2x2x2 (8 blocks of 64 threads, exp / sincos recomputed in the
// inner loop); the sums over k are taken in a different order, so results agree to rounding, not bit for bit.
__global__ void __launch_bounds__(256) k_logdet_combine_c_warp(const float* __restrict__ det_ld,
                                                               const float* __restrict__ det_grad,
                                                               const float* __restrict__ det_lap, int W, int n, int D,
                                                               int track, float* __restrict__ logpsi_re,
                                                               float* __restrict__ logpsi_im, float* __restrict__ grad,
                                                               float* __restrict__ lap, float* __restrict__ e_kin) {
  const unsigned full = 0xffffffffu;
  const int K = 3 * n;
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= W) return;
  const bool has = lane < D;
  const float lr_d = has ? det_ld[(w * D + lane) * 2] : -INFINITY;
  const float li_d = has ? det_ld[(w * D + lane) * 2 + 1] : 0.f;
  float lmax = lr_d;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(full, lmax, o));
  float er = 0.f, ei = 0.f;
  if (has) {
    const float m = expf(lr_d - lmax);
    float sn, cs;
    sincosf_(li_d, &sn, &cs);
    er = m * cs;
    ei = m * sn;
  }
  float sr = er, si = ei;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(full, sr, o);
    si += __shfl_xor_sync(full, si, o);
  }
  const float s2 = sr * sr + si * si;
  if (lane == 0) {
    logpsi_re[w] = 0.5f * logf(s2) + lmax;
    logpsi_im[w] = atan2f(si, sr);
  }
  if (!track) return;
  const float inv2 = 1.0f / s2;
  const float wr_d = (er * sr + ei * si) * inv2, wi_d = (ei * sr - er * si) * inv2;   // exp(ld_d - lmax) / s
  // sum_d w_d lap_d (lane d's term), then + sum_d w_d sum_k g_dk^2 below
  float alr = 0.f, ali = 0.f;
  if (has) {
    const float qr = det_lap[(w * D + lane) * 2], qi = det_lap[(w * D + lane) * 2 + 1];
    alr = wr_d * qr - wi_d * qi;
    ali = wr_d * qi + wi_d * qr;
  }
  const float2* g = reinterpret_cast<const float2*>(det_grad) + w * D * K;
  float2* gout = reinterpret_cast<float2*>(grad) + w * K;
  float ggr = 0.f, ggi = 0.f;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    const bool kv = k < K;
    float gr = 0.f, gi = 0.f;
    for (int d = 0; d < D; ++d) {
      const float wr = __shfl_sync(full, wr_d, d), wi = __shfl_sync(full, wi_d, d);
      const float2 x = kv ? g[d * K + k] : make_float2(0.f, 0.f);
      gr += wr * x.x - wi * x.y;
      gi += wr * x.y + wi * x.x;
      const float x2r = x.x * x.x - x.y * x.y, x2i = 2.f * x.x * x.y;
      alr += wr * x2r - wi * x2i;
      ali += wr * x2i + wi * x2r;
    }
    if (kv) gout[k] = make_float2(gr, gi);
    ggr += gr * gr - gi * gi;
    ggi += 2.f * gr * gi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    alr += __shfl_xor_sync(full, alr, o);
    ali += __shfl_xor_sync(full, ali, o);
    ggr += __shfl_xor_sync(full, ggr, o);
    ggi += __shfl_xor_sync(full, ggi, o);
  }
  if (lane == 0) {
    const float lr = alr - ggr, li = ali - ggi;
    lap[2 * w] = lr;
    lap[2 * w + 1] = li;
    e_kin[2 * w] = -0.5f * (lr + ggr);
    e_kin[2 * w + 1] = -0.5f * (li + ggi);
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// pipeline
// ------------------------------------------------------------------------------------------------
namespace {
struct SolidBufs {
  FermiBufs f;
  float *r_ae, *orb_r, *orb_i, *det_ld, *det_grad, *det_lap;
};

void solid_carve(const FermiDims& d, long long W, JqArena& ar, SolidBufs* b) {
  const long long n = d.n;
  jq_fermi_carve_backbone(d, W, ar, &b->f);
  b->r_ae = ar.take<float>(W * n * d.C1 * d.A);
  b->orb_r = ar.take<float>(W * n * d.C * d.D * n);
  b->orb_i = ar.take<float>(W * n * d.C * d.D * n);
  b->det_ld = ar.take<float>(W * d.D * 2);
  b->det_grad = ar.take<float>(W * d.D * (d.C > 1 ? 3 * n : 1) * 2);
  b->det_lap = ar.take<float>(W * d.D * 2);
}

int solid_dims(const jaqmc_solid_config* c, int track, FermiDims* d) {
  const int fw = c->distance_type == JAQMC_DISTANCE_NU ? 4 : 7;   // features per electron-atom / electron-electron pair
  return jq_fermi_dims(&c->net, track, fw, fw, d);
}

size_t logdet_c_smem(int db, int n) {
  const size_t nn = (size_t)n * n;
  const size_t nb = (size_t)(n + 3) / 4;
  const size_t np = nb * nb > (size_t)n ? nb * nb : (size_t)n;
  const size_t np4 = (size_t)((n + 3) & ~3);
  const size_t mt = nb * (20 * nb + 4);   // tile-major M (see k_logdet_c)
  return 16 * (size_t)db + sizeof(float) * (2 * db * nn + 2 * db * mt + 4 * (size_t)db * n * np4 + 3 * (size_t)db * n + 4 * (size_t)db * np + 6 * (size_t)db) + 32;
}
}  // namespace

size_t jq_solid_ws_bytes(const jaqmc_solid_config* c, long long W, int track) {
  FermiDims d;
  if (solid_dims(c, track, &d) != JQ_OK) return 0;
  JqArena ar(nullptr, 0);
  SolidBufs b;
  solid_carve(d, W, ar, &b);
  return ar.off;
}

int jq_solid_forward(const jaqmc_solid_config* c, const jaqmc_solid_params* p, const jaqmc_system* sys,
                     const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOutC out,
                     cudaStream_t st) {
  FermiDims d;
  int rc = solid_dims(c, track, &d);
  if (rc) return rc;
  JQ_REQUIRE(sys && sys->atoms && sys->n_atoms == d.A, JQ_ERR_INVALID_ARGUMENT,
             "solid: system must hold the %d primitive-cell atoms", d.A);
  JQ_REQUIRE(p->klist, JQ_ERR_INVALID_ARGUMENT, "solid: null klist");
  JQ_REQUIRE(!c->net.use_last_layer, JQ_ERR_UNSUPPORTED, "solid: use_last_layer is not implemented");
  JQ_REQUIRE(c->net.envelope_type != JAQMC_ENVELOPE_DIAGONAL, JQ_ERR_UNSUPPORTED,
             "solid: the diagonal envelope (6-component tri displacement) is not implemented");
  const bool split = c->net.orbitals_spin_split && d.nch == 2;
  JQ_REQUIRE(p->real_orbital_kernel[0] && p->imag_orbital_kernel[0] &&
                 (!split || (p->real_orbital_kernel[1] && p->imag_orbital_kernel[1])),
             JQ_ERR_INVALID_ARGUMENT, "solid: null orbital kernel");
  JQ_REQUIRE(c->net.envelope_type == JAQMC_ENVELOPE_NULL ||
                 (p->net.env_pi[0] && p->net.env_sigma[0] && (!split || (p->net.env_pi[1] && p->net.env_sigma[1]))),
             JQ_ERR_INVALID_ARGUMENT, "solid: null envelope parameter");
  JqArena ar(ws, ws_bytes);
  SolidBufs b;
  solid_carve(d, W, ar, &b);
  JQ_REQUIRE(ar.ok(), JQ_ERR_WORKSPACE_TOO_SMALL, "solid: workspace %zu < %zu bytes", ws_bytes, ar.off);
  const int n = d.n;
  if ((rc = jq_launch_solid_features(electrons, sys->atoms, c->simulation_lattice, c->primitive_lattice, c->distance_type,
                                     c->sym_type, (int)W, n, d.A,
                                     track, b.f.ae, b.r_ae, b.f.h2a, st)))
    return rc;
  float* h = nullptr;
  if ((rc = jq_fermi_backbone(d, &p->net, W, track, b.f, st, &h))) return rc;
  // real and imaginary orbital projections (per spin channel DenseGeneral, no bias)
  const int nchan = split ? 2 : 1;
  for (int part = 0; part < 2; ++part)
    for (int s = 0; s < nchan; ++s) {
      JqDenseArgs a;
      memset(&a, 0, sizeof(a));
      a.src0 = h;
      a.k0 = d.d1[d.L - 1];
      a.w0 = part ? p->imag_orbital_kernel[s] : p->real_orbital_kernel[s];
      a.out = part ? b.orb_i : b.orb_r;
      a.wscratch = b.f.wscr;
      a.N = d.D * n;
      a.C = d.C;
      a.n_tot = n;
      a.j0 = split ? d.sp.lo(s) : 0;
      a.n_sub = split ? d.sp.hi(s) - d.sp.lo(s) : n;
      a.G = W * a.n_sub;
      if ((rc = jq_launch_dense(a, st))) return rc;
    }
  JqEnvelopeArgs env;
  env.type = c->net.envelope_type;
  env.pi[0] = p->net.env_pi[0];
  env.sigma[0] = p->net.env_sigma[0];
  env.pi[1] = split ? p->net.env_pi[1] : nullptr;
  env.sigma[1] = split ? p->net.env_sigma[1] : nullptr;
  {
    long long items = W * n * d.D * n;
    int grid = jq_cdiv(items, 256);
    if (grid > 148 * 32) grid = 148 * 32;
    jq_prof_work(0.0, 16.0 * (double)items * d.C);
    JQ_LAUNCH(k_solid_orb_factor, dim3(grid), dim3(256), 0, st, b.orb_r, b.orb_i, electrons, b.r_ae, p->klist, env, items,
              d.sp, d.A, d.D, track);
    JQ_CHECK_LAUNCH();
  }
  if (out.orbitals) {   // wf.orbitals: complex matrices, no determinant
    JQ_REQUIRE(!track, JQ_ERR_INVALID_ARGUMENT, "solid: orbitals are emitted on the value path only");
    return jq_launch_orbitals_out(b.orb_r, b.orb_i, out.orbitals, W, n, d.D, st);
  }
  {
    int DB = d.D;
    // several resident blocks per SM hide each other's slab loads and barriers: at most ~72 KB per block
    // (n = 32: two determinants, 128 register tiles, 128 threads)
    while (DB > 1 && logdet_c_smem(DB, n) > 74 * 1024) DB = (DB + 1) / 2;   // three blocks per SM: 3 x (74 + 1) KB <= 228 KB
    size_t smem = logdet_c_smem(DB, n);
    JQ_REQUIRE(smem <= 200 * 1024, JQ_ERR_UNSUPPORTED, "solid: %d electrons need %zu bytes of shared memory", n, smem);
#ifndef JAQMC_HOST_EMU
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_logdet_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "solid: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
#endif
    const long long blocks = W * ((d.D + DB - 1) / DB);
    jq_prof_work((double)W * d.D * 8.0 * n * n * n * (track ? 2 * (d.C - 1) + 1 : 0.34), 8.0 * (double)W * d.D * d.C * n * n);
    const int tiles_blk = DB * ((n + 3) / 4) * ((n + 3) / 4);
    const int nthr = (track && tiles_blk <= 128) ? 128 : 256;
    // A/B switches: 1 = plain load / store staging, 2 = 4-byte cp.async staging (r2's first version), 0 = default
    static const int plain_staging = getenv("JAQMC_B200_LOGDET_PLAIN_STAGING") ? 1 : getenv("JAQMC_B200_LOGDET_C_SCALAR_STAGING") ? 2 : 0;
    JQ_LAUNCH(k_logdet_c, dim3((unsigned)blocks), dim3(nthr), smem, st, b.orb_r, b.orb_i, n, d.D, d.C, DB, b.det_ld,
              b.det_grad, b.det_lap, plain_staging);
    JQ_CHECK_LAUNCH();
  }
#ifndef JAQMC_HOST_EMU
  static const bool combine_thread = getenv("JAQMC_B200_COMBINE_C_THREAD") != nullptr;   // A/B switch
  if (d.D <= 32 && !combine_thread) {
    JQ_LAUNCH(k_logdet_combine_c_warp, dim3((unsigned)jq_cdiv(W, 8)), dim3(256), 0, st, b.det_ld, b.det_grad, b.det_lap,
              (int)W, n, d.D, track, out.logpsi_re, out.logpsi_im, out.grad, out.lap, out.e_kin);
  } else
#endif
  JQ_LAUNCH(k_logdet_combine_c, dim3(jq_cdiv(W, 64)), dim3(64), 0, st, b.det_ld, b.det_grad, b.det_lap, (int)W, n, d.D,
            track, out.logpsi_re, out.logpsi_im, out.grad, out.lap, out.e_kin);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}
