// Parameter-gradient path of the FermiNet (molecule): the VJP of  theta -> vmap(log|psi|)(theta, walkers)  with a
// per-walker cotangent.
//
// Reference semantics: estimator/loss_grad.py:70-128 (LossAndGrad).  The reference materialises the per-walker score
// d log|psi| / d theta (a W x P tensor, `jax.vmap(jax.value_and_grad)`), multiplies it by the clipped local energies and
// averages; its `grads` is
//     2 ( <score * E_clip> - <E_clip> <score> )  =  sum_w c_w * score_w,    c_w = 2 (E_clip,w - <E_clip>) / W.
// The right-hand side is ONE reverse pass with the cotangent c: no W x P tensor exists here.  (`<score>` alone, which
// the reference also reports, is the same call with c_w = 1 / W.)
//
// Forward: the value-only pipeline of ferminet.cu (same restructured layers: per-walker spin-mean part contracted once)
// with every activation kept.  Reverse: dense layers by  dX = dZ W^T  (the forward dense kernels on the transposed
// weights, tcgen05 where the shape allows) and  dW = X^T dZ  (row-reduction GEMM, split over row chunks and reduced in a
// fixed order: bit-reproducible); tanh / residual / spin-mean / pair-mean adjoints elementwise; the determinant head by
// d log|det M| / dM = M^-T weighted with the log-sum-exp weights.  n <= 32 electrons (half-warp / warp inversion in registers).
#include "wf.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// small elementwise / reduction kernels (block-stride: also run by the host emulation)
// ---------------------------------------------------------------------------------------------------------------
// env[w][e][d*n + o] = sum_I pi[o][I][d] exp(-s r),  s = |sigma| (abs_isotropic) or sigma (isotropic)
__global__ void k_envelope_values(const float* __restrict__ el, const float* __restrict__ atoms, JqEnvelopeArgs env,
                                  long long items, JqSpins sp, int A, int D, float* __restrict__ out) {
  const int n = sp.n(), DN = D * n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(it % DN);
    const long long g = it / DN;   // (w, e)
    const int e = (int)(g % n);
    const int d = col / n, o = col % n;
    const int ch = env.pi[1] ? sp.chan_of(e) : 0;
    const float* x = el + g * 3;
    float acc = 0.f;
    for (int I = 0; I < A; ++I) {
      const float dx = x[0] - atoms[3 * I], dy = x[1] - atoms[3 * I + 1], dz = x[2] - atoms[3 * I + 2];
      const float r = sqrtf(dx * dx + dy * dy + dz * dz);
      float s = env.sigma[ch][(o * A + I) * D + d];
      if (env.type == 1) s = fabsf(s);
      acc += env.pi[ch][(o * A + I) * D + d] * expf(-s * r);
    }
    out[it] = acc;
  }
}

__global__ void k_mul(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    o[i] = a[i] * b[i];
}

// One item per matrix (walker, determinant): Gauss-Jordan with partial pivoting on a local copy.
//   M[w][e][d*n + o]  ->  minv[w][d][o][e] = (M_d^-1)[o][e],  sign, log|det|
#define BW_NMAX 32   // r2, late: 16 before (a full warp per matrix for 17 ... 32 electrons)
__global__ void k_minv(const float* __restrict__ M, long long MT, int n, int D, float* __restrict__ minv,
                       float* __restrict__ sign, float* __restrict__ logabs) {
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < MT; m += (long long)gridDim.x * blockDim.x) {
    const long long w = m / D;
    const int d = (int)(m - w * D);
    float a[BW_NMAX][BW_NMAX], b[BW_NMAX][BW_NMAX];
    for (int e = 0; e < n; ++e)
      for (int o = 0; o < n; ++o) {
        a[e][o] = M[(w * n + e) * (long long)D * n + d * n + o];
        b[e][o] = (e == o) ? 1.f : 0.f;
      }
    float sg = 1.f;
    double la = 0.0;
    for (int p = 0; p < n; ++p) {
      int r = p;
      float best = fabsf(a[p][p]);
      for (int q = p + 1; q < n; ++q)
        if (fabsf(a[q][p]) > best) {
          best = fabsf(a[q][p]);
          r = q;
        }
      if (r != p) {
        for (int c = 0; c < n; ++c) {
          float t = a[p][c];
          a[p][c] = a[r][c];
          a[r][c] = t;
          t = b[p][c];
          b[p][c] = b[r][c];
          b[r][c] = t;
        }
        sg = -sg;
      }
      const float pv = a[p][p];
      if (pv < 0.f) sg = -sg;
      if (pv == 0.f) sg = 0.f;
      la += log((double)fabsf(pv));
      const float pinv = 1.0f / pv;
      for (int c = 0; c < n; ++c) {
        a[p][c] *= pinv;
        b[p][c] *= pinv;
      }
      for (int q = 0; q < n; ++q) {
        if (q == p) continue;
        const float f = a[q][p];
        for (int c = 0; c < n; ++c) {
          a[q][c] = fmaf(-f, a[p][c], a[q][c]);
          b[q][c] = fmaf(-f, b[p][c], b[q][c]);
        }
      }
    }
    // b = M_d^-1 with M_d[e][o]: b[o][e]
    for (int o = 0; o < n; ++o)
      for (int e = 0; e < n; ++e) minv[(m * n + o) * n + e] = b[o][e];
    sign[m] = sg;
    logabs[m] = (float)la;
  }
}

#ifndef JAQMC_HOST_EMU
// Device version (r2): one HALF-warp per matrix, lane r = row r of [A | I] in registers, Gauss-Jordan with shuffle
// pivoting and no row movement (the scheme of k_logdet_small): at step p the pivot is the largest |a[r][p]| among the
// rows not used yet -- the same choice as the row-swapping elimination above -- and row p of A^-1 is the right half of
// the row that served as pivot p.  The thread-per-matrix kernel kept 2 KB of local arrays per thread and ran with 256
// blocks of 64 threads: 1.50 ms of a 10.5 ms reverse pass (ncu, N2, 4096 walkers).
// G lanes per matrix: 16 (two matrices per warp, n <= 16) or 32 (one per warp, n <= 32).
template <int G>
__global__ void __launch_bounds__(256) k_minv_small(const float* __restrict__ M, long long MT, int n, int D,
                                                    float* __restrict__ minv, float* __restrict__ sign,
                                                    float* __restrict__ logabs) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, half = lane / G, hl = lane & (G - 1);
  const long long m0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / G) + half;
  const bool on = m0 < MT;
  const long long m = on ? m0 : MT - 1;
  const long long w = m / D;
  const int d = (int)(m - w * D);
  bool used = !on || hl >= n;
  float a[G], bi[G];
  {
    const float* row = M + (w * n + (hl < n ? hl : 0)) * (long long)D * n + d * n;
#pragma unroll
    for (int c = 0; c < G; ++c) {
      a[c] = (!used && c < n) ? row[c] : 0.f;
      bi[c] = (c == hl) ? 1.0f : 0.f;
    }
  }
  int step_of = 0;
  float sg = 1.0f;
  double mant = 1.0;
  int expo = 0;
#pragma unroll
  for (int p = 0; p < G; ++p) {
    if (p < n) {
      const unsigned key = used ? 0u : __float_as_uint(fabsf(a[p]));
      unsigned mx = key;
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(full, mx, o));
      const unsigned cand = (__ballot_sync(full, !used && key == mx) >> (G * half)) & (G == 32 ? 0xffffffffu : 0xffffu);
      const int pl = cand ? __ffs(cand) - 1 : 0;
      const float pv = __shfl_sync(full, a[p], pl, G);
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
      const float f = is_p ? 0.f : a[p];
#pragma unroll
      for (int c = 0; c < G; ++c)
        if (c < n) {
          const float ap = __shfl_sync(full, a[c], pl, G) * pinv;
          const float bp = __shfl_sync(full, bi[c], pl, G) * pinv;
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
    const int si = __shfl_sync(full, step_of, i, G);
    if (i < hl && hl < n && si > step_of) ++invc;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) invc += __shfl_xor_sync(full, invc, o);
  if (on && hl < n) {
    float* dst = minv + (m * n + step_of) * n;
#pragma unroll
    for (int c = 0; c < G; ++c)
      if (c < n) dst[c] = bi[c];
  }
  if (on && hl == 0) {
    sign[m] = (invc & 1) ? -sg : sg;
    logabs[m] = (float)(log(mant) + (double)expo * 0.69314718055994530942);
  }
}

// out[p] = sum_s part[s * stride + p] for few outputs and many partials: a block takes 32 outputs (lanes) x 8 slices of
// the partial index (warps); slice sums and the final sum over the slices both run in a fixed order.  The simple kernel
// below walks all S partials in one thread: with S ~ 1000 and one block (the column sums of the bias gradients) that was
// 0.32 - 0.78 ms per call, 3.9 ms of the reverse pass.
__global__ void __launch_bounds__(256) k_reduce_partials_wide(const float* __restrict__ part, int S, long long stride,
                                                              long long P, float* __restrict__ out) {
  __shared__ float sl[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long p0 = (long long)blockIdx.x * 32; p0 < P; p0 += (long long)gridDim.x * 32) {
    const long long p = p0 + lane;
    float acc = 0.f;
    if (p < P)
      for (int s2 = warp; s2 < S; s2 += 8) acc += part[(long long)s2 * stride + p];
    sl[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && p < P) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += sl[q][lane];
      out[p] = t;
    }
    __syncthreads();
  }
}
#endif

// log-sum-exp over determinants (output/logdet.py:65-79) and its adjoint weights:
//   wdet[w][d] = cot[w] * s_d e^{ld_d - max} / sum_e s_e e^{ld_e - max}
__global__ void k_det_weights(const float* __restrict__ sign, const float* __restrict__ logabs,
                              const float* __restrict__ cot, long long W, int D, float* __restrict__ wdet,
                              float* __restrict__ logpsi, float* __restrict__ psign) {
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (long long)gridDim.x * blockDim.x) {
    float mx = -INFINITY;
    for (int d = 0; d < D; ++d) mx = fmaxf(mx, logabs[w * D + d]);
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += sign[w * D + d] * expf(logabs[w * D + d] - mx);
    const float inv = cot[w] / s;
    for (int d = 0; d < D; ++d) wdet[w * D + d] = sign[w * D + d] * expf(logabs[w * D + d] - mx) * inv;
    if (logpsi) logpsi[w] = logf(fabsf(s)) + mx;
    if (psign) psign[w] = (s > 0.f) ? 1.f : (s < 0.f ? -1.f : 0.f);
  }
}

// dM[w][e][d*n+o] = wdet[w][d] * minv[w][d][o][e];  dorb = dM * env,  denv = dM * orb
__global__ void k_head_bwd(const float* __restrict__ wdet, const float* __restrict__ minv, const float* __restrict__ orb,
                           const float* __restrict__ env, long long items, int n, int D, float* __restrict__ dorb,
                           float* __restrict__ denv) {
  const int DN = D * n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(it % DN);
    const long long g = it / DN;
    const int e = (int)(g % n);
    const long long w = g / n;
    const int d = col / n, o = col % n;
    const float dM = wdet[w * D + d] * minv[((w * D + d) * n + o) * n + e];
    dorb[it] = env ? dM * env[it] : dM;
    if (denv) denv[it] = dM * orb[it];
  }
}

// Envelope-parameter adjoints, partial over a chunk of walkers: one item per (chunk, channel, o, I, d).
//   dpi[o][I][d] += denv * exp(-s r);   dsigma[o][I][d] += denv * pi * exp(-s r) * (-r) * (abs: sign(sigma))
__global__ void k_env_param_grad(const float* __restrict__ denv, const float* __restrict__ el,
                                 const float* __restrict__ atoms, JqEnvelopeArgs env, long long W, JqSpins sp, int A,
                                 int D, int chunks, float* __restrict__ part_pi, float* __restrict__ part_sigma) {
  const int n = sp.n(), DN = D * n;
  const int nchp = env.pi[1] ? 2 : 1;
  const long long P = (long long)nchp * n * A * D;
  const long long items = (long long)chunks * P;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const long long pidx = it % P;
    const int chunk = (int)(it / P);
    const int d = (int)(pidx % D);
    long long t = pidx / D;
    const int I = (int)(t % A);
    t /= A;
    const int o = (int)(t % n);
    const int ch = (int)(t / n);
    const long long w0 = W * chunk / chunks, w1 = W * (chunk + 1) / chunks;
    const int e0 = (nchp == 2) ? sp.lo(ch) : 0, e1 = (nchp == 2) ? sp.hi(ch) : n;
    const float sg_raw = env.sigma[ch][(o * A + I) * D + d];
    const float s = (env.type == 1) ? fabsf(sg_raw) : sg_raw;
    const float dsds = (env.type == 1) ? (sg_raw > 0.f ? 1.f : (sg_raw < 0.f ? -1.f : 0.f)) : 1.f;
    const float pv = env.pi[ch][(o * A + I) * D + d];
    const float ax = atoms[3 * I], ay = atoms[3 * I + 1], az = atoms[3 * I + 2];
    float gp = 0.f, gs = 0.f;
    for (long long w = w0; w < w1; ++w)
      for (int e = e0; e < e1; ++e) {
        const float* x = el + (w * n + e) * 3;
        const float dx = x[0] - ax, dy = x[1] - ay, dz = x[2] - az;
        const float r = sqrtf(dx * dx + dy * dy + dz * dz);
        const float ex = expf(-s * r);
        const float de = denv[(w * n + e) * (long long)DN + d * n + o];
        gp = fmaf(de, ex, gp);
        gs = fmaf(de * pv * ex, -r, gs);
      }
    part_pi[it] = gp;
    part_sigma[it] = gs * dsds;
  }
}

// out[p] = sum_s part[s * stride + p], p < P  (fixed order)
__global__ void k_reduce_partials(const float* __restrict__ part, int S, long long stride, long long P,
                                  float* __restrict__ out) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += part[(long long)s * stride + p];
    out[p] = acc;
  }
}

// out[(w * ns + e)][f] = in[(w * n + lo + e)][f]: the rows of one spin channel, compacted
__global__ void k_gather_channel(const float* __restrict__ in, long long W, int n, int lo, int ns, int F,
                                 float* __restrict__ out) {
  const long long items = W * ns * F;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(it % F);
    const long long r = it / F;
    const int e = (int)(r % ns);
    const long long w = r / ns;
    out[it] = in[(w * n + lo + e) * (long long)F + f];
  }
}

// Adjoint of  h_next = res ? (h_prev + y)/sqrt2 : y,  y = tanh(z):
//   dz = (res ? dh/sqrt2 : dh) * (1 - y^2),  y recovered from the stored activations;  dres = dh/sqrt2 (res only)
__global__ void k_tanh_bwd(const float* __restrict__ dh, const float* __restrict__ h_next, const float* __restrict__ h_prev,
                           int res, long long count, float* __restrict__ dz, float* __restrict__ dres) {
  const float is2 = 0.70710678118654752440f, s2 = 1.41421356237309504880f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float g = dh[i];
    float y, dy;
    if (res) {
      y = fmaf(s2, h_next[i], -h_prev[i]);
      dy = g * is2;
      if (dres) dres[i] = dy;
    } else {
      y = h_next[i];
      dy = g;
    }
    dz[i] = dy * (1.0f - y * y);
  }
}

// dzw[w][f] = sum_e dz[w][e][f]
__global__ void k_walker_sum(const float* __restrict__ dz, long long W, int n, int F, float* __restrict__ out) {
  const long long items = W * F;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(it % F);
    const long long w = it / F;
    float acc = 0.f;
    for (int e = 0; e < n; ++e) acc += dz[(w * n + e) * (long long)F + f];
    out[it] = acc;
  }
}

// adjoint of the spin means: dh[w][e][f] += dm[w][s(e)*F + f] / n_s
__global__ void k_spin_mean_bwd(const float* __restrict__ dm, long long W, JqSpins sp, int F, float* __restrict__ dh) {
  const int n = sp.n(), nch = sp.nch();
  const long long items = W * n * F;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(it % F);
    const long long g = it / F;
    const int e = (int)(g % n);
    const long long w = g / n;
    const int s = sp.chan_of(e);
    dh[it] += dm[(w * nch + s) * (long long)F + f] / (float)(sp.hi(s) - sp.lo(s));
  }
}

// adjoint of the pair means g2[w][j][s*F + f] = mean_{i in s} h2[w][i][j][f]:
//   dh2[w][i][j][f] (+)= dg2[w][j][s(i)*F + f] / n_{s(i)}
__global__ void k_pair_mean_bwd(const float* __restrict__ dg2, long long W, JqSpins sp, int F, int accumulate,
                                float* __restrict__ dh2) {
  const int n = sp.n(), nch = sp.nch();
  const long long items = W * n * n * F;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(it % F);
    long long t = it / F;
    const int j = (int)(t % n);
    t /= n;
    const int i = (int)(t % n);
    const long long w = t / n;
    const int s = sp.chan_of(i);
    const float v = dg2[((w * n + j) * nch + s) * (long long)F + f] / (float)(sp.hi(s) - sp.lo(s));
    dh2[it] = accumulate ? dh2[it] + v : v;
  }
}

// WT[nn][k] = Wm[k][nn]   (Wm: K rows of stride ldw, N columns)
__global__ void k_transpose(const float* __restrict__ Wm, int K, int N, int ldw, float* __restrict__ WT) {
  const long long items = (long long)K * N;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(it % K);
    const int nn = (int)(it / K);
    WT[(long long)nn * K + k] = Wm[(long long)k * ldw + nn];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dW = X^T dZ :  out[k][nn] = sum_r X[r*ldx + k] * Z[r*ldz + nn],  k < Kd, nn < Nd, r < R.
// The row range is split into S chunks; chunk s writes part[s][k][nn]; k_reduce_partials adds them in order.
// ---------------------------------------------------------------------------------------------------------------
#ifdef JAQMC_HOST_EMU
__global__ void k_gemm_tn(const float* __restrict__ X, int ldx, const float* __restrict__ Z, int ldz, long long R, int Kd,
                          int Nd, int S, float* __restrict__ part) {
  const long long items = (long long)S * Kd * Nd;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int nn = (int)(it % Nd);
    long long t = it / Nd;
    const int k = (int)(t % Kd);
    const int s = (int)(t / Kd);
    const long long r0 = R * s / S, r1 = R * (s + 1) / S;
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc = fmaf(X[r * ldx + k], Z[r * ldz + nn], acc);
    part[it] = acc;
  }
}
#else
#define GT 64   // output tile
#define GR 16   // rows per stage
__global__ void __launch_bounds__(256) k_gemm_tn(const float* __restrict__ X, int ldx, const float* __restrict__ Z, int ldz,
                                                 long long R, int Kd, int Nd, int S, float* __restrict__ part) {
  __shared__ float Xs[GR][GT + 4], Zs[GR][GT + 4];
  const int k0 = blockIdx.x * GT, n0 = blockIdx.y * GT, s = blockIdx.z;
  const long long r0 = R * s / S, r1 = R * (s + 1) / S;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // thread -> 4 x 4 outputs (k = k0 + 4ty.., n = n0 + 4tx..)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;   // staging: row lr, columns lc..lc+3
  for (long long rb = r0; rb < r1; rb += GR) {
    const long long r = rb + lr;
    const bool okr = r < r1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      Xs[lr][lc + q] = (okr && k0 + lc + q < Kd) ? X[r * ldx + k0 + lc + q] : 0.f;
      Zs[lr][lc + q] = (okr && n0 + lc + q < Nd) ? Z[r * ldz + n0 + lc + q] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < GR; ++rr) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[rr][4 * ty]);
      const float4 zv = *reinterpret_cast<const float4*>(&Zs[rr][4 * tx]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, za[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], za[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* o = part + (long long)s * Kd * Nd;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + 4 * ty + i, nn = n0 + 4 * tx + j;
      if (k < Kd && nn < Nd) o[(long long)k * Nd + nn] = acc[i][j];
    }
}
#endif

#ifndef JAQMC_HOST_EMU
// dW = X^T dZ on the tensor cores (r2): mma.sync.m16n8k8 TF32, operands split x = hi + lo by truncation (three MMAs per
// product, FP32 accumulation), same grid and partial layout as k_gemm_tn.  A = X^T is read from the row-major shared tile
// as A[m = feature][k = row]; the row stride 72 (= 8 mod 32) makes both fragment loads conflict-free.
// Block tile 64 x 64, 8 warps = 4 feature tiles x 2 halves of the output columns, 32 rows per stage, the next stage's
// global loads in flight in registers while the current one is multiplied.
#define GTM_R 32
#define GTM_LD 72
__device__ __forceinline__ void gtm_split(float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void gtm_mma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void __launch_bounds__(256) k_gemm_tn_mma(const float* __restrict__ X, int ldx, const float* __restrict__ Z, int ldz,
                                                     long long R, int Kd, int Nd, int S, float* __restrict__ part) {
  __shared__ float Xs[GTM_R][GTM_LD], Zs[GTM_R][GTM_LD];
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64, s = blockIdx.z;
  const long long r0 = R * s / S, r1 = R * (s + 1) / S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int mt = warp & 3, nh = warp >> 2;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  // staging: thread -> rows lr, lr + 16; columns lc .. lc + 3 of each operand
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;
  const bool vx = (ldx % 4 == 0) && (k0 + lc + 3 < Kd) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const bool vz = (ldz % 4 == 0) && (n0 + lc + 3 < Nd) && ((reinterpret_cast<uintptr_t>(Z) & 15) == 0);
  float4 px[2], pz[2];
  auto fetch = [&](long long rb) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = rb + lr + 16 * h;
      const bool okr = r < r1;
      if (okr && vx) px[h] = *reinterpret_cast<const float4*>(X + r * ldx + k0 + lc);
      else {
        px[h].x = (okr && k0 + lc + 0 < Kd) ? X[r * ldx + k0 + lc + 0] : 0.f;
        px[h].y = (okr && k0 + lc + 1 < Kd) ? X[r * ldx + k0 + lc + 1] : 0.f;
        px[h].z = (okr && k0 + lc + 2 < Kd) ? X[r * ldx + k0 + lc + 2] : 0.f;
        px[h].w = (okr && k0 + lc + 3 < Kd) ? X[r * ldx + k0 + lc + 3] : 0.f;
      }
      if (okr && vz) pz[h] = *reinterpret_cast<const float4*>(Z + r * ldz + n0 + lc);
      else {
        pz[h].x = (okr && n0 + lc + 0 < Nd) ? Z[r * ldz + n0 + lc + 0] : 0.f;
        pz[h].y = (okr && n0 + lc + 1 < Nd) ? Z[r * ldz + n0 + lc + 1] : 0.f;
        pz[h].z = (okr && n0 + lc + 2 < Nd) ? Z[r * ldz + n0 + lc + 2] : 0.f;
        pz[h].w = (okr && n0 + lc + 3 < Nd) ? Z[r * ldz + n0 + lc + 3] : 0.f;
      }
    }
  };
  fetch(r0);
  for (long long rb = r0; rb < r1; rb += GTM_R) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(&Xs[lr + 16 * h][lc]) = px[h];
      *reinterpret_cast<float4*>(&Zs[lr + 16 * h][lc]) = pz[h];
    }
    __syncthreads();
    if (rb + GTM_R < r1) fetch(rb + GTM_R);
#pragma unroll
    for (int ks = 0; ks < GTM_R / 8; ++ks) {
      unsigned ah[4], al[4];
      gtm_split(Xs[8 * ks + t][16 * mt + g], ah[0], al[0]);
      gtm_split(Xs[8 * ks + t][16 * mt + g + 8], ah[1], al[1]);
      gtm_split(Xs[8 * ks + t + 4][16 * mt + g], ah[2], al[2]);
      gtm_split(Xs[8 * ks + t + 4][16 * mt + g + 8], ah[3], al[3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        unsigned bh[2], bl[2];
        gtm_split(Zs[8 * ks + t][32 * nh + 8 * nt + g], bh[0], bl[0]);
        gtm_split(Zs[8 * ks + t + 4][32 * nh + 8 * nt + g], bh[1], bl[1]);
        gtm_mma(acc[nt], al, bh);
        gtm_mma(acc[nt], ah, bl);
        gtm_mma(acc[nt], ah, bh);
      }
    }
    __syncthreads();
  }
  float* o = part + (long long)s * Kd * Nd;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int nn = n0 + 32 * nh + 8 * nt + 2 * t;
    const int ka = k0 + 16 * mt + g, kb = ka + 8;
    if (ka < Kd) {
      if (nn < Nd) o[(long long)ka * Nd + nn] = acc[nt][0];
      if (nn + 1 < Nd) o[(long long)ka * Nd + nn + 1] = acc[nt][1];
    }
    if (kb < Kd) {
      if (nn < Nd) o[(long long)kb * Nd + nn] = acc[nt][2];
      if (nn + 1 < Nd) o[(long long)kb * Nd + nn + 1] = acc[nt][3];
    }
  }
}
#endif

int grid_for(long long items);
// fixed-order sum over S partials; wide = parallel over the partial index (device only)
int launch_reduce_partials(const float* part, int S, long long stride, long long P, float* out, cudaStream_t st);

int grid_for(long long items) {
  int g = jq_cdiv(items, 256);
  if (g > 148 * 32) g = 148 * 32;
  return g < 1 ? 1 : g;
}

int launch_reduce_partials(const float* part, int S, long long stride, long long P, float* out, cudaStream_t st) {
#ifndef JAQMC_HOST_EMU
  if (S >= 16) {
    long long g = jq_cdiv(P, 32);
    if (g > 148 * 16) g = 148 * 16;
    JQ_LAUNCH(k_reduce_partials_wide, dim3((unsigned)g), dim3(256), 0, st, part, S, stride, P, out);
    JQ_CHECK_LAUNCH();
    return JQ_OK;
  }
#endif
  JQ_LAUNCH(k_reduce_partials, dim3(grid_for(P)), dim3(256), 0, st, part, S, stride, P, out);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// Row chunks of the reduction GEMM: about four blocks per SM, at least 64 rows per chunk, and the partial results
// [S][Kd][Nd] must fit the scratch area (GEMM_PART_FLOATS, carved in bwd_carve)
#define GEMM_PART_TILES 96
int gemm_splits(long long R, int Kd, int Nd, long long part_floats) {
  const long long tiles = (long long)jq_cdiv(Kd, 64) * jq_cdiv(Nd, 64);
  long long S = (148 * 4 + tiles - 1) / tiles;
  if (S > (R + 63) / 64) S = (R + 63) / 64;
  if (S * Kd * Nd > part_floats) S = part_floats / ((long long)Kd * Nd);
  return S < 1 ? 1 : (int)S;
}

// out[Kd][Nd] = X^T Z (row stride of out: ldo)
int launch_gemm_tn(const float* X, int ldx, const float* Z, int ldz, long long R, int Kd, int Nd, float* part,
                   long long part_floats, float* out, int ldo, cudaStream_t st) {
  if (R <= 0 || Kd <= 0 || Nd <= 0) return JQ_OK;
  const int S = gemm_splits(R, Kd, Nd, part_floats);
  jq_prof_work(2.0 * (double)R * Kd * Nd, 4.0 * (double)R * (Kd + Nd));
#ifdef JAQMC_HOST_EMU
  JQ_LAUNCH(k_gemm_tn, dim3(grid_for((long long)S * Kd * Nd)), dim3(256), 0, st, X, ldx, Z, ldz, R, Kd, Nd, S, part);
#else
  static const bool simt = getenv("JAQMC_B200_GEMM_TN_SIMT") != nullptr;   // A/B switch: the FP32 CUDA-core kernel
  if (simt) JQ_LAUNCH(k_gemm_tn, dim3(jq_cdiv(Kd, 64), jq_cdiv(Nd, 64), S), dim3(256), 0, st, X, ldx, Z, ldz, R, Kd, Nd, S, part);
  else JQ_LAUNCH(k_gemm_tn_mma, dim3(jq_cdiv(Kd, 64), jq_cdiv(Nd, 64), S), dim3(256), 0, st, X, ldx, Z, ldz, R, Kd, Nd, S, part);
#endif
  JQ_CHECK_LAUNCH();
  // reduce into out (possibly strided rows): ldo == Nd for every caller but the layer-1 kernel, which is contiguous too
  JQ_REQUIRE(ldo == Nd, JQ_ERR_INVALID_ARGUMENT, "gemm_tn: strided output is not supported");
  return launch_reduce_partials(part, S, (long long)Kd * Nd, (long long)Kd * Nd, out, st);
}

// column sums: out[f] = sum_r Z[r][f]  == gemm_tn with X = ones; done as a two-stage reduction
__global__ void k_colsum_partial(const float* __restrict__ Z, long long R, int F, int S, float* __restrict__ part) {
  const long long items = (long long)S * F;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(it % F);
    const int s = (int)(it / F);
    const long long r0 = R * s / S, r1 = R * (s + 1) / S;
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc += Z[r * F + f];
    part[it] = acc;
  }
}

int launch_colsum(const float* Z, long long R, int F, float* part, long long part_floats, float* out, cudaStream_t st) {
  long long S = (R + 63) / 64;                       // 64 rows per item: enough items to fill the machine
  const long long smax = (148LL * 2048) / F;         // ... but no more than ~8 items per resident thread
  if (S > smax) S = smax;
  if (S * F > part_floats) S = part_floats / F;
  if (S < 1) S = 1;
  JQ_LAUNCH(k_colsum_partial, dim3(grid_for(S * F)), dim3(256), 0, st, Z, R, F, (int)S, part);
  JQ_CHECK_LAUNCH();
  return launch_reduce_partials(part, (int)S, (long long)F, (long long)F, out, st);
}

// value-only dense through the forward launcher: out[G][N] = act(x [G][k0] W[k0][N] (+ x2 W2) + bias + cadd) (+ residual)
int dense_value(const float* x, int k0, const float* w, int ldw, const float* x2, int k1, const float* w2,
                const float* bias, const float* cadd, int n_per_walker, int act, const float* res, int res_mode,
                float* out, int N, long long G, float* wscr, int k0_valid, cudaStream_t st) {
  JqDenseArgs a;
  memset(&a, 0, sizeof(a));
  a.src0 = x;
  a.k0 = k0;
  a.k0_valid = k0_valid;
  a.w0 = w;
  a.ldw = ldw;
  a.src1 = x2;
  a.k1 = k1;
  a.w1 = w2;
  a.bias = bias;
  a.cadd = cadd;
  a.out = out;
  a.N = N;
  a.C = 1;
  a.n_sub = a.n_tot = n_per_walker;
  a.G = G;
  a.act = act;
  a.res = res;
  a.res_mode = res_mode;
  a.wscratch = wscr;
  return jq_launch_dense(a, st);
}

struct BwdBufs {
  float *ae, *x1, *m, *cadd, *g2[JQ_MAX_LAYERS], *h[JQ_MAX_LAYERS + 1], *h2[JQ_MAX_LAYERS + 1];
  float *orb, *env, *M, *minv, *dsign, *dlogabs, *wdet, *dorb, *denv;
  float *dh_a, *dh_b, *dz, *dres, *dzw, *dm, *dg2, *dh2_a, *dh2_b, *dz2, *dres2;
  float *wt, *part, *wscr;
  long long part_floats;
};

void bwd_carve(const FermiDims& d, long long W, JqArena& ar, BwdBufs* b) {
  const long long n = d.n, nn = n * n, DN = (long long)d.D * n;
  b->ae = ar.take<float>(W * n * d.f1);
  b->x1 = ar.take<float>(W * n * d.in1p);
  b->m = ar.take<float>(W * d.nch * d.d1max);
  b->cadd = ar.take<float>(W * d.d1max);
  for (int l = 0; l < d.L; ++l) b->g2[l] = ar.take<float>(W * n * d.nch * d.d2max);
  for (int l = 0; l <= d.L; ++l) b->h[l] = (l == 0) ? nullptr : ar.take<float>(W * n * d.d1max);
  for (int l = 0; l < d.L; ++l) b->h2[l] = ar.take<float>(W * nn * d.d2max);
  b->orb = ar.take<float>(W * n * DN);
  b->env = ar.take<float>(W * n * DN);
  b->M = ar.take<float>(W * n * DN);
  b->minv = ar.take<float>(W * DN * n);
  b->dsign = ar.take<float>(W * d.D);
  b->dlogabs = ar.take<float>(W * d.D);
  b->wdet = ar.take<float>(W * d.D);
  b->dorb = ar.take<float>(W * n * DN);
  b->denv = ar.take<float>(W * n * DN);
  b->dh_a = ar.take<float>(W * n * d.d1max);
  b->dh_b = ar.take<float>(W * n * d.d1max);
  b->dz = ar.take<float>(W * n * d.d1max);
  b->dres = ar.take<float>(W * n * d.d1max);
  b->dzw = ar.take<float>(W * d.d1max);
  b->dm = ar.take<float>(W * d.nch * d.d1max);
  b->dg2 = ar.take<float>(W * n * d.nch * d.d2max);
  b->dh2_a = ar.take<float>(W * nn * d.d2max);
  b->dh2_b = ar.take<float>(W * nn * d.d2max);
  b->dz2 = ar.take<float>(W * nn * d.d2max);
  b->dres2 = ar.take<float>(W * nn * d.d2max);
  const long long kmax = (long long)d.d1max * (1 + d.nch) + (long long)d.nch * d.d2max > d.in1p
                             ? (long long)d.d1max * (1 + d.nch) + (long long)d.nch * d.d2max
                             : d.in1p;
  const long long nmax = d.d1max > DN ? d.d1max : DN;
  b->wt = ar.take<float>(kmax * nmax);
  b->part_floats = GEMM_PART_TILES * kmax * nmax > 256 * 2 * n * d.A * d.D * 2 ? GEMM_PART_TILES * kmax * nmax
                                                                                 : 256 * 2 * n * d.A * d.D * 2;
  b->part = ar.take<float>(b->part_floats);
  b->wscr = ar.take<float>(jq_dense_tc_scratch_floats((int)kmax, (int)nmax));
}

}  // namespace

extern "C" size_t jaqmc_b200_ferminet_vjp_workspace_bytes(const jaqmc_ferminet_config* c, int64_t n_walkers) {
  FermiDims d;
  if (!c || n_walkers < 0 || jq_fermi_dims(c, 0, 4, 4, &d) != JQ_OK) return 0;
  JqArena ar(nullptr, 0);
  BwdBufs b;
  bwd_carve(d, n_walkers, ar, &b);
  return ar.off + 256;
}

extern "C" int jaqmc_b200_ferminet_logpsi_vjp(const jaqmc_ferminet_config* c, const jaqmc_ferminet_params* p,
                                              const jaqmc_system* sys, const float* electrons, int64_t n_walkers,
                                              const float* cotangent, const jaqmc_ferminet_grads* grads, float* logpsi,
                                              float* sign, void* workspace, size_t workspace_bytes,
                                              jaqmc_stream_t stream) {
  JQ_REQUIRE(c && p && sys && grads, JQ_ERR_INVALID_ARGUMENT, "logpsi_vjp: null descriptor");
  FermiDims d;
  int rc = jq_fermi_dims(c, 0, 4, 4, &d);
  if (rc) return rc;
  JQ_REQUIRE(!c->use_last_layer, JQ_ERR_UNSUPPORTED, "logpsi_vjp: use_last_layer is not implemented");
  JQ_REQUIRE(d.n <= BW_NMAX, JQ_ERR_UNSUPPORTED, "logpsi_vjp: at most %d electrons", BW_NMAX);
  JQ_REQUIRE(c->envelope_type != JAQMC_ENVELOPE_DIAGONAL, JQ_ERR_UNSUPPORTED, "logpsi_vjp: diagonal envelope is not implemented");
  JQ_REQUIRE(sys->atoms && sys->n_atoms == d.A, JQ_ERR_INVALID_ARGUMENT, "logpsi_vjp: system/atoms mismatch");
  JQ_REQUIRE(n_walkers >= 0 && (n_walkers == 0 || (electrons && cotangent)), JQ_ERR_INVALID_ARGUMENT, "logpsi_vjp: null buffer");
  const long long W = n_walkers;
  if (W == 0) return JQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  JqArena ar(workspace, workspace_bytes);
  BwdBufs b;
  bwd_carve(d, W, ar, &b);
  JQ_REQUIRE(workspace && ar.ok(), JQ_ERR_WORKSPACE_TOO_SMALL, "logpsi_vjp: workspace %zu < %zu bytes (no walker tiling on this path)",
             workspace_bytes, ar.off);
  const int n = d.n, L = d.L, nch = d.nch;
  const long long G = W * n, G2 = W * n * n;
  const int DN = d.D * n;
  const bool split = c->orbitals_spin_split && nch == 2;
  const int nchan = split ? 2 : 1;
  for (int l = 0; l < L; ++l) {
    JQ_REQUIRE(p->single_kernel[l] && p->single_bias[l] && grads->single_kernel[l] && grads->single_bias[l],
               JQ_ERR_INVALID_ARGUMENT, "logpsi_vjp: null single-stream parameter / gradient buffer in layer %d", l);
    JQ_REQUIRE(l == L - 1 || (p->double_kernel[l] && p->double_bias[l] && grads->double_kernel[l] && grads->double_bias[l]),
               JQ_ERR_INVALID_ARGUMENT, "logpsi_vjp: null double-stream parameter / gradient buffer in layer %d", l);
  }
  for (int s = 0; s < nchan; ++s) {
    JQ_REQUIRE(p->orbital_kernel[s] && grads->orbital_kernel[s], JQ_ERR_INVALID_ARGUMENT, "logpsi_vjp: null orbital kernel");
    JQ_REQUIRE(c->envelope_type == JAQMC_ENVELOPE_NULL || (p->env_pi[s] && p->env_sigma[s] && grads->env_pi[s] && grads->env_sigma[s]),
               JQ_ERR_INVALID_ARGUMENT, "logpsi_vjp: null envelope parameter / gradient buffer");
  }

  // ================================================= forward (value path, activations kept) ======================
  // h2[0] = ee features, ae = atom features
  if ((rc = jq_launch_mol_features(electrons, sys->atoms, (int)W, d.sp, d.A, 0, 0, 0, b.ae, b.h2[0], st))) return rc;
  int d1[JQ_MAX_LAYERS + 1], d2[JQ_MAX_LAYERS + 1];
  bool res1[JQ_MAX_LAYERS], res2[JQ_MAX_LAYERS];
  d1[0] = d.f1;
  d2[0] = d.fee;
  for (int l = 0; l < L; ++l) {
    d1[l + 1] = d.d1[l];
    d2[l + 1] = (l < L - 1) ? d.d2[l] : d2[l];
    res1[l] = (l > 0) && d1[l] == d1[l + 1];
    res2[l] = (l < L - 1) && d2[l] == d2[l + 1];
  }
  for (int l = 0; l < L; ++l) {
    const int fg = nch * d2[l];
    if ((rc = jq_launch_pair_mean(b.h2[l], b.g2[l], (int)W, d.sp, d2[l], 0, st))) return rc;
    if (l == 0) {
      if ((rc = jq_launch_concat_layer1(b.ae, b.g2[0], b.x1, (int)W, d.sp, d.f1, fg, 0, d.in1p, st))) return rc;
      if ((rc = dense_value(b.x1, d.in1p, p->single_kernel[0], 0, nullptr, 0, nullptr, p->single_bias[0], nullptr, n, 1,
                            nullptr, 0, b.h[1], d1[1], G, b.wscr, d.in1, st)))
        return rc;
    } else {
      if ((rc = jq_launch_spin_mean(b.h[l], b.m, (int)W, d.sp, 1, d1[l], st))) return rc;
      if ((rc = dense_value(b.m, nch * d1[l], p->single_kernel[l] + (size_t)d1[l] * d1[l + 1], 0, nullptr, 0, nullptr, nullptr,
                            nullptr, 1, 0, nullptr, 0, b.cadd, d1[l + 1], W, b.wscr, 0, st)))
        return rc;
      if ((rc = dense_value(b.h[l], d1[l], p->single_kernel[l], 0, b.g2[l], fg,
                            p->single_kernel[l] + (size_t)d1[l] * (1 + nch) * d1[l + 1], p->single_bias[l], b.cadd, n, 1,
                            res1[l] ? b.h[l] : nullptr, res1[l] ? 1 : 0, b.h[l + 1], d1[l + 1], G, b.wscr, 0, st)))
        return rc;
    }
    if (l < L - 1) {
      if ((rc = dense_value(b.h2[l], d2[l], p->double_kernel[l], 0, nullptr, 0, nullptr, p->double_bias[l], nullptr, n * n, 1,
                            res2[l] ? b.h2[l] : nullptr, res2[l] ? 1 : 0, b.h2[l + 1], d2[l + 1], G2, b.wscr, 0, st)))
        return rc;
    }
  }
  // orbitals, envelope, determinants
  const int hid = d1[L];
  for (int s = 0; s < nchan; ++s) {
    JqDenseArgs a;
    memset(&a, 0, sizeof(a));
    a.src0 = b.h[L];
    a.k0 = hid;
    a.w0 = p->orbital_kernel[s];
    a.out = b.orb;
    a.wscratch = b.wscr;
    a.N = DN;
    a.C = 1;
    a.n_tot = n;
    a.j0 = split ? d.sp.lo(s) : 0;
    a.n_sub = split ? d.sp.hi(s) - d.sp.lo(s) : n;
    a.G = W * a.n_sub;
    if ((rc = jq_launch_dense(a, st))) return rc;
  }
  JqEnvelopeArgs env;
  env.type = c->envelope_type;
  env.pi[0] = p->env_pi[0];
  env.sigma[0] = p->env_sigma[0];
  env.pi[1] = split ? p->env_pi[1] : nullptr;
  env.sigma[1] = split ? p->env_sigma[1] : nullptr;
  const long long items = G * DN;
  const bool has_env = c->envelope_type != JAQMC_ENVELOPE_NULL;
  const float* Mv = b.orb;
  if (has_env) {
    JQ_LAUNCH(k_envelope_values, dim3(grid_for(items)), dim3(256), 0, st, electrons, sys->atoms, env, items, d.sp, d.A, d.D, b.env);
    JQ_CHECK_LAUNCH();
    JQ_LAUNCH(k_mul, dim3(grid_for(items)), dim3(256), 0, st, b.orb, b.env, b.M, items);
    JQ_CHECK_LAUNCH();
    Mv = b.M;
  }
#ifndef JAQMC_HOST_EMU
  if (n <= 16) {
    auto k_minv_halfwarp = k_minv_small<16>;
    JQ_LAUNCH(k_minv_halfwarp, dim3((unsigned)jq_cdiv(W * d.D, 16)), dim3(256), 0, st, Mv, W * d.D, n, d.D, b.minv, b.dsign, b.dlogabs);
  } else {
    auto k_minv_warp = k_minv_small<32>;
    JQ_LAUNCH(k_minv_warp, dim3((unsigned)jq_cdiv(W * d.D, 8)), dim3(256), 0, st, Mv, W * d.D, n, d.D, b.minv, b.dsign, b.dlogabs);
  }
#else
  JQ_LAUNCH(k_minv, dim3(grid_for(W * d.D)), dim3(64), 0, st, Mv, W * d.D, n, d.D, b.minv, b.dsign, b.dlogabs);
#endif
  JQ_CHECK_LAUNCH();
  JQ_LAUNCH(k_det_weights, dim3(grid_for(W)), dim3(128), 0, st, b.dsign, b.dlogabs, cotangent, W, d.D, b.wdet, logpsi, sign);
  JQ_CHECK_LAUNCH();

  // ================================================= reverse =====================================================
  JQ_LAUNCH(k_head_bwd, dim3(grid_for(items)), dim3(256), 0, st, b.wdet, b.minv, b.orb, has_env ? b.env : nullptr, items, n,
            d.D, b.dorb, has_env ? b.denv : nullptr);
  JQ_CHECK_LAUNCH();
  if (has_env) {
    const int chunks = W >= 256 ? 256 : (W >= 64 ? 64 : (int)W);   // walker chunks: ~230 k items at N2 (r2: 64 chunks left the kernel at 224 blocks, 0.53 ms)
    const long long P = (long long)nchan * n * d.A * d.D;
    float* part_pi = b.part;
    float* part_sg = b.part + (long long)chunks * P;
    JQ_LAUNCH(k_env_param_grad, dim3(grid_for((long long)chunks * P)), dim3(128), 0, st, b.denv, electrons, sys->atoms, env, W,
              d.sp, d.A, d.D, chunks, part_pi, part_sg);
    JQ_CHECK_LAUNCH();
    for (int s = 0; s < nchan; ++s) {
      const long long Ps = (long long)n * d.A * d.D;
      // partials are laid out [chunk][channel][o][I][d]: reduce one channel at a time with stride P
      if ((rc = launch_reduce_partials(part_pi + s * Ps, chunks, P, Ps, grads->env_pi[s], st))) return rc;
      if ((rc = launch_reduce_partials(part_sg + s * Ps, chunks, P, Ps, grads->env_sigma[s], st))) return rc;
    }
  }
  // orbital kernels: dK_s = h_L(rows of channel s)^T dorb(rows of channel s); dh_L = dorb K_s^T
  float* dh = b.dh_a;
  float* dh_other = b.dh_b;
  for (int s = 0; s < nchan; ++s) {
    const int lo = split ? d.sp.lo(s) : 0, ns = split ? d.sp.hi(s) - d.sp.lo(s) : n;
    if (ns == n) {
      if ((rc = launch_gemm_tn(b.h[L], hid, b.dorb, DN, G, hid, DN, b.part, b.part_floats, grads->orbital_kernel[s], DN, st))) return rc;
    } else {
      // rows of one spin channel (electrons [lo, lo + ns) of every walker), compacted; dz and M are free here
      float* hc = b.dz;
      float* dc = b.M;
      JQ_LAUNCH(k_gather_channel, dim3(grid_for(W * ns * hid)), dim3(256), 0, st, b.h[L], W, n, lo, ns, hid, hc);
      JQ_CHECK_LAUNCH();
      JQ_LAUNCH(k_gather_channel, dim3(grid_for(W * ns * DN)), dim3(256), 0, st, b.dorb, W, n, lo, ns, DN, dc);
      JQ_CHECK_LAUNCH();
      if ((rc = launch_gemm_tn(hc, hid, dc, DN, W * ns, hid, DN, b.part, b.part_floats, grads->orbital_kernel[s], DN, st))) return rc;
    }
    // dh_L rows of this channel = dorb K_s^T : K_s (hid, DN) -> K_s^T (DN, hid)
    JQ_LAUNCH(k_transpose, dim3(grid_for((long long)hid * DN)), dim3(256), 0, st, p->orbital_kernel[s], hid, DN, DN, b.wt);
    JQ_CHECK_LAUNCH();
    JqDenseArgs a;
    memset(&a, 0, sizeof(a));
    a.src0 = b.dorb;
    a.k0 = DN;
    a.w0 = b.wt;
    a.out = dh;
    a.wscratch = b.wscr;
    a.N = hid;
    a.C = 1;
    a.n_tot = n;
    a.j0 = lo;
    a.n_sub = ns;
    a.G = W * ns;
    if ((rc = jq_launch_dense(a, st))) return rc;
  }
  // layers, top down.  dh = adjoint of h[l + 1]
  bool have_dh2 = false;   // dh2 (adjoint of h2[l]) accumulated so far
  float* dh2 = b.dh2_a;
  float* dh2_other = b.dh2_b;
  for (int l = L - 1; l >= 0; --l) {
    const int N1 = d1[l + 1], K1 = d1[l], fg = nch * d2[l];
    const long long cnt = G * N1;
    JQ_LAUNCH(k_tanh_bwd, dim3(grid_for(cnt)), dim3(256), 0, st, dh, b.h[l + 1], res1[l] ? b.h[l] : nullptr, res1[l] ? 1 : 0, cnt,
              b.dz, res1[l] ? b.dres : nullptr);
    JQ_CHECK_LAUNCH();
    if ((rc = launch_colsum(b.dz, G, N1, b.part, b.part_floats, grads->single_bias[l], st))) return rc;
    if (l == 0) {
      // dW_0 = x1^T dz over the in1 rows that exist (x1 rows are in1p wide, zero padded)
      if ((rc = launch_gemm_tn(b.x1, d.in1p, b.dz, N1, G, d.in1, N1, b.part, b.part_floats, grads->single_kernel[0], N1, st))) return rc;
      break;   // the input features carry no parameters
    }
    float* gW = grads->single_kernel[l];
    // recompute this layer's spin means (b.m is reused by every layer in the forward pass)
    if ((rc = jq_launch_spin_mean(b.h[l], b.m, (int)W, d.sp, 1, K1, st))) return rc;
    JQ_LAUNCH(k_walker_sum, dim3(grid_for(W * N1)), dim3(256), 0, st, b.dz, W, n, N1, b.dzw);
    JQ_CHECK_LAUNCH();
    if ((rc = launch_gemm_tn(b.h[l], K1, b.dz, N1, G, K1, N1, b.part, b.part_floats, gW, N1, st))) return rc;
    if ((rc = launch_gemm_tn(b.m, nch * K1, b.dzw, N1, W, nch * K1, N1, b.part, b.part_floats, gW + (size_t)K1 * N1, N1, st))) return rc;
    if ((rc = launch_gemm_tn(b.g2[l], fg, b.dz, N1, G, fg, N1, b.part, b.part_floats, gW + (size_t)K1 * (1 + nch) * N1, N1, st))) return rc;
    // dX = dz W^T, block by block of the transposed kernel WT [N1][fan_in]
    const int fan_in = K1 * (1 + nch) + fg;
    JQ_LAUNCH(k_transpose, dim3(grid_for((long long)fan_in * N1)), dim3(256), 0, st, p->single_kernel[l], fan_in, N1, N1, b.wt);
    JQ_CHECK_LAUNCH();
    // dh[l] = dz Wh^T (+ residual path)
    if ((rc = dense_value(b.dz, N1, b.wt, fan_in, nullptr, 0, nullptr, nullptr, nullptr, n, 0, res1[l] ? b.dres : nullptr,
                          res1[l] ? 2 : 0, dh_other, K1, G, b.wscr, 0, st)))
      return rc;
    // dm = dzw Wm^T, spread back over the electrons of each channel
    if ((rc = dense_value(b.dzw, N1, b.wt + K1, fan_in, nullptr, 0, nullptr, nullptr, nullptr, 1, 0, nullptr, 0, b.dm, nch * K1,
                          W, b.wscr, 0, st)))
      return rc;
    JQ_LAUNCH(k_spin_mean_bwd, dim3(grid_for(G * K1)), dim3(256), 0, st, b.dm, W, d.sp, K1, dh_other);
    JQ_CHECK_LAUNCH();
    // dg2 = dz Wg^T -> adjoint of h2[l]
    if ((rc = dense_value(b.dz, N1, b.wt + (size_t)K1 * (1 + nch), fan_in, nullptr, 0, nullptr, nullptr, nullptr, n, 0, nullptr,
                          0, b.dg2, fg, G, b.wscr, 0, st)))
      return rc;
    JQ_LAUNCH(k_pair_mean_bwd, dim3(grid_for(G2 * d2[l])), dim3(256), 0, st, b.dg2, W, d.sp, d2[l], have_dh2 ? 1 : 0, dh2);
    JQ_CHECK_LAUNCH();
    have_dh2 = true;
    {
      float* t = dh;
      dh = dh_other;
      dh_other = t;
    }
    // two-electron layer l - 1 maps h2[l-1] -> h2[l]: push dh2 (adjoint of h2[l]) through it
    const int lp = l - 1;
    const int N2 = d2[lp + 1], K2 = d2[lp];
    const long long cnt2 = G2 * N2;
    JQ_LAUNCH(k_tanh_bwd, dim3(grid_for(cnt2)), dim3(256), 0, st, dh2, b.h2[lp + 1], res2[lp] ? b.h2[lp] : nullptr,
              res2[lp] ? 1 : 0, cnt2, b.dz2, res2[lp] ? b.dres2 : nullptr);
    JQ_CHECK_LAUNCH();
    if ((rc = launch_colsum(b.dz2, G2, N2, b.part, b.part_floats, grads->double_bias[lp], st))) return rc;
    if ((rc = launch_gemm_tn(b.h2[lp], K2, b.dz2, N2, G2, K2, N2, b.part, b.part_floats, grads->double_kernel[lp], N2, st))) return rc;
    if (lp > 0) {   // h2[0] are the input features: nothing further below
      JQ_LAUNCH(k_transpose, dim3(grid_for((long long)K2 * N2)), dim3(256), 0, st, p->double_kernel[lp], K2, N2, N2, b.wt);
      JQ_CHECK_LAUNCH();
      if ((rc = dense_value(b.dz2, N2, b.wt, K2, nullptr, 0, nullptr, nullptr, nullptr, n * n, 0, res2[lp] ? b.dres2 : nullptr,
                            res2[lp] ? 2 : 0, dh2_other, K2, G2, b.wscr, 0, st)))
        return rc;
      float* t = dh2;
      dh2 = dh2_other;
      dh2_other = t;
    } else {
      have_dh2 = false;
    }
  }
  return JQ_OK;
}
