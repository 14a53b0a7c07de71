// Shared output head of the molecular wavefunctions: orbital projection -> envelope product -> multi-determinant
// LogDet, plus the optional electron-electron Jastrow factor added to log|psi|.
//
// Reference semantics: wavefunction/output/orbital.py:59-78 (per-spin DenseGeneral), output/envelope.py:98-140,
// output/logdet.py:53-79, wavefunction/jastrow.py:47-122 (SimpleEEJastrow), and the composition in
// app/molecule/wavefunction/{ferminet.py:76-96, lapnet.py:117-135, psiformer.py:137-167}.
#include "wf.cuh"

// ------------------------------------------------------------------------------------------------
// SimpleEEJastrow with its gradient and Laplacian in closed form:
//   J = sum_{i<j} f(r_ij),  f(r) = -c a^2 / (a + r),  c = 1/4, a = alpha_par for parallel spins, c = 1/2, a = alpha_anti
//   f'(r) = c a^2 / (a + r)^2,  f''(r) = -2 c a^2 / (a + r)^3,
//   grad_i J = sum_j f'(r_ij) (r_i - r_j)/r_ij,   lap J = sum_i sum_{j != i} (f'' + 2 f'/r)
// (what the reference's interpreter obtains from the norm / div rules on the Local2 r_ee tensor).
// A block owns JT walkers: phase 1 one item per (walker, electron), phase 2 one item per walker.
// extra [W][C]: component 0 value, 1..3n gradient, 3n+1 Laplacian (C == 1: value only).
// ------------------------------------------------------------------------------------------------
#define JT 8
__global__ void k_jastrow(const float* __restrict__ el, const float* __restrict__ alpha_par,
                          const float* __restrict__ alpha_anti, int W, JqSpins sp, int track,
                          float* __restrict__ extra) {
  JQ_DYN_SMEM(float, part);  // [JT][n][2]
  const int n = sp.n();
  const int C = track ? 3 * n + 2 : 1;
  const int w0 = blockIdx.x * JT;
  const int nw = (W - w0 < JT) ? W - w0 : JT;
  const float ap = alpha_par[0], aa = alpha_anti[0];
  for (int q = threadIdx.x; q < nw * n; q += blockDim.x) {
    const int t = q / n, i = q % n;
    const long long w = w0 + t;
    const float* e = el + w * n * 3;
    const float xi = e[i * 3], yi = e[i * 3 + 1], zi = e[i * 3 + 2];
    float val = 0.f, lap = 0.f, g[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      const bool same = sp.chan_of(i) == sp.chan_of(j) || sp.nch() == 1;
      const float c = same ? 0.25f : 0.5f, a = same ? ap : aa;
      const float dx = xi - e[j * 3], dy = yi - e[j * 3 + 1], dz = zi - e[j * 3 + 2];
      const float r = sqrtf(dx * dx + dy * dy + dz * dz);
      const float inv = 1.0f / (a + r);
      const float ca2 = c * a * a;
      const float f1 = ca2 * inv * inv;
      val -= 0.5f * ca2 * inv;  // each pair is visited from both ends
      if (track) {
        const float rinv = 1.0f / r;
        const float s = f1 * rinv;
        g[0] = fmaf(s, dx, g[0]);
        g[1] = fmaf(s, dy, g[1]);
        g[2] = fmaf(s, dz, g[2]);
        lap += -2.0f * f1 * inv + 2.0f * s;
      }
    }
    part[(t * n + i) * 2] = val;
    part[(t * n + i) * 2 + 1] = lap;
    if (track) {
      float* o = extra + w * C + 1 + 3 * i;
      o[0] = g[0];
      o[1] = g[1];
      o[2] = g[2];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nw; t += blockDim.x) {
    float v = 0.f, l = 0.f;
    for (int i = 0; i < n; ++i) {
      v += part[(t * n + i) * 2];
      l += part[(t * n + i) * 2 + 1];
    }
    extra[(long long)(w0 + t) * C] = v;
    if (track) extra[(long long)(w0 + t) * C + C - 1] = l;
  }
}

int jq_launch_jastrow(const float* electrons, const float* alpha_par, const float* alpha_anti, int W, JqSpins sp,
                      int track, float* extra, cudaStream_t st) {
  if (W <= 0) return JQ_OK;
  size_t smem = sizeof(float) * JT * sp.n() * 2;
  JQ_LAUNCH(k_jastrow, dim3(jq_cdiv(W, JT)), dim3(128), smem, st, electrons, alpha_par, alpha_anti, W, sp, track,
            extra);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// wf.orbitals output: (ndets, n_electrons, n_orbitals) per walker (output/orbital.py:78 transposes the electron-major
// DenseGeneral output the same way); complex orbitals are written interleaved.
// ------------------------------------------------------------------------------------------------
__global__ void k_orbitals_out(const float* __restrict__ in_re, const float* __restrict__ in_im, float* __restrict__ out,
                               long long total, int n, int D) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % n);
    long long t = i / n;
    const int e = (int)(t % n);
    t /= n;
    const int dd = (int)(t % D);
    const long long w = t / D;
    const long long src = ((w * n + e) * D + dd) * n + o;
    if (in_im) {
      out[2 * i] = in_re[src];
      out[2 * i + 1] = in_im[src];
    } else {
      out[i] = in_re[src];
    }
  }
}

int jq_launch_orbitals_out(const float* in_re, const float* in_im, float* out, long long W, int n, int D,
                           cudaStream_t st) {
  const long long total = W * D * n * n;
  if (total <= 0) return JQ_OK;
  int grid = jq_cdiv(total, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  JQ_LAUNCH(k_orbitals_out, dim3(grid), dim3(256), 0, st, in_re, in_im, out, total, n, D);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// head pipeline
// ------------------------------------------------------------------------------------------------
void jq_head_carve(const JqHeadDims& d, long long W, JqArena& ar, JqHeadBufs* b) {
  const long long n = d.sp.n();
  b->orb = ar.take<float>(W * n * d.C * d.D * n);
  b->det_sign = ar.take<float>(W * d.D);
  b->det_logabs = ar.take<float>(W * d.D);
  b->det_grad = ar.take<float>(W * d.D * (d.C > 1 ? 3 * n : 1));
  b->det_lap = ar.take<float>(W * d.D);
  b->extra = d.jastrow ? ar.take<float>(W * d.C) : nullptr;
}

int jq_head_forward(const JqHeadDims& d, const jaqmc_head_params* p, const float* h, const float* electrons,
                    const float* atoms, long long W, const JqHeadBufs& b, float* wscratch, JqWfOut out,
                    cudaStream_t st) {
  const int n = d.sp.n();
  const int track = d.C > 1;
  const bool split = d.split && d.sp.nch() == 2;
  JQ_REQUIRE(p->orbital_kernel[0] && (!split || p->orbital_kernel[1]), JQ_ERR_INVALID_ARGUMENT,
             "head: null orbital kernel");
  JQ_REQUIRE(d.envelope_type == JAQMC_ENVELOPE_NULL ||
                 (p->env_pi[0] && p->env_sigma[0] && (!split || (p->env_pi[1] && p->env_sigma[1]))),
             JQ_ERR_INVALID_ARGUMENT, "head: null envelope parameter");
  JQ_REQUIRE(!d.jastrow || (p->jastrow_alpha_par && p->jastrow_alpha_anti), JQ_ERR_INVALID_ARGUMENT,
             "head: null jastrow parameter");
  int rc;
  const int nchan = split ? 2 : 1;
  // The envelope is fused into the orbital GEMM's epilogue under the forward Laplacian (tcgen05 kernel); on the
  // value-only sampling path (one row per electron) the separate elementwise pass is cheaper than a fused epilogue
  // variant.  Eligibility depends on the channel (its electron count enters the launch size), and the separate pass
  // runs over the whole orbital buffer: fuse only when EVERY channel is eligible, decided before anything is launched.
  JqDenseArgs args[2];
  JqEnvFuse ef[2];
  bool env_fused = track && (d.envelope_type == JAQMC_ENVELOPE_ISOTROPIC || d.envelope_type == JAQMC_ENVELOPE_ABS_ISOTROPIC);
  for (int s = 0; s < nchan; ++s) {
    JqDenseArgs& a = args[s];
    memset(&a, 0, sizeof(a));
    a.src0 = h;
    a.k0 = d.hidden;
    a.k0_valid = d.hidden_valid;
    a.w0 = p->orbital_kernel[s];
    a.bias = p->orbital_bias[s];
    a.out = b.orb;
    a.wscratch = wscratch;
    a.N = d.D * n;
    a.C = d.C;
    a.n_tot = n;
    if (split) {
      a.j0 = d.sp.lo(s);
      a.n_sub = d.sp.hi(s) - d.sp.lo(s);
    } else {
      a.j0 = 0;
      a.n_sub = n;
    }
    a.G = W * a.n_sub;
    if (env_fused && !jq_dense_tc_eligible(a)) env_fused = false;
  }
  for (int s = 0; s < nchan; ++s) {
    JqDenseArgs& a = args[s];
    if (env_fused) {
      ef[s].electrons = electrons;
      ef[s].atoms = atoms;
      ef[s].pi = p->env_pi[s];
      ef[s].sigma = p->env_sigma[s];
      ef[s].A = d.A;
      ef[s].D = d.D;
      ef[s].n = n;
      ef[s].type = d.envelope_type;
      a.env = &ef[s];
      a.act = 2;
    }
    if ((rc = jq_launch_dense(a, st))) return rc;
  }
  if (jq_prep.collect) return JQ_OK;   // dry pass of the weight-split cache: only the dense launches above matter
  JqEnvelopeArgs env;
  env.type = d.envelope_type;
  env.pi[0] = p->env_pi[0];
  env.sigma[0] = p->env_sigma[0];
  env.pi[1] = split ? p->env_pi[1] : nullptr;
  env.sigma[1] = split ? p->env_sigma[1] : nullptr;
  // sampling path (value only, n <= 32): the envelope is applied inside the determinant kernel as it loads the rows
  const bool env_in_logdet = !track && !out.orbitals && jq_logdet_value_env_eligible(n, d.A, d.envelope_type);
  if (!env_fused && !env_in_logdet &&
      (rc = jq_launch_orb_envelope(b.orb, electrons, atoms, env, (int)W, d.sp, d.A, d.D, track, st)))
    return rc;
  if (out.orbitals) {   // wf.orbitals (pretraining head): the matrices themselves, no determinant
    JQ_REQUIRE(!track, JQ_ERR_INVALID_ARGUMENT, "head: orbitals are emitted on the value path only");
    return jq_launch_orbitals_out(b.orb, nullptr, out.orbitals, W, n, d.D, st);
  }
  if (env_in_logdet)
    rc = jq_launch_logdet_value_env(b.orb, electrons, atoms, env, 1, (int)W, d.sp, d.A, d.D, b.det_sign, b.det_logabs, st);
  else
    rc = jq_launch_logdet(b.orb, (int)W, n, d.D, track, b.det_sign, b.det_logabs, b.det_grad, b.det_lap, st);
  if (rc) return rc;
  if (d.jastrow &&
      (rc = jq_launch_jastrow(electrons, p->jastrow_alpha_par, p->jastrow_alpha_anti, (int)W, d.sp, track, b.extra, st)))
    return rc;
  return jq_launch_logdet_combine(b.det_sign, b.det_logabs, b.det_grad, b.det_lap, (int)W, n, d.D, track,
                                  d.jastrow ? b.extra : nullptr, out.logpsi, out.sign, out.grad, out.lap, out.e_kin, st);
}

// ------------------------------------------------------------------------------------------------
// HydrogenAtom demo wavefunction (app/hydrogen_atom.py:28-35): log psi = alpha * ||electrons|| (norm of the whole
// (n, 3) array; the app uses n = 1).  grad = alpha x / r, lap = alpha (3n - 1) / r.  One item per walker.
// ------------------------------------------------------------------------------------------------
__global__ void k_hydrogen(const float* __restrict__ el, const float* __restrict__ alpha, int W, int n, int track,
                           float* __restrict__ logpsi, float* __restrict__ sign, float* __restrict__ grad,
                           float* __restrict__ lap, float* __restrict__ e_kin) {
  const int K = 3 * n;
  const float a = alpha[0];
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W;
       w += (long long)gridDim.x * blockDim.x) {
    const float* e = el + w * K;
    float r2 = 0.f;
    for (int k = 0; k < K; ++k) r2 = fmaf(e[k], e[k], r2);
    const float r = sqrtf(r2);
    logpsi[w] = a * r;
    if (sign) sign[w] = 1.0f;
    if (!track) continue;
    const float rinv = 1.0f / r;
    for (int k = 0; k < K; ++k) grad[w * K + k] = a * e[k] * rinv;
    const float l = a * (float)(K - 1) * rinv;
    lap[w] = l;
    e_kin[w] = -0.5f * (l + a * a);
  }
}

int jq_hydrogen_forward(const jaqmc_hydrogen_config* c, const jaqmc_hydrogen_params* p, const float* electrons,
                        long long W, int track, JqWfOut out, cudaStream_t st) {
  JQ_REQUIRE(c->n_electrons >= 1, JQ_ERR_INVALID_ARGUMENT, "hydrogen: n_electrons=%d", c->n_electrons);
  JQ_REQUIRE(p->alpha, JQ_ERR_INVALID_ARGUMENT, "hydrogen: null alpha");
  if (W <= 0) return JQ_OK;
  JQ_LAUNCH(k_hydrogen, dim3(jq_cdiv(W, 128)), dim3(128), 0, st, electrons, p->alpha, (int)W, c->n_electrons, track,
            out.logpsi, out.sign, out.grad, out.lap, out.e_kin);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}
