// Periodic input features of the solid-state FermiNet with their Local1 / Local2 Jacobians and Laplacians.
//
// Reference semantics: wavefunction/input/atomic.py:108-147 (SolidFeatures, `tri` distance, minimal symmetry),
// geometry/pbc.py:97-111 (wrap_positions), :282-324 (tri_distance), :347-381 (get_symmetry_lat).
//   x        displacement: wrapped electron - primitive atom (ae) or wrapped electron i - wrapped electron j (ee)
//   w_l      = b_l . x                      (b = 2 pi inv(lattice)^T rows, a = pinv(b)^T rows)
//   rel      = [ sum_l sin(w_l) a_l , sum_l cos(w_l) a_l ]                                     (6 numbers)
//   sd       = sqrt( sum_lk (a a^T)_lk [ sin w_l sin w_k + (1 - cos w_l)(1 - cos w_k) ] )
//   features = [ sd, rel ]  (7 per atom / per pair)
// wrap_positions subtracts a constant integer number of lattice vectors, so under derivative tracking it is the
// identity; all derivatives below are with respect to x (d/dr_i = +d/dx, d/dr_j = -d/dx).
//   u = sd^2,  p = M s,  r = M (1 - c)   (M = a a^T, s = sin w, c = cos w)
//   du/dx_m = 2 sum_l b_lm (c_l p_l + s_l r_l)
//   lap u   = 2 sum_l |b_l|^2 (c_l r_l - s_l p_l) + 2 sum_lk M_lk (b_l . b_k)(c_l c_k + s_l s_k)
//   d sd    = du / (2 sd),   lap sd = lap u / (2 sd) - |du|^2 / (4 sd^3)
#include "wf.cuh"

struct JqTri {
  float a[9], b[9], M[9], bb[9], b2[3];  // a, b rows; M = a a^T; bb_lk = b_l . b_k; b2_l = |b_l|^2
  float lat[9], inv[9];                   // lattice rows and its inverse (wrap_positions)
};

__device__ __forceinline__ void jq_wrap(const JqTri& t, const float* p, float* o) {
  float f[3];
  for (int k = 0; k < 3; ++k) {
    float v = p[0] * t.inv[k] + p[1] * t.inv[3 + k] + p[2] * t.inv[6 + k];
    f[k] = v - floorf(v);
  }
  for (int m = 0; m < 3; ++m) o[m] = f[0] * t.lat[m] + f[1] * t.lat[3 + m] + f[2] * t.lat[6 + m];
}

// value[7], jac[3][7] (d/dx_m), lap[7]
__device__ __forceinline__ void jq_tri_features(const JqTri& t, const float* x, int track, float* val, float* jac,
                                                float* lap) {
  float s[3], c[3];
  for (int l = 0; l < 3; ++l) {
    float w = t.b[3 * l] * x[0] + t.b[3 * l + 1] * x[1] + t.b[3 * l + 2] * x[2];
    sincosf_(w, &s[l], &c[l]);
  }
  float p[3], r[3];
  for (int l = 0; l < 3; ++l) {
    p[l] = t.M[3 * l] * s[0] + t.M[3 * l + 1] * s[1] + t.M[3 * l + 2] * s[2];
    r[l] = t.M[3 * l] * (1.f - c[0]) + t.M[3 * l + 1] * (1.f - c[1]) + t.M[3 * l + 2] * (1.f - c[2]);
  }
  float u = 0.f;
  for (int l = 0; l < 3; ++l) u += s[l] * p[l] + (1.f - c[l]) * r[l];
  const float sd = sqrtf(u);
  val[0] = sd;
  for (int m = 0; m < 3; ++m) {
    val[1 + m] = s[0] * t.a[m] + s[1] * t.a[3 + m] + s[2] * t.a[6 + m];
    val[4 + m] = c[0] * t.a[m] + c[1] * t.a[3 + m] + c[2] * t.a[6 + m];
  }
  if (!track) return;
  float du[3], du2 = 0.f;
  for (int m = 0; m < 3; ++m) {
    float v = 0.f;
    for (int l = 0; l < 3; ++l) v += t.b[3 * l + m] * (c[l] * p[l] + s[l] * r[l]);
    du[m] = 2.f * v;
    du2 += du[m] * du[m];
  }
  float lu = 0.f;
  for (int l = 0; l < 3; ++l) {
    lu += t.b2[l] * (c[l] * r[l] - s[l] * p[l]);
    for (int k = 0; k < 3; ++k) lu += t.M[3 * l + k] * t.bb[3 * l + k] * (c[l] * c[k] + s[l] * s[k]);
  }
  lu *= 2.f;
  const float inv_sd = 1.0f / sd;
  for (int m = 0; m < 3; ++m) jac[m * 7] = 0.5f * du[m] * inv_sd;
  lap[0] = 0.5f * lu * inv_sd - 0.25f * du2 * inv_sd * inv_sd * inv_sd;
  for (int mp = 0; mp < 3; ++mp) {
    float ls = 0.f, lc = 0.f;
    for (int l = 0; l < 3; ++l) {
      ls -= s[l] * t.b2[l] * t.a[3 * l + mp];
      lc -= c[l] * t.b2[l] * t.a[3 * l + mp];
    }
    lap[1 + mp] = ls;
    lap[4 + mp] = lc;
    for (int m = 0; m < 3; ++m) {
      float js = 0.f, jc = 0.f;
      for (int l = 0; l < 3; ++l) {
        js += c[l] * t.b[3 * l + m] * t.a[3 * l + mp];
        jc -= s[l] * t.b[3 * l + m] * t.a[3 * l + mp];
      }
      jac[m * 7 + 1 + mp] = js;
      jac[m * 7 + 4 + mp] = jc;
    }
  }
}

// one item per (walker, electron, primitive atom): ae [W][n][C1][7*A] Local1, r_ae [W][n][C1][A] (sd with derivatives)
__global__ void k_solid_ae_features(const float* __restrict__ el, const float* __restrict__ prim_atoms, JqTri tri,
                                    long long items, int n, int A, int C, float* __restrict__ ae,
                                    float* __restrict__ r_ae) {
  const int F = 7 * A;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int I = (int)(it % A);
    const long long g = it / A;
    float pe[3], x[3], val[7], jac[21], lap[7];
    jq_wrap(tri, el + g * 3, pe);
    for (int m = 0; m < 3; ++m) x[m] = pe[m] - prim_atoms[3 * I + m];
    jq_tri_features(tri, x, C > 1, val, jac, lap);
    float* o = ae + g * C * F + 7 * I;
    for (int f = 0; f < 7; ++f) o[f] = val[f];
    r_ae[g * C * A + I] = val[0];
    if (C > 1) {
      for (int m = 0; m < 3; ++m) {
        for (int f = 0; f < 7; ++f) o[(1 + m) * F + f] = jac[m * 7 + f];
        r_ae[(g * C + 1 + m) * A + I] = jac[m * 7];
      }
      for (int f = 0; f < 7; ++f) o[4 * F + f] = lap[f];
      r_ae[(g * C + 4) * A + I] = lap[0];
    }
  }
}

// one item per (walker, i, j): ee [W][n*n][C2][7] Local2 (rows 1-3 d/dr_i, 4-6 d/dr_j, 7 Laplacian); zero on the diagonal
__global__ void k_solid_ee_features(const float* __restrict__ el, JqTri tri, long long items, int n, int C,
                                    float* __restrict__ ee) {
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(it % n);
    long long t = it / n;
    const int i = (int)(t % n);
    const long long w = t / n;
    float* o = ee + it * C * 7;
    if (i == j) {
      for (int q = 0; q < C * 7; ++q) o[q] = 0.f;
      continue;
    }
    float si[3], sj[3], x[3], val[7], jac[21], lap[7];
    jq_wrap(tri, el + (w * n + i) * 3, si);
    jq_wrap(tri, el + (w * n + j) * 3, sj);
    for (int m = 0; m < 3; ++m) x[m] = si[m] - sj[m];
    jq_tri_features(tri, x, C > 1, val, jac, lap);
    for (int f = 0; f < 7; ++f) o[f] = val[f];
    if (C > 1) {
      for (int m = 0; m < 3; ++m)
        for (int f = 0; f < 7; ++f) {
          o[(1 + m) * 7 + f] = jac[m * 7 + f];
          o[(4 + m) * 7 + f] = -jac[m * 7 + f];
        }
      for (int f = 0; f < 7; ++f) o[7 * 7 + f] = 2.0f * lap[f];
    }
  }
}

// a, b from the lattice as get_symmetry_lat does (minimal symmetry): b = 2 pi inv(lattice)^T, a = pinv(b)^T = lattice / 2 pi
static void make_tri(const float* lattice, JqTri* t) {
  double L[9], inv[9];
  for (int i = 0; i < 9; ++i) L[i] = lattice[i];
  const double det = L[0] * (L[4] * L[8] - L[5] * L[7]) - L[1] * (L[3] * L[8] - L[5] * L[6]) + L[2] * (L[3] * L[7] - L[4] * L[6]);
  inv[0] = (L[4] * L[8] - L[5] * L[7]) / det;
  inv[1] = (L[2] * L[7] - L[1] * L[8]) / det;
  inv[2] = (L[1] * L[5] - L[2] * L[4]) / det;
  inv[3] = (L[5] * L[6] - L[3] * L[8]) / det;
  inv[4] = (L[0] * L[8] - L[2] * L[6]) / det;
  inv[5] = (L[2] * L[3] - L[0] * L[5]) / det;
  inv[6] = (L[3] * L[7] - L[4] * L[6]) / det;
  inv[7] = (L[1] * L[6] - L[0] * L[7]) / det;
  inv[8] = (L[0] * L[4] - L[1] * L[3]) / det;
  const double two_pi = 6.283185307179586476925286766559;
  double a[9], b[9];
  for (int l = 0; l < 3; ++l)
    for (int m = 0; m < 3; ++m) {
      b[3 * l + m] = two_pi * inv[3 * m + l];  // (inv^T)[l][m]
      a[3 * l + m] = L[3 * l + m] / two_pi;    // pinv(b)^T = lattice / 2 pi for an invertible cell
    }
  for (int i = 0; i < 9; ++i) {
    t->a[i] = (float)a[i];
    t->b[i] = (float)b[i];
    t->lat[i] = (float)L[i];
    t->inv[i] = (float)inv[i];
  }
  for (int l = 0; l < 3; ++l) {
    for (int k = 0; k < 3; ++k) {
      double m = 0, bb = 0;
      for (int q = 0; q < 3; ++q) {
        m += a[3 * l + q] * a[3 * k + q];
        bb += b[3 * l + q] * b[3 * k + q];
      }
      t->M[3 * l + k] = (float)m;
      t->bb[3 * l + k] = (float)bb;
    }
    t->b2[l] = t->bb[3 * l + l];
  }
}

// `sim_lattice` / `prim_lattice` are HOST pointers to 9 floats (rows are lattice vectors)
int jq_launch_solid_features(const float* electrons, const float* prim_atoms, const float* sim_lattice,
                             const float* prim_lattice, int W, int n, int A, int track, float* ae, float* r_ae, float* ee,
                             cudaStream_t st) {
  JqTri tp, ts;
  make_tri(prim_lattice, &tp);
  make_tri(sim_lattice, &ts);
  long long items = (long long)W * n * A;
  if (items > 0) {
    int grid = jq_cdiv(items, 128);
    if (grid > 148 * 16) grid = 148 * 16;
    JQ_LAUNCH(k_solid_ae_features, dim3(grid), dim3(128), 0, st, electrons, prim_atoms, tp, items, n, A, track ? 5 : 1, ae,
              r_ae);
    JQ_CHECK_LAUNCH();
  }
  long long pitems = (long long)W * n * n;
  if (ee != nullptr && pitems > 0) {
    int grid = jq_cdiv(pitems, 128);
    if (grid > 148 * 32) grid = 148 * 32;
    JQ_LAUNCH(k_solid_ee_features, dim3(grid), dim3(128), 0, st, electrons, ts, pitems, n, track ? 8 : 1, ee);
    JQ_CHECK_LAUNCH();
  }
  return JQ_OK;
}
