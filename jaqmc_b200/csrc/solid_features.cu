// Periodic input features of the solid-state FermiNet with their Local1 / Local2 Jacobians and Laplacians.
//
// Reference semantics: wavefunction/input/atomic.py:108-147 (SolidFeatures), geometry/pbc.py:97-111 (wrap_positions),
// :282-324 (tri_distance), :205-279 (nu_distance, 4 features per pair), :347-381 (get_symmetry_lat: 3, 4 or 6 directions).
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

constexpr int JQ_LMAX = 6;   // lattice directions after get_symmetry_lat: 3 (minimal), 4 (fcc, hexagonal), 6 (bcc)
struct JqTri {
  int L;                       // number of directions
  int nu;                      // 0: `tri` distance (7 features per pair), 1: `nu` polynomial distance (4 features)
  float a[3 * JQ_LMAX], b[3 * JQ_LMAX];             // a, b rows
  float M[JQ_LMAX * JQ_LMAX], bb[JQ_LMAX * JQ_LMAX];  // M = a a^T; bb_lk = b_l . b_k
  float b2[JQ_LMAX];           // |b_l|^2
  float lat[9], inv[9];        // lattice rows and its inverse (wrap_positions)
  __host__ __device__ int fw() const { return nu ? 4 : 7; }
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
  const int L = t.L;
  float s[JQ_LMAX], c[JQ_LMAX];
  for (int l = 0; l < L; ++l) {
    float w = t.b[3 * l] * x[0] + t.b[3 * l + 1] * x[1] + t.b[3 * l + 2] * x[2];
    sincosf_(w, &s[l], &c[l]);
  }
  float p[JQ_LMAX], r[JQ_LMAX];
  for (int l = 0; l < L; ++l) {
    float pa = 0.f, ra = 0.f;
    for (int k = 0; k < L; ++k) {
      pa += t.M[JQ_LMAX * l + k] * s[k];
      ra += t.M[JQ_LMAX * l + k] * (1.f - c[k]);
    }
    p[l] = pa;
    r[l] = ra;
  }
  float u = 0.f;
  for (int l = 0; l < L; ++l) u += s[l] * p[l] + (1.f - c[l]) * r[l];
  const float sd = sqrtf(u);
  val[0] = sd;
  for (int m = 0; m < 3; ++m) {
    float vs = 0.f, vc = 0.f;
    for (int l = 0; l < L; ++l) {
      vs += s[l] * t.a[3 * l + m];
      vc += c[l] * t.a[3 * l + m];
    }
    val[1 + m] = vs;
    val[4 + m] = vc;
  }
  if (!track) return;
  float du[3], du2 = 0.f;
  for (int m = 0; m < 3; ++m) {
    float v = 0.f;
    for (int l = 0; l < L; ++l) v += t.b[3 * l + m] * (c[l] * p[l] + s[l] * r[l]);
    du[m] = 2.f * v;
    du2 += du[m] * du[m];
  }
  float lu = 0.f;
  for (int l = 0; l < L; ++l) {
    lu += t.b2[l] * (c[l] * r[l] - s[l] * p[l]);
    for (int k = 0; k < L; ++k) lu += t.M[JQ_LMAX * l + k] * t.bb[JQ_LMAX * l + k] * (c[l] * c[k] + s[l] * s[k]);
  }
  lu *= 2.f;
  const float inv_sd = 1.0f / sd;
  for (int m = 0; m < 3; ++m) jac[m * 7] = 0.5f * du[m] * inv_sd;
  lap[0] = 0.5f * lu * inv_sd - 0.25f * du2 * inv_sd * inv_sd * inv_sd;
  for (int mp = 0; mp < 3; ++mp) {
    float ls = 0.f, lc = 0.f;
    for (int l = 0; l < L; ++l) {
      ls -= s[l] * t.b2[l] * t.a[3 * l + mp];
      lc -= c[l] * t.b2[l] * t.a[3 * l + mp];
    }
    lap[1 + mp] = ls;
    lap[4 + mp] = lc;
    for (int m = 0; m < 3; ++m) {
      float js = 0.f, jc = 0.f;
      for (int l = 0; l < L; ++l) {
        js += c[l] * t.b[3 * l + m] * t.a[3 * l + mp];
        jc -= s[l] * t.b[3 * l + m] * t.a[3 * l + mp];
      }
      jac[m * 7 + 1 + mp] = js;
      jac[m * 7 + 4 + mp] = jc;
    }
  }
}

// `nu` polynomial distance (geometry/pbc.py:205-279): w_l = b_l . x reduced to [-pi, pi) (a constant shift under
// derivative tracking),
//   f(w) = |w| (1 - |w / pi|^3 / 4),  g(w) = w (1 - 3/2 |w / pi| + 1/2 (w / pi)^2),
//   u = sd^2 = sum_l |a_l|^2 f(w_l)^2 + sum_{l != k} M_lk g_l g_k,   rel = sum_l g_l a_l
//   F = f^2 = w^2 q^2 with q = 1 - |w|^3 / (4 pi^3):  F' = 2 w q^2 + 2 w^2 q q',  q' = -3 w |w| / (4 pi^3),
//        F'' = 2 q^2 + 8 w q q' + 2 w^2 (q'^2 + q q''),  q'' = -3 |w| / (2 pi^3)
//   g' = 1 - 3 |w| / pi + 3 w^2 / (2 pi^2),   g'' = -3 sign(w) / pi + 3 w / pi^2
//   du/dw_l = M_ll F'_l + 2 g'_l sum_{k != l} M_lk g_k
//   d2u/dw_l dw_k = 2 M_lk g'_l g'_k (l != k),   d2u/dw_l^2 = M_ll F''_l + 2 g''_l sum_{k != l} M_lk g_k
//   du/dx_m = sum_l b_lm du/dw_l,   lap u = sum_lk (b_l . b_k) d2u/dw_l dw_k
// value[4], jac[3][4] (row stride 7, like the tri features), lap[4]
__device__ __forceinline__ void jq_nu_features(const JqTri& t, const float* x, int track, float* val, float* jac,
                                               float* lap) {
  const int L = t.L;
  const float pi = 3.14159265358979323846f, ipi = 1.0f / pi;
  float g[JQ_LMAX], g1[JQ_LMAX], g2[JQ_LMAX], F[JQ_LMAX], F1[JQ_LMAX], F2[JQ_LMAX];
  for (int l = 0; l < L; ++l) {
    float w = t.b[3 * l] * x[0] + t.b[3 * l + 1] * x[1] + t.b[3 * l + 2] * x[2];
    w -= floorf((w + pi) * (0.5f * ipi)) * (2.0f * pi);
    const float aw = fabsf(w), sw = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : 0.f);
    const float z = aw * ipi;
    g[l] = w * (1.f - 1.5f * z + 0.5f * z * z);
    g1[l] = 1.f - 3.f * z + 1.5f * z * z;
    g2[l] = (-3.f * sw + 3.f * w * ipi) * ipi;
    const float k3 = 0.25f * ipi * ipi * ipi;   // 1 / (4 pi^3)
    const float q = 1.f - aw * aw * aw * k3;
    const float q1 = -3.f * w * aw * k3, q2 = -6.f * aw * k3;
    F[l] = w * w * q * q;
    F1[l] = 2.f * w * q * q + 2.f * w * w * q * q1;
    F2[l] = 2.f * q * q + 8.f * w * q * q1 + 2.f * w * w * (q1 * q1 + q * q2);
  }
  float u = 0.f, off[JQ_LMAX];   // off_l = sum_{k != l} M_lk g_k
  for (int l = 0; l < L; ++l) {
    float o = 0.f;
    for (int k = 0; k < L; ++k)
      if (k != l) o += t.M[JQ_LMAX * l + k] * g[k];
    off[l] = o;
    u += t.M[JQ_LMAX * l + l] * F[l] + g[l] * o;
  }
  const float sd = sqrtf(u);
  val[0] = sd;
  for (int m = 0; m < 3; ++m) {
    float v = 0.f;
    for (int l = 0; l < L; ++l) v += g[l] * t.a[3 * l + m];
    val[1 + m] = v;
  }
  if (!track) return;
  float uw[JQ_LMAX];
  for (int l = 0; l < L; ++l) uw[l] = t.M[JQ_LMAX * l + l] * F1[l] + 2.f * g1[l] * off[l];
  float du[3], du2 = 0.f;
  for (int m = 0; m < 3; ++m) {
    float v = 0.f;
    for (int l = 0; l < L; ++l) v += t.b[3 * l + m] * uw[l];
    du[m] = v;
    du2 += v * v;
  }
  float lu = 0.f;
  for (int l = 0; l < L; ++l) {
    lu += t.b2[l] * (t.M[JQ_LMAX * l + l] * F2[l] + 2.f * g2[l] * off[l]);
    for (int k = 0; k < L; ++k)
      if (k != l) lu += t.bb[JQ_LMAX * l + k] * 2.f * t.M[JQ_LMAX * l + k] * g1[l] * g1[k];
  }
  const float inv_sd = 1.0f / sd;
  for (int m = 0; m < 3; ++m) jac[m * 7] = 0.5f * du[m] * inv_sd;
  lap[0] = 0.5f * lu * inv_sd - 0.25f * du2 * inv_sd * inv_sd * inv_sd;
  for (int mp = 0; mp < 3; ++mp) {
    float lr = 0.f;
    for (int l = 0; l < L; ++l) lr += g2[l] * t.b2[l] * t.a[3 * l + mp];
    lap[1 + mp] = lr;
    for (int m = 0; m < 3; ++m) {
      float jr = 0.f;
      for (int l = 0; l < L; ++l) jr += g1[l] * t.b[3 * l + m] * t.a[3 * l + mp];
      jac[m * 7 + 1 + mp] = jr;
    }
  }
}

__device__ __forceinline__ void jq_pbc_features(const JqTri& t, const float* x, int track, float* val, float* jac,
                                                float* lap) {
  if (t.nu) jq_nu_features(t, x, track, val, jac, lap);
  else jq_tri_features(t, x, track, val, jac, lap);
}

// one item per (walker, electron, primitive atom): ae [W][n][C1][7*A] Local1, r_ae [W][n][C1][A] (sd with derivatives)
__global__ void k_solid_ae_features(const float* __restrict__ el, const float* __restrict__ prim_atoms, JqTri tri,
                                    long long items, int n, int A, int C, float* __restrict__ ae,
                                    float* __restrict__ r_ae) {
  const int FW = tri.fw(), F = FW * A;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int I = (int)(it % A);
    const long long g = it / A;
    float pe[3], x[3], val[7], jac[21], lap[7];
    jq_wrap(tri, el + g * 3, pe);
    for (int m = 0; m < 3; ++m) x[m] = pe[m] - prim_atoms[3 * I + m];
    jq_pbc_features(tri, x, C > 1, val, jac, lap);
    float* o = ae + g * C * F + FW * I;
    for (int f = 0; f < FW; ++f) o[f] = val[f];
    r_ae[g * C * A + I] = val[0];
    if (C > 1) {
      for (int m = 0; m < 3; ++m) {
        for (int f = 0; f < FW; ++f) o[(1 + m) * F + f] = jac[m * 7 + f];
        r_ae[(g * C + 1 + m) * A + I] = jac[m * 7];
      }
      for (int f = 0; f < FW; ++f) o[4 * F + f] = lap[f];
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
    const int FW = tri.fw();
    float* o = ee + it * C * FW;
    if (i == j) {
      for (int q = 0; q < C * FW; ++q) o[q] = 0.f;
      continue;
    }
    float si[3], sj[3], x[3], val[7], jac[21], lap[7];
    jq_wrap(tri, el + (w * n + i) * 3, si);
    jq_wrap(tri, el + (w * n + j) * 3, sj);
    for (int m = 0; m < 3; ++m) x[m] = si[m] - sj[m];
    jq_pbc_features(tri, x, C > 1, val, jac, lap);
    for (int f = 0; f < FW; ++f) o[f] = val[f];
    if (C > 1) {
      for (int m = 0; m < 3; ++m)
        for (int f = 0; f < FW; ++f) {
          o[(1 + m) * FW + f] = jac[m * 7 + f];
          o[(4 + m) * FW + f] = -jac[m * 7 + f];
        }
      for (int f = 0; f < FW; ++f) o[7 * FW + f] = 2.0f * lap[f];
    }
  }
}

// a, b from the lattice as get_symmetry_lat does (geometry/pbc.py:347-381): b = mat (2 pi inv(lattice)^T) with the integer
// direction table of the symmetry type, a = pinv(b)^T = b (b^T b)^-1 (for three directions: lattice / 2 pi).
static int make_tri(const float* lattice, int distance_type, int sym_type, JqTri* t) {
  static const int MAT_MIN[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  static const int MAT_FCC[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 1}};
  static const int MAT_BCC[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, -1, 0}, {1, 0, -1}, {0, 1, -1}};
  static const int MAT_HEX[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, -1, 0}};
  const int (*mat)[3] = MAT_MIN;
  int Ld = 3;
  if (sym_type == JAQMC_SYMMETRY_FCC) { mat = MAT_FCC; Ld = 4; }
  else if (sym_type == JAQMC_SYMMETRY_BCC) { mat = MAT_BCC; Ld = 6; }
  else if (sym_type == JAQMC_SYMMETRY_HEXAGONAL) { mat = MAT_HEX; Ld = 4; }
  else JQ_REQUIRE(sym_type == JAQMC_SYMMETRY_MINIMAL, JQ_ERR_INVALID_ARGUMENT, "solid: unknown sym_type %d", sym_type);
  JQ_REQUIRE(distance_type == JAQMC_DISTANCE_TRI || distance_type == JAQMC_DISTANCE_NU, JQ_ERR_INVALID_ARGUMENT,
             "solid: unknown distance_type %d", distance_type);
  memset(t, 0, sizeof(*t));
  t->L = Ld;
  t->nu = distance_type == JAQMC_DISTANCE_NU;
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
  double b0[9], a[3 * JQ_LMAX], b[3 * JQ_LMAX];
  for (int l = 0; l < 3; ++l)
    for (int m = 0; m < 3; ++m) b0[3 * l + m] = two_pi * inv[3 * m + l];  // (inv^T)[l][m]
  for (int l = 0; l < Ld; ++l)
    for (int m = 0; m < 3; ++m) b[3 * l + m] = mat[l][0] * b0[m] + mat[l][1] * b0[3 + m] + mat[l][2] * b0[6 + m];
  // S = b^T b (3 x 3, symmetric positive definite), a = b S^-1
  double S[9] = {0}, Si[9];
  for (int m = 0; m < 3; ++m)
    for (int q = 0; q < 3; ++q)
      for (int l = 0; l < Ld; ++l) S[3 * m + q] += b[3 * l + m] * b[3 * l + q];
  const double ds = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  Si[0] = (S[4] * S[8] - S[5] * S[7]) / ds;
  Si[1] = (S[2] * S[7] - S[1] * S[8]) / ds;
  Si[2] = (S[1] * S[5] - S[2] * S[4]) / ds;
  Si[3] = (S[5] * S[6] - S[3] * S[8]) / ds;
  Si[4] = (S[0] * S[8] - S[2] * S[6]) / ds;
  Si[5] = (S[2] * S[3] - S[0] * S[5]) / ds;
  Si[6] = (S[3] * S[7] - S[4] * S[6]) / ds;
  Si[7] = (S[1] * S[6] - S[0] * S[7]) / ds;
  Si[8] = (S[0] * S[4] - S[1] * S[3]) / ds;
  for (int l = 0; l < Ld; ++l)
    for (int m = 0; m < 3; ++m) a[3 * l + m] = b[3 * l] * Si[m] + b[3 * l + 1] * Si[3 + m] + b[3 * l + 2] * Si[6 + m];
  for (int i = 0; i < 3 * Ld; ++i) {
    t->a[i] = (float)a[i];
    t->b[i] = (float)b[i];
  }
  for (int i = 0; i < 9; ++i) {
    t->lat[i] = (float)L[i];
    t->inv[i] = (float)inv[i];
  }
  for (int l = 0; l < Ld; ++l) {
    for (int k = 0; k < Ld; ++k) {
      double m = 0, bb = 0;
      for (int q = 0; q < 3; ++q) {
        m += a[3 * l + q] * a[3 * k + q];
        bb += b[3 * l + q] * b[3 * k + q];
      }
      t->M[JQ_LMAX * l + k] = (float)m;
      t->bb[JQ_LMAX * l + k] = (float)bb;
    }
    t->b2[l] = t->bb[JQ_LMAX * l + l];
  }
  return JQ_OK;
}

// `sim_lattice` / `prim_lattice` are HOST pointers to 9 floats (rows are lattice vectors)
int jq_launch_solid_features(const float* electrons, const float* prim_atoms, const float* sim_lattice,
                             const float* prim_lattice, int distance_type, int sym_type, int W, int n, int A, int track,
                             float* ae, float* r_ae, float* ee, cudaStream_t st) {
  JqTri tp, ts;
  int rc;
  if ((rc = make_tri(prim_lattice, distance_type, sym_type, &tp))) return rc;
  if ((rc = make_tri(sim_lattice, distance_type, sym_type, &ts))) return rc;
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
