// Electron-atom / electron-electron input features and the molecular Coulomb potential.
//
// Reference semantics: wavefunction/input/atomic.py:46-79 (MoleculeFeatures), geometry/obc.py:7-88
// (pair displacements with the eye-guarded diagonal), app/molecule/hamiltonian.py:9-22 (potential).
// Derivatives are emitted in closed form (what the reference's interpreter obtains by composing the
// square / sum / sqrt / log1p / div rules): ae features are Local1 (C = 5), ee features Local2 (C = 8).
#include "aug.cuh"

// one item per (walker, electron j, atom I)
// spin_up >= 0 appends the spin-encoding column (+1 for the first spin_up electrons, -1 after; untracked constant):
// backbone/psiformer.py:166-170, backbone/lapnet/_backbone.py:259-261.
__global__ void k_mol_ae_features(const float* __restrict__ el, const float* __restrict__ atoms, long long items,
                                  int n, int A, int rescale, int C, int spin_up, float* __restrict__ ae) {
  const int F = 4 * A + (spin_up >= 0 ? 1 : 0);
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    int I = (int)(it % A);
    long long g = it / A;  // (w, j)
    const float* e = el + g * 3;
    float d[3] = {e[0] - atoms[I * 3 + 0], e[1] - atoms[I * 3 + 1], e[2] - atoms[I * 3 + 2]};
    float r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    float r = sqrtf(r2);
    float rinv = 1.0f / r;
    float* o = ae + g * C * F + 4 * I;  // component c lives at o + c*F
    if (spin_up >= 0 && I == 0) {
      float* sc = ae + g * C * F + 4 * A;
      sc[0] = ((int)(g % n) < spin_up) ? 1.f : -1.f;
      for (int c = 1; c < C; ++c) sc[c * F] = 0.f;
    }
    if (!rescale) {
      // features (r, dx, dy, dz)
      o[0] = r;
      o[1] = d[0];
      o[2] = d[1];
      o[3] = d[2];
      if (C > 1) {
        for (int a = 0; a < 3; ++a) {
          float* oj = o + (1 + a) * F;
          oj[0] = d[a] * rinv;
          oj[1] = (a == 0) ? 1.f : 0.f;
          oj[2] = (a == 1) ? 1.f : 0.f;
          oj[3] = (a == 2) ? 1.f : 0.f;
        }
        float* ol = o + 4 * F;
        ol[0] = 2.0f * rinv;
        ol[1] = ol[2] = ol[3] = 0.f;
      }
    } else {
      // features (log(1+r), d * log(1+r)/r)   (atomic.py:59-67)
      float lg = log1pf(r);
      float s = lg * rinv;            // s(r) = log1p(r)/r
      float q = 1.0f / (1.0f + r);    // d/dr log1p
      o[0] = lg;
      o[1] = d[0] * s;
      o[2] = d[1] * s;
      o[3] = d[2] * s;
      if (C > 1) {
        float sp = (q - s) * rinv;                  // s'(r)
        float spp = (-q * q - 2.0f * sp) * rinv;    // s''(r)
        for (int a = 0; a < 3; ++a) {
          float* oj = o + (1 + a) * F;
          float ua = d[a] * rinv;
          oj[0] = q * ua;
          for (int b = 0; b < 3; ++b) oj[1 + b] = ((a == b) ? s : 0.f) + d[b] * sp * ua;
        }
        float* ol = o + 4 * F;
        ol[0] = -q * q + 2.0f * q * rinv;
        float lf = spp + 4.0f * sp * rinv;
        ol[1] = d[0] * lf;
        ol[2] = d[1] * lf;
        ol[3] = d[2] * lf;
      }
    }
  }
}

// one item per (walker, i, j): disp = r_i - r_j, r = |disp + eye| * (1 - eye)   (obc.py:39-48)
__global__ void k_mol_ee_features(const float* __restrict__ el, long long items, int n, int C,
                                  float* __restrict__ ee) {
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    int j = (int)(it % n);
    long long t = it / n;
    int i = (int)(t % n);
    long long w = t / n;
    const float* ei = el + (w * n + i) * 3;
    const float* ej = el + (w * n + j) * 3;
    float* o = ee + it * C * 4;
    if (i == j) {
      for (int q = 0; q < C * 4; ++q) o[q] = 0.f;
      continue;
    }
    float d[3] = {ei[0] - ej[0], ei[1] - ej[1], ei[2] - ej[2]};
    float r = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    float rinv = 1.0f / r;
    o[0] = r;
    o[1] = d[0];
    o[2] = d[1];
    o[3] = d[2];
    if (C > 1) {
      for (int a = 0; a < 3; ++a) {
        float* oi = o + (1 + a) * 4;  // d/d r_i[a]
        float* oj = o + (4 + a) * 4;  // d/d r_j[a]
        float u = d[a] * rinv;
        oi[0] = u;
        oj[0] = -u;
        for (int b = 0; b < 3; ++b) {
          oi[1 + b] = (a == b) ? 1.f : 0.f;
          oj[1 + b] = (a == b) ? -1.f : 0.f;
        }
      }
      float* ol = o + 7 * 4;
      ol[0] = 4.0f * rinv;  // 2/r from each of the two electrons
      ol[1] = ol[2] = ol[3] = 0.f;
    }
  }
}

int jq_launch_mol_features(const float* electrons, const float* atoms, int W, JqSpins sp, int A, int rescale,
                           int track, int spin_column, float* ae, float* ee, cudaStream_t st) {
  int n = sp.n();
  long long items = (long long)W * n * A;
  if (items > 0) {
    int grid = jq_cdiv(items, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    JQ_LAUNCH(k_mol_ae_features, dim3(grid), dim3(256), 0, st, electrons, atoms, items, n, A, rescale,
              track ? 5 : 1, spin_column ? sp.n_up : -1, ae);
    JQ_CHECK_LAUNCH();
  }
  if (ee != nullptr) {
    long long pitems = (long long)W * n * n;
    if (pitems > 0) {
      int grid = jq_cdiv(pitems, 256);
      if (grid > 148 * 16) grid = 148 * 16;
      JQ_LAUNCH(k_mol_ee_features, dim3(grid), dim3(256), 0, st, electrons, pitems, n, track ? 8 : 1, ee);
      JQ_CHECK_LAUNCH();
    }
  }
  return JQ_OK;
}

// Coulomb potential, one walker per item (hamiltonian.py:9-22).  12n bytes in, 4 bytes out per walker.
__global__ void k_coulomb(const float* __restrict__ el, const float* __restrict__ atoms,
                          const float* __restrict__ charges, int W, int n, int A, float* __restrict__ e_pot) {
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W;
       w += (long long)gridDim.x * blockDim.x) {
    const float* e = el + w * n * 3;
    float v = 0.f;
    for (int i = 0; i < n; ++i) {
      float xi = e[i * 3], yi = e[i * 3 + 1], zi = e[i * 3 + 2];
      for (int I = 0; I < A; ++I) {
        float dx = xi - atoms[I * 3], dy = yi - atoms[I * 3 + 1], dz = zi - atoms[I * 3 + 2];
        v -= charges[I] / sqrtf(dx * dx + dy * dy + dz * dz);
      }
      for (int j = i + 1; j < n; ++j) {
        float dx = xi - e[j * 3], dy = yi - e[j * 3 + 1], dz = zi - e[j * 3 + 2];
        v += 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
      }
    }
    for (int I = 0; I < A; ++I)
      for (int J = I + 1; J < A; ++J) {
        float dx = atoms[I * 3] - atoms[J * 3], dy = atoms[I * 3 + 1] - atoms[J * 3 + 1],
              dz = atoms[I * 3 + 2] - atoms[J * 3 + 2];
        v += charges[I] * charges[J] / sqrtf(dx * dx + dy * dy + dz * dz);
      }
    e_pot[w] = v;
  }
}

#ifndef JAQMC_HOST_EMU
// One warp per walker (r2): the thread-per-walker kernel above walks n (n - 1) / 2 + n A reciprocal square roots serially,
// 25 us of pure latency at any batch size -- visible in the 512-walker shard of the 8-GPU run.  Lanes take the
// electron-electron pairs (row-major upper triangle), the electron-nucleus pairs and the nucleus-nucleus pairs in
// strides of 32; the lane sums are combined by a fixed-order shuffle tree.
__global__ void __launch_bounds__(256) k_coulomb_warp(const float* __restrict__ el, const float* __restrict__ atoms,
                                                      const float* __restrict__ charges, int W, int n, int A,
                                                      float* __restrict__ e_pot) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= W) return;
  const float* e = el + w * n * 3;
  float v = 0.f;
  const int npair = n * (n - 1) / 2;
  for (int q = lane; q < npair; q += 32) {
    // q -> (i, j), i < j: row i starts at i (2n - i - 1) / 2
    int i = (int)floorf(((float)(2 * n - 1) - sqrtf((float)((2 * n - 1) * (2 * n - 1) - 8 * q))) * 0.5f);
    while (i * (2 * n - i - 1) / 2 > q) --i;
    while ((i + 1) * (2 * n - i - 2) / 2 <= q) ++i;
    const int j = i + 1 + (q - i * (2 * n - i - 1) / 2);
    const float dx = e[i * 3] - e[j * 3], dy = e[i * 3 + 1] - e[j * 3 + 1], dz = e[i * 3 + 2] - e[j * 3 + 2];
    v += 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
  }
  for (int q = lane; q < n * A; q += 32) {
    const int i = q / A, I = q - i * A;
    const float dx = e[i * 3] - atoms[I * 3], dy = e[i * 3 + 1] - atoms[I * 3 + 1], dz = e[i * 3 + 2] - atoms[I * 3 + 2];
    v -= charges[I] / sqrtf(dx * dx + dy * dy + dz * dz);
  }
  for (int q = lane; q < A * A; q += 32) {
    const int I = q / A, J = q - I * A;
    if (J > I) {
      const float dx = atoms[I * 3] - atoms[J * 3], dy = atoms[I * 3 + 1] - atoms[J * 3 + 1],
                  dz = atoms[I * 3 + 2] - atoms[J * 3 + 2];
      v += charges[I] * charges[J] / sqrtf(dx * dx + dy * dy + dz * dz);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) e_pot[w] = v;
}
#endif

int jq_launch_coulomb(const float* electrons, const float* atoms, const float* charges, int W, int n, int A,
                      float* e_pot, cudaStream_t st) {
  if (W <= 0) return JQ_OK;
#ifndef JAQMC_HOST_EMU
  JQ_LAUNCH(k_coulomb_warp, dim3(jq_cdiv(W, 8)), dim3(256), 0, st, electrons, atoms, charges, W, n, A, e_pot);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
#endif
  JQ_LAUNCH(k_coulomb, dim3(jq_cdiv(W, 128)), dim3(128), 0, st, electrons, atoms, charges, W, n, A, e_pot);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}
