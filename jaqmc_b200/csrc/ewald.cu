// Ewald-summed electrostatic energy of point charges in a periodic cell (solid-state potential).
//
// Reference semantics: estimator/ewald.py:112-173 (EwaldSum.energy): minimum-image displacements
// (geometry/pbc.py:114-184, the three branches: diagonal, orthogonal, general 27-image search with argmin's
// first-minimum tie rule), real-space sum over the lattice images with the centre-image self term masked and
// r_safe = max(r, 1e-7) (:156), reciprocal sum over the selected half-space G vectors, self and charged-background
// constants.  The electrons (charge -1) and the ions are one set of point charges (app/solid/hamiltonian.py:28-56).
//
// One block per walker.  Phase 1: one item per ordered pair (i, j), loop over the images.  Phase 2: one item per
// G vector, loop over the particles.  Partial sums go through shared memory (one slot per thread).
#include "wf.cuh"

#define EW_THREADS 256
__global__ void k_ewald(jaqmc_ewald ew, const float* __restrict__ electrons, const float* __restrict__ atoms,
                        const float* __restrict__ charges, int n_el, int n_at, float* __restrict__ e_pot) {
  JQ_DYN_SMEM(float, sm);
  const int P = n_el + n_at;
  float* pos = sm;              // [P][3]
  float* q = pos + 3 * P;       // [P]
  double* part = reinterpret_cast<double*>(q + P);  // [blockDim.x]; float offset 4P keeps it 16-byte aligned
  const long long w = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < 3 * P; t += nt)
    pos[t] = (t < 3 * n_el) ? electrons[w * 3 * n_el + t] : atoms[t - 3 * n_el];
  for (int t = tid; t < P; t += nt) q[t] = (t < n_el) ? -1.0f : charges[t - n_el];
  __syncthreads();

  double acc = 0.0;  // the sums run over ~10^4-10^5 terms per walker: accumulate them in double
  const float* L = ew.lattice;  // rows are lattice vectors
  for (int pr = tid; pr < P * P; pr += nt) {
    const int i = pr / P, j = pr % P;
    float d0 = pos[3 * i] - pos[3 * j], d1 = pos[3 * i + 1] - pos[3 * j + 1], d2 = pos[3 * i + 2] - pos[3 * j + 2];
    float m0, m1, m2;
    if (ew.mic_kind == 0) {  // diagonal cell: (d + L/2) mod L - L/2, floor-mod like jnp's %
      const float l0 = L[0], l1 = L[4], l2 = L[8];
      float t0 = d0 + 0.5f * l0, t1 = d1 + 0.5f * l1, t2 = d2 + 0.5f * l2;
      m0 = t0 - floorf(t0 / l0) * l0 - 0.5f * l0;
      m1 = t1 - floorf(t1 / l1) * l1 - 0.5f * l1;
      m2 = t2 - floorf(t2 / l2) * l2 - 0.5f * l2;
    } else if (ew.mic_kind == 1) {  // orthogonal cell: wrap the fractional coordinates
      const float* R = ew.inv_lattice;
      float f0 = d0 * R[0] + d1 * R[3] + d2 * R[6] + 0.5f;
      float f1 = d0 * R[1] + d1 * R[4] + d2 * R[7] + 0.5f;
      float f2 = d0 * R[2] + d1 * R[5] + d2 * R[8] + 0.5f;
      f0 = f0 - floorf(f0) - 0.5f;
      f1 = f1 - floorf(f1) - 0.5f;
      f2 = f2 - floorf(f2) - 0.5f;
      m0 = f0 * L[0] + f1 * L[3] + f2 * L[6];
      m1 = f0 * L[1] + f1 * L[4] + f2 * L[7];
      m2 = f0 * L[2] + f1 * L[5] + f2 * L[8];
    } else {  // general cell: nearest of the 27 neighbouring images, first minimum wins
      float best = 3.0e38f;
      m0 = d0;
      m1 = d1;
      m2 = d2;
      for (int s = 0; s < 27; ++s) {
        float c0 = d0 + ew.mic_shifts[3 * s], c1 = d1 + ew.mic_shifts[3 * s + 1], c2 = d2 + ew.mic_shifts[3 * s + 2];
        float r2 = c0 * c0 + c1 * c1 + c2 * c2;
        if (r2 < best) {
          best = r2;
          m0 = c0;
          m1 = c1;
          m2 = c2;
        }
      }
    }
    const float qq = q[i] * q[j];
    float s = 0.f;
    for (int im = 0; im < ew.n_images; ++im) {
      if (i == j && im == ew.center_image) continue;  // masked self term
      float r0 = m0 + ew.images[3 * im], r1 = m1 + ew.images[3 * im + 1], r2 = m2 + ew.images[3 * im + 2];
      float r = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
      r = (r < 1e-7f) ? 1e-7f : r;
      s += erfcf(ew.alpha * r) / r;
    }
    acc += (double)(0.5f * qq * s);
  }
  for (int g = tid; g < ew.n_g; g += nt) {
    const float g0 = ew.gpoints[3 * g], g1 = ew.gpoints[3 * g + 1], g2 = ew.gpoints[3 * g + 2];
    float sr = 0.f, si = 0.f;
    for (int i = 0; i < P; ++i) {
      float ph = g0 * pos[3 * i] + g1 * pos[3 * i + 1] + g2 * pos[3 * i + 2];
      float sn, cs;
      sincosf_(ph, &sn, &cs);
      sr = fmaf(q[i], cs, sr);
      si = fmaf(q[i], sn, si);
    }
    acc += (double)(ew.gweight[g] * (sr * sr + si * si));
  }
  part[tid] = acc;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    float q2 = 0.f, qs = 0.f;
    for (int t = 0; t < nt; ++t) tot += part[t];
    for (int i = 0; i < P; ++i) {
      q2 = fmaf(q[i], q[i], q2);
      qs += q[i];
    }
    e_pot[w] = (float)(tot + (double)(ew.self_const_factor * q2) + (double)(0.5f * ew.ijconst * qs * qs));
  }
}

int jq_launch_ewald(const jaqmc_ewald* ew, const float* electrons, long long W, int n_el, const float* atoms,
                    const float* charges, int n_at, float* e_pot, cudaStream_t st) {
  if (W <= 0) return JQ_OK;
  JQ_REQUIRE(ew && ew->lattice && ew->images && ew->gpoints && ew->gweight, JQ_ERR_INVALID_ARGUMENT,
             "ewald: null descriptor field");
  JQ_REQUIRE(ew->mic_kind >= 0 && ew->mic_kind <= 2 && (ew->mic_kind != 1 || ew->inv_lattice) &&
                 (ew->mic_kind != 2 || ew->mic_shifts),
             JQ_ERR_INVALID_ARGUMENT, "ewald: minimum-image data missing for mic_kind %d", ew->mic_kind);
  JQ_REQUIRE(ew->n_images >= 1 && ew->center_image >= 0 && ew->center_image < ew->n_images && ew->n_g >= 0,
             JQ_ERR_INVALID_ARGUMENT, "ewald: bad image / G-vector counts");
  const int P = n_el + n_at;
  size_t smem = sizeof(float) * ((size_t)4 * P + 2) + sizeof(double) * EW_THREADS;
  JQ_REQUIRE(smem <= 48 * 1024, JQ_ERR_UNSUPPORTED, "ewald: %d particles need %zu bytes of shared memory", P, smem);
  jq_prof_work((double)W * (40.0 * P * P * ew->n_images + 30.0 * (double)ew->n_g * P), 4.0 * (double)W * (3 * n_el + 1));
  JQ_LAUNCH(k_ewald, dim3((unsigned)W), dim3(EW_THREADS), smem, st, *ew, electrons, atoms, charges, n_el, n_at, e_pot);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}
