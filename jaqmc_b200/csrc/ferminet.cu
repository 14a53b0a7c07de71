// FermiNet forward / forward-Laplacian pipeline (molecule).
//
// Composition follows app/molecule/wavefunction/ferminet.py:76-96 (features -> FermiLayers -> orbitals x
// envelope -> LogDet) with the layer structure of wavefunction/backbone/ferminet.py:29-63.
//
// Restructuring relative to the reference's traced graph (same function, fewer FLOPs):
//   the aggregated input [h_j | mean_up h | mean_dn h | g2_j] is never materialised for layers >= 2.  The Dense
//   kernel is split by rows; the spin-mean part is identical for every electron of a walker, so it is contracted
//   once per walker (rows W*C) and broadcast-added inside the main GEMM's epilogue, and the main GEMM contracts
//   only [h_j | g2_j]  (256+64 wide instead of 832 wide for the default network).
#include "wf.cuh"

// out[w][j][c][:] = [h[w][j][c][0:f1] | m[w][c][0:fm] | g2[w][j][c][0:fg] | 0 ...]   (row length ld)
__global__ void k_concat_agg(const float* __restrict__ h, const float* __restrict__ m, const float* __restrict__ g2,
                             float* __restrict__ out, long long rows, int n, int C, int f1, int fm, int fg, int ld) {
  const long long total = rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % ld);
    const long long row = i / ld;          // (w, j, c)
    float v = 0.f;
    if (col < f1) {
      v = h[row * f1 + col];
    } else if (col < f1 + fm) {
      const int c = (int)(row % C);
      const long long w = row / ((long long)C * n);
      v = m[(w * C + c) * fm + (col - f1)];
    } else if (col < f1 + fm + fg) {
      v = g2[row * fg + (col - f1 - fm)];
    }
    out[i] = v;
  }
}

static int jq_launch_concat_agg(const float* h, const float* m, const float* g2, float* out, long long W, int n, int C,
                                int f1, int fm, int fg, int ld, cudaStream_t st) {
  const long long rows = W * n * C;
  if (rows <= 0) return JQ_OK;
  int grid = jq_cdiv(rows * ld, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  JQ_LAUNCH(k_concat_agg, dim3(grid), dim3(256), 0, st, h, m, g2, out, rows, n, C, f1, fm, fg, ld);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

int jq_fermi_dims(const jaqmc_ferminet_config* c, int track, int fat, int fee, FermiDims* o) {
  JQ_REQUIRE(c->n_layers >= 1 && c->n_layers <= JQ_MAX_LAYERS, JQ_ERR_INVALID_ARGUMENT, "ferminet: n_layers=%d",
             c->n_layers);
  JQ_REQUIRE(c->n_up >= 0 && c->n_dn >= 0 && c->n_up + c->n_dn >= 1, JQ_ERR_INVALID_ARGUMENT, "ferminet: nspins");
  JQ_REQUIRE(c->n_atoms >= 1 && c->n_atoms <= JQ_MAX_ATOMS, JQ_ERR_INVALID_ARGUMENT, "ferminet: n_atoms=%d",
             c->n_atoms);
  JQ_REQUIRE(c->ndets >= 1, JQ_ERR_INVALID_ARGUMENT, "ferminet: ndets=%d", c->ndets);
  o->sp.n_up = c->n_up;
  o->sp.n_dn = c->n_dn;
  o->n = c->n_up + c->n_dn;
  o->A = c->n_atoms;
  o->D = c->ndets;
  o->L = c->n_layers;
  o->nch = o->sp.nch();
  o->C = track ? 3 * o->n + 2 : 1;
  o->C1 = track ? 5 : 1;
  o->C2 = track ? 8 : 1;
  o->fee = fee;
  o->f1 = fat * o->A;
  o->d1max = 0;
  o->d2max = fee;
  for (int l = 0; l < o->L; ++l) {
    o->d1[l] = c->hidden_single[l];
    o->d2[l] = c->hidden_double[l];
    JQ_REQUIRE(o->d1[l] >= 1 && ((l == o->L - 1 && !c->use_last_layer) || o->d2[l] >= 1), JQ_ERR_INVALID_ARGUMENT,
               "ferminet: hidden dims");
    if (o->d1[l] > o->d1max) o->d1max = o->d1[l];
    if ((l < o->L - 1 || c->use_last_layer) && o->d2[l] > o->d2max) o->d2max = o->d2[l];
  }
  o->use_last = c->use_last_layer ? 1 : 0;
  o->agg_w = o->d1[o->L - 1] * (1 + o->nch) + o->nch * o->d2[o->L - 1];
  o->agg_wp = (o->agg_w + 31) / 32 * 32;
  o->in1 = o->f1 * (1 + o->nch) + fee * o->nch;
  o->in1p = (o->in1 + 31) / 32 * 32;   // layer-1 input row length: zero padded to the tensor-core path's K granularity
  JQ_REQUIRE(o->f1 != o->d1[0], JQ_ERR_UNSUPPORTED,
             "ferminet: input feature width == hidden_dims_single[0] (input-layer residual) is not supported");
  return JQ_OK;
}

static JqHeadDims fermi_head_dims(const FermiDims& d, const jaqmc_ferminet_config* c) {
  JqHeadDims hd;
  hd.sp = d.sp;
  hd.A = d.A;
  hd.D = d.D;
  hd.C = d.C;
  hd.hidden = d.use_last ? d.agg_wp : d.d1[d.L - 1];
  hd.hidden_valid = d.use_last ? d.agg_w : 0;
  hd.envelope_type = c->envelope_type;
  hd.split = c->orbitals_spin_split;
  hd.jastrow = 0;
  return hd;
}

void jq_fermi_carve_backbone(const FermiDims& d, long long W, JqArena& ar, FermiBufs* b) {
  long long n = d.n, nn = (long long)d.n * d.n;
  bool pairs = d.L > 1 || d.use_last;
  b->ae = ar.take<float>(W * n * d.C1 * d.f1);
  b->h2a = ar.take<float>(W * nn * d.C2 * d.d2max);
  b->h2b = pairs ? ar.take<float>(W * nn * d.C2 * d.d2max) : nullptr;
  b->g2 = ar.take<float>(W * n * d.C * d.nch * d.d2max);
  b->x1 = ar.take<float>(W * n * d.C * d.in1p);
  b->ha = ar.take<float>(W * n * d.C * d.d1max);
  b->hb = ar.take<float>(W * n * d.C * d.d1max);
  b->m = ar.take<float>(W * d.C * d.nch * d.d1max);
  b->cadd = ar.take<float>(W * d.C * d.d1max);
  b->agg = d.use_last ? ar.take<float>(W * n * d.C * d.agg_wp) : nullptr;
  {
    int kmax = d.d1max * (1 + d.nch) + d.nch * d.d2max;
    if (d.use_last && d.agg_wp > kmax) kmax = d.agg_wp;
    if (d.in1p > kmax) kmax = d.in1p;
    int nmax = d.d1max > d.D * d.n ? d.d1max : d.D * d.n;
    b->wscr = ar.take<float>(jq_dense_tc_scratch_floats(kmax, nmax));
  }
}

static void fermi_carve(const FermiDims& d, const jaqmc_ferminet_config* c, long long W, JqArena& ar, FermiBufs* b) {
  jq_fermi_carve_backbone(d, W, ar, b);
  jq_head_carve(fermi_head_dims(d, c), W, ar, &b->head);
}

// FermiLayers on precomputed features: b.ae [W][n][C1][f1] (Local1), b.h2a [W][n*n][C2][fee] (Local2) -> *h_out
// [W][n][C][hidden_single[-1]].  Shared by the molecular and the periodic FermiNet.
int jq_fermi_backbone(const FermiDims& d, const jaqmc_ferminet_params* p, long long W, int track, const FermiBufs& b,
                      cudaStream_t st, float** h_out) {
  int rc;
  const int n = d.n, C = d.C;
  for (int l = 0; l < d.L; ++l)
    JQ_REQUIRE(p->single_kernel[l] && p->single_bias[l] &&
                   ((l == d.L - 1 && !d.use_last) || (p->double_kernel[l] && p->double_bias[l])),
               JQ_ERR_INVALID_ARGUMENT, "ferminet: null parameter in layer %d", l);
  const int n_double = d.use_last ? d.L : d.L - 1;   // two-electron layers that exist
  float* h2 = b.h2a;
  float* h2n = b.h2b;
  float* h = b.ha;
  float* hn = b.hb;
  int d2prev = d.fee, d1prev = d.f1;
  bool g2_ready = false;   // the previous two-electron layer already produced this layer's spin-channel means
  for (int l = 0; l < d.L; ++l) {
    const int fg = d.nch * d2prev;
    if (!g2_ready && !jq_prep.collect && (rc = jq_launch_pair_mean(h2, b.g2, (int)W, d.sp, d2prev, track, st))) return rc;
    g2_ready = false;
    JqDenseArgs a;
    memset(&a, 0, sizeof(a));
    a.N = d.d1[l];
    a.C = C;
    a.n_sub = n;
    a.n_tot = n;
    a.j0 = 0;
    a.G = W * n;
    a.bias = p->single_bias[l];
    a.act = 1;
    a.wscratch = b.wscr;
    if (l == 0) {
      if (!jq_prep.collect && (rc = jq_launch_concat_layer1(b.ae, b.g2, b.x1, (int)W, d.sp, d.f1, fg, track, d.in1p, st))) return rc;
      a.src0 = b.x1;
      a.k0 = d.in1p;
      a.k0_valid = d.in1;
      a.w0 = p->single_kernel[0];
      a.out = h;
      if ((rc = jq_launch_dense(a, st))) return rc;
    } else {
      // walker-wide part: cadd[w][c] = [mean_up h | mean_dn h] . K[d1prev : d1prev*(1+nch)]
      if (!jq_prep.collect && (rc = jq_launch_spin_mean(h, b.m, (int)W, d.sp, C, d1prev, st))) return rc;
      JqDenseArgs am;
      memset(&am, 0, sizeof(am));
      am.src0 = b.m;
      am.k0 = d.nch * d1prev;
      am.w0 = p->single_kernel[l] + (size_t)d1prev * d.d1[l];
      am.out = b.cadd;
      am.N = d.d1[l];
      am.C = C;
      am.n_sub = 1;
      am.n_tot = 1;
      am.G = W;
      am.wscratch = b.wscr;
      if ((rc = jq_launch_dense(am, st))) return rc;
      a.src0 = h;
      a.k0 = d1prev;
      a.w0 = p->single_kernel[l];
      a.src1 = b.g2;
      a.k1 = fg;
      a.w1 = p->single_kernel[l] + (size_t)d1prev * (1 + d.nch) * d.d1[l];
      a.cadd = b.cadd;
      a.out = hn;
      if (d1prev == d.d1[l]) {
        a.res = h;
        a.res_mode = 1;
      }
      if ((rc = jq_launch_dense(a, st))) return rc;
      float* t = h;
      h = hn;
      hn = t;
    }
    d1prev = d.d1[l];
#ifndef JAQMC_HOST_EMU
    if (l < n_double) {
      // layer + means of its output in one pass over the pair tensor; the last two-electron layer's output is only
      // used through its means and is not written
      const bool last2 = (l == n_double - 1);
      int frc = JQ_OK;
      if (jq_launch_pair_layer_fused(h2, d2prev, p->double_kernel[l], p->double_bias[l], last2 ? nullptr : h2n, b.g2,
                                     (int)W, d.sp, d.d2[l], d2prev == d.d2[l], track, st, &frc)) {
        if (frc) return frc;
        float* t = h2;
        h2 = h2n;
        h2n = t;
        d2prev = d.d2[l];
        g2_ready = true;
        continue;
      }
    }
#endif
    if (l < n_double) {
      JqDenseArgs a2;
      memset(&a2, 0, sizeof(a2));
      a2.src0 = h2;
      a2.k0 = d2prev;
      a2.w0 = p->double_kernel[l];
      a2.bias = p->double_bias[l];
      a2.out = h2n;
      a2.act = 1;
      a2.wscratch = b.wscr;
      if (d2prev == d.d2[l]) {
        a2.res = h2;
        a2.res_mode = 1;
      }
      a2.N = d.d2[l];
      a2.C = d.C2;
      a2.n_sub = n * n;
      a2.n_tot = n * n;
      a2.G = W * n * n;
      if ((rc = jq_launch_dense(a2, st))) return rc;
      float* t = h2;
      h2 = h2n;
      h2n = t;
      d2prev = d.d2[l];
    }
  }

  if (d.use_last) {
    // use_last_layer: the orbitals see aggregate_features(h_one, h_two) = [h_j | spin means of h | pair means of h_two]
    // (backbone/ferminet.py:45-47,65-90), materialised once with zero padding to the tensor-core K granularity
    if (jq_prep.collect) {
      *h_out = b.agg;
      return JQ_OK;
    }
    if (!g2_ready && (rc = jq_launch_pair_mean(h2, b.g2, (int)W, d.sp, d2prev, track, st))) return rc;
    if ((rc = jq_launch_spin_mean(h, b.m, (int)W, d.sp, C, d1prev, st))) return rc;
    if ((rc = jq_launch_concat_agg(h, b.m, b.g2, b.agg, W, n, C, d1prev, d.nch * d1prev, d.nch * d2prev, d.agg_wp, st)))
      return rc;
    *h_out = b.agg;
    return JQ_OK;
  }
  *h_out = h;
  return JQ_OK;
}

size_t jq_ferminet_ws_bytes(const jaqmc_ferminet_config* c, long long W, int track) {
  FermiDims d;
  if (jq_fermi_dims(c, track, 4, 4, &d) != JQ_OK) return 0;
  JqArena ar(nullptr, 0);
  FermiBufs b;
  fermi_carve(d, c, W, ar, &b);
  return ar.off;
}

int jq_ferminet_forward(const jaqmc_ferminet_config* c, const jaqmc_ferminet_params* p, const jaqmc_system* sys,
                        const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOut out,
                        cudaStream_t st) {
  FermiDims d;
  int rc = jq_fermi_dims(c, track, 4, 4, &d);
  if (rc != JQ_OK) return rc;
  JQ_REQUIRE(sys && sys->atoms && sys->n_atoms == d.A, JQ_ERR_INVALID_ARGUMENT, "ferminet: system/atoms mismatch");
  JqArena ar(ws, ws_bytes);
  FermiBufs b;
  fermi_carve(d, c, W, ar, &b);
  JQ_REQUIRE(ar.ok(), JQ_ERR_WORKSPACE_TOO_SMALL, "ferminet: workspace %zu < %zu bytes", ws_bytes, ar.off);
  jaqmc_head_params hp;
  memset(&hp, 0, sizeof(hp));
  for (int s = 0; s < 2; ++s) {
    hp.orbital_kernel[s] = p->orbital_kernel[s];
    hp.env_pi[s] = p->env_pi[s];
    hp.env_sigma[s] = p->env_sigma[s];
  }
  auto run = [&]() -> int {
    int r;
    // features: ae Local1 [W][n][C1][4A], ee Local2 [W][n*n][C2][4] (into h2a)
    if (!jq_prep.collect &&
        (r = jq_launch_mol_features(electrons, sys->atoms, (int)W, d.sp, d.A, /*rescale=*/0, track, /*spin_column=*/0, b.ae, b.h2a, st)))
      return r;
    float* h = nullptr;
    if ((r = jq_fermi_backbone(d, p, W, track, b, st, &h))) return r;
    return jq_head_forward(fermi_head_dims(d, c), &hp, h, electrons, sys->atoms, W, b.head, b.wscr, out, st);
  };
#ifndef JAQMC_HOST_EMU
  if (jq_prep.base && jq_prep.mode == 1 && jq_prep.n == 0) {
    // the launch sequence is known: collect the tensor-core launches' weights in a dry pass, split them all with one
    // kernel, and let the real pass find the splits by key
    jq_prep.collect = 1;
    rc = run();
    jq_prep.collect = 0;
    if (rc) return rc;
    if ((rc = jq_prep_flush(st))) return rc;
    jq_prep.mode = 2;
  }
#endif
  return run();
}
