// LapNet and Psiformer forward / forward-Laplacian pipelines (molecule).
//
// Composition follows app/molecule/wavefunction/lapnet.py:117-135 with backbone/lapnet/_backbone.py:191-263, and
// app/molecule/wavefunction/psiformer.py:137-167 with backbone/psiformer.py:60-99,143-187.  Both end in the shared
// output head (head.cu).
//
// Sparsity kept from the reference (Appendix A of SURVEY.md): input features and LapNet's individual stream stay
// Local1 (5 components per electron); the dense stream carries 3n+2 components.  The dense stream is expanded
// to 3n+2 components right after the input projection (the reference densifies it at the first attention).
#include "wf.cuh"

namespace {

struct AttnDims {
  JqSpins sp;
  int n, A, D, L, H, dh, hid, C, C1, fin;
};

int dense(const float* x, int C, long long G, int n_tot, const float* w, int k, int n_out, int ldw, const float* bias,
          int act, const float* res, int res_mode, float* out, float* wscr, cudaStream_t st) {
  JqDenseArgs a;
  memset(&a, 0, sizeof(a));
  a.src0 = x;
  a.k0 = k;
  a.w0 = w;
  a.ldw = ldw;
  a.bias = bias;
  a.out = out;
  a.N = n_out;
  a.C = C;
  a.n_sub = a.n_tot = n_tot;
  a.G = G;
  a.act = act;
  a.res = res;
  a.res_mode = res_mode;
  a.wscratch = wscr;
  return jq_launch_dense(a, st);
}

JqHeadDims head_dims(const AttnDims& d, int envelope_type, int split, int jastrow) {
  JqHeadDims hd;
  hd.sp = d.sp;
  hd.A = d.A;
  hd.D = d.D;
  hd.C = d.C;
  hd.hidden = d.hid;
  hd.hidden_valid = 0;
  hd.envelope_type = envelope_type;
  hd.split = split;
  hd.jastrow = jastrow;
  return hd;
}

int common_dims(int n_up, int n_dn, int n_atoms, int ndets, int L, int H, int dh, int track, AttnDims* o,
                const char* who) {
  JQ_REQUIRE(L >= 1 && L <= JQ_MAX_LAYERS, JQ_ERR_INVALID_ARGUMENT, "%s: num_layers=%d", who, L);
  JQ_REQUIRE(n_up >= 0 && n_dn >= 0 && n_up + n_dn >= 1, JQ_ERR_INVALID_ARGUMENT, "%s: nspins", who);
  JQ_REQUIRE(n_atoms >= 1 && n_atoms <= JQ_MAX_ATOMS, JQ_ERR_INVALID_ARGUMENT, "%s: n_atoms=%d", who, n_atoms);
  JQ_REQUIRE(ndets >= 1 && H >= 1 && dh >= 1, JQ_ERR_INVALID_ARGUMENT, "%s: ndets/heads", who);
  o->sp.n_up = n_up;
  o->sp.n_dn = n_dn;
  o->n = n_up + n_dn;
  o->A = n_atoms;
  o->D = ndets;
  o->L = L;
  o->H = H;
  o->dh = dh;
  o->hid = H * dh;
  o->C = track ? 3 * o->n + 2 : 1;
  o->C1 = track ? 5 : 1;
  o->fin = 4 * n_atoms + 1;
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// LapNet
// ------------------------------------------------------------------------------------------------
struct LapBufs {
  float *feat, *hs_a, *hs_b, *q[JQ_MAX_LAYERS], *k[JQ_MAX_LAYERS], *hd_a, *hd_b, *v, *att, *wscr, *ln_s, *ln_d;
  JqHeadBufs head;
};

void lap_carve(const AttnDims& d, const jaqmc_lapnet_config* c, long long W, JqArena& ar, LapBufs* b) {
  const long long n = d.n;
  b->feat = ar.take<float>(W * n * d.C1 * d.fin);
  b->hs_a = ar.take<float>(W * n * d.C1 * d.hid);
  b->hs_b = ar.take<float>(W * n * d.C1 * d.hid);
  for (int l = 0; l < d.L; ++l) {
    b->q[l] = ar.take<float>(W * n * d.C1 * d.hid);
    b->k[l] = ar.take<float>(W * n * d.C1 * d.hid);
  }
  b->hd_a = ar.take<float>(W * n * d.C * d.hid);
  b->hd_b = ar.take<float>(W * n * d.C * d.hid);
  b->v = ar.take<float>(W * n * d.C * d.hid);
  b->att = ar.take<float>(W * n * d.C * d.hid);
  b->ln_s = c->use_layernorm ? ar.take<float>(W * n * d.C1 * d.hid) : nullptr;
  b->ln_d = c->use_layernorm ? ar.take<float>(W * n * d.C * d.hid) : nullptr;
  int nmax = d.hid > d.D * d.n ? d.hid : d.D * d.n;
  b->wscr = ar.take<float>(jq_dense_tc_scratch_floats(d.hid, nmax));
  jq_head_carve(head_dims(d, c->envelope_type, 1, 1), W, ar, &b->head);
}

}  // namespace

size_t jq_lapnet_ws_bytes(const jaqmc_lapnet_config* c, long long W, int track) {
  AttnDims d;
  if (common_dims(c->n_up, c->n_dn, c->n_atoms, c->ndets, c->num_layers, c->num_heads, c->heads_dim, track, &d,
                  "lapnet") != JQ_OK)
    return 0;
  JqArena ar(nullptr, 0);
  LapBufs b;
  lap_carve(d, c, W, ar, &b);
  return ar.off;
}

int jq_lapnet_forward(const jaqmc_lapnet_config* c, const jaqmc_lapnet_params* p, const jaqmc_system* sys,
                      const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOut out,
                      cudaStream_t st) {
  AttnDims d;
  int rc = common_dims(c->n_up, c->n_dn, c->n_atoms, c->ndets, c->num_layers, c->num_heads, c->heads_dim, track, &d,
                       "lapnet");
  if (rc) return rc;
  JQ_REQUIRE(c->num_local_updates >= 0 && c->num_local_updates <= 4, JQ_ERR_INVALID_ARGUMENT,
             "lapnet: num_local_updates=%d", c->num_local_updates);
  JQ_REQUIRE(sys && sys->atoms && sys->n_atoms == d.A, JQ_ERR_INVALID_ARGUMENT, "lapnet: system/atoms mismatch");
  JQ_REQUIRE(p->input_kernel, JQ_ERR_INVALID_ARGUMENT, "lapnet: null input projection");
  for (int l = 0; l < d.L; ++l) {
    JQ_REQUIRE(p->qk_kernel[l] && p->value_kernel[l] && p->output_kernel[l] && p->update_kernel[l],
               JQ_ERR_INVALID_ARGUMENT, "lapnet: null kernel in layer %d", l);
    if (l < d.L - 1)
      for (int j = 0; j < c->num_local_updates; ++j)
        JQ_REQUIRE(p->qk_update_kernel[l][j], JQ_ERR_INVALID_ARGUMENT, "lapnet: null qk_update kernel %d/%d", l, j);
    JQ_REQUIRE(!c->use_layernorm || (p->qk_ln_scale[l] && p->qk_ln_bias[l] && p->value_ln_scale[l] && p->value_ln_bias[l] &&
                                     p->post_ln_scale[l] && p->post_ln_bias[l]),
               JQ_ERR_INVALID_ARGUMENT, "lapnet: null LayerNorm parameter in layer %d", l);
  }
  const float ln_eps = 1e-6f;   // _backbone.py:62-64
  JqArena ar(ws, ws_bytes);
  LapBufs b;
  lap_carve(d, c, W, ar, &b);
  JQ_REQUIRE(ar.ok(), JQ_ERR_WORKSPACE_TOO_SMALL, "lapnet: workspace %zu < %zu bytes", ws_bytes, ar.off);
  const int n = d.n, hid = d.hid;
  const long long G = W * n;

  if ((rc = jq_launch_mol_features(electrons, sys->atoms, (int)W, d.sp, d.A, c->rescale, track, /*spin_column=*/1,
                                   b.feat, nullptr, st)))
    return rc;
  // individual stream (Local1): hs, per-layer q / k
  float* hs = b.hs_a;
  float* hs_n = b.hs_b;
  if ((rc = dense(b.feat, d.C1, G, n, p->input_kernel, d.fin, hid, 0, p->input_bias, 0, nullptr, 0, hs, b.wscr, st)))
    return rc;
  // dense stream starts as a copy of the projected input (hd0 = hs, _backbone.py:216-218)
  float* hd = b.hd_a;
  float* hd_n = b.hd_b;
  if (track) {
    if ((rc = jq_launch_densify_local1(hs, hd, W, n, hid, st))) return rc;
  } else {
    cudaMemcpyAsyncD2D(hd, hs, sizeof(float) * G * hid, st);
  }
  // The first value projection sees hd0 = hs0, still a one-electron (5-row) tensor: project in that form and expand the
  // result (r2; 1.64 -> 0.27 + 0.5 ms at N2).  Not with use_layernorm (the value LayerNorm comes first).
  static const bool dense_v0 = getenv("JAQMC_B200_LAPNET_DENSE_VALUE0") != nullptr;   // A/B switch
  const bool local_v0 = track && !c->use_layernorm && !dense_v0;
  if (local_v0) {
    if ((rc = dense(hs, d.C1, G, n, p->value_kernel[0], hid, hid, 0, p->value_bias[0], 0, nullptr, 0, b.att, b.wscr, st)))
      return rc;
    if ((rc = jq_launch_densify_local1(b.att, b.v, W, n, hid, st))) return rc;
  }
  for (int l = 0; l < d.L; ++l) {
    const float* hs_in = hs;
    if (c->use_layernorm) {   // project_qk_stream: qk_layernorm first (_backbone.py:121)
      if ((rc = jq_launch_layernorm_fl(hs, p->qk_ln_scale[l], p->qk_ln_bias[l], b.ln_s, G, d.C1, hid, ln_eps, st))) return rc;
      hs_in = b.ln_s;
    }
    if ((rc = dense(hs_in, d.C1, G, n, p->qk_kernel[l], hid, hid, 2 * hid, p->qk_bias[l], 0, nullptr, 0, b.q[l], b.wscr, st)))
      return rc;
    if ((rc = dense(hs_in, d.C1, G, n, p->qk_kernel[l] + hid, hid, hid, 2 * hid, p->qk_bias[l] ? p->qk_bias[l] + hid : nullptr,
                    0, nullptr, 0, b.k[l], b.wscr, st)))
      return rc;
    if (l < d.L - 1)
      for (int j = 0; j < c->num_local_updates; ++j) {
        if ((rc = dense(hs, d.C1, G, n, p->qk_update_kernel[l][j], hid, hid, 0, p->qk_update_bias[l][j], 1, hs, 2, hs_n,
                        b.wscr, st)))
          return rc;
        float* t = hs;
        hs = hs_n;
        hs_n = t;
      }
  }
  for (int l = 0; l < d.L; ++l) {
    const float* v_in = hd;
    if (c->use_layernorm) {   // value_layernorm (_backbone.py:92-95)
      if ((rc = jq_launch_layernorm_fl(hd, p->value_ln_scale[l], p->value_ln_bias[l], b.ln_d, G, d.C, hid, ln_eps, st)))
        return rc;
      v_in = b.ln_d;
    }
    if (!(l == 0 && local_v0) &&
        (rc = dense(v_in, d.C, G, n, p->value_kernel[l], hid, hid, 0, p->value_bias[l], 0, nullptr, 0, b.v, b.wscr, st)))
      return rc;
    JqAttnOperand q = {b.q[l], d.C1, hid, 0}, k = {b.k[l], d.C1, hid, 0}, v = {b.v, d.C, hid, 0};
    if ((rc = jq_launch_attention_fl(q, k, v, b.att, hid, W, n, d.H, d.dh, track, st))) return rc;
    // residual_value = hd + output_projection(att)
    if ((rc = dense(b.att, d.C, G, n, p->output_kernel[l], hid, hid, 0, p->output_bias[l], 0, hd, 2, hd_n, b.wscr, st)))
      return rc;
    // hd = residual_value + tanh(value_update(post_attention_layernorm(residual_value)))   (_backbone.py:105-111)
    const float* u_in = hd_n;
    if (c->use_layernorm) {
      if ((rc = jq_launch_layernorm_fl(hd_n, p->post_ln_scale[l], p->post_ln_bias[l], b.ln_d, G, d.C, hid, ln_eps, st)))
        return rc;
      u_in = b.ln_d;
    }
    if ((rc = dense(u_in, d.C, G, n, p->update_kernel[l], hid, hid, 0, p->update_bias[l], 1, hd_n, 2, hd, b.wscr, st)))
      return rc;
  }
  return jq_head_forward(head_dims(d, c->envelope_type, 1, p->head.jastrow_alpha_par != nullptr), &p->head, hd,
                         electrons, sys->atoms, W, b.head, b.wscr, out, st);
}

// ------------------------------------------------------------------------------------------------
// Psiformer
// ------------------------------------------------------------------------------------------------
namespace {
struct PsiBufs {
  float *feat, *x0, *xa, *xb, *ln, *q, *k, *v, *att, *m1, *m2, *wscr;
  JqHeadBufs head;
};

int psi_mlp_max(const jaqmc_psiformer_config* c, int hid) {
  int m = hid;
  for (int j = 0; j < c->n_mlp_hidden; ++j)
    if (c->mlp_hidden[j] > m) m = c->mlp_hidden[j];
  return m;
}

void psi_carve(const AttnDims& d, const jaqmc_psiformer_config* c, long long W, JqArena& ar, PsiBufs* b) {
  const long long n = d.n;
  const int mmax = psi_mlp_max(c, d.hid);
  b->feat = ar.take<float>(W * n * d.C1 * d.fin);
  b->x0 = ar.take<float>(W * n * d.C1 * d.hid);
  b->xa = ar.take<float>(W * n * d.C * d.hid);
  b->xb = ar.take<float>(W * n * d.C * d.hid);
  b->ln = ar.take<float>(W * n * d.C * d.hid);
  b->q = ar.take<float>(W * n * d.C * d.hid);
  b->k = ar.take<float>(W * n * d.C * d.hid);
  b->v = ar.take<float>(W * n * d.C * d.hid);
  b->att = ar.take<float>(W * n * d.C * d.hid);
  b->m1 = ar.take<float>(W * n * d.C * mmax);
  b->m2 = ar.take<float>(W * n * d.C * mmax);
  int nmax = mmax > d.D * d.n ? mmax : d.D * d.n;
  b->wscr = ar.take<float>(jq_dense_tc_scratch_floats(mmax, nmax));
  jq_head_carve(head_dims(d, c->envelope_type, c->orbitals_spin_split, 1), W, ar, &b->head);
}
}  // namespace

size_t jq_psiformer_ws_bytes(const jaqmc_psiformer_config* c, long long W, int track) {
  AttnDims d;
  if (common_dims(c->n_up, c->n_dn, c->n_atoms, c->ndets, c->num_layers, c->num_heads, c->heads_dim, track, &d,
                  "psiformer") != JQ_OK)
    return 0;
  if (c->n_mlp_hidden < 0 || c->n_mlp_hidden > JAQMC_MAX_MLP - 1) return 0;
  JqArena ar(nullptr, 0);
  PsiBufs b;
  psi_carve(d, c, W, ar, &b);
  return ar.off;
}

int jq_psiformer_forward(const jaqmc_psiformer_config* c, const jaqmc_psiformer_params* p, const jaqmc_system* sys,
                         const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOut out,
                         cudaStream_t st) {
  AttnDims d;
  int rc = common_dims(c->n_up, c->n_dn, c->n_atoms, c->ndets, c->num_layers, c->num_heads, c->heads_dim, track, &d,
                       "psiformer");
  if (rc) return rc;
  JQ_REQUIRE(c->n_mlp_hidden >= 0 && c->n_mlp_hidden <= JAQMC_MAX_MLP - 1, JQ_ERR_INVALID_ARGUMENT,
             "psiformer: %d hidden MLP layers (max %d)", c->n_mlp_hidden, JAQMC_MAX_MLP - 1);
  JQ_REQUIRE(c->layer_norm_mode >= 0 && c->layer_norm_mode <= 2, JQ_ERR_INVALID_ARGUMENT, "psiformer: layer_norm_mode");
  JQ_REQUIRE(sys && sys->atoms && sys->n_atoms == d.A, JQ_ERR_INVALID_ARGUMENT, "psiformer: system/atoms mismatch");
  JQ_REQUIRE(p->input_kernel, JQ_ERR_INVALID_ARGUMENT, "psiformer: null input projection");
  const bool use_ln = c->layer_norm_mode != JAQMC_LAYERNORM_NULL;
  for (int l = 0; l < d.L; ++l) {
    JQ_REQUIRE(p->q_kernel[l] && p->k_kernel[l] && p->v_kernel[l] && p->out_kernel[l], JQ_ERR_INVALID_ARGUMENT,
               "psiformer: null attention kernel in layer %d", l);
    JQ_REQUIRE(!use_ln || (p->ln0_scale[l] && p->ln0_bias[l] && p->ln1_scale[l] && p->ln1_bias[l]),
               JQ_ERR_INVALID_ARGUMENT, "psiformer: null LayerNorm parameter in layer %d", l);
    JQ_REQUIRE(c->layer_norm_mode != JAQMC_LAYERNORM_POST || (p->ln0_scale[l] != nullptr), JQ_ERR_INVALID_ARGUMENT,
               "psiformer: post-LN needs LayerNorm parameters");
    for (int j = 0; j <= c->n_mlp_hidden; ++j)
      JQ_REQUIRE(p->mlp_kernel[l][j], JQ_ERR_INVALID_ARGUMENT, "psiformer: null MLP kernel %d/%d", l, j);
  }
  JqArena ar(ws, ws_bytes);
  PsiBufs b;
  psi_carve(d, c, W, ar, &b);
  JQ_REQUIRE(ar.ok(), JQ_ERR_WORKSPACE_TOO_SMALL, "psiformer: workspace %zu < %zu bytes", ws_bytes, ar.off);
  const int n = d.n, hid = d.hid, C = d.C;
  const long long G = W * n;
  const float eps = 1e-5f;

  if ((rc = jq_launch_mol_features(electrons, sys->atoms, (int)W, d.sp, d.A, c->rescale, track, /*spin_column=*/1,
                                   b.feat, nullptr, st)))
    return rc;
  float* x = b.xa;
  float* xn = b.xb;
  if (track) {
    if ((rc = dense(b.feat, d.C1, G, n, p->input_kernel, d.fin, hid, 0, p->input_bias, 0, nullptr, 0, b.x0, b.wscr, st)))
      return rc;
    if ((rc = jq_launch_densify_local1(b.x0, x, W, n, hid, st))) return rc;
  } else {
    if ((rc = dense(b.feat, 1, G, n, p->input_kernel, d.fin, hid, 0, p->input_bias, 0, nullptr, 0, x, b.wscr, st)))
      return rc;
  }
  for (int l = 0; l < d.L; ++l) {
    // attention block: x = x + out(MHA(LN0(x)))            (pre-LN; post / null: MHA(x))
    static const bool dense_l0 = getenv("JAQMC_B200_PSIFORMER_DENSE_LAYER0") != nullptr;   // A/B switch
    if (l == 0 && track && !dense_l0) {
      // The first block sees the projected one-electron features: electron i's row depends on r_i only (Local1, 5 rows
      // per group instead of 3n + 2).  LayerNorm and the q / k / v projections act row-wise, so they run on the 5-row
      // groups (r2: 44 -> 5 rows at N2 for one LayerNorm and three dense launches); q and k enter the attention kernel
      // in that form (its one-electron phase), v is expanded first.  The residual uses the expanded x as before.
      const float* xin = b.x0;
      if (c->layer_norm_mode == JAQMC_LAYERNORM_PRE) {
        if ((rc = jq_launch_layernorm_fl(b.x0, p->ln0_scale[l], p->ln0_bias[l], b.ln, G, d.C1, hid, eps, st))) return rc;
        xin = b.ln;
      }
      if ((rc = dense(xin, d.C1, G, n, p->q_kernel[l], hid, hid, 0, p->q_bias[l], 0, nullptr, 0, b.q, b.wscr, st))) return rc;
      if ((rc = dense(xin, d.C1, G, n, p->k_kernel[l], hid, hid, 0, p->k_bias[l], 0, nullptr, 0, b.k, b.wscr, st))) return rc;
      if ((rc = dense(xin, d.C1, G, n, p->v_kernel[l], hid, hid, 0, p->v_bias[l], 0, nullptr, 0, b.m1, b.wscr, st))) return rc;
      if ((rc = jq_launch_densify_local1(b.m1, b.v, W, n, hid, st))) return rc;
      JqAttnOperand q = {b.q, d.C1, hid, 0}, k = {b.k, d.C1, hid, 0}, v = {b.v, C, hid, 0};
      if ((rc = jq_launch_attention_fl(q, k, v, b.att, hid, W, n, d.H, d.dh, track, st))) return rc;
    } else {
    const float* xin = x;
    if (c->layer_norm_mode == JAQMC_LAYERNORM_PRE) {
      if ((rc = jq_launch_layernorm_fl(x, p->ln0_scale[l], p->ln0_bias[l], b.ln, G, C, hid, eps, st))) return rc;
      xin = b.ln;
    }
    if ((rc = dense(xin, C, G, n, p->q_kernel[l], hid, hid, 0, p->q_bias[l], 0, nullptr, 0, b.q, b.wscr, st))) return rc;
    if ((rc = dense(xin, C, G, n, p->k_kernel[l], hid, hid, 0, p->k_bias[l], 0, nullptr, 0, b.k, b.wscr, st))) return rc;
    if ((rc = dense(xin, C, G, n, p->v_kernel[l], hid, hid, 0, p->v_bias[l], 0, nullptr, 0, b.v, b.wscr, st))) return rc;
    JqAttnOperand q = {b.q, C, hid, 0}, k = {b.k, C, hid, 0}, v = {b.v, C, hid, 0};
    if ((rc = jq_launch_attention_fl(q, k, v, b.att, hid, W, n, d.H, d.dh, track, st))) return rc;
    }
    if ((rc = dense(b.att, C, G, n, p->out_kernel[l], hid, hid, 0, p->out_bias[l], 0, x, 2, xn, b.wscr, st))) return rc;
    {
      float* t = x;
      x = xn;
      xn = t;
    }
    // MLP block
    const float* m = x;
    if (c->layer_norm_mode == JAQMC_LAYERNORM_PRE) {
      if ((rc = jq_launch_layernorm_fl(x, p->ln1_scale[l], p->ln1_bias[l], b.ln, G, C, hid, eps, st))) return rc;
      m = b.ln;
    } else if (c->layer_norm_mode == JAQMC_LAYERNORM_POST) {
      // mlp_out = x = LayerNorm(x)   (backbone/psiformer.py:87-88)
      if ((rc = jq_launch_layernorm_fl(x, p->ln0_scale[l], p->ln0_bias[l], xn, G, C, hid, eps, st))) return rc;
      float* t = x;
      x = xn;
      xn = t;
      m = x;
    }
    int kin = hid;
    float* mo = b.m1;
    float* mo2 = b.m2;
    for (int j = 0; j < c->n_mlp_hidden; ++j) {
      if ((rc = dense(m, C, G, n, p->mlp_kernel[l][j], kin, c->mlp_hidden[j], 0, p->mlp_bias[l][j], 1, nullptr, 0, mo,
                      b.wscr, st)))
        return rc;
      m = mo;
      float* t = mo;
      mo = mo2;
      mo2 = t;
      kin = c->mlp_hidden[j];
    }
    // x = x + tanh(Dense(m))
    if ((rc = dense(m, C, G, n, p->mlp_kernel[l][c->n_mlp_hidden], kin, hid, 0, p->mlp_bias[l][c->n_mlp_hidden], 1, x, 2,
                    xn, b.wscr, st)))
      return rc;
    {
      float* t = x;
      x = xn;
      xn = t;
    }
    if (c->layer_norm_mode == JAQMC_LAYERNORM_POST) {
      if ((rc = jq_launch_layernorm_fl(x, p->ln1_scale[l], p->ln1_bias[l], xn, G, C, hid, eps, st))) return rc;
      float* t = x;
      x = xn;
      xn = t;
    }
  }
  return jq_head_forward(head_dims(d, c->envelope_type, c->orbitals_spin_split, p->head.jastrow_alpha_par != nullptr),
                         &p->head, x, electrons, sys->atoms, W, b.head, b.wscr, out, st);
}
