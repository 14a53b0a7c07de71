// Parameter leaves -> descriptor structs: the ONE place that knows how the reference's Flax parameter trees
// (verified against the reference's own modules by tests/test_reference_fixtures.py) map onto the
// pointer fields of include/jaqmc_b200.h.
//
// An XLA-FFI handler receives the parameters as a flat operand list in `jax.tree.leaves(params)` order: dictionary
// keys sorted (Python string order) at every level of the tree.  jaqmc_b200_param_leaf_count / _param_leaf_info /
// _bind_param_leaves enumerate the leaves of a wavefunction kind + config in exactly that order, so that the shim
// (ffi/xla_ffi_shim.cc) and the ctypes host (jaqmc_b200/_marshal.py is cross-checked against this in
// tests/test_leaf_binding.py) never hand-roll the order.  Note the traps this removes: "_env_down" sorts before
// "_env_up", "bias" before "kernel", "Dense_10" before "Dense_2".
// Host-only code (no kernels).
#include <algorithm>
#include <string>
#include <vector>

#include "wf.cuh"

namespace {

struct Leaf {
  std::vector<std::string> path;   // key components below "params"
  const float** slot;              // field of the params struct that receives the pointer
  long long count;                 // number of float32 elements
  int rank;
  long long dims[4];
};

struct LeafList {
  std::vector<Leaf> leaves;
  void add(std::vector<std::string> path, const float** slot, std::initializer_list<long long> dims) {
    Leaf l;
    l.path = std::move(path);
    l.slot = slot;
    l.rank = (int)dims.size();
    l.count = 1;
    int i = 0;
    for (long long d : dims) {
      l.dims[i++] = d;
      l.count *= d;
    }
    leaves.push_back(std::move(l));
  }
  void sort() {
    std::sort(leaves.begin(), leaves.end(), [](const Leaf& a, const Leaf& b) { return a.path < b.path; });
  }
};

std::string idx(const char* base, int i) { return std::string(base) + std::to_string(i); }

// orbital_layer / envelope_layer / jastrow_layer of the shared head (output/orbital.py:59-78, output/envelope.py:98-163,
// jastrow.py:61-63)
void head_leaves(LeafList& L, const char* orbital_name, const float** okern, const float** obias, bool with_bias,
                 const float** pi, const float** sigma, const float** a_par, const float** a_anti, int n_up, int n_dn,
                 int A, int D, int hidden, bool split_flag, int envelope_type, bool jastrow) {
  const long long n = n_up + n_dn;
  const bool two = n_up > 0 && n_dn > 0;
  const bool split = split_flag && two;
  if (split) {
    for (int s = 0; s < 2; ++s) {
      L.add({orbital_name, "SplitChannelDense_0", idx("DenseGeneral_", s), "kernel"}, okern + s, {hidden, D, n});
      if (with_bias) L.add({orbital_name, "SplitChannelDense_0", idx("DenseGeneral_", s), "bias"}, obias + s, {D, n});
    }
  } else {
    L.add({orbital_name, "DenseGeneral_0", "kernel"}, okern, {hidden, D, n});
    if (with_bias) L.add({orbital_name, "DenseGeneral_0", "bias"}, obias, {D, n});
  }
  if (pi && envelope_type != JAQMC_ENVELOPE_NULL) {
    const char* names2[2] = {"_env_up", "_env_down"};
    for (int s = 0; s < (split ? 2 : 1); ++s) {
      const char* nm = split ? names2[s] : "_env";
      L.add({"envelope_layer", nm, "pi"}, pi + s, {n, A, D});
      if (envelope_type == JAQMC_ENVELOPE_DIAGONAL) L.add({"envelope_layer", nm, "sigma"}, sigma + s, {n, A, 3, D});
      else L.add({"envelope_layer", nm, "sigma"}, sigma + s, {n, A, D});
    }
  }
  if (jastrow) {
    L.add({"jastrow_layer", "alpha_par"}, a_par, {1});
    L.add({"jastrow_layer", "alpha_anti"}, a_anti, {1});
  }
}

void fermi_backbone_leaves(LeafList& L, const jaqmc_ferminet_config* c, jaqmc_ferminet_params* p, int fat, int fee) {
  const int nch = (c->n_up > 0 && c->n_dn > 0) ? 2 : 1;
  long long d1 = (long long)fat * c->n_atoms, d2 = fee;
  int k = 0;
  for (int l = 0; l < c->n_layers; ++l) {
    const long long h1 = c->hidden_single[l];
    L.add({"backbone_layer", idx("Dense_", k), "kernel"}, &p->single_kernel[l], {d1 * (1 + nch) + d2 * nch, h1});
    L.add({"backbone_layer", idx("Dense_", k), "bias"}, &p->single_bias[l], {h1});
    ++k;
    if (l < c->n_layers - 1 || c->use_last_layer) {
      const long long h2 = c->hidden_double[l];
      L.add({"backbone_layer", idx("Dense_", k), "kernel"}, &p->double_kernel[l], {d2, h2});
      L.add({"backbone_layer", idx("Dense_", k), "bias"}, &p->double_bias[l], {h2});
      ++k;
      d2 = h2;
    }
    d1 = h1;
  }
}

int fermi_orbital_width(const jaqmc_ferminet_config* c) {
  const int nch = (c->n_up > 0 && c->n_dn > 0) ? 2 : 1;
  const int L = c->n_layers;
  int d1 = c->hidden_single[L - 1];
  if (c->use_last_layer) d1 = d1 * (1 + nch) + nch * c->hidden_double[L - 1];
  return d1;
}

int build(int kind, const void* config, void* params, LeafList& L) {
  switch (kind) {
    case JAQMC_WF_FERMINET: {
      auto* c = (const jaqmc_ferminet_config*)config;
      auto* p = (jaqmc_ferminet_params*)params;
      JQ_REQUIRE(c->n_layers >= 1 && c->n_layers <= JQ_MAX_LAYERS, JQ_ERR_INVALID_ARGUMENT, "leaves: n_layers=%d", c->n_layers);
      fermi_backbone_leaves(L, c, p, 4, 4);
      head_leaves(L, "orbital_layer", p->orbital_kernel, nullptr, false, p->env_pi, p->env_sigma, nullptr, nullptr, c->n_up,
                  c->n_dn, c->n_atoms, c->ndets, fermi_orbital_width(c), c->orbitals_spin_split, c->envelope_type, false);
      break;
    }
    case JAQMC_WF_SOLID_FERMINET: {
      auto* c = (const jaqmc_solid_config*)config;
      auto* p = (jaqmc_solid_params*)params;
      JQ_REQUIRE(c->net.n_layers >= 1 && c->net.n_layers <= JQ_MAX_LAYERS, JQ_ERR_INVALID_ARGUMENT, "leaves: n_layers");
      fermi_backbone_leaves(L, &c->net, &p->net, c->distance_type == JAQMC_DISTANCE_NU ? 4 : 7,
                            c->distance_type == JAQMC_DISTANCE_NU ? 4 : 7);
      const int hid = fermi_orbital_width(&c->net);
      head_leaves(L, "real_orbital_layer", p->real_orbital_kernel, nullptr, false, p->net.env_pi, p->net.env_sigma, nullptr,
                  nullptr, c->net.n_up, c->net.n_dn, c->net.n_atoms, c->net.ndets, hid, c->net.orbitals_spin_split,
                  c->net.envelope_type, false);
      head_leaves(L, "imag_orbital_layer", p->imag_orbital_kernel, nullptr, false, nullptr, nullptr, nullptr, nullptr,
                  c->net.n_up, c->net.n_dn, c->net.n_atoms, c->net.ndets, hid, c->net.orbitals_spin_split,
                  JAQMC_ENVELOPE_NULL, false);
      break;   // klist is a module attribute, not a parameter: the caller sets params->klist
    }
    case JAQMC_WF_LAPNET: {
      auto* c = (const jaqmc_lapnet_config*)config;
      auto* p = (jaqmc_lapnet_params*)params;
      JQ_REQUIRE(c->num_layers >= 1 && c->num_layers <= JQ_MAX_LAYERS, JQ_ERR_INVALID_ARGUMENT, "leaves: num_layers");
      const long long hid = (long long)c->num_heads * c->heads_dim;
      const bool ib = p->input_bias != nullptr;   // bias presence is part of the tree: flagged by the caller (see header)
      L.add({"backbone_layer", "input_projection", "kernel"}, &p->input_kernel, {4LL * c->n_atoms + 1, hid});
      if (ib) L.add({"backbone_layer", "input_projection", "bias"}, &p->input_bias, {hid});
      for (int l = 0; l < c->num_layers; ++l) {
        const std::string ly = idx("layers_", l);
        const bool bb = p->qk_bias[l] != nullptr;
        L.add({"backbone_layer", ly, "qk_projection", "kernel"}, &p->qk_kernel[l], {hid, 2 * hid});
        if (bb) L.add({"backbone_layer", ly, "qk_projection", "bias"}, &p->qk_bias[l], {2 * hid});
        const char* nm[3] = {"value_projection", "output_projection", "value_update"};
        const float** ks[3] = {&p->value_kernel[l], &p->output_kernel[l], &p->update_kernel[l]};
        const float** bs[3] = {&p->value_bias[l], &p->output_bias[l], &p->update_bias[l]};
        for (int q = 0; q < 3; ++q) {
          L.add({"backbone_layer", ly, nm[q], "kernel"}, ks[q], {hid, hid});
          if (bb) L.add({"backbone_layer", ly, nm[q], "bias"}, bs[q], {hid});
        }
        if (l < c->num_layers - 1)
          for (int j = 0; j < c->num_local_updates; ++j) {
            L.add({"backbone_layer", ly, idx("qk_update_layers_", j), "kernel"}, &p->qk_update_kernel[l][j], {hid, hid});
            if (bb) L.add({"backbone_layer", ly, idx("qk_update_layers_", j), "bias"}, &p->qk_update_bias[l][j], {hid});
          }
        if (c->use_layernorm) {
          const char* ln[3] = {"qk_layernorm", "value_layernorm", "post_attention_layernorm"};
          const float** sc[3] = {&p->qk_ln_scale[l], &p->value_ln_scale[l], &p->post_ln_scale[l]};
          const float** bi[3] = {&p->qk_ln_bias[l], &p->value_ln_bias[l], &p->post_ln_bias[l]};
          for (int q = 0; q < 3; ++q) {
            L.add({"backbone_layer", ly, ln[q], "scale"}, sc[q], {hid});
            L.add({"backbone_layer", ly, ln[q], "bias"}, bi[q], {hid});
          }
        }
      }
      head_leaves(L, "orbital_layer", p->head.orbital_kernel, p->head.orbital_bias, p->head.orbital_bias[0] != nullptr,
                  p->head.env_pi, p->head.env_sigma, &p->head.jastrow_alpha_par, &p->head.jastrow_alpha_anti, c->n_up,
                  c->n_dn, c->n_atoms, c->ndets, (int)hid, true, c->envelope_type, p->head.jastrow_alpha_par != nullptr);
      break;
    }
    case JAQMC_WF_PSIFORMER: {
      auto* c = (const jaqmc_psiformer_config*)config;
      auto* p = (jaqmc_psiformer_params*)params;
      JQ_REQUIRE(c->num_layers >= 1 && c->num_layers <= JQ_MAX_LAYERS && c->n_mlp_hidden >= 0 &&
                     c->n_mlp_hidden < JAQMC_MAX_MLP,
                 JQ_ERR_INVALID_ARGUMENT, "leaves: num_layers / mlp");
      const long long H = c->num_heads, dh = c->heads_dim, hid = H * dh;
      L.add({"backbone_layer", "Dense_0", "kernel"}, &p->input_kernel, {4LL * c->n_atoms + 1, hid});
      if (p->input_bias) L.add({"backbone_layer", "Dense_0", "bias"}, &p->input_bias, {hid});
      for (int l = 0; l < c->num_layers; ++l) {
        const std::string ly = idx("PsiformerLayer_", l);
        const bool wb = p->q_bias[l] != nullptr;
        if (c->layer_norm_mode != JAQMC_LAYERNORM_NULL) {
          L.add({"backbone_layer", ly, "LayerNorm_0", "scale"}, &p->ln0_scale[l], {hid});
          L.add({"backbone_layer", ly, "LayerNorm_0", "bias"}, &p->ln0_bias[l], {hid});
          L.add({"backbone_layer", ly, "LayerNorm_1", "scale"}, &p->ln1_scale[l], {hid});
          L.add({"backbone_layer", ly, "LayerNorm_1", "bias"}, &p->ln1_bias[l], {hid});
        }
        const char* nm[3] = {"query", "key", "value"};
        const float** ks[3] = {&p->q_kernel[l], &p->k_kernel[l], &p->v_kernel[l]};
        const float** bs[3] = {&p->q_bias[l], &p->k_bias[l], &p->v_bias[l]};
        for (int q = 0; q < 3; ++q) {
          L.add({"backbone_layer", ly, "MultiHeadDotProductAttention_0", nm[q], "kernel"}, ks[q], {hid, H, dh});
          if (wb) L.add({"backbone_layer", ly, "MultiHeadDotProductAttention_0", nm[q], "bias"}, bs[q], {H, dh});
        }
        L.add({"backbone_layer", ly, "MultiHeadDotProductAttention_0", "out", "kernel"}, &p->out_kernel[l], {H, dh, hid});
        if (wb) L.add({"backbone_layer", ly, "MultiHeadDotProductAttention_0", "out", "bias"}, &p->out_bias[l], {hid});
        long long fan = hid;
        for (int j = 0; j <= c->n_mlp_hidden; ++j) {
          const long long w = (j < c->n_mlp_hidden) ? c->mlp_hidden[j] : hid;
          L.add({"backbone_layer", ly, idx("Dense_", j), "kernel"}, &p->mlp_kernel[l][j], {fan, w});
          L.add({"backbone_layer", ly, idx("Dense_", j), "bias"}, &p->mlp_bias[l][j], {w});
          fan = w;
        }
      }
      head_leaves(L, "orbital_layer", p->head.orbital_kernel, p->head.orbital_bias, p->head.orbital_bias[0] != nullptr,
                  p->head.env_pi, p->head.env_sigma, &p->head.jastrow_alpha_par, &p->head.jastrow_alpha_anti, c->n_up,
                  c->n_dn, c->n_atoms, c->ndets, (int)hid, c->orbitals_spin_split, c->envelope_type,
                  p->head.jastrow_alpha_par != nullptr);
      break;
    }
    case JAQMC_WF_HYDROGEN: {
      auto* p = (jaqmc_hydrogen_params*)params;
      L.add({"alpha"}, &p->alpha, {1});
      break;
    }
    default:
      jq_set_error("leaves: wavefunction kind %d is not implemented", kind);
      return JQ_ERR_UNSUPPORTED;
  }
  L.sort();
  return JQ_OK;
}

}  // namespace

// Optional leaves (biases, Jastrow) are part of the tree or not depending on the reference class's flags.  The caller
// states which exist by pre-setting the corresponding pointer fields of `params` to any non-NULL value (e.g. (float*)1)
// before the call: input_bias, qk_bias[l] (all backbone biases of LapNet layer l), q_bias[l] (with_bias of Psiformer
// layer l), head.orbital_bias[0], head.jastrow_alpha_par.  Every other field is ignored on entry and overwritten.
extern "C" int jaqmc_b200_param_leaf_count(int32_t kind, const void* config, void* params) {
  LeafList L;
  if (!config || !params || build(kind, config, params, L) != JQ_OK) return -1;
  return (int)L.leaves.size();
}

extern "C" int jaqmc_b200_param_leaf_info(int32_t kind, const void* config, void* params, int32_t index, char* path,
                                          size_t path_cap, int64_t* n_elements, int32_t* rank, int64_t* dims) {
  LeafList L;
  JQ_REQUIRE(config && params, JQ_ERR_INVALID_ARGUMENT, "leaf_info: null descriptor");
  int rc = build(kind, config, params, L);
  if (rc) return rc;
  JQ_REQUIRE(index >= 0 && index < (int)L.leaves.size(), JQ_ERR_INVALID_ARGUMENT, "leaf_info: index %d of %zu", index,
             L.leaves.size());
  const Leaf& l = L.leaves[index];
  if (path && path_cap) {
    std::string s = "params";
    for (auto& k : l.path) s += "/" + k;
    snprintf(path, path_cap, "%s", s.c_str());
  }
  if (n_elements) *n_elements = l.count;
  if (rank) *rank = l.rank;
  if (dims)
    for (int i = 0; i < l.rank; ++i) dims[i] = l.dims[i];
  return JQ_OK;
}

extern "C" int jaqmc_b200_bind_param_leaves(int32_t kind, const void* config, void* params, const float* const* leaves,
                                            const int64_t* leaf_elements, int32_t n_leaves) {
  LeafList L;
  JQ_REQUIRE(config && params && (leaves || n_leaves == 0), JQ_ERR_INVALID_ARGUMENT, "bind_leaves: null argument");
  int rc = build(kind, config, params, L);
  if (rc) return rc;
  JQ_REQUIRE(n_leaves == (int)L.leaves.size(), JQ_ERR_INVALID_ARGUMENT,
             "bind_leaves: the tree of this wavefunction has %zu leaves, %d were passed", L.leaves.size(), n_leaves);
  for (int i = 0; i < n_leaves; ++i) {
    const Leaf& l = L.leaves[i];
    if (leaf_elements && leaf_elements[i] != l.count) {
      std::string s = "params";
      for (auto& k : l.path) s += "/" + k;
      jq_set_error("bind_leaves: leaf %d (%s) has %lld elements, expected %lld", i, s.c_str(), (long long)leaf_elements[i],
                   l.count);
      return JQ_ERR_INVALID_ARGUMENT;
    }
    JQ_REQUIRE(leaves[i] != nullptr, JQ_ERR_INVALID_ARGUMENT, "bind_leaves: leaf %d is null", i);
  }
  for (int i = 0; i < n_leaves; ++i) *L.leaves[i].slot = leaves[i];
  return JQ_OK;
}
