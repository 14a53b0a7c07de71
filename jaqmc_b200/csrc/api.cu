// C-ABI entry points (include/jaqmc_b200.h): argument validation, walker tiling over the caller-owned
// workspace, and the glue kernels (local-energy finalisation, MH bookkeeping).
#include <cstdarg>

#include "wf.cuh"

thread_local long long jq_launch_counter = 0;
thread_local JqPrepCache jq_prep = {};
static thread_local char jq_err[512] = "";

void jq_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(jq_err, sizeof(jq_err), fmt, ap);
  va_end(ap);
}

static JqWrap make_wrap(const float* lattice) {
  JqWrap wr;
  memset(&wr, 0, sizeof(wr));
  if (!lattice) return wr;
  wr.on = 1;
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
  for (int i = 0; i < 9; ++i) {
    wr.lat[i] = (float)L[i];
    wr.inv[i] = (float)inv[i];
  }
  return wr;
}

// ------------------------------------------------------------------------------------------------
// per-kernel profiler (device build only)
// ------------------------------------------------------------------------------------------------
#include <map>
#include <string>
#include <vector>
thread_local int jq_prof_enabled = 0;
static thread_local double jq_prof_next_flops = 0.0, jq_prof_next_bytes = 0.0;
void jq_prof_work(double flops, double bytes) {
  jq_prof_next_flops = flops;
  jq_prof_next_bytes = bytes;
}
#ifndef JAQMC_HOST_EMU
struct JqProfRec {
  const char* name;
  cudaEvent_t e0, e1;
  double flops, bytes;
};
static thread_local std::vector<JqProfRec> jq_prof_recs;
void jq_prof_before(const char* name, cudaStream_t st) {
  JqProfRec r;
  r.name = name;
  r.flops = jq_prof_next_flops;
  r.bytes = jq_prof_next_bytes;
  jq_prof_next_flops = jq_prof_next_bytes = 0.0;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  jq_prof_recs.push_back(r);
}
void jq_prof_after(cudaStream_t st) { cudaEventRecord(jq_prof_recs.back().e1, st); }
#endif

extern "C" void jaqmc_b200_profile_enable(int on) { jq_prof_enabled = on; }

// Synchronises the recorded events and writes one line per kernel: "name launches total_ms flops bytes\n".
// Returns the number of bytes written (0 when nothing was recorded); clears the records.
extern "C" size_t jaqmc_b200_profile_fetch(char* buf, size_t cap) {
  size_t off = 0;
#ifndef JAQMC_HOST_EMU
  struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : jq_prof_recs) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    // launches of one kernel with different declared work are different shapes: keep them apart
    char key[160];
    snprintf(key, sizeof(key), "%s@%.4g", r.name, r.flops > 0 ? r.flops : r.bytes);
    Agg& a = agg[key];
    a.n += 1;
    a.ms += ms;
    a.flops += r.flops;
    a.bytes += r.bytes;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  jq_prof_recs.clear();
  for (auto& kv : agg) {
    int w = snprintf(buf + off, off < cap ? cap - off : 0, "%s %lld %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.n,
                     kv.second.ms, kv.second.flops, kv.second.bytes);
    if (w < 0 || off + (size_t)w >= cap) break;
    off += (size_t)w;
  }
#else
  (void)buf;
  (void)cap;
#endif
  return off;
}

#ifdef JAQMC_HOST_EMU
thread_local jq_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local unsigned char* jq_emu_dyn_smem = nullptr;
static thread_local std::vector<unsigned char> jq_emu_smem_store;
void jq_emu_set_smem(size_t bytes) {
  if (jq_emu_smem_store.size() < bytes + 64) jq_emu_smem_store.resize(bytes + 64);
  jq_emu_dyn_smem = jq_emu_smem_store.data();
}
#endif

// ------------------------------------------------------------------------------------------------
// per-kind dispatch
// ------------------------------------------------------------------------------------------------
static int wf_n_electrons(const jaqmc_wavefunction* wf) {
  switch (wf->kind) {
    case JAQMC_WF_FERMINET: {
      auto* c = (const jaqmc_ferminet_config*)wf->config;
      return c->n_up + c->n_dn;
    }
    case JAQMC_WF_LAPNET: {
      auto* c = (const jaqmc_lapnet_config*)wf->config;
      return c->n_up + c->n_dn;
    }
    case JAQMC_WF_PSIFORMER: {
      auto* c = (const jaqmc_psiformer_config*)wf->config;
      return c->n_up + c->n_dn;
    }
    case JAQMC_WF_SOLID_FERMINET: {
      auto* c = (const jaqmc_solid_config*)wf->config;
      return c->net.n_up + c->net.n_dn;
    }
    case JAQMC_WF_HYDROGEN:
      return ((const jaqmc_hydrogen_config*)wf->config)->n_electrons;
    default:
      return -1;
  }
}

static size_t wf_ws_bytes(const jaqmc_wavefunction* wf, long long W, int track) {
  switch (wf->kind) {
    case JAQMC_WF_FERMINET:
      return jq_ferminet_ws_bytes((const jaqmc_ferminet_config*)wf->config, W, track);
    case JAQMC_WF_LAPNET:
      return jq_lapnet_ws_bytes((const jaqmc_lapnet_config*)wf->config, W, track);
    case JAQMC_WF_PSIFORMER:
      return jq_psiformer_ws_bytes((const jaqmc_psiformer_config*)wf->config, W, track);
    case JAQMC_WF_SOLID_FERMINET:
      return jq_solid_ws_bytes((const jaqmc_solid_config*)wf->config, W, track);
    case JAQMC_WF_HYDROGEN:
      return 256;
    default:
      return 0;
  }
}

static int wf_forward_impl(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons, long long W,
                           int track, void* ws, size_t ws_bytes, JqWfOut out, cudaStream_t st);

// Weight-split cache scope (JqPrepCache): `begin` at the start of an API call that may run several forwards of the same
// network with the same walker-tile shape, `end` before returning.  The first forward fills, later ones reuse.
#define JQ_PREP_BYTES ((size_t)24 << 20)
static void prep_begin(void* base, size_t bytes) {
  jq_prep.collect = 0;
  jq_prep.base = (float*)base;
  jq_prep.cap = (long long)(bytes / sizeof(float));
  jq_prep.used = 0;
  jq_prep.n = jq_prep.cur = 0;
  jq_prep.mode = base ? 1 : 0;
}
static void prep_end() { jq_prep.mode = 0; }

static int wf_forward(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons, long long W,
                      int track, void* ws, size_t ws_bytes, JqWfOut out, cudaStream_t st) {
  jq_prep.cur = 0;
  const int rc = wf_forward_impl(wf, sys, electrons, W, track, ws, ws_bytes, out, st);
  if (jq_prep.mode == 1) jq_prep.mode = 2;   // the launch sequence is recorded: reuse from the next forward on
  return rc;
}

static int wf_forward_impl(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons, long long W,
                           int track, void* ws, size_t ws_bytes, JqWfOut out, cudaStream_t st) {
  switch (wf->kind) {
    case JAQMC_WF_FERMINET:
      return jq_ferminet_forward((const jaqmc_ferminet_config*)wf->config, (const jaqmc_ferminet_params*)wf->params,
                                 sys, electrons, W, track, ws, ws_bytes, out, st);
    case JAQMC_WF_LAPNET:
      return jq_lapnet_forward((const jaqmc_lapnet_config*)wf->config, (const jaqmc_lapnet_params*)wf->params, sys,
                               electrons, W, track, ws, ws_bytes, out, st);
    case JAQMC_WF_PSIFORMER:
      return jq_psiformer_forward((const jaqmc_psiformer_config*)wf->config, (const jaqmc_psiformer_params*)wf->params,
                                  sys, electrons, W, track, ws, ws_bytes, out, st);
    case JAQMC_WF_HYDROGEN:
      return jq_hydrogen_forward((const jaqmc_hydrogen_config*)wf->config, (const jaqmc_hydrogen_params*)wf->params,
                                 electrons, W, track, out, st);
    case JAQMC_WF_SOLID_FERMINET: {
      // value path only (sampling): real part -> logpsi, phase angle -> sign
      JQ_REQUIRE(track == 0, JQ_ERR_UNSUPPORTED, "solid: use jaqmc_b200_local_energy_complex for the tracked path");
      JqWfOutC oc = {out.logpsi, out.sign, nullptr, nullptr, nullptr, out.orbitals};
      return jq_solid_forward((const jaqmc_solid_config*)wf->config, (const jaqmc_solid_params*)wf->params, sys,
                              electrons, W, 0, ws, ws_bytes, oc, st);
    }
    default:
      jq_set_error("wavefunction kind %d is not implemented", wf->kind);
      return JQ_ERR_UNSUPPORTED;
  }
}

static int check_wf(const jaqmc_wavefunction* wf) {
  JQ_REQUIRE(wf && wf->config && wf->params, JQ_ERR_INVALID_ARGUMENT, "null wavefunction descriptor");
  JQ_REQUIRE(wf_n_electrons(wf) > 0, JQ_ERR_UNSUPPORTED, "wavefunction kind %d is not implemented", wf->kind);
  return JQ_OK;
}

// API-level scratch per walker (besides the pipeline's own): grad, lap, e_kin, e_pot, x2, lp2
static size_t api_scratch_bytes(int n, long long W) {
  JqArena ar(nullptr, 0);
  ar.take<float>(W * 3 * n);  // grad
  ar.take<float>(W);          // lap
  ar.take<float>(W);          // e_kin
  ar.take<float>(W);          // e_pot
  ar.take<float>(W);          // sign
  return ar.off;
}

// Largest walker tile whose pipeline workspace fits into `avail` bytes.
static long long fit_tile(const jaqmc_wavefunction* wf, long long W, int track, size_t avail) {
  if (wf_ws_bytes(wf, W, track) <= avail) return W;
  long long lo = 0, hi = W;  // invariant: lo fits (0 = nothing), hi does not
  while (hi - lo > 1) {
    long long mid = (lo + hi) / 2;
    if (wf_ws_bytes(wf, mid, track) <= avail) lo = mid; else hi = mid;
  }
  return lo;
}

extern "C" size_t jaqmc_b200_workspace_bytes(const jaqmc_wavefunction* wf, int64_t n_walkers, int track) {
  if (check_wf(wf) != JQ_OK || n_walkers < 0) return 0;
  int n = wf_n_electrons(wf);
  size_t mh = 0;
  {
    JqArena ar(nullptr, 0);
    ar.take<float>(n_walkers * 3 * n);  // x2
    ar.take<float>(n_walkers);          // lp2
    ar.take<float>(n_walkers);          // sign
    mh = ar.off;
  }
  size_t api = api_scratch_bytes(n, n_walkers);
  if (wf->kind == JAQMC_WF_SOLID_FERMINET) api = 2 * api + 4 * 256 + (size_t)n_walkers * 16;  // complex grad / lap / e_kin, logpsi planes
  return wf_ws_bytes(wf, n_walkers, track) + (api > mh ? api : mh) + JQ_PREP_BYTES + 1024;
}

extern "C" int jaqmc_b200_logpsi(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons,
                                 int64_t n_walkers, float* logpsi, float* sign, void* workspace,
                                 size_t workspace_bytes, jaqmc_stream_t stream) {
  int rc = check_wf(wf);
  if (rc) return rc;
  JQ_REQUIRE(n_walkers >= 0 && (n_walkers == 0 || (electrons && logpsi)), JQ_ERR_INVALID_ARGUMENT, "logpsi: null buffer");
  if (n_walkers == 0) return JQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = wf_n_electrons(wf);
  JqArena ar(workspace, workspace_bytes);
  float* sign_tmp = sign ? sign : ar.take<float>(n_walkers);
  JQ_REQUIRE(workspace && ar.off <= workspace_bytes, JQ_ERR_WORKSPACE_TOO_SMALL, "logpsi: workspace too small");
  size_t avail = workspace_bytes - ar.off;
  long long tile = fit_tile(wf, n_walkers, 0, avail);
  JQ_REQUIRE(tile >= 1, JQ_ERR_WORKSPACE_TOO_SMALL, "logpsi: workspace of %zu bytes cannot hold one walker", workspace_bytes);
  for (long long w0 = 0; w0 < n_walkers; w0 += tile) {
    long long wc = (n_walkers - w0 < tile) ? n_walkers - w0 : tile;
    JqWfOut out = {logpsi + w0, sign_tmp + w0, nullptr, nullptr, nullptr};
    rc = wf_forward(wf, sys, electrons + w0 * 3 * n, wc, 0, ar.base + ar.off, avail, out, st);
    if (rc) return rc;
  }
  return JQ_OK;
}

static int wf_ndets(const jaqmc_wavefunction* wf) {
  switch (wf->kind) {
    case JAQMC_WF_FERMINET: return ((const jaqmc_ferminet_config*)wf->config)->ndets;
    case JAQMC_WF_LAPNET: return ((const jaqmc_lapnet_config*)wf->config)->ndets;
    case JAQMC_WF_PSIFORMER: return ((const jaqmc_psiformer_config*)wf->config)->ndets;
    case JAQMC_WF_SOLID_FERMINET: return ((const jaqmc_solid_config*)wf->config)->net.ndets;
    default: return 0;
  }
}

extern "C" int jaqmc_b200_orbitals(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons,
                                   int64_t n_walkers, float* orbitals, void* workspace, size_t workspace_bytes,
                                   jaqmc_stream_t stream) {
  int rc = check_wf(wf);
  if (rc) return rc;
  const int D = wf_ndets(wf);
  JQ_REQUIRE(D > 0, JQ_ERR_UNSUPPORTED, "orbitals: wavefunction kind %d has no orbital matrices", wf->kind);
  JQ_REQUIRE(n_walkers >= 0 && (n_walkers == 0 || (electrons && orbitals)), JQ_ERR_INVALID_ARGUMENT, "orbitals: null buffer");
  if (n_walkers == 0) return JQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = wf_n_electrons(wf);
  const size_t per_walker = (size_t)D * n * n * (wf->kind == JAQMC_WF_SOLID_FERMINET ? 2 : 1);
  JqArena ar(workspace, workspace_bytes);
  float* lp = ar.take<float>(n_walkers);
  float* sg = ar.take<float>(n_walkers);
  JQ_REQUIRE(workspace && ar.off <= workspace_bytes, JQ_ERR_WORKSPACE_TOO_SMALL, "orbitals: workspace too small");
  size_t avail = workspace_bytes - ar.off;
  long long tile = fit_tile(wf, n_walkers, 0, avail);
  JQ_REQUIRE(tile >= 1, JQ_ERR_WORKSPACE_TOO_SMALL, "orbitals: workspace of %zu bytes cannot hold one walker", workspace_bytes);
  for (long long w0 = 0; w0 < n_walkers; w0 += tile) {
    long long wc = (n_walkers - w0 < tile) ? n_walkers - w0 : tile;
    JqWfOut out = {lp + w0, sg + w0, nullptr, nullptr, nullptr, orbitals + (size_t)w0 * per_walker};
    rc = wf_forward(wf, sys, electrons + w0 * 3 * n, wc, 0, ar.base + ar.off, avail, out, st);
    if (rc) return rc;
  }
  return JQ_OK;
}

// e_loc = e_kin + e_pot and the per-device partial sums {sum, sum of squares, finite count}.
#ifdef JAQMC_HOST_EMU
#define JQ_ATOMIC_ADD(p, v) (*(p) += (v))
#else
#define JQ_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#endif
#define FIN_TILE 256
__global__ void k_energy_finalize(const float* __restrict__ e_kin, const float* __restrict__ e_pot,
                                  float* __restrict__ e_loc, float* __restrict__ sums, long long W) {
  __shared__ float buf[FIN_TILE];
  const long long w0 = (long long)blockIdx.x * FIN_TILE;
  const int nw = (int)((W - w0 < FIN_TILE) ? W - w0 : FIN_TILE);
  for (int t = threadIdx.x; t < nw; t += blockDim.x) {
    float e = e_kin[w0 + t] + e_pot[w0 + t];
    if (e_loc) e_loc[w0 + t] = e;
    buf[t] = e;
  }
  __syncthreads();
  if (threadIdx.x == 0 && sums) {
    float s = 0.f, s2 = 0.f, c = 0.f;
    for (int t = 0; t < nw; ++t) {
      float e = buf[t];
      if (isfinite(e)) {
        s += e;
        s2 = fmaf(e, e, s2);
        c += 1.f;
      }
    }
    JQ_ATOMIC_ADD(sums + 0, s);
    JQ_ATOMIC_ADD(sums + 1, s2);
    JQ_ATOMIC_ADD(sums + 2, c);
  }
}

extern "C" int jaqmc_b200_coulomb(const jaqmc_system* sys, const float* electrons, int64_t n_walkers,
                                  int32_t n_electrons, float* e_pot, jaqmc_stream_t stream) {
  JQ_REQUIRE(sys && sys->atoms && sys->charges && sys->n_atoms >= 1, JQ_ERR_INVALID_ARGUMENT, "coulomb: null system");
  JQ_REQUIRE(n_walkers >= 0 && n_electrons >= 1 && (n_walkers == 0 || (electrons && e_pot)), JQ_ERR_INVALID_ARGUMENT,
             "coulomb: bad arguments");
  return jq_launch_coulomb(electrons, sys->atoms, sys->charges, (int)n_walkers, n_electrons, sys->n_atoms, e_pot,
                           (cudaStream_t)stream);
}

extern "C" int jaqmc_b200_ewald(const jaqmc_ewald* ewald, const float* electrons, int64_t n_walkers, int32_t n_electrons,
                                const float* atoms, const float* charges, int32_t n_atoms, float* e_pot,
                                jaqmc_stream_t stream) {
  JQ_REQUIRE(ewald != nullptr, JQ_ERR_INVALID_ARGUMENT, "ewald: null descriptor");
  JQ_REQUIRE(n_walkers >= 0 && n_electrons >= 0 && n_atoms >= 0 && n_electrons + n_atoms >= 1 &&
                 (n_walkers == 0 || ((electrons || n_electrons == 0) && e_pot && (n_atoms == 0 || (atoms && charges)))),
             JQ_ERR_INVALID_ARGUMENT, "ewald: bad arguments");
  return jq_launch_ewald(ewald, electrons, n_walkers, n_electrons, atoms, charges, n_atoms, e_pot, (cudaStream_t)stream);
}

extern "C" int jaqmc_b200_local_energy(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons,
                                       int64_t n_walkers, float* logpsi, float* sign, float* grad, float* lap,
                                       float* e_kin, float* e_pot, float* e_loc, float* sums, void* workspace,
                                       size_t workspace_bytes, jaqmc_stream_t stream) {
  int rc = check_wf(wf);
  if (rc) return rc;
  JQ_REQUIRE(n_walkers >= 0 && (n_walkers == 0 || (electrons && logpsi)), JQ_ERR_INVALID_ARGUMENT,
             "local_energy: null buffer");
  JQ_REQUIRE(sys && sys->atoms && sys->charges, JQ_ERR_INVALID_ARGUMENT, "local_energy: system needs atoms and charges");
  if (n_walkers == 0) return JQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = wf_n_electrons(wf);
  const long long W = n_walkers;
  JqArena ar(workspace, workspace_bytes);
  float* g = grad ? grad : ar.take<float>(W * 3 * n);
  float* lp = lap ? lap : ar.take<float>(W);
  float* ek = e_kin ? e_kin : ar.take<float>(W);
  float* ep = e_pot ? e_pot : ar.take<float>(W);
  float* sg = sign ? sign : ar.take<float>(W);
  JQ_REQUIRE(workspace && ar.off <= workspace_bytes, JQ_ERR_WORKSPACE_TOO_SMALL, "local_energy: workspace too small");
  // weight-split cache (one split kernel per evaluation for the pipelines that can list their dense launches; shared by
  // the walker tiles otherwise), when the workspace has room for it next to a full-batch pass
  struct PrepScope {
    ~PrepScope() { prep_end(); }
  } prep_scope;
  if (workspace_bytes >= ar.off + wf_ws_bytes(wf, W, 1) + JQ_PREP_BYTES + 256) {
    float* prep = ar.take<float>(JQ_PREP_BYTES / sizeof(float));
    prep_begin(prep, JQ_PREP_BYTES);
  } else {
    prep_begin(nullptr, 0);
  }
  size_t avail = workspace_bytes - ar.off;
  long long tile = fit_tile(wf, W, 1, avail);
  JQ_REQUIRE(tile >= 1, JQ_ERR_WORKSPACE_TOO_SMALL, "local_energy: workspace of %zu bytes cannot hold one walker",
             workspace_bytes);
  for (long long w0 = 0; w0 < W; w0 += tile) {
    long long wc = (W - w0 < tile) ? W - w0 : tile;
    JqWfOut out = {logpsi + w0, sg + w0, g + w0 * 3 * n, lp + w0, ek + w0};
    rc = wf_forward(wf, sys, electrons + w0 * 3 * n, wc, 1, ar.base + ar.off, avail, out, st);
    if (rc) return rc;
  }
  rc = jq_launch_coulomb(electrons, sys->atoms, sys->charges, (int)W, n, sys->n_atoms, ep, st);
  if (rc) return rc;
  if (e_loc || sums) {
    JQ_LAUNCH(k_energy_finalize, dim3(jq_cdiv(W, FIN_TILE)), dim3(FIN_TILE), 0, st, ek, ep, e_loc, sums, W);
    JQ_CHECK_LAUNCH();
  }
  return JQ_OK;
}

// e_loc = e_kin (complex) + e_pot and the partial sums of its real part
__global__ void k_energy_finalize_c(const float* __restrict__ e_kin, const float* __restrict__ e_pot,
                                    float* __restrict__ e_loc, float* __restrict__ sums, long long W) {
  __shared__ float buf[FIN_TILE];
  const long long w0 = (long long)blockIdx.x * FIN_TILE;
  const int nw = (int)((W - w0 < FIN_TILE) ? W - w0 : FIN_TILE);
  for (int t = threadIdx.x; t < nw; t += blockDim.x) {
    const float er = e_kin[2 * (w0 + t)] + (e_pot ? e_pot[w0 + t] : 0.f);
    if (e_loc) {
      e_loc[2 * (w0 + t)] = er;
      e_loc[2 * (w0 + t) + 1] = e_kin[2 * (w0 + t) + 1];
    }
    buf[t] = er;
  }
  __syncthreads();
  if (threadIdx.x == 0 && sums) {
    float s = 0.f, s2 = 0.f, c = 0.f;
    for (int t = 0; t < nw; ++t) {
      float e = buf[t];
      if (isfinite(e)) {
        s += e;
        s2 = fmaf(e, e, s2);
        c += 1.f;
      }
    }
    JQ_ATOMIC_ADD(sums + 0, s);
    JQ_ATOMIC_ADD(sums + 1, s2);
    JQ_ATOMIC_ADD(sums + 2, c);
  }
}

// interleave two planes into (re, im) pairs
__global__ void k_interleave2(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    o[2 * i] = a[i];
    o[2 * i + 1] = b[i];
  }
}

extern "C" int jaqmc_b200_local_energy_complex(const jaqmc_wavefunction* wf, const jaqmc_system* sys,
                                               const jaqmc_ewald* ewald, const float* cell_atoms,
                                               const float* cell_charges, int32_t n_cell_atoms, const float* electrons,
                                               int64_t n_walkers, float* logpsi, float* grad, float* lap, float* e_kin,
                                               float* e_pot, float* e_loc, float* sums, void* workspace,
                                               size_t workspace_bytes, jaqmc_stream_t stream) {
  JQ_REQUIRE(wf && wf->config && wf->params, JQ_ERR_INVALID_ARGUMENT, "null wavefunction descriptor");
  JQ_REQUIRE(wf->kind == JAQMC_WF_SOLID_FERMINET, JQ_ERR_UNSUPPORTED,
             "local_energy_complex: wavefunction kind %d has a real log psi; use jaqmc_b200_local_energy", wf->kind);
  JQ_REQUIRE(n_walkers >= 0 && (n_walkers == 0 || (electrons && logpsi)), JQ_ERR_INVALID_ARGUMENT,
             "local_energy_complex: null buffer");
  JQ_REQUIRE(!ewald || (cell_atoms && cell_charges && n_cell_atoms >= 1), JQ_ERR_INVALID_ARGUMENT,
             "local_energy_complex: the Ewald potential needs the simulation-cell atoms and charges");
  JQ_REQUIRE(!e_pot || ewald, JQ_ERR_INVALID_ARGUMENT, "local_energy_complex: e_pot requested without an Ewald descriptor");
  if (n_walkers == 0) return JQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = wf_n_electrons(wf);
  const long long W = n_walkers;
  JqArena ar(workspace, workspace_bytes);
  float* lp_re = ar.take<float>(W);
  float* lp_im = ar.take<float>(W);
  float* g = grad ? grad : ar.take<float>(W * 3 * n * 2);
  float* lpc = lap ? lap : ar.take<float>(W * 2);
  float* ek = e_kin ? e_kin : ar.take<float>(W * 2);
  float* ep = e_pot ? e_pot : (ewald ? ar.take<float>(W) : nullptr);
  JQ_REQUIRE(workspace && ar.off <= workspace_bytes, JQ_ERR_WORKSPACE_TOO_SMALL, "local_energy_complex: workspace too small");
  size_t avail = workspace_bytes - ar.off;
  long long tile = fit_tile(wf, W, 1, avail);
  JQ_REQUIRE(tile >= 1, JQ_ERR_WORKSPACE_TOO_SMALL,
             "local_energy_complex: workspace of %zu bytes cannot hold one walker", workspace_bytes);
  for (long long w0 = 0; w0 < W; w0 += tile) {
    long long wc = (W - w0 < tile) ? W - w0 : tile;
    JqWfOutC out = {lp_re + w0, lp_im + w0, g + w0 * 3 * n * 2, lpc + w0 * 2, ek + w0 * 2, nullptr};
    int rc = jq_solid_forward((const jaqmc_solid_config*)wf->config, (const jaqmc_solid_params*)wf->params, sys,
                              electrons + w0 * 3 * n, wc, 1, ar.base + ar.off, avail, out, st);
    if (rc) return rc;
  }
  JQ_LAUNCH(k_interleave2, dim3(jq_cdiv(W, 256)), dim3(256), 0, st, lp_re, lp_im, logpsi, W);
  JQ_CHECK_LAUNCH();
  if (ewald) {
    int rc = jq_launch_ewald(ewald, electrons, W, n, cell_atoms, cell_charges, n_cell_atoms, ep, st);
    if (rc) return rc;
  }
  if (e_loc || sums) {
    JQ_LAUNCH(k_energy_finalize_c, dim3(jq_cdiv(W, FIN_TILE)), dim3(FIN_TILE), 0, st, ek, ep, e_loc, sums, W);
    JQ_CHECK_LAUNCH();
  }
  return JQ_OK;
}

extern "C" int jaqmc_b200_layernorm_fl(const float* x, const float* scale, const float* bias, float* out, int64_t n_groups,
                                       int32_t n_components, int32_t n_features, float epsilon, int32_t kernel,
                                       jaqmc_stream_t stream) {
  JQ_REQUIRE(n_groups >= 0 && n_components >= 1 && n_features >= 1, JQ_ERR_INVALID_ARGUMENT, "layernorm_fl: bad sizes");
  JQ_REQUIRE(kernel >= 0 && kernel <= 3, JQ_ERR_INVALID_ARGUMENT, "layernorm_fl: kernel selector %d", kernel);
  JQ_REQUIRE(x && out, JQ_ERR_INVALID_ARGUMENT, "layernorm_fl: null operand");
  return jq_launch_layernorm_fl_sel(x, scale, bias, out, n_groups, n_components, n_features, epsilon, kernel,
                                    (cudaStream_t)stream);
}

extern "C" int jaqmc_b200_attention_fl(const float* q, const float* k, const float* v, float* out, int64_t n_walkers,
                                       int32_t n_electrons, int32_t n_heads, int32_t head_dim, int32_t q_components,
                                       int32_t k_components, int32_t kernel, jaqmc_stream_t stream) {
  JQ_REQUIRE(n_walkers >= 0 && n_electrons >= 1 && n_heads >= 1 && head_dim >= 1, JQ_ERR_INVALID_ARGUMENT,
             "attention_fl: bad sizes");
  JQ_REQUIRE(kernel >= 0 && kernel <= 5, JQ_ERR_INVALID_ARGUMENT, "attention_fl: kernel selector %d", kernel);
  JQ_REQUIRE(q && k && v && out, JQ_ERR_INVALID_ARGUMENT, "attention_fl: null operand");
  const int F = n_heads * head_dim, Cd = 3 * n_electrons + 2;
  JqAttnOperand qo{q, q_components, F, 0}, ko{k, k_components, F, 0}, vo{v, Cd, F, 0};
  return jq_launch_attention_fl_sel(qo, ko, vo, out, F, n_walkers, n_electrons, n_heads, head_dim, 1, kernel,
                                    (cudaStream_t)stream);
}

extern "C" int jaqmc_b200_dense_fl(const float* x, const float* x2, const float* kernel, const float* kernel2,
                                   const float* bias, const float* addend, const float* residual, float* out,
                                   int64_t n_groups, int32_t n_components, int32_t k0, int32_t k1, int32_t n_out,
                                   int32_t groups_per_walker, int32_t activation, int32_t residual_mode,
                                   int32_t use_tensor_cores, void* workspace, size_t workspace_bytes,
                                   jaqmc_stream_t stream) {
  JQ_REQUIRE(n_groups >= 0 && n_components >= 1 && k0 >= 1 && k1 >= 0 && n_out >= 1 && groups_per_walker >= 1,
             JQ_ERR_INVALID_ARGUMENT, "dense_fl: bad sizes");
  JQ_REQUIRE(n_groups % groups_per_walker == 0, JQ_ERR_INVALID_ARGUMENT, "dense_fl: n_groups %% groups_per_walker != 0");
  JqDenseArgs a;
  memset(&a, 0, sizeof(a));
  a.src0 = x;
  a.k0 = k0;
  a.src1 = x2;
  a.k1 = k1;
  a.w0 = kernel;
  a.w1 = kernel2;
  a.bias = bias;
  a.cadd = addend;
  a.res = residual;
  a.out = out;
  a.N = n_out;
  a.C = n_components;
  a.n_sub = a.n_tot = groups_per_walker;
  a.G = n_groups;
  a.act = activation;
  a.res_mode = residual_mode;
  if (use_tensor_cores) {
    size_t need = jq_dense_tc_scratch_floats(k0 + k1, n_out) * sizeof(float);
    JQ_REQUIRE(workspace && workspace_bytes >= need, JQ_ERR_WORKSPACE_TOO_SMALL, "dense_fl: workspace %zu < %zu bytes",
               workspace_bytes, need);
    a.wscratch = (float*)workspace;
    a.tc_mode = (use_tensor_cores == 2) ? 1 : 0;
    a.tc_force = 1;   // the caller asked for the tensor-core kernel: no size heuristic
  }
  return jq_launch_dense(a, (cudaStream_t)stream);
}

extern "C" int jaqmc_b200_mh_propose(const float* x1, const float* normals, const float* stddev, float* x2,
                                     int64_t count, jaqmc_stream_t stream) {
  JQ_REQUIRE(count >= 0 && (count == 0 || (x1 && normals && stddev && x2)), JQ_ERR_INVALID_ARGUMENT, "mh_propose: null buffer");
  return jq_launch_mh_propose(x1, normals, stddev, x2, count, make_wrap(nullptr), (cudaStream_t)stream);
}

extern "C" int jaqmc_b200_mh_accept(float* x1, const float* x2, float* logprob1, const float* logprob2,
                                    const float* uniforms, int64_t n_walkers, int32_t row, float* n_accept,
                                    uint8_t* accepted, jaqmc_stream_t stream) {
  JQ_REQUIRE(n_walkers >= 0 && row >= 1 && (n_walkers == 0 || (x1 && x2 && logprob1 && logprob2 && uniforms && n_accept)),
             JQ_ERR_INVALID_ARGUMENT, "mh_accept: null buffer");
  return jq_launch_mh_accept(x1, x2, logprob1, logprob2, uniforms, nullptr, nullptr, nullptr, (int)n_walkers, row, 1.0f,
                             n_accept, accepted, make_wrap(nullptr), (cudaStream_t)stream);
}

static int mh_step_impl(const jaqmc_wavefunction* wf, const jaqmc_system* sys, float* electrons, float* logpsi,
                        int32_t logpsi_valid, const float* normals, const float* uniforms, const float* stddev,
                        int32_t n_steps, int64_t n_walkers, float* n_accept, uint8_t* accepted, const JqWrap& wrap,
                        void* workspace, size_t workspace_bytes, jaqmc_stream_t stream);

extern "C" int jaqmc_b200_mh_step(const jaqmc_wavefunction* wf, const jaqmc_system* sys, float* electrons,
                                  float* logpsi, int32_t logpsi_valid, const float* normals, const float* uniforms,
                                  const float* stddev, int32_t n_steps, int64_t n_walkers, float* n_accept,
                                  uint8_t* accepted, void* workspace, size_t workspace_bytes, jaqmc_stream_t stream) {
  return mh_step_impl(wf, sys, electrons, logpsi, logpsi_valid, normals, uniforms, stddev, n_steps, n_walkers, n_accept,
                      accepted, make_wrap(nullptr), workspace, workspace_bytes, stream);
}

extern "C" int jaqmc_b200_mh_step_pbc(const jaqmc_wavefunction* wf, const jaqmc_system* sys, float* electrons,
                                      float* logpsi, int32_t logpsi_valid, const float* normals, const float* uniforms,
                                      const float* stddev, int32_t n_steps, int64_t n_walkers, float* n_accept,
                                      uint8_t* accepted, const float* lattice, void* workspace, size_t workspace_bytes,
                                      jaqmc_stream_t stream) {
  JQ_REQUIRE(lattice != nullptr, JQ_ERR_INVALID_ARGUMENT, "mh_step_pbc: null lattice");
  return mh_step_impl(wf, sys, electrons, logpsi, logpsi_valid, normals, uniforms, stddev, n_steps, n_walkers, n_accept,
                      accepted, make_wrap(lattice), workspace, workspace_bytes, stream);
}

static int mh_step_impl(const jaqmc_wavefunction* wf, const jaqmc_system* sys, float* electrons, float* logpsi,
                        int32_t logpsi_valid, const float* normals, const float* uniforms, const float* stddev,
                        int32_t n_steps, int64_t n_walkers, float* n_accept, uint8_t* accepted, const JqWrap& wrap,
                        void* workspace, size_t workspace_bytes, jaqmc_stream_t stream) {
  int rc = check_wf(wf);
  if (rc) return rc;
  JQ_REQUIRE(n_walkers >= 0 && n_steps >= 0, JQ_ERR_INVALID_ARGUMENT, "mh_step: negative size");
  if (n_walkers == 0) return JQ_OK;
  JQ_REQUIRE(electrons && logpsi && stddev && n_accept && (n_steps == 0 || (normals && uniforms)),
             JQ_ERR_INVALID_ARGUMENT, "mh_step: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = wf_n_electrons(wf);
  const long long W = n_walkers;
  const int row = 3 * n;
  JqArena ar(workspace, workspace_bytes);
  float* x2 = ar.take<float>(W * row);
  float* lp2 = ar.take<float>(W);
  float* sg = ar.take<float>(W);
  JQ_REQUIRE(workspace && ar.off <= workspace_bytes, JQ_ERR_WORKSPACE_TOO_SMALL, "mh_step: workspace too small");
  // weight-split cache for the n_steps + 1 forwards, when the workspace has room for it next to a full-batch pass
  struct PrepScope {
    ~PrepScope() { prep_end(); }
  } prep_scope;
  if (workspace_bytes >= ar.off + wf_ws_bytes(wf, W, 0) + JQ_PREP_BYTES + 256) {
    float* prep = ar.take<float>(JQ_PREP_BYTES / sizeof(float));
    prep_begin(prep, JQ_PREP_BYTES);
  } else {
    prep_begin(nullptr, 0);
  }
  size_t avail = workspace_bytes - ar.off;
  long long tile = fit_tile(wf, W, 0, avail);
  JQ_REQUIRE(tile >= 1, JQ_ERR_WORKSPACE_TOO_SMALL, "mh_step: workspace of %zu bytes cannot hold one walker", workspace_bytes);
  auto forward_all = [&](const float* x, float* lp) -> int {
    for (long long w0 = 0; w0 < W; w0 += tile) {
      long long wc = (W - w0 < tile) ? W - w0 : tile;
      JqWfOut out = {lp + w0, sg + w0, nullptr, nullptr, nullptr};
      int r = wf_forward(wf, sys, x + w0 * row, wc, 0, ar.base + ar.off, avail, out, st);
      if (r) return r;
    }
    return JQ_OK;
  };
  if (!logpsi_valid && (rc = forward_all(electrons, logpsi))) return rc;
  if (n_steps == 0) return JQ_OK;
  if ((rc = jq_launch_mh_propose(electrons, normals, stddev, x2, W * row, wrap, st))) return rc;
  for (int s = 0; s < n_steps; ++s) {
    if ((rc = forward_all(x2, lp2))) return rc;
    const float* next = (s + 1 < n_steps) ? normals + (size_t)(s + 1) * W * row : nullptr;
    // fused: accept test, select, and the next proposal (x2 is rewritten in place)
    rc = jq_launch_mh_accept(electrons, x2, logpsi, lp2, uniforms + (size_t)s * W, next, stddev, x2, (int)W, row, 2.0f,
                             n_accept, accepted ? accepted + (size_t)s * W : nullptr, wrap, st);
    if (rc) return rc;
  }
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// psi-ratio consumers: the ECP non-local integral (estimator/ecp/nonlocal_integral.py:23-165) and
// SpinSquared (estimator/spin.py:75-146) evaluate phase_logpsi at many configurations per walker that differ from the
// walker's in one or two electrons.  One call: build the moved configurations tile by tile, run the value-only forward
// pass (the sampling path's kernels), return log|psi'/psi| and the sign (phase) of the ratio.
// ------------------------------------------------------------------------------------------------
// cfg[t][e][:] = electrons[w][e][:] with electron idx[q][s] replaced by pos[w][q][s][:]  (t = local index of (w, q))
__global__ void k_build_moved(const float* __restrict__ el, const int* __restrict__ idx, const float* __restrict__ pos,
                              long long t0, long long count, int Q, int n, float* __restrict__ cfg) {
  const long long items = count * n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(it % n);
    const long long t = it / n;
    const long long wq = t0 + t;
    const long long w = wq / Q;
    const int q = (int)(wq - w * Q);
    const float* src = el + (w * n + e) * 3;
    for (int s = 0; s < 2; ++s)
      if (idx[2 * q + s] == e) src = pos + ((w * Q + q) * 2 + s) * 3;   // the later entry wins if both name e
    float* o = cfg + it * 3;
    o[0] = src[0];
    o[1] = src[1];
    o[2] = src[2];
  }
}

// complex_phase: sign arrays hold phase angles (periodic network): difference wrapped into (-pi, pi]; else signs: product
__global__ void k_ratio_finish(const float* __restrict__ lp0, const float* __restrict__ sg0, float* __restrict__ lp,
                               float* __restrict__ sg, long long t0, long long count, int Q, int complex_phase) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (long long)gridDim.x * blockDim.x) {
    const long long w = (t0 + t) / Q;
    lp[t] = lp[t] - lp0[w];
    if (complex_phase) {
      float dph = sg[t] - sg0[w];
      const float two_pi = 6.28318530717958647692f;
      dph -= two_pi * rintf(dph / two_pi);
      sg[t] = dph;
    } else {
      sg[t] = sg[t] * sg0[w];
    }
  }
}

extern "C" int jaqmc_b200_psi_ratios(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons,
                                     int64_t n_walkers, int32_t n_moves, const int32_t* move_index, const float* move_pos,
                                     float* log_ratio, float* sign_ratio, void* workspace, size_t workspace_bytes,
                                     jaqmc_stream_t stream) {
  int rc = check_wf(wf);
  if (rc) return rc;
  JQ_REQUIRE(n_walkers >= 0 && n_moves >= 0, JQ_ERR_INVALID_ARGUMENT, "psi_ratios: negative size");
  if (n_walkers == 0 || n_moves == 0) return JQ_OK;
  JQ_REQUIRE(electrons && move_index && move_pos && log_ratio && sign_ratio, JQ_ERR_INVALID_ARGUMENT, "psi_ratios: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = wf_n_electrons(wf);
  const long long W = n_walkers, Q = n_moves, T = W * Q;
  JqArena ar(workspace, workspace_bytes);
  float* lp0 = ar.take<float>(W);
  float* sg0 = ar.take<float>(W);
  JQ_REQUIRE(workspace && ar.off <= workspace_bytes, JQ_ERR_WORKSPACE_TOO_SMALL, "psi_ratios: workspace too small");
  struct PrepScope {
    ~PrepScope() { prep_end(); }
  } prep_scope;
  if (workspace_bytes >= ar.off + JQ_PREP_BYTES + 64 * wf_ws_bytes(wf, 1, 0) + (size_t)64 * 3 * n * 4 + 1024) {
    float* prep = ar.take<float>(JQ_PREP_BYTES / sizeof(float));
    prep_begin(prep, JQ_PREP_BYTES);
  } else {
    prep_begin(nullptr, 0);
  }
  // largest tile of configurations: per configuration 3n floats of positions + the pipeline's value-path workspace
  const size_t fixed = ar.off;
  auto fits = [&](long long tile) {
    JqArena probe(nullptr, 0);
    probe.take<float>(tile * 3 * n);
    return fixed + probe.off + wf_ws_bytes(wf, tile, 0) + 256 <= workspace_bytes;
  };
  long long tile = T;
  if (!fits(tile)) {
    long long lo = 0, hi = T;
    while (hi - lo > 1) {
      const long long mid = (lo + hi) / 2;
      if (fits(mid)) lo = mid; else hi = mid;
    }
    tile = lo;
  }
  JQ_REQUIRE(tile >= 1, JQ_ERR_WORKSPACE_TOO_SMALL, "psi_ratios: workspace of %zu bytes cannot hold one configuration", workspace_bytes);
  float* cfg = ar.take<float>(tile * 3 * n);
  void* wsn = ar.base ? ar.base + ar.off : nullptr;
  const size_t avail = workspace_bytes - ar.off;
  // reference values
  {
    const long long wt = (tile < W) ? tile : W;
    for (long long w0 = 0; w0 < W; w0 += wt) {
      const long long wc = (W - w0 < wt) ? W - w0 : wt;
      JqWfOut out = {lp0 + w0, sg0 + w0, nullptr, nullptr, nullptr, nullptr};
      if ((rc = wf_forward(wf, sys, electrons + w0 * 3 * n, wc, 0, wsn, avail, out, st))) return rc;
    }
  }
  const int complex_phase = wf->kind == JAQMC_WF_SOLID_FERMINET;
  for (long long t0 = 0; t0 < T; t0 += tile) {
    const long long tc = (T - t0 < tile) ? T - t0 : tile;
    int grid = jq_cdiv(tc * n, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    JQ_LAUNCH(k_build_moved, dim3(grid), dim3(256), 0, st, electrons, move_index, move_pos, t0, tc, (int)Q, n, cfg);
    JQ_CHECK_LAUNCH();
    JqWfOut out = {log_ratio + t0, sign_ratio + t0, nullptr, nullptr, nullptr, nullptr};
    if ((rc = wf_forward(wf, sys, cfg, tc, 0, wsn, avail, out, st))) return rc;
    JQ_LAUNCH(k_ratio_finish, dim3(jq_cdiv(tc, 256)), dim3(256), 0, st, lp0, sg0, log_ratio + t0, sign_ratio + t0, t0, tc,
              (int)Q, complex_phase);
    JQ_CHECK_LAUNCH();
  }
  return JQ_OK;
}

extern "C" int64_t jaqmc_b200_launch_count(void) { return jq_launch_counter; }
extern "C" void jaqmc_b200_reset_launch_count(void) { jq_launch_counter = 0; }
extern "C" const char* jaqmc_b200_last_error(void) { return jq_err; }
extern "C" const char* jaqmc_b200_version(void) {
#ifdef JAQMC_HOST_EMU
  return "jaqmc_b200 0.1 (HOST EMULATION BUILD - test infrastructure only)";
#else
  return "jaqmc_b200 0.1 (sm_100a)";
#endif
}
