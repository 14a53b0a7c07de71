// Internal interface between the C-ABI dispatcher (api.cu) and the per-network pipelines.
#pragma once
#include "../../include/jaqmc_b200.h"
#include "aug.cuh"

struct JqWfOut {
  float* logpsi;  // [W]
  float* sign;    // [W]
  float* grad;    // [W][3n]   (track only)
  float* lap;     // [W]       (track only)
  float* e_kin;   // [W]       (track only)
  float* orbitals;  // [W][D][n][n] or null: value path only -- stop after orbitals x envelope and emit the matrices
};
// out[w][d][e][o] = in[w][e][d*n + o]  (+ imaginary plane interleaved when in_im != null): the layout of wf.orbitals
int jq_launch_orbitals_out(const float* in_re, const float* in_im, float* out, long long W, int n, int D, cudaStream_t st);

size_t jq_ferminet_ws_bytes(const jaqmc_ferminet_config* c, long long W, int track);
int jq_ferminet_forward(const jaqmc_ferminet_config* c, const jaqmc_ferminet_params* p, const jaqmc_system* sys,
                        const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOut out,
                        cudaStream_t st);

// ---- shared output head (head.cu) ---------------------------------------------------------------
struct JqHeadDims {
  JqSpins sp;
  int A, D, C, hidden;
  int hidden_valid;   // 0 (= hidden) or the number of orbital-kernel rows: input columns [hidden_valid, hidden) are zero padding
  int envelope_type, split, jastrow;
};
struct JqHeadBufs {
  float *orb, *det_sign, *det_logabs, *det_grad, *det_lap, *extra;
};
void jq_head_carve(const JqHeadDims& d, long long W, JqArena& ar, JqHeadBufs* b);
// h [W][n][C][hidden] -> logpsi, sign (+ grad, lap, e_kin when C > 1)
int jq_head_forward(const JqHeadDims& d, const jaqmc_head_params* p, const float* h, const float* electrons,
                    const float* atoms, long long W, const JqHeadBufs& b, float* wscratch, JqWfOut out,
                    cudaStream_t st);
int jq_launch_jastrow(const float* electrons, const float* alpha_par, const float* alpha_anti, int W, JqSpins sp,
                      int track, float* extra, cudaStream_t st);

// ---- LapNet / Psiformer pipelines (attnnets.cu) ---------------------------------------------------
size_t jq_lapnet_ws_bytes(const jaqmc_lapnet_config* c, long long W, int track);
int jq_lapnet_forward(const jaqmc_lapnet_config* c, const jaqmc_lapnet_params* p, const jaqmc_system* sys,
                      const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOut out,
                      cudaStream_t st);
size_t jq_psiformer_ws_bytes(const jaqmc_psiformer_config* c, long long W, int track);
int jq_psiformer_forward(const jaqmc_psiformer_config* c, const jaqmc_psiformer_params* p, const jaqmc_system* sys,
                         const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOut out,
                         cudaStream_t st);

// ---- Ewald (ewald.cu) ---------------------------------------------------------------------------
int jq_launch_ewald(const jaqmc_ewald* ew, const float* electrons, long long W, int n_el, const float* atoms,
                    const float* charges, int n_at, float* e_pot, cudaStream_t st);

// ---- FermiNet backbone shared by the molecular and periodic networks (ferminet.cu) -----------------
struct FermiDims {
  JqSpins sp;
  int n, A, D, L, nch, C, C1, C2, f1, fee;  // f1 = input features per electron, fee = per pair
  int d1[JQ_MAX_LAYERS], d2[JQ_MAX_LAYERS];  // widths after layer l
  int d1max, d2max, in1, in1p;
  int use_last, agg_w, agg_wp;   // use_last_layer: aggregated feature width fed to the orbitals (and its padding to 32)
};

struct FermiBufs {
  float *ae, *h2a, *h2b, *g2, *x1, *ha, *hb, *m, *cadd, *wscr, *agg;
  JqHeadBufs head;
};

int jq_fermi_dims(const jaqmc_ferminet_config* c, int track, int fat, int fee, FermiDims* o);
void jq_fermi_carve_backbone(const FermiDims& d, long long W, JqArena& ar, FermiBufs* b);
int jq_fermi_backbone(const FermiDims& d, const jaqmc_ferminet_params* p, long long W, int track, const FermiBufs& b,
                      cudaStream_t st, float** h_out);
int jq_launch_solid_features(const float* electrons, const float* prim_atoms, const float* sim_lattice,
                             const float* prim_lattice, int distance_type, int sym_type, int W, int n, int A, int track,
                             float* ae, float* r_ae, float* ee, cudaStream_t st);

// ---- periodic FermiNet (solid.cu) ------------------------------------------------------------------
struct JqWfOutC {
  float* logpsi_re;  // [W]
  float* logpsi_im;  // [W]
  float* grad;       // [W][3n][2]  (track only)
  float* lap;        // [W][2]
  float* e_kin;      // [W][2]
  float* orbitals;   // [W][D][n][n][2] or null (value path only)
};
size_t jq_solid_ws_bytes(const jaqmc_solid_config* c, long long W, int track);
int jq_solid_forward(const jaqmc_solid_config* c, const jaqmc_solid_params* p, const jaqmc_system* sys,
                     const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOutC out,
                     cudaStream_t st);

int jq_hydrogen_forward(const jaqmc_hydrogen_config* c, const jaqmc_hydrogen_params* p, const float* electrons,
                        long long W, int track, JqWfOut out, cudaStream_t st);
