// Augmented ("forward-Laplacian") tensor conventions shared by all kernels.
//
// A tracked activation is stored as  T[group][C][F]  (row-major, F contiguous) where
//   group  = (walker, electron)  for single-electron streams,   [W][n]
//          = (walker, i, j)      for pair streams,              [W][n*n], pair p = i*n + j
//   C      = number of components carried per group:
//              C == 1        value only (sampling / psi-ratio path),
//              C == ncoord+2 value, ncoord Jacobian columns, Laplacian      (c = 0, 1..ncoord, ncoord+1)
//            with ncoord = 3 for "Local1" tensors (each row depends on its own electron only,
//            reference laplacian/sparse.py Local1Jacobian), 6 for "Local2" pair tensors (columns
//            0-2 w.r.t. electron i, 3-5 w.r.t. electron j, Local2Jacobian) and 3n for dense tensors.
// One GEMM over the rows {x, J_1..J_K, L} is the forward-Laplacian rule of an affine map (reference
// laplacian/primitives/dot_general.py:377-407); the elementwise rule f'(x)J, f'(x)L + f''(x) sum J^2
// (primitives/elementwise.py:42-72) needs all C rows of one group, which is why groups are contiguous.
#pragma once
#include "common.cuh"

#define JQ_MAX_LAYERS 8
#define JQ_MAX_ATOMS 64

// Non-empty spin channels (reference utils/array.py:24-45 split_nonempty_channels).
struct JqSpins {
  int n_up, n_dn;
  __host__ __device__ int n() const { return n_up + n_dn; }
  __host__ __device__ int nch() const { return (n_up > 0 && n_dn > 0) ? 2 : 1; }
  // channel s covers electrons [lo, hi)
  __host__ __device__ int lo(int s) const { return (nch() == 2 && s == 1) ? n_up : 0; }
  __host__ __device__ int hi(int s) const { return (nch() == 2 && s == 0) ? n_up : n_up + n_dn; }
  __host__ __device__ int chan_of(int e) const { return (nch() == 2 && e >= n_up) ? 1 : 0; }
};

// ---- kernels implemented in features.cu ---------------------------------------------------------
// ae [W][n][C1][4A (+1 with spin_column)] Local1, ee [W][n*n][C2][4] Local2 (skipped when ee == nullptr)
int jq_launch_mol_features(const float* electrons, const float* atoms, int W, JqSpins sp, int A, int rescale,
                           int track, int spin_column, float* ae, float* ee, cudaStream_t st);
int jq_launch_coulomb(const float* electrons, const float* atoms, const float* charges, int W, int n, int A,
                      float* e_pot, cudaStream_t st);

// ---- kernels implemented in dense.cu ------------------------------------------------------------
struct JqEnvFuse {
  const float* electrons;  // [W][n][3]
  const float* atoms;      // [A][3]
  const float* pi;         // [n_orb][A][D] of the launch's spin channel
  const float* sigma;
  int A, D, n, type;       // type 0 isotropic, 1 abs_isotropic
};
struct JqDenseArgs {
  const float* src0;  // [groups_total][C][k0]
  int k0;
  const float* src1;  // optional second source concatenated along the contraction axis, [..][C][k1]
  int k1;
  const float* w0;    // [k0][ldw]  rows of the flax kernel that multiply src0 (columns [0, N) of each row)
  const float* w1;    // [k1][ldw]
  int ldw;            // row stride of w0 / w1 (0 -> N); lets a launch take a column block of a wider kernel
  const float* bias;  // [N] or null; added to the value row only
  const float* cadd;  // [W][C][N] or null; per-walker addend broadcast over the walker's groups
  float* out;         // [groups_total][C][N]
  int N, C;
  // group mapping: launch covers G = W*n_sub groups; sub-group g' -> group (g'/n_sub)*n_tot + j0 + g'%n_sub
  int n_sub, n_tot, j0;
  long long G;
  // fused epilogue: act 0 = none, 1 = tanh with the forward-Laplacian rule; then the optional residual
  // res_mode 0 none, 1 (res + y)/sqrt(2) (FermiNet), 2 res + y; res has the layout of out
  int act, res_mode;
  const float* res;
  // act 2 (tensor-core path only): the output is multiplied by the orbital envelope with the product rule
  // (wavefunction/output/envelope.py:98-140): features are (determinant, orbital) pairs, groups are electrons
  const struct JqEnvFuse* env;
  // scratch for the tensor-core path's transposed hi/lo weight split: jq_dense_tc_scratch_floats(k0+k1, N) floats,
  // or null to force the CUDA-core kernel
  float* wscratch;
  int k0_valid; // 0 (= k0), or the number of kernel rows that exist: src0 columns [k0_valid, k0) are zero padding
                // (lets a 20- or 56-wide first layer take the tensor-core path, which needs multiples of 32)
  int small_gt; // set by the launcher: groups per block of k_dense_small
  int tc_force; // 1: take the tensor-core path whenever the shape allows, however small the launch (tests, dense_fl)
  int tc_mode;  // 0: CTA-pair kernel (weights resident) when the shape allows, else the streaming one; 1: streaming only
};
int jq_launch_dense(const JqDenseArgs& a, cudaStream_t st);

// Per-API-call cache of the tensor-core path's hi/lo weight splits.  The weights are constant within one call, but a
// call may run the same network many times (jaqmc_b200_mh_step: 11 value-only forwards, psi_ratios and tiled calls:
// one per tile): the first forward fills the cache launch by launch (mode 1), later forwards walk the same launch
// sequence and reuse the slots (mode 2) instead of re-running k_weight_split_t -- 9 of the ~32 launches of a
// FermiNet forward on the launch-latency-bound sampling path.  Thread-local: set and cleared by api.cu around a call.
// A pipeline that knows its dense launches up front (FermiNet) can do better still: a COLLECT pass walks the launch
// sequence without launching anything (jq_launch_dense only records the tensor-core launches' weights), one
// multi-segment kernel splits them all (jq_prep_flush), and the real pass finds every split by key -- one launch
// instead of nine per local-energy evaluation, which matters for the 512-walker shard of the 8-GPU run.
struct JqPrepSlot {
  const float* w0;
  const float* w1;
  int k0, k1, N, k0_valid, ldw;
  long long off;   // float offset into base, or -1: no room, this launch splits into its own scratch
  int pending;     // recorded by the collect pass, not split yet
};
struct JqPrepCache {
  float* base;
  long long cap, used;   // floats
  int mode;              // 0 off, 1 fill (launch by launch), 2 reuse (lookup by key)
  int collect;           // 1: dry pass -- launchers record / skip instead of launching
  int n, cur;
  JqPrepSlot slot[96];
};
extern thread_local JqPrepCache jq_prep;
int jq_prep_flush(cudaStream_t st);   // splits every pending slot with one kernel (device build; no-op in the emulation)
bool jq_dense_tc_eligible(const JqDenseArgs& a);  // device build: will this launch take the tcgen05 kernel?
size_t jq_dense_tc_scratch_floats(int k_total, int n_out);
// out = act(y) or (res + act(y))/sqrt(2); act = tanh with the forward-Laplacian rule.  In-place allowed.
int jq_launch_tanh_fl(const float* y, const float* res, float* out, long long G, int C, int F, int residual_mode,
                      cudaStream_t st);
int jq_launch_pair_mean(const float* h2, float* g2, int W, JqSpins sp, int d2, int track, cudaStream_t st);
#ifndef JAQMC_HOST_EMU
// one two-electron layer fused with the spin-channel means of its output (32 output features); returns
// false when the shape is not covered (the caller then takes jq_launch_dense + jq_launch_pair_mean)
bool jq_launch_pair_layer_fused(const float* h2, int K, const float* w0, const float* bias, float* h2n, float* g2, int W,
                                JqSpins sp, int N, int residual, int track, cudaStream_t st, int* rc);
#endif
int jq_launch_concat_layer1(const float* ae, const float* g2, float* out, int W, JqSpins sp, int f1, int fg,
                            int track, int ld_out, cudaStream_t st);   // ld_out >= f1 (1 + nch) + fg: zero padded
int jq_launch_spin_mean(const float* h, float* m, int W, JqSpins sp, int C, int F, cudaStream_t st);

// ---- kernels implemented in logdet.cu -----------------------------------------------------------
struct JqEnvelopeArgs {
  const float* pi[2];     // per spin channel: [n_orb][A][D]   (reference output/envelope.py:131-135)
  const float* sigma[2];
  int type;               // 0 isotropic, 1 abs_isotropic, 2 null
};
int jq_launch_orb_envelope(float* orb, const float* electrons, const float* atoms, const JqEnvelopeArgs& env,
                           int W, JqSpins sp, int A, int D, int track, cudaStream_t st);
int jq_launch_logdet(const float* orb, int W, int n, int D, int track, float* det_sign, float* det_logabs,
                     float* det_grad, float* det_lap, cudaStream_t st);
// value path, n <= 32: envelope (isotropic / abs_isotropic) applied while the rows are loaded, then the LU
bool jq_logdet_value_env_eligible(int n, int A, int env_type);
int jq_launch_logdet_value_env(const float* orb, const float* electrons, const float* atoms, const JqEnvelopeArgs& env,
                               int has_env, int W, JqSpins sp, int A, int D, float* det_sign, float* det_logabs,
                               cudaStream_t st);
int jq_launch_logdet_combine(const float* det_sign, const float* det_logabs, const float* det_grad,
                             const float* det_lap, int W, int n, int D, int track, const float* extra_logpsi,
                             float* logpsi, float* sign, float* grad, float* lap, float* e_kin, cudaStream_t st);

// ---- kernels implemented in mcmc.cu ---------------------------------------------------------------
struct JqWrap {   // periodic wrap of proposals into a cell: lattice rows and inverse (on == 0: open boundaries)
  int on;
  float lat[9], inv[9];
};
int jq_launch_mh_propose(const float* x1, const float* normals, const float* stddev, float* x2, long long count,
                         const JqWrap& wr, cudaStream_t st);
int jq_launch_mh_accept(float* x1, const float* x2, float* lp1, const float* lp2, const float* log_u,
                        const float* next_normals, const float* stddev, float* x2_next, int W, int row,
                        float scale, float* n_accept, unsigned char* accepted, const JqWrap& wr, cudaStream_t st);

// ---- kernels implemented in attention.cu ----------------------------------------------------------
struct JqAttnOperand {
  const float* p;
  int C;    // components stored: 3n+2 (dense), 5 (Local1) or 1 (value only)
  int ld;   // row length
  int off;  // first column of head 0
};
int jq_launch_densify_local1(const float* in, float* out, long long W, int n, int F, cudaStream_t st);
int jq_launch_layernorm_fl(const float* x, const float* scale, const float* bias, float* out, long long G, int C, int F,
                           float eps, cudaStream_t st);
// force: 0 library's choice | 1 three-pass kernel | 2 group in shared memory | 3 streaming row-per-warp kernel
int jq_launch_layernorm_fl_sel(const float* x, const float* scale, const float* bias, float* out, long long G, int C, int F,
                               float eps, int force, cudaStream_t st);
int jq_launch_attention_fl(const JqAttnOperand& q, const JqAttnOperand& k, const JqAttnOperand& v, float* out, int ldo,
                           long long W, int n, int H, int dh, int track, cudaStream_t st);
// force: 0 library's choice | 1 block kernel | 2 warp kernel | 3 mma.sync kernel, truncating split | 4 mma.sync kernel,
// round-to-nearest split | 5 mma.sync kernel, sparse phase for one-electron q / k (JQ_ERR_UNSUPPORTED when not eligible)
int jq_launch_attention_fl_sel(const JqAttnOperand& q, const JqAttnOperand& k, const JqAttnOperand& v, float* out, int ldo,
                               long long W, int n, int H, int dh, int track, int force, cudaStream_t st);
