/* jaqmc_b200 -- C ABI of the B200-native JaQMC local-energy + sampling hot path.
 *
 * The reference (bytedance/jaqmc) has no native code and no operator ABI: its hot path is Python/JAX
 * behind three protocols (Wavefunction, SamplerLike, Estimator).  This header is the C boundary a
 * thin XLA-FFI shim (jax.ffi.ffi_call) or ctypes binds; each entry point names the
 * reference interface whose per-walker work it replaces.  All citations are relative to the
 * reference's src/jaqmc/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to float32 unless stated; arrays are dense, row-major;
 *   - the caller (XLA / torch) owns every buffer including the workspace; the library never allocates,
 *     frees, retains a pointer past return, or synchronises: it only enqueues on `stream`;
 *   - walkers are independent units: `n_walkers` is the local shard (walker axis = axis 0);
 *   - parameters keep the reference's Flax layouts (Dense kernel (in,out), DenseGeneral kernel
 *     (in, ndets, n), envelope pi/sigma (n_orb, n_atoms, ndets));
 *   - functions return 0 on success, a JAQMC_ERR_* code otherwise; jaqmc_b200_last_error() returns the
 *     message (thread-local).  Numerical failure (singular determinant) is reported in-band as
 *     -inf / NaN exactly like jnp.linalg.slogdet, never as an error code;
 *   - there is no CPU path: the library is built for sm_100a only.
 */
#ifndef JAQMC_B200_H
#define JAQMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JAQMC_OK 0
#define JAQMC_ERR_INVALID_ARGUMENT 1
#define JAQMC_ERR_WORKSPACE_TOO_SMALL 2
#define JAQMC_ERR_CUDA 3
#define JAQMC_ERR_UNSUPPORTED 4

#define JAQMC_MAX_LAYERS 8

/* Envelope types: wavefunction/output/envelope.py:18-40 (EnvelopeType).  pi (n_orb, n_atoms, ndets); sigma the same
 * shape, except (n_orb, n_atoms, 3, ndets) for the diagonal envelope. */
#define JAQMC_ENVELOPE_ISOTROPIC 0
#define JAQMC_ENVELOPE_ABS_ISOTROPIC 1
#define JAQMC_ENVELOPE_NULL 2
#define JAQMC_ENVELOPE_DIAGONAL 3 /* output/envelope.py:143-163; sigma has shape (n_orb, n_atoms, 3, ndets) */

/* Wavefunction kinds (one descriptor type per reference class). */
#define JAQMC_WF_FERMINET 1 /* app/molecule/wavefunction/ferminet.py:21-96  FermiNetWavefunction */
#define JAQMC_WF_LAPNET 2   /* app/molecule/wavefunction/lapnet.py          LapNetWavefunction   */
#define JAQMC_WF_PSIFORMER 3 /* app/molecule/wavefunction/psiformer.py      PsiformerWavefunction */
#define JAQMC_WF_SOLID_FERMINET 4 /* app/solid/wavefunction.py:91-147       SolidWavefunction    */
#define JAQMC_WF_HYDROGEN 5 /* app/hydrogen_atom.py:28-35                   HydrogenAtom          */

typedef void* jaqmc_stream_t; /* cudaStream_t */

/* ---- FermiNet ---------------------------------------------------------------------------------
 * Fields mirror FermiNetWavefunction's dataclass fields (ferminet.py:43-51) plus the system sizes. */
typedef struct {
  int32_t n_up, n_dn;     /* nspins */
  int32_t n_atoms;
  int32_t ndets;
  int32_t n_layers;       /* len(hidden_dims_single) */
  int32_t hidden_single[JAQMC_MAX_LAYERS];
  int32_t hidden_double[JAQMC_MAX_LAYERS];
  int32_t envelope_type;  /* JAQMC_ENVELOPE_* */
  int32_t orbitals_spin_split;
  int32_t use_last_layer; /* backbone/ferminet.py:45-47: also update the double stream in the last layer and feed the
                             orbitals the aggregated features (width d1 (1 + nch) + nch d2); the tree then holds one more
                             double layer, Dense_{2 n_layers - 1} */
} jaqmc_ferminet_config;

/* Leaves of the Flax tree `params/…`.  backbone_layer/Dense_{2l} is single layer l,
 * Dense_{2l+1} double layer l (backbone/ferminet.py:37-43); only n_layers-1 double layers exist unless
 * use_last_layer is set. */
typedef struct {
  const float* single_kernel[JAQMC_MAX_LAYERS]; /* (fan_in_l, hidden_single[l]) */
  const float* single_bias[JAQMC_MAX_LAYERS];   /* (hidden_single[l],) */
  const float* double_kernel[JAQMC_MAX_LAYERS]; /* (d2_{l-1}, hidden_double[l]) */
  const float* double_bias[JAQMC_MAX_LAYERS];
  const float* orbital_kernel[2]; /* orbital_layer/SplitChannelDense_0/DenseGeneral_{0,1}/kernel (hidden, ndets, n);
                                     [1] is NULL when not spin-split (single DenseGeneral_0) */
  const float* env_pi[2];         /* envelope_layer/{_env_up,_env_down}/pi (n, n_atoms, ndets); [1] NULL -> `_env` */
  const float* env_sigma[2];
} jaqmc_ferminet_params;

/* Gradient buffers of the FermiNet parameters: same leaves and shapes as jaqmc_ferminet_params, DEVICE pointers the
 * library writes (jaqmc_b200_ferminet_logpsi_vjp). */
typedef struct {
  float* single_kernel[JAQMC_MAX_LAYERS];
  float* single_bias[JAQMC_MAX_LAYERS];
  float* double_kernel[JAQMC_MAX_LAYERS];
  float* double_bias[JAQMC_MAX_LAYERS];
  float* orbital_kernel[2];
  float* env_pi[2];
  float* env_sigma[2];
} jaqmc_ferminet_grads;

/* ---- shared output head: orbitals x envelope -> LogDet (+ Jastrow) ----------------------------
 * OrbitalProjection (wavefunction/output/orbital.py:59-78), Envelope (output/envelope.py:98-140),
 * SimpleEEJastrow (wavefunction/jastrow.py:47-122). */
typedef struct {
  const float* orbital_kernel[2]; /* orbital_layer/SplitChannelDense_0/DenseGeneral_{0,1}/kernel (hidden, ndets, n) */
  const float* orbital_bias[2];   /* .../bias (ndets, n) or NULL (use_bias=False, the default) */
  const float* env_pi[2];         /* envelope_layer/{_env_up,_env_down}/pi (n, n_atoms, ndets); [1] NULL -> `_env` */
  const float* env_sigma[2];
  const float* jastrow_alpha_par;  /* jastrow_layer/alpha_par (1,) or NULL (jastrow = none) */
  const float* jastrow_alpha_anti; /* jastrow_layer/alpha_anti (1,) */
} jaqmc_head_params;

/* ---- LapNet ------------------------------------------------------------------------------------
 * Fields mirror LapNetWavefunction (app/molecule/wavefunction/lapnet.py:63-78). */
typedef struct {
  int32_t n_up, n_dn;
  int32_t n_atoms;
  int32_t ndets;
  int32_t num_layers, num_heads, heads_dim;
  int32_t num_local_updates; /* per layer except the last (backbone/lapnet/_backbone.py:184-186) */
  int32_t envelope_type;
  int32_t rescale;           /* log-scaled input features (wavefunction/input/atomic.py:59-67) */
  int32_t use_layernorm;     /* LayerNorm (epsilon 1e-6) before the Q/K projection, before the value projection and
                                before value_update (_backbone.py:62-64,81-111,121) */
} jaqmc_lapnet_config;

/* backbone_layer/input_projection and backbone_layer/layers_{l}/... (_backbone.py:40-64,169-189).
 * Every bias may be NULL (use_input_bias / use_backbone_bias = False). */
typedef struct {
  const float* input_kernel; /* (4*n_atoms+1, hidden) */
  const float* input_bias;
  const float* qk_kernel[JAQMC_MAX_LAYERS];     /* qk_projection (hidden, 2*hidden) */
  const float* qk_bias[JAQMC_MAX_LAYERS];
  const float* value_kernel[JAQMC_MAX_LAYERS];  /* value_projection (hidden, hidden) */
  const float* value_bias[JAQMC_MAX_LAYERS];
  const float* output_kernel[JAQMC_MAX_LAYERS]; /* output_projection */
  const float* output_bias[JAQMC_MAX_LAYERS];
  const float* update_kernel[JAQMC_MAX_LAYERS]; /* value_update */
  const float* update_bias[JAQMC_MAX_LAYERS];
  const float* qk_update_kernel[JAQMC_MAX_LAYERS][4]; /* qk_update_layers_{j}, j < num_local_updates <= 4 */
  const float* qk_update_bias[JAQMC_MAX_LAYERS][4];
  /* layers_{l}/{qk_layernorm, value_layernorm, post_attention_layernorm}/{scale,bias} (hidden,): use_layernorm only */
  const float* qk_ln_scale[JAQMC_MAX_LAYERS];
  const float* qk_ln_bias[JAQMC_MAX_LAYERS];
  const float* value_ln_scale[JAQMC_MAX_LAYERS];
  const float* value_ln_bias[JAQMC_MAX_LAYERS];
  const float* post_ln_scale[JAQMC_MAX_LAYERS];
  const float* post_ln_bias[JAQMC_MAX_LAYERS];
  jaqmc_head_params head;
} jaqmc_lapnet_params;

/* ---- Psiformer ---------------------------------------------------------------------------------
 * Fields mirror PsiformerWavefunction (app/molecule/wavefunction/psiformer.py:77-94). */
#define JAQMC_LAYERNORM_PRE 0
#define JAQMC_LAYERNORM_POST 1
#define JAQMC_LAYERNORM_NULL 2
#define JAQMC_MAX_MLP 4
typedef struct {
  int32_t n_up, n_dn;
  int32_t n_atoms;
  int32_t ndets;
  int32_t num_layers, num_heads, heads_dim;
  int32_t n_mlp_hidden;                 /* len(mlp_hidden_dims) <= JAQMC_MAX_MLP - 1 */
  int32_t mlp_hidden[JAQMC_MAX_MLP];
  int32_t layer_norm_mode;              /* JAQMC_LAYERNORM_* (backbone/psiformer.py:57,72-98) */
  int32_t envelope_type;
  int32_t orbitals_spin_split;
  int32_t rescale;
} jaqmc_psiformer_config;

/* backbone_layer/Dense_0 and backbone_layer/PsiformerLayer_{l}/... (backbone/psiformer.py:60-99,175-185);
 * flax MultiHeadDotProductAttention kernels: query/key/value (hidden, heads, head_dim) == row-major (hidden, hidden),
 * out (heads, head_dim, hidden) == (hidden, hidden). LayerNorm epsilon is 1e-5. */
typedef struct {
  const float* input_kernel; /* (4*n_atoms+1, hidden) */
  const float* input_bias;   /* or NULL */
  const float* ln0_scale[JAQMC_MAX_LAYERS];
  const float* ln0_bias[JAQMC_MAX_LAYERS];
  const float* q_kernel[JAQMC_MAX_LAYERS];
  const float* q_bias[JAQMC_MAX_LAYERS];
  const float* k_kernel[JAQMC_MAX_LAYERS];
  const float* k_bias[JAQMC_MAX_LAYERS];
  const float* v_kernel[JAQMC_MAX_LAYERS];
  const float* v_bias[JAQMC_MAX_LAYERS];
  const float* out_kernel[JAQMC_MAX_LAYERS];
  const float* out_bias[JAQMC_MAX_LAYERS];
  const float* ln1_scale[JAQMC_MAX_LAYERS];
  const float* ln1_bias[JAQMC_MAX_LAYERS];
  const float* mlp_kernel[JAQMC_MAX_LAYERS][JAQMC_MAX_MLP]; /* Dense_{j}: j < n_mlp_hidden hidden layers, then Dense_{n_mlp_hidden} -> hidden */
  const float* mlp_bias[JAQMC_MAX_LAYERS][JAQMC_MAX_MLP];
  jaqmc_head_params head;
} jaqmc_psiformer_params;

/* ---- periodic FermiNet (solid) ------------------------------------------------------------------
 * SolidWavefunction (app/solid/wavefunction.py:40-147): `tri` distance features with minimal symmetry (7 per primitive
 * atom / per pair), FermiLayers, real + imaginary orbital projections, envelope on the periodic distance, Bloch
 * phases exp(i k.r).  log psi is complex. */
typedef struct {
  jaqmc_ferminet_config net;      /* n_atoms = atoms of the PRIMITIVE cell (SolidData.primitive_atoms) */
  float simulation_lattice[9];    /* rows are lattice vectors (host values) */
  float primitive_lattice[9];
  int32_t distance_type;          /* JAQMC_DISTANCE_* (geometry/pbc.py DistanceType): tri -> 7, nu -> 4 features per pair */
  int32_t sym_type;               /* JAQMC_SYMMETRY_* (geometry/pbc.py:347-381 get_symmetry_lat) */
} jaqmc_solid_config;

typedef struct {
  jaqmc_ferminet_params net;             /* backbone_layer/Dense_*, envelope_layer; net.orbital_kernel is unused */
  const float* real_orbital_kernel[2];   /* real_orbital_layer/.../kernel (hidden, ndets, n) per spin channel */
  const float* imag_orbital_kernel[2];   /* imag_orbital_layer/.../kernel */
  const float* klist;                    /* (n, 3) k-point of every orbital (module attribute `klist`) */
} jaqmc_solid_params;

#define JAQMC_DISTANCE_TRI 0
#define JAQMC_DISTANCE_NU 1
#define JAQMC_SYMMETRY_MINIMAL 0
#define JAQMC_SYMMETRY_FCC 1
#define JAQMC_SYMMETRY_BCC 2
#define JAQMC_SYMMETRY_HEXAGONAL 3

/* ---- HydrogenAtom demo wavefunction (app/hydrogen_atom.py:28-35): log psi = alpha * |electrons| ------------------ */
typedef struct {
  int32_t n_electrons; /* 1 in the reference app */
} jaqmc_hydrogen_config;
typedef struct {
  const float* alpha;  /* params/alpha, scalar */
} jaqmc_hydrogen_params;

/* ---- generic wavefunction descriptor ---------------------------------------------------------- */
typedef struct {
  int32_t kind;       /* JAQMC_WF_* */
  const void* config; /* HOST pointer to the kind's config struct */
  const void* params; /* HOST pointer to the kind's params struct (of device pointers) */
} jaqmc_wavefunction;

/* System data replicated across walkers: MoleculeData.atoms / charges (app/molecule/data.py:46-53). */
typedef struct {
  const float* atoms;   /* (n_atoms, 3) */
  const float* charges; /* (n_atoms,)  may be NULL where unused */
  int32_t n_atoms;
} jaqmc_system;

/* ---- Ewald sum (solid-state potential) ---------------------------------------------------------
 * Precomputed by the host exactly as EwaldSum.__init__ does (estimator/ewald.py:50-110): all pointers are DEVICE
 * arrays of float32. */
typedef struct {
  const float* lattice;     /* (3,3) supercell lattice vectors, rows */
  const float* inv_lattice; /* (3,3) inverse (orthogonal-cell minimum image), or NULL */
  const float* mic_shifts;  /* (27,3) neighbouring-image shifts in the reference's meshgrid order (general cell), or NULL */
  const float* images;      /* (n_images,3) lattice_displacements of the real-space sum */
  const float* gpoints;     /* (n_g,3) selected reciprocal vectors (half space) */
  const float* gweight;     /* (n_g,) 4 pi exp(-G^2/4 alpha^2) / (V G^2) */
  int32_t n_images, center_image, n_g;
  int32_t mic_kind;         /* 0 diagonal, 1 orthogonal, 2 general (geometry/pbc.py:114-184) */
  float alpha, self_const_factor, ijconst;
} jaqmc_ewald;

/* Replaces PotentialEnergy.evaluate_single_walker (app/solid/hamiltonian.py:28-56) / EwaldSum.energy
 * (estimator/ewald.py:112-173) vmapped over walkers: electrons (n_walkers, n_electrons, 3) with charge -1 and the
 * supercell ions atoms (n_atoms,3) / charges (n_atoms,) -> e_pot (n_walkers,). */
int jaqmc_b200_ewald(const jaqmc_ewald* ewald, const float* electrons, int64_t n_walkers, int32_t n_electrons,
                     const float* atoms, const float* charges, int32_t n_atoms, float* e_pot, jaqmc_stream_t stream);


/* Workspace (bytes) that lets a call process all `n_walkers` walkers in one pass.  A smaller workspace is legal:
 * the library then tiles the walker axis, as long as one walker fits (else JAQMC_ERR_WORKSPACE_TOO_SMALL).
 * `track` = 0 for log|psi| only, 1 for value + gradient + Laplacian. */
size_t jaqmc_b200_workspace_bytes(const jaqmc_wavefunction* wf, int64_t n_walkers, int track);

/* Replaces vmap(wf.logpsi / wf.phase_logpsi) over the walker axis
 * (sampler/base.py:136-138; app/molecule/wavefunction/ferminet.py:98-124).
 *   electrons (n_walkers, n, 3) -> logpsi (n_walkers,), sign (n_walkers,) in {-1, 0, +1}.
 * For JAQMC_WF_SOLID_FERMINET log psi is complex: logpsi receives its real part (what the sampler uses,
 * sampler/mcmc.py:127) and `sign` the phase angle Im log psi in (-pi, pi]. */
int jaqmc_b200_logpsi(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons,
                      int64_t n_walkers, float* logpsi, float* sign, void* workspace, size_t workspace_bytes,
                      jaqmc_stream_t stream);

/* Replaces vmap(wf.orbitals) -- the pretraining head (app/molecule/wavefunction/base.py:60-72, ferminet.py:126-138,
 * app/solid/wavefunction.py `orbitals`; consumer utils/atomic/pretrain.py:92-143): the orbital matrices after the
 * envelope (and Bloch phase), before the determinant.
 *   electrons (n_walkers, n, 3) -> orbitals (n_walkers, ndets, n, n) [electron, orbital]; JAQMC_WF_SOLID_FERMINET:
 *   complex, interleaved (n_walkers, ndets, n, n, 2).  Workspace as for jaqmc_b200_logpsi (track = 0) + 2 floats per walker. */
int jaqmc_b200_orbitals(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons,
                        int64_t n_walkers, float* orbitals, void* workspace, size_t workspace_bytes,
                        jaqmc_stream_t stream);

/* psi-ratio consumers.  Replaces the vmapped `phase_logpsi` evaluations of the ECP non-local
 * integral (estimator/ecp/nonlocal_integral.py:43-110: one electron displaced to quadrature points around an atom) and
 * of SpinSquared (estimator/spin.py:87-146: a minority-spin electron exchanged with each majority-spin electron): for
 * every walker, `n_moves` configurations that differ from the walker's in at most two electrons.
 *   move_index (n_moves, 2) int32 DEVICE: the electrons replaced in move q (second entry -1: only one);
 *   move_pos   (n_walkers, n_moves, 2, 3): their new positions (entries of unused slots are ignored);
 *   log_ratio  (n_walkers, n_moves) = log|psi(moved)| - log|psi(walker)|;
 *   sign_ratio (n_walkers, n_moves) = sign(psi(moved)) * sign(psi(walker))  (periodic network: the phase difference
 *              Im log psi(moved) - Im log psi(walker), wrapped into (-pi, pi]).
 * A workspace smaller than the single-pass size tiles the configurations; it must hold one configuration. */
int jaqmc_b200_psi_ratios(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons,
                          int64_t n_walkers, int32_t n_moves, const int32_t* move_index, const float* move_pos,
                          float* log_ratio, float* sign_ratio, void* workspace, size_t workspace_bytes,
                          jaqmc_stream_t stream);

/* Replaces the estimator half of EvaluationWorkStage.compute_step for the energy keys
 * (workflow/stage/evaluation.py:190-192): EuclideanKinetic in forward_laplacian mode
 * (estimator/kinetic/euclidean.py:114-135, laplacian/interpreter.py:392-438), the potential
 * (app/molecule/hamiltonian.py:9-22) and TotalEnergy (estimator/total_energy.py:36-60), vmapped over walkers.
 * Outputs (any may be NULL except logpsi/sign):
 *   grad (n_walkers, 3n) = d log|psi| / d r,   lap (n_walkers,) = laplacian of log|psi|,
 *   e_kin = -1/2 (lap + |grad|^2),  e_pot,  e_loc = e_kin + e_pot,
 *   sums (3,) += {sum e_loc, sum e_loc^2, count of finite e_loc} over this call's walkers (the per-device partial
 *   sums that precede the pmean of estimator/base.py:27-53); the caller zeroes it. */
int jaqmc_b200_local_energy(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const float* electrons,
                            int64_t n_walkers, float* logpsi, float* sign, float* grad, float* lap, float* e_kin,
                            float* e_pot, float* e_loc, float* sums, void* workspace, size_t workspace_bytes,
                            jaqmc_stream_t stream);

/* Complex-valued counterpart of jaqmc_b200_local_energy for JAQMC_WF_SOLID_FERMINET: EuclideanKinetic on the complex
 * log psi (estimator/kinetic/euclidean.py:114-135 with the complex square of kinetic/_common.py:61-73), the Ewald
 * potential (app/solid/hamiltonian.py:18-56) and TotalEnergy.  `sys` holds the primitive-cell atoms the network sees;
 * `ewald` / `cell_atoms` (n_cell_atoms,3) / `cell_charges` describe the simulation cell for the potential (ewald NULL:
 * e_pot is not computed).  Complex outputs are interleaved (re, im):
 *   logpsi (W,2), grad (W,3n,2), lap (W,2), e_kin (W,2), e_pot (W,), e_loc (W,2) = e_kin + e_pot,
 *   sums (3,) += {sum Re e_loc, sum (Re e_loc)^2, finite count}.  Any output except logpsi may be NULL. */
int jaqmc_b200_local_energy_complex(const jaqmc_wavefunction* wf, const jaqmc_system* sys, const jaqmc_ewald* ewald,
                                    const float* cell_atoms, const float* cell_charges, int32_t n_cell_atoms,
                                    const float* electrons, int64_t n_walkers, float* logpsi, float* grad, float* lap,
                                    float* e_kin, float* e_pot, float* e_loc, float* sums, void* workspace,
                                    size_t workspace_bytes, jaqmc_stream_t stream);

/* Parameter-gradient path.  Replaces the per-walker `jax.value_and_grad(wf.logpsi)` of
 * LossAndGrad.evaluate_single_walker and the contraction with the clipped local energies of LossAndGrad.reduce /
 * finalize_stats (estimator/loss_grad.py:70-128) by ONE reverse pass: the VJP of theta -> vmap(log|psi|)(theta, walkers),
 *   grads[leaf] = sum_w cotangent[w] * d log|psi|(x_w) / d leaf        (every leaf of `grads` is OVERWRITTEN),
 * so that LossAndGrad's `grads` is the call with cotangent[w] = 2 (E_clip[w] - mean E_clip) / W and the mean score the
 * call with 1 / W; the W x P per-walker score tensor of the reference is never formed.  Also returns log|psi| / sign of
 * the walkers (either may be NULL).  FermiNet kind, n <= 32 electrons, isotropic / abs_isotropic / null envelope.
 * Needs the whole workspace of jaqmc_b200_ferminet_vjp_workspace_bytes (activations are kept; no walker tiling).
 * Results are bit-reproducible (row reductions are split and summed in a fixed order). */
size_t jaqmc_b200_ferminet_vjp_workspace_bytes(const jaqmc_ferminet_config* config, int64_t n_walkers);
int jaqmc_b200_ferminet_logpsi_vjp(const jaqmc_ferminet_config* config, const jaqmc_ferminet_params* params,
                                   const jaqmc_system* sys, const float* electrons, int64_t n_walkers,
                                   const float* cotangent, const jaqmc_ferminet_grads* grads, float* logpsi, float* sign,
                                   void* workspace, size_t workspace_bytes, jaqmc_stream_t stream);

/* Replaces potential_energy (app/molecule/hamiltonian.py:9-22) vmapped over walkers. */
int jaqmc_b200_coulomb(const jaqmc_system* sys, const float* electrons, int64_t n_walkers, int32_t n_electrons,
                       float* e_pot, jaqmc_stream_t stream);

/* Replaces MCMCSampler.step's fori_loop of _mh_update (sampler/mcmc.py:96-137,167-180) for `n_steps` all-electron
 * moves with host-supplied noise:
 *   electrons (n_walkers, n, 3) in/out; logpsi (n_walkers,) in/out (log|psi| of `electrons`; computed on entry
 *   when `logpsi_valid` == 0); normals (n_steps, n_walkers, n, 3); uniforms (n_steps, n_walkers) in (0,1);
 *   stddev (1,) device scalar; n_accept (1,) += accepted moves (caller zeroes); accepted (n_steps, n_walkers) u8 or NULL.
 * accept iff 2*(logpsi' - logpsi) > log(u)  (sampler/base.py:173-178 supplies the factor 2). */
int jaqmc_b200_mh_step(const jaqmc_wavefunction* wf, const jaqmc_system* sys, float* electrons, float* logpsi,
                       int32_t logpsi_valid, const float* normals, const float* uniforms, const float* stddev,
                       int32_t n_steps, int64_t n_walkers, float* n_accept, uint8_t* accepted, void* workspace,
                       size_t workspace_bytes, jaqmc_stream_t stream);

/* Same with proposals wrapped into the periodic cell (geometry/pbc.py:187-201 make_pbc_gaussian_proposal, used by
 * `jaqmc solid train`): x2 = wrap_positions(x1 + normal * stddev, lattice); `lattice` is a HOST pointer to 9 floats. */
int jaqmc_b200_mh_step_pbc(const jaqmc_wavefunction* wf, const jaqmc_system* sys, float* electrons, float* logpsi,
                           int32_t logpsi_valid, const float* normals, const float* uniforms, const float* stddev,
                           int32_t n_steps, int64_t n_walkers, float* n_accept, uint8_t* accepted, const float* lattice,
                           void* workspace, size_t workspace_bytes, jaqmc_stream_t stream);

/* One dense layer under the forward Laplacian: replaces nn.Dense traced by forward_laplacian
 * (laplacian/primitives/dot_general.py:377-407: one GEMM over the rows {x, J_1..J_K, L}, bias on the value row) fused
 * with the tanh rule (laplacian/primitives/elementwise.py:42-72) and FermiNet's residual (backbone/ferminet.py:59-63).
 *   x (n_groups, n_components, k0) [, x2 (.., k1) concatenated along the contraction axis with kernel2 (k1, n_out)],
 *   kernel (k0, n_out), bias (n_out,) or NULL, addend (n_groups/groups_per_walker, n_components, n_out) or NULL
 *   (broadcast over a walker's groups), residual / out (n_groups, n_components, n_out).
 *   n_components = 1 (value only) or K+2 (value, K Jacobian columns, Laplacian).
 *   activation 0 none | 1 tanh;  residual_mode 0 none | 1 (res + y)/sqrt(2) | 2 res + y.
 *   use_tensor_cores != 0 routes eligible shapes (k0, k1 multiples of 32, 64 <= n_out <= 256) to the tcgen05 3xTF32
 *   kernel and needs workspace >= 2*(k0+k1)*n_out floats; 0 forces the CUDA-core FP32 kernel; 2 forces the
 *   weight-streaming single-CTA variant of the tcgen05 kernel (1 prefers the CTA-pair variant with resident weights). */
int jaqmc_b200_dense_fl(const float* x, const float* x2, const float* kernel, const float* kernel2, const float* bias,
                        const float* addend, const float* residual, float* out, int64_t n_groups, int32_t n_components,
                        int32_t k0, int32_t k1, int32_t n_out, int32_t groups_per_walker, int32_t activation,
                        int32_t residual_mode, int32_t use_tensor_cores, void* workspace, size_t workspace_bytes,
                        jaqmc_stream_t stream);

/* Softmax self-attention core under the forward Laplacian: replaces `attention_core` traced by forward_laplacian
 * (wavefunction/backbone/psiformer.py attention block; backbone/lapnet/_attention.py:20-34, :113-196 for the
 * one-electron query / key structure).  Operands are augmented tensors (n_walkers, n_electrons, C, n_heads * head_dim)
 * with C = 3 n + 2 components {value, 3n Jacobian columns, Laplacian}; q and k may instead carry C = 5 ("local":
 * electron i holds only its own three Jacobian columns, the rest are zero).  out is always dense.
 *   kernel 0: the library's choice for the shape; 1: CUDA-core block kernel; 2: CUDA-core warp-per-component kernel
 *   (n <= 14 at head_dim 64: shared-memory limit); 3: mma.sync 3xTF32 tensor-core kernel (n <= 48, head_dim 64), operands split by truncation; 4: the same with
 *   round-to-nearest operand splits; 5: kernel 3 with the sparse logit-Jacobian phase for one-electron q / k operands
 *   (q_components = k_components = 5 only).  A forced kernel that does not support the shape
 *   returns JAQMC_ERR_UNSUPPORTED. */
int jaqmc_b200_attention_fl(const float* q, const float* k, const float* v, float* out, int64_t n_walkers,
                            int32_t n_electrons, int32_t n_heads, int32_t head_dim, int32_t q_components,
                            int32_t k_components, int32_t kernel, jaqmc_stream_t stream);

/* LayerNorm over the feature axis under the forward Laplacian (flax nn.LayerNorm traced by forward_laplacian in the
 * Psiformer block, backbone/psiformer.py, and LapNet's optional layer norms, backbone/lapnet/_backbone.py:62-64):
 *   x, out (n_groups, n_components, n_features), n_components = 1 or K+2; scale / bias (n_features,) or NULL.
 *   kernel 0: the library's choice; 1: three-pass kernel (any shape); 2: group cached in shared memory; 3: streaming
 *   row-per-warp kernel (n_features 128, 256 or 512, 16-byte aligned).  In-place is not supported. */
int jaqmc_b200_layernorm_fl(const float* x, const float* scale, const float* bias, float* out, int64_t n_groups,
                            int32_t n_components, int32_t n_features, float epsilon, int32_t kernel,
                            jaqmc_stream_t stream);

/* Building blocks of the MH step, exported for samplers that drive their own loop
 * (sampler/mcmc.py:53-54 gaussian_proposal; :128-137 accept/select). */
int jaqmc_b200_mh_propose(const float* x1, const float* normals, const float* stddev, float* x2, int64_t count,
                          jaqmc_stream_t stream);
int jaqmc_b200_mh_accept(float* x1, const float* x2, float* logprob1, const float* logprob2, const float* uniforms,
                         int64_t n_walkers, int32_t row, float* n_accept, uint8_t* accepted, jaqmc_stream_t stream);

/* ---- parameter leaves -> descriptor structs -----------------------------------------------------------------------
 * An XLA-FFI handler receives the parameters as a flat operand list in `jax.tree.leaves(params)` order (dictionary
 * keys sorted at every level; reference trees: SURVEY.md Appendix B / tests/golden/ref_*.npz).  These three functions
 * are the single definition of that order for every wavefunction kind (host-only code, no CUDA calls):
 *   param_leaf_count  -> number of leaves of (kind, config), or -1;
 *   param_leaf_info   -> path ("params/backbone_layer/Dense_0/bias"), element count and shape of leaf `index`;
 *   bind_param_leaves -> writes leaves[i] into the field of `params` that leaf i belongs to, after checking count and
 *                        element counts (leaf_elements may be NULL to skip the size check).
 * Optional leaves (biases, Jastrow) exist or not depending on the reference class's flags; the caller states which
 * exist by pre-setting these fields of `params` to any non-NULL value before the call: input_bias, qk_bias[l] (all
 * backbone biases of LapNet layer l), q_bias[l] (with_bias, Psiformer layer l), head.orbital_bias[0],
 * head.jastrow_alpha_par.  All other fields are outputs.  `params` is the kind's params struct; for the periodic
 * network `klist` is not a parameter (module attribute) and stays the caller's to set. */
int jaqmc_b200_param_leaf_count(int32_t kind, const void* config, void* params);
int jaqmc_b200_param_leaf_info(int32_t kind, const void* config, void* params, int32_t index, char* path,
                               size_t path_cap, int64_t* n_elements, int32_t* rank, int64_t* dims);
int jaqmc_b200_bind_param_leaves(int32_t kind, const void* config, void* params, const float* const* leaves,
                                 const int64_t* leaf_elements, int32_t n_leaves);

/* Number of kernels the library has launched on this thread since the last reset (bench.py's gpu_launches). */
int64_t jaqmc_b200_launch_count(void);
void jaqmc_b200_reset_launch_count(void);

/* Per-kernel device timing for the roofline report (bench.py).  While enabled, every kernel launch on the calling
 * thread is bracketed by CUDA events on its stream.  `fetch` waits for them and writes one text line per kernel,
 * "name launches total_ms algorithmic_flops algorithmic_bytes", returning the bytes written. */
void jaqmc_b200_profile_enable(int on);
size_t jaqmc_b200_profile_fetch(char* buf, size_t cap);

const char* jaqmc_b200_last_error(void);
const char* jaqmc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* JAQMC_B200_H */
