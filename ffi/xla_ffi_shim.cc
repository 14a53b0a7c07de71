// XLA-FFI binding of the jaqmc_b200 C ABI (include/jaqmc_b200.h) -- the file a JAX-side install compiles:
//
//   make -C ffi FFI_INCLUDE="$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')"
//       (g++ -shared xla_ffi_shim.cc -I include -I $FFI_INCLUDE -ljaqmc_b200 -> jaqmc_b200/_C/libjaqmc_b200_ffi.so)
//
// and jaqmc_b200_jax/_ffi.py registers (jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(lib.<name>), platform="CUDA")).
// In the build container there is no jaxlib: `make -C ffi check` compiles this file against ffi/stub/xla/ffi/api/ffi.h,
// a declaration-only stand-in of the parts of the header-only XLA FFI API used here (syntax / type check only).
//
// One handler per entry point; the wavefunction KIND is an attribute, so each handler serves the five reference classes
// (FermiNetWavefunction, LapNetWavefunction, PsiformerWavefunction, SolidWavefunction, HydrogenAtom):
//
//   attribute  kind      : i32            JAQMC_WF_*
//   attribute  config    : i32[]          the kind's config struct, int32 word by word (include/jaqmc_b200.h)
//   attribute  fconfig   : f32[]          float tail of the config (periodic network: simulation + primitive lattice)
//   attribute  optional  : i32            bit 0 input bias, bit 1 backbone / attention biases, bit 2 orbital bias,
//                                         bit 3 Jastrow (which optional leaves the tree holds)
//   operands   electrons [..., n, 3], atoms [..., A, 3] (+ charges, noise ...), then EVERY parameter leaf in
//              jax.tree.leaves(params) order -- bound to the descriptor by jaqmc_b200_bind_param_leaves, the single
//              definition of that order (csrc/leaves.cu).
//
// Leading batch dimensions: the reference calls one-walker functions under jax.vmap; with
// ffi_call(..., vmap_method="expand_dims") the handler sees electrons [W, n, 3] and size-1 leading axes on the replicated
// operands.  W = product of the leading dimensions of `electrons`.
//
// Contract: XLA owns every buffer; scratch comes from ffi::ScratchAllocator; the handler only enqueues
// on the stream it is given, never synchronises or allocates, and reports shape / attribute errors as ffi::Error.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "jaqmc_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F32 = ffi::Buffer<ffi::F32>;
using F32Out = ffi::ResultBuffer<ffi::F32>;
using U8Out = ffi::ResultBuffer<ffi::U8>;

ffi::Error Invalid(const std::string& msg) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, msg); }

ffi::Error FromStatus(int rc) {
  if (rc == JAQMC_OK) return ffi::Error::Success();
  const ffi::ErrorCode code = rc == JAQMC_ERR_INVALID_ARGUMENT  ? ffi::ErrorCode::kInvalidArgument
                              : rc == JAQMC_ERR_WORKSPACE_TOO_SMALL ? ffi::ErrorCode::kResourceExhausted
                              : rc == JAQMC_ERR_UNSUPPORTED       ? ffi::ErrorCode::kUnimplemented
                                                                  : ffi::ErrorCode::kInternal;
  return ffi::Error(code, std::string("jaqmc_b200: ") + jaqmc_b200_last_error());
}

// Wavefunction descriptor assembled from the attributes and the trailing parameter operands.
struct Descriptor {
  jaqmc_wavefunction wf;
  union {
    jaqmc_ferminet_config ferminet;
    jaqmc_lapnet_config lapnet;
    jaqmc_psiformer_config psiformer;
    jaqmc_solid_config solid;
    jaqmc_hydrogen_config hydrogen;
  } config;
  union {
    jaqmc_ferminet_params ferminet;
    jaqmc_lapnet_params lapnet;
    jaqmc_psiformer_params psiformer;
    jaqmc_solid_params solid;
    jaqmc_hydrogen_params hydrogen;
  } params;
  int n_electrons = 0;
};

size_t ConfigWords(int kind) {
  switch (kind) {
    case JAQMC_WF_FERMINET: return sizeof(jaqmc_ferminet_config) / 4;
    case JAQMC_WF_LAPNET: return sizeof(jaqmc_lapnet_config) / 4;
    case JAQMC_WF_PSIFORMER: return sizeof(jaqmc_psiformer_config) / 4;
    case JAQMC_WF_SOLID_FERMINET: return sizeof(jaqmc_ferminet_config) / 4 + 2;   // + distance_type, sym_type; 18 floats in fconfig
    case JAQMC_WF_HYDROGEN: return sizeof(jaqmc_hydrogen_config) / 4;
    default: return 0;
  }
}

ffi::Error Assemble(int32_t kind, ffi::Span<const int32_t> config, ffi::Span<const float> fconfig, int32_t optional,
                    ffi::RemainingArgs leaves, size_t first_leaf, const float* klist, Descriptor* d) {
  std::memset(static_cast<void*>(d), 0, sizeof(*d));
  const size_t words = ConfigWords(kind);
  if (words == 0) return Invalid("jaqmc_b200: unknown wavefunction kind " + std::to_string(kind));
  if (config.size() != words)
    return Invalid("jaqmc_b200: attribute `config` has " + std::to_string(config.size()) + " words, the kind needs " +
                   std::to_string(words));
  std::memcpy(&d->config, config.begin(), (kind == JAQMC_WF_SOLID_FERMINET ? words - 2 : words) * 4);
  const float* present = reinterpret_cast<const float*>(1);   // "this optional leaf exists" marker (see header)
  switch (kind) {
    case JAQMC_WF_FERMINET:
      d->n_electrons = d->config.ferminet.n_up + d->config.ferminet.n_dn;
      break;
    case JAQMC_WF_SOLID_FERMINET:
      if (fconfig.size() != 18) return Invalid("jaqmc_b200: attribute `fconfig` must hold the two 3x3 lattices");
      std::memcpy(d->config.solid.simulation_lattice, fconfig.begin(), 9 * 4);
      std::memcpy(d->config.solid.primitive_lattice, fconfig.begin() + 9, 9 * 4);
      d->config.solid.distance_type = config.begin()[words - 2];
      d->config.solid.sym_type = config.begin()[words - 1];
      d->n_electrons = d->config.solid.net.n_up + d->config.solid.net.n_dn;
      break;
    case JAQMC_WF_LAPNET:
      d->n_electrons = d->config.lapnet.n_up + d->config.lapnet.n_dn;
      if (optional & 1) d->params.lapnet.input_bias = present;
      if (optional & 2)
        for (int l = 0; l < JAQMC_MAX_LAYERS; ++l) d->params.lapnet.qk_bias[l] = present;
      if (optional & 4) d->params.lapnet.head.orbital_bias[0] = present;
      if (optional & 8) d->params.lapnet.head.jastrow_alpha_par = present;
      break;
    case JAQMC_WF_PSIFORMER:
      d->n_electrons = d->config.psiformer.n_up + d->config.psiformer.n_dn;
      if (optional & 1) d->params.psiformer.input_bias = present;
      if (optional & 2)
        for (int l = 0; l < JAQMC_MAX_LAYERS; ++l) d->params.psiformer.q_bias[l] = present;
      if (optional & 4) d->params.psiformer.head.orbital_bias[0] = present;
      if (optional & 8) d->params.psiformer.head.jastrow_alpha_par = present;
      break;
    case JAQMC_WF_HYDROGEN:
      d->n_electrons = d->config.hydrogen.n_electrons;
      break;
  }
  const size_t n_leaves = leaves.size() - first_leaf;
  std::vector<const float*> ptrs(n_leaves);
  std::vector<int64_t> sizes(n_leaves);
  for (size_t i = 0; i < n_leaves; ++i) {
    auto buf = leaves.get<ffi::AnyBuffer>(first_leaf + i);
    if (!buf.has_value()) return buf.error();
    if (buf->element_type() != ffi::F32) return Invalid("jaqmc_b200: parameter leaf " + std::to_string(i) + " is not float32");
    ptrs[i] = static_cast<const float*>(buf->untyped_data());
    sizes[i] = static_cast<int64_t>(buf->element_count());   // size-1 leading vmap axes do not change the count
  }
  int rc = jaqmc_b200_bind_param_leaves(kind, &d->config, &d->params, ptrs.data(), sizes.data(), (int32_t)n_leaves);
  if (rc) return FromStatus(rc);
  if (kind == JAQMC_WF_SOLID_FERMINET) {
    if (!klist) return Invalid("jaqmc_b200: the periodic network needs the klist operand");
    d->params.solid.klist = klist;
  }
  d->wf.kind = kind;
  d->wf.config = &d->config;
  d->wf.params = &d->params;
  return ffi::Error::Success();
}

// electrons [..., n, 3] -> number of walkers
ffi::Error Walkers(const F32& electrons, int n, int64_t* W) {
  auto dims = electrons.dimensions();
  if (dims.size() < 2 || dims[dims.size() - 1] != 3 || dims[dims.size() - 2] != n)
    return Invalid("jaqmc_b200: electrons must have shape [..., " + std::to_string(n) + ", 3]");
  int64_t w = 1;
  for (size_t i = 0; i + 2 < dims.size(); ++i) w *= dims[i];
  *W = w;
  return ffi::Error::Success();
}

// Scratch from XLA's allocator: the full single-pass size if available, else smaller (the library tiles the walker axis).
ffi::Error Workspace(ffi::ScratchAllocator& scratch, const jaqmc_wavefunction* wf, int64_t W, int track, void** ws,
                     size_t* bytes) {
  size_t need = jaqmc_b200_workspace_bytes(wf, W, track);
  if (need == 0) return FromStatus(JAQMC_ERR_INVALID_ARGUMENT);
  const size_t floor = jaqmc_b200_workspace_bytes(wf, 1, track) + (size_t)W * 64;
  for (size_t ask = need; ask >= floor; ask /= 2) {
    auto p = scratch.Allocate(ask, 256);
    if (p.has_value()) {
      *ws = *p;
      *bytes = ask;
      return ffi::Error::Success();
    }
  }
  return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaqmc_b200: no scratch memory for one walker");
}

jaqmc_system System(const F32& atoms, const float* charges) {
  auto dims = atoms.dimensions();
  jaqmc_system s;
  s.atoms = atoms.typed_data();
  s.charges = charges;
  s.n_atoms = (int32_t)dims[dims.size() - 2];
  return s;
}

// ---- vmap(wf.logpsi / wf.phase_logpsi)  (sampler/base.py:136-138) -------------------------------------------------
ffi::Error LogPsiImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int32_t kind, ffi::Span<const int32_t> config,
                      ffi::Span<const float> fconfig, int32_t optional, F32 electrons, F32 atoms, ffi::RemainingArgs rest,
                      F32Out logpsi, F32Out sign) {
  Descriptor d;
  const bool solid = kind == JAQMC_WF_SOLID_FERMINET;
  const float* klist = nullptr;
  if (solid) {
    auto k = rest.get<F32>(0);
    if (!k.has_value()) return k.error();
    klist = k->typed_data();
  }
  if (auto e = Assemble(kind, config, fconfig, optional, rest, solid ? 1 : 0, klist, &d); e.failure()) return e;
  int64_t W = 0;
  if (auto e = Walkers(electrons, d.n_electrons, &W); e.failure()) return e;
  void* ws = nullptr;
  size_t bytes = 0;
  if (auto e = Workspace(scratch, &d.wf, W, 0, &ws, &bytes); e.failure()) return e;
  jaqmc_system sys = System(atoms, nullptr);
  return FromStatus(jaqmc_b200_logpsi(&d.wf, &sys, electrons.typed_data(), W, logpsi->typed_data(), sign->typed_data(), ws,
                                      bytes, stream));
}

// ---- vmap(wf.orbitals)  (app/molecule/wavefunction/base.py:60-72) ---------------------------------------------------
ffi::Error OrbitalsImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int32_t kind, ffi::Span<const int32_t> config,
                        ffi::Span<const float> fconfig, int32_t optional, F32 electrons, F32 atoms,
                        ffi::RemainingArgs rest, F32Out orbitals) {
  Descriptor d;
  const bool solid = kind == JAQMC_WF_SOLID_FERMINET;
  const float* klist = nullptr;
  if (solid) {
    auto k = rest.get<F32>(0);
    if (!k.has_value()) return k.error();
    klist = k->typed_data();
  }
  if (auto e = Assemble(kind, config, fconfig, optional, rest, solid ? 1 : 0, klist, &d); e.failure()) return e;
  int64_t W = 0;
  if (auto e = Walkers(electrons, d.n_electrons, &W); e.failure()) return e;
  void* ws = nullptr;
  size_t bytes = 0;
  if (auto e = Workspace(scratch, &d.wf, W, 0, &ws, &bytes); e.failure()) return e;
  jaqmc_system sys = System(atoms, nullptr);
  return FromStatus(jaqmc_b200_orbitals(&d.wf, &sys, electrons.typed_data(), W, orbitals->typed_data(), ws, bytes, stream));
}

// ---- forward-Laplacian local energy (estimator/kinetic/euclidean.py:114-135 + app/molecule/hamiltonian.py:9-22) --------
ffi::Error LocalEnergyImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int32_t kind,
                           ffi::Span<const int32_t> config, ffi::Span<const float> fconfig, int32_t optional,
                           F32 electrons, F32 atoms, F32 charges, ffi::RemainingArgs rest, F32Out logpsi, F32Out sign,
                           F32Out grad, F32Out lap, F32Out e_kin, F32Out e_pot, F32Out e_loc, F32Out sums) {
  if (kind == JAQMC_WF_SOLID_FERMINET) return Invalid("jaqmc_b200: use jaqmc_b200_ffi_local_energy_complex for the periodic network");
  Descriptor d;
  if (auto e = Assemble(kind, config, fconfig, optional, rest, 0, nullptr, &d); e.failure()) return e;
  int64_t W = 0;
  if (auto e = Walkers(electrons, d.n_electrons, &W); e.failure()) return e;
  void* ws = nullptr;
  size_t bytes = 0;
  if (auto e = Workspace(scratch, &d.wf, W, 1, &ws, &bytes); e.failure()) return e;
  jaqmc_system sys = System(atoms, charges.typed_data());
  if (cudaMemsetAsync(sums->typed_data(), 0, 3 * sizeof(float), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "jaqmc_b200: cudaMemsetAsync failed");
  return FromStatus(jaqmc_b200_local_energy(&d.wf, &sys, electrons.typed_data(), W, logpsi->typed_data(),
                                            sign->typed_data(), grad->typed_data(), lap->typed_data(), e_kin->typed_data(),
                                            e_pot->typed_data(), e_loc->typed_data(), sums->typed_data(), ws, bytes, stream));
}

// Ewald descriptor from operands precomputed by the host exactly as EwaldSum.__init__ does (estimator/ewald.py:50-110)
jaqmc_ewald MakeEwald(const F32& lattice, const F32& mic_shifts, const F32& images, const F32& gpoints, const F32& gweight,
                      ffi::Span<const float> consts, ffi::Span<const int32_t> iconsts) {
  jaqmc_ewald ew;
  std::memset(&ew, 0, sizeof(ew));
  ew.lattice = lattice.typed_data();
  ew.mic_shifts = mic_shifts.typed_data();
  ew.images = images.typed_data();
  ew.gpoints = gpoints.typed_data();
  ew.gweight = gweight.typed_data();
  ew.n_images = (int32_t)images.dimensions()[0];
  ew.n_g = (int32_t)gweight.dimensions()[0];
  ew.center_image = iconsts[0];
  ew.mic_kind = iconsts[1];
  ew.alpha = consts[0];
  ew.self_const_factor = consts[1];
  ew.ijconst = consts[2];
  return ew;
}

// ---- periodic network: complex local energy + Ewald (app/solid/hamiltonian.py:18-56) -----------------------------------
ffi::Error LocalEnergyComplexImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Span<const int32_t> config,
                                  ffi::Span<const float> fconfig, ffi::Span<const float> ewald_consts,
                                  ffi::Span<const int32_t> ewald_iconsts, F32 electrons, F32 prim_atoms, F32 cell_atoms,
                                  F32 cell_charges, F32 lattice, F32 mic_shifts, F32 images, F32 gpoints, F32 gweight,
                                  F32 klist, ffi::RemainingArgs rest, F32Out logpsi, F32Out grad, F32Out lap, F32Out e_kin,
                                  F32Out e_pot, F32Out e_loc, F32Out sums) {
  if (ewald_consts.size() != 3 || ewald_iconsts.size() != 2) return Invalid("jaqmc_b200: ewald constants");
  Descriptor d;
  if (auto e = Assemble(JAQMC_WF_SOLID_FERMINET, config, fconfig, 0, rest, 0, klist.typed_data(), &d); e.failure()) return e;
  int64_t W = 0;
  if (auto e = Walkers(electrons, d.n_electrons, &W); e.failure()) return e;
  void* ws = nullptr;
  size_t bytes = 0;
  if (auto e = Workspace(scratch, &d.wf, W, 1, &ws, &bytes); e.failure()) return e;
  jaqmc_system sys = System(prim_atoms, nullptr);
  jaqmc_ewald ew = MakeEwald(lattice, mic_shifts, images, gpoints, gweight, ewald_consts, ewald_iconsts);
  auto cd = cell_atoms.dimensions();
  if (cudaMemsetAsync(sums->typed_data(), 0, 3 * sizeof(float), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "jaqmc_b200: cudaMemsetAsync failed");
  return FromStatus(jaqmc_b200_local_energy_complex(
      &d.wf, &sys, &ew, cell_atoms.typed_data(), cell_charges.typed_data(), (int32_t)cd[cd.size() - 2],
      electrons.typed_data(), W, logpsi->typed_data(), grad->typed_data(), lap->typed_data(), e_kin->typed_data(),
      e_pot->typed_data(), e_loc->typed_data(), sums->typed_data(), ws, bytes, stream));
}

// ---- MCMCSampler.step's fori_loop of _mh_update (sampler/mcmc.py:96-137,167-180) with host-supplied noise ---------------
// lattice: empty -> gaussian_proposal; 9 floats -> make_pbc_gaussian_proposal(lattice) (geometry/pbc.py:187-201).
// electrons / logpsi are updated through input_output_aliases {electrons -> electrons_out}; XLA copies when not donated.
ffi::Error MhStepImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int32_t kind, ffi::Span<const int32_t> config,
                      ffi::Span<const float> fconfig, int32_t optional, ffi::Span<const float> lattice, F32 electrons,
                      F32 atoms, F32 normals, F32 uniforms, F32 stddev, ffi::RemainingArgs rest, F32Out electrons_out,
                      F32Out logpsi_out, F32Out n_accept, U8Out accepted) {
  Descriptor d;
  const bool solid = kind == JAQMC_WF_SOLID_FERMINET;
  const float* klist = nullptr;
  if (solid) {
    auto k = rest.get<F32>(0);
    if (!k.has_value()) return k.error();
    klist = k->typed_data();
  }
  if (auto e = Assemble(kind, config, fconfig, optional, rest, solid ? 1 : 0, klist, &d); e.failure()) return e;
  int64_t W = 0;
  if (auto e = Walkers(electrons, d.n_electrons, &W); e.failure()) return e;
  auto nd = normals.dimensions();
  if (nd.size() < 3) return Invalid("jaqmc_b200: normals must have shape [steps, ..., n, 3]");
  const int32_t steps = (int32_t)nd[0];
  if ((int64_t)normals.element_count() != (int64_t)steps * W * d.n_electrons * 3 ||
      (int64_t)uniforms.element_count() != (int64_t)steps * W)
    return Invalid("jaqmc_b200: normals / uniforms do not match [steps, walkers, n, 3] / [steps, walkers]");
  if (lattice.size() != 0 && lattice.size() != 9) return Invalid("jaqmc_b200: attribute `lattice` must be empty or 3x3");
  void* ws = nullptr;
  size_t bytes = 0;
  if (auto e = Workspace(scratch, &d.wf, W, 0, &ws, &bytes); e.failure()) return e;
  jaqmc_system sys = System(atoms, nullptr);
  float* x = electrons_out->typed_data();
  if (x != electrons.typed_data() &&
      cudaMemcpyAsync(x, electrons.typed_data(), electrons.size_bytes(), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "jaqmc_b200: cudaMemcpyAsync failed");
  if (cudaMemsetAsync(n_accept->typed_data(), 0, sizeof(float), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "jaqmc_b200: cudaMemsetAsync failed");
  if (lattice.size() == 9)
    return FromStatus(jaqmc_b200_mh_step_pbc(&d.wf, &sys, x, logpsi_out->typed_data(), 0, normals.typed_data(),
                                             uniforms.typed_data(), stddev.typed_data(), steps, W, n_accept->typed_data(),
                                             accepted->typed_data(), lattice.begin(), ws, bytes, stream));
  return FromStatus(jaqmc_b200_mh_step(&d.wf, &sys, x, logpsi_out->typed_data(), 0, normals.typed_data(),
                                       uniforms.typed_data(), stddev.typed_data(), steps, W, n_accept->typed_data(),
                                       accepted->typed_data(), ws, bytes, stream));
}

// ---- potential_energy (app/molecule/hamiltonian.py:9-22) ---------------------------------------------------------------
ffi::Error CoulombImpl(cudaStream_t stream, F32 electrons, F32 atoms, F32 charges, F32Out e_pot) {
  auto dims = electrons.dimensions();
  if (dims.size() < 2 || dims[dims.size() - 1] != 3) return Invalid("jaqmc_b200: electrons must have shape [..., n, 3]");
  const int32_t n = (int32_t)dims[dims.size() - 2];
  const int64_t W = (int64_t)electrons.element_count() / (3 * n);
  jaqmc_system sys = System(atoms, charges.typed_data());
  return FromStatus(jaqmc_b200_coulomb(&sys, electrons.typed_data(), W, n, e_pot->typed_data(), stream));
}

// ---- EwaldSum.energy / solid PotentialEnergy (estimator/ewald.py:112-173, app/solid/hamiltonian.py:28-56) ----------------
ffi::Error EwaldImpl(cudaStream_t stream, ffi::Span<const float> ewald_consts, ffi::Span<const int32_t> ewald_iconsts,
                     F32 electrons, F32 cell_atoms, F32 cell_charges, F32 lattice, F32 mic_shifts, F32 images, F32 gpoints,
                     F32 gweight, F32Out e_pot) {
  if (ewald_consts.size() != 3 || ewald_iconsts.size() != 2) return Invalid("jaqmc_b200: ewald constants");
  auto dims = electrons.dimensions();
  if (dims.size() < 2 || dims[dims.size() - 1] != 3) return Invalid("jaqmc_b200: electrons must have shape [..., n, 3]");
  const int32_t n = (int32_t)dims[dims.size() - 2];
  const int64_t W = (int64_t)electrons.element_count() / (3 * n);
  jaqmc_ewald ew = MakeEwald(lattice, mic_shifts, images, gpoints, gweight, ewald_consts, ewald_iconsts);
  auto cd = cell_atoms.dimensions();
  return FromStatus(jaqmc_b200_ewald(&ew, electrons.typed_data(), W, n, cell_atoms.typed_data(), cell_charges.typed_data(),
                                     (int32_t)cd[cd.size() - 2], e_pot->typed_data(), stream));
}

}  // namespace

#define JQ_WF_ATTRS()                                   \
  Attr<int32_t>("kind")                                 \
      .Attr<ffi::Span<const int32_t>>("config")         \
      .Attr<ffi::Span<const float>>("fconfig")          \
      .Attr<int32_t>("optional")

XLA_FFI_DEFINE_HANDLER_SYMBOL(jaqmc_b200_ffi_logpsi, LogPsiImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .JQ_WF_ATTRS()
                                  .Arg<F32>()   // electrons
                                  .Arg<F32>()   // atoms
                                  .RemainingArgs()
                                  .Ret<F32>()   // logpsi
                                  .Ret<F32>()); // sign

XLA_FFI_DEFINE_HANDLER_SYMBOL(jaqmc_b200_ffi_orbitals, OrbitalsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .JQ_WF_ATTRS()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .RemainingArgs()
                                  .Ret<F32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(jaqmc_b200_ffi_local_energy, LocalEnergyImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .JQ_WF_ATTRS()
                                  .Arg<F32>()   // electrons
                                  .Arg<F32>()   // atoms
                                  .Arg<F32>()   // charges
                                  .RemainingArgs()
                                  .Ret<F32>()   // logpsi
                                  .Ret<F32>()   // sign
                                  .Ret<F32>()   // grad
                                  .Ret<F32>()   // lap
                                  .Ret<F32>()   // e_kin
                                  .Ret<F32>()   // e_pot
                                  .Ret<F32>()   // e_loc
                                  .Ret<F32>()); // sums

XLA_FFI_DEFINE_HANDLER_SYMBOL(jaqmc_b200_ffi_local_energy_complex, LocalEnergyComplexImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Attr<ffi::Span<const int32_t>>("config")
                                  .Attr<ffi::Span<const float>>("fconfig")
                                  .Attr<ffi::Span<const float>>("ewald_consts")
                                  .Attr<ffi::Span<const int32_t>>("ewald_iconsts")
                                  .Arg<F32>()   // electrons
                                  .Arg<F32>()   // primitive atoms
                                  .Arg<F32>()   // cell atoms
                                  .Arg<F32>()   // cell charges
                                  .Arg<F32>()   // lattice
                                  .Arg<F32>()   // mic shifts
                                  .Arg<F32>()   // images
                                  .Arg<F32>()   // gpoints
                                  .Arg<F32>()   // gweight
                                  .Arg<F32>()   // klist
                                  .RemainingArgs()
                                  .Ret<F32>()   // logpsi (re, im)
                                  .Ret<F32>()   // grad
                                  .Ret<F32>()   // lap
                                  .Ret<F32>()   // e_kin
                                  .Ret<F32>()   // e_pot
                                  .Ret<F32>()   // e_loc
                                  .Ret<F32>()); // sums

XLA_FFI_DEFINE_HANDLER_SYMBOL(jaqmc_b200_ffi_mh_step, MhStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .JQ_WF_ATTRS()
                                  .Attr<ffi::Span<const float>>("lattice")
                                  .Arg<F32>()   // electrons
                                  .Arg<F32>()   // atoms
                                  .Arg<F32>()   // normals
                                  .Arg<F32>()   // uniforms
                                  .Arg<F32>()   // stddev
                                  .RemainingArgs()
                                  .Ret<F32>()   // electrons_out (aliased to electrons)
                                  .Ret<F32>()   // logpsi_out
                                  .Ret<F32>()   // n_accept
                                  .Ret<ffi::Buffer<ffi::U8>>());  // accepted

XLA_FFI_DEFINE_HANDLER_SYMBOL(jaqmc_b200_ffi_coulomb, CoulombImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Ret<F32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(jaqmc_b200_ffi_ewald, EwaldImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const float>>("ewald_consts")
                                  .Attr<ffi::Span<const int32_t>>("ewald_iconsts")
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Arg<F32>()
                                  .Ret<F32>());
