"""CPU oracle for the JaQMC local-energy + sampling hot path.  TEST INFRASTRUCTURE ONLY.

This package is a float64 PyTorch-CPU / NumPy restatement of the reference algorithm
(bytedance/jaqmc, files cited per function).  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and
there only as the checker or the timed CPU baseline -- never by ``jaqmc_b200`` (the product),
which fails loudly when its CUDA library is missing.

Pinning status: the reference is pure Python on jax/flax, neither of which exists
in the build container or its wheelhouse, so the reference itself could not be run to generate
golden vectors.  The oracle is pinned instead against every *known-answer* check the reference's
own tests hold for this path:

* analytic Gaussian kinetic energy ``E_kin = c*N*d - 2 c^2 sum x^2``
  (``tests/estimator/kinetic_forward_laplacian_test.py:42-85``),
* Madelung constants NaCl -1.74756 (primitive + conventional cell) and CaF2 -5.03879
  (``tests/estimator/ewald_test.py:72-152``),
* slogdet forward-Laplacian rule against a full-Hessian contraction
  (``tests/laplacian/primitives/slogdet_test.py``, oracle ``tests/laplacian/helpers.py:45-125``),
* forward-Laplacian == brute-force Hessian trace on the real FermiNet / LapNet / Psiformer nets at the
  reference's fixed two-electron walker (``tests/estimator/kinetic_forward_laplacian_test.py:31-39,254-524``),
* antisymmetry sign flip (``tests/wavefunction/molecule_wavefunction_test.py:52-82``),
* LapNet attention == softmax formula (``:291-309``), PBC translation invariance
  (``tests/wavefunction/solid_test.py:112-137``), hydrogen-atom closed form.

``MCMCSampler`` and the molecular ``potential_energy`` have no direct reference test and no stored
vector anywhere in the reference: for those two functions parity is **unpinned** (restated from
``sampler/mcmc.py:96-137`` and ``app/molecule/hamiltonian.py:9-22`` only).
"""
