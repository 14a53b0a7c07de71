"""JAX-side plugin of jaqmc_b200: the classes a ``jaqmc molecule train`` run loads with ``wf.module=...`` so that the
reference's unchanged workflow (sampler, estimators, optimizers) reaches the B200 kernels through XLA-FFI custom calls.

    jaqmc molecule train wf.module=jaqmc_b200_jax.wavefunction:FermiNetWavefunction \\
        train.optim.module=jaqmc_b200_jax.kfac:kfac          # only when the optimizer is KFAC (graph tagging)

STATUS: written against jax==0.9.1 / flax==0.12.6 / the reference at /root/reference (the versions its requirements.txt
pins), but NOT executed in the build container, which has neither jax nor a way to install it; `tests/test_jax_plugin.py`
only byte-compiles these files and checks the pieces that do not need jax (attribute packing against the ctypes structs).
The kernels, the C ABI, the leaf order (`jaqmc_b200_bind_param_leaves`) and the XLA-FFI shim's signatures ARE tested; see
INTEGRATION.md for the bring-up checklist on a JAX box.

Layering (SURVEY.md Appendix C -- which transform reaches ``wf.logpsi`` and what each needs):

* ``MCMCSampler`` / ``SamplePlan``: ``vmap(logpsi)`` inside ``fori_loop`` -> ``ffi_call`` of ``jaqmc_b200_ffi_logpsi`` with
  ``vmap_method="expand_dims"`` (parameters are not replicated along the walker axis);
* ``EuclideanKinetic`` (forward Laplacian): ``logpsi`` is a ``custom_laplacian`` function whose rule -- valid for the
  identity Local1 seed ``EuclideanKinetic`` plants -- returns ``LapTuple(logpsi, grad, lap)`` from
  ``jaqmc_b200_ffi_local_energy``; any other seed raises ``AutoLaplacianFallback`` and the interpreter takes the Flax graph;
* ``LossAndGrad`` / SR: ``jax.custom_vjp`` whose backward pass is the VJP of the reference's Flax graph (the twin);
* KFAC: needs the plain Flax graph for its tag registration -> ``kfac.py`` hands it ``wf.twin_logpsi``.
"""

from ._config import pack_config  # noqa: F401
