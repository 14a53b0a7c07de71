// tcgen05 / TMEM / TMA split-TF32 augmented-row GEMM (sm_100a).  Placeholder until the kernel lands:
// reports "not handled" so jq_launch_dense uses the CUDA-core kernel.
#include "aug.cuh"

int jq_launch_dense_tc(const JqDenseArgs& a, cudaStream_t st, bool* handled) {
  (void)a;
  (void)st;
  *handled = false;
  return JQ_OK;
}
