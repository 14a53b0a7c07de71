// jaqmc_b200 -- common definitions for the sm_100a kernels.
//
// Every non-tensor-core kernel in this library is written in "block-stride" style: a block's work is
// a sequence of phases separated by __syncthreads(), and inside each phase threads stride over the
// items with `for (i = threadIdx.x; i < N; i += blockDim.x)`.  The result therefore does not depend
// on blockDim, which gives two things: launch shapes can be tuned freely, and the very same source
// compiles with g++ (-DJAQMC_HOST_EMU) into a one-thread-per-block host build that the CPU test-suite
// uses to check the kernel arithmetic against the float64 oracle before any GPU time is spent.
// The host build is TEST INFRASTRUCTURE (tests/emu/): the product library is the nvcc build only, and
// the Python package refuses to load anything else.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>

extern thread_local long long jq_launch_counter;

#ifdef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// host emulation: one "thread" per block, blocks run sequentially
// ------------------------------------------------------------------------------------------------
struct jq_dim3 {
  unsigned x, y, z;
  jq_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef jq_dim3 dim3;
struct float4 {
  float x, y, z, w;
};
typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
extern thread_local jq_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local unsigned char* jq_emu_dyn_smem;
void jq_emu_set_smem(size_t bytes);
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __syncthreads() ((void)0)
#define JQ_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(jq_emu_dyn_smem)
using std::isfinite;
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline void sincosf_(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
// (a0, a1) += s * (y0, y1): two fmaf on the host; one packed FFMA2 on the device (below)
static inline void jq_fma2(float& a0, float& a1, float s, float y0, float y1) {
  a0 = fmaf(s, y0, a0);
  a1 = fmaf(s, y1, a1);
}
static inline const char* cudaGetErrorString(int) { return "emu"; }
static inline int cudaGetLastError() { return 0; }
static inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline int cudaMemcpyAsyncD2D(void* d, const void* s, size_t n, cudaStream_t) { memcpy(d, s, n); return 0; }

#define JQ_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
  do {                                                                          \
    ++jq_launch_counter;                                                        \
    jq_dim3 g_ = (grid);                                                        \
    gridDim = g_;                                                               \
    blockDim = jq_dim3(1, 1, 1);                                                \
    threadIdx = jq_dim3(0, 0, 0);                                               \
    jq_emu_set_smem(smem);                                                      \
    for (unsigned bz_ = 0; bz_ < g_.z; ++bz_)                                   \
      for (unsigned by_ = 0; by_ < g_.y; ++by_)                                 \
        for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                             \
          blockIdx = jq_dim3(bx_, by_, bz_);                                    \
          kernel(__VA_ARGS__);                                                  \
        }                                                                       \
  } while (0)

#else
// ------------------------------------------------------------------------------------------------
// device build
// ------------------------------------------------------------------------------------------------
#include <cuda_runtime.h>
#define JQ_DYN_SMEM(type, name)                                  \
  extern __shared__ __align__(16) unsigned char jq_dyn_smem_[]; \
  type* name = reinterpret_cast<type*>(jq_dyn_smem_)
static __device__ __forceinline__ void sincosf_(float x, float* s, float* c) { sincosf(x, s, c); }
// (a0, a1) += s * (y0, y1) as ONE packed instruction (sm_100: fma.rn.f32x2 -> FFMA2 with a broadcast scalar operand; ptxas
// folds the mov.b64 packing into register-pair allocation when y0 / y1 and a0 / a1 are adjacent, e.g. halves of a float4).
// Two IEEE round-to-nearest FMAs: bit-identical to two fmaf, at half the issue slots -- for the issue-bound CUDA-core
// contractions (r2, late).  -DJQ_NO_FFMA2 restores the scalar form (A/B).
static __device__ __forceinline__ void jq_fma2(float& a0, float& a1, float s, float y0, float y1) {
#ifdef JQ_NO_FFMA2
  a0 = fmaf(s, y0, a0);
  a1 = fmaf(s, y1, a1);
#else
  unsigned long long ss, yy, zz;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(yy) : "f"(y0), "f"(y1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(zz) : "f"(a0), "f"(a1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(zz) : "l"(ss), "l"(yy));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(zz));
#endif
}
static inline cudaError_t cudaMemcpyAsyncD2D(void* d, const void* s, size_t n, cudaStream_t st) {
  return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st);
}
// One-off per-DEVICE setup (cudaFuncSetAttribute opt-ins, SM count): function attributes belong to the device that is
// current when they are set, so a process that drives several GPUs must repeat them per device.  Flags are written
// after the setup they guard and only ever go false -> true, so a race between two threads repeats the (idempotent)
// setup at worst.
#define JQ_MAX_DEVICES 64
static inline int jq_current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < JQ_MAX_DEVICES) ? dev : 0;
}
struct JqPerDeviceFlag {
  volatile bool done[JQ_MAX_DEVICES];
};
static inline int jq_sm_count() {
  static volatile int count[JQ_MAX_DEVICES];
  const int dev = jq_current_device();
  if (!count[dev]) {
    int c = 0;
    cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
    count[dev] = c;
  }
  return count[dev];
}
// Optional per-kernel device timing (bench.py's roofline leg): when enabled, every launch is bracketed by
// two cudaEvents on the launching stream and tagged with the work the launcher declared via jq_prof_work().
void jq_prof_before(const char* name, cudaStream_t st);
void jq_prof_after(cudaStream_t st);
extern thread_local int jq_prof_enabled;
#define JQ_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
  do {                                                                          \
    ++jq_launch_counter;                                                        \
    if (jq_prof_enabled) jq_prof_before(#kernel, (stream));                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                 \
    if (jq_prof_enabled) jq_prof_after((stream));                               \
  } while (0)
#endif
// Declares the algorithmic work (flops, bytes) of the NEXT launch for the profiler; no-op when disabled.
void jq_prof_work(double flops, double bytes);

// ------------------------------------------------------------------------------------------------
// status codes returned through the C ABI (0 == ok); see include/jaqmc_b200.h
// ------------------------------------------------------------------------------------------------
#define JQ_OK 0
#define JQ_ERR_INVALID_ARGUMENT 1
#define JQ_ERR_WORKSPACE_TOO_SMALL 2
#define JQ_ERR_CUDA 3
#define JQ_ERR_UNSUPPORTED 4

void jq_set_error(const char* fmt, ...);
extern thread_local long long jq_launch_counter;

#define JQ_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      jq_set_error(__VA_ARGS__);     \
      return (code);                 \
    }                                \
  } while (0)

#define JQ_CHECK_LAUNCH()                                                            \
  do {                                                                               \
    cudaError_t e_ = (cudaError_t)cudaGetLastError();                                \
    if (e_ != cudaSuccess) {                                                         \
      jq_set_error("CUDA launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                   cudaGetErrorString(e_));                                          \
      return JQ_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

static inline int jq_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t jq_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-owned workspace (XLA owns every buffer; the library never mallocs).
struct JqArena {
  unsigned char* base;
  size_t cap, off;
  bool dry;  // dry run: only measure
  JqArena(void* p, size_t c) : base((unsigned char*)p), cap(c), off(0), dry(p == nullptr) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = jq_align_up(count * sizeof(T), 256);
    size_t o = off;
    off += bytes;
    if (dry) return nullptr;
    return reinterpret_cast<T*>(base + o);
  }
  bool ok() const { return dry || off <= cap; }
};
