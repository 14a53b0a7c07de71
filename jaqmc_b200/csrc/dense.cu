// Augmented-row dense layers (CUDA-core path), tanh forward-Laplacian epilogue, and the FermiNet
// aggregation kernels.
//
// Reference semantics:
//   * nn.Dense under forward_laplacian = one GEMM over the rows {x, J_1..J_K, L}; bias on the value row
//     only (laplacian/primitives/dot_general.py:377-407)
//   * tanh rule y_J = (1-y^2) J, y_L = (1-y^2) L - 2y(1-y^2) sum_k J_k^2 (primitives/elementwise.py:42-72)
//   * FermiNet aggregate_features / residual (wavefunction/backbone/ferminet.py:51-90)
//
// The CUDA-core GEMM below serves the narrow layers (contraction width < 64: input layers, the
// two-electron stream) and is the numerical cross-check of the tcgen05 kernel in dense_tc.cu, which takes
// the wide (>= 64-deep) layers on device builds.
#include <cstdlib>

#include "aug.cuh"

// ------------------------------------------------------------------------------------------------
// GEMM: out[row][col] = sum_k [src0|src1][row][k] * [w0;w1][k][col] (+cadd) (+bias on value rows)
// ------------------------------------------------------------------------------------------------
#define GM_BM 128
#define GM_BN 64
#define GM_BK 16

__device__ __forceinline__ long long jq_group_of(long long gsub, int n_sub, int n_tot, int j0) {
  return (gsub / n_sub) * (long long)n_tot + j0 + (gsub % n_sub);
}

#ifndef JAQMC_HOST_EMU
__global__ void __launch_bounds__(256) k_dense_simt(JqDenseArgs a) {
  __shared__ float As[GM_BK][GM_BM + 4];
  __shared__ float Bs[GM_BK][GM_BN];
  const long long R = a.G * a.C;
  const long long row0 = (long long)blockIdx.x * GM_BM;
  const int col0 = blockIdx.y * GM_BN;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int kt = a.k0 + a.k1;
  const int ldw = a.ldw ? a.ldw : a.N;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // each thread loads 8 A elements: element q -> (row = (tid + q*256)/16, k = (tid + q*256)%16)
  long long arow[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    long long r = row0 + (tid + q * 256) / GM_BK;
    if (r < R) {
      long long gs = r / a.C;
      int c = (int)(r % a.C);
      arow[q] = jq_group_of(gs, a.n_sub, a.n_tot, a.j0) * a.C + c;
    } else {
      arow[q] = -1;
    }
  }
  for (int kb = 0; kb < kt; kb += GM_BK) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int idx = tid + q * 256;
      int rl = idx / GM_BK, kl = idx % GM_BK;
      int kk = kb + kl;
      float v = 0.f;
      if (arow[q] >= 0 && kk < kt)
        v = (kk < a.k0) ? a.src0[arow[q] * a.k0 + kk] : a.src1[arow[q] * a.k1 + (kk - a.k0)];
      As[kl][rl] = v;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int idx = tid + q * 256;
      int kl = idx / GM_BN, cl = idx % GM_BN;
      int kk = kb + kl, col = col0 + cl;
      float v = 0.f;
      if (kk < kt && col < a.N)
        v = (kk < a.k0) ? ((a.k0_valid == 0 || kk < a.k0_valid) ? a.w0[(long long)kk * ldw + col] : 0.f)
                        : a.w1[(long long)(kk - a.k0) * ldw + col];
      Bs[kl][cl] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GM_BK; ++k) {
      float av[8], bv[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = As[k][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long r = row0 + ty * 8 + i;
    if (r >= R) continue;
    long long gs = r / a.C;
    int c = (int)(r % a.C);
    long long g = jq_group_of(gs, a.n_sub, a.n_tot, a.j0);
    long long orow = g * a.C + c;
    long long w = g / a.n_tot;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = col0 + tx * 4 + j;
      if (col >= a.N) continue;
      float v = acc[i][j];
      if (a.cadd) v += a.cadd[(w * a.C + c) * a.N + col];
      if (a.bias && c == 0) v += a.bias[col];
      if (a.act == 0 && a.res_mode) {  // linear layer with a residual connection (no activation)
        float r = a.res[orow * a.N + col];
        v = (a.res_mode == 1) ? (r + v) * 0.70710678118654752440f : r + v;
      }
      a.out[orow * a.N + col] = v;
    }
  }
}
#else
__global__ void k_dense_simt(JqDenseArgs a) {
  const long long R = a.G * a.C;
  const int ldw = a.ldw ? a.ldw : a.N;
  for (long long r = (long long)blockIdx.x * GM_BM; r < R && r < (long long)(blockIdx.x + 1) * GM_BM; ++r) {
    long long gs = r / a.C;
    int c = (int)(r % a.C);
    long long g = jq_group_of(gs, a.n_sub, a.n_tot, a.j0);
    long long row = g * a.C + c;
    long long w = g / a.n_tot;
    for (int col = blockIdx.y * GM_BN; col < a.N && col < (int)(blockIdx.y + 1) * GM_BN; ++col) {
      float v = 0.f;
      const int kv = a.k0_valid ? a.k0_valid : a.k0;   // src0 columns [kv, k0) are zero padding without kernel rows
      for (int k = 0; k < kv; ++k) v = fmaf(a.src0[row * a.k0 + k], a.w0[(long long)k * ldw + col], v);
      for (int k = 0; k < a.k1; ++k) v = fmaf(a.src1[row * a.k1 + k], a.w1[(long long)k * ldw + col], v);
      if (a.cadd) v += a.cadd[(w * a.C + c) * a.N + col];
      if (a.bias && c == 0) v += a.bias[col];
      if (a.act == 0 && a.res_mode) {
        float r = a.res[row * a.N + col];
        v = (a.res_mode == 1) ? (r + v) * 0.70710678118654752440f : r + v;
      }
      a.out[row * a.N + col] = v;
    }
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// Narrow layers with few components per group (FermiNet's two-electron stream: Local2, C = 8, 4/32 -> 32 features;
// Local1 layers, C = 5): one fused kernel per layer.  A block stages a tile of SM_GT groups [C][k0] in shared memory
// with coalesced loads; each item owns one (group, output feature) with all C components in registers, so the bias,
// the tanh forward-Laplacian rule and the residual are private to the item, and the tile doubles as the residual
// source when the layer is square.  HBM traffic = read x once + write out once.
// ------------------------------------------------------------------------------------------------
#define SM_GT 32
#define SM_CMAX 8
// bias + tanh forward-Laplacian rule + residual for one (group, feature) item of k_dense_small
__device__ __forceinline__ void small_epilogue(const JqDenseArgs& a, float (&ac)[8], int f, long long g, const float* xr,
                                               bool res_from_tile) {
  const float inv_sqrt2 = 0.70710678118654752440f;
  const int C = a.C, K = a.k0, N = a.N;
  if (a.bias) ac[0] += a.bias[f];
  if (a.act == 1) {
    const float t = tanhf(ac[0]);
    const float d1 = 1.0f - t * t;
    float s2 = 0.f;
#pragma unroll
    for (int c = 1; c < 7; ++c)
      if (c < C - 1) {
        s2 = fmaf(ac[c], ac[c], s2);
        ac[c] *= d1;
      }
    if (C > 1) {
#pragma unroll
      for (int c = 1; c < 8; ++c)
        if (c == C - 1) ac[c] = d1 * ac[c] - 2.0f * t * d1 * s2;
    }
    ac[0] = t;
  }
  float* o = a.out + (g * C) * N + f;
  const float* rg = a.res ? a.res + (g * C) * N + f : nullptr;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < C) {
      float v = ac[c];
      if (a.res_mode) {
        const float r = res_from_tile ? xr[c * K + f] : rg[(long long)c * N];
        v = (a.res_mode == 1) ? (r + v) * inv_sqrt2 : r + v;
      }
      o[(long long)c * N] = v;
    }
}

// TC / TK / TN fix the components, contraction depth and width at compile time (0 = runtime value): with constants
// the address arithmetic folds into immediate offsets and the loops unroll, which is what bounds this kernel.
template <int TC, int TK, int TN>
__global__ void k_dense_small(JqDenseArgs a) {
  JQ_DYN_SMEM(float, sm);
  const int C = TC ? TC : a.C, K = TK ? TK : a.k0, N = TN ? TN : a.N;
  const int ldw = a.ldw ? a.ldw : N;
  float* Ws = sm;                 // [K][N]
  float* Xs = Ws + K * N;         // [GT][C][K]
  const int GT = a.small_gt;       // groups per block: SM_GT for 8-component groups, more for value-only launches
  const long long g0 = (long long)blockIdx.x * GT;
  const int ng = (int)((a.G - g0 < GT) ? a.G - g0 : GT);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int q = tid; q < K * N; q += nt) Ws[q] = a.w0[(long long)(q / N) * ldw + (q % N)];
  const float* xg = a.src0 + g0 * C * K;
  for (int q = tid; q < ng * C * K; q += nt) Xs[q] = xg[q];
  __syncthreads();
  const float inv_sqrt2 = 0.70710678118654752440f;
  const bool res_from_tile = (a.res == a.src0) && (K == N);
  // an item owns NF output features of one group: f and f + N/2 when N is even (the tile reads are shared by both)
  const int NF = (N % 2 == 0) ? 2 : 1;
  const int NH = N / NF;
  const bool vec4 = (K % 4 == 0);
  for (int q = tid; q < ng * NH; q += nt) {
    const int gl = q / NH, f0 = q % NH;
    const float* xr = Xs + gl * C * K;
    float acc0[SM_CMAX], acc1[SM_CMAX];
#pragma unroll
    for (int c = 0; c < SM_CMAX; ++c) acc0[c] = acc1[c] = 0.f;
    if (vec4) {
#pragma unroll
      for (int k = 0; k < K; k += 4) {
        float w0[4], w1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          w0[u] = Ws[(k + u) * N + f0];
          w1[u] = (NF == 2) ? Ws[(k + u) * N + f0 + NH] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < SM_CMAX; ++c)
          if (c < C) {
            const float4 x4 = *reinterpret_cast<const float4*>(xr + c * K + k);
            acc0[c] = fmaf(x4.x, w0[0], acc0[c]);
            acc0[c] = fmaf(x4.y, w0[1], acc0[c]);
            acc0[c] = fmaf(x4.z, w0[2], acc0[c]);
            acc0[c] = fmaf(x4.w, w0[3], acc0[c]);
            acc1[c] = fmaf(x4.x, w1[0], acc1[c]);
            acc1[c] = fmaf(x4.y, w1[1], acc1[c]);
            acc1[c] = fmaf(x4.z, w1[2], acc1[c]);
            acc1[c] = fmaf(x4.w, w1[3], acc1[c]);
          }
      }
    } else {
      for (int k = 0; k < K; ++k) {
        const float w0 = Ws[k * N + f0];
        const float w1 = (NF == 2) ? Ws[k * N + f0 + NH] : 0.f;
#pragma unroll
        for (int c = 0; c < SM_CMAX; ++c)
          if (c < C) {
            const float x = xr[c * K + k];
            acc0[c] = fmaf(x, w0, acc0[c]);
            acc1[c] = fmaf(x, w1, acc1[c]);
          }
      }
    }
    small_epilogue(a, acc0, f0, g0 + gl, xr, res_from_tile);
    if (NF == 2) small_epilogue(a, acc1, f0 + NH, g0 + gl, xr, res_from_tile);
  }
}

static bool dense_small_eligible(const JqDenseArgs& a) {
  return a.C <= SM_CMAX && a.k1 == 0 && a.k0_valid == 0 && a.cadd == nullptr && a.n_sub == a.n_tot && a.k0 <= 64 && a.N <= 64 &&
         a.out != a.src0;
}

#ifndef JAQMC_HOST_EMU
int jq_launch_dense_tc(const JqDenseArgs& a, cudaStream_t st, bool* handled);  // dense_tc.cu
#endif

int jq_launch_tanh_fl_mapped(const JqDenseArgs& a, cudaStream_t st);
#ifdef JAQMC_HOST_EMU
size_t jq_dense_tc_scratch_floats(int k_total, int n_out) { return (size_t)2 * k_total * n_out; }
bool jq_dense_tc_eligible(const JqDenseArgs&) { return false; }
#endif

#ifndef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// Value-only narrow layer with 32 output features (the FermiNet two-electron stream on the sampling path: 11 of the
// 12 forward passes of a VMC iteration).  k_dense_small re-reads the weights from shared memory for every
// multiply-add when a group has a single component; here a lane IS an output feature and keeps its weight column in
// registers, a warp stages 32 groups with coalesced 16-byte loads and every lane reads a group's inputs as float4
// broadcasts: 8 shared-memory reads per 32 multiply-adds, one coalesced 128-byte store per group.
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) k_dense_small_value(JqDenseArgs a) {
  __shared__ float4 tile_all[8][32 * K / 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* tile = tile_all[warp];
  const float* tile_f = reinterpret_cast<const float*>(tile);
  const int ldw = a.ldw ? a.ldw : 32;
  float wreg[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wreg[k] = a.w0[(long long)k * ldw + lane];
  const float bias = a.bias ? a.bias[lane] : 0.f;
  const bool res_from_tile = (a.res == a.src0) && (K == 32);
  const float inv_sqrt2 = 0.70710678118654752440f;
  const float4* src4 = reinterpret_cast<const float4*>(a.src0);
  for (long long g0 = ((long long)blockIdx.x * 8 + warp) * 32; g0 < a.G; g0 += (long long)gridDim.x * 8 * 32) {
    const int ng = (int)((a.G - g0 < 32) ? a.G - g0 : 32);
    for (int q = lane; q < ng * (K / 4); q += 32) tile[q] = src4[g0 * (K / 4) + q];
    __syncwarp();
    for (int gl = 0; gl < ng; ++gl) {
      float acc = bias;
#pragma unroll
      for (int k = 0; k < K; k += 4) {
        const float4 x = tile[gl * (K / 4) + k / 4];
        acc = fmaf(x.x, wreg[k], acc);
        acc = fmaf(x.y, wreg[k + 1], acc);
        acc = fmaf(x.z, wreg[k + 2], acc);
        acc = fmaf(x.w, wreg[k + 3], acc);
      }
      if (a.act == 1) acc = tanhf(acc);
      if (a.res_mode) {
        const float r = res_from_tile ? tile_f[gl * K + lane] : a.res[(g0 + gl) * 32 + lane];
        acc = (a.res_mode == 1) ? (r + acc) * inv_sqrt2 : r + acc;
      }
      a.out[(g0 + gl) * 32 + lane] = acc;
    }
    __syncwarp();
  }
}
#endif

static int dense_small_launch(const JqDenseArgs& a, cudaStream_t st);
static int dense_generic_launch(const JqDenseArgs& a, cudaStream_t st);

int jq_launch_dense(const JqDenseArgs& a, cudaStream_t st) {
  if (a.G <= 0 || a.N <= 0) return JQ_OK;
#ifdef JAQMC_HOST_EMU
  JQ_REQUIRE(a.act == 0 || a.act == 1, JQ_ERR_INVALID_ARGUMENT, "dense: unknown activation %d", a.act);
#else
  JQ_REQUIRE(a.act == 0 || a.act == 1 || (a.act == 2 && a.env && jq_dense_tc_eligible(a)), JQ_ERR_INVALID_ARGUMENT,
             "dense: unknown activation %d", a.act);
#endif
  JQ_REQUIRE(a.res_mode == 0 || a.res != nullptr, JQ_ERR_INVALID_ARGUMENT, "dense: residual without source");
  JQ_REQUIRE(a.k0 > 0 && a.src0 && a.w0 && a.out, JQ_ERR_INVALID_ARGUMENT, "dense: null operand");
  JQ_REQUIRE(a.k1 == 0 || (a.src1 && a.w1), JQ_ERR_INVALID_ARGUMENT, "dense: null second operand");
#ifndef JAQMC_HOST_EMU
  bool handled = false;
  int rc = jq_launch_dense_tc(a, st, &handled);
  if (rc != JQ_OK) return rc;
  if (handled) return JQ_OK;
  if (jq_prep.collect) return JQ_OK;   // dry pass of the weight-split cache: CUDA-core launches have nothing to record
#endif
  if (dense_small_eligible(a)) {
    // ~256 rows per block: with one component per group (sampling path) a 32-group tile would be smaller than the
    // weights every block stages
    JqDenseArgs b = a;
    b.small_gt = (SM_GT * SM_CMAX) / a.C;
    if (b.small_gt < SM_GT) b.small_gt = SM_GT;
    return dense_small_launch(b, st);
  }
  return dense_generic_launch(a, st);
}

static int dense_small_launch(const JqDenseArgs& a, cudaStream_t st) {
#ifndef JAQMC_HOST_EMU
  if (a.C == 1 && a.N == 32 && (a.k0 == 4 || a.k0 == 32) && (reinterpret_cast<uintptr_t>(a.src0) & 15) == 0 &&
      a.out != a.src0) {
    jq_prof_work(2.0 * (double)a.G * a.k0 * a.N, 4.0 * (double)a.G * (a.k0 + a.N));
    long long blocks = jq_cdiv(a.G, 8 * 32);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (a.k0 == 4) JQ_LAUNCH(k_dense_small_value<4>, dim3((unsigned)blocks), dim3(256), 0, st, a);
    else JQ_LAUNCH(k_dense_small_value<32>, dim3((unsigned)blocks), dim3(256), 0, st, a);
    JQ_CHECK_LAUNCH();
    return JQ_OK;
  }
#endif
  {
    size_t smem = sizeof(float) * ((size_t)a.k0 * a.N + (size_t)a.small_gt * a.C * a.k0);
    jq_prof_work(2.0 * (double)a.G * a.C * a.k0 * a.N, 4.0 * (double)a.G * a.C * (a.k0 + a.N));
#ifndef JAQMC_HOST_EMU
    static JqPerDeviceFlag attr_set;
    const int dev = jq_current_device();
    if (!attr_set.done[dev]) {
      cudaFuncSetAttribute(k_dense_small<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr_set.done[dev] = true;
    }
#endif
    const dim3 grid((unsigned)jq_cdiv(a.G, a.small_gt));
    // the FermiNet two-electron stream (Local2: 8 components; value path: 1) and Local1 layers at the default widths
    if (a.C == 8 && a.k0 == 32 && a.N == 32) JQ_LAUNCH((k_dense_small<8, 32, 32>), grid, dim3(256), smem, st, a);
    else if (a.C == 8 && a.k0 == 4 && a.N == 32) JQ_LAUNCH((k_dense_small<8, 4, 32>), grid, dim3(256), smem, st, a);
    else if (a.C == 1 && a.k0 == 32 && a.N == 32) JQ_LAUNCH((k_dense_small<1, 32, 32>), grid, dim3(256), smem, st, a);
    else if (a.C == 1 && a.k0 == 4 && a.N == 32) JQ_LAUNCH((k_dense_small<1, 4, 32>), grid, dim3(256), smem, st, a);
    else if (a.C == 8 && a.k0 == 7 && a.N == 32) JQ_LAUNCH((k_dense_small<8, 7, 32>), grid, dim3(256), smem, st, a);
    else JQ_LAUNCH((k_dense_small<0, 0, 0>), grid, dim3(256), smem, st, a);
    JQ_CHECK_LAUNCH();
    return JQ_OK;
  }
}

static int dense_generic_launch(const JqDenseArgs& a, cudaStream_t st) {
  long long R = a.G * a.C;
  dim3 grid(jq_cdiv(R, GM_BM), jq_cdiv(a.N, GM_BN));
  jq_prof_work(2.0 * (double)R * (a.k0 + a.k1) * a.N, 4.0 * (double)R * (a.k0 + a.k1 + a.N));
  JQ_LAUNCH(k_dense_simt, grid, dim3(256), 0, st, a);
  JQ_CHECK_LAUNCH();
  if (a.act == 1) return jq_launch_tanh_fl_mapped(a, st);  // unfused epilogue, in place on `out`
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// tanh with the forward-Laplacian rule (+ FermiNet residual).  One item per (group, feature).
//   residual_mode 0: out = tanh(y)     1: out = (res + tanh(y)) / sqrt(2)     2: out = res + tanh(y)
// ------------------------------------------------------------------------------------------------
__global__ void k_tanh_fl(const float* y, const float* res, float* out, long long items, int C, int F, int mode,
                          int n_sub, int n_tot, int j0) {
  const float inv_sqrt2 = 0.70710678118654752440f;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    long long g = jq_group_of(it / F, n_sub, n_tot, j0);
    int f = (int)(it % F);
    const float* yp = y + g * C * F + f;
    const float* rp = res ? res + g * C * F + f : nullptr;
    float* op = out + g * C * F + f;
    float t = tanhf(yp[0]);
    float d1 = 1.0f - t * t;
    float v = t;
    if (mode == 1) v = (rp[0] + t) * inv_sqrt2;
    else if (mode == 2) v = rp[0] + t;
    if (C > 1) {
      float s2 = 0.f;
      for (int c = 1; c < C - 1; ++c) {
        float j = yp[(long long)c * F];
        s2 = fmaf(j, j, s2);
        float o = d1 * j;
        if (mode == 1) o = (rp[(long long)c * F] + o) * inv_sqrt2;
        else if (mode == 2) o = rp[(long long)c * F] + o;
        op[(long long)c * F] = o;
      }
      float l = d1 * yp[(long long)(C - 1) * F] - 2.0f * t * d1 * s2;
      if (mode == 1) l = (rp[(long long)(C - 1) * F] + l) * inv_sqrt2;
      else if (mode == 2) l = rp[(long long)(C - 1) * F] + l;
      op[(long long)(C - 1) * F] = l;
    }
    op[0] = v;
  }
}

int jq_launch_tanh_fl(const float* y, const float* res, float* out, long long G, int C, int F, int residual_mode,
                      cudaStream_t st) {
  long long items = G * F;
  if (items <= 0) return JQ_OK;
  JQ_REQUIRE(residual_mode == 0 || res != nullptr, JQ_ERR_INVALID_ARGUMENT, "tanh_fl: residual without source");
  int grid = jq_cdiv(items, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  jq_prof_work(0.0, 4.0 * (double)G * C * F * (res ? 3 : 2));
  JQ_LAUNCH(k_tanh_fl, dim3(grid), dim3(256), 0, st, y, res, out, items, C, F, residual_mode, 1, 1, 0);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// epilogue of an unfused dense launch: in place on a.out, honouring the launch's group mapping
int jq_launch_tanh_fl_mapped(const JqDenseArgs& a, cudaStream_t st) {
  long long items = a.G * a.N;
  int grid = jq_cdiv(items, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  jq_prof_work(0.0, 4.0 * (double)a.G * a.C * a.N * (a.res ? 3 : 2));
  JQ_LAUNCH(k_tanh_fl, dim3(grid), dim3(256), 0, st, a.out, a.res, a.out, items, a.C, a.N, a.res_mode, a.n_sub, a.n_tot,
            a.j0);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// FermiNet pair-stream spin means: h2 [W][n*n][C2][d2] (Local2, pair p = i*n + j) -> g2 [W][n][C][nch*d2]
// g2[j][s] = mean_{i in s} h2[i][j]  (mean over the FIRST electron axis, ferminet.py:86-89).  This is the
// Local2 -> dense transition of the reference (laplacian/primitives/reductions.py:52-61).
// One item per (walker, j, s, feature).
// ------------------------------------------------------------------------------------------------
__global__ void k_pair_mean(const float* __restrict__ h2, float* __restrict__ g2, long long items, JqSpins sp,
                            int d2, int track) {
  const int n = sp.n(), nch = sp.nch();
  const int C2 = track ? 8 : 1, C = track ? 3 * n + 2 : 1, FO = nch * d2;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    int f = (int)(it % d2);
    long long t = it / d2;
    int s = (int)(t % nch);
    t /= nch;
    int j = (int)(t % n);
    long long w = t / n;
    const int lo = sp.lo(s), hi = sp.hi(s);
    const float inv = 1.0f / (float)(hi - lo);
    const float* base = h2 + (w * n * n) * (long long)C2 * d2 + f;  // pair p comp c at base + (p*C2 + c)*d2
    float* o = g2 + ((w * n + j) * (long long)C) * FO + s * d2 + f;  // comp c at o + c*FO
    float sx = 0.f, sl = 0.f, sj[3] = {0.f, 0.f, 0.f};
    for (int i = lo; i < hi; ++i) {
      const float* p = base + ((long long)(i * n + j) * C2) * d2;
      sx += p[0];
      if (track) {
        sl += p[7 * d2];
        sj[0] += p[4 * d2];
        sj[1] += p[5 * d2];
        sj[2] += p[6 * d2];
      }
    }
    o[0] = sx * inv;
    if (track) {
      o[(long long)(C - 1) * FO] = sl * inv;
      for (int e = 0; e < n; ++e) {
        bool in_s = (e >= lo && e < hi);
        const float* p = base + ((long long)(e * n + j) * C2) * d2;
        for (int a = 0; a < 3; ++a) {
          float v = in_s ? p[(1 + a) * d2] : 0.f;
          if (e == j) v += sj[a];
          o[(long long)(1 + 3 * e + a) * FO] = v * inv;
        }
      }
    }
  }
}

#ifndef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// FermiNet two-electron stream, one layer fused with the spin-channel means that feed the next one-electron layer
// (wavefunction/backbone/ferminet.py:57-63 followed by aggregate_features :82-90 of the next layer):
//     h2'[i,j] = res( tanh-FL( h2[i,j] . K + b ) ),      g2'[j] = mean_{i in channel} h2'[i,j]   (Local2 -> dense)
// k_dense_small + k_pair_mean read / write the pair tensor three times; here one WARP owns a column j of a walker
// (the n pairs (i, j)), a lane is an output feature (32 of them) with its weight column in registers, and the means
// are accumulated in registers while the pairs stream through, so the pair tensor is read once and written once --
// or not written at all for the last two-electron layer, whose output is only ever used through its means.
// Local2 components: 0 value, 1-3 d/dr_i, 4-6 d/dr_j, 7 Laplacian.  g2' component of electron e, axis a:
//     ([e in channel] h2'[e,j][1+a] + [e == j] sum_{i in channel} h2'[i,j][4+a]) / |channel|     (k_pair_mean).
// ------------------------------------------------------------------------------------------------
template <int K, int C2>   // C2 = 8: Local2 components (forward Laplacian); C2 = 1: value only (sampling path)
__global__ void __launch_bounds__(256) k_pair_layer_fused(const float* __restrict__ h2, const float* __restrict__ w0,
                                                         const float* __restrict__ bias, float* __restrict__ h2n,
                                                         float* __restrict__ g2, long long WJ, JqSpins sp, int residual) {
  __shared__ float4 tile_all[8][C2 * K / 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* tile = tile_all[warp];
  const float* tile_f = reinterpret_cast<const float*>(tile);
  const int n = sp.n(), nch = sp.nch();
  const int C = (C2 == 8) ? 3 * n + 2 : 1, FO = nch * 32;
  float wreg[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wreg[k] = w0[k * 32 + lane];
  const float b = bias ? bias[lane] : 0.f;
  const float inv_sqrt2 = 0.70710678118654752440f;
  const float4* src4 = reinterpret_cast<const float4*>(h2);
  constexpr int NV = (C2 * K / 4 + 31) / 32;   // float4 per lane of one pair tile
  for (long long wj = (long long)blockIdx.x * 8 + warp; wj < WJ; wj += (long long)gridDim.x * 8) {
    const long long w = wj / n;
    const int j = (int)(wj - w * n);
    // per-channel sums kept in scalar registers (no dynamically indexed arrays): channel 0 / channel 1
    float sx0 = 0.f, sx1 = 0.f, sl0 = 0.f, sl1 = 0.f, sj0[3] = {0.f, 0.f, 0.f}, sj1[3] = {0.f, 0.f, 0.f};
    float own[3] = {0.f, 0.f, 0.f};
    float* o = g2 + (wj * C) * FO + lane;   // component c, channel s at o[c * FO + s * 32]
    // the next pair's tile travels in registers while the current one is processed
    float4 nxt[NV];
    {
      const long long pair = (w * n) * n + j;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (lane + 32 * v < C2 * K / 4) nxt[v] = src4[pair * (C2 * K / 4) + lane + 32 * v];
    }
    for (int i = 0; i < n; ++i) {
      const long long pair = (w * n + i) * n + j;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (lane + 32 * v < C2 * K / 4) tile[lane + 32 * v] = nxt[v];
      if (i + 1 < n) {
        const long long pn = pair + n;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (lane + 32 * v < C2 * K / 4) nxt[v] = src4[pn * (C2 * K / 4) + lane + 32 * v];
      }
      __syncwarp();
      float y[8];
#pragma unroll
      for (int c = 0; c < C2; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; k += 4) {
          const float4 x = tile[(c * K + k) / 4];
          acc = fmaf(x.x, wreg[k], acc);
          acc = fmaf(x.y, wreg[k + 1], acc);
          acc = fmaf(x.z, wreg[k + 2], acc);
          acc = fmaf(x.w, wreg[k + 3], acc);
        }
        y[c] = acc;
      }
      // tanh with the forward-Laplacian rule (elementwise.py:42-72)
      const float t = tanhf(y[0] + b);
      if (C2 == 8) {
        const float d1 = 1.0f - t * t;
        float s2 = 0.f;
#pragma unroll
        for (int c = 1; c < 7; ++c) {
          s2 = fmaf(y[c], y[c], s2);
          y[c] *= d1;
        }
        y[7] = d1 * y[7] - 2.0f * t * d1 * s2;
      }
      y[0] = t;
      if (residual) {
#pragma unroll
        for (int c = 0; c < C2; ++c) y[c] = (tile_f[c * K + lane] + y[c]) * inv_sqrt2;   // K == 32 here
      }
      if (h2n) {
        float* out = h2n + pair * (C2 * 32) + lane;
#pragma unroll
        for (int c = 0; c < C2; ++c) out[c * 32] = y[c];
      }
      const int s = sp.chan_of(i);
      const bool c1 = (s == 1);
      const float inv_s = 1.0f / (float)(sp.hi(s) - sp.lo(s));
      sx0 += c1 ? 0.f : y[0];
      sx1 += c1 ? y[0] : 0.f;
      if (C2 == 8) {
        sl0 += c1 ? 0.f : y[7];
        sl1 += c1 ? y[7] : 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          sj0[a] += c1 ? 0.f : y[4 + a];
          sj1[a] += c1 ? y[4 + a] : 0.f;
          if (i == j) {
            own[a] = y[1 + a];
          } else {
            o[(long long)(1 + 3 * i + a) * FO + s * 32] = y[1 + a] * inv_s;
            if (nch == 2) o[(long long)(1 + 3 * i + a) * FO + (1 - s) * 32] = 0.f;
          }
        }
      }
      __syncwarp();
    }
    const int sjn = sp.chan_of(j);
    {
      const float inv0 = 1.0f / (float)(sp.hi(0) - sp.lo(0));
      o[0] = sx0 * inv0;
      if (C2 == 8) {
        o[(long long)(C - 1) * FO] = sl0 * inv0;
#pragma unroll
        for (int a = 0; a < 3; ++a) o[(long long)(1 + 3 * j + a) * FO] = ((sjn == 0 ? own[a] : 0.f) + sj0[a]) * inv0;
      }
    }
    if (nch == 2) {
      const float inv1 = 1.0f / (float)(sp.hi(1) - sp.lo(1));
      o[32] = sx1 * inv1;
      if (C2 == 8) {
        o[(long long)(C - 1) * FO + 32] = sl1 * inv1;
#pragma unroll
        for (int a = 0; a < 3; ++a) o[(long long)(1 + 3 * j + a) * FO + 32] = ((sjn == 1 ? own[a] : 0.f) + sj1[a]) * inv1;
      }
    }
  }
}

// Value-only (sampling path) version of the same layer, r2: with one pair per iteration (k_pair_layer_fused<K, 1>) only
// K / 4 lanes of the warp had a load in flight -- 128 bytes per warp, 1.5 TB/s (ncu).  Here a batch of PB = 128 / K pairs
// is fetched with one float4 per lane (512 bytes per warp in flight, the next batch prefetched), then the pairs of the
// batch are contracted one after the other from the shared tile.
template <int K>
__global__ void __launch_bounds__(256) k_pair_layer_value(const float* __restrict__ h2, const float* __restrict__ w0,
                                                         const float* __restrict__ bias, float* __restrict__ h2n,
                                                         float* __restrict__ g2, long long WJ, JqSpins sp, int residual) {
  constexpr int V4 = K / 4, PB = 32 / V4;
  __shared__ float4 tile_all[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* tile = tile_all[warp];
  const float* tile_f = reinterpret_cast<const float*>(tile);
  const int n = sp.n(), nch = sp.nch(), FO = nch * 32;
  float wreg[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wreg[k] = w0[k * 32 + lane];
  const float b = bias ? bias[lane] : 0.f;
  const float inv_sqrt2 = 0.70710678118654752440f;
  const float4* src4 = reinterpret_cast<const float4*>(h2);
  const int pp_l = lane / V4, v_l = lane - pp_l * V4;   // this lane's pair of the batch and float4 of that pair
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long wj = (long long)blockIdx.x * 8 + warp; wj < WJ; wj += (long long)gridDim.x * 8) {
    const long long w = wj / n;
    const int j = (int)(wj - w * n);
    float sx0 = 0.f, sx1 = 0.f;
    float4 nxt = (pp_l < n) ? src4[((w * n + pp_l) * n + j) * V4 + v_l] : zero4;
    for (int i0 = 0; i0 < n; i0 += PB) {
      tile[lane] = nxt;
      const int in = i0 + PB + pp_l;
      nxt = (in < n) ? src4[((w * n + in) * n + j) * V4 + v_l] : zero4;
      __syncwarp();
      const int np = (n - i0 < PB) ? n - i0 : PB;
      for (int pp = 0; pp < np; ++pp) {
        const int i = i0 + pp;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; k += 4) {
          const float4 x = tile[pp * V4 + k / 4];
          acc = fmaf(x.x, wreg[k], acc);
          acc = fmaf(x.y, wreg[k + 1], acc);
          acc = fmaf(x.z, wreg[k + 2], acc);
          acc = fmaf(x.w, wreg[k + 3], acc);
        }
        float y = tanhf(acc + b);
        if (residual) y = (tile_f[pp * K + lane] + y) * inv_sqrt2;   // K == 32 here
        if (h2n) h2n[((w * n + i) * n + j) * 32 + lane] = y;
        const bool c1 = sp.chan_of(i) == 1;
        sx0 += c1 ? 0.f : y;
        sx1 += c1 ? y : 0.f;
      }
      __syncwarp();
    }
    float* o = g2 + wj * FO + lane;
    o[0] = sx0 * (1.0f / (float)(sp.hi(0) - sp.lo(0)));
    if (nch == 2) o[32] = sx1 * (1.0f / (float)(sp.hi(1) - sp.lo(1)));
  }
}

// h2 [W][n*n][C2][K] -> h2n [W][n*n][C2][32] (or null) and g2 [W][n][C][nch*32] (C2 = 8, C = 3n+2 tracked; 1, 1 value
// only); false when the shape is not covered
bool jq_launch_pair_layer_fused(const float* h2, int K, const float* w0, const float* bias, float* h2n, float* g2, int W,
                                JqSpins sp, int N, int residual, int track, cudaStream_t st, int* rc) {
  *rc = JQ_OK;
  static const bool disabled = getenv("JAQMC_B200_UNFUSED_PAIR_LAYER") != nullptr;   // A/B switch
  if (disabled || N != 32 || (K != 4 && K != 32) || (residual && K != 32) || (reinterpret_cast<uintptr_t>(h2) & 15)) return false;
  if (jq_prep.collect) return true;   // dry pass: same decision, nothing launched
  const long long WJ = (long long)W * sp.n();
  if (WJ <= 0) return true;
  long long blocks = jq_cdiv(WJ, 8);   // one (walker, j) column per warp
  const double pairs = (double)W * sp.n() * sp.n();
  const int C2 = track ? 8 : 1, C = track ? 3 * sp.n() + 2 : 1;
  jq_prof_work(2.0 * pairs * C2 * K * 32, 4.0 * (pairs * C2 * (K + (h2n ? 32 : 0)) + (double)WJ * C * sp.nch() * 32));
#define JQ_PLF(KK, CC) JQ_LAUNCH((k_pair_layer_fused<KK, CC>), dim3((unsigned)blocks), dim3(256), 0, st, h2, w0, bias, h2n, g2, WJ, sp, residual)
  if (track) {
    if (K == 4) JQ_PLF(4, 8);
    else JQ_PLF(32, 8);
  } else {
    static const bool one_pair = getenv("JAQMC_B200_PAIR_LAYER_ONE_PAIR") != nullptr;   // A/B switch: one pair per iteration
    if (one_pair) {
      if (K == 4) JQ_PLF(4, 1);
      else JQ_PLF(32, 1);
    } else if (K == 4) {
      JQ_LAUNCH(k_pair_layer_value<4>, dim3((unsigned)blocks), dim3(256), 0, st, h2, w0, bias, h2n, g2, WJ, sp, residual);
    } else {
      JQ_LAUNCH(k_pair_layer_value<32>, dim3((unsigned)blocks), dim3(256), 0, st, h2, w0, bias, h2n, g2, WJ, sp, residual);
    }
  }
#undef JQ_PLF
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    jq_set_error("CUDA launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e));
    *rc = JQ_ERR_CUDA;
  }
  return true;
}
#endif

int jq_launch_pair_mean(const float* h2, float* g2, int W, JqSpins sp, int d2, int track, cudaStream_t st) {
  long long items = (long long)W * sp.n() * sp.nch() * d2;
  if (items <= 0) return JQ_OK;
  int grid = jq_cdiv(items, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  JQ_LAUNCH(k_pair_mean, dim3(grid), dim3(256), 0, st, h2, g2, items, sp, d2, track);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// First single-stream layer input: concat [ae_j | mean_s ae | g2_j]  (ferminet.py:65-90) as a dense
// augmented tensor [W][n][C][f1*(1+nch) + fg];  ae is Local1 [W][n][C1][f1], g2 dense [W][n][C][fg].
// One item per (walker, j, output column).
// ------------------------------------------------------------------------------------------------
__global__ void k_concat_layer1(const float* __restrict__ ae, const float* __restrict__ g2, float* __restrict__ out,
                                long long items, JqSpins sp, int f1, int fg, int track, int FO) {
  const int n = sp.n(), nch = sp.nch();
  const int C1 = track ? 5 : 1, C = track ? 3 * n + 2 : 1;
  const int FV = f1 * (1 + nch) + fg;   // columns that exist; [FV, FO) is zero padding
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    int col = (int)(it % FO);
    long long t = it / FO;
    int j = (int)(t % n);
    long long w = t / n;
    float* o = out + ((w * n + j) * (long long)C) * FO + col;
    if (col >= FV) {
      for (int c = 0; c < C; ++c) o[(long long)c * FO] = 0.f;
    } else if (col < f1) {
      const float* p = ae + ((w * n + j) * (long long)C1) * f1 + col;
      o[0] = p[0];
      if (track) {
        for (int c = 1; c < C - 1; ++c) o[(long long)c * FO] = 0.f;
        for (int a = 0; a < 3; ++a) o[(long long)(1 + 3 * j + a) * FO] = p[(1 + a) * f1];
        o[(long long)(C - 1) * FO] = p[4 * f1];
      }
    } else if (col < f1 * (1 + nch)) {
      int s = (col - f1) / f1, f = (col - f1) % f1;
      int lo = sp.lo(s), hi = sp.hi(s);
      float inv = 1.0f / (float)(hi - lo);
      float sx = 0.f, sl = 0.f;
      if (track)
        for (int c = 1; c < C - 1; ++c) o[(long long)c * FO] = 0.f;
      for (int e = lo; e < hi; ++e) {
        const float* p = ae + ((w * n + e) * (long long)C1) * f1 + f;
        sx += p[0];
        if (track) {
          sl += p[4 * f1];
          for (int a = 0; a < 3; ++a) o[(long long)(1 + 3 * e + a) * FO] = p[(1 + a) * f1] * inv;
        }
      }
      o[0] = sx * inv;
      if (track) o[(long long)(C - 1) * FO] = sl * inv;
    } else {
      int f = col - f1 * (1 + nch);
      const float* p = g2 + ((w * n + j) * (long long)C) * fg + f;
      for (int c = 0; c < C; ++c) o[(long long)c * FO] = p[(long long)c * fg];
    }
  }
}

int jq_launch_concat_layer1(const float* ae, const float* g2, float* out, int W, JqSpins sp, int f1, int fg,
                            int track, int ld_out, cudaStream_t st) {
  int FO = ld_out;
  long long items = (long long)W * sp.n() * FO;
  if (items <= 0) return JQ_OK;
  int grid = jq_cdiv(items, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  JQ_LAUNCH(k_concat_layer1, dim3(grid), dim3(256), 0, st, ae, g2, out, items, sp, f1, fg, track, FO);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// Spin means of a dense single stream: h [W][n][C][F] -> m [W][C][nch*F], m[w][c][s*F+f] = mean_{e in s} h
// (the walker-wide part of aggregate_features; it is the same for every electron of the walker, so its
// Dense contribution is computed once per walker and broadcast, instead of once per electron).
// ------------------------------------------------------------------------------------------------
__global__ void k_spin_mean(const float* __restrict__ h, float* __restrict__ m, long long items, JqSpins sp, int C,
                            int F, int vec) {
  // one item per (walker, component, channel, `vec` consecutive features); vec = 4 when F % 4 == 0, else 1
  const int n = sp.n(), nch = sp.nch();
  const int FV = F / vec;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    int fv = (int)(it % FV);
    long long t = it / FV;
    int s = (int)(t % nch);
    t /= nch;
    int c = (int)(t % C);
    long long w = t / C;
    int lo = sp.lo(s), hi = sp.hi(s);
    const float* base = h + (((w * n + lo) * (long long)C) + c) * F + vec * fv;
    const long long stride = (long long)C * F;
    const float inv = 1.0f / (float)(hi - lo);
    float* o = m + ((w * C + c) * (long long)nch + s) * F + vec * fv;
    if (vec == 4) {
      float4 a = {0.f, 0.f, 0.f, 0.f};
      for (int e = lo; e < hi; ++e) {
        float4 v = *reinterpret_cast<const float4*>(base + (e - lo) * stride);
        a.x += v.x;
        a.y += v.y;
        a.z += v.z;
        a.w += v.w;
      }
      a.x *= inv;
      a.y *= inv;
      a.z *= inv;
      a.w *= inv;
      *reinterpret_cast<float4*>(o) = a;
    } else {
      float a = 0.f;
      for (int e = lo; e < hi; ++e) a += base[(e - lo) * stride];
      o[0] = a * inv;
    }
  }
}

int jq_launch_spin_mean(const float* h, float* m, int W, JqSpins sp, int C, int F, cudaStream_t st) {
  const int vec = (F % 4 == 0) ? 4 : 1;
  long long items = (long long)W * C * sp.nch() * (F / vec);
  if (items <= 0) return JQ_OK;
  int grid = jq_cdiv(items, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  jq_prof_work(0.0, 4.0 * (double)W * C * F * (sp.n() + sp.nch()));
  JQ_LAUNCH(k_spin_mean, dim3(grid), dim3(256), 0, st, h, m, items, sp, C, F, vec);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}
