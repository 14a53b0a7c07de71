// Attention-network building blocks under the forward Laplacian: Local1 -> dense expansion, LayerNorm, and
// softmax attention (LapNet cross-stream attention and Psiformer self-attention share one kernel).
//
// Reference semantics:
//   * flax nn.LayerNorm(epsilon) as used by backbone/psiformer.py:74,86,88,98: mean / fast variance over the last
//     axis, y = (x - mu) rsqrt(var + eps) scale + bias; traced by the interpreter through the mean / square / rsqrt /
//     product rules (laplacian/primitives/{reductions,elementwise,arithmetic}.py)
//   * attention  softmax(q k^T / sqrt(d)) v  with operands (n, heads, d): backbone/lapnet/_attention.py:20-34 and
//     its hand rule :113-196 (q, k Local1, v dense); flax MultiHeadDotProductAttention for Psiformer
//     (backbone/psiformer.py:76-82), traced through dot_general (laplacian/primitives/dot_general.py:410-449) and softmax
// Both are evaluated here in closed form (Appendix A of SURVEY.md): same function, same derivatives.
#include <cstdlib>

#include "aug.cuh"

// ------------------------------------------------------------------------------------------------
// Local1 [W][n][5][F] -> dense [W][n][3n+2][F]: electron i owns Jacobian columns 3i..3i+2
// (laplacian/sparse.py Local1Jacobian.to_dense).  One item per output element.
// ------------------------------------------------------------------------------------------------
__global__ void k_densify_local1(const float* __restrict__ in, float* __restrict__ out, long long G, int n, int F) {
  // one block per group g = (walker, electron i); a thread owns features f = tid, tid + blockDim, ... and writes the
  // 3n + 2 rows from its five inputs (no per-element index arithmetic: the pass is a pure 2.6 GB write)
  const int C = 3 * n + 2;
  for (long long g = blockIdx.x; g < G; g += gridDim.x) {
    const int i = (int)(g % n);
    const float* p = in + g * 5 * F;
    float* o = out + g * (long long)C * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      const float v0 = p[f], j0 = p[F + f], j1 = p[2 * F + f], j2 = p[3 * F + f], l = p[4 * F + f];
      o[f] = v0;
      for (int e = 0; e < n; ++e) {
        const bool own = (e == i);
        o[(long long)(1 + 3 * e) * F + f] = own ? j0 : 0.f;
        o[(long long)(2 + 3 * e) * F + f] = own ? j1 : 0.f;
        o[(long long)(3 + 3 * e) * F + f] = own ? j2 : 0.f;
      }
      o[(long long)(C - 1) * F + f] = l;
    }
  }
}

int jq_launch_densify_local1(const float* in, float* out, long long W, int n, int F, cudaStream_t st) {
  const long long G = W * n;
  if (G <= 0 || F <= 0) return JQ_OK;
  long long grid = G;
  if (grid > 148LL * 64) grid = 148LL * 64;
  jq_prof_work(0.0, 4.0 * (double)G * (3 * n + 2) * F);
  JQ_LAUNCH(k_densify_local1, dim3((unsigned)grid), dim3(256), 0, st, in, out, G, n, F);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm with the forward-Laplacian rule.  One block per group (walker, electron); x, out [G][C][F].
//   mu = mean x, var = max(0, mean x^2 - mu^2), s = rsqrt(var + eps), xc = x - mu, y = xc s scale + bias
//   J rows:  muJ = mean J, xcJ = J - muJ, varJ = 2 mean(xc xcJ), sJ = -1/2 s^3 varJ,  yJ = (xcJ s + xc sJ) scale
//   L row:   varL = 2 mean(xc xcL) + 2 sum_k mean(xcJ_k^2),  sL = -1/2 s^3 varL + 3/4 s^5 sum_k varJ_k^2,
//            yL = (xcL s + xc sL + 2 sum_k xcJ_k sJ_k) scale
// Shared: mu[C] | dot[C] | sq[C] | sj[C] | part[C*LN_T] | scal[8]
// ------------------------------------------------------------------------------------------------
#define LN_T 32
__global__ void k_layernorm_fl(const float* __restrict__ x, const float* __restrict__ scale,
                               const float* __restrict__ bias, float* __restrict__ out, int C, int F, float eps) {
  JQ_DYN_SMEM(float, sm);
  float* mu = sm;
  float* dot = mu + C;
  float* sq = dot + C;
  float* sj = sq + C;
  float* part = sj + C;
  float* part2 = part + C * LN_T;
  float* scal = part2 + C * LN_T;
  const long long g = blockIdx.x;
  const float* xg = x + g * (long long)C * F;
  float* og = out + g * (long long)C * F;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float invF = 1.0f / (float)F;
  const int K = C - 2;  // Jacobian rows (C == 1: value only)

  // pass 1: row sums (+ sum of squares of the value row)
  for (int q = tid; q < C * LN_T; q += nt) {
    int c = q / LN_T, t = q % LN_T;
    const float* r = xg + (long long)c * F;
    float s = 0.f, s2 = 0.f;
    for (int f = t; f < F; f += LN_T) {
      float v = r[f];
      s += v;
      if (c == 0) s2 = fmaf(v, v, s2);
    }
    part[q] = s;
    if (c == 0) part2[t] = s2;
  }
  __syncthreads();
  for (int c = tid; c < C; c += nt) {
    float s = 0.f;
    for (int t = 0; t < LN_T; ++t) s += part[c * LN_T + t];
    mu[c] = s * invF;
  }
  if (tid == 0) {
    float s2 = 0.f;
    for (int t = 0; t < LN_T; ++t) s2 += part2[t];
    scal[7] = s2 * invF;  // mean x^2
  }
  __syncthreads();
  if (tid == 0) {
    float var = scal[7] - mu[0] * mu[0];
    if (var < 0.f) var = 0.f;
    scal[0] = rsqrtf(var + eps);
  }
  // pass 2: dot[c] = sum xc (row_c - mu_c), sq[c] = sum (row_c - mu_c)^2 for derivative rows
  if (C > 1) {
    const float mu0 = mu[0];
    for (int q = tid; q < (C - 1) * LN_T; q += nt) {
      int c = 1 + q / LN_T, t = q % LN_T;
      const float* r = xg + (long long)c * F;
      const float m = mu[c];
      float d = 0.f, s2 = 0.f;
      for (int f = t; f < F; f += LN_T) {
        float v = r[f] - m;
        d = fmaf(xg[f] - mu0, v, d);
        s2 = fmaf(v, v, s2);
      }
      part[c * LN_T + t] = d;
      part2[c * LN_T + t] = s2;
    }
  }
  __syncthreads();
  const float s = scal[0];
  if (C > 1) {
    for (int c = 1 + tid; c < C; c += nt) {
      float d = 0.f, s2 = 0.f;
      for (int t = 0; t < LN_T; ++t) {
        d += part[c * LN_T + t];
        s2 += part2[c * LN_T + t];
      }
      dot[c] = d;
      sq[c] = s2;
      sj[c] = -0.5f * s * s * s * (2.0f * d * invF);  // sJ (for c == C-1 this is the first part of sL)
    }
    __syncthreads();
    if (tid == 0) {
      float sumq = 0.f, sumv2 = 0.f;
      for (int k = 1; k <= K; ++k) {
        sumq += sq[k];
        float vj = 2.0f * dot[k] * invF;
        sumv2 = fmaf(vj, vj, sumv2);
      }
      float s3 = s * s * s;
      float varL = 2.0f * dot[C - 1] * invF + 2.0f * sumq * invF;
      scal[1] = -0.5f * s3 * varL + 0.75f * s3 * s * s * sumv2;  // sL
    }
    __syncthreads();
  }
  // pass 3: outputs
  const float mu0 = mu[0];
  for (int q = tid; q < (C > 1 ? C - 1 : 1) * F; q += nt) {
    int c = q / F, f = q % F;
    float xc = xg[f] - mu0;
    float sc = scale ? scale[f] : 1.0f;
    if (c == 0) {
      float y = xc * s * sc;
      if (bias) y += bias[f];
      og[f] = y;
    } else {
      float v = xg[(long long)c * F + f] - mu[c];
      og[(long long)c * F + f] = (v * s + xc * sj[c]) * sc;
    }
  }
  if (C > 1) {
    const float sL = scal[1];
    for (int f = tid; f < F; f += nt) {
      float xc = xg[f] - mu0;
      float acc = 0.f;
      for (int k = 1; k <= K; ++k) acc = fmaf(xg[(long long)k * F + f] - mu[k], sj[k], acc);
      float v = xg[(long long)(C - 1) * F + f] - mu[C - 1];
      float sc = scale ? scale[f] : 1.0f;
      og[(long long)(C - 1) * F + f] = (v * s + xc * sL + 2.0f * acc) * sc;
    }
  }
}

#ifndef JAQMC_HOST_EMU
// Same rule with the group [C][F] cached in shared memory: global memory is read once and written once (k_layernorm_fl
// walks it three times).  Row statistics: one warp per row, float4 reads, shuffle reduction.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) k_layernorm_fl_smem(const float* __restrict__ x, const float* __restrict__ scale,
                                                          const float* __restrict__ bias, float* __restrict__ out, int C,
                                                          int F, float eps) {
  JQ_DYN_SMEM(float, sm);
  float* xs = sm;                    // [C][F]
  float* mu = xs + (size_t)C * F;    // [C]
  float* dot = mu + C;
  float* sq = dot + C;
  float* sj = sq + C;
  float* scal = sj + C;              // [8]
  const long long g = blockIdx.x;
  const float4* xg4 = reinterpret_cast<const float4*>(x + g * (long long)C * F);
  float* og = out + g * (long long)C * F;
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const float invF = 1.0f / (float)F;
  const int K = C - 2;
  const int F4 = F >> 2;
  float4* xs4 = reinterpret_cast<float4*>(xs);
  for (int q = tid; q < C * F4; q += nt) xs4[q] = xg4[q];
  __syncthreads();
  // row means (+ mean square of the value row)
  for (int c = warp; c < C; c += nw) {
    float s = 0.f, s2 = 0.f;
    for (int f4 = lane; f4 < F4; f4 += 32) {
      const float4 v = xs4[c * F4 + f4];
      s += (v.x + v.y) + (v.z + v.w);
      if (c == 0) s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
    }
    s = warp_sum(s);
    if (c == 0) s2 = warp_sum(s2);
    if (lane == 0) {
      mu[c] = s * invF;
      if (c == 0) scal[7] = s2 * invF;
    }
  }
  __syncthreads();
  if (tid == 0) {
    float var = scal[7] - mu[0] * mu[0];
    if (var < 0.f) var = 0.f;
    scal[0] = rsqrtf(var + eps);
  }
  __syncthreads();
  const float s = scal[0];
  const float mu0 = mu[0];
  if (C > 1) {
    for (int c = 1 + warp; c < C; c += nw) {
      const float m = mu[c];
      float d = 0.f, s2 = 0.f;
      for (int f4 = lane; f4 < F4; f4 += 32) {
        const float4 a = xs4[f4], v = xs4[c * F4 + f4];
        const float v0 = v.x - m, v1 = v.y - m, v2 = v.z - m, v3 = v.w - m;
        d = fmaf(a.x - mu0, v0, d);
        d = fmaf(a.y - mu0, v1, d);
        d = fmaf(a.z - mu0, v2, d);
        d = fmaf(a.w - mu0, v3, d);
        s2 = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, s2))));
      }
      d = warp_sum(d);
      s2 = warp_sum(s2);
      if (lane == 0) {
        dot[c] = d;
        sq[c] = s2;
        sj[c] = -0.5f * s * s * s * (2.0f * d * invF);
      }
    }
    __syncthreads();
    if (tid == 0) {
      float sumq = 0.f, sumv2 = 0.f;
      for (int k = 1; k <= K; ++k) {
        sumq += sq[k];
        const float vj = 2.0f * dot[k] * invF;
        sumv2 = fmaf(vj, vj, sumv2);
      }
      const float s3 = s * s * s;
      const float varL = 2.0f * dot[C - 1] * invF + 2.0f * sumq * invF;
      scal[1] = -0.5f * s3 * varL + 0.75f * s3 * s * s * sumv2;  // sL
    }
    __syncthreads();
  }
  // outputs: a thread owns feature columns f = tid, tid + blockDim, ... and walks the rows (no index division; the
  // Laplacian row's sum over the Jacobian rows is accumulated on the way)
  for (int f = tid; f < F; f += nt) {
    const float xc = xs[f] - mu0;
    const float sc = scale ? scale[f] : 1.0f;
    float y = xc * s * sc;
    if (bias) y += bias[f];
    og[f] = y;
    if (C > 1) {
      float acc = 0.f;
      for (int k = 1; k <= K; ++k) {
        const float v = xs[k * F + f] - mu[k];
        const float sjk = sj[k];
        og[(long long)k * F + f] = (v * s + xc * sjk) * sc;
        acc = fmaf(v, sjk, acc);
      }
      const float v = xs[(C - 1) * F + f] - mu[C - 1];
      og[(long long)(C - 1) * F + f] = (v * s + xc * scal[1] + 2.0f * acc) * sc;
    }
  }
}
#endif

#ifndef JAQMC_HOST_EMU
// Same rule, streaming (r2): a derivative row's statistics depend only on the row itself and on the value row, so each
// WARP takes rows c = 1 + warp, 1 + warp + nw, ...: loads the row into registers (F / 32 values per lane, float4), reduces
// mu_c, dot_c, sq_c with shuffles and writes the normalised row at once -- the group never sits in shared memory, several
// blocks are resident per SM and every warp keeps the next row in flight while it works on the current one.  The value
// row's statistics are recomputed by every warp (1 KB from L1); the Laplacian row needs sum_k xcJ_k sJ_k per feature and
// two scalars, accumulated per warp and reduced in a fixed order.   F = 128 NV.
constexpr int LNR_WARPS = 8;
__device__ __forceinline__ float ln_hsum(const float4& v) { return (v.x + v.y) + (v.z + v.w); }

template <int NV>
__global__ void __launch_bounds__(LNR_WARPS * 32) k_layernorm_fl_rows(const float* __restrict__ x, const float* __restrict__ scale,
                                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                                      int C, float eps) {
  constexpr int F = 128 * NV, F4 = 32 * NV;
  __shared__ float4 acc_s[LNR_WARPS][F4];
  __shared__ float red_s[LNR_WARPS][2];
  const long long g = blockIdx.x;
  const float4* xg4 = reinterpret_cast<const float4*>(x + g * (long long)C * F);
  float4* og4 = reinterpret_cast<float4*>(out + g * (long long)C * F);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invF = 1.0f / (float)F;
  float4 xc[NV], sc[NV];
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    xc[u] = xg4[lane + 32 * u];
    s += ln_hsum(xc[u]);
    s2 = fmaf(xc[u].x, xc[u].x, fmaf(xc[u].y, xc[u].y, fmaf(xc[u].z, xc[u].z, fmaf(xc[u].w, xc[u].w, s2))));
    sc[u] = scale ? reinterpret_cast<const float4*>(scale)[lane + 32 * u] : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  const float mu0 = s * invF;
  float var = s2 * invF - mu0 * mu0;
  if (var < 0.f) var = 0.f;
  const float sinv = rsqrtf(var + eps);
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    xc[u].x -= mu0; xc[u].y -= mu0; xc[u].z -= mu0; xc[u].w -= mu0;
  }
  if (warp == 0) {
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      float4 y = make_float4(xc[u].x * sinv * sc[u].x, xc[u].y * sinv * sc[u].y, xc[u].z * sinv * sc[u].z, xc[u].w * sinv * sc[u].w);
      if (bias) {
        const float4 b = reinterpret_cast<const float4*>(bias)[lane + 32 * u];
        y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
      }
      og4[lane + 32 * u] = y;
    }
  }
  if (C == 1) return;
  const int K = C - 2;
  const float s3 = sinv * sinv * sinv;
  float4 acc[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  float sumq = 0.f, sumv2 = 0.f;
  float4 nxt[NV];
  int c = 1 + warp;
  if (c <= K) {
#pragma unroll
    for (int u = 0; u < NV; ++u) nxt[u] = xg4[(long long)c * F4 + lane + 32 * u];
  }
  for (; c <= K; c += LNR_WARPS) {
    float4 v[NV];
    float m = 0.f;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      v[u] = nxt[u];
      m += ln_hsum(v[u]);
    }
    if (c + LNR_WARPS <= K) {
#pragma unroll
      for (int u = 0; u < NV; ++u) nxt[u] = xg4[(long long)(c + LNR_WARPS) * F4 + lane + 32 * u];
    }
    m = warp_sum(m) * invF;
    float d = 0.f, q = 0.f;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      v[u].x -= m; v[u].y -= m; v[u].z -= m; v[u].w -= m;
      d = fmaf(xc[u].x, v[u].x, fmaf(xc[u].y, v[u].y, fmaf(xc[u].z, v[u].z, fmaf(xc[u].w, v[u].w, d))));
      q = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, fmaf(v[u].w, v[u].w, q))));
    }
    d = warp_sum(d);
    q = warp_sum(q);
    const float vj = 2.0f * d * invF;
    const float sj = -0.5f * s3 * vj;
    sumq += q;
    sumv2 = fmaf(vj, vj, sumv2);
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      og4[(long long)c * F4 + lane + 32 * u] =
          make_float4((v[u].x * sinv + xc[u].x * sj) * sc[u].x, (v[u].y * sinv + xc[u].y * sj) * sc[u].y,
                      (v[u].z * sinv + xc[u].z * sj) * sc[u].z, (v[u].w * sinv + xc[u].w * sj) * sc[u].w);
      acc[u].x = fmaf(v[u].x, sj, acc[u].x);
      acc[u].y = fmaf(v[u].y, sj, acc[u].y);
      acc[u].z = fmaf(v[u].z, sj, acc[u].z);
      acc[u].w = fmaf(v[u].w, sj, acc[u].w);
    }
  }
#pragma unroll
  for (int u = 0; u < NV; ++u) acc_s[warp][lane + 32 * u] = acc[u];
  if (lane == 0) {
    red_s[warp][0] = sumq;
    red_s[warp][1] = sumv2;
  }
  __syncthreads();
  if (warp == 0) {
    float tq = 0.f, tv = 0.f;
    for (int ww = 0; ww < LNR_WARPS; ++ww) {
      tq += red_s[ww][0];
      tv += red_s[ww][1];
    }
    float4 v[NV];
    float m = 0.f;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      v[u] = xg4[(long long)(C - 1) * F4 + lane + 32 * u];
      m += ln_hsum(v[u]);
    }
    m = warp_sum(m) * invF;
    float d = 0.f;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      v[u].x -= m; v[u].y -= m; v[u].z -= m; v[u].w -= m;
      d = fmaf(xc[u].x, v[u].x, fmaf(xc[u].y, v[u].y, fmaf(xc[u].z, v[u].z, fmaf(xc[u].w, v[u].w, d))));
    }
    d = warp_sum(d);
    const float varL = 2.0f * d * invF + 2.0f * tq * invF;
    const float sL = -0.5f * s3 * varL + 0.75f * s3 * sinv * sinv * tv;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ww = 0; ww < LNR_WARPS; ++ww) {
        const float4 p = acc_s[ww][lane + 32 * u];
        a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
      }
      og4[(long long)(C - 1) * F4 + lane + 32 * u] =
          make_float4((v[u].x * sinv + xc[u].x * sL + 2.0f * a.x) * sc[u].x, (v[u].y * sinv + xc[u].y * sL + 2.0f * a.y) * sc[u].y,
                      (v[u].z * sinv + xc[u].z * sL + 2.0f * a.z) * sc[u].z, (v[u].w * sinv + xc[u].w * sL + 2.0f * a.w) * sc[u].w);
    }
  }
}
#endif

int jq_launch_layernorm_fl(const float* x, const float* scale, const float* bias, float* out, long long G, int C, int F,
                           float eps, cudaStream_t st) {
  return jq_launch_layernorm_fl_sel(x, scale, bias, out, G, C, F, eps, 0, st);
}

int jq_launch_layernorm_fl_sel(const float* x, const float* scale, const float* bias, float* out, long long G, int C, int F,
                               float eps, int force, cudaStream_t st) {
  if (G <= 0) return JQ_OK;
  JQ_REQUIRE(x != out, JQ_ERR_INVALID_ARGUMENT, "layernorm: in-place is not supported");
#ifndef JAQMC_HOST_EMU
  {
    static const bool env_no_rows = getenv("JAQMC_B200_LAYERNORM_SMEM") != nullptr;   // A/B switch: the shared-memory kernel
    const bool no_rows = force ? force != 3 : env_no_rows;
    const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(scale) |
                        reinterpret_cast<uintptr_t>(bias)) & 15) == 0;
    if (!no_rows && al16 && (F == 128 || F == 256 || F == 512)) {
      jq_prof_work(0.0, 8.0 * (double)G * C * F);
      if (F == 128) JQ_LAUNCH(k_layernorm_fl_rows<1>, dim3((unsigned)G), dim3(LNR_WARPS * 32), 0, st, x, scale, bias, out, C, eps);
      else if (F == 256) JQ_LAUNCH(k_layernorm_fl_rows<2>, dim3((unsigned)G), dim3(LNR_WARPS * 32), 0, st, x, scale, bias, out, C, eps);
      else JQ_LAUNCH(k_layernorm_fl_rows<4>, dim3((unsigned)G), dim3(LNR_WARPS * 32), 0, st, x, scale, bias, out, C, eps);
      JQ_CHECK_LAUNCH();
      return JQ_OK;
    }
  }
  {
    const size_t sc = sizeof(float) * ((size_t)C * F + 4 * (size_t)C + 8);
    if (F % 4 == 0 && sc <= 200 * 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (force == 0 || force == 2)) {
      cudaError_t e = cudaFuncSetAttribute(k_layernorm_fl_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc);
      JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "layernorm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      jq_prof_work(0.0, 8.0 * (double)G * C * F);
      JQ_LAUNCH(k_layernorm_fl_smem, dim3((unsigned)G), dim3(256), sc, st, x, scale, bias, out, C, F, eps);
      JQ_CHECK_LAUNCH();
      return JQ_OK;
    }
  }
#endif
  JQ_REQUIRE(force == 0 || force == 1, JQ_ERR_UNSUPPORTED, "layernorm: kernel %d does not support C=%d F=%d", force, C, F);
  size_t smem = sizeof(float) * ((size_t)4 * C + (size_t)2 * C * LN_T + 8);
  JQ_REQUIRE(smem <= 200 * 1024, JQ_ERR_UNSUPPORTED, "layernorm: %d components need %zu bytes of shared memory", C, smem);
#ifndef JAQMC_HOST_EMU
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_layernorm_fl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "layernorm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
#endif
  jq_prof_work(0.0, 8.0 * (double)G * C * F);
  JQ_LAUNCH(k_layernorm_fl, dim3((unsigned)G), dim3(256), smem, st, x, scale, bias, out, C, F, eps);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}

// ------------------------------------------------------------------------------------------------
// softmax attention with the forward-Laplacian rule.  One block per (walker, head).
//   a_ij = q_i . k_j / sqrt(d),  w = softmax_j(a),  o_i = sum_j w_ij v_j
//   per coordinate k:  aJ = (qJ.k + q.kJ)/sqrt(d),  abar_i = sum_m w_im aJ_im,  wJ_ij = w_ij (aJ_ij - abar_i),
//                      oJ_i = sum_j (wJ_ij v_j + w_ij vJ_j)
//   Laplacian:  aL = (qL.k + q.kL + 2 sum_k qJ_k.kJ_k)/sqrt(d),
//               wL_ij = sum_k wJ_ij (aJ_ij - abar_i) + w_ij (aL_ij - sum_m w_im aL_im - sum_k sum_m wJ_im aJ_im),
//               oL_i = sum_j (wL_ij v_j + w_ij vL_j) + 2 sum_k sum_j wJ_ij vJ_kj
// Operand layouts: t[W][n][Ct][ld] with Ct = 3n+2 (dense) or 5 (Local1: electron i carries only its own 3
// Jacobian columns); head h occupies columns [off + h*dh, off + (h+1)*dh).
// Shared (floats, rows padded to dh+1): q0 k0 v0 qJ kJ vJ oL [n][dh+1] | w aJ aL t1 wJ [n][n] | abar t2 [n]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float attn_fetch(const JqAttnOperand& t, long long w, int n, int i, int comp_dense, int col,
                                            int Cd) {
  // comp_dense in [0, Cd): 0 value, 1..3n Jacobian column k = comp-1, Cd-1 Laplacian
  const float* base = t.p + ((w * n + i) * (long long)t.C) * t.ld + col;
  if (t.C == Cd) return base[(long long)comp_dense * t.ld];
  if (comp_dense == 0) return base[0];
  if (comp_dense == Cd - 1) return base[(long long)4 * t.ld];
  int k = comp_dense - 1;
  if (k / 3 != i) return 0.f;
  return base[(long long)(1 + k % 3) * t.ld];
}

// One component of one head of an operand -> shared tile dst[i][d] (row stride ldh).  Device build: dense operands go
// through 4-byte cp.async, so that a thread's ~30 element loads are all in flight together (the plain
// load -> store loop paid the global-memory latency once per element: 30 % of the kernel's stall samples, r2 profile);
// the caller commits / waits.  Local1 operands (a single electron's columns) are mostly zeros: stored directly.
__device__ __forceinline__ void attn_stage(float* dst, const JqAttnOperand& t, long long w, int n, int comp, int col0,
                                           int Cd, int dh, int ldh, int tid, int nt) {
  const int ndh = n * dh;
#ifndef JAQMC_HOST_EMU
  if (t.C == Cd) {
    const float* base = t.p + ((w * n) * (long long)t.C + comp) * t.ld + col0;
    const long long istride = (long long)t.C * t.ld;
    const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
    for (int x = tid; x < ndh; x += nt) {
      const int i = x / dh, d = x - i * dh;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * (unsigned)(i * ldh + d)), "l"(base + i * istride + d));
    }
    return;
  }
#endif
  for (int x = tid; x < ndh; x += nt) {
    const int i = x / dh, d = x % dh;
    dst[i * ldh + d] = attn_fetch(t, w, n, i, comp, col0 + d, Cd);
  }
}
__device__ __forceinline__ void attn_stage_wait() {
#ifndef JAQMC_HOST_EMU
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 0;");
#endif
}

template <int TAI, int TAJ, int TOI, int TOD>
__global__ void k_attention_fl(JqAttnOperand q, JqAttnOperand k, JqAttnOperand v, float* __restrict__ out, int ldo,
                               int n, int H, int dh, int track) {
  JQ_DYN_SMEM(float, sm);
  const int ldh = dh + 1;
  const int nd = n * ldh, nn = n * n;
  float* q0 = sm;
  float* k0 = q0 + nd;
  float* v0 = k0 + nd;
  float* qJ = v0 + nd;
  float* kJ = qJ + nd;
  float* vJ = kJ + nd;
  float* oL = vJ + nd;
  float* wgt = oL + nd;
  float* aJ = wgt + nn;
  float* aL = aJ + nn;
  float* t1 = aL + nn;
  float* wJ = t1 + nn;
  float* abar = wJ + nn;
  float* t2 = abar + n;
  const long long w = blockIdx.x / H;
  const int h = blockIdx.x % H;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int Cd = track ? 3 * n + 2 : 1;
  const int K = track ? 3 * n : 0;
  const float scale = rsqrtf((float)dh);
  const int ndh = n * dh;
  // register tiles: TAI x TAJ logits per item in the Jacobian phase, TOI electrons x TOD features in the output phase
  const int nbi = (n + TAI - 1) / TAI, nbj = (n + TAJ - 1) / TAJ, nbo = (n + TOI - 1) / TOI, nbd = (dh + TOD - 1) / TOD;
  const int tiles_a = nbi * nbj, tiles_o = nbo * nbd;

  for (int x = tid; x < ndh; x += nt) {
    int i = x / dh, d = x % dh;
    q0[i * ldh + d] = attn_fetch(q, w, n, i, 0, q.off + h * dh + d, Cd);
    k0[i * ldh + d] = attn_fetch(k, w, n, i, 0, k.off + h * dh + d, Cd);
    v0[i * ldh + d] = attn_fetch(v, w, n, i, 0, v.off + h * dh + d, Cd);
    oL[i * ldh + d] = 0.f;
  }
  __syncthreads();
  for (int x = tid; x < nn; x += nt) {
    int i = x / n, j = x % n;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(q0[i * ldh + d], k0[j * ldh + d], acc);
    wgt[x] = acc * scale;
    aL[x] = 0.f;
    t1[x] = 0.f;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    float m = wgt[i * n];
    for (int j = 1; j < n; ++j) m = fmaxf(m, wgt[i * n + j]);
    float z = 0.f;
    for (int j = 0; j < n; ++j) {
      float e = expf(wgt[i * n + j] - m);
      wgt[i * n + j] = e;
      z += e;
    }
    float zi = 1.0f / z;
    for (int j = 0; j < n; ++j) wgt[i * n + j] *= zi;
    t2[i] = 0.f;
  }
  __syncthreads();
  // value row
  for (int x = tid; x < ndh; x += nt) {
    int i = x / dh, d = x % dh;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = fmaf(wgt[i * n + j], v0[j * ldh + d], acc);
    out[((w * n + i) * (long long)Cd) * ldo + h * dh + d] = acc;
  }
  if (!track) return;

  for (int kk = 0; kk < K; ++kk) {
    const int comp = 1 + kk;
    __syncthreads();
    attn_stage(qJ, q, w, n, comp, q.off + h * dh, Cd, dh, ldh, tid, nt);
    attn_stage(kJ, k, w, n, comp, k.off + h * dh, Cd, dh, ldh, tid, nt);
    attn_stage(vJ, v, w, n, comp, v.off + h * dh, Cd, dh, ldh, tid, nt);
    attn_stage_wait();
    __syncthreads();
    // logit Jacobians in 4 x 4 register tiles (r2): 16 shared-memory words per 48 multiply-adds instead of 4 per 3 --
    // the kernel is bound by the shared-memory pipe
    for (int x = tid; x < tiles_a; x += nt) {
      const int i0 = TAI * (x / nbj), j0 = TAJ * (x % nbj);
      int ir[TAI], jr[TAJ];
#pragma unroll
      for (int a = 0; a < TAI; ++a) ir[a] = ((i0 + a < n) ? i0 + a : n - 1) * ldh;
#pragma unroll
      for (int b = 0; b < TAJ; ++b) jr[b] = ((j0 + b < n) ? j0 + b : n - 1) * ldh;
      float a1[TAI][TAJ], a2[TAI][TAJ];
#pragma unroll
      for (int a = 0; a < TAI; ++a)
#pragma unroll
        for (int b = 0; b < TAJ; ++b) a1[a][b] = a2[a][b] = 0.f;
      for (int d = 0; d < dh; ++d) {
        float qj[TAI], qv[TAI], kj[TAJ], kv[TAJ];
#pragma unroll
        for (int a = 0; a < TAI; ++a) {
          qj[a] = qJ[ir[a] + d];
          qv[a] = q0[ir[a] + d];
        }
#pragma unroll
        for (int b = 0; b < TAJ; ++b) {
          kj[b] = kJ[jr[b] + d];
          kv[b] = k0[jr[b] + d];
        }
#pragma unroll
        for (int a = 0; a < TAI; ++a)
#pragma unroll
          for (int b = 0; b < TAJ; ++b) {
            a1[a][b] = fmaf(qj[a], kv[b], a1[a][b]);
            a1[a][b] = fmaf(qv[a], kj[b], a1[a][b]);
            a2[a][b] = fmaf(qj[a], kj[b], a2[a][b]);
          }
      }
#pragma unroll
      for (int a = 0; a < TAI; ++a)
#pragma unroll
        for (int b = 0; b < TAJ; ++b)
          if (i0 + a < n && j0 + b < n) {
            const int xx = (i0 + a) * n + j0 + b;
            aJ[xx] = a1[a][b] * scale;
            aL[xx] = fmaf(2.0f * scale, a2[a][b], aL[xx]);
          }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      float ab = 0.f;
      for (int m = 0; m < n; ++m) ab = fmaf(wgt[i * n + m], aJ[i * n + m], ab);
      abar[i] = ab;
    }
    __syncthreads();
    for (int x = tid; x < nn; x += nt) {
      int i = x / n;
      float c = aJ[x] - abar[i];
      float wj = wgt[x] * c;
      t1[x] = fmaf(wj, c, t1[x]);
      wJ[x] = wj;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      float acc = 0.f;
      for (int m = 0; m < n; ++m) acc = fmaf(wJ[i * n + m], aJ[i * n + m], acc);
      t2[i] += acc;
    }
    // oJ = wJ v + w vJ, oL += 2 wJ vJ in 4 (electrons) x 4 (features) register tiles
    for (int x = tid; x < tiles_o; x += nt) {
      const int i0 = TOI * (x / nbd), d0 = TOD * (x % nbd);
      int ir[TOI], dc[TOD];
#pragma unroll
      for (int a = 0; a < TOI; ++a) ir[a] = ((i0 + a < n) ? i0 + a : n - 1) * n;
#pragma unroll
      for (int b = 0; b < TOD; ++b) dc[b] = (d0 + b < dh) ? d0 + b : dh - 1;
      float acc[TOI][TOD], acc2[TOI][TOD];
#pragma unroll
      for (int a = 0; a < TOI; ++a)
#pragma unroll
        for (int b = 0; b < TOD; ++b) acc[a][b] = acc2[a][b] = 0.f;
      for (int j = 0; j < n; ++j) {
        float wj[TOI], wv[TOI], vv[TOD], vj[TOD];
#pragma unroll
        for (int a = 0; a < TOI; ++a) {
          wj[a] = wJ[ir[a] + j];
          wv[a] = wgt[ir[a] + j];
        }
#pragma unroll
        for (int b = 0; b < TOD; ++b) {
          vv[b] = v0[j * ldh + dc[b]];
          vj[b] = vJ[j * ldh + dc[b]];
        }
#pragma unroll
        for (int a = 0; a < TOI; ++a)
#pragma unroll
          for (int b = 0; b < TOD; ++b) {
            acc[a][b] = fmaf(wj[a], vv[b], acc[a][b]);
            acc[a][b] = fmaf(wv[a], vj[b], acc[a][b]);
            acc2[a][b] = fmaf(wj[a], vj[b], acc2[a][b]);
          }
      }
#pragma unroll
      for (int a = 0; a < TOI; ++a)
#pragma unroll
        for (int b = 0; b < TOD; ++b)
          if (i0 + a < n && d0 + b < dh) {
            const int i = i0 + a, d = d0 + b;
            out[((w * n + i) * (long long)Cd + comp) * ldo + h * dh + d] = acc[a][b];
            oL[i * ldh + d] = fmaf(2.0f, acc2[a][b], oL[i * ldh + d]);
          }
    }
  }
  // Laplacian row
  __syncthreads();
  const int cl = Cd - 1;
  attn_stage(qJ, q, w, n, cl, q.off + h * dh, Cd, dh, ldh, tid, nt);
  attn_stage(kJ, k, w, n, cl, k.off + h * dh, Cd, dh, ldh, tid, nt);
  attn_stage(vJ, v, w, n, cl, v.off + h * dh, Cd, dh, ldh, tid, nt);
  attn_stage_wait();
  __syncthreads();
  for (int x = tid; x < nn; x += nt) {
    int i = x / n, j = x % n;
    float a1 = 0.f;
    for (int d = 0; d < dh; ++d) {
      a1 = fmaf(qJ[i * ldh + d], k0[j * ldh + d], a1);
      a1 = fmaf(q0[i * ldh + d], kJ[j * ldh + d], a1);
    }
    aL[x] = fmaf(a1, scale, aL[x]);
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    float ab = 0.f;
    for (int m = 0; m < n; ++m) ab = fmaf(wgt[i * n + m], aL[i * n + m], ab);
    abar[i] = ab;
  }
  __syncthreads();
  for (int x = tid; x < nn; x += nt) {
    int i = x / n;
    aJ[x] = t1[x] + wgt[x] * (aL[x] - abar[i] - t2[i]);  // wL
  }
  __syncthreads();
  for (int x = tid; x < ndh; x += nt) {
    int i = x / dh, d = x % dh;
    float acc = oL[i * ldh + d];
    for (int j = 0; j < n; ++j) {
      acc = fmaf(aJ[i * n + j], v0[j * ldh + d], acc);
      acc = fmaf(wgt[i * n + j], vJ[j * ldh + d], acc);
    }
    out[((w * n + i) * (long long)Cd + cl) * ldo + h * dh + d] = acc;
  }
}

#ifndef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// Same rule for small molecules (n <= 16 electrons, head dimension 64): the Jacobian components are independent
// given (q, k, v, w), so instead of walking them one by one with five block-wide barriers each (k_attention_fl: 2 % of
// the FP32 peak), every WARP of the block takes components c = warp, warp + 8, ... and runs the three phases on its own
// staging area with warp-level synchronisation only:
//   A  aJ = (qJ.k + q.kJ)/sqrt(d), aL += 2 qJ.kJ/sqrt(d)   lane = 2 x 4 tile of (i, j) pairs, float4 operand reads
//   B  abar, wJ, t1 += wJ (aJ - abar), t2 += sum wJ aJ       lane = row i
//   C  oJ = wJ v + w vJ (stored), oL += 2 wJ vJ             lane = features (lane, lane + 32), 8 rows per pass
// The Laplacian accumulators (aL, t1, t2, oL) are per-warp partial sums, reduced in a fixed order at the end, so the
// result is deterministic.  Shared memory (floats): q0 k0 v0 [n][68] | w [n][n] | per warp: qJ kJ vJ [n][68],
// aJ aL t1 [n][n], t2 abar [16], oL [n][64].
// ------------------------------------------------------------------------------------------------
constexpr int AW_LD = 68;      // row stride of the [n][64] operand tiles (16-byte aligned, rows 4 banks apart)
constexpr int AW_NS = 16;      // row stride of the [n][n] matrices (n <= 16): float4 reads along j
constexpr int AW_WARPS = 8;

// 16-byte asynchronous copy global -> shared (LDGSTS), or a zero fill when the source element does not exist
__device__ __forceinline__ void attn_stage4(float* dst, const JqAttnOperand& t, long long w, int n, int i, int comp_dense,
                                            int col, int Cd) {
  const float* base = t.p + ((w * n + i) * (long long)t.C) * t.ld + col;
  const float* src;
  if (t.C == Cd) src = base + (long long)comp_dense * t.ld;
  else if (comp_dense == 0) src = base;
  else if (comp_dense == Cd - 1) src = base + (long long)4 * t.ld;
  else {
    const int k = comp_dense - 1;
    src = (k / 3 == i) ? base + (long long)(1 + k % 3) * t.ld : nullptr;
  }
  if (src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  } else {
    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void attn_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void attn_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// QKL: q and k are Local1 operands (LapNet: queries / keys come from the one-electron stream, so component (e, a) of
// their Jacobian is non-zero for electron e only).  Then aJ has one non-zero row and one non-zero column, and
//   sum_j wJ_ij v_j = -abar_i o_i + w_ie aJ_ie v_e   for i != e   (o = w v, the value output),
// likewise with vJ for the Laplacian term: of the three n^2 d products per component only  w vJ  remains
// (the reference's hand rule exploits the same structure, backbone/lapnet/_attention.py:113-196).
template <bool QKL>
__global__ void __launch_bounds__(AW_WARPS * 32, 1)
k_attention_fl_warp(JqAttnOperand q, JqAttnOperand k, JqAttnOperand v, float* __restrict__ out, int ldo, int n, int H) {
  constexpr int dh = 64;
  JQ_DYN_SMEM(float, sm);
  const int nn = n * AW_NS, nt = n * AW_LD;
  float* q0 = sm;
  float* k0 = q0 + nt;
  float* v0 = k0 + nt;
  float* wgt = v0 + nt;            // [n][16]
  float* red = wgt + nn;           // block-level: aL t1 [n][16] | t2 abL [16] | oL [n][64] | o0 [n][64] (value output)
  const int red_floats = 2 * nn + 32 + 2 * n * dh;
  const int per_warp = 6 * nt + 3 * nn + 48;   // two staging sets {qJ kJ vJ} | aJ aL t1 | t2 [16] abar [16] col [16]
  float* wbase = red + red_floats;
  const long long w = blockIdx.x / H;
  const int h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Cd = 3 * n + 2, K = 3 * n;
  const float scale = 0.125f;   // 1 / sqrt(64)
  float* stage = wbase + (size_t)warp * per_warp;
  float* aJ = stage + 6 * nt;
  float* aLp = aJ + nn;
  float* t1p = aLp + nn;
  float* t2p = t1p + nn;           // [16]
  float* abar_w = t2p + 16;        // QKL: abar [16] | column el of aJ [16] of the current component
  float* o0s = red + 2 * nn + 32 + n * dh;

  // ---- value row: operands, logits, softmax, o = w v ----
  for (int x = tid; x < n * 16; x += blockDim.x) {
    const int i = x >> 4, d4 = (x & 15) * 4;
    attn_stage4(q0 + i * AW_LD + d4, q, w, n, i, 0, q.off + h * dh + d4, Cd);
    attn_stage4(k0 + i * AW_LD + d4, k, w, n, i, 0, k.off + h * dh + d4, Cd);
    attn_stage4(v0 + i * AW_LD + d4, v, w, n, i, 0, v.off + h * dh + d4, Cd);
  }
  attn_async_commit();
  // first component of this warp: staged while the block computes the softmax
  auto stage_comp = [&](int kk, int buf) {
    float* qJ = stage + buf * 3 * nt;
    if (QKL) {   // only row kk / 3 of qJ and kJ is non-zero, and phase A reads nothing else
      const int e = kk / 3;
      if (lane < 16) attn_stage4(qJ + e * AW_LD + lane * 4, q, w, n, e, 1 + kk, q.off + h * dh + lane * 4, Cd);
      else attn_stage4(qJ + nt + e * AW_LD + (lane - 16) * 4, k, w, n, e, 1 + kk, k.off + h * dh + (lane - 16) * 4, Cd);
    }
    for (int x = lane; x < n * 16; x += 32) {
      const int i = x >> 4, d4 = (x & 15) * 4;
      if (!QKL) {
        attn_stage4(qJ + i * AW_LD + d4, q, w, n, i, 1 + kk, q.off + h * dh + d4, Cd);
        attn_stage4(qJ + nt + i * AW_LD + d4, k, w, n, i, 1 + kk, k.off + h * dh + d4, Cd);
      }
      attn_stage4(qJ + 2 * nt + i * AW_LD + d4, v, w, n, i, 1 + kk, v.off + h * dh + d4, Cd);
    }
    attn_async_commit();
  };
  if (warp < K) stage_comp(warp, 0);
  for (int x = lane; x < 2 * nn + 32; x += 32) aLp[x] = 0.f;   // aL, t1 partials, t2
  attn_async_wait<1>();   // the value operands (first group) have landed
  __syncthreads();
  for (int x = tid; x < n * n; x += blockDim.x) {
    const int i = x / n, j = x - i * n;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(q0[i * AW_LD + d], k0[j * AW_LD + d], acc);
    wgt[i * AW_NS + j] = acc * scale;
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    float m = wgt[i * AW_NS];
    for (int j = 1; j < n; ++j) m = fmaxf(m, wgt[i * AW_NS + j]);
    float z = 0.f;
    for (int j = 0; j < n; ++j) {
      const float e = expf(wgt[i * AW_NS + j] - m);
      wgt[i * AW_NS + j] = e;
      z += e;
    }
    const float zi = 1.0f / z;
    for (int j = 0; j < AW_NS; ++j) wgt[i * AW_NS + j] = (j < n) ? wgt[i * AW_NS + j] * zi : 0.f;
  }
  __syncthreads();
  for (int x = tid; x < n * dh; x += blockDim.x) {
    const int i = x >> 6, d = x & 63;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = fmaf(wgt[i * AW_NS + j], v0[j * AW_LD + d], acc);
    out[((w * n + i) * (long long)Cd) * ldo + h * dh + d] = acc;
    o0s[x] = acc;
  }
  if (QKL) __syncthreads();   // o0 is read by every warp below

  // ---- Jacobian components: one per warp at a time, operands of the next one in flight (cp.async) ----
  const int ti = (lane >> 2) * 2, tj = (lane & 3) * 4;          // phase A tile: rows ti, ti+1; columns tj .. tj+3
  const bool tile_on = ti < n && tj < n;
  const int i0 = min(ti, n - 1), i1 = min(ti + 1, n - 1);
  int jr[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) jr[u] = min(tj + u, n - 1);
  float oL[AW_NS][2];   // this lane's features (lane, lane + 32) of the 2 sum_k wJ vJ term, all rows
#pragma unroll
  for (int r = 0; r < AW_NS; ++r) oL[r][0] = oL[r][1] = 0.f;
  int buf = 0;
  for (int kk = warp; kk < K; kk += AW_WARPS, buf ^= 1) {
    const int comp = 1 + kk;
    if (kk + AW_WARPS < K) {
      stage_comp(kk + AW_WARPS, buf ^ 1);
      attn_async_wait<1>();
    } else {
      attn_async_wait<0>();
    }
    __syncwarp();
    const float* qJ = stage + buf * 3 * nt;
    const float* kJ = qJ + nt;
    const float* vJ = kJ + nt;
    const int el = kk / 3;   // electron of this component (QKL: the only non-zero row of qJ / kJ)
    float col_e = 0.f;       // QKL: aJ[lane][el] before phase B overwrites aJ with wJ
    // phase A
    if (QKL) {
      // row el: aJ[el][j] = qJ_el . k_j / sqrt(d);  column el: aJ[i][el] = q_i . kJ_el / sqrt(d);  both at [el][el]
      for (int x = lane; x < nn; x += 32) aJ[x] = 0.f;
      __syncwarp();
      if (lane < n) {
        float r = 0.f, c = 0.f, rc = 0.f;
#pragma unroll 4
        for (int d4 = 0; d4 < dh; d4 += 4) {
          const float4 qe = *reinterpret_cast<const float4*>(qJ + el * AW_LD + d4);
          const float4 ke = *reinterpret_cast<const float4*>(kJ + el * AW_LD + d4);
          const float4 kj = *reinterpret_cast<const float4*>(k0 + lane * AW_LD + d4);
          const float4 qi = *reinterpret_cast<const float4*>(q0 + lane * AW_LD + d4);
          r = fmaf(qe.x, kj.x, r); r = fmaf(qe.y, kj.y, r); r = fmaf(qe.z, kj.z, r); r = fmaf(qe.w, kj.w, r);
          c = fmaf(qi.x, ke.x, c); c = fmaf(qi.y, ke.y, c); c = fmaf(qi.z, ke.z, c); c = fmaf(qi.w, ke.w, c);
          rc = fmaf(qe.x, ke.x, rc); rc = fmaf(qe.y, ke.y, rc); rc = fmaf(qe.z, ke.z, rc); rc = fmaf(qe.w, ke.w, rc);
        }
        if (lane == el) {
          aJ[el * AW_NS + el] = (r + c) * scale;
          aLp[el * AW_NS + el] = fmaf(2.0f * scale, rc, aLp[el * AW_NS + el]);
          col_e = (r + c) * scale;
        } else {
          aJ[el * AW_NS + lane] = r * scale;
          aJ[lane * AW_NS + el] = c * scale;
          col_e = c * scale;
        }
      }
    } else if (tile_on) {
      float a1[2][4], a2[2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int u = 0; u < 4; ++u) a1[r][u] = a2[r][u] = 0.f;
#pragma unroll 4
      for (int d4 = 0; d4 < dh; d4 += 4) {
        float4 qa[2], qb[2], ka[4], kb[4];
        qa[0] = *reinterpret_cast<const float4*>(q0 + i0 * AW_LD + d4);
        qa[1] = *reinterpret_cast<const float4*>(q0 + i1 * AW_LD + d4);
        qb[0] = *reinterpret_cast<const float4*>(qJ + i0 * AW_LD + d4);
        qb[1] = *reinterpret_cast<const float4*>(qJ + i1 * AW_LD + d4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ka[u] = *reinterpret_cast<const float4*>(k0 + jr[u] * AW_LD + d4);
          kb[u] = *reinterpret_cast<const float4*>(kJ + jr[u] * AW_LD + d4);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float s1 = a1[r][u], s2 = a2[r][u];
            s1 = fmaf(qb[r].x, ka[u].x, s1); s1 = fmaf(qa[r].x, kb[u].x, s1); s2 = fmaf(qb[r].x, kb[u].x, s2);
            s1 = fmaf(qb[r].y, ka[u].y, s1); s1 = fmaf(qa[r].y, kb[u].y, s1); s2 = fmaf(qb[r].y, kb[u].y, s2);
            s1 = fmaf(qb[r].z, ka[u].z, s1); s1 = fmaf(qa[r].z, kb[u].z, s1); s2 = fmaf(qb[r].z, kb[u].z, s2);
            s1 = fmaf(qb[r].w, ka[u].w, s1); s1 = fmaf(qa[r].w, kb[u].w, s1); s2 = fmaf(qb[r].w, kb[u].w, s2);
            a1[r][u] = s1;
            a2[r][u] = s2;
          }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ti + r < n && tj + u < n) {
            const int x = (ti + r) * AW_NS + tj + u;
            aJ[x] = a1[r][u] * scale;
            aLp[x] = fmaf(2.0f * scale, a2[r][u], aLp[x]);
          }
    }
    __syncwarp();
    // phase B
    if (lane < n) {
      const int i = lane;
      float ab = 0.f;
      for (int m = 0; m < n; ++m) ab = fmaf(wgt[i * AW_NS + m], aJ[i * AW_NS + m], ab);
      float t2 = 0.f;
      for (int j = 0; j < AW_NS; ++j) {
        const float a = (j < n) ? aJ[i * AW_NS + j] : 0.f;
        const float c = a - ab;
        const float wj = wgt[i * AW_NS + j] * c;   // w = 0 in the padding columns
        t1p[i * AW_NS + j] = fmaf(wj, c, t1p[i * AW_NS + j]);
        t2 = fmaf(wj, a, t2);
        aJ[i * AW_NS + j] = wj;   // aJ now holds wJ
      }
      t2p[i] += t2;
      if (QKL) {
        abar_w[i] = ab;
        abar_w[16 + i] = col_e;   // aJ[i][el]
      }
    }
    __syncwarp();
    // phase C: rows in two passes of 8, j in steps of 4 (float4 broadcasts of wJ and w)
#pragma unroll
    for (int ib = 0; ib < AW_NS; ib += 8) {
      if (ib < n) {
        float acc[8][2], acc2[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = acc2[r][0] = acc2[r][1] = 0.f;
        if (QKL) {
          // only  w vJ  is an n^2 d product; the wJ terms follow from abar, column el of aJ and row el of wJ
          for (int j4 = 0; j4 < n; j4 += 4) {
            float ja[4], jb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = min(j4 + u, n - 1);
              ja[u] = vJ[j * AW_LD + lane];
              jb[u] = vJ[j * AW_LD + lane + 32];
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const int i = min(ib + r, n - 1);
              const float4 ww4 = *reinterpret_cast<const float4*>(wgt + i * AW_NS + j4);
              const float ww[4] = {ww4.x, ww4.y, ww4.z, ww4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                acc[r][0] = fmaf(ww[u], ja[u], acc[r][0]);
                acc[r][1] = fmaf(ww[u], jb[u], acc[r][1]);
              }
            }
          }
          const float ve_a = v0[el * AW_LD + lane], ve_b = v0[el * AW_LD + lane + 32];
          const float je_a = vJ[el * AW_LD + lane], je_b = vJ[el * AW_LD + lane + 32];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int i = min(ib + r, n - 1);
            if (i != el) {
              const float ab = abar_w[i], g = wgt[i * AW_NS + el] * abar_w[16 + i];   // w_ie aJ_ie
              const float wv_a = acc[r][0], wv_b = acc[r][1];
              acc[r][0] = wv_a - ab * o0s[i * dh + lane] + g * ve_a;
              acc[r][1] = wv_b - ab * o0s[i * dh + lane + 32] + g * ve_b;
              acc2[r][0] = g * je_a - ab * wv_a;
              acc2[r][1] = g * je_b - ab * wv_b;
            } else {
              float t_a = 0.f, t_b = 0.f, u_a = 0.f, u_b = 0.f;   // row el of wJ is dense
              for (int j = 0; j < n; ++j) {
                const float wj = aJ[el * AW_NS + j];
                t_a = fmaf(wj, v0[j * AW_LD + lane], t_a);
                t_b = fmaf(wj, v0[j * AW_LD + lane + 32], t_b);
                u_a = fmaf(wj, vJ[j * AW_LD + lane], u_a);
                u_b = fmaf(wj, vJ[j * AW_LD + lane + 32], u_b);
              }
              acc[r][0] += t_a;
              acc[r][1] += t_b;
              acc2[r][0] = u_a;
              acc2[r][1] = u_b;
            }
          }
        } else {
        for (int j4 = 0; j4 < n; j4 += 4) {
            float va[4], vb[4], ja[4], jb[4];
  #pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = min(j4 + u, n - 1);   // w = wJ = 0 beyond n: the clamped operand does not contribute
              va[u] = v0[j * AW_LD + lane];
              vb[u] = v0[j * AW_LD + lane + 32];
              ja[u] = vJ[j * AW_LD + lane];
              jb[u] = vJ[j * AW_LD + lane + 32];
            }
  #pragma unroll
            for (int r = 0; r < 8; ++r) {
              const int i = min(ib + r, n - 1);
              const float4 wj4 = *reinterpret_cast<const float4*>(aJ + i * AW_NS + j4);
              const float4 ww4 = *reinterpret_cast<const float4*>(wgt + i * AW_NS + j4);
              const float wj[4] = {wj4.x, wj4.y, wj4.z, wj4.w}, ww[4] = {ww4.x, ww4.y, ww4.z, ww4.w};
  #pragma unroll
              for (int u = 0; u < 4; ++u) {
                acc[r][0] = fmaf(wj[u], va[u], acc[r][0]);
                acc[r][0] = fmaf(ww[u], ja[u], acc[r][0]);
                acc[r][1] = fmaf(wj[u], vb[u], acc[r][1]);
                acc[r][1] = fmaf(ww[u], jb[u], acc[r][1]);
                acc2[r][0] = fmaf(wj[u], ja[u], acc2[r][0]);
                acc2[r][1] = fmaf(wj[u], jb[u], acc2[r][1]);
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (ib + r < n) {
            const int i = ib + r;
            float* o = out + ((w * n + i) * (long long)Cd + comp) * ldo + h * dh;
            o[lane] = acc[r][0];
            o[lane + 32] = acc[r][1];
            oL[ib + r][0] = fmaf(2.0f, acc2[r][0], oL[ib + r][0]);
            oL[ib + r][1] = fmaf(2.0f, acc2[r][1], oL[ib + r][1]);
          }
      }
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- reduce the per-warp partial sums in a fixed order ----
  float* aL = red;
  float* t1 = aL + nn;
  float* t2 = t1 + nn;
  float* abL = t2 + 16;
  float* oLs = abL + 16;
  // oL lives in registers: each warp parks its partial in its own (now free) staging area first
#pragma unroll
  for (int r = 0; r < AW_NS; ++r)
    if (r < n) {
      stage[r * dh + lane] = oL[r][0];
      stage[r * dh + lane + 32] = oL[r][1];
    }
  __syncthreads();
  for (int x = tid; x < nn; x += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int ww = 0; ww < AW_WARPS; ++ww) {
      const float* b = wbase + (size_t)ww * per_warp + 6 * nt + nn;
      s1 += b[x];
      s2 += b[nn + x];
    }
    aL[x] = s1;
    t1[x] = s2;
  }
  for (int x = tid; x < n * dh; x += blockDim.x) {
    float s = 0.f;
    for (int ww = 0; ww < AW_WARPS; ++ww) s += (wbase + (size_t)ww * per_warp)[x];
    oLs[x] = s;
  }
  if (tid < n) {
    float s = 0.f;
    for (int ww = 0; ww < AW_WARPS; ++ww) s += (wbase + (size_t)ww * per_warp + 6 * nt + 3 * nn)[tid];
    t2[tid] = s;
  }
  __syncthreads();
  // ---- Laplacian row ----
  float* qL = wbase + 3 * nt;   // warp 0's second staging set (not read by the reductions above)
  float* kL = qL + nt;
  float* vL = kL + nt;
  float* wL = wbase + 6 * nt;   // warp 0's aJ
  const int cl = Cd - 1;
  for (int x = tid; x < n * 16; x += blockDim.x) {
    const int i = x >> 4, d4 = (x & 15) * 4;
    attn_stage4(qL + i * AW_LD + d4, q, w, n, i, cl, q.off + h * dh + d4, Cd);
    attn_stage4(kL + i * AW_LD + d4, k, w, n, i, cl, k.off + h * dh + d4, Cd);
    attn_stage4(vL + i * AW_LD + d4, v, w, n, i, cl, v.off + h * dh + d4, Cd);
  }
  attn_async_commit();
  attn_async_wait<0>();
  __syncthreads();
  for (int x = tid; x < n * n; x += blockDim.x) {
    const int i = x / n, j = x - i * n;
    float a1 = 0.f;
    for (int d = 0; d < dh; ++d) {
      a1 = fmaf(qL[i * AW_LD + d], k0[j * AW_LD + d], a1);
      a1 = fmaf(q0[i * AW_LD + d], kL[j * AW_LD + d], a1);
    }
    aL[i * AW_NS + j] = fmaf(a1, scale, aL[i * AW_NS + j]);
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    float ab = 0.f;
    for (int m = 0; m < n; ++m) ab = fmaf(wgt[i * AW_NS + m], aL[i * AW_NS + m], ab);
    abL[i] = ab;
  }
  __syncthreads();
  for (int x = tid; x < n * n; x += blockDim.x) {
    const int i = x / n, j = x - i * n;
    wL[i * AW_NS + j] = t1[i * AW_NS + j] + wgt[i * AW_NS + j] * (aL[i * AW_NS + j] - abL[i] - t2[i]);
  }
  __syncthreads();
  for (int x = tid; x < n * dh; x += blockDim.x) {
    const int i = x >> 6, d = x & 63;
    float acc = oLs[x];
    for (int j = 0; j < n; ++j) {
      acc = fmaf(wL[i * AW_NS + j], v0[j * AW_LD + d], acc);
      acc = fmaf(wgt[i * AW_NS + j], vL[j * AW_LD + d], acc);
    }
    out[((w * n + i) * (long long)Cd + cl) * ldo + h * dh + d] = acc;
  }
}
#endif


#ifndef JAQMC_HOST_EMU
// ------------------------------------------------------------------------------------------------
// Same rule on the tensor cores for 16 < n <= 48, head dimension 64 (r2): mma.sync.m16n8k8 TF32 with the 3xTF32 split
// (a = a_hi + a_lo, three products, FP32 accumulation), so the products keep FP32-level accuracy.
// One block per (walker, head); a block holds SLOTS component slots of RT = ceil(n / 16) warps each.  Warp (slot, r)
// owns rows 16r .. 16r+15 of every [n x n] / [n x d] matrix of the slot's current component and ALL their columns, so
// the three phases run in its registers with no block-level barrier:
//   A  aJ = (qJ k^T + q kJ^T)/sqrt(d), a2 = qJ kJ^T          A operands: qJ rows straight from global, q from shared;
//                                                              B operands: k, kJ rows from shared
//   B  abar (quad shuffles), wJ = w (aJ - abar), X += wJ (aJ - abar) + 2 w a2 / sqrt(d)
//   C  oJ = wJ v + w vJ (stored), oL2 += wJ vJ                the accumulator fragment of wJ IS the A fragment of the
//      next product once the contraction index is permuted (slot t <-> column 2t, slot t+4 <-> column 2t+1); the B
//      fragments of v / vJ are read with the same permutation.
// The Laplacian row needs only X and oL2:  with Y = X + w aL' (aL' = (qL k^T + q kL^T)/sqrt(d)),
//   wL = Y - w rowsum(Y)   (the t2 term of the SIMT kernels cancels: sum_j wJ_ij = 0),   oL = 2 oL2 + wL v + w vL.
// The RT warps of a slot share the staged kJ / vJ tiles (cp.async, named barrier per slot).
// Shared (floats): q0 k0 v0 [NP][68] | w [NP][56] | per slot: kJ vJ [NP][68], NP = 16 RT the padded electron count.
// Instantiated as <NJ, NP, SLOTS> = <2, 16, 8> (n <= 16, two blocks per SM), <4, 32, 4>, <6, 48, 4>.
// ------------------------------------------------------------------------------------------------
constexpr int AM_LD = 68;     // 68 mod 32 = 4: the (g, t) fragment loads touch 32 distinct banks
constexpr int AM_LW = 56;     // 56 mod 32 = 24: the float2 accumulator-layout loads of w are conflict-free per half-warp
constexpr int AM_NP = 48;     // largest padded electron count

// x = hi + lo for the 3xTF32 products.  RNA: both parts rounded to nearest with cvt.rna.tf32.f32 -- which sm_100a has no
// single instruction for: ncu shows ~11 integer / predicate instructions per split, 63 % of everything the first version
// of this kernel issued, against 14.5 % HMMA.  !RNA: hi = x with the 13 low mantissa bits cleared (one LOP3),
// lo = x - hi (exact, one FADD) handed to the tensor core as is, which ignores its low 13 bits: two instructions, at the
// price of a one-sided 2^-22 relative truncation of each operand (measured error: DESIGN.md section 4).
template <bool RNA>
__device__ __forceinline__ void am_split(float x, unsigned& hi, unsigned& lo) {
  if (RNA) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
  } else {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
  }
}
__device__ __forceinline__ void am_mma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// c += (ah + al)(bh + bl) without the al bl term; small products first
__device__ __forceinline__ void am_mma3(float (&c)[4], const unsigned (&ah)[4], const unsigned (&al)[4],
                                        const unsigned (&bh)[2], const unsigned (&bl)[2]) {
  am_mma(c, al, bh);
  am_mma(c, ah, bl);
  am_mma(c, ah, bh);
}
// pointer to (walker w, electron i, dense component comp) of one head of an operand, or null where it is zero
__device__ __forceinline__ const float* am_row(const JqAttnOperand& t, long long w, int n, int i, int comp, int col, int Cd) {
  if (i >= n) return nullptr;
  const float* base = t.p + ((w * n + i) * (long long)t.C) * t.ld + col;
  if (t.C == Cd) return base + (long long)comp * t.ld;
  if (comp == 0) return base;
  if (comp == Cd - 1) return base + (long long)4 * t.ld;
  const int k = comp - 1;
  return (k / 3 == i) ? base + (long long)(1 + k % 3) * t.ld : nullptr;
}

template <int NJ, int NP, int SLOTS, bool RNA, bool QKL>
__global__ void __launch_bounds__(SLOTS * (NP / 16) * 32, NP == 16 ? 2 : 1)
k_attention_fl_mma(JqAttnOperand q, JqAttnOperand k, JqAttnOperand v, float* __restrict__ out, int ldo, int n, int H) {
  constexpr int dh = 64;
  JQ_DYN_SMEM(float, sm);
  constexpr int NT = NP * AM_LD;
  float* q0 = sm;
  float* k0 = q0 + NT;
  float* v0 = k0 + NT;
  float* wgt = v0 + NT;                   // [48][56]
  float* slots = wgt + NP * AM_LW;     // per slot: kJ | vJ
  const long long w = blockIdx.x / H;
  const int h = blockIdx.x % H;
  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
  constexpr int RT = NP / 16;
  const int slot = warp / RT, r = warp - slot * RT;
  const int g = lane >> 2, t = lane & 3;
  const int Cd = 3 * n + 2, K = 3 * n;
  const float scale = 0.125f;   // 1 / sqrt(64)
  float* kJs = slots + (size_t)slot * 2 * NT;
  float* vJs = kJs + NT;

  // ---- value row: operands (zero padded), logits, softmax, o = w v ----
  for (int x = tid; x < 3 * NT + NP * AM_LW + SLOTS * 2 * NT; x += nthr) sm[x] = 0.f;
  __syncthreads();
  for (int x = tid; x < n * 16; x += nthr) {
    const int i = x >> 4, d4 = (x & 15) * 4;
    attn_stage4(q0 + i * AM_LD + d4, q, w, n, i, 0, q.off + h * dh + d4, Cd);
    attn_stage4(k0 + i * AM_LD + d4, k, w, n, i, 0, k.off + h * dh + d4, Cd);
    attn_stage4(v0 + i * AM_LD + d4, v, w, n, i, 0, v.off + h * dh + d4, Cd);
  }
  attn_async_commit();
  attn_async_wait<0>();
  __syncthreads();
  for (int x = tid; x < n * n; x += nthr) {
    const int i = x / n, j = x - i * n;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(q0[i * AM_LD + d], k0[j * AM_LD + d], acc);
    wgt[i * AM_LW + j] = acc * scale;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nthr) {
    float m = wgt[i * AM_LW];
    for (int j = 1; j < n; ++j) m = fmaxf(m, wgt[i * AM_LW + j]);
    float z = 0.f;
    for (int j = 0; j < n; ++j) {
      const float e = expf(wgt[i * AM_LW + j] - m);
      wgt[i * AM_LW + j] = e;
      z += e;
    }
    const float zi = 1.0f / z;
    for (int j = 0; j < n; ++j) wgt[i * AM_LW + j] *= zi;
  }
  __syncthreads();
  for (int x = tid; x < n * dh; x += nthr) {
    const int i = x >> 6, d = x & 63;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = fmaf(wgt[i * AM_LW + j], v0[j * AM_LD + d], acc);
    out[((w * n + i) * (long long)Cd) * ldo + h * dh + d] = acc;
  }

  // ---- Jacobian components ----
  const int i0 = 16 * r + g, i1 = i0 + 8;            // this lane's fragment rows
  float X[NJ][4], oL2[8][4];
#pragma unroll
  for (int a = 0; a < NJ; ++a) X[a][0] = X[a][1] = X[a][2] = X[a][3] = 0.f;
#pragma unroll
  for (int a = 0; a < 8; ++a) oL2[a][0] = oL2[a][1] = oL2[a][2] = oL2[a][3] = 0.f;
  const int sthreads = 32 * RT, sl = r * 32 + lane;
  const bool slot_on = slot < SLOTS;
  if (slot_on) {
    // softmax weights of this warp's rows in accumulator layout: re-read from shared memory where they are used (24
    // registers fewer than keeping them; the first version of this kernel spilled 300 bytes per thread)
    const float* wrow0 = wgt + i0 * AM_LW + 2 * t;
    const float* wrow1 = wgt + i1 * AM_LW + 2 * t;
    // (Measured r2: fetching kJ of the next component during phases B / C and vJ during phase A -- one more named
    // barrier per component -- changed the launch time by 1-2 %, inside the box-to-box spread; the simpler form is kept.)
    for (int kk = slot; kk < K; kk += SLOTS) {
      const int comp = 1 + kk;
      // (a slot is a single warp when NP == 16: __syncwarp instead of a named barrier)
      if (RT == 1) __syncwarp();
      else asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(sthreads) : "memory");   // previous tiles fully consumed
      float aJ[NJ][4], a2[NJ][4];
#pragma unroll
      for (int a = 0; a < NJ; ++a) aJ[a][0] = aJ[a][1] = aJ[a][2] = aJ[a][3] = a2[a][0] = a2[a][1] = a2[a][2] = a2[a][3] = 0.f;
      if constexpr (QKL) {
        // One-electron queries / keys (LapNet, backbone/lapnet/_attention.py:113-196): component (el, axis) of qJ and kJ is
        // non-zero for electron el only, so the logit Jacobian has one non-zero row and one non-zero column,
        //   aJ[el][j] = qJ_el . k_j,   aJ[i][el] = q_i . kJ_el,   a2[el][el] = qJ_el . kJ_el      (before the 1/sqrt(d)),
        // i.e. 16 RT + 16 dot products of length 64 on the CUDA cores instead of 144 HMMAs per row tile; only vJ is a
        // full tile.  The two rows travel in the first two rows of the slot's kJ area.
        const int el = kk / 3;
        for (int x = sl; x < n * 16; x += sthreads) {
          const int i = x >> 4, d4 = (x & 15) * 4;
          attn_stage4(vJs + i * AM_LD + d4, v, w, n, i, comp, v.off + h * dh + d4, Cd);
        }
        if (sl < 16) attn_stage4(kJs + sl * 4, q, w, n, el, comp, q.off + h * dh + sl * 4, Cd);
        else if (sl < 32) attn_stage4(kJs + AM_LD + (sl - 16) * 4, k, w, n, el, comp, k.off + h * dh + (sl - 16) * 4, Cd);
        attn_async_commit();
        attn_async_wait<0>();
        if (RT == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(sthreads) : "memory");
        const float4* qe4 = reinterpret_cast<const float4*>(kJs);
        const float4* ke4 = reinterpret_cast<const float4*>(kJs + AM_LD);
        auto dot64 = [](const float4* x4, const float4* y4) {
          float acc = 0.f;
#pragma unroll
          for (int d4 = 0; d4 < 16; ++d4) {
            const float4 xv = x4[d4], yv = y4[d4];
            acc = fmaf(xv.x, yv.x, acc); acc = fmaf(xv.y, yv.y, acc); acc = fmaf(xv.z, yv.z, acc); acc = fmaf(xv.w, yv.w, acc);
          }
          return acc;
        };
        float cval = 0.f, rv[RT];
#pragma unroll
        for (int m = 0; m < RT; ++m) rv[m] = 0.f;
        if (lane < 16) {
          cval = dot64(reinterpret_cast<const float4*>(q0 + (16 * r + lane) * AM_LD), ke4);     // column el, row 16 r + lane
        } else {
#pragma unroll
          for (int m = 0; m < RT; ++m)
            rv[m] = dot64(qe4, reinterpret_cast<const float4*>(k0 + (lane - 16 + 16 * m) * AM_LD));   // row el, column
        }
        float dee = 0.f;
        if (lane < 16) {
          const float4 xv = qe4[lane], yv = ke4[lane];
          dee = xv.x * yv.x + xv.y * yv.y + xv.z * yv.z + xv.w * yv.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dee += __shfl_xor_sync(0xffffffffu, dee, o);
        const float c0 = __shfl_sync(0xffffffffu, cval, g), c1 = __shfl_sync(0xffffffffu, cval, g + 8);
#pragma unroll
        for (int a = 0; a < NJ; ++a) {
          const int colA = 8 * a + 2 * t, colB = colA + 1;
          const float rA = __shfl_sync(0xffffffffu, rv[a / 2], 16 + (colA & 15));
          const float rB = __shfl_sync(0xffffffffu, rv[a / 2], 16 + (colB & 15));
          aJ[a][0] = (i0 == el ? rA : 0.f) + (colA == el ? c0 : 0.f);
          aJ[a][1] = (i0 == el ? rB : 0.f) + (colB == el ? c0 : 0.f);
          aJ[a][2] = (i1 == el ? rA : 0.f) + (colA == el ? c1 : 0.f);
          aJ[a][3] = (i1 == el ? rB : 0.f) + (colB == el ? c1 : 0.f);
          a2[a][0] = (i0 == el && colA == el) ? dee : 0.f;
          a2[a][1] = (i0 == el && colB == el) ? dee : 0.f;
          a2[a][2] = (i1 == el && colA == el) ? dee : 0.f;
          a2[a][3] = (i1 == el && colB == el) ? dee : 0.f;
        }
      } else {
        for (int x = sl; x < n * 16; x += sthreads) {
          const int i = x >> 4, d4 = (x & 15) * 4;
          attn_stage4(kJs + i * AM_LD + d4, k, w, n, i, comp, k.off + h * dh + d4, Cd);
          attn_stage4(vJs + i * AM_LD + d4, v, w, n, i, comp, v.off + h * dh + d4, Cd);
        }
        attn_async_commit();
        const float* pJ0 = am_row(q, w, n, i0, comp, q.off + h * dh, Cd);
        const float* pJ1 = am_row(q, w, n, i1, comp, q.off + h * dh, Cd);
        // qJ fragments straight from global.  NP == 16 (few accumulators): all eight k-steps are requested here, before
        // the wait on the staged tiles (r2 profile of this variant: 17 % of the stall samples at the first use of a
        // fragment fetched one step ahead); the larger variants are at their register cap and keep the one-step prefetch.
        // Measured: 4.11-4.19 -> 4.09 ms (Psiformer-N2), 3.46 -> 3.42 ms (LapNet-N2) at equal dense-kernel times: ~1 %, the
        // stalls moved elsewhere (tensor pipe 41 % active, 7.8 instructions per HMMA with only two column tiles per A fragment).
        constexpr int XQ = (NP == 16) ? dh / 8 : 1;
        float xq[XQ][4];
  #pragma unroll
        for (int u = 0; u < XQ; ++u) {
          const int c = 8 * u + t;
          xq[u][0] = pJ0 ? pJ0[c] : 0.f; xq[u][1] = pJ1 ? pJ1[c] : 0.f; xq[u][2] = pJ0 ? pJ0[c + 4] : 0.f; xq[u][3] = pJ1 ? pJ1[c + 4] : 0.f;
        }
        attn_async_wait<0>();
        if (RT == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(sthreads) : "memory");
        // phase A
  #pragma unroll (NP == 16 ? 8 : 2)
        for (int ks = 0; ks < dh / 8; ++ks) {
          unsigned qJh[4], qJl[4], q0h[4], q0l[4];
  #pragma unroll
          for (int e = 0; e < 4; ++e) am_split<RNA>(xq[XQ == 1 ? 0 : ks][e], qJh[e], qJl[e]);
          if (XQ == 1 && ks + 1 < dh / 8) {
            const int c = 8 * (ks + 1) + t;
            xq[0][0] = pJ0 ? pJ0[c] : 0.f; xq[0][1] = pJ1 ? pJ1[c] : 0.f; xq[0][2] = pJ0 ? pJ0[c + 4] : 0.f; xq[0][3] = pJ1 ? pJ1[c + 4] : 0.f;
          }
          am_split<RNA>(q0[i0 * AM_LD + 8 * ks + t], q0h[0], q0l[0]);
          am_split<RNA>(q0[i1 * AM_LD + 8 * ks + t], q0h[1], q0l[1]);
          am_split<RNA>(q0[i0 * AM_LD + 8 * ks + t + 4], q0h[2], q0l[2]);
          am_split<RNA>(q0[i1 * AM_LD + 8 * ks + t + 4], q0h[3], q0l[3]);
  #pragma unroll
          for (int a = 0; a < NJ; ++a) {
            unsigned k0h[2], k0l[2], kJh[2], kJl[2];
            const int o = (8 * a + g) * AM_LD + 8 * ks + t;
            am_split<RNA>(k0[o], k0h[0], k0l[0]);
            am_split<RNA>(k0[o + 4], k0h[1], k0l[1]);
            am_split<RNA>(kJs[o], kJh[0], kJl[0]);
            am_split<RNA>(kJs[o + 4], kJh[1], kJl[1]);
            am_mma3(aJ[a], qJh, qJl, k0h, k0l);
            am_mma3(aJ[a], q0h, q0l, kJh, kJl);
            am_mma3(a2[a], qJh, qJl, kJh, kJl);
          }
        }
      }
      // phase B (registers): rows i0 (elements 0, 1) and i1 (elements 2, 3)
      float ab0 = 0.f, ab1 = 0.f;
#pragma unroll
      for (int a = 0; a < NJ; ++a) {
#pragma unroll
        for (int e = 0; e < 4; ++e) aJ[a][e] *= scale;
        const float2 u0 = *reinterpret_cast<const float2*>(wrow0 + 8 * a);
        const float2 u1 = *reinterpret_cast<const float2*>(wrow1 + 8 * a);
        ab0 = fmaf(u0.x, aJ[a][0], ab0);
        ab0 = fmaf(u0.y, aJ[a][1], ab0);
        ab1 = fmaf(u1.x, aJ[a][2], ab1);
        ab1 = fmaf(u1.y, aJ[a][3], ab1);
      }
      ab0 += __shfl_xor_sync(0xffffffffu, ab0, 1);
      ab0 += __shfl_xor_sync(0xffffffffu, ab0, 2);
      ab1 += __shfl_xor_sync(0xffffffffu, ab1, 1);
      ab1 += __shfl_xor_sync(0xffffffffu, ab1, 2);
#pragma unroll
      for (int a = 0; a < NJ; ++a) {
        const float2 u0 = *reinterpret_cast<const float2*>(wrow0 + 8 * a);
        const float2 u1 = *reinterpret_cast<const float2*>(wrow1 + 8 * a);
        const float wv[4] = {u0.x, u0.y, u1.x, u1.y};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float c = aJ[a][e] - (e < 2 ? ab0 : ab1);
          const float wj = wv[e] * c;
          X[a][e] = fmaf(wj, c, X[a][e]);
          X[a][e] = fmaf(2.0f * scale * wv[e], a2[a][e], X[a][e]);
          aJ[a][e] = wj;   // aJ now holds wJ
        }
      }
      // phase C: contraction over j in steps of 8 (slot t <-> j = 8 ks + 2t, slot t+4 <-> j = 8 ks + 2t + 1)
      float oJ[8][4];
#pragma unroll
      for (int a = 0; a < 8; ++a) oJ[a][0] = oJ[a][1] = oJ[a][2] = oJ[a][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < NJ; ++ks) {
        unsigned wJh[4], wJl[4], wh[4], wl[4];
        am_split<RNA>(aJ[ks][0], wJh[0], wJl[0]);
        am_split<RNA>(aJ[ks][2], wJh[1], wJl[1]);
        am_split<RNA>(aJ[ks][1], wJh[2], wJl[2]);
        am_split<RNA>(aJ[ks][3], wJh[3], wJl[3]);
        {
          const float2 u0 = *reinterpret_cast<const float2*>(wrow0 + 8 * ks);
          const float2 u1 = *reinterpret_cast<const float2*>(wrow1 + 8 * ks);
          am_split<RNA>(u0.x, wh[0], wl[0]);
          am_split<RNA>(u1.x, wh[1], wl[1]);
          am_split<RNA>(u0.y, wh[2], wl[2]);
          am_split<RNA>(u1.y, wh[3], wl[3]);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          unsigned v0h[2], v0l[2], vJh[2], vJl[2];
          const int o = (8 * ks + 2 * t) * AM_LD + 8 * a + g;
          am_split<RNA>(v0[o], v0h[0], v0l[0]);
          am_split<RNA>(v0[o + AM_LD], v0h[1], v0l[1]);
          am_split<RNA>(vJs[o], vJh[0], vJl[0]);
          am_split<RNA>(vJs[o + AM_LD], vJh[1], vJl[1]);
          am_mma3(oJ[a], wJh, wJl, v0h, v0l);
          am_mma3(oJ[a], wh, wl, vJh, vJl);
          am_mma3(oL2[a], wJh, wJl, vJh, vJl);
        }
      }
      if (i0 < n) {
        float* o = out + ((w * n + i0) * (long long)Cd + comp) * ldo + h * dh + 2 * t;
#pragma unroll
        for (int a = 0; a < 8; ++a) *reinterpret_cast<float2*>(o + 8 * a) = make_float2(oJ[a][0], oJ[a][1]);
      }
      if (i1 < n) {
        float* o = out + ((w * n + i1) * (long long)Cd + comp) * ldo + h * dh + 2 * t;
#pragma unroll
        for (int a = 0; a < 8; ++a) *reinterpret_cast<float2*>(o + 8 * a) = make_float2(oJ[a][2], oJ[a][3]);
      }
    }
  }
  __syncthreads();

  // ---- per-warp partial X and oL2 -> the slot's (now free) tiles, then a fixed-order sum over the slots ----
  if (slot_on) {
#pragma unroll
    for (int a = 0; a < NJ; ++a) {
      *reinterpret_cast<float2*>(kJs + i0 * AM_LW + 8 * a + 2 * t) = make_float2(X[a][0], X[a][1]);
      *reinterpret_cast<float2*>(kJs + i1 * AM_LW + 8 * a + 2 * t) = make_float2(X[a][2], X[a][3]);
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      *reinterpret_cast<float2*>(vJs + i0 * AM_LD + 8 * a + 2 * t) = make_float2(oL2[a][0], oL2[a][1]);
      *reinterpret_cast<float2*>(vJs + i1 * AM_LD + 8 * a + 2 * t) = make_float2(oL2[a][2], oL2[a][3]);
    }
  }
  __syncthreads();
  float* Xs = slots;            // slot 0: X [48][56] | oL2 [48][68]
  float* oLs = slots + NT;
  for (int x = tid; x < NP * AM_LW; x += nthr) {
    float s2 = 0.f;
    for (int ss = 0; ss < SLOTS; ++ss) s2 += slots[(size_t)ss * 2 * NT + x];
    Xs[x] = s2;
  }
  for (int x = tid; x < NT; x += nthr) {
    float s2 = 0.f;
    for (int ss = 0; ss < SLOTS; ++ss) s2 += slots[(size_t)ss * 2 * NT + NT + x];
    oLs[x] = s2;
  }
  __syncthreads();
  // ---- Laplacian row ----
  float* qL = slots + 2 * NT;   // slot 1
  float* kL = qL + NT;
  float* vL = kL + NT;          // slot 2
  float* rs = vL + NT;          // row sums of Y [48]
  const int cl = Cd - 1;
  for (int x = tid; x < n * 16; x += nthr) {
    const int i = x >> 4, d4 = (x & 15) * 4;
    attn_stage4(qL + i * AM_LD + d4, q, w, n, i, cl, q.off + h * dh + d4, Cd);
    attn_stage4(kL + i * AM_LD + d4, k, w, n, i, cl, k.off + h * dh + d4, Cd);
    attn_stage4(vL + i * AM_LD + d4, v, w, n, i, cl, v.off + h * dh + d4, Cd);
  }
  attn_async_commit();
  attn_async_wait<0>();
  __syncthreads();
  for (int x = tid; x < n * n; x += nthr) {
    const int i = x / n, j = x - i * n;
    float a1 = 0.f;
    for (int d = 0; d < dh; ++d) {
      a1 = fmaf(qL[i * AM_LD + d], k0[j * AM_LD + d], a1);
      a1 = fmaf(q0[i * AM_LD + d], kL[j * AM_LD + d], a1);
    }
    Xs[i * AM_LW + j] = fmaf(wgt[i * AM_LW + j] * scale, a1, Xs[i * AM_LW + j]);   // Y
  }
  __syncthreads();
  for (int i = tid; i < n; i += nthr) {
    float s2 = 0.f;
    for (int j = 0; j < n; ++j) s2 += Xs[i * AM_LW + j];
    rs[i] = s2;
  }
  __syncthreads();
  for (int x = tid; x < n * n; x += nthr) {
    const int i = x / n, j = x - i * n;
    Xs[i * AM_LW + j] -= wgt[i * AM_LW + j] * rs[i];   // wL
  }
  __syncthreads();
  for (int x = tid; x < n * dh; x += nthr) {
    const int i = x >> 6, d = x & 63;
    float acc = 2.0f * oLs[i * AM_LD + d];
    for (int j = 0; j < n; ++j) {
      acc = fmaf(Xs[i * AM_LW + j], v0[j * AM_LD + d], acc);
      acc = fmaf(wgt[i * AM_LW + j], vL[j * AM_LD + d], acc);
    }
    out[((w * n + i) * (long long)Cd + cl) * ldo + h * dh + d] = acc;
  }
}
#endif

int jq_launch_attention_fl(const JqAttnOperand& q, const JqAttnOperand& k, const JqAttnOperand& v, float* out, int ldo,
                           long long W, int n, int H, int dh, int track, cudaStream_t st) {
  return jq_launch_attention_fl_sel(q, k, v, out, ldo, W, n, H, dh, track, 0, st);
}

int jq_launch_attention_fl_sel(const JqAttnOperand& q, const JqAttnOperand& k, const JqAttnOperand& v, float* out, int ldo,
                               long long W, int n, int H, int dh, int track, int force, cudaStream_t st) {
  if (W <= 0) return JQ_OK;
  const int Cd = track ? 3 * n + 2 : 1;
  JQ_REQUIRE(v.C == Cd && (q.C == Cd || (track && q.C == 5)) && (k.C == Cd || (track && k.C == 5)),
             JQ_ERR_INVALID_ARGUMENT, "attention: operand components %d/%d/%d do not match %d", q.C, k.C, v.C, Cd);
  size_t smem = sizeof(float) * ((size_t)7 * n * (dh + 1) + (size_t)5 * n * n + 2 * n);
  JQ_REQUIRE(smem <= 220 * 1024, JQ_ERR_UNSUPPORTED, "attention: n=%d head_dim=%d need %zu bytes of shared memory", n, dh,
             smem);
#ifndef JAQMC_HOST_EMU
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_attention_fl<4, 4, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attention_fl<2, 4, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attention_fl<2, 2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
#endif
  double K = track ? 3.0 * n + 1.0 : 0.0;
  jq_prof_work((double)W * H * (4.0 * n * n * dh * (1.0 + 2.0 * K)), 4.0 * (double)W * n * Cd * H * dh * 4);
#ifndef JAQMC_HOST_EMU
  {
    static const bool env_block = getenv("JAQMC_B200_ATTENTION_BLOCK") != nullptr;   // A/B switch
    const bool old_kernel = force ? force == 1 : env_block;
    const bool aligned = (q.ld % 4 == 0) && (k.ld % 4 == 0) && (v.ld % 4 == 0) && (q.off % 4 == 0) && (k.off % 4 == 0) &&
                         (v.off % 4 == 0) && ((reinterpret_cast<uintptr_t>(q.p) | reinterpret_cast<uintptr_t>(k.p) |
                                               reinterpret_cast<uintptr_t>(v.p)) % 16 == 0);
    const int nt = n * AW_LD, nn = n * AW_NS;
    const size_t sw = sizeof(float) * ((size_t)3 * nt + nn + (2 * nn + 32 + 2 * n * 64) + (size_t)AW_WARPS * (6 * nt + 3 * nn + 48));
    static const bool env_warp = getenv("JAQMC_B200_ATTENTION_WARP") != nullptr;   // A/B switch: warp kernel for n <= 14
    const bool mma_ok = dh == 64 && n >= 2 && n <= AM_NP && ldo % 2 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0;
    const bool prefer_warp = force == 2 || env_warp || !mma_ok || getenv("JAQMC_B200_ATTENTION_SIMT") != nullptr;
    if (track && dh == 64 && n >= 2 && n <= AW_NS && sw <= 227 * 1024 && aligned && !old_kernel && (force == 0 || force == 2) && prefer_warp) {
      const bool qkl = (q.C == 5 && k.C == 5 && Cd != 5);
      cudaError_t e = qkl ? cudaFuncSetAttribute(k_attention_fl_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw)
                          : cudaFuncSetAttribute(k_attention_fl_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw);
      JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      if (qkl) JQ_LAUNCH(k_attention_fl_warp<true>, dim3((unsigned)(W * H)), dim3(AW_WARPS * 32), sw, st, q, k, v, out, ldo, n, H);
      else JQ_LAUNCH(k_attention_fl_warp<false>, dim3((unsigned)(W * H)), dim3(AW_WARPS * 32), sw, st, q, k, v, out, ldo, n, H);
      JQ_CHECK_LAUNCH();
      return JQ_OK;
    }
    // n <= 48 and not taken by the warp kernel above: the tensor-core kernel (A/B switch: JAQMC_B200_ATTENTION_SIMT keeps the CUDA-core block kernel)
    static const bool env_simt = getenv("JAQMC_B200_ATTENTION_SIMT") != nullptr;
    const bool simt_kernel = force ? (force != 3 && force != 4 && force != 5) : env_simt;
    // (also n = 15, 16, where the warp kernel's per-warp staging no longer fits in shared memory)
    if (track && dh == 64 && n >= 2 && n <= AM_NP && aligned && ldo % 2 == 0 &&
        reinterpret_cast<uintptr_t>(out) % 8 == 0 && !old_kernel && !simt_kernel) {
      static const bool env_rna = getenv("JAQMC_B200_ATTENTION_RNA") != nullptr;   // A/B switch: round-to-nearest operand split
      const bool rna = force == 4 || (force == 0 && env_rna);
      const int NP = 16 * ((n + 15) / 16), SL = NP == 16 ? 8 : 4;
      const size_t sm_mma = sizeof(float) * ((size_t)3 * NP * AM_LD + NP * AM_LW + (size_t)SL * 2 * NP * AM_LD);
      const dim3 grid((unsigned)(W * H)), block(32 * (NP / 16) * SL);
#define JQ_AM_LAUNCH(NJ_, NP_, SL_, RNA_, QKL_)                                                                               \
  do {                                                                                                                   \
    static JqPerDeviceFlag attr_set;                                                                                     \
    const int dev = jq_current_device();                                                                                 \
    if (!attr_set.done[dev]) {                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(k_attention_fl_mma<NJ_, NP_, SL_, RNA_, QKL_>,                                \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_mma);                   \
      JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));           \
      attr_set.done[dev] = true;                                                                                         \
    }                                                                                                                    \
    JQ_LAUNCH((k_attention_fl_mma<NJ_, NP_, SL_, RNA_, QKL_>), grid, block, sm_mma, st, q, k, v, out, ldo, n, H);       \
  } while (0)
      // one-electron queries / keys (LapNet): sparse logit-Jacobian phase (A/B switch JAQMC_B200_ATTENTION_DENSE_QK)
      static const bool env_dense_qk = getenv("JAQMC_B200_ATTENTION_DENSE_QK") != nullptr;
      const bool qkl = q.C == 5 && k.C == 5 && Cd != 5 && !rna && ((force == 0 && !env_dense_qk) || force == 5);
      JQ_REQUIRE(force != 5 || qkl, JQ_ERR_UNSUPPORTED, "attention: kernel 5 needs one-electron (5-component) q and k operands");
      if (NP == 16 && qkl) JQ_AM_LAUNCH(2, 16, 8, false, true);
      else if (NP == 32 && qkl) JQ_AM_LAUNCH(4, 32, 4, false, true);
      else if (qkl) JQ_AM_LAUNCH(6, 48, 4, false, true);
      else if (NP == 16 && !rna) JQ_AM_LAUNCH(2, 16, 8, false, false);
      else if (NP == 16) JQ_AM_LAUNCH(2, 16, 8, true, false);
      else if (NP == 32 && !rna) JQ_AM_LAUNCH(4, 32, 4, false, false);
      else if (NP == 32) JQ_AM_LAUNCH(4, 32, 4, true, false);
      else if (!rna) JQ_AM_LAUNCH(6, 48, 4, false, false);
      else JQ_AM_LAUNCH(6, 48, 4, true, false);
#undef JQ_AM_LAUNCH
      JQ_CHECK_LAUNCH();
      return JQ_OK;
    }
  }
#endif
  JQ_REQUIRE(force == 0 || force == 1, JQ_ERR_UNSUPPORTED, "attention: kernel %d does not support n=%d head_dim=%d", force, n, dh);
#ifndef JAQMC_HOST_EMU
  static const int tile_sel = getenv("JAQMC_B200_ATTENTION_TILES") ? atoi(getenv("JAQMC_B200_ATTENTION_TILES")) : 0;   // tuning
  if (tile_sel == 1) JQ_LAUNCH((k_attention_fl<2, 4, 2, 4>), dim3((unsigned)(W * H)), dim3(256), smem, st, q, k, v, out, ldo, n, H, dh, track);
  else if (tile_sel == 2) JQ_LAUNCH((k_attention_fl<2, 2, 2, 2>), dim3((unsigned)(W * H)), dim3(256), smem, st, q, k, v, out, ldo, n, H, dh, track);
  else
#endif
  JQ_LAUNCH((k_attention_fl<4, 4, 4, 4>), dim3((unsigned)(W * H)), dim3(256), smem, st, q, k, v, out, ldo, n, H, dh, track);
  JQ_CHECK_LAUNCH();
  return JQ_OK;
}
