// Augmented-row dense layer on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with the
// accumulator in TMEM, operands staged in shared memory by TMA, FP32-faithful 3xTF32 split arithmetic, and the
// bias / per-walker addend / tanh forward-Laplacian / residual epilogue fused.
//
// Orientation ("swap AB"): the MMA's M dimension is the layer's OUTPUT FEATURES and its N dimension is the
// tile's ROWS {x, J_1..J_K, L} of a few whole (walker, electron) groups:
//     D[f, r] = sum_k Wt[f, k] * X[r, k]          A = Wt (out x in, K-major), B = X (rows x in, K-major)
// so a TMEM lane holds one output feature and its columns are the rows of the groups.  The epilogue thread that
// owns lane f reads x, every J_k and L of a group from its own lane: the tanh rule's sum_k J_k^2 is a private
// serial reduction (no shuffles), and for a fixed row the 32 lanes of a warp write 32 consecutive floats.
//
// 3xTF32: X = Xh + Xl, W = Wh + Wl with every part exactly representable in TF32 (cvt.rna), and
//     D = Wh.Xh + Wl.Xh + Wh.Xl      (Wl.Xl ~ 2^-22 relative, dropped)
// accumulated in FP32 in TMEM.  Wh/Wl are produced once per call by a small transpose+split kernel; Xh/Xl are
// produced in shared memory by a converter warpgroup from the raw FP32 tile that TMA landed, so activations cross
// L2->SMEM once.
//
// Work decomposition: a work item is (row tile, block of 128 output features); a persistent CTA (one per SM) walks
// items blockIdx.x, blockIdx.x + gridDim.x, ...  The two feature blocks of a 256-wide layer are adjacent items, so
// the second read of the activation tile hits L2.  One 128-feature block per CTA keeps a pipeline stage at 64 KB
// (Xh, Xl, Wh, Wl: 16 KB each -> 3 stages) and leaves TMEM room for two accumulator buffers (the epilogue of item i
// overlaps the MMAs of item i+1), each split in two:
//     main  += Wh.Xh          (one truncating TMEM accumulation per 8-deep K step)
//     cross += Wl.Xh + Wh.Xl  (2^-11 smaller: its truncation error is negligible)
// The tensor core adds into TMEM with round-toward-zero, a coherent bias that grows with the number of accumulation
// steps; keeping the small products out of the large accumulator cuts those steps by three (DESIGN.md "Numerics").
//
// Warp roles (512 threads, 1 CTA/SM):
//   warp 0      TMA producer            warp 1      MMA issuer (+ TMEM alloc/dealloc)
//   warps 4-7   converter (lo part)     warps 8-15  epilogue: lane quarter = warp % 4; the two warps of a quarter
//                                                   take alternate groups of the tile
// Pipelines: smem ring  full[s] (TMA->converter) -> ready[s] (converter->MMA) -> empty[s] (MMA commit->TMA);
//            accumulators acc_full[b] (MMA commit->epilogue) / acc_empty[b] (epilogue->MMA), b = item parity.
#include <cuda.h>

#include <cstdlib>

#include "aug.cuh"

namespace {

constexpr int TC_BK = 32;          // K chunk: 32 floats = 128 B = one SWIZZLE_128B row
constexpr int TC_NMAX = 128;       // max MMA N (rows per tile)
constexpr int TC_MAX_STAGES = 4;   // ring depth is chosen per launch from the tile's row count
constexpr int TC_MBLK = 128;       // features per work item (MMA M)
constexpr int TC_X_BYTES = TC_NMAX * 128;   // 16384 (multiple of 1024)
constexpr int TC_W_BYTES = TC_MBLK * 128;   // 16384
constexpr int TC_SMEM_LIMIT = 227 * 1024;
constexpr int TC_SMEM_EXTRA = 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_THREADS = 512;
constexpr int TC_TMEM_COLS = 512;  // 2 buffers x {main, cross} x 128 columns
constexpr int TC_CH = 8;           // epilogue column chunk

struct TcParams {
  // problem
  int n_rows_tile;   // G_t * C
  int n_mma;         // roundup16(n_rows_tile)
  int G_t, C, N_out, mblocks;
  int kchunks0, kchunks1;
  int stages, x_bytes, stage_bytes;  // smem ring: per stage X (raw fp32 = tf32 hi by truncation) | Xl | Wh | Wl
  long long tiles, tiles_per_w, items;  // items = tiles * mblocks
  int n_sub, n_tot, j0;    // TMA-side grouping (flat launches use n_sub = n_tot = all groups, one "walker")
  int n_tot_true;          // electrons per walker (for the per-walker addend)
  long long G_sub_total;   // total sub-groups covered (W * n_sub)
  const float* bias;
  const float* cadd;
  const float* res;
  float* out;
  int act;       // 0 raw, 1 tanh forward-Laplacian, 2 envelope product
  int res_mode;  // 0 none, 1 (res + y)/sqrt2, 2 res + y
  JqEnvFuse env; // act == 2
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, M=128
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
// K-major, SWIZZLE_128B operand descriptor: 8-row atoms of 1024 B (SBO), rows of 128 B.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // LBO (ignored for swizzled K-major), 16 B
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO = 1024 B between 8-row atoms
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t tc_idesc(int M, int N) {
  // c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, K-major A and B, N>>3 @17, M>>4 @24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __uint_as_float(r);
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// One elected lane of a converged warp (warp-uniform control flow keeps descriptors in uniform registers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n_stages) {
    if (++stage == n_stages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// Epilogue of one CTA (all its items), specialised on the fused operations.  A thread owns one output feature (its
// TMEM lane) and walks the tile's groups; a group's C columns are fetched in chunks of TC_CH from both accumulators
// (tcgen05.ld x8) and summed.  All global addresses are a warp-uniform 64-bit base plus one 32-bit per-thread offset
// (row * N + f) shared by the addend, the residual and the output, so a row costs one integer add.
__device__ __forceinline__ float tmem_sum1(uint32_t tcol) { return tmem_ld1(tcol) + tmem_ld1(tcol + TC_NMAX); }

template <int ACT, int RES, bool CADD>
__device__ __forceinline__ void epilogue_loop(const TcParams& p, uint32_t tlane0, int q, int sub, uint64_t* acc_full,
                                              uint64_t* acc_empty, int lane) {
  const float inv_sqrt2 = 0.70710678118654752440f;
  const int C = p.C;
  const uint32_t N = (uint32_t)p.N_out;
  uint32_t it = 0;
  for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
    const long long t = item / p.mblocks;
    const int hf = (int)(item % p.mblocks);
    const int f = hf * TC_MBLK + q * 32 + lane;
    const bool f_ok = f < p.N_out;
    const uint32_t fo = f_ok ? (uint32_t)f : 0u;  // lanes beyond N_out read a valid column and never store
    const float bias_f = (p.bias && f_ok) ? p.bias[f] : 0.f;
    const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
    const uint32_t tbuf = tlane0 + buf * 2 * TC_NMAX;
    const long long w_tma = t / p.tiles_per_w;
    const int gsub0 = (int)(t % p.tiles_per_w) * p.G_t;
    mbar_wait(&acc_full[buf], acc_phase);
    tc_fence_after();
    for (int gi = sub; gi < p.G_t; gi += 2) {
      const int gsub = gsub0 + gi;
      if (gsub >= p.n_sub) break;  // warp-uniform
      const long long g = w_tma * p.n_tot + p.j0 + gsub;   // actual group index
      const float* __restrict__ cadd_b = CADD ? p.cadd + ((g / p.n_tot_true) * C) * (long long)N : nullptr;
      const float* __restrict__ res_b = RES ? p.res + g * C * (long long)N : nullptr;
      float* __restrict__ out_b = p.out + g * C * (long long)N;
      const uint32_t tcol = tbuf + gi * C;
      // Single pass: the Jacobian rows of tanh only need d1 = 1 - tanh(x)^2 (value column, read first); sum_k y_k^2
      // is accumulated on the way and enters the Laplacian row, which is emitted last.
      float th = 0.f, d1 = 1.f, s2 = 0.f;
      const int r0 = (ACT == 0) ? 0 : 1;
      const int r1 = (ACT == 0) ? C : ((ACT == 1) ? C - 1 : ((C > 1) ? C - 1 : 1));
      // Rows r0..r1 are walked in chunks of TC_CH, software-pipelined: the addend / residual loads of chunk j+1 are
      // in flight while chunk j is read from TMEM, transformed and stored (the epilogue is latency-bound on those
      // loads); chunk 0 is requested before the value column is handled.
      const bool chunked = r1 - r0 >= TC_CH;
      float caA[TC_CH], rrA[TC_CH], caB[TC_CH], rrB[TC_CH];
#define TC_CHUNK_START(j) ((r0 + (j) * TC_CH + TC_CH > r1) ? (r1 - TC_CH) : (r0 + (j) * TC_CH))
#define TC_CHUNK_LOAD(j, ca, rr)                                  \
  {                                                               \
    uint32_t o_ = (uint32_t)TC_CHUNK_START(j) * N + fo;           \
    _Pragma("unroll") for (int i = 0; i < TC_CH; ++i, o_ += N) { \
      if (CADD) ca[i] = cadd_b[o_];                               \
      if (RES) rr[i] = res_b[o_];                                 \
    }                                                             \
  }
      if ((CADD || RES) && chunked) TC_CHUNK_LOAD(0, caA, rrA)
      // ACT == 2: orbital x envelope.  E = sum_I pi exp(-s r_I) for (electron j, orbital i, determinant d) with
      // dE_a = sum_I -s t (r_j - R_I)_a / r_I and lap E = sum_I t (s^2 - 2 s / r_I); product rule per row:
      //   out_0 = y_0 E,  out_c = y_c E (+ y_0 dE_a on the electron's own three rows),
      //   out_L = y_L E + y_0 lap E + 2 sum_a y_{own a} dE_a
      float y0 = 0.f, ev = 1.f, e_d[3] = {0.f, 0.f, 0.f}, e_l = 0.f;
      int own0 = -8;
      if (ACT == 2) {
        const int nel = p.env.n;
        const int j = (int)(g % nel);
        own0 = 1 + 3 * j;
        const int dd = (int)fo / nel, io = (int)fo % nel;
        const float* e = p.env.electrons + g * 3;
        ev = 0.f;
        for (int I = 0; I < p.env.A; ++I) {
          const float dx = e[0] - p.env.atoms[3 * I], dy = e[1] - p.env.atoms[3 * I + 1], dz = e[2] - p.env.atoms[3 * I + 2];
          const float r = sqrtf(dx * dx + dy * dy + dz * dz);
          float sg = p.env.sigma[(io * p.env.A + I) * p.env.D + dd];
          if (p.env.type == 1) sg = fabsf(sg);
          const float t = p.env.pi[(io * p.env.A + I) * p.env.D + dd] * expf(-sg * r);
          const float rinv = 1.0f / r;
          const float c1 = -sg * t * rinv;
          ev += t;
          e_d[0] += c1 * dx;
          e_d[1] += c1 * dy;
          e_d[2] += c1 * dz;
          e_l += t * (sg * sg - 2.0f * sg * rinv);
        }
        y0 = tmem_sum1(tcol) + bias_f;
        if (f_ok) out_b[fo] = y0 * ev;
      }
      if (ACT == 1) {
        float x = tmem_sum1(tcol);
        if (CADD) x += cadd_b[fo];
        x += bias_f;
        th = tanhf(x);
        d1 = 1.0f - th * th;
      }
      if (chunked) {
        const int nch = (r1 - r0 + TC_CH - 1) / TC_CH;
        auto process = [&](int j, const float* ca, const float* rr) {
          const int cs = TC_CHUNK_START(j);
          const int skip = r0 + j * TC_CH - cs;  // last chunk: shifted back to stay inside the group's columns
          float v[TC_CH], v2[TC_CH];
          tmem_ld8_nowait(tcol + cs, v);
          tmem_ld8_nowait(tcol + TC_NMAX + cs, v2);
          tmem_wait_ld();
          uint32_t o = (uint32_t)cs * N + fo;
#pragma unroll
          for (int i = 0; i < TC_CH; ++i, o += N) {
            float y = v[i] + v2[i];
            if (CADD) y += ca[i];
            if (ACT == 1) {
              if (i >= skip) s2 = fmaf(y, y, s2);
              y *= d1;
            }
            if (ACT == 2) {
              const int a = cs + i - own0;
              float o2 = y * ev;
              if (a >= 0 && a < 3) {
                const float da = (a == 0) ? e_d[0] : ((a == 1) ? e_d[1] : e_d[2]);
                if (i >= skip) s2 = fmaf(y, da, s2);  // cross term for the Laplacian row
                o2 = fmaf(y0, da, o2);
              }
              y = o2;
            }
            if (ACT == 0 && cs + i == 0) y += bias_f;
            if (RES == 1) y = (rr[i] + y) * inv_sqrt2;
            if (RES == 2) y = rr[i] + y;
            if (f_ok && i >= skip) out_b[o] = y;
          }
        };
        for (int j = 0; j < nch; j += 2) {
          if ((CADD || RES) && j + 1 < nch) TC_CHUNK_LOAD(j + 1, caB, rrB)
          process(j, caA, rrA);
          if (j + 1 < nch) {
            if ((CADD || RES) && j + 2 < nch) TC_CHUNK_LOAD(j + 2, caA, rrA)
            process(j + 1, caB, rrB);
          }
        }
#undef TC_CHUNK_LOAD
#undef TC_CHUNK_START
      } else {
        for (int c = r0; c < r1; ++c) {
          float y = tmem_sum1(tcol + c);
          const uint32_t o = (uint32_t)c * N + fo;
          if (CADD) y += cadd_b[o];
          if (ACT == 1) {
            s2 = fmaf(y, y, s2);
            y *= d1;
          }
          if (ACT == 2) {
            const int a = c - own0;
            float o2 = y * ev;
            if (a >= 0 && a < 3) {
              const float da = (a == 0) ? e_d[0] : ((a == 1) ? e_d[1] : e_d[2]);
              s2 = fmaf(y, da, s2);
              o2 = fmaf(y0, da, o2);
            }
            y = o2;
          }
          if (ACT == 0 && c == 0) y += bias_f;
          if (RES == 1) y = (res_b[o] + y) * inv_sqrt2;
          if (RES == 2) y = res_b[o] + y;
          if (f_ok) out_b[o] = y;
        }
      }
      if (ACT == 2 && C > 1) {
        const float yl = tmem_sum1(tcol + C - 1);
        if (f_ok) out_b[(uint32_t)(C - 1) * N + fo] = yl * ev + y0 * e_l + 2.0f * s2;
      }
      if (ACT == 1) {
        float yl = (C > 1) ? tmem_sum1(tcol + C - 1) : 0.f;
        if (f_ok) {
          if (C > 1) {
            const uint32_t o = (uint32_t)(C - 1) * N + fo;
            if (CADD) yl += cadd_b[o];
            float l = d1 * yl - 2.0f * th * d1 * s2;
            if (RES == 1) l = (res_b[o] + l) * inv_sqrt2;
            if (RES == 2) l = res_b[o] + l;
            out_b[o] = l;
          }
          float o0 = th;
          if (RES == 1) o0 = (res_b[fo] + th) * inv_sqrt2;
          if (RES == 2) o0 = res_b[fo] + th;
          out_b[fo] = o0;
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[buf]);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_dense_tc(const __grid_constant__ CUtensorMap mapX0, const __grid_constant__ CUtensorMap mapX1,
           const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl, TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* full = bars;                       // [stages]  TMA -> converter and MMA (raw X, W)
  uint64_t* ready = bars + TC_MAX_STAGES;      // [stages]  converter -> MMA (Xl)
  uint64_t* empty = bars + 2 * TC_MAX_STAGES;  // [stages]  MMA -> TMA
  uint64_t* acc_full = bars + 3 * TC_MAX_STAGES;   // [2]  MMA commit -> epilogue, per accumulator buffer
  uint64_t* acc_empty = acc_full + 2;          // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kchunks = p.kchunks0 + p.kchunks1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], 4);   // one arrive per converter warp
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      PipeState ps;
      const uint32_t stage_tx = (uint32_t)(p.n_mma * 128 + 2 * TC_W_BYTES);
      for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const long long t = item / p.mblocks;
        const int hf = (int)(item % p.mblocks);
        long long w = t / p.tiles_per_w;
        int gsub0 = (int)(t % p.tiles_per_w) * p.G_t;
        int row0 = (p.j0 + gsub0) * p.C;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[ps.stage], ps.phase ^ 1);
          unsigned char* st = smem + ps.stage * p.stage_bytes;
          mbar_expect_tx(&full[ps.stage], stage_tx);
          if (kc < p.kchunks0)
            tma_load_3d(st, &mapX0, &full[ps.stage], kc * TC_BK, row0, (int)w);
          else
            tma_load_3d(st, &mapX1, &full[ps.stage], (kc - p.kchunks0) * TC_BK, row0, (int)w);
          tma_load_2d(st + 2 * p.x_bytes, &mapWh, &full[ps.stage], kc * TC_BK, hf * TC_MBLK);
          tma_load_2d(st + 2 * p.x_bytes + TC_W_BYTES, &mapWl, &full[ps.stage], kc * TC_BK, hf * TC_MBLK);
          ps.advance(p.stages);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop; one elected lane issues) =====================
    {
      PipeState ps;
      uint32_t it = 0;
      const uint32_t idesc = tc_idesc(TC_MBLK, p.n_mma);
      // Operand descriptors differ between stages and K steps only in the 14-bit start-address field (bytes >> 4):
      // build them once and advance with integer adds -- the issuing thread is the pipeline's critical resource.
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t dx_h0 = tc_smem_desc(smem0);
      const uint64_t dx_l0 = tc_smem_desc(smem0 + p.x_bytes);
      const uint64_t dw_h0 = tc_smem_desc(smem0 + 2 * p.x_bytes);
      const uint64_t dw_l0 = tc_smem_desc(smem0 + 2 * p.x_bytes + TC_W_BYTES);
      const uint64_t stage_step = (uint64_t)(p.stage_bytes >> 4);
      for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&acc_empty[buf], acc_phase ^ 1);   // epilogue has drained this buffer (two items ago)
        tc_fence_after();
        const uint32_t d_main = tmem_base + buf * 2 * TC_NMAX, d_cross = d_main + TC_NMAX;
        for (int kc = 0; kc < kchunks; ++kc) {
          // the raw FP32 tile is consumed as TF32 directly (the tensor core ignores the low 13 mantissa bits: hi =
          // trunc(x)), so two of the three products start as soon as TMA lands; only Wh.Xl waits for the converter
          const uint64_t so = stage_step * (uint64_t)ps.stage;
          const uint64_t xh = dx_h0 + so, xl = dx_l0 + so, wh = dw_h0 + so, wl = dw_l0 + so;
          const uint32_t acc0 = kc ? 1u : 0u;
          mbar_wait(&full[ps.stage], ps.phase);
          tc_fence_after();
          // 8 tf32 = 32 bytes along K inside the 128 B swizzle row: +2 in the address field per K step
          if (elect_one()) {
          tc_mma_tf32(d_main, wh, xh, idesc, acc0);
          tc_mma_tf32(d_cross, wl, xh, idesc, acc0);
          tc_mma_tf32(d_main, wh + 2, xh + 2, idesc, 1u);
          tc_mma_tf32(d_cross, wl + 2, xh + 2, idesc, 1u);
          tc_mma_tf32(d_main, wh + 4, xh + 4, idesc, 1u);
          tc_mma_tf32(d_cross, wl + 4, xh + 4, idesc, 1u);
          tc_mma_tf32(d_main, wh + 6, xh + 6, idesc, 1u);
          tc_mma_tf32(d_cross, wl + 6, xh + 6, idesc, 1u);
          }
          __syncwarp();
          mbar_wait(&ready[ps.stage], ps.phase);
          tc_fence_after();
          if (elect_one()) {
          tc_mma_tf32(d_cross, wh, xl, idesc, 1u);
          tc_mma_tf32(d_cross, wh + 2, xl + 2, idesc, 1u);
          tc_mma_tf32(d_cross, wh + 4, xl + 4, idesc, 1u);
          tc_mma_tf32(d_cross, wh + 6, xl + 6, idesc, 1u);
          tc_commit(&empty[ps.stage]);   // stage reusable once these MMAs have read it
          if (kc == kchunks - 1) tc_commit(&acc_full[buf]);
          }
          __syncwarp();
          ps.advance(p.stages);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== converter: Xl = x - trunc_tf32(x) =====================
    PipeState ps;
    const int ct = threadIdx.x - 128;  // 0..127
    const int nvec = p.n_mma * 8;      // float4 per X chunk (128 B rows)
    for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&full[ps.stage], ps.phase);
        const float4* xh = reinterpret_cast<const float4*>(smem + ps.stage * p.stage_bytes);
        float4* xl = reinterpret_cast<float4*>(smem + ps.stage * p.stage_bytes + p.x_bytes);
        for (int i = ct; i < nvec; i += 128) {
          float4 v = xh[i];
          float4 l;  // lo = x - trunc_tf32(x): exact, at most 13 significant bits
          l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          xl[i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[ps.stage]);
        ps.advance(p.stages);
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int sub = (warp - 8) >> 2;   // which of the quarter's two warps
    const uint32_t tlane0 = tmem_base + ((uint32_t)(q * 32) << 16);
    // one instantiation per (activation, residual mode, addend) combination: no per-element predicates
#define TC_EPI(ACT, RES, CADD) epilogue_loop<ACT, RES, CADD>(p, tlane0, q, sub, acc_full, acc_empty, lane)
    const int key = (p.act == 2) ? 16 : ((p.act ? 8 : 0) | (p.res_mode << 1) | (p.cadd ? 1 : 0));
    switch (key) {
      case 16: TC_EPI(2, 0, false); break;
      case 0: TC_EPI(0, 0, false); break;
      case 1: TC_EPI(0, 0, true); break;
      case 2: TC_EPI(0, 1, false); break;
      case 3: TC_EPI(0, 1, true); break;
      case 4: TC_EPI(0, 2, false); break;
      case 5: TC_EPI(0, 2, true); break;
      case 8: TC_EPI(1, 0, false); break;
      case 9: TC_EPI(1, 0, true); break;
      case 10: TC_EPI(1, 1, false); break;
      case 11: TC_EPI(1, 1, true); break;
      case 12: TC_EPI(1, 2, false); break;
      default: TC_EPI(1, 2, true); break;
    }
#undef TC_EPI
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// Wt_hi / Wt_lo [N_out][Kt] (K-major) from the flax kernel rows w0 [k0][N], w1 [k1][N].
__global__ void k_weight_split_t(const float* __restrict__ w0, int k0, const float* __restrict__ w1, int k1, int N,
                                 int ldw, float* __restrict__ wh, float* __restrict__ wl) {
  const int kt = k0 + k1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * kt;
       i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % kt);
    int f = (int)(i / kt);
    float v = (k < k0) ? w0[(long long)k * ldw + f] : w1[(long long)(k - k0) * ldw + f];
    float h = tf32_rna(v);
    wh[i] = h;
    wl[i] = tf32_rna(v - h);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// dims are innermost-first; strides (bytes) for dims 1..rank-1
int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
             const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  JQ_REQUIRE(enc != nullptr, JQ_ERR_CUDA, "dense_tc: cuTensorMapEncodeTiled is unavailable");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  JQ_REQUIRE(r == CUDA_SUCCESS, JQ_ERR_CUDA, "dense_tc: cuTensorMapEncodeTiled failed with %d", (int)r);
  return JQ_OK;
}

int pick_groups_per_tile(int C) {
  int g = TC_NMAX / C;
  return g < 1 ? 0 : g;
}

}  // namespace

size_t jq_dense_tc_scratch_floats(int k_total, int n_out) { return (size_t)2 * k_total * n_out; }

bool jq_dense_tc_eligible(const JqDenseArgs& a) {
  // debugging switch (bisecting a numerical difference between the two dense kernels); both are sm_100a CUDA
  static const bool disabled = getenv("JAQMC_B200_DISABLE_TC") != nullptr;
  if (disabled) return false;
  if (a.C > TC_NMAX) return false;
  if (a.k0 % TC_BK || a.k1 % TC_BK) return false;
  if (a.k0 + a.k1 < 32) return false;
  if (a.N < 64 || a.N > 2 * TC_MBLK) return false;
  if (!a.wscratch) return false;
  if ((reinterpret_cast<uintptr_t>(a.src0) & 15) || (a.src1 && (reinterpret_cast<uintptr_t>(a.src1) & 15))) return false;
  return true;
}

int jq_launch_dense_tc(const JqDenseArgs& a, cudaStream_t st, bool* handled) {
  *handled = false;
  if (!jq_dense_tc_eligible(a)) return JQ_OK;
  static int sm_count = 0;
  static bool attr_set = false;
  if (!attr_set) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(k_dense_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    JQ_REQUIRE(e == cudaSuccess, JQ_ERR_CUDA, "dense_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int kt = a.k0 + a.k1;
  float* wh = a.wscratch;
  float* wl = a.wscratch + (size_t)kt * a.N;
  {
    long long items = (long long)kt * a.N;
    int grid = jq_cdiv(items, 256);
    if (grid > 148 * 4) grid = 148 * 4;
    JQ_LAUNCH(k_weight_split_t, dim3(grid), dim3(256), 0, st, a.w0, a.k0, a.w1, a.k1, a.N, a.ldw ? a.ldw : a.N, wh, wl);
    JQ_CHECK_LAUNCH();
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.C = a.C;
  p.G_t = pick_groups_per_tile(a.C);
  p.N_out = a.N;
  p.mblocks = jq_cdiv(a.N, TC_MBLK);
  p.kchunks0 = a.k0 / TC_BK;
  p.kchunks1 = a.k1 / TC_BK;
  p.n_tot_true = a.n_tot;
  long long Wn;  // number of TMA "walkers" (outer dimension)
  if (a.n_sub == a.n_tot) {
    // flat: all groups are contiguous
    p.n_sub = p.n_tot = 0;  // set below (may exceed int for huge problems: guarded)
    JQ_REQUIRE(a.G <= 0x7fffffffLL / a.C, JQ_ERR_UNSUPPORTED, "dense_tc: too many rows");
    p.n_sub = p.n_tot = (int)a.G;
    p.j0 = 0;
    Wn = 1;
  } else {
    p.n_sub = a.n_sub;
    p.n_tot = a.n_tot;
    p.j0 = a.j0;
    Wn = a.G / a.n_sub;
  }
  if (p.G_t > p.n_sub) p.G_t = p.n_sub;
  p.n_rows_tile = p.G_t * a.C;
  p.n_mma = (p.n_rows_tile + 15) / 16 * 16;
  p.tiles_per_w = jq_cdiv(p.n_sub, p.G_t);
  p.tiles = p.tiles_per_w * Wn;
  p.items = p.tiles * p.mblocks;
  p.x_bytes = p.n_mma * 128;  // multiple of 2048: keeps every operand 1024-byte aligned
  p.stage_bytes = 2 * p.x_bytes + 2 * TC_W_BYTES;
  p.stages = (TC_SMEM_LIMIT - TC_SMEM_EXTRA) / p.stage_bytes;
  if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
  const int smem_bytes = p.stages * p.stage_bytes + TC_SMEM_EXTRA;
  p.G_sub_total = a.G;
  p.bias = a.bias;
  p.cadd = a.cadd;
  p.res = a.res;
  p.out = a.out;
  p.act = a.act;
  p.res_mode = a.res_mode;
  if (a.act == 2) {
    JQ_REQUIRE(a.env && !a.cadd && !a.res && a.env->n == a.n_tot && a.N == a.env->D * a.env->n, JQ_ERR_INVALID_ARGUMENT,
               "dense_tc: envelope epilogue needs (determinant, orbital) features and no addend / residual");
    p.env = *a.env;
  }

  CUtensorMap mX0, mX1, mWh, mWl;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.k0, (cuuint64_t)p.n_tot * a.C, (cuuint64_t)Wn};
    cuuint64_t str[2] = {(cuuint64_t)a.k0 * 4, (cuuint64_t)p.n_tot * a.C * a.k0 * 4};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)p.n_mma, 1};
    int rc = make_map(&mX0, a.src0, 3, dims, str, box);
    if (rc) return rc;
    if (a.k1 > 0) {
      cuuint64_t dims1[3] = {(cuuint64_t)a.k1, (cuuint64_t)p.n_tot * a.C, (cuuint64_t)Wn};
      cuuint64_t str1[2] = {(cuuint64_t)a.k1 * 4, (cuuint64_t)p.n_tot * a.C * a.k1 * 4};
      rc = make_map(&mX1, a.src1, 3, dims1, str1, box);
      if (rc) return rc;
    } else {
      mX1 = mX0;
    }
    cuuint64_t wd[2] = {(cuuint64_t)kt, (cuuint64_t)a.N};
    cuuint64_t ws[1] = {(cuuint64_t)kt * 4};
    cuuint32_t wbox[2] = {TC_BK, TC_MBLK};
    rc = make_map(&mWh, wh, 2, wd, ws, wbox);
    if (rc) return rc;
    rc = make_map(&mWl, wl, 2, wd, ws, wbox);
    if (rc) return rc;
  }
  long long grid = p.items < sm_count ? p.items : sm_count;
  double R = (double)a.G * a.C;
  jq_prof_work(2.0 * R * kt * a.N, 4.0 * R * (kt + a.N * (a.res ? 2 : 1)));
  JQ_LAUNCH(k_dense_tc, dim3((unsigned)grid), dim3(TC_THREADS), smem_bytes, st, mX0, mX1, mWh, mWl, p);
  JQ_CHECK_LAUNCH();
  *handled = true;
  return JQ_OK;
}
