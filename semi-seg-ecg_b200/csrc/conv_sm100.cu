// Blackwell-native Conv1d for the wide-channel stages: implicit GEMM on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA into 128B-swizzled
// shared memory, mbarrier producer/consumer pipeline, warp-specialised roles.
//
//   fprop / dgrad  (conv_tn_kernel):  D[rows, N] = sum_taps A_tap[rows, K] * B_tap[K, N]
//       A (activations) K-major.  The weights live in ONE layout, [k][Cin][Cout]: for dgrad
//       (N = Cin, K = Cout) that is a K-major B operand, for fprop (N = Cout, K = Cin) an
//       MN-major one (template B_MN) -- no transposed weight copy exists.
//       A k=3 conv is three row-shifted TMA boxes over the flat padded
//       NLC activation (halo rows / TMA out-of-bounds zero fill implement the padding); a
//       stride-2 conv reads the input through a [rows/2, 2C] "row pair" view.
//   wgrad          (conv_wgrad_kernel): dW_tap[ci, co] = sum_rows X_tap[rows, ci] * dY[rows, co]
//       both operands MN-major (the reduction runs over rows), split over rows across CTAs.
//
// Replaces cuDNN fprop/dgrad/wgrad for resnet.py:32-49,283-289 and fcn_head.py:40-47 (K4, K6 in
// SURVEY.md 2.2).  bf16 operands, fp32 accumulation.  One CTA = one 128 x BN output tile;
// warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;          // UMMA_M
constexpr int BK = 64;           // bf16 elements per 128-byte swizzle row
constexpr int A_BYTES = BM * BK * 2;
constexpr int NTHREADS = 192;      // wgrad kernels: TMA warp, MMA warp, 4 epilogue warps
constexpr int TN_THREADS = 320;    // fprop / dgrad kernels: TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant,
                                   // each owning one half of the tile's columns)
constexpr int EPI_THREADS = 256;

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef SSB_BOUNDED_WAIT
  // development builds (make NVFLAGS+=-DSSB_BOUNDED_WAIT): a pipeline-protocol error becomes a launch failure after
  // ~2 s instead of a hung device.  Not in the product build: the clock reads in the single-thread producer / MMA
  // issue loops cost 1.7 % of the conv kernels at large batch and 0.3 % of the config-2 step (measured).
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) asm volatile("trap;");
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the mbarrier once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (lane i of the warp <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, 128-byte swizzle, Blackwell version bit
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// NOTE (measured on B200, tests/test_kernels_gpu.py): a start address that is offset by whole 128-byte
// rows inside an 8-row swizzle group needs NO matrix-base-offset field -- the 128B swizzle is applied on
// absolute shared-memory address bits, exactly as TMA wrote the tile; setting (addr >> 7) & 7 in bits
// 49-51 gives wrong products.  The tap-reuse kernel below relies on this.
// instruction descriptor: bf16 x bf16 -> f32, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major ? (1u << 15) : 0u) | (b_mn_major ? (1u << 16) : 0u) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct TnParams {
  int M, N, K, ntaps;
  int a_row_off[3], a_col_off[3], w_tap[3];
  int w_rows_per_tap;
  int o_mul, o_off, o_rows, o_pitch, o_len;
  int accumulate;
  double* stats;   // STATS epilogue: [2N] per-channel sum / sum of squares of the stored (rounded) output
  // EPI epilogue (eval-mode BatchNorm folded into the conv): y = [relu](acc*scale + shift [+ res])
  const float* ep_gamma;
  const float* ep_beta;
  const float* ep_mean;
  const float* ep_var;
  const bf16* ep_res;   // residual in the output geometry, may be NULL
  int ep_relu;
  // DUAL (STATS and EPI together): rows < split_row are TRAIN rows (raw output to `out`, statistics), rows >=
  // split_row are EVAL rows of the same conv (EPI transform, written to out2 with the same row indexing) --
  // FixMatch's pseudo-label forward rides in the student's launches (same weights, more rows)
  int split_row;
  bf16* out2;
  // RED epilogue (dgrad): BatchNorm-backward reduce of the gradient this launch produces, g = dx * (red_y > 0):
  // red_sums[0:N] += sum g, red_sums[N:2N] += sum g * xhat(red_x); RED == 2 also the residual-branch BN (red_xr)
  const bf16* red_y;
  const bf16* red_x;
  const bf16* red_xr;
  const float* red_mi;      // [2N] mean / invstd saved by the forward
  const float* red_mi_r;
  double* red_sums;
  double* red_sums_r;
  // BNF (with STATS, train mode): after the statistics are complete ACROSS the launch (grid-wide barrier on
  // bnf_barrier), a second pass over the TMEM accumulator writes out2 = [relu](bn(y) [+ res | + bn_r(res)]) --
  // the BatchNorm apply pass without its own launch and without re-reading y.  bnf / bnf_r: the BN layers
  // (batch statistics from their sums; the block row 0 tiles update running stats and save mean / invstd).
  unsigned int* bnf_barrier;
  unsigned int bnf_expected;
  ssb_bn bnf, bnf_r;
  const bf16* bnf_res;
  int bnf_res_mode;   // 0 none, 1 identity residual, 2 residual with its own BN
  int bnf_relu;
  double bnf_n;       // elements per channel (B * len * world)
};

// Column sums across the 32 lanes of a warp: every lane holds 32 column values x[0..31] of its own row;
// on return x[0] of lane l is the sum over all lanes of column l (31 shuffles instead of 160).
__device__ __forceinline__ void warp_transpose_sum(float (&x)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = upper ? x[i] : x[i + o];
      const float keep = upper ? x[i + o] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

struct SmemLayout {
  uint8_t* a;
  uint8_t* b;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* done;
  uint32_t* tmem_slot;
};

template <int B_BYTES, int STAGES>
__device__ __forceinline__ SmemLayout carve(uint8_t* raw) {
  const uint32_t base = smem_u32(raw);
  uint8_t* al = raw + (((base + 1023u) & ~1023u) - base);
  SmemLayout s;
  s.a = al;
  s.b = al + STAGES * A_BYTES;
  s.full = reinterpret_cast<uint64_t*>(s.b + STAGES * B_BYTES);
  s.empty = s.full + STAGES;
  s.done = s.empty + STAGES;
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.done + 1);
  return s;
}
template <int B_BYTES, int STAGES>
constexpr int smem_bytes() {
  return STAGES * (A_BYTES + B_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024 + 4 * 128 * 4;   // + per-column coefficients (EPI / RED)
}

// ---------------------------------------------------------------------------------------------
// epilogue shared by the conv kernels: TMEM -> registers -> (BN affine / residual / ReLU | accumulate)
// -> bf16 rows in global memory, plus the optional per-channel statistics of the stored values.
// Called by the four epilogue warps (threads 64..191); `scratch` is operand smem that is free once
// `done` has fired, `ep_scale` a dedicated [2*BN] float area.
// ---------------------------------------------------------------------------------------------
template <int BN, bool STATS, bool EPI, int RED = 0>
__device__ __forceinline__ void tn_epilogue(const TnParams& p, bf16* __restrict__ out, uint32_t tmem_base, uint64_t* done,
                                          uint32_t done_parity, uint64_t* release, float* ep_scale, float* scratch, int m0,
                                          int n0, int warp, int lane) {
  // RED template parameter: 0 none, 1 / 2 = BN-backward reduce of the produced gradient (2: + residual BN), 3 = BNF, the
  // train-mode BatchNorm apply pass fused behind the statistics (compiled only into the variants that use it, so that
  // the default kernels stay small)
  constexpr bool REDC = RED == 1 || RED == 2;
  constexpr bool BNFC = RED == 3;
  // epilogue warps 2..9: warp w may only touch TMEM lanes 32*(w%4) .. +31; the two warps of a lane quadrant split the
  // tile's columns in halves
  const int q = warp & 3;
  const int wq = warp - 2;                 // 0..7
  const int half = wq >> 2;
  constexpr int CH = BN / 2;               // columns per half (a multiple of 32)
  const int cbeg = half * CH, cend = cbeg + CH;
  const int e = wq * 32 + lane;            // 0..255
  float* ep_shift = ep_scale + BN;
  if (EPI) {   // eval-mode BN coefficients of this CTA's columns, computed while the main loop runs
    for (int col = e; col < BN; col += EPI_THREADS) {
      const float sc = p.ep_gamma[n0 + col] * (1.0f / sqrtf(p.ep_var[n0 + col] + 1e-5f));
      ep_scale[col] = sc;
      ep_shift[col] = p.ep_beta[n0 + col] - p.ep_mean[n0 + col] * sc;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  if (REDC) {   // saved mean / invstd of this CTA's columns (ep_scale area: mean, inv, [mean_r, inv_r])
    for (int col = e; col < BN; col += EPI_THREADS) {
      ep_scale[col] = p.red_mi[n0 + col];
      ep_scale[BN + col] = p.red_mi[p.N + n0 + col];
      if (RED == 2) {
        ep_scale[2 * BN + col] = p.red_mi_r[n0 + col];
        ep_scale[3 * BN + col] = p.red_mi_r[p.N + n0 + col];
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  const int row = q * 32 + lane;
  const int m = m0 + row;
  const long long orow = (long long)p.o_mul * m + p.o_off;
  const bool in_range = m < p.M && orow >= 0 && orow < p.o_rows;
  const bool valid = in_range && row_valid((int)orow, p.o_pitch, p.o_len);
  constexpr bool DUAL = STATS && EPI;
  const bool eval_row = !DUAL || m >= p.split_row;    // this thread's row takes the EPI transform
  bf16* optr = ((DUAL && eval_row) ? p.out2 : out) + (size_t)(in_range ? orow : 0) * p.N + n0;
  // RED: the operands of the BatchNorm-backward reduce (y for the ReLU mask, x for xhat[, x_res]) are requested NOW, while
  // the main loop still runs -- as first written they were loaded chunk by chunk after the accumulator, two dependent L2
  // round trips per 32 columns on the critical tail of the kernel (measured slower than a separate reduce launch)
  constexpr int NCHK = CH / 32;                          // 32-column chunks of this warp's column half
  constexpr int PF = REDC ? (NCHK <= 2 ? NCHK : 2) : 0;  // chunks whose operands are held in registers
  uint4 pf_y[PF > 0 ? PF : 1][4], pf_x[PF > 0 ? PF : 1][4], pf_r[(PF > 0 && RED == 2) ? PF : 1][4];
  if (REDC && valid) {
#pragma unroll
    for (int ch = 0; ch < PF; ++ch) {
      const size_t roff = (size_t)orow * p.N + n0 + cbeg + ch * 32;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        pf_y[ch][v] = *reinterpret_cast<const uint4*>(p.red_y + roff + v * 8);
        pf_x[ch][v] = *reinterpret_cast<const uint4*>(p.red_x + roff + v * 8);
        if (RED == 2) pf_r[ch][v] = *reinterpret_cast<const uint4*>(p.red_xr + roff + v * 8);
      }
    }
  }
  mbar_wait(done, done_parity);
  if (threadIdx.x == 64) SSB_MARK();   // accumulator complete (MMAs retired)
  tc_fence_after();
  // all MMAs have retired: the operand stages are free, stage 0 of A is reused as reduction scratch
  float* red = scratch;   // [4 warps][2][BN]
  // (RED: fully unrolled, so that the prefetched operand arrays are indexed by constants and stay in registers)
  constexpr int UNR = REDC ? NCHK : 1;
#pragma unroll UNR
  for (int ci = 0; ci < NCHK; ++ci) {
    const int c = cbeg + ci * 32;
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
    tmem_ld_wait();
    if (REDC && threadIdx.x == 64) SSB_MARK();   // chunk in registers
    float sv[32];
    if (in_range && !(!valid && p.accumulate)) {
      uint4* dst = reinterpret_cast<uint4*>(optr + c);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = valid ? __uint_as_float(r[v * 8 + i]) : 0.f;
        if (EPI && eval_row) {
          if (valid) {
            float rs[8];
            if (p.ep_res) {
              Vec<bf16> rv;
              rv.raw = *reinterpret_cast<const uint4*>(p.ep_res + (size_t)orow * p.N + n0 + c + v * 8);
              rv.get(rs);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float o = fmaf(f[i], ep_scale[c + v * 8 + i], ep_shift[c + v * 8 + i]);
              if (p.ep_res) o += rs[i];
              f[i] = p.ep_relu ? fmaxf(o, 0.f) : o;
            }
          }
        } else if (!EPI && p.accumulate) {
          Vec<bf16> prev;
          prev.raw = dst[v];
          float g[8];
          prev.get(g);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += g[i];
        }
        Vec<bf16> o;
        o.set(f);
        dst[v] = o.raw;
        if (REDC) o.get(&sv[v * 8]);   // the gradient as stored
        if (STATS) {   // statistics of the values as stored (train rows only)
          if (DUAL && eval_row) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sv[v * 8 + i] = 0.f;
          } else {
            o.get(&sv[v * 8]);
          }
        }
      }
    } else if (STATS || REDC) {
#pragma unroll
      for (int i = 0; i < 32; ++i) sv[i] = 0.f;
    }
    if (REDC && threadIdx.x == 64) SSB_MARK();     // chunk stored
    if (REDC) {
      // g = dx * (y > 0); per-column sums of g and g * xhat over this tile's rows, 32 columns at a time: warp
      // transpose-reduce into the scratch (one warp per (lane quadrant, column)); combined and added after the loop
      const size_t roff = (size_t)(in_range ? orow : 0) * p.N + n0 + c;
      constexpr int NQR = RED == 2 ? 3 : 2;
      float t[32];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float fy[8];
        if (valid) {
          Vec<bf16> vy;
          vy.raw = ci < PF ? pf_y[ci < PF ? ci : 0][v] : *reinterpret_cast<const uint4*>(p.red_y + roff + v * 8);
          vy.get(fy);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sv[v * 8 + i] = (valid && fy[i] > 0.f) ? sv[v * 8 + i] : 0.f;
          t[v * 8 + i] = sv[v * 8 + i];
        }
      }
      warp_transpose_sum(t, lane);
      red[(q * NQR + 0) * BN + c + lane] = t[0];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float fx[8];
        if (valid) {
          Vec<bf16> vx;
          vx.raw = ci < PF ? pf_x[ci < PF ? ci : 0][v] : *reinterpret_cast<const uint4*>(p.red_x + roff + v * 8);
          vx.get(fx);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          t[v * 8 + i] = valid ? sv[v * 8 + i] * ((fx[i] - ep_scale[c + v * 8 + i]) * ep_scale[BN + c + v * 8 + i]) : 0.f;
      }
      warp_transpose_sum(t, lane);
      red[(q * NQR + 1) * BN + c + lane] = t[0];
      if (RED == 2) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float fx[8];
          if (valid) {
            Vec<bf16> vx;
            vx.raw = ci < PF ? pf_r[(ci < PF && RED == 2) ? ci : 0][v] : *reinterpret_cast<const uint4*>(p.red_xr + roff + v * 8);
            vx.get(fx);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            t[v * 8 + i] = valid ? sv[v * 8 + i] * ((fx[i] - ep_scale[2 * BN + c + v * 8 + i]) * ep_scale[3 * BN + c + v * 8 + i]) : 0.f;
        }
        warp_transpose_sum(t, lane);
        red[(q * NQR + 2) * BN + c + lane] = t[0];
      }
      if (threadIdx.x == 64) SSB_MARK();           // chunk reduced
    }
    if (STATS) {
      float sq[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) sq[i] = sv[i] * sv[i];
      warp_transpose_sum(sv, lane);
      warp_transpose_sum(sq, lane);
      red[(q * 2 + 0) * BN + c + lane] = sv[0];
      red[(q * 2 + 1) * BN + c + lane] = sq[0];
    }
  }
  const bool bnf = STATS && !EPI && BNFC && p.bnf_barrier != nullptr;
  if (release && !bnf) {   // persistent kernel: this thread is done reading the accumulator buffer
    tc_fence_before();
    mbar_arrive(release);
  }
  if (threadIdx.x == 64) SSB_MARK();     // tile stored
  if (REDC) {
    constexpr int NQR = RED == 2 ? 3 : 2;
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps only
    for (int col = e; col < BN; col += EPI_THREADS) {
      float a[NQR];
#pragma unroll
      for (int wh = 0; wh < NQR; ++wh) a[wh] = (red[(0 * NQR + wh) * BN + col] + red[(1 * NQR + wh) * BN + col]) +
                                               (red[(2 * NQR + wh) * BN + col] + red[(3 * NQR + wh) * BN + col]);
      const int gc = n0 + col;
      atomicAdd(&p.red_sums[gc], (double)a[0]);
      atomicAdd(&p.red_sums[p.N + gc], (double)a[1]);
      if (RED == 2) {
        atomicAdd(&p.red_sums_r[gc], (double)a[0]);
        atomicAdd(&p.red_sums_r[p.N + gc], (double)a[2]);
      }
    }
  }
  if (STATS) {
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps only
    for (int col = e; col < BN; col += EPI_THREADS) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        a += red[(w * 2 + 0) * BN + col];
        b += red[(w * 2 + 1) * BN + col];
      }
      atomicAdd(&p.stats[n0 + col], (double)a);
      atomicAdd(&p.stats[p.N + n0 + col], (double)b);
    }
    if (bnf) {
      // ---- train-mode BatchNorm apply fused in: wait until every tile of this launch has added its statistics ----
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (e == 0) {
        atomicAdd(p.bnf_barrier, 1u);
        unsigned int v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.bnf_barrier) : "memory");
        } while (v < p.bnf_expected);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // per-column coefficients (one thread per column does the fp64 part); the first row tile of each column
      // tile updates the running statistics and saves mean / invstd for the backward
      const bool writer = m0 == 0;
      for (int col = e; col < BN; col += EPI_THREADS) {
        float sc, sh;
        bn_coeffs(p.bnf, n0 + col, p.N, 1, 1.0 / p.bnf_n, p.bnf_n, writer, sc, sh);
        ep_scale[col] = sc;
        ep_scale[BN + col] = sh;
        if (p.bnf_res_mode == 2) {
          bn_coeffs(p.bnf_r, n0 + col, p.N, 1, 1.0 / p.bnf_n, p.bnf_n, writer, sc, sh);
          ep_scale[2 * BN + col] = sc;
          ep_scale[3 * BN + col] = sh;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bf16* optr2 = p.out2 + (size_t)(in_range ? orow : 0) * p.N + n0;
#pragma unroll 1
      for (int c = cbeg; c < cend; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        tmem_ld_wait();
        if (!in_range) continue;
        uint4* dst = reinterpret_cast<uint4*>(optr2 + c);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float f[8];
          Vec<bf16> o;
          if (valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[v * 8 + i]);
            o.set(f);
            o.get(f);   // the normalisation sees the value as stored (bf16), like the separate pass did
            float rs[8];
            if (p.bnf_res_mode) {
              Vec<bf16> rv;
              rv.raw = *reinterpret_cast<const uint4*>(p.bnf_res + (size_t)orow * p.N + n0 + c + v * 8);
              rv.get(rs);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int cc = c + v * 8 + i;
              float y = fmaf(f[i], ep_scale[cc], ep_scale[BN + cc]);
              if (p.bnf_res_mode == 1) y += rs[i];
              if (p.bnf_res_mode == 2) y += fmaf(rs[i], ep_scale[2 * BN + cc], ep_scale[3 * BN + cc]);
              f[i] = p.bnf_relu ? fmaxf(y, 0.f) : y;
            }
            o.set(f);
          } else {
            o.zero();
          }
          dst[v] = o.raw;
        }
      }
      if (release) {
        tc_fence_before();
        mbar_arrive(release);
      }
    }
  }
  }

// ---------------------------------------------------------------------------------------------
// fprop / dgrad
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool STATS, bool B_MN, bool EPI, int RED = 0>
__global__ void __launch_bounds__(TN_THREADS, 1)
conv_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, bf16* __restrict__ out,
               const TnParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  SmemLayout s = carve<B_BYTES, STAGES>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(s.done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;
  pdl_wait();   // everything above is on-chip setup; global memory is first touched below
  const int KC = p.K / BK;
  const int iters = p.ntaps * KC;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int st = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&s.empty[st], ph ^ 1u);
        mbar_expect_tx(&s.full[st], A_BYTES + B_BYTES);
        const int tap = it / KC, kc = it - tap * KC;
        tma_load_2d(s.a + st * A_BYTES, &tmA, &s.full[st], kc * BK + p.a_col_off[tap], m0 + p.a_row_off[tap]);
        if (B_MN) {
          // weights [tap][K][N], N contiguous: one [BK rows x 64 columns] box per 64-column atom
#pragma unroll
          for (int b = 0; b < BN / 64; ++b)
            tma_load_2d(s.b + st * B_BYTES + b * (BK * 128), &tmB, &s.full[st], n0 + b * 64,
                        p.w_tap[tap] * p.w_rows_per_tap + kc * BK);
        } else {
          tma_load_2d(s.b + st * B_BYTES, &tmB, &s.full[st], kc * BK, p.w_tap[tap] * p.w_rows_per_tap + n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, false, B_MN);
      for (int it = 0; it < iters; ++it) {
        const int st = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&s.full[st], ph);
        if (it == 0) SSB_MARK();   // first operands landed
        tc_fence_after();
        const uint32_t a0 = smem_u32(s.a + st * A_BYTES), b0 = smem_u32(s.b + st * B_BYTES);
#pragma unroll
        for (int k4 = 0; k4 < BK / 16; ++k4) {
          // K-major, 128B swizzle: 8-row groups are 1024 B apart; advance 16 elements = 32 B along K.
          // MN-major B: 64-column atoms BK*128 B apart (LBO), 8-row K groups 1024 B apart (SBO); 16 K rows = 2048 B
          const uint64_t bdesc = B_MN ? make_smem_desc(b0 + k4 * 2048, BK * 128, 1024) : make_smem_desc(b0 + k4 * 32, 0, 1024);
          umma_bf16(tmem_base, make_smem_desc(a0 + k4 * 32, 0, 1024), bdesc, idesc, (uint32_t)((it | k4) != 0));
        }
        umma_commit(&s.empty[st]);   // frees the smem stage when these MMAs retire
      }
      umma_commit(s.done);           // accumulator complete
      SSB_MARK();                    // last MMA issued
    }
  } else {
    tn_epilogue<BN, STATS, EPI, RED>(p, out, tmem_base, s.done, 0u, nullptr, reinterpret_cast<float*>(s.tmem_slot + 4),
                                reinterpret_cast<float*>(s.a), m0, n0, warp, lane);
    if (threadIdx.x == 64) SSB_MARK();   // epilogue done
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// fprop / dgrad of the stride-1 k=3 convs with TAP REUSE: the three taps read the same activation
// rows shifted by one, so the A tile (BM + 2 halo rows, padded to 136) is loaded ONCE per 64-channel
// K chunk and the three MMAs address it at row offsets 0/1/2 through the smem descriptor start address
// (128 B per row; the swizzle follows the absolute address).  A traffic from L2 drops 3x; the B (weight) tiles
// have their own, deeper ring.  Tiles up to 128 x 256.
// ---------------------------------------------------------------------------------------------
constexpr int A3_ROWS = 136;
constexpr int A3_BYTES = A3_ROWS * 128;   // 17408 = 17 * 1024
constexpr int A3_SLOTS = 3;
template <int BN> __host__ __device__ constexpr int b3_slots() { return BN == 256 ? 5 : (BN == 128 ? 8 : 9); }
template <int BN> constexpr int smem3_bytes() {
  return A3_SLOTS * A3_BYTES + b3_slots<BN>() * BN * 128 + (2 * A3_SLOTS + 2 * b3_slots<BN>() + 4) * 8 + 16 + 4 * BN * 4 +
         ((BN == 256 ? 8 : 12) * BN < 768 ? 768 : (BN == 256 ? 8 : 12) * BN) * 4 + 1024;   // scratch: [4][2 | 3][BN] floats (three quantities:
                                                                                       // RED == 2, never launched with 256-wide tiles)
}

// PERSISTENT: gridDim.x CTAs walk the tile list (m fastest, so that CTAs running together share the weight
// tiles in L2); the accumulator is double-buffered in TMEM (2 x BN columns), so the MMAs of tile i+1 run under
// the epilogue of tile i; the TMA producer runs ahead across tile boundaries through the same rings.
template <int BN, bool STATS, bool B_MN, bool EPI, int RED = 0>
__global__ void __launch_bounds__(TN_THREADS, 1)
conv_tn3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, bf16* __restrict__ out,
                const TnParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int NB = b3_slots<BN>();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* sa = smem_raw + (((base + 1023u) & ~1023u) - base);
  uint8_t* sb = sa + A3_SLOTS * A3_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sb + NB * B_BYTES);
  uint64_t* a_empty = a_full + A3_SLOTS;
  uint64_t* b_full = a_empty + A3_SLOTS;
  uint64_t* b_empty = b_full + NB;
  uint64_t* t_full = b_empty + NB;      // [2] accumulator buffer complete (MMA -> epilogue)
  uint64_t* t_empty = t_full + 2;       // [2] accumulator buffer drained (epilogue -> MMA), one arrival per epilogue thread
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* ep_scale = reinterpret_cast<float*>(tmem_slot + 4);   // [4 * BN] per-column coefficients
  float* red = ep_scale + 4 * BN;                               // [4][2][BN] statistics scratch
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < A3_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], EPI_THREADS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int KC = p.K / BK;
  const int MT = (p.M + BM - 1) / BM;
  const int ntiles = MT * (p.N / BN);

  if (warp == 0) {
    if (lane == 0) {
      int ai = 0, bi = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = (t % MT) * BM, n0 = (t / MT) * BN;
        for (int kc = 0; kc < KC; ++kc, ++ai) {
          const int sl = ai % A3_SLOTS;
          mbar_wait(&a_empty[sl], (((uint32_t)(ai / A3_SLOTS)) & 1u) ^ 1u);
          mbar_expect_tx(&a_full[sl], A3_BYTES);
          tma_load_2d(sa + sl * A3_BYTES, &tmA, &a_full[sl], kc * BK, m0 - 1);   // rows m0-1 .. m0+134 (zero fill outside)
          for (int tap = 0; tap < 3; ++tap, ++bi) {
            const int bs = bi % NB;
            mbar_wait(&b_empty[bs], (((uint32_t)(bi / NB)) & 1u) ^ 1u);
            mbar_expect_tx(&b_full[bs], B_BYTES);
            if (B_MN) {
#pragma unroll
              for (int b = 0; b < BN / 64; ++b)
                tma_load_2d(sb + bs * B_BYTES + b * (BK * 128), &tmB, &b_full[bs], n0 + b * 64,
                            p.w_tap[tap] * p.w_rows_per_tap + kc * BK);
            } else {
              tma_load_2d(sb + bs * B_BYTES, &tmB, &b_full[bs], kc * BK, p.w_tap[tap] * p.w_rows_per_tap + n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, false, B_MN);
      int ai = 0, bi = 0, it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&t_empty[buf], (((uint32_t)(it >> 1)) & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
        for (int kc = 0; kc < KC; ++kc, ++ai) {
          const int sl = ai % A3_SLOTS;
          mbar_wait(&a_full[sl], ((uint32_t)(ai / A3_SLOTS)) & 1u);
          const uint32_t a0 = smem_u32(sa + sl * A3_BYTES);
          for (int tap = 0; tap < 3; ++tap, ++bi) {
            const int bs = bi % NB;
            mbar_wait(&b_full[bs], ((uint32_t)(bi / NB)) & 1u);
            if (bi == 0) SSB_MARK();   // first operands landed
            tc_fence_after();
            const uint32_t b0 = smem_u32(sb + bs * B_BYTES);
            const uint32_t at = a0 + (uint32_t)(p.a_row_off[tap] + 1) * 128u;   // tap's row shift inside the haloed tile
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {
              const uint64_t bdesc = B_MN ? make_smem_desc(b0 + k4 * 2048, BK * 128, 1024) : make_smem_desc(b0 + k4 * 32, 0, 1024);
              umma_bf16(acc, make_smem_desc(at + k4 * 32, 0, 1024), bdesc, idesc, (uint32_t)((kc | tap | k4) != 0));
            }
            umma_commit(&b_empty[bs]);
          }
          umma_commit(&a_empty[sl]);
        }
        umma_commit(&t_full[buf]);
        SSB_MARK();                    // last MMA of a tile issued
      }
    }
  } else {
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const int m0 = (t % MT) * BM, n0 = (t / MT) * BN;
      if (it > 0) asm volatile("bar.sync 1, 256;" ::: "memory");   // previous tile's scratch / coefficient readers are done
      tn_epilogue<BN, STATS, EPI, RED>(p, out, tmem_base + (uint32_t)(buf * BN), &t_full[buf], ((uint32_t)(it >> 1)) & 1u,
                                  &t_empty[buf], ep_scale, red, m0, n0, warp, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * BN>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[tap][ci][co] += sum_rows X[rows + shift(tap)][ci] * dY[rows][co]
// GEMM rows (TMEM lanes) = ci tile of 128, columns = co tile of BN, reduction over rows.
// grid = (ci tiles * co tiles, taps, splits)
// ---------------------------------------------------------------------------------------------
struct WgParams {
  int M;            // output rows of the conv (reduction length)
  int Cin, Cout, k;
  int n_ci_tiles;
  int a_row_off[3], a_col_off[3], w_tap[3];
  int blocks_per_split;   // reduction blocks (of BK rows) per split
  int nsplit;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  float* __restrict__ dw, const WgParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  SmemLayout s = carve<B_BYTES, STAGES>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci0 = (blockIdx.x % p.n_ci_tiles) * BM;
  const int co0 = (blockIdx.x / p.n_ci_tiles) * BN;
  const int tap = blockIdx.y;
  const int nblk_total = (p.M + BK - 1) / BK;
  const int blk0 = blockIdx.z * p.blocks_per_split;
  const int blk1 = min(nblk_total, blk0 + p.blocks_per_split);
  const int iters = blk1 - blk0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(s.done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;
  pdl_wait();   // everything above is on-chip setup; global memory is first touched below

  if (iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&s.empty[st], ph ^ 1u);
          mbar_expect_tx(&s.full[st], A_BYTES + B_BYTES);
          const int r0 = (blk0 + it) * BK;
          // MN-major operand tiles: [64-channel atom][BK rows][128 B]
#pragma unroll
          for (int a = 0; a < BM / 64; ++a)
            tma_load_2d(s.a + st * A_BYTES + a * (BK * 128), &tmX, &s.full[st], p.a_col_off[tap] + ci0 + a * 64,
                        r0 + p.a_row_off[tap]);
#pragma unroll
          for (int b = 0; b < BN / 64; ++b)
            tma_load_2d(s.b + st * B_BYTES + b * (BK * 128), &tmDY, &s.full[st], co0 + b * 64, r0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(BN, true, true);
        for (int it = 0; it < iters; ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&s.full[st], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(s.a + st * A_BYTES), b0 = smem_u32(s.b + st * B_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4) {
            // MN-major, 128B swizzle: 64-element MN atoms are BK*128 B apart (LBO), 8-row K groups
            // 1024 B apart (SBO); advancing 16 reduction rows = 2048 B
            umma_bf16(tmem_base, make_smem_desc(a0 + k4 * 2048, BK * 128, 1024),
                      make_smem_desc(b0 + k4 * 2048, BK * 128, 1024), idesc, (uint32_t)((it | k4) != 0));
          }
          umma_commit(&s.empty[st]);
        }
        umma_commit(s.done);
      }
    } else {
      const int q = warp & 3;
      mbar_wait(s.done, 0);
      tc_fence_after();
      const int ci = ci0 + q * 32 + lane;
      const int wt = p.w_tap[tap];
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        tmem_ld_wait();
        if (ci >= p.Cin) continue;
        // dw is [tap][Cin][Cout]: this thread's 32 columns are 128 contiguous bytes
        float* dst = dw + ((size_t)wt * p.Cin + ci) * p.Cout + co0 + c;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (p.nsplit > 1) {
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                         "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                         : "memory");
          } else {
            float4 o = *reinterpret_cast<float4*>(dst + j);
            o.x += __uint_as_float(r[j]);
            o.y += __uint_as_float(r[j + 1]);
            o.z += __uint_as_float(r[j + 2]);
            o.w += __uint_as_float(r[j + 3]);
            *reinterpret_cast<float4*>(dst + j) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// wgrad of the stride-1 k=3 convs with TAP FUSION: the three taps' weight gradients are
//   dW_tap[ci, co] = sum_r X[r + tap - 1][ci] * dY[r][co]
// i.e. the SAME dY tile against the X tile shifted by one row.  One CTA keeps three accumulators
// (3 x BN TMEM columns), loads dY once and X once (BK + 2 halo rows, padded to 72) per reduction block
// and issues the three taps' MMAs with the A descriptor offset by 0/1/2 rows of 128 B (MN-major, the
// swizzle follows the absolute address as in conv_tn3_kernel).  Operand traffic per FLOP drops ~2.5x.
// grid = (ci tiles * co tiles, 1, splits)
// ---------------------------------------------------------------------------------------------
constexpr int XW_ROWS = BK + 8;                 // 72-row TMA box: rows r0-1 .. r0+70
constexpr int XW_ATOM = XW_ROWS * 128;          // bytes of one 64-channel atom of the X tile (9 * 1024)
constexpr int XW_BYTES = 2 * XW_ATOM;           // 128 ci
constexpr int WG3_STAGES = 4;
template <int BN> constexpr int smem_wg3_bytes() {
  return WG3_STAGES * (XW_BYTES + BN * BK * 2) + (2 * WG3_STAGES + 1) * 8 + 16 + 1024;
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_wgrad3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                   float* __restrict__ dw, const WgParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGES = WG3_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* sa = smem_raw + (((base + 1023u) & ~1023u) - base);
  uint8_t* sb = sa + STAGES * XW_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sb + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci0 = (blockIdx.x % p.n_ci_tiles) * BM;
  const int co0 = (blockIdx.x / p.n_ci_tiles) * BN;
  const int nblk_total = (p.M + BK - 1) / BK;
  const int blk0 = blockIdx.z * p.blocks_per_split;
  const int blk1 = min(nblk_total, blk0 + p.blocks_per_split);
  const int iters = blk1 - blk0;
  constexpr int TCOLS = (3 * BN <= 256) ? 256 : 512;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&empty[st], ph ^ 1u);
          mbar_expect_tx(&full[st], XW_BYTES + B_BYTES);
          const int r0 = (blk0 + it) * BK;
#pragma unroll
          for (int a = 0; a < BM / 64; ++a)     // X rows r0-1 .. r0+70 of a 64-channel atom (zero fill outside)
            tma_load_2d(sa + st * XW_BYTES + a * XW_ATOM, &tmX, &full[st], ci0 + a * 64, r0 - 1);
#pragma unroll
          for (int b = 0; b < BN / 64; ++b)
            tma_load_2d(sb + st * B_BYTES + b * (BK * 128), &tmDY, &full[st], co0 + b * 64, r0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(BN, true, true);
        for (int it = 0; it < iters; ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sa + st * XW_BYTES), b0 = smem_u32(sb + st * B_BYTES);
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {
              // tap t pairs dY row r with X row r + t - 1 = box row (r - r0) + t
              umma_bf16(tmem_base + (uint32_t)(tap * BN), make_smem_desc(a0 + tap * 128 + k4 * 2048, XW_ATOM, 1024),
                        make_smem_desc(b0 + k4 * 2048, BK * 128, 1024), idesc, (uint32_t)((it | k4) != 0));
            }
          }
          umma_commit(&empty[st]);
        }
        umma_commit(done);
      }
    } else {
      const int q = warp & 3;
      mbar_wait(done, 0);
      tc_fence_after();
      const int ci = ci0 + q * 32 + lane;
#pragma unroll 1
      for (int tap = 0; tap < 3; ++tap) {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * BN + c), r);
          tmem_ld_wait();
          if (ci >= p.Cin) continue;
          float* dst = dw + ((size_t)tap * p.Cin + ci) * p.Cout + co0 + c;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (p.nsplit > 1) {
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                           "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                           : "memory");
            } else {
              float4 o = *reinterpret_cast<float4*>(dst + j);
              o.x += __uint_as_float(r[j]);
              o.y += __uint_as_float(r[j + 1]);
              o.z += __uint_as_float(r[j + 2]);
              o.w += __uint_as_float(r[j + 3]);
              *reinterpret_cast<float4*>(dst + j) = o;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// Stem conv (k7, s2, p3) of MULTI-LEAD inputs as an implicit GEMM on the tensor cores (round 2).
//   D[rows, Cs] = A[rows, K] * W[K, Cs],  K = 7 * leads (84 at 12 leads, padded to a multiple of 16)
// The CUDA-core direct conv is FMA-bound there: 94-177 us per launch at 12 x 5000 (ncu: tensor pipe 0 %, 7 % of the HBM
// roofline, 9 % of the width-128 step), against ~10 us of memory time.  The input is fp32 NCL -- the reference's layout,
// not TMA-tileable as an im2col operand -- so the A tile is BUILT in shared memory by the threads: row r of a 128-row
// tile is one output position, its 7 * leads samples x[b][lead][2t-3 .. 2t+3] are converted to bf16 and stored in the
// 128-byte-swizzled K-major layout the UMMA descriptors of the other kernels read (byte offset r*128 + ((k/8) ^ (r%8))*16
// + (k%8)*2 inside a 64-column chunk; the swizzle follows the absolute address, tiles are 1024-byte aligned), then made
// visible to the async proxy (fence.proxy.async) before ONE thread issues K/16 MMAs.  Weights: converted from the fp32
// master [Cs][leads][7] = K-major rows once per CTA.  Persistent over the row tiles; two CTAs per SM overlap one CTA's
// tile build with the other's MMAs / epilogue.  Epilogue = tn_epilogue (halo rows zeroed, optional BN statistics).
// ---------------------------------------------------------------------------------------------
template <int BN, bool STATS>
__global__ void __launch_bounds__(TN_THREADS, 1)
stem_tn_kernel(const float* __restrict__ x, const float* __restrict__ w, bf16* __restrict__ out, const TnParams p, int Cl, int L,
               int KP) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* sa = smem_raw + (((base + 1023u) & ~1023u) - base);
  const int nchunk = (KP + 63) / 64;
  uint8_t* sb = sa + nchunk * A_BYTES;
  uint64_t* done = reinterpret_cast<uint64_t*>(sb + nchunk * BN * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  float* ep_scale = reinterpret_cast<float*>(tmem_slot + 4);   // [4 * BN] (unused by this epilogue mode)
  float* red = ep_scale + 4 * BN;                               // [8 * BN] statistics scratch
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int K = 7 * Cl;
  // weights -> bf16, swizzled K-major rows (row n = output channel); pad columns of A zeroed once
  for (int i = threadIdx.x; i < BN * KP; i += TN_THREADS) {
    const int n = i / KP, k = i - n * KP;
    const float v = k < K ? w[(size_t)n * K + k] : 0.f;
    const int kk = k & 63;
    *reinterpret_cast<bf16*>(sb + (k >> 6) * (BN * 128) + n * 128 + ((((kk >> 3) ^ (n & 7))) << 4) + (kk & 7) * 2) = __float2bfloat16_rn(v);
  }
  for (int i = threadIdx.x; i < BM * (KP - K); i += TN_THREADS) {
    const int r = i % BM, k = K + i / BM;
    const int kk = k & 63;
    *reinterpret_cast<bf16*>(sa + (k >> 6) * A_BYTES + r * 128 + ((((kk >> 3) ^ (r & 7))) << 4) + (kk & 7) * 2) = __float2bfloat16_rn(0.f);
  }
  const int MT = (p.M + BM - 1) / BM;
  constexpr uint32_t idesc = make_idesc(BN, false, false);
  int it = 0;
  for (int t = blockIdx.x; t < MT; t += gridDim.x, ++it) {
    const int m0 = t * BM;
    // ---- build the im2col tile: (row, lead) pairs, rows fastest (neighbouring threads read neighbouring samples) ----
    for (int i = threadIdx.x; i < BM * Cl; i += TN_THREADS) {
      const int r = i % BM, lead = i / BM;
      const int m = m0 + r;
      float v[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) v[j] = 0.f;
      if (m < p.M) {
        const int b = m / p.o_pitch, pos = m - b * p.o_pitch;
        if (pos >= 1 && pos <= p.o_len) {
          const float* src = x + ((size_t)b * Cl + lead) * L;
          const int s0 = 2 * (pos - 1) - 3;
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            const int sidx = s0 + j;
            if (sidx >= 0 && sidx < L) v[j] = __ldg(src + sidx);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int k = lead * 7 + j, kk = k & 63;
        *reinterpret_cast<bf16*>(sa + (k >> 6) * A_BYTES + r * 128 + ((((kk >> 3) ^ (r & 7))) << 4) + (kk & 7) * 2) = __float2bfloat16_rn(v[j]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (warp == 1) {
      if (lane == 0) {
        tc_fence_after();
        const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
        for (int kc = 0; kc < KP / 16; ++kc) {
          const int ch = kc >> 2, k4 = kc & 3;
          umma_bf16(tmem_base, make_smem_desc(a0 + ch * A_BYTES + k4 * 32, 0, 1024),
                    make_smem_desc(b0 + ch * (BN * 128) + k4 * 32, 0, 1024), idesc, (uint32_t)(kc != 0));
        }
        umma_commit(done);
      }
    } else if (warp >= 2) {
      tn_epilogue<BN, STATS, false, 0>(p, out, tmem_base, done, (uint32_t)(it & 1), nullptr, ep_scale, red, m0, 0, warp, lane);
    }
    tc_fence_before();
    __syncthreads();     // accumulator drained, operand tile free
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// Stem weight gradient of multi-lead inputs on the tensor cores:  dW[co][k] += sum_rows dY[row][co] * A[row][k],
// k = lead * 7 + tap.  Same operand roles as conv_wgrad_kernel (both MN-major, reduction over rows, split over rows across
// CTAs): the accumulator has the 7*leads <= 112 im2col columns on its lanes and the stem channels on its columns; dY
// tiles come by TMA, the im2col tile [64 rows][128 columns] is built in shared memory from the fp32 NCL input like in
// stem_tn_kernel.  Epilogue: the accumulator is transposed through shared memory so that each output channel's 7*leads
// contiguous gradients go out as 16-byte vector REDs (scalar REDs when 7*leads is not a multiple of 4).
// The CUDA-core kernel took 173 us at 32+32 x 12 x 5000 (ncu), ~51 % of the remaining stem time.
// ---------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
stem_wgrad_tn_kernel(const float* __restrict__ x, const __grid_constant__ CUtensorMap tmDY, float* __restrict__ dw, int Cl, int L,
                     int M, int pitch, int len, int blocks_per_split) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int X_BYTES = 2 * BK * 128;           // [2 atoms of 64 columns][64 rows][128 B]
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* sa = smem_raw + (((base + 1023u) & ~1023u) - base);
  uint8_t* sb = sa + X_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sb + B_BYTES);
  uint64_t* mma_done = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 1);
  float* stage = reinterpret_cast<float*>(tmem_slot + 4);     // [32 channels][132] transposed accumulator chunk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = 7 * Cl;
  const int nblk_total = (M + BK - 1) / BK;
  const int blk0 = blockIdx.x * blocks_per_split;
  const int blk1 = min(nblk_total, blk0 + blocks_per_split);
  const int iters = blk1 - blk0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    mbar_init(full, 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (iters > 0) {
    // pad columns K .. 127 of the im2col tile: zero, once
    for (int i = threadIdx.x; i < BK * (128 - K); i += NTHREADS) {
      const int r = i % BK, k = K + i / BK;
      const int kk = k & 63;
      *reinterpret_cast<bf16*>(sa + (k >> 6) * (BK * 128) + r * 128 + ((((kk >> 3) ^ (r & 7))) << 4) + (kk & 7) * 2) = __float2bfloat16_rn(0.f);
    }
    constexpr uint32_t idesc = make_idesc(BN, true, true);
    for (int it = 0; it < iters; ++it) {
      const int r0 = (blk0 + it) * BK;
      if (warp == 0 && lane == 0) {
        mbar_expect_tx(full, B_BYTES);
#pragma unroll
        for (int b = 0; b < BN / 64; ++b) tma_load_2d(sb + b * (BK * 128), &tmDY, full, b * 64, r0);
      }
      for (int i = threadIdx.x; i < BK * Cl; i += NTHREADS) {
        const int r = i % BK, lead = i / BK;
        const int m = r0 + r;
        float v[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) v[j] = 0.f;
        if (m < M) {
          const int b = m / pitch, pos = m - b * pitch;
          if (pos >= 1 && pos <= len) {
            const float* src = x + ((size_t)b * Cl + lead) * L;
            const int s0 = 2 * (pos - 1) - 3;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
              const int sidx = s0 + j;
              if (sidx >= 0 && sidx < L) v[j] = __ldg(src + sidx);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const int k = lead * 7 + j, kk = k & 63;
          *reinterpret_cast<bf16*>(sa + (k >> 6) * (BK * 128) + r * 128 + ((((kk >> 3) ^ (r & 7))) << 4) + (kk & 7) * 2) = __float2bfloat16_rn(v[j]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (warp == 1 && lane == 0) {
        mbar_wait(full, (uint32_t)(it & 1));
        tc_fence_after();
        const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
#pragma unroll
        for (int k4 = 0; k4 < BK / 16; ++k4)
          umma_bf16(tmem_base, make_smem_desc(a0 + k4 * 2048, BK * 128, 1024), make_smem_desc(b0 + k4 * 2048, BK * 128, 1024), idesc,
                    (uint32_t)((it | k4) != 0));
        umma_commit(mma_done);
      }
      mbar_wait(mma_done, (uint32_t)(it & 1));     // the MMAs have read both tiles: they may be overwritten
    }
    // ---- epilogue: lanes = im2col columns k, TMEM columns = channels; transposed through shared memory ----
    tc_fence_after();
    const bool vec = (K & 3) == 0;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (warp >= 2) {
        const int q = warp & 3;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        tmem_ld_wait();
        const int k = q * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[j * 132 + k] = __uint_as_float(r[j]);
      }
      __syncthreads();
      if (vec) {
        const int nv = K / 4;
        for (int i = threadIdx.x; i < 32 * nv; i += NTHREADS) {
          const int j = i / nv, v4 = i - j * nv;
          const float* sp_ = stage + j * 132 + v4 * 4;
          float* dst = dw + (size_t)(c + j) * K + v4 * 4;
          asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(sp_[0]), "f"(sp_[1]), "f"(sp_[2]),
                       "f"(sp_[3]) : "memory");
        }
      } else {
        for (int i = threadIdx.x; i < 32 * K; i += NTHREADS) {
          const int j = i / K, k = i - j * K;
          atomicAdd(dw + (size_t)(c + j) * K + k, stage[j * 132 + k]);
        }
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

template <int BN> constexpr int smem_stem_wgrad_bytes() { return 2 * BK * 128 + BN * BK * 2 + 16 + 16 + 32 * 132 * 4 + 1024 + 64; }

template <int BN> constexpr int smem_stem_bytes(int nchunk) { return nchunk * (A_BYTES + BN * 128) + 8 + 16 + 12 * BN * 4 + 1024 + 64; }

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows of `pitch_elems` elements,
// box = box_inner x box_outer, 128-byte swizzle, out-of-bounds elements read as zero
int make_map(CUtensorMap* map, const void* base, long long inner, long long outer, long long pitch_elems, int box_inner,
             int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    ssb_set_error("cuTensorMapEncodeTiled entry point not available");
    return SSB_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ssb_set_error("cuTensorMapEncodeTiled failed (%d): inner=%lld outer=%lld pitch=%lld box=%dx%d base=%p", (int)r, inner,
                  outer, pitch_elems, box_inner, box_outer, base);
    return SSB_ERR_CUDA;
  }
  return SSB_OK;
}

constexpr int TN_STAGES = 4;
constexpr int WG_STAGES = 4;

// every opt-in shared-memory size must fit the 227 KB a CTA can have (cudaFuncSetAttribute fails at ssb_prepare otherwise --
// which no CPU-side build step notices)
constexpr int SMEM_MAX = 227 * 1024;
static_assert(smem3_bytes<64>() <= SMEM_MAX && smem3_bytes<128>() <= SMEM_MAX && smem3_bytes<256>() <= SMEM_MAX, "conv_tn3 smem");
static_assert(smem_bytes<128 * BK * 2, TN_STAGES>() <= SMEM_MAX && smem_bytes<128 * BK * 2, WG_STAGES>() <= SMEM_MAX, "conv_tn / wgrad smem");
static_assert(smem_wg3_bytes<128>() <= SMEM_MAX && smem_stem_bytes<128>(2) <= SMEM_MAX && smem_stem_wgrad_bytes<128>() <= SMEM_MAX, "wgrad3 / stem smem");

#define g_num_sms g_ssb_num_sms      // (set by ssb_prepare, api.cu)

template <int BN, bool B_MN>
int launch_tn(const CUtensorMap& tmA, const CUtensorMap& tmB, bf16* out, const TnParams& p, cudaStream_t st) {
  constexpr int smem = smem_bytes<BN * BK * 2, TN_STAGES>();
  dim3 grid(ceil_div(p.M, BM), p.N / BN);
  if (p.ep_gamma && p.stats) {
    if constexpr (B_MN)
      ssb_launch_pro(conv_tn_kernel<BN, TN_STAGES, true, true, true>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
  } else if (p.ep_gamma) {
    if constexpr (B_MN)
      ssb_launch_pro(conv_tn_kernel<BN, TN_STAGES, false, true, true>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
  } else if (p.stats) {
    if constexpr (B_MN) {
      TnParams q = p;
      if (p.bnf_barrier) {   // grid-wide barrier inside: every CTA of the launch must be resident at once
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_tn_kernel<BN, TN_STAGES, true, true, false, 3>, TN_THREADS, smem) != cudaSuccess)
          occ = 0;
        q.bnf_expected = grid.x * grid.y;
        if ((long long)q.bnf_expected > (long long)occ * g_num_sms) {
          ssb_set_error("conv + BN fusion: %u CTAs exceed the co-resident capacity (%d per SM)", q.bnf_expected, occ);
          return SSB_ERR_UNSUPPORTED;
        }
      }
      if (p.bnf_barrier)
        ssb_launch_pro(conv_tn_kernel<BN, TN_STAGES, true, true, false, 3>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, q);
      else
        ssb_launch_pro(conv_tn_kernel<BN, TN_STAGES, true, true, false>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, q);
    }
  } else if (p.red_sums) {
    if constexpr (!B_MN) {
      if (p.red_xr)
        ssb_launch_pro(conv_tn_kernel<BN, TN_STAGES, false, false, false, 2>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
      else
        ssb_launch_pro(conv_tn_kernel<BN, TN_STAGES, false, false, false, 1>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
    }
  } else {
    ssb_launch_pro(conv_tn_kernel<BN, TN_STAGES, false, B_MN, false>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
  }
  SSB_LAUNCH_CHECK("conv_tn_kernel");
  return SSB_OK;
}

template <int BN, bool B_MN>
int launch_tn3(const CUtensorMap& tmA, const CUtensorMap& tmB, bf16* out, const TnParams& p, cudaStream_t st) {
  constexpr int smem = smem3_bytes<BN>();
  const int ntiles = ceil_div(p.M, BM) * (p.N / BN);
  dim3 grid(ntiles < g_num_sms ? ntiles : g_num_sms);
  if (p.ep_gamma && p.stats) {
    if constexpr (B_MN)
      ssb_launch_pro(conv_tn3_kernel<BN, true, true, true>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
  } else if (p.ep_gamma) {
    if constexpr (B_MN)
      ssb_launch_pro(conv_tn3_kernel<BN, false, true, true>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
  } else if (p.stats) {
    if constexpr (B_MN) {
      TnParams q = p;
      if (p.bnf_barrier) {   // grid-wide barrier inside: one tile per CTA, every CTA resident (1 per SM)
        q.bnf_expected = (unsigned int)ntiles;
        if (ntiles > g_num_sms) {
          ssb_set_error("conv + BN fusion: %d tiles exceed the %d co-resident CTAs of the persistent kernel", ntiles, g_num_sms);
          return SSB_ERR_UNSUPPORTED;
        }
      }
      if (p.bnf_barrier)
        ssb_launch_pro(conv_tn3_kernel<BN, true, true, false, 3>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, q);
      else
        ssb_launch_pro(conv_tn3_kernel<BN, true, true, false>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, q);
    }
  } else if (p.red_sums) {
    if constexpr (!B_MN) {
      if (p.red_xr)
        ssb_launch_pro(conv_tn3_kernel<BN, false, false, false, 2>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
      else
        ssb_launch_pro(conv_tn3_kernel<BN, false, false, false, 1>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
    }
  } else {
    ssb_launch_pro(conv_tn3_kernel<BN, false, B_MN, false>, dim3(grid), dim3(TN_THREADS), smem, st, tmA, tmB, out, p);
  }
  SSB_LAUNCH_CHECK("conv_tn3_kernel");
  return SSB_OK;
}

// Split count of a weight-gradient launch: `tiles` independent output tiles, each reducing over `nblk` row blocks, one
// (BN=128) or two (BN=64) CTAs per SM.  ceil(#SMs / tiles) splits -- the round-1 rule -- often lands just above a
// whole number of waves (16 tiles x 3 taps x 4 splits = 192 CTAs on 148 SMs = two waves of 11 blocks instead of one wave
// of 14: ncu, profiles/r2_conv_w128_launches.md).  Pick the count that minimises waves x (blocks per CTA + fixed cost),
// the fixed cost (pipeline fill + the RED epilogue of the partial tile) expressed in reduction blocks.
int pick_splits(int tiles, int nblk, int ctas_per_sm, int fixed_blocks) {
  const int cap = g_num_sms * ctas_per_sm;
  int max_ns = nblk / 4;
  if (max_ns < 1) max_ns = 1;
  int best = 1;
  long long best_cost = -1;
  for (int ns = 1; ns <= max_ns; ++ns) {
    const long long ctas = (long long)tiles * ns;
    if (ctas > 4LL * cap && ns > 1) break;
    const long long waves = (ctas + cap - 1) / cap;
    const long long cost = waves * (ceil_div(nblk, ns) + fixed_blocks);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ns; }
  }
  return best;
}
int g_wg_split_rule = 1;   // env SSB_WG_SPLIT=0: the round-1 rule (A/B comparison)
int g_wg_fixed = 4, g_wg3_fixed = 8;   // env SSB_WG_FIXED / SSB_WG3_FIXED: fixed per-CTA cost in reduction blocks

int g_tn3 = 1;   // env SSB_TN3=0 disables the tap-reuse kernel (A/B comparison)
int g_wg3 = 1;   // env SSB_WG3=0 disables the tap-fused wgrad kernel

// b_mn: the weight matrix of one tap is [K rows][N contiguous] (fprop) instead of [N rows][K contiguous] (dgrad)
// tap3: stride-1 k=3 conv whose taps are the row shifts -1/0/+1 of the same operand -> tap-reuse kernel
int run_tn(const void* a_base, long long a_inner, long long a_outer, long long a_pitch, const void* w_base, int w_rows,
           bool b_mn, bool tap3, bf16* out, const TnParams& p, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  int rc;
  // (K = 64: a single K chunk -- nothing to reuse, and the smaller v1 footprint lets two CTAs share an SM)
  if (tap3 && g_tn3 && p.K >= 128) {
    // widest tile that still gives every SM a CTA
    const int mt = ceil_div(p.M, BM);
    int BN = 64;
    if (p.N % 256 == 0 && (long long)mt * (p.N / 256) >= g_num_sms && !p.red_xr) BN = 256;   // (RED == 2 needs 12 x BN floats of scratch)
    else if (p.N % 128 == 0) BN = 128;
    rc = make_map(&tmA, a_base, a_inner, a_outer, a_pitch, BK, A3_ROWS);
    if (rc) return rc;
    if (b_mn) {
      rc = make_map(&tmB, w_base, p.N, w_rows, p.N, 64, BK);
      if (rc) return rc;
      return BN == 256 ? launch_tn3<256, true>(tmA, tmB, out, p, st)
                       : (BN == 128 ? launch_tn3<128, true>(tmA, tmB, out, p, st) : launch_tn3<64, true>(tmA, tmB, out, p, st));
    }
    rc = make_map(&tmB, w_base, p.K, w_rows, p.K, BK, BN);
    if (rc) return rc;
    return BN == 256 ? launch_tn3<256, false>(tmA, tmB, out, p, st)
                     : (BN == 128 ? launch_tn3<128, false>(tmA, tmB, out, p, st) : launch_tn3<64, false>(tmA, tmB, out, p, st));
  }
  const int BN = (p.N % 128 == 0) ? 128 : 64;
  rc = make_map(&tmA, a_base, a_inner, a_outer, a_pitch, BK, BM);
  if (rc) return rc;
  if (b_mn) {
    rc = make_map(&tmB, w_base, p.N, w_rows, p.N, 64, BK);
    if (rc) return rc;
    return BN == 128 ? launch_tn<128, true>(tmA, tmB, out, p, st) : launch_tn<64, true>(tmA, tmB, out, p, st);
  }
  rc = make_map(&tmB, w_base, p.K, w_rows, p.K, BK, BN);
  if (rc) return rc;
  return BN == 128 ? launch_tn<128, false>(tmA, tmB, out, p, st) : launch_tn<64, false>(tmA, tmB, out, p, st);
}

int check_sm100_shape(const char* who, const ssb_geom& gi, const ssb_geom& go) {
  SSB_REQUIRE(gi.C % 64 == 0 && go.C % 64 == 0, "%s: tcgen05 path needs Cin, Cout multiples of 64 (got %d, %d)", who, gi.C, go.C);
  return SSB_OK;
}

}  // namespace

template <typename T>
__global__ void zero_parity_rows_sm100(T* out, int rows, int N, int parity) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)(rows / 2) * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long q = idx / N;
    const int n = (int)(idx - q * N);
    out[(size_t)(2 * q + parity) * N + n] = from_f<T>(0.f);
  }
}

int ssb_sm100_prepare() {
  if (!get_encode()) {   // resolve the driver entry point outside of any stream capture
    ssb_set_error("ssb_sm100_prepare: cuTensorMapEncodeTiled entry point not available");
    return SSB_ERR_CUDA;
  }
  cudaError_t e = cudaSuccess;
#define SSB_TN_ATTR(BN_, ST_, MN_, EP_)                                                                                    \
  if (e == cudaSuccess)                                                                                                      \
    e = cudaFuncSetAttribute(conv_tn_kernel<BN_, TN_STAGES, ST_, MN_, EP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             smem_bytes<BN_ * BK * 2, TN_STAGES>());
  SSB_TN_ATTR(128, false, false, false) SSB_TN_ATTR(64, false, false, false) SSB_TN_ATTR(128, false, true, false)
  SSB_TN_ATTR(64, false, true, false) SSB_TN_ATTR(128, true, true, false) SSB_TN_ATTR(64, true, true, false)
  SSB_TN_ATTR(128, false, true, true) SSB_TN_ATTR(64, false, true, true)
  SSB_TN_ATTR(128, true, true, true) SSB_TN_ATTR(64, true, true, true)
#undef SSB_TN_ATTR
#define SSB_TNR_ATTR(BN_, R_)                                                                                               \
  if (e == cudaSuccess)                                                                                                     \
    e = cudaFuncSetAttribute(conv_tn_kernel<BN_, TN_STAGES, false, false, false, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             smem_bytes<BN_ * BK * 2, TN_STAGES>());
  SSB_TNR_ATTR(128, 1) SSB_TNR_ATTR(64, 1) SSB_TNR_ATTR(128, 2) SSB_TNR_ATTR(64, 2)
#undef SSB_TNR_ATTR
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_tn_kernel<128, TN_STAGES, true, true, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_bytes<128 * BK * 2, TN_STAGES>());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_tn_kernel<64, TN_STAGES, true, true, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_bytes<64 * BK * 2, TN_STAGES>());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_tn3_kernel<64, true, true, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3_bytes<64>());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_tn3_kernel<128, true, true, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3_bytes<128>());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_tn3_kernel<256, true, true, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3_bytes<256>());
#define SSB_TN3_ATTR(BN_, ST_, MN_, EP_)                                                                               \
  if (e == cudaSuccess)                                                                                              \
    e = cudaFuncSetAttribute(conv_tn3_kernel<BN_, ST_, MN_, EP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3_bytes<BN_>());
  SSB_TN3_ATTR(64, false, false, false) SSB_TN3_ATTR(128, false, false, false) SSB_TN3_ATTR(256, false, false, false)
  SSB_TN3_ATTR(64, false, true, false) SSB_TN3_ATTR(128, false, true, false) SSB_TN3_ATTR(256, false, true, false)
  SSB_TN3_ATTR(64, true, true, false) SSB_TN3_ATTR(128, true, true, false) SSB_TN3_ATTR(256, true, true, false)
  SSB_TN3_ATTR(64, false, true, true) SSB_TN3_ATTR(128, false, true, true) SSB_TN3_ATTR(256, false, true, true)
  SSB_TN3_ATTR(64, true, true, true) SSB_TN3_ATTR(128, true, true, true) SSB_TN3_ATTR(256, true, true, true)
#undef SSB_TN3_ATTR
#define SSB_TN3R_ATTR(BN_, R_)                                                                                        \
  if (e == cudaSuccess)                                                                                               \
    e = cudaFuncSetAttribute(conv_tn3_kernel<BN_, false, false, false, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3_bytes<BN_>());
  SSB_TN3R_ATTR(64, 1) SSB_TN3R_ATTR(128, 1) SSB_TN3R_ATTR(256, 1) SSB_TN3R_ATTR(64, 2) SSB_TN3R_ATTR(128, 2) SSB_TN3R_ATTR(256, 2)
#undef SSB_TN3R_ATTR
  if (const char* t3 = getenv("SSB_TN3")) g_tn3 = atoi(t3);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_wgrad_kernel<128, WG_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_bytes<128 * BK * 2, WG_STAGES>());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_wgrad_kernel<64, WG_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_bytes<64 * BK * 2, WG_STAGES>());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_wgrad3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_wg3_bytes<128>());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_wgrad3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_wg3_bytes<64>());
  if (const char* w3 = getenv("SSB_WG3")) g_wg3 = atoi(w3);
  if (const char* ws = getenv("SSB_WG_SPLIT")) g_wg_split_rule = atoi(ws);
  if (const char* wf = getenv("SSB_WG_FIXED")) g_wg_fixed = atoi(wf);
  if (const char* wf = getenv("SSB_WG3_FIXED")) g_wg3_fixed = atoi(wf);
  if (e != cudaSuccess) {
    ssb_set_error("ssb_sm100_prepare: %s", cudaGetErrorString(e));
    return SSB_ERR_CUDA;
  }
  return SSB_OK;
}

// stem conv on the tensor cores (bf16 output); 0 = launched, 1 = shape not covered (caller keeps the direct kernel)
int ssb_stem_conv_fwd_sm100(const float* x, const float* w, void* y, int Cl, int L, ssb_geom g, double* stats, cudaStream_t st) {
  static const bool on = !(getenv("SSB_STEM_TC") && atoi(getenv("SSB_STEM_TC")) == 0);
  // one lead stays on the CUDA-core kernel: it multiplies the fp32 signal by the fp32 master weights (this kernel rounds both
  // to bf16, which at one lead is the whole receptive field of 7 values -- the step-parity tests hold the stem to the
  // fp32-operand figure), and it is at the launch floor already (fusing the statistics saved 2 us of 690)
  if (!on || Cl < 2 || (g.C != 64 && g.C != 128)) return 1;
  const int KP = (7 * Cl + 15) / 16 * 16;
  const int nchunk = (KP + 63) / 64;
  TnParams p = {};
  p.M = g.B * g.pitch;
  p.N = g.C;
  p.K = KP;
  p.o_mul = 1;
  p.o_off = 0;
  p.o_rows = p.M;
  p.o_pitch = g.pitch;
  p.o_len = g.len;
  p.accumulate = 0;
  p.stats = stats;
  const int mt = ceil_div(p.M, BM);
  cudaError_t e = cudaSuccess;
#define SSB_STEM_TC_LAUNCH(BN_, ST_)                                                                                          \
  {                                                                                                                           \
    const int smem = smem_stem_bytes<BN_>(nchunk);                                                                            \
    e = cudaFuncSetAttribute(stem_tn_kernel<BN_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                    \
    dim3 grid(mt < 2 * g_num_sms ? mt : 2 * g_num_sms);                                                                       \
    if (e == cudaSuccess)                                                                                                     \
      ssb_launch_pro(stem_tn_kernel<BN_, ST_>, dim3(grid), dim3(TN_THREADS), smem, st, x, w, (bf16*)y, p, Cl, L, KP);         \
  }
  if (g.C == 128) { if (stats) SSB_STEM_TC_LAUNCH(128, true) else SSB_STEM_TC_LAUNCH(128, false) }
  else { if (stats) SSB_STEM_TC_LAUNCH(64, true) else SSB_STEM_TC_LAUNCH(64, false) }
#undef SSB_STEM_TC_LAUNCH
  if (e != cudaSuccess) {
    ssb_set_error("ssb_stem_conv_fwd: %s", cudaGetErrorString(e));
    return SSB_ERR_CUDA;
  }
  SSB_LAUNCH_CHECK("stem_tn_kernel");
  return SSB_OK;
}

// stem weight gradient on the tensor cores; 0 = launched, 1 = shape not covered (caller keeps the CUDA-core kernel)
int ssb_stem_conv_wgrad_sm100(const float* x, const void* dy, float* dw, int Cl, int L, ssb_geom g, cudaStream_t st) {
  static const bool on = !(getenv("SSB_STEM_TC") && atoi(getenv("SSB_STEM_TC")) == 0);
  if (!on || Cl < 2 || (g.C != 64 && g.C != 128)) return 1;
  const int M = g.B * g.pitch;
  const int nblk = ceil_div(M, BK);
  int ns = pick_splits(1, nblk, 1, 12);
  const int bps = ceil_div(nblk, ns);
  ns = ceil_div(nblk, bps);
  CUtensorMap tmDY;
  int rc = make_map(&tmDY, dy, g.C, M, g.C, 64, BK);
  if (rc) return rc;
  cudaError_t e;
  if (g.C == 128) {
    e = cudaFuncSetAttribute(stem_wgrad_tn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_stem_wgrad_bytes<128>());
    if (e == cudaSuccess)
      ssb_launch_pro(stem_wgrad_tn_kernel<128>, dim3(ns), dim3(NTHREADS), smem_stem_wgrad_bytes<128>(), st, x, tmDY, dw, Cl, L, M, g.pitch, g.len, bps);
  } else {
    e = cudaFuncSetAttribute(stem_wgrad_tn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_stem_wgrad_bytes<64>());
    if (e == cudaSuccess)
      ssb_launch_pro(stem_wgrad_tn_kernel<64>, dim3(ns), dim3(NTHREADS), smem_stem_wgrad_bytes<64>(), st, x, tmDY, dw, Cl, L, M, g.pitch, g.len, bps);
  }
  if (e != cudaSuccess) {
    ssb_set_error("ssb_stem_conv_wgrad: %s", cudaGetErrorString(e));
    return SSB_ERR_CUDA;
  }
  SSB_LAUNCH_CHECK("stem_wgrad_tn_kernel");
  return SSB_OK;
}

int ssb_conv1d_fwd_sm100(const void* x, const void* w, void* y, ssb_geom gin, ssb_geom gout, int k, int stride,
                         double* stats, const ssb_bn* ep_bn, const void* ep_res, int ep_relu, int train_samples,
                         void* y_eval, const ssb_bnf_args* bnf, cudaStream_t st) {
  int rc = check_sm100_shape("ssb_conv1d_fwd", gin, gout);
  if (rc) return rc;
  TnParams p = {};
  p.M = gout.B * gout.pitch;
  p.N = gout.C;
  p.K = gin.C;
  p.ntaps = k;
  p.w_rows_per_tap = gin.C;   // weight rows of one tap in [k][Cin][Cout]
  p.o_mul = 1;
  p.o_off = 0;
  p.o_rows = p.M;
  p.o_pitch = gout.pitch;
  p.o_len = gout.len;
  p.accumulate = 0;
  p.stats = stats;
  if (ep_bn) {
    if (stats) {   // dual mode: the first train_samples samples are train rows, the rest eval rows
      SSB_REQUIRE(train_samples > 0 && train_samples < gout.B && y_eval, "ssb_conv1d_fwd: dual mode needs 0 < train_samples < B and y_eval");
      p.split_row = train_samples * gout.pitch;
      p.out2 = (bf16*)y_eval;
    }
    p.ep_gamma = ep_bn->gamma;
    p.ep_beta = ep_bn->beta;
    p.ep_mean = ep_bn->running_mean;
    p.ep_var = ep_bn->running_var;
    p.ep_res = (const bf16*)ep_res;
    p.ep_relu = ep_relu;
  }
  if (bnf) {
    p.bnf_barrier = bnf->barrier;
    p.bnf = *bnf->bn;
    if (bnf->bn_res) p.bnf_r = *bnf->bn_res;
    p.bnf_res = (const bf16*)bnf->res;
    p.bnf_res_mode = bnf->res ? (bnf->bn_res ? 2 : 1) : 0;
    p.bnf_relu = bnf->relu;
    p.bnf_n = (double)gout.B * (double)gout.len * (double)(bnf->bn->count_mul > 1 ? bnf->bn->count_mul : 1);
    p.out2 = (bf16*)bnf->y_act;
    p.stats = bnf->bn->sums;
  }
  const long long rows_in = (long long)gin.B * gin.pitch;
  long long a_inner, a_outer, a_pitch;
  if (stride == 1) {
    a_inner = gin.C; a_outer = rows_in; a_pitch = gin.C;
    for (int j = 0; j < k; ++j) { p.a_row_off[j] = (k == 3) ? j - 1 : 0; p.a_col_off[j] = 0; p.w_tap[j] = j; }
  } else {
    // row-pair view [rows/2, 2C]: input row 2q+par <-> (q, par*C); output row R <-> q = R-1
    a_inner = 2LL * gin.C; a_outer = rows_in / 2; a_pitch = 2LL * gin.C;
    if (k == 3) {
      p.a_row_off[0] = -1; p.a_col_off[0] = 0;     p.w_tap[0] = 0;
      p.a_row_off[1] = -1; p.a_col_off[1] = gin.C; p.w_tap[1] = 1;
      p.a_row_off[2] = 0;  p.a_col_off[2] = 0;     p.w_tap[2] = 2;
    } else {
      p.a_row_off[0] = -1; p.a_col_off[0] = gin.C; p.w_tap[0] = 0;
    }
  }
  return run_tn(x, a_inner, a_outer, a_pitch, w, k * gin.C, true, stride == 1 && k == 3, (bf16*)y, p, st);
}

// same kernel / tile selection as run_tn: can the conv + BN fusion keep every CTA of this conv resident?
int ssb_conv1d_fwd_bnf_fits_sm100(ssb_geom gin, ssb_geom gout, int k, int stride) {
  if (gin.C % 64 || gout.C % 64) return 0;
  const int M = gout.B * gout.pitch, N = gout.C, K = gin.C;
  const int mt = ceil_div(M, BM);
  if (stride == 1 && k == 3 && g_tn3 && K >= 128) {
    int BN = 64;
    if (N % 256 == 0 && (long long)mt * (N / 256) >= g_num_sms) BN = 256;
    else if (N % 128 == 0) BN = 128;
    return mt * (N / BN) <= g_num_sms ? 1 : 0;
  }
  int occ = 0;
  if (N % 128 == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_tn_kernel<128, TN_STAGES, true, true, false, 3>, TN_THREADS,
                                                      smem_bytes<128 * BK * 2, TN_STAGES>()) != cudaSuccess) return 0;
    return (long long)mt * (N / 128) <= (long long)occ * g_num_sms ? 1 : 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_tn_kernel<64, TN_STAGES, true, true, false, 3>, TN_THREADS,
                                                    smem_bytes<64 * BK * 2, TN_STAGES>()) != cudaSuccess) return 0;
  return (long long)mt * (N / 64) <= (long long)occ * g_num_sms ? 1 : 0;
}

int ssb_conv1d_dgrad_sm100(const void* dy, const void* w, void* dx, ssb_geom gin, ssb_geom gout, int k, int stride,
                           int accumulate, const void* red_y, const void* red_x, const ssb_bn* red_bn, const void* red_xr,
                           const ssb_bn* red_bn_r, cudaStream_t st) {
  int rc = check_sm100_shape("ssb_conv1d_dgrad", gin, gout);
  if (rc) return rc;
  const int rows_in = gin.B * gin.pitch, rows_out = gout.B * gout.pitch;
  TnParams p = {};
  p.N = gin.C;
  p.K = gout.C;
  p.w_rows_per_tap = gin.C;
  p.o_rows = rows_in;
  p.o_pitch = gin.pitch;
  p.o_len = gin.len;
  p.accumulate = accumulate;
  if (red_bn) {
    SSB_REQUIRE(stride == 1, "ssb_conv1d_dgrad: the fused BN-backward reduce needs a stride-1 conv (one launch writes dx)");
    p.red_y = (const bf16*)red_y;
    p.red_x = (const bf16*)red_x;
    p.red_mi = red_bn->mean_invstd;
    p.red_sums = red_bn->bwd_sums;
    if (red_bn_r) {
      p.red_xr = (const bf16*)red_xr;
      p.red_mi_r = red_bn_r->mean_invstd;
      p.red_sums_r = red_bn_r->bwd_sums;
    }
  }
  if (stride == 1) {
    p.M = rows_in;
    p.ntaps = k;
    p.o_mul = 1;
    p.o_off = 0;
    for (int j = 0; j < k; ++j) { p.a_row_off[j] = (k == 3) ? 1 - j : 0; p.a_col_off[j] = 0; p.w_tap[j] = j; }
    return run_tn(dy, gout.C, rows_out, gout.C, w, k * gin.C, false, k == 3, (bf16*)dx, p, st);
  }
  // stride 2: dx row 2q+par; par 0 <- dy[q+1]*W0 + dy[q]*W2; par 1 <- dy[q+1]*W1 (k=1: par 1 <- dy[q+1]*W0)
  p.M = rows_in / 2;
  p.o_mul = 2;
  for (int par = 0; par < 2; ++par) {
    p.o_off = par;
    p.ntaps = 0;
    if (k == 3) {
      if (par == 0) { p.ntaps = 2; p.a_row_off[0] = 1; p.w_tap[0] = 0; p.a_row_off[1] = 0; p.w_tap[1] = 2; }
      else { p.ntaps = 1; p.a_row_off[0] = 1; p.w_tap[0] = 1; }
    } else if (par == 1) {
      p.ntaps = 1; p.a_row_off[0] = 1; p.w_tap[0] = 0;
    }
    p.a_col_off[0] = p.a_col_off[1] = p.a_col_off[2] = 0;
    if (p.ntaps == 0) {
      if (!accumulate) {
        ssb_launch(zero_parity_rows_sm100<bf16>, dim3(148 * 2), dim3(256), 0, st, (bf16*)dx, rows_in, gin.C, par);
        SSB_LAUNCH_CHECK("zero_parity_rows_sm100");
      }
      continue;
    }
    rc = run_tn(dy, gout.C, rows_out, gout.C, w, k * gin.C, false, false, (bf16*)dx, p, st);
    if (rc) return rc;
  }
  return SSB_OK;
}

int ssb_conv1d_wgrad_sm100(const void* x, const void* dy, float* dw, ssb_geom gin, ssb_geom gout, int k, int stride,
                           cudaStream_t st) {
  int rc = check_sm100_shape("ssb_conv1d_wgrad", gin, gout);
  if (rc) return rc;
  WgParams p = {};
  p.M = gout.B * gout.pitch;
  p.Cin = gin.C;
  p.Cout = gout.C;
  p.k = k;
  p.n_ci_tiles = ceil_div(gin.C, BM);
  const long long rows_in = (long long)gin.B * gin.pitch;
  long long x_inner, x_outer, x_pitch;
  if (stride == 1) {
    x_inner = gin.C; x_outer = rows_in; x_pitch = gin.C;
    for (int j = 0; j < k; ++j) { p.a_row_off[j] = (k == 3) ? j - 1 : 0; p.a_col_off[j] = 0; p.w_tap[j] = j; }
  } else {
    x_inner = 2LL * gin.C; x_outer = rows_in / 2; x_pitch = 2LL * gin.C;
    if (k == 3) {
      p.a_row_off[0] = -1; p.a_col_off[0] = 0;     p.w_tap[0] = 0;
      p.a_row_off[1] = -1; p.a_col_off[1] = gin.C; p.w_tap[1] = 1;
      p.a_row_off[2] = 0;  p.a_col_off[2] = 0;     p.w_tap[2] = 2;
    } else {
      p.a_row_off[0] = -1; p.a_col_off[0] = gin.C; p.w_tap[0] = 0;
    }
  }
  const int BN = (gout.C % 128 == 0) ? 128 : 64;
  const int nblk = ceil_div(p.M, BK);
  const int tiles3 = p.n_ci_tiles * (gout.C / BN);
  // tap-fused kernel: one CTA = (ci tile, co tile), all three taps.  Worth it when, after splitting the reduction
  // to fill the GPU, a CTA still owns >= 16 reduction blocks (else the extra splits' REDs cost more than the
  // operand traffic saved: measured at the 16+16 x 2500 shapes, 147 -> 189 us per step)
  if (stride == 1 && k == 3 && g_wg3 && nblk / ceil_div(g_num_sms, tiles3) >= 16) {
    int ns = ceil_div(g_num_sms, tiles3);
    if (ns > nblk / 4) ns = nblk / 4;
    if (ns < 1) ns = 1;
    if (g_wg_split_rule) ns = pick_splits(tiles3, nblk, 1, g_wg3_fixed);
    p.blocks_per_split = ceil_div(nblk, ns);
    p.nsplit = ceil_div(nblk, p.blocks_per_split);
    CUtensorMap tmX3, tmDY3;
    rc = make_map(&tmX3, x, x_inner, x_outer, x_pitch, 64, XW_ROWS);
    if (rc) return rc;
    rc = make_map(&tmDY3, dy, gout.C, p.M, gout.C, 64, BK);
    if (rc) return rc;
    dim3 grid3(tiles3, 1, p.nsplit);
    if (BN == 128)
      ssb_launch_pro(conv_wgrad3_kernel<128>, dim3(grid3), dim3(NTHREADS), smem_wg3_bytes<128>(), st, tmX3, tmDY3, dw, p);
    else
      ssb_launch_pro(conv_wgrad3_kernel<64>, dim3(grid3), dim3(NTHREADS), smem_wg3_bytes<64>(), st, tmX3, tmDY3, dw, p);
    SSB_LAUNCH_CHECK("conv_wgrad3_kernel");
    return SSB_OK;
  }
  const int tiles = p.n_ci_tiles * (gout.C / BN) * k;
  int nsplit = ceil_div(g_num_sms, tiles);
  if (nsplit > nblk / 4) nsplit = nblk / 4;
  if (nsplit < 1) nsplit = 1;
  if (g_wg_split_rule) nsplit = pick_splits(tiles, nblk, BN == 64 ? 2 : 1, g_wg_fixed);
  p.blocks_per_split = ceil_div(nblk, nsplit);
  p.nsplit = ceil_div(nblk, p.blocks_per_split);
  CUtensorMap tmX, tmDY;
  // a 128-wide ci tile may run past Cin (into the other parity's columns of the pair view or out of
  // bounds = zeros); those accumulator rows are skipped by the epilogue (ci >= Cin)
  rc = make_map(&tmX, x, x_inner, x_outer, x_pitch, 64, BK);
  if (rc) return rc;
  rc = make_map(&tmDY, dy, gout.C, p.M, gout.C, 64, BK);
  if (rc) return rc;
  dim3 grid(p.n_ci_tiles * (gout.C / BN), k, p.nsplit);
  if (BN == 128) {
    ssb_launch_pro(conv_wgrad_kernel<128, WG_STAGES>, dim3(grid), dim3(NTHREADS), smem_bytes<128 * BK * 2, WG_STAGES>(), st, tmX, tmDY, dw, p);
  } else {
    ssb_launch_pro(conv_wgrad_kernel<64, WG_STAGES>, dim3(grid), dim3(NTHREADS), smem_bytes<64 * BK * 2, WG_STAGES>(), st, tmX, tmDY, dw, p);
  }
  SSB_LAUNCH_CHECK("conv_wgrad_kernel");
  return SSB_OK;
}

SSB_TRACE_DEFINE(conv_sm100)
