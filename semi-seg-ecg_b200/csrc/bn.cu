// BatchNorm1d (+ReLU, +residual, +MaxPool) forward/backward over flat padded NLC tensors.
// Replaces cuDNN/ATen native_batch_norm + relu_ + add_ + max_pool (+ their backward
// kernels) at resnet.py:41-70,254-257,354-355 and fcn_head.py:48-49.  HBM-bound: every
// kernel moves 16-byte vectors, channels fastest (coalesced), one pass per tensor.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <utility>

#include "common.cuh"

#define BN_THREADS 256

// ---------------------------------------------------------------------------------------
// statistics: sums[c] += sum_r x[r][c]; sums[C+c] += sum_r x[r][c]^2   (fp64 accumulators)
// halo / pad rows are zero by the layout invariant, so no masking is needed.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const T* __restrict__ x, int rows, int C,
                                                              double* __restrict__ sums, int rpb) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int ncg = C / V;
  const int cgb = ncg < BN_THREADS ? ncg : BN_THREADS;
  const int nrl = BN_THREADS / cgb;
  const int tid = threadIdx.x;
  const int cgl = tid % cgb, rl = tid / cgb;
  const int cg = blockIdx.y * cgb + cgl;
  __shared__ float sS[BN_THREADS * V];
  __shared__ float sQ[BN_THREADS * V];
  float s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = q[i] = 0.f;
  if (rl < nrl && cg < ncg) {
    const int r0 = blockIdx.x * rpb;
    const int r1 = min(rows, r0 + rpb);
    for (int r = r0 + rl; r < r1; r += nrl) {
      Vec<T> v;
      v.load(x + (size_t)r * C + (size_t)cg * V);
      float f[V];
      v.get(f);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s[i] += f[i];
        q[i] += f[i] * f[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sS[tid * V + i] = s[i];
    sQ[tid * V + i] = q[i];
  }
  __syncthreads();
  // first row-lane reduces over the row lanes in fp64, one thread per (cg, i)
  for (int o = tid; o < cgb * V; o += BN_THREADS) {
    const int l = o / V, i = o % V;
    const int cgo = blockIdx.y * cgb + l;
    if (cgo >= ncg) continue;
    // block-level partials in fp32 (<= 32 terms each already summed over a few rows); only the
    // cross-block accumulation is fp64 -- the fp64 pipe is slow and this loop is on every block's tail
    float ds = 0.f, dq = 0.f;
    for (int k = 0; k < nrl; ++k) {
      ds += sS[(k * cgb + l) * V + i];
      dq += sQ[(k * cgb + l) * V + i];
    }
    const int c = cgo * V + i;
    atomicAdd(&sums[c], (double)ds);
    atomicAdd(&sums[C + c], (double)dq);
  }
}

// ---------------------------------------------------------------------------------------
// y = [relu]( bn(x) [+ bn_res(res) | + res] ), zero on halo/pad rows
// RES: 0 none, 1 identity residual, 2 residual with its own BN
// grid = (row blocks, channel chunks of <= 64 channels).  A thread keeps ONE 16-byte channel
// group for all its rows, so the per-channel coefficients live in registers: no shared memory,
// no barrier, and the data loads are in flight while the coefficients are being computed.
// ---------------------------------------------------------------------------------------
template <typename T, int RES, bool SYNC>
__global__ void __launch_bounds__(BN_THREADS) bn_act_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                                T* __restrict__ y, ssb_bn bn, ssb_bn bnr,
                                                                ssb_geom g, int relu, int train, int cgpc, int rpb) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int C = g.C;
  const int ncg = C / V;
  const int rpp = BN_THREADS / cgpc;            // rows per pass
  const int cgl = threadIdx.x % cgpc, rl = threadIdx.x / cgpc;
  const int cg = blockIdx.y * cgpc + cgl;
  const int rows = g.B * g.pitch;
  const int r0 = blockIdx.x * rpb;
  const int r1 = min(rows, r0 + rpb);
  // the first row's operands are requested BEFORE the coefficient phase (statistics load -> fp64 arithmetic -> shared
  // memory -> barrier): two dependent memory round trips become one (most launches have one row per thread)
  const bool active = rl < rpp && cg < ncg;
  const int row0 = r0 + rl;
  const bool first_ok = active && row0 < r1 && row_valid(row0, g.pitch, g.len);
  Vec<T> pv, prv;
  if (first_ok) {
    const size_t off = (size_t)row0 * C + (size_t)cg * V;
    pv.load(x + off);
    if (RES != 0) prv.load(res + off);
  }
  // per-channel coefficients of this block's <= 64 channels: ONE thread per channel does the fp64
  // statistics arithmetic (fp64 issue rate is low: never replicate it across row lanes)
  __shared__ float sCo[4][64];
  const int nch = cgpc * V;
  if (threadIdx.x < nch) {
    const int c = blockIdx.y * nch + threadIdx.x;
    if (c < C) {
      const double n = (double)g.B * (double)g.len * (double)(bn.count_mul > 1 ? bn.count_mul : 1);
      const double inv_n = 1.0 / n;
      const bool writer = blockIdx.x == 0;
      float a, b;
      bn_coeffs<SYNC>(bn, c, C, train, inv_n, n, writer, a, b);
      sCo[0][threadIdx.x] = a;
      sCo[1][threadIdx.x] = b;
      if (RES == 2) {
        bn_coeffs<SYNC>(bnr, c, C, train, inv_n, n, writer, a, b);
        sCo[2][threadIdx.x] = a;
        sCo[3][threadIdx.x] = b;
      }
    }
  }
  __syncthreads();
  if (!active) return;
  float sc[V], sh[V], rsc[V], rsh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sc[i] = sCo[0][cgl * V + i];
    sh[i] = sCo[1][cgl * V + i];
    if (RES == 2) {
      rsc[i] = sCo[2][cgl * V + i];
      rsh[i] = sCo[3][cgl * V + i];
    }
  }
  for (int row = row0; row < r1; row += rpp) {
    const size_t off = (size_t)row * C + (size_t)cg * V;
    Vec<T> out;
    if (row_valid(row, g.pitch, g.len)) {
      Vec<T> v, rv;
      if (row == row0) {
        v = pv;
        if (RES != 0) rv = prv;
      } else {
        v.load(x + off);
        if (RES != 0) rv.load(res + off);
      }
      float f[V];
      v.get(f);
      float r[V];
      if (RES != 0) rv.get(r);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float o = fmaf(f[i], sc[i], sh[i]);
        if (RES == 1) o += r[i];
        if (RES == 2) o += fmaf(r[i], rsc[i], rsh[i]);
        if (relu) o = fmaxf(o, 0.f);
        f[i] = o;
      }
      out.set(f);
    } else {
      out.zero();
    }
    out.store(y + off);
  }
}

// ---------------------------------------------------------------------------------------
// stem tail: y[b, t] = max_{l in {2t-1,2t,2t+1} valid} relu(bn(c0[b, l]))
// arg (optional, train): per pooled element the window slot 0..2 of the first maximum (torch's
// max-pool tie rule), or 3 when the maximum is <= 0 (the ReLU kills the gradient there), so the
// backward pass routes gradients without recomputing the windows.
// ---------------------------------------------------------------------------------------
template <typename T, bool SYNC>
__global__ void __launch_bounds__(BN_THREADS) stem_bn_relu_pool_kernel(const T* __restrict__ c0, T* __restrict__ y,
                                                                       uint8_t* __restrict__ arg, ssb_bn bn,
                                                                       ssb_geom gi, ssb_geom go, int train) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  extern __shared__ float sm[];
  const int C = gi.C;
  float* sScale = sm;
  float* sShift = sm + C;
  const double n = (double)gi.B * (double)gi.len * (double)(bn.count_mul > 1 ? bn.count_mul : 1);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sc, sh;
    bn_coeffs<SYNC>(bn, c, C, train, 1.0 / n, n, blockIdx.x == 0, sc, sh);
    sScale[c] = sc;
    sShift[c] = sh;
  }
  __syncthreads();
  const int ncg = C / V;
  const long long total = (long long)go.B * go.pitch * ncg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / ncg);
    const int cg = (int)(idx - (long long)row * ncg);
    const int b = row / go.pitch, pos = row - b * go.pitch;
    Vec<T> out;
    uint8_t am[V];
    if (pos >= 1 && pos <= go.len) {
      const int t = pos - 1;
      float m[V];
#pragma unroll
      for (int i = 0; i < V; ++i) { m[i] = 0.f; am[i] = 3; }
#pragma unroll
      for (int j = -1; j <= 1; ++j) {
        const int l = 2 * t + j;
        if (l < 0 || l >= gi.len) continue;
        Vec<T> v;
        v.load(c0 + ((size_t)b * gi.pitch + 1 + l) * C + (size_t)cg * V);
        float f[V];
        v.get(f);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const int c = cg * V + i;
          const float a = fmaf(f[i], sScale[c], sShift[c]);
          if (a > m[i]) { m[i] = a; am[i] = (uint8_t)(j + 1); }   // strict >: first maximum wins
        }
      }
      out.set(m);
    } else {
      out.zero();
#pragma unroll
      for (int i = 0; i < V; ++i) am[i] = 3;
    }
    out.store(y + (size_t)row * C + (size_t)cg * V);
    if (arg) {
      uint8_t* ap = arg + (size_t)row * C + (size_t)cg * V;
#pragma unroll
      for (int i = 0; i < V; i += 4)
        *reinterpret_cast<uint32_t*>(ap + i) = (uint32_t)am[i] | ((uint32_t)am[i + 1] << 8) | ((uint32_t)am[i + 2] << 16) | ((uint32_t)am[i + 3] << 24);
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward pass 1 (reduce): g = (g1 [+g2]) * (y>0);  bwd_sums += (sum g, sum g*xhat)
// ---------------------------------------------------------------------------------------
template <typename T, bool HAS_G2, bool HAS_Y, bool HAS_RES>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const T* __restrict__ g1, const T* __restrict__ g2,
                                                                   const T* __restrict__ y, const T* __restrict__ x,
                                                                   const T* __restrict__ xr, ssb_bn bn, ssb_bn bnr,
                                                                   int rows, int C, int rpb) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int ncg = C / V;
  const int cgb = ncg < BN_THREADS ? ncg : BN_THREADS;
  const int nrl = BN_THREADS / cgb;
  const int tid = threadIdx.x;
  const int cgl = tid % cgb, rl = tid / cgb;
  const int cg = blockIdx.y * cgb + cgl;
  __shared__ float sA[BN_THREADS * V];
  __shared__ float sB[BN_THREADS * V];
  __shared__ float sC[HAS_RES ? BN_THREADS * V : 1];
  float a[V], bq[V], cq[V];
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = bq[i] = cq[i] = 0.f;
  if (rl < nrl && cg < ncg) {
    float mean[V], inv[V], meanr[V], invr[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = cg * V + i;
      mean[i] = bn.mean_invstd[c];
      inv[i] = bn.mean_invstd[C + c];
      if (HAS_RES) {
        meanr[i] = bnr.mean_invstd[c];
        invr[i] = bnr.mean_invstd[C + c];
      }
    }
    const int r0 = blockIdx.x * rpb;
    const int r1 = min(rows, r0 + rpb);
    for (int r = r0 + rl; r < r1; r += nrl) {
      const size_t off = (size_t)r * C + (size_t)cg * V;
      Vec<T> vg, vx;
      vg.load(g1 + off);
      vx.load(x + off);
      float fg[V], fx[V];
      vg.get(fg);
      vx.get(fx);
      if (HAS_G2) {
        Vec<T> v2;
        v2.load(g2 + off);
        float f2[V];
        v2.get(f2);
#pragma unroll
        for (int i = 0; i < V; ++i) fg[i] += f2[i];
      }
      if (HAS_Y) {
        Vec<T> vy;
        vy.load(y + off);
        float fy[V];
        vy.get(fy);
#pragma unroll
        for (int i = 0; i < V; ++i) fg[i] = fy[i] > 0.f ? fg[i] : 0.f;
      }
      float fr[V];
      if (HAS_RES) {
        Vec<T> vr;
        vr.load(xr + off);
        vr.get(fr);
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        a[i] += fg[i];
        bq[i] += fg[i] * ((fx[i] - mean[i]) * inv[i]);
        if (HAS_RES) cq[i] += fg[i] * ((fr[i] - meanr[i]) * invr[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sA[tid * V + i] = a[i];
    sB[tid * V + i] = bq[i];
    if (HAS_RES) sC[tid * V + i] = cq[i];
  }
  __syncthreads();
  for (int o = tid; o < cgb * V; o += BN_THREADS) {
    const int l = o / V, i = o % V;
    const int cgo = blockIdx.y * cgb + l;
    if (cgo >= ncg) continue;
    float da = 0.f, db = 0.f, dc = 0.f;
    for (int k = 0; k < nrl; ++k) {
      da += sA[(k * cgb + l) * V + i];
      db += sB[(k * cgb + l) * V + i];
      if (HAS_RES) dc += sC[(k * cgb + l) * V + i];
    }
    const int c = cgo * V + i;
    atomicAdd(&bn.bwd_sums[c], (double)da);
    atomicAdd(&bn.bwd_sums[C + c], (double)db);
    if (HAS_RES) {
      atomicAdd(&bnr.bwd_sums[c], (double)da);
      atomicAdd(&bnr.bwd_sums[C + c], (double)dc);
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward pass 2 (apply): dx = gamma*invstd*(g - sum_g/n - xhat*sum_gx/n)
// RES: 0 none; 1 identity residual -> g_ident = g; 2 residual BN -> dx_res
// same decomposition as bn_act_fwd_kernel: per-channel constants in registers.
// ---------------------------------------------------------------------------------------
template <typename T, bool HAS_G2, bool HAS_Y, int RES, bool SYNC>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const T* __restrict__ g1, const T* __restrict__ g2,
                                                                  const T* __restrict__ y, const T* __restrict__ x,
                                                                  const T* __restrict__ xr, T* __restrict__ dx,
                                                                  T* __restrict__ dxr, T* __restrict__ gid, ssb_bn bn,
                                                                  ssb_bn bnr, ssb_geom g, int cgpc, int rpb) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int C = g.C;
  const int ncg = C / V;
  const int rpp = BN_THREADS / cgpc;
  const int cgl = threadIdx.x % cgpc, rl = threadIdx.x / cgpc;
  const int cg = blockIdx.y * cgpc + cgl;
  const int rows = g.B * g.pitch;
  const int r0 = blockIdx.x * rpb;
  const int r1 = min(rows, r0 + rpb);
  // per channel: mean, invstd, k0 = gamma*invstd, k1 = sum_g/n, k2 = sum_gx/n (+ residual-BN mean, invstd, k0, k2);
  // one thread per channel does the fp64 part
  __shared__ float sCo[9][64];
  const int nch = cgpc * V;
  if (threadIdx.x < nch) {
    const int c = blockIdx.y * nch + threadIdx.x;
    if (c < C) {
      const double inv_n = 1.0 / ((double)g.B * (double)g.len * (double)(bn.count_mul > 1 ? bn.count_mul : 1));
      // SyncBN: the sums are global totals and the gradient arena is averaged over ranks afterwards,
      // so each rank contributes total / world
      const double gsc = 1.0 / (double)(bn.count_mul > 1 ? bn.count_mul : 1);
      const bool writer = blockIdx.x == 0;
      const float mean_ = bn.mean_invstd[c], inv_ = bn.mean_invstd[C + c];
      double sb[2] = {bn.bwd_sums[c], bn.bwd_sums[C + c]};
      if (SYNC && bn.sync_peers) {      // SyncBN: totals over the ranks, exchanged here
        const unsigned int ix[2] = {bn.sync_bwd_off + (unsigned int)c, bn.sync_bwd_off + (unsigned int)(C + c)};
        sbx_allsum<2>(bn, ix, sb, writer);
      }
      const double sg = sb[0], sgx = sb[1];
      sCo[0][threadIdx.x] = mean_;
      sCo[1][threadIdx.x] = inv_;
      sCo[2][threadIdx.x] = bn.gamma[c] * inv_;
      sCo[3][threadIdx.x] = (float)(sg * inv_n);
      sCo[4][threadIdx.x] = (float)(sgx * inv_n);
      if (writer) {
        bn.dgamma[c] = (float)(sgx * gsc);
        bn.dbeta[c] = (float)(sg * gsc);
      }
      if (RES == 2) {
        const float meanr = bnr.mean_invstd[c], invr = bnr.mean_invstd[C + c];
        double sr[1] = {bnr.bwd_sums[C + c]};
        if (SYNC && bnr.sync_peers) {
          const unsigned int ix[1] = {bnr.sync_bwd_off + (unsigned int)(C + c)};
          sbx_allsum<1>(bnr, ix, sr, writer);
        }
        const double sgxr = sr[0];
        sCo[5][threadIdx.x] = meanr;
        sCo[6][threadIdx.x] = invr;
        sCo[7][threadIdx.x] = bnr.gamma[c] * invr;
        sCo[8][threadIdx.x] = (float)(sgxr * inv_n);
        if (writer) {
          bnr.dgamma[c] = (float)(sgxr * gsc);
          bnr.dbeta[c] = (float)(sg * gsc);
        }
      }
    }
  }
  __syncthreads();
  if (rl >= rpp || cg >= ncg) return;
  float mean[V], inv[V], k0[V], k1[V], k2[V], rmean[V], rinv[V], rk0[V], rk2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = cgl * V + i;
    mean[i] = sCo[0][j]; inv[i] = sCo[1][j]; k0[i] = sCo[2][j]; k1[i] = sCo[3][j]; k2[i] = sCo[4][j];
    if (RES == 2) { rmean[i] = sCo[5][j]; rinv[i] = sCo[6][j]; rk0[i] = sCo[7][j]; rk2[i] = sCo[8][j]; }
  }
  for (int row = r0 + rl; row < r1; row += rpp) {
    const size_t off = (size_t)row * C + (size_t)cg * V;
    Vec<T> odx, odr, ogi;
    if (row_valid(row, g.pitch, g.len)) {
      Vec<T> vg, vx;
      vg.load(g1 + off);
      vx.load(x + off);
      float fg[V], fx[V];
      vg.get(fg);
      vx.get(fx);
      if (HAS_G2) {
        Vec<T> v2;
        v2.load(g2 + off);
        float f2[V];
        v2.get(f2);
#pragma unroll
        for (int i = 0; i < V; ++i) fg[i] += f2[i];
      }
      if (HAS_Y) {
        Vec<T> vy;
        vy.load(y + off);
        float fy[V];
        vy.get(fy);
#pragma unroll
        for (int i = 0; i < V; ++i) fg[i] = fy[i] > 0.f ? fg[i] : 0.f;
      }
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xh = (fx[i] - mean[i]) * inv[i];
        o[i] = k0[i] * (fg[i] - k1[i] - xh * k2[i]);
      }
      odx.set(o);
      if (RES == 1) ogi.set(fg);
      if (RES == 2) {
        Vec<T> vr;
        vr.load(xr + off);
        float fr[V];
        vr.get(fr);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xh = (fr[i] - rmean[i]) * rinv[i];
          o[i] = rk0[i] * (fg[i] - k1[i] - xh * rk2[i]);
        }
        odr.set(o);
      }
    } else {
      odx.zero();
      odr.zero();
      ogi.zero();
    }
    odx.store(dx + off);
    if (RES == 1) ogi.store(gid + off);
    if (RES == 2) odr.store(dxr + off);
  }
}

// ---------------------------------------------------------------------------------------
// backward, both passes in ONE launch: reduce -> grid-wide barrier -> apply.  The grid is sized to the kernel's
// co-resident capacity (host check with the occupancy API) and walks its rows twice (the second read hits L2 / L1);
// the barrier is a counter in the step's zero arena (one per BN layer, zeroed by the step's memset).
// RES: 0 none; 1 identity residual -> g_ident = g; 2 residual BN -> dx_res
// ---------------------------------------------------------------------------------------
#define BWF_ROWS 2
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < expected);
  }
  __syncthreads();
}

#ifndef BWF_KEEP
// Rows per thread whose (masked gradient, x[, x_res]) stay in registers across the barrier.  Measured on B200 at the
// config-2 shapes: 0 is fastest -- keeping 1 / 2 rows costs 16 / 28 registers, the occupancy-capped grid shrinks
// (444 -> 296 -> 148 blocks) and the step slows by 3 % / 5 %; forcing the register count with min-blocks spills.
// The second read comes from L2 anyway.
#define BWF_KEEP 0
#endif
#ifndef BWF_MINB
#define BWF_MINB 2     // <= 128 registers: two blocks per SM with the batched loads (forcing 64 registers -- 4 blocks per SM,
                       // 324-block grids -- was measured SLOWER in round 1: 0.710 vs 0.700 ms)
#endif
#define BWF_UR 4       // rows per trip of the row loops (all loads of a trip are in flight together)
#define BWF_REP 8      // replicated accumulators: block b adds into replica b % 8 (8x less same-address contention)

template <typename T, bool HAS_Y, int RES, bool SYNC>
__global__ void __launch_bounds__(BN_THREADS, BWF_MINB) bn_bwd_fused_kernel(const T* __restrict__ g1, const T* __restrict__ y,
                                                                  const T* __restrict__ x, const T* __restrict__ xr,
                                                                  T* __restrict__ dx, T* __restrict__ dxr, T* __restrict__ gid,
                                                                  ssb_bn bn, ssb_bn bnr, ssb_geom g, int cgpc, int rpb,
                                                                  unsigned int* __restrict__ barrier, double* __restrict__ rep,
                                                                  double* __restrict__ rep_r, long long rep_stride) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  constexpr int UR = RES == 2 ? 2 : BWF_UR;     // (three operand tensors x 4 rows do not fit 128 registers: ptxas spills)
  const int C = g.C;
  const int ncg = C / V;
  const int rpp = BN_THREADS / cgpc;
  const int cgl = threadIdx.x % cgpc, rl = threadIdx.x / cgpc;
  const int cg = blockIdx.y * cgpc + cgl;
  const bool owner = rl < rpp && cg < ncg;
  const int rows = g.B * g.pitch;
  const int r0 = blockIdx.x * rpb;
  const int r1 = min(rows, r0 + rpb);
  __shared__ float sA[BN_THREADS * V];
  __shared__ float sB[BN_THREADS * V];
  __shared__ float sC[RES == 2 ? BN_THREADS * V : 1];
  __shared__ float sCo[5][64];
  float mean[V], inv[V], meanr[V], invr[V];
  // ---- pass 1: reduce ----
  {
    float a[V], bq[V], cq[V];
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] = bq[i] = cq[i] = 0.f;
    if (owner) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = cg * V + i;
        mean[i] = bn.mean_invstd[c];
        inv[i] = bn.mean_invstd[C + c];
        if (RES == 2) {
          meanr[i] = bnr.mean_invstd[c];
          invr[i] = bnr.mean_invstd[C + c];
        }
      }
      // UR rows per trip, every load of the trip issued before the first use: the trace marks showed this pass at
      // ~5 us for 3 rows per thread -- one DRAM/L2 round trip per row, back to back (profiles/r2b_trace_config2.md)
      for (int base = r0 + rl; base < r1; base += UR * rpp) {
        Vec<T> vg[UR], vx[UR], vy[HAS_Y ? UR : 1], vr[RES == 2 ? UR : 1];
        bool ok[UR];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
          const int row = base + u * rpp;
          ok[u] = row < r1 && row_valid(row, g.pitch, g.len);       // halo / pad rows carry zero gradient
          if (ok[u]) {
            const size_t off = (size_t)row * C + (size_t)cg * V;
            vg[u].load(g1 + off);
            vx[u].load(x + off);
            if (HAS_Y) vy[u].load(y + off);
            if (RES == 2) vr[u].load(xr + off);
          }
        }
#pragma unroll
        for (int u = 0; u < UR; ++u) {
          if (!ok[u]) continue;
          float fg[V], fx[V], fr[V];
          vg[u].get(fg);
          vx[u].get(fx);
          if (HAS_Y) {
            float fy[V];
            vy[u].get(fy);
#pragma unroll
            for (int i = 0; i < V; ++i) fg[i] = fy[i] > 0.f ? fg[i] : 0.f;
          }
          if (RES == 2) vr[u].get(fr);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            a[i] += fg[i];
            bq[i] += fg[i] * ((fx[i] - mean[i]) * inv[i]);
            if (RES == 2) cq[i] += fg[i] * ((fr[i] - meanr[i]) * invr[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sA[threadIdx.x * V + i] = a[i];
      sB[threadIdx.x * V + i] = bq[i];
      if (RES == 2) sC[threadIdx.x * V + i] = cq[i];
    }
  }
  __syncthreads();
  const long long rsel = (long long)(blockIdx.x % BWF_REP) * rep_stride;
  for (int o = threadIdx.x; o < cgpc * V; o += BN_THREADS) {
    const int l = o / V, i = o % V;
    const int cgo = blockIdx.y * cgpc + l;
    if (cgo >= ncg) continue;
    float da = 0.f, db = 0.f, dc = 0.f;
    for (int k = 0; k < rpp; ++k) {
      da += sA[(k * cgpc + l) * V + i];
      db += sB[(k * cgpc + l) * V + i];
      if (RES == 2) dc += sC[(k * cgpc + l) * V + i];
    }
    const int c = cgo * V + i;
    atomicAdd(&rep[rsel + c], (double)da);
    atomicAdd(&rep[rsel + C + c], (double)db);
    if (RES == 2) atomicAdd(&rep_r[rsel + C + c], (double)dc);
    __threadfence();   // this thread's reductions are performed before the block signals the barrier
  }
  if (threadIdx.x == 0) SSB_MARK();                 // pass 1 + atomics done
  grid_barrier(barrier, gridDim.x * gridDim.y);
  if (threadIdx.x == 0) SSB_MARK();                 // past the grid barrier
  // ---- pass 2: apply; per-channel constants from the now complete sums (one thread per channel) ----
  // The operands of the first trip are requested BEFORE the constants are worked out: the second read (L2) then runs
  // under the reload of the sums, the fp64 arithmetic and -- with SyncBN -- the exchange of the sums with the peers.
  // (only where an exchange follows: at N = 1 the early requests made the kernel SLOWER, 8.05 -> 8.4 us in isolation)
  constexpr int PU = !SYNC ? 0 : (RES == 2 ? 0 : (HAS_Y ? 1 : 2));   // rows requested early (more does not fit 128 registers next to the exchange)
  constexpr int PA = PU > 0 ? PU : 1;
  Vec<T> pg[PA], px[PA], py[PA], pr[PA];
  bool pok[PA];
#pragma unroll
  for (int u = 0; u < PU; ++u) {
    const int row = r0 + rl + u * rpp;
    pok[u] = owner && row < r1 && row_valid(row, g.pitch, g.len);
    if (pok[u]) {
      const size_t off = (size_t)row * C + (size_t)cg * V;
      pg[u].load(g1 + off);
      px[u].load(x + off);
      if (HAS_Y) py[u].load(y + off);
      if (RES == 2) pr[u].load(xr + off);
    }
  }
  const int nch = cgpc * V;
  if (threadIdx.x < nch) {
    const int c = blockIdx.y * nch + threadIdx.x;
    if (c < C) {
      const double inv_n = 1.0 / ((double)g.B * (double)g.len * (double)(bn.count_mul > 1 ? bn.count_mul : 1));
      const double gsc = 1.0 / (double)(bn.count_mul > 1 ? bn.count_mul : 1);
      const bool writer = blockIdx.x == 0;
      double sg = 0.0, sgx = 0.0, sgxr = 0.0;
#pragma unroll
      for (int k = 0; k < BWF_REP; ++k) {
        sg += __ldcg(&rep[k * rep_stride + c]);
        sgx += __ldcg(&rep[k * rep_stride + C + c]);
        if (RES == 2) sgxr += __ldcg(&rep_r[k * rep_stride + C + c]);
      }
      if (SYNC && bn.sync_peers) {      // SyncBN: this rank's sums are complete (grid barrier above) -- totals over the ranks
        double sb[2] = {sg, sgx};
        const unsigned int ix[2] = {bn.sync_bwd_off + (unsigned int)c, bn.sync_bwd_off + (unsigned int)(C + c)};
        sbx_allsum<2>(bn, ix, sb, writer);
        sg = sb[0];
        sgx = sb[1];
        if (RES == 2) {
          double sr[1] = {sgxr};
          const unsigned int ixr[1] = {bnr.sync_bwd_off + (unsigned int)(C + c)};
          sbx_allsum<1>(bnr, ixr, sr, writer);
          sgxr = sr[0];
        }
      }
      sCo[0][threadIdx.x] = bn.gamma[c] * bn.mean_invstd[C + c];
      sCo[1][threadIdx.x] = (float)(sg * inv_n);
      sCo[2][threadIdx.x] = (float)(sgx * inv_n);
      if (writer) {
        bn.dgamma[c] = (float)(sgx * gsc);
        bn.dbeta[c] = (float)(sg * gsc);
        bn.bwd_sums[c] = sg;             // the totals, where the two-launch path leaves them
        bn.bwd_sums[C + c] = sgx;
      }
      if (RES == 2) {
        sCo[3][threadIdx.x] = bnr.gamma[c] * bnr.mean_invstd[C + c];
        sCo[4][threadIdx.x] = (float)(sgxr * inv_n);
        if (writer) {
          bnr.dgamma[c] = (float)(sgxr * gsc);
          bnr.dbeta[c] = (float)(sg * gsc);
          bnr.bwd_sums[c] = sg;
          bnr.bwd_sums[C + c] = sgxr;
        }
      }
    }
  }
  __syncthreads();
  if (!owner) return;
  float k0[V], k1[V], k2[V], rk0[V], rk2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = cgl * V + i;
    k0[i] = sCo[0][j]; k1[i] = sCo[1][j]; k2[i] = sCo[2][j];
    if (RES == 2) { rk0[i] = sCo[3][j]; rk2[i] = sCo[4][j]; }
  }
  for (int base = r0 + rl; base < r1; base += UR * rpp) {
    Vec<T> vg[UR], vx[UR], vy[HAS_Y ? UR : 1], vr[RES == 2 ? UR : 1];
    bool ok[UR];
    const bool first = base == r0 + rl;       // the trip whose first PU rows were requested before the constants
#pragma unroll
    for (int u = 0; u < UR; ++u) {            // second read (L2 / L1): again all loads of the trip first
      if (first && u < PU) {
        ok[u] = pok[u < PA ? u : 0];
        vg[u] = pg[u < PA ? u : 0];
        vx[u] = px[u < PA ? u : 0];
        if (HAS_Y) vy[u] = py[u < PA ? u : 0];
        if (RES == 2) vr[u] = pr[u < PA ? u : 0];
        continue;
      }
      const int row = base + u * rpp;
      ok[u] = row < r1 && row_valid(row, g.pitch, g.len);
      if (ok[u]) {
        const size_t off = (size_t)row * C + (size_t)cg * V;
        vg[u].load(g1 + off);
        vx[u].load(x + off);
        if (HAS_Y) vy[u].load(y + off);
        if (RES == 2) vr[u].load(xr + off);
      }
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int row = base + u * rpp;
      if (row >= r1) continue;
      const size_t off = (size_t)row * C + (size_t)cg * V;
      Vec<T> odx, odr, ogi;
      if (ok[u]) {
        float fg[V], fx[V], o[V];
        vg[u].get(fg);
        vx[u].get(fx);
        if (HAS_Y) {
          float fy[V];
          vy[u].get(fy);
#pragma unroll
          for (int i = 0; i < V; ++i) fg[i] = fy[i] > 0.f ? fg[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = k0[i] * (fg[i] - k1[i] - ((fx[i] - mean[i]) * inv[i]) * k2[i]);
        odx.set(o);
        if (RES == 1) ogi.set(fg);
        if (RES == 2) {
          float fr[V];
          vr[u].get(fr);
#pragma unroll
          for (int i = 0; i < V; ++i) o[i] = rk0[i] * (fg[i] - k1[i] - ((fr[i] - meanr[i]) * invr[i]) * rk2[i]);
          odr.set(o);
        }
      } else {
        odx.zero();
        odr.zero();
        ogi.zero();
      }
      odx.store(dx + off);
      if (RES == 1) ogi.store(gid + off);
      if (RES == 2) odr.store(dxr + off);
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward, both passes in ONE launch and ONE read of the operands: thread-block CLUSTERS + distributed shared memory.
// The in-kernel timeline of the config-2 step (profiles/r2_trace_config2.md) shows the grid-barrier kernel above at
// 10-14 us per BatchNorm layer on the critical path of the backward (17 layers): load -> shared reduce -> fp64 atomics
// -> fence -> grid barrier -> reload of the sums -> second read of the operands -> store, every arrow a round trip to
// L2.  Here a cluster of CL blocks owns ALL rows of a chunk of <= 64 channels: every thread keeps its <= NR rows of
// (masked gradient, x[, x_res]) in REGISTERS, the block partials meet in shared memory, the CL blocks read one another's
// partials through DSMEM after one hardware cluster barrier (no global atomics, no fence, no spin), and the apply pass
// runs from the registers.  Summation order is fixed (warp tree -> 8 warps -> CL ranks in rank order): deterministic.
// Fits when rows <= CL * NR * (256 / vector groups per chunk); larger tensors keep the kernels above.
// MEASURED (round 2, config-2 step, B200): 0.689 ms per step with the grid-barrier kernel, 0.739 with clusters of 8,
// 0.772 with clusters of 16 -- in-kernel marks: 2-4 us until a block has its operands and partials, ~0.9 us cluster
// barrier, ~3 us DSMEM gather + finalise + stores.  The chain of dependent memory round trips is as long as before and
// cluster launches schedule later.  Kept behind SSB_BN_CLUSTER=8|16 (tested), OFF by default.
// RES: 0 none; 1 identity residual -> g_ident = g; 2 residual BN -> dx_res
// ---------------------------------------------------------------------------------------
template <typename T, bool HAS_Y, int RES, int NR>
__global__ void __launch_bounds__(BN_THREADS, 1) bn_bwd_cluster_kernel(const T* __restrict__ g1, const T* __restrict__ y,
                                                                      const T* __restrict__ x, const T* __restrict__ xr,
                                                                      T* __restrict__ dx, T* __restrict__ dxr, T* __restrict__ gid,
                                                                      ssb_bn bn, ssb_bn bnr, ssb_geom g, int cgpc, int rows_per_cta) {
  namespace cgs = cooperative_groups;
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  constexpr int NQ = RES == 2 ? 3 : 2;
  cgs::cluster_group cluster = cgs::this_cluster();
  const unsigned CL = cluster.num_blocks(), rank = cluster.block_rank();
  const int C = g.C;
  const int chunk = blockIdx.x / CL;
  const int cpc = cgpc * V;                       // channels of this chunk (<= 64)
  const int rpp = BN_THREADS / cgpc;              // row lanes
  const int cgl = threadIdx.x % cgpc, rl = threadIdx.x / cgpc;
  const int cg = chunk * cgpc + cgl;              // 16-byte channel group
  const int rows = g.B * g.pitch;
  const int r0 = (int)rank * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  __shared__ float s_warp[BN_THREADS / 32][NQ][64];
  __shared__ float s_part[NQ][64];                // this block's partial sums: read by the whole cluster
  __shared__ float s_co[5][64];
  float mean[V], inv[V], meanr[V], invr[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cg * V + i;
    mean[i] = bn.mean_invstd[c];
    inv[i] = bn.mean_invstd[C + c];
    if (RES == 2) {
      meanr[i] = bnr.mean_invstd[c];
      invr[i] = bnr.mean_invstd[C + c];
    }
  }
  // ---- one read of the operands: all loads of a thread's rows are issued before the first use ----
  Vec<T> kg[NR], kx[NR], kr[RES == 2 ? NR : 1], ky[HAS_Y ? NR : 1];
  bool ok[NR];
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    const int row = r0 + rl + k * rpp;
    ok[k] = row < r1 && row_valid(row, g.pitch, g.len);
    if (ok[k]) {
      const size_t off = (size_t)row * C + (size_t)cg * V;
      kg[k].load(g1 + off);
      kx[k].load(x + off);
      if (HAS_Y) ky[k].load(y + off);
      if (RES == 2) kr[k].load(xr + off);
    }
  }
  float a[V], bq[V], cq[V];
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = bq[i] = cq[i] = 0.f;
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    if (!ok[k]) continue;
    float fg[V], fx[V];
    kg[k].get(fg);
    kx[k].get(fx);
    if (HAS_Y) {
      float fy[V];
      ky[k].get(fy);
#pragma unroll
      for (int i = 0; i < V; ++i) fg[i] = fy[i] > 0.f ? fg[i] : 0.f;
      kg[k].set(fg);            // exact: masking keeps the stored values or zero
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      a[i] += fg[i];
      bq[i] += fg[i] * ((fx[i] - mean[i]) * inv[i]);
    }
    if (RES == 2) {
      float fr[V];
      kr[k].get(fr);
#pragma unroll
      for (int i = 0; i < V; ++i) cq[i] += fg[i] * ((fr[i] - meanr[i]) * invr[i]);
    }
  }
  // ---- block partials: lanes of a warp that share a channel group (lane % cgpc) are summed by xor shuffles ----
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    if (o >= cgpc) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
        bq[i] += __shfl_xor_sync(0xffffffffu, bq[i], o);
        if (RES == 2) cq[i] += __shfl_xor_sync(0xffffffffu, cq[i], o);
      }
    }
  }
  if (lane < cgpc) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s_warp[warp][0][lane * V + i] = a[i];
      s_warp[warp][1][lane * V + i] = bq[i];
      if (RES == 2) s_warp[warp][2][lane * V + i] = cq[i];
    }
  }
  __syncthreads();
  float pre_gi = 0.f;       // gamma * invstd of the channel this thread finalises (loaded before the barrier)
  if ((int)threadIdx.x < NQ * cpc) {
    const int q = threadIdx.x / cpc, ch = threadIdx.x - q * cpc;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < BN_THREADS / 32; ++w) t += s_warp[w][q][ch];
    s_part[q][ch] = t;
    const int c = chunk * cpc + ch;
    if (q == 0) pre_gi = bn.gamma[c] * bn.mean_invstd[C + c];
    if (RES == 2 && q == 2) pre_gi = bnr.gamma[c] * bnr.mean_invstd[C + c];
  }
  if (threadIdx.x == 0) SSB_MARK();                 // operands read, block partials done
  cluster.sync();                                   // every block's partials are in its shared memory
  if (threadIdx.x == 0) SSB_MARK();                 // past the cluster barrier
  // ---- totals over the cluster (rank order, fp64), per-channel constants; rank 0 publishes the parameter gradients ----
  double tot = 0.0;
  if ((int)threadIdx.x < NQ * cpc) {
    const int q = threadIdx.x / cpc, ch = threadIdx.x - q * cpc;
    float pv[16];
#pragma unroll
    for (unsigned r = 0; r < 16; ++r) pv[r] = r < CL ? cluster.map_shared_rank(&s_part[0][0], r)[q * 64 + ch] : 0.f;   // loads in flight together
#pragma unroll
    for (unsigned r = 0; r < 16; ++r) tot += (double)pv[r];
  }
  cgs::cluster_group::arrival_token token = cluster.barrier_arrive();   // done reading the peers' shared memory
  {
    const double inv_n = 1.0 / ((double)g.B * (double)g.len * (double)(bn.count_mul > 1 ? bn.count_mul : 1));
    const double gsc = 1.0 / (double)(bn.count_mul > 1 ? bn.count_mul : 1);
    if ((int)threadIdx.x < NQ * cpc) {
      const int q = threadIdx.x / cpc, ch = threadIdx.x - q * cpc;
      const int c = chunk * cpc + ch;
      if (q == 0) {
        s_co[0][ch] = pre_gi;
        s_co[1][ch] = (float)(tot * inv_n);
        if (rank == 0) {
          bn.dbeta[c] = (float)(tot * gsc);
          bn.bwd_sums[c] = tot;
          if (RES == 2) {
            bnr.dbeta[c] = (float)(tot * gsc);
            bnr.bwd_sums[c] = tot;
          }
        }
      } else if (q == 1) {
        s_co[2][ch] = (float)(tot * inv_n);
        if (rank == 0) {
          bn.dgamma[c] = (float)(tot * gsc);
          bn.bwd_sums[C + c] = tot;
        }
      } else {
        s_co[3][ch] = pre_gi;
        s_co[4][ch] = (float)(tot * inv_n);
        if (rank == 0) {
          bnr.dgamma[c] = (float)(tot * gsc);
          bnr.bwd_sums[C + c] = tot;
        }
      }
    }
  }
  __syncthreads();
  float k0[V], k1[V], k2[V], rk0[V], rk2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = cgl * V + i;
    k0[i] = s_co[0][j]; k1[i] = s_co[1][j]; k2[i] = s_co[2][j];
    if (RES == 2) { rk0[i] = s_co[3][j]; rk2[i] = s_co[4][j]; }
  }
  // ---- apply from the registers ----
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    const int row = r0 + rl + k * rpp;
    if (row >= r1) continue;
    const size_t off = (size_t)row * C + (size_t)cg * V;
    Vec<T> odx, odr, ogi;
    if (ok[k]) {
      float fg[V], fx[V], o[V];
      kg[k].get(fg);
      kx[k].get(fx);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = k0[i] * (fg[i] - k1[i] - ((fx[i] - mean[i]) * inv[i]) * k2[i]);
      odx.set(o);
      if (RES == 1) ogi = kg[k];
      if (RES == 2) {
        float fr[V];
        kr[k].get(fr);
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = rk0[i] * (fg[i] - k1[i] - ((fr[i] - meanr[i]) * invr[i]) * rk2[i]);
        odr.set(o);
      }
    } else {
      odx.zero();
      odr.zero();
      ogi.zero();
    }
    odx.store(dx + off);
    if (RES == 1) ogi.store(gid + off);
    if (RES == 2) odr.store(dxr + off);
  }
  if (threadIdx.x == 0) SSB_MARK();                 // applied and stored
  cluster.barrier_wait(std::move(token));           // no block leaves while a peer may still read its partials
}

// ---------------------------------------------------------------------------------------
// stem tail backward: the pooled gradient is routed through the max-pool by the slot index the
// forward saved (slot 3 = ReLU-dead), then the two BN-backward passes on c0.
// ---------------------------------------------------------------------------------------
template <typename T, int V>
__device__ __forceinline__ void stem_routed_grad(const T* __restrict__ gp, const uint8_t* __restrict__ arg,
                                                 const ssb_geom& gi, const ssb_geom& go, int b, int l, int cg, float* g) {
  const int C = gi.C;
#pragma unroll
  for (int i = 0; i < V; ++i) g[i] = 0.f;
  // windows containing l (centre 2t): l even -> t = l/2, slot 1; l odd -> t = (l-1)/2 slot 2 and t = (l+1)/2 slot 0
  int tw[2], slot[2], nw;
  if ((l & 1) == 0) { tw[0] = l >> 1; slot[0] = 1; nw = 1; }
  else { tw[0] = (l - 1) >> 1; slot[0] = 2; tw[1] = (l + 1) >> 1; slot[1] = 0; nw = 2; }
  for (int w = 0; w < nw; ++w) {
    if (tw[w] >= go.len) continue;
    const size_t off = ((size_t)b * go.pitch + 1 + tw[w]) * C + (size_t)cg * V;
    Vec<T> vg;
    vg.load(gp + off);
    float fg[V];
    vg.get(fg);
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const uint32_t a4 = *reinterpret_cast<const uint32_t*>(arg + off + i);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (((a4 >> (8 * u)) & 0xffu) == (uint32_t)slot[w]) g[i + u] += fg[i + u];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS) stem_bwd_reduce_kernel(const T* __restrict__ gp, const T* __restrict__ c0,
                                                                     const uint8_t* __restrict__ arg, ssb_bn bn,
                                                                     ssb_geom gi, ssb_geom go, int rpb) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int C = gi.C;
  const int ncg = C / V;
  const int cgb = ncg < BN_THREADS ? ncg : BN_THREADS;
  const int nrl = BN_THREADS / cgb;
  const int tid = threadIdx.x;
  const int cgl = tid % cgb, rl = tid / cgb;
  const int cg = blockIdx.y * cgb + cgl;
  __shared__ float sA[BN_THREADS * V];
  __shared__ float sB[BN_THREADS * V];
  float a[V], bq[V];
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = bq[i] = 0.f;
  if (rl < nrl && cg < ncg) {
    float mean[V], inv[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      mean[i] = bn.mean_invstd[cg * V + i];
      inv[i] = bn.mean_invstd[C + cg * V + i];
    }
    const int rows = gi.B * gi.pitch;
    const int r0 = blockIdx.x * rpb;
    const int r1 = min(rows, r0 + rpb);
    for (int r = r0 + rl; r < r1; r += nrl) {
      const int b = r / gi.pitch, pos = r - b * gi.pitch;
      if (pos < 1 || pos > gi.len) continue;
      float g[V];
      stem_routed_grad<T, V>(gp, arg, gi, go, b, pos - 1, cg, g);
      Vec<T> vx;
      vx.load(c0 + (size_t)r * C + (size_t)cg * V);
      float fx[V];
      vx.get(fx);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        a[i] += g[i];
        bq[i] += g[i] * ((fx[i] - mean[i]) * inv[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sA[tid * V + i] = a[i];
    sB[tid * V + i] = bq[i];
  }
  __syncthreads();
  for (int o = tid; o < cgb * V; o += BN_THREADS) {
    const int l = o / V, i = o % V;
    const int cgo = blockIdx.y * cgb + l;
    if (cgo >= ncg) continue;
    float da = 0.f, db = 0.f;
    for (int k = 0; k < nrl; ++k) {
      da += sA[(k * cgb + l) * V + i];
      db += sB[(k * cgb + l) * V + i];
    }
    const int c = cgo * V + i;
    atomicAdd(&bn.bwd_sums[c], (double)da);
    atomicAdd(&bn.bwd_sums[C + c], (double)db);
  }
}

template <typename T, bool SYNC>
__global__ void __launch_bounds__(BN_THREADS) stem_bwd_apply_kernel(const T* __restrict__ gp, const T* __restrict__ c0,
                                                                    const uint8_t* __restrict__ arg, T* __restrict__ dc0,
                                                                    ssb_bn bn, ssb_geom gi, ssb_geom go) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  extern __shared__ float sm[];
  const int C = gi.C;
  float* sScale = sm;
  float* sMean = sm + C;
  float* sInv = sm + 2 * C;
  float* sK1 = sm + 3 * C;
  float* sK2 = sm + 4 * C;
  const double inv_n = 1.0 / ((double)gi.B * (double)gi.len * (double)(bn.count_mul > 1 ? bn.count_mul : 1));
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = bn.mean_invstd[c], inv = bn.mean_invstd[C + c];
    sMean[c] = mean;
    sInv[c] = inv;
    sScale[c] = bn.gamma[c] * inv;
    double sb[2] = {bn.bwd_sums[c], bn.bwd_sums[C + c]};
    if (SYNC && bn.sync_peers) {      // SyncBN: totals over the ranks, exchanged here
      const unsigned int ix[2] = {bn.sync_bwd_off + (unsigned int)c, bn.sync_bwd_off + (unsigned int)(C + c)};
      sbx_allsum<2>(bn, ix, sb, blockIdx.x == 0);
    }
    sK1[c] = (float)(sb[0] * inv_n);
    sK2[c] = (float)(sb[1] * inv_n);
    if (blockIdx.x == 0) {
      const double gsc = 1.0 / (double)(bn.count_mul > 1 ? bn.count_mul : 1);
      bn.dgamma[c] = (float)(sb[1] * gsc);
      bn.dbeta[c] = (float)(sb[0] * gsc);
    }
  }
  __syncthreads();
  const int ncg = C / V;
  const long long total = (long long)gi.B * gi.pitch * ncg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / ncg);
    const int cg = (int)(idx - (long long)row * ncg);
    const int b = row / gi.pitch, pos = row - b * gi.pitch;
    Vec<T> out;
    if (pos >= 1 && pos <= gi.len) {
      float g[V];
      stem_routed_grad<T, V>(gp, arg, gi, go, b, pos - 1, cg, g);
      Vec<T> vx;
      vx.load(c0 + (size_t)row * C + (size_t)cg * V);
      float fx[V];
      vx.get(fx);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = cg * V + i;
        const float xh = (fx[i] - sMean[c]) * sInv[c];
        g[i] = sScale[c] * (g[i] - sK1[c] - xh * sK2[c]);
      }
      out.set(g);
    } else {
      out.zero();
    }
    out.store(dc0 + (size_t)row * C + (size_t)cg * V);
  }
}

// ---------------------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------------------
static int check_geom(const char* who, const ssb_geom& g, int vec) {
  SSB_REQUIRE(g.B > 0 && g.len > 0 && g.C > 0, "%s: empty geometry (B=%d len=%d C=%d)", who, g.B, g.len, g.C);
  SSB_REQUIRE(g.pitch >= g.len + 2, "%s: pitch %d < len %d + 2", who, g.pitch, g.len);
  SSB_REQUIRE(g.C % 8 == 0, "%s: C=%d must be a multiple of 8", who, g.C);
  SSB_REQUIRE((long long)g.B * g.pitch < (1ll << 31), "%s: too many rows", who);
  (void)vec;
  return SSB_OK;
}

static int ew_blocks(long long total_vec) {
  long long b = ceil_div_ll(total_vec, (long long)BN_THREADS * 2);
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

// 2-D launch geometry of the per-row elementwise BN kernels: channel chunks of <= 64 channels
// (cgpc 16-byte groups), rpb rows per block chosen so that the grid has a few blocks per SM
struct EwGeom {
  int cgpc, rpb;
  dim3 grid;
};
static EwGeom ew_geom(int rows, int ncg, int vec) {
  EwGeom e;
  const int per_chunk = 64 / vec;
  e.cgpc = ncg < per_chunk ? ncg : per_chunk;
  const int ny = ceil_div(ncg, e.cgpc);
  const int rpp = BN_THREADS / e.cgpc;
  int nx = ceil_div(148 * 6, ny);
  int rpb = ceil_div(rows, nx);
  rpb = ceil_div(rpb, rpp) * rpp;
  if (rpb < rpp) rpb = rpp;
  e.rpb = rpb;
  e.grid = dim3(ceil_div(rows, rpb), ny);
  return e;
}

static const ssb_bn kNoBn = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};

// grid of the fused backward kernel and whether all of it can be resident at once (the blocks meet at a grid-wide
// barrier): bounded by the measured occupancy of the instantiation, one block per SM left as slack
template <typename T>
static bool bwd_fused_plan(const ssb_geom& g, int mode, bool has_y, int* cgpc_o, int* rpb_o, dim3* grid_o) {
  constexpr int V = Vec<T>::N;
  const int rows = g.B * g.pitch;
  const int ncg = g.C / V;
  const int per_chunk = 64 / V;
  const int cgpc = ncg < per_chunk ? ncg : per_chunk;
  const int ny = ceil_div(ncg, cgpc);
  const int rpp = BN_THREADS / cgpc;
  int occ = 0;
  cudaError_t oe = cudaSuccess;
  if (mode == 0) oe = has_y ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_kernel<T, true, 0, false>, BN_THREADS, 0)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_kernel<T, false, 0, false>, BN_THREADS, 0);
  else if (mode == 1) oe = has_y ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_kernel<T, true, 1, false>, BN_THREADS, 0)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_kernel<T, false, 1, false>, BN_THREADS, 0);
  else oe = has_y ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_kernel<T, true, 2, false>, BN_THREADS, 0)
                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_kernel<T, false, 2, false>, BN_THREADS, 0);
  if (oe != cudaSuccess) occ = 0;
  // co-resident capacity with one block per SM left as slack; the rows are spread over at most that many blocks
  // (measured at config 2: 216 blocks -- this cap at 3 blocks per SM -- 0.700 ms per step; 148 blocks 0.716; 324 blocks
  // at a forced 64 registers 0.710)
  static const int slack = getenv("SSB_BWF_SLACK") ? atoi(getenv("SSB_BWF_SLACK")) : 1;
  const long long cap = occ - slack >= 1 ? (long long)g_ssb_num_sms * (occ - slack) : 0;
  int nx = cap / ny > 0 ? (int)(cap / ny) : 0;
  const int want = ceil_div(rows, rpp * BWF_ROWS);
  if (nx > want) nx = want;
  int rpb = nx > 0 ? ceil_div(ceil_div(rows, nx), rpp) * rpp : rows;
  if (nx > 0) nx = ceil_div(rows, rpb);
  *cgpc_o = cgpc;
  *rpb_o = rpb;
  *grid_o = dim3(nx, ny);
  return nx > 0 && (long long)nx * ny <= cap;
}

// ---- cluster / DSMEM backward (bn_bwd_cluster_kernel) ----
// largest cluster size to use: 16 (non-portable, opt-in), 8, or 0 = off; env SSB_BN_CLUSTER (read per call: launches are
// captured once, and the tests switch it)
static int bn_cluster_max() {
  const char* e = getenv("SSB_BN_CLUSTER");
  const int v = e ? atoi(e) : 0;      // measured SLOWER than the grid-barrier kernel: off by default
  return (v == 0 || v == 8 || v == 16) ? v : 0;
}
struct ClusterPlan {
  int cgpc, rows_per_cta, cl, nr, nchunks;
};
template <typename T>
static bool bwd_cluster_plan(const ssb_geom& g, int mode, ClusterPlan* p) {
  constexpr int V = Vec<T>::N;
  const int cmax = bn_cluster_max();
  if (cmax == 0) return false;
  const int ncg = g.C / V;
  int cgpc = 1;
  while (cgpc * 2 <= ncg / 8 && cgpc * 2 <= 64 / V) cgpc *= 2;      // aim at >= 8 channel chunks, <= 64 channels each
  while (ncg % cgpc) cgpc >>= 1;
  const int rpp = BN_THREADS / cgpc;
  const int rows = g.B * g.pitch;
  for (int cl = cmax; cl >= 8; cl >>= 1) {
    const int rpc = ceil_div(rows, cl);
    const int need = ceil_div(rpc, rpp);
    // (12 register-resident rows of three bf16 operands -- the residual-BN mode -- do not fit 255 registers: ptxas spills)
    if (need <= 6 || (need <= 12 && !(mode == 2 && V == 8))) {
      p->cgpc = cgpc; p->rows_per_cta = rpc; p->cl = cl; p->nr = need <= 6 ? 6 : 12; p->nchunks = ncg / cgpc;
      return true;
    }
  }
  return false;
}
template <typename Kern, typename... Args>
static cudaError_t launch_cluster(Kern kern, int nblocks, int cl, cudaStream_t st, Args... args) {
  // clusters of 16 are a per-kernel opt-in.  (All instantiations share one function-pointer TYPE, so a flag that is
  // static in this template would be shared by them: set the attribute on every launch -- launches are captured once.)
  if (cl > 8) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)nblocks);
  cfg.blockDim = dim3(BN_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_ssb_pdl == 1 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
#define SSB_BWC(Y, R, N)                                                                                                   \
  launch_cluster(bn_bwd_cluster_kernel<T, Y, R, N>, cp.nchunks * cp.cl, cp.cl, st, (const T*)g1, (const T*)y, (const T*)x, \
                 (const T*)x_res, (T*)dx, (T*)dx_res, (T*)g_ident, *bn, br, g, cp.cgpc, cp.rows_per_cta)
#define SSB_BWC_NR(Y, R) (cp.nr == 6 ? SSB_BWC(Y, R, 6) : SSB_BWC(Y, R, 12))

#define SSB_BWF_S(Y, R, S)                                                                                                     \
  ssb_launch(bn_bwd_fused_kernel<T, Y, R, S>, dim3(grid), dim3(BN_THREADS), 0, st, (const T*)g1, (const T*)y, (const T*)x,    \
             (const T*)x_res, (T*)dx, (T*)dx_res, (T*)g_ident, *bn, br, g, cgpc, rpb, barrier, rep, rep_r, (long long)rep_stride)
#define SSB_BWF(Y, R)                                    \
  do {                                                   \
    if (bn->sync_peers) SSB_BWF_S(Y, R, true);           \
    else SSB_BWF_S(Y, R, false);                         \
  } while (0)
#define SSB_RED(G2, Y, R) \
  ssb_launch(bn_bwd_reduce_kernel<T, G2, Y, R>, dim3(grid), dim3(BN_THREADS), 0, st, (const T*)g1, (const T*)g2, (const T*)y, (const T*)x, (const T*)x_res, *bn, br, rows, g.C, rpb)
#define SSB_APP_S(G2, Y, R, S) \
  ssb_launch(bn_bwd_apply_kernel<T, G2, Y, R, S>, dim3(eg.grid), dim3(BN_THREADS), 0, st, (const T*)g1, (const T*)g2, (const T*)y, (const T*)x, (const T*)x_res, (T*)dx, (T*)dx_res, (T*)g_ident, *bn, br, g, eg.cgpc, eg.rpb)
#define SSB_APP(G2, Y, R)                                  \
  do {                                                     \
    if (bn->sync_peers) SSB_APP_S(G2, Y, R, true);         \
    else SSB_APP_S(G2, Y, R, false);                       \
  } while (0)

extern "C" {

int ssb_bn_stats(const void* x, ssb_geom g, double* sums, int dtype, ssb_stream_t stream) {
  int rc = check_geom("ssb_bn_stats", g, 0);
  if (rc) return rc;
  SSB_REQUIRE(x && sums, "ssb_bn_stats: null pointer");
  const int rows = g.B * g.pitch;
  SSB_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    const int ncg = g.C / V;
    const int cgb = ncg < BN_THREADS ? ncg : BN_THREADS;
    const int nrl = BN_THREADS / cgb;
    int rpb = ceil_div(rows, 148 * 4);
    if (rpb < nrl * 4) rpb = nrl * 4;
    dim3 grid(ceil_div(rows, rpb), ceil_div(ncg, cgb));
    ssb_launch(bn_stats_kernel<T>, dim3(grid), dim3(BN_THREADS), 0, to_stream(stream), (const T*)x, rows, g.C, sums, rpb);
  })
  SSB_LAUNCH_CHECK("ssb_bn_stats");
  return SSB_OK;
}

int ssb_bn_act_fwd(const void* x, const ssb_bn* bn, const void* res, const ssb_bn* bn_res, void* y, ssb_geom g,
                   int relu, int train, int dtype, ssb_stream_t stream) {
  int rc = check_geom("ssb_bn_act_fwd", g, 0);
  if (rc) return rc;
  SSB_REQUIRE(x && bn && y, "ssb_bn_act_fwd: null pointer");
  SSB_REQUIRE(!(bn_res && !res), "ssb_bn_act_fwd: bn_res given without res");
  const int mode = res ? (bn_res ? 2 : 1) : 0;
  SSB_DISPATCH_DTYPE(dtype, T, {
    const EwGeom eg = ew_geom(g.B * g.pitch, g.C / Vec<T>::N, Vec<T>::N);
    cudaStream_t st = to_stream(stream);
    if (mode == 0)
      if (bn->sync_peers) ssb_launch(bn_act_fwd_kernel<T, 0, true>, dim3(eg.grid), dim3(BN_THREADS), 0, st, (const T*)x, nullptr, (T*)y, *bn, kNoBn, g, relu, train, eg.cgpc, eg.rpb);
      else ssb_launch(bn_act_fwd_kernel<T, 0, false>, dim3(eg.grid), dim3(BN_THREADS), 0, st, (const T*)x, nullptr, (T*)y, *bn, kNoBn, g, relu, train, eg.cgpc, eg.rpb);
    else if (mode == 1)
      if (bn->sync_peers) ssb_launch(bn_act_fwd_kernel<T, 1, true>, dim3(eg.grid), dim3(BN_THREADS), 0, st, (const T*)x, (const T*)res, (T*)y, *bn, kNoBn, g, relu, train, eg.cgpc, eg.rpb);
      else ssb_launch(bn_act_fwd_kernel<T, 1, false>, dim3(eg.grid), dim3(BN_THREADS), 0, st, (const T*)x, (const T*)res, (T*)y, *bn, kNoBn, g, relu, train, eg.cgpc, eg.rpb);
    else
      if (bn->sync_peers) ssb_launch(bn_act_fwd_kernel<T, 2, true>, dim3(eg.grid), dim3(BN_THREADS), 0, st, (const T*)x, (const T*)res, (T*)y, *bn, *bn_res, g, relu, train, eg.cgpc, eg.rpb);
      else ssb_launch(bn_act_fwd_kernel<T, 2, false>, dim3(eg.grid), dim3(BN_THREADS), 0, st, (const T*)x, (const T*)res, (T*)y, *bn, *bn_res, g, relu, train, eg.cgpc, eg.rpb);
  })
  SSB_LAUNCH_CHECK("ssb_bn_act_fwd");
  return SSB_OK;
}

int ssb_stem_bn_relu_pool_fwd(const void* c0, const ssb_bn* bn, void* y, uint8_t* arg, ssb_geom gin, ssb_geom gout,
                              int train, int dtype, ssb_stream_t stream) {
  int rc = check_geom("ssb_stem_bn_relu_pool_fwd(in)", gin, 0);
  if (rc) return rc;
  rc = check_geom("ssb_stem_bn_relu_pool_fwd(out)", gout, 0);
  if (rc) return rc;
  SSB_REQUIRE(c0 && bn && y, "ssb_stem_bn_relu_pool_fwd: null pointer");
  SSB_REQUIRE(gin.C == gout.C && gin.B == gout.B, "ssb_stem_bn_relu_pool_fwd: geometry mismatch");
  SSB_REQUIRE(gout.len == (gin.len - 1) / 2 + 1, "ssb_stem_bn_relu_pool_fwd: len_out %d != pool(len_in %d)", gout.len, gin.len);
  const size_t smem = (size_t)2 * gin.C * sizeof(float);
  SSB_DISPATCH_DTYPE(dtype, T, {
    const long long total = (long long)gout.B * gout.pitch * (gout.C / Vec<T>::N);
    if (bn->sync_peers) ssb_launch(stem_bn_relu_pool_kernel<T, true>, dim3(ew_blocks(total)), dim3(BN_THREADS), smem, to_stream(stream), (const T*)c0, (T*)y, arg, *bn, gin, gout, train);
    else ssb_launch(stem_bn_relu_pool_kernel<T, false>, dim3(ew_blocks(total)), dim3(BN_THREADS), smem, to_stream(stream), (const T*)c0, (T*)y, arg, *bn, gin, gout, train);
  })
  SSB_LAUNCH_CHECK("ssb_stem_bn_relu_pool_fwd");
  return SSB_OK;
}

int ssb_bn_bwd_reduce(const void* g1, const void* g2, const void* y, const void* x, const ssb_bn* bn,
                      const void* x_res, const ssb_bn* bn_res, ssb_geom g, int dtype, ssb_stream_t stream) {
  int rc = check_geom("ssb_bn_bwd_reduce", g, 0);
  if (rc) return rc;
  SSB_REQUIRE(g1 && x && bn, "ssb_bn_bwd_reduce: null pointer");
  SSB_REQUIRE((x_res != nullptr) == (bn_res != nullptr), "ssb_bn_bwd_reduce: x_res and bn_res go together");
  const int rows = g.B * g.pitch;
  SSB_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    const int ncg = g.C / V;
    const int cgb = ncg < BN_THREADS ? ncg : BN_THREADS;
    const int nrl = BN_THREADS / cgb;
    int rpb = ceil_div(rows, 148 * 4);
    if (rpb < nrl * 4) rpb = nrl * 4;
    dim3 grid(ceil_div(rows, rpb), ceil_div(ncg, cgb));
    cudaStream_t st = to_stream(stream);
    const ssb_bn br = bn_res ? *bn_res : kNoBn;
    const int sel = (g2 ? 4 : 0) | (y ? 2 : 0) | (x_res ? 1 : 0);
    switch (sel) {
      case 0: SSB_RED(false, false, false); break;
      case 1: SSB_RED(false, false, true); break;
      case 2: SSB_RED(false, true, false); break;
      case 3: SSB_RED(false, true, true); break;
      case 4: SSB_RED(true, false, false); break;
      case 5: SSB_RED(true, false, true); break;
      case 6: SSB_RED(true, true, false); break;
      default: SSB_RED(true, true, true); break;
    }
  })
  SSB_LAUNCH_CHECK("ssb_bn_bwd_reduce");
  return SSB_OK;
}

int ssb_bn_bwd_apply(const void* g1, const void* g2, const void* y, const void* x, const ssb_bn* bn, void* dx,
                     const void* x_res, const ssb_bn* bn_res, void* dx_res, void* g_ident, ssb_geom g, int dtype,
                     ssb_stream_t stream) {
  int rc = check_geom("ssb_bn_bwd_apply", g, 0);
  if (rc) return rc;
  SSB_REQUIRE(g1 && x && bn && dx, "ssb_bn_bwd_apply: null pointer");
  SSB_REQUIRE(!(bn_res && (!x_res || !dx_res)), "ssb_bn_bwd_apply: residual BN needs x_res and dx_res");
  SSB_REQUIRE(!(bn_res && g_ident), "ssb_bn_bwd_apply: g_ident and bn_res are exclusive");
  const int mode = bn_res ? 2 : (g_ident ? 1 : 0);
  SSB_DISPATCH_DTYPE(dtype, T, {
    const EwGeom eg = ew_geom(g.B * g.pitch, g.C / Vec<T>::N, Vec<T>::N);
    cudaStream_t st = to_stream(stream);
    const ssb_bn br = bn_res ? *bn_res : kNoBn;
    const int sel = (g2 ? 2 : 0) | (y ? 1 : 0);
    if (mode == 0) {
      switch (sel) { case 0: SSB_APP(false, false, 0); break; case 1: SSB_APP(false, true, 0); break;
                     case 2: SSB_APP(true, false, 0); break; default: SSB_APP(true, true, 0); break; }
    } else if (mode == 1) {
      switch (sel) { case 0: SSB_APP(false, false, 1); break; case 1: SSB_APP(false, true, 1); break;
                     case 2: SSB_APP(true, false, 1); break; default: SSB_APP(true, true, 1); break; }
    } else {
      switch (sel) { case 0: SSB_APP(false, false, 2); break; case 1: SSB_APP(false, true, 2); break;
                     case 2: SSB_APP(true, false, 2); break; default: SSB_APP(true, true, 2); break; }
    }
  })
  SSB_LAUNCH_CHECK("ssb_bn_bwd_apply");
  return SSB_OK;
}

int ssb_bn_bwd_fused_fits(ssb_geom g, int res_mode, int has_y, int dtype) {
  if (check_geom("ssb_bn_bwd_fused_fits", g, 0)) return 0;
  int cgpc, rpb;
  dim3 grid;
  ClusterPlan cp;
  if (dtype == SSB_F32) return (bwd_cluster_plan<float>(g, res_mode, &cp) || bwd_fused_plan<float>(g, res_mode, has_y != 0, &cgpc, &rpb, &grid)) ? 1 : 0;
  if (dtype == SSB_BF16) return (bwd_cluster_plan<bf16>(g, res_mode, &cp) || bwd_fused_plan<bf16>(g, res_mode, has_y != 0, &cgpc, &rpb, &grid)) ? 1 : 0;
  return 0;
}

int ssb_bn_bwd_fused(const void* g1, const void* y, const void* x, const ssb_bn* bn, void* dx, const void* x_res,
                     const ssb_bn* bn_res, void* dx_res, void* g_ident, ssb_geom g, uint32_t* barrier, double* rep,
                     double* rep_res, size_t rep_stride, int dtype, ssb_stream_t stream) {
  int rc = check_geom("ssb_bn_bwd_fused", g, 0);
  if (rc) return rc;
  SSB_REQUIRE(g1 && x && bn && dx && barrier && rep, "ssb_bn_bwd_fused: null pointer");
  SSB_REQUIRE(!bn_res || rep_res, "ssb_bn_bwd_fused: residual BN needs its replica scratch");
  SSB_REQUIRE(rep_stride >= (size_t)2 * g.C, "ssb_bn_bwd_fused: replica stride %zu < 2C", rep_stride);
  double* rep_r = rep_res;
  SSB_REQUIRE(!(bn_res && (!x_res || !dx_res)), "ssb_bn_bwd_fused: residual BN needs x_res and dx_res");
  SSB_REQUIRE(!(bn_res && g_ident), "ssb_bn_bwd_fused: g_ident and bn_res are exclusive");
  const int mode = bn_res ? 2 : (g_ident ? 1 : 0);
  SSB_DISPATCH_DTYPE(dtype, T, {
    int cgpc, rpb;
    dim3 grid;
    ClusterPlan cp;
    if (bwd_cluster_plan<T>(g, mode, &cp)) {     // clusters + DSMEM: one read of the operands, no grid barrier
      cudaStream_t st = to_stream(stream);
      const ssb_bn br = bn_res ? *bn_res : kNoBn;
      cudaError_t ce;
      if (mode == 0) ce = y ? SSB_BWC_NR(true, 0) : SSB_BWC_NR(false, 0);
      else if (mode == 1) ce = y ? SSB_BWC_NR(true, 1) : SSB_BWC_NR(false, 1);
      else ce = y ? SSB_BWC_NR(true, 2) : SSB_BWC_NR(false, 2);
      if (ce != cudaSuccess) {
        ssb_set_error("ssb_bn_bwd_fused: cluster launch failed: %s", cudaGetErrorString(ce));
        return SSB_ERR_CUDA;
      }
      SSB_LAUNCH_CHECK("ssb_bn_bwd_fused");
      return SSB_OK;
    }
    if (!bwd_fused_plan<T>(g, mode, y != nullptr, &cgpc, &rpb, &grid)) {
      ssb_set_error("ssb_bn_bwd_fused: %u x %u blocks exceed the co-resident capacity; use ssb_bn_bwd_reduce + ssb_bn_bwd_apply",
                    grid.x, grid.y);
      return SSB_ERR_UNSUPPORTED;
    }
    cudaStream_t st = to_stream(stream);
    const ssb_bn br = bn_res ? *bn_res : kNoBn;
    if (mode == 0) { if (y) SSB_BWF(true, 0); else SSB_BWF(false, 0); }
    else if (mode == 1) { if (y) SSB_BWF(true, 1); else SSB_BWF(false, 1); }
    else { if (y) SSB_BWF(true, 2); else SSB_BWF(false, 2); }
  })
  SSB_LAUNCH_CHECK("ssb_bn_bwd_fused");
  return SSB_OK;
}

int ssb_stem_bwd_reduce(const void* gp, const void* c0, const uint8_t* arg, const ssb_bn* bn, ssb_geom gin,
                        ssb_geom gout, int dtype, ssb_stream_t stream) {
  int rc = check_geom("ssb_stem_bwd_reduce(in)", gin, 0);
  if (rc) return rc;
  rc = check_geom("ssb_stem_bwd_reduce(out)", gout, 0);
  if (rc) return rc;
  SSB_REQUIRE(gp && c0 && bn && arg, "ssb_stem_bwd_reduce: null pointer");
  SSB_REQUIRE(gin.C == gout.C && gin.B == gout.B, "ssb_stem_bwd_reduce: geometry mismatch");
  const int rows = gin.B * gin.pitch;
  SSB_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    const int ncg = gin.C / V;
    const int cgb = ncg < BN_THREADS ? ncg : BN_THREADS;
    const int nrl = BN_THREADS / cgb;
    int rpb = ceil_div(rows, 148 * 4);
    if (rpb < nrl * 4) rpb = nrl * 4;
    dim3 grid(ceil_div(rows, rpb), ceil_div(ncg, cgb));
    ssb_launch(stem_bwd_reduce_kernel<T>, dim3(grid), dim3(BN_THREADS), 0, to_stream(stream), (const T*)gp, (const T*)c0, arg, *bn, gin, gout, rpb);
  })
  SSB_LAUNCH_CHECK("ssb_stem_bwd_reduce");
  return SSB_OK;
}

int ssb_stem_bwd_apply(const void* gp, const void* c0, const uint8_t* arg, const ssb_bn* bn, void* dc0, ssb_geom gin,
                       ssb_geom gout, int dtype, ssb_stream_t stream) {
  int rc = check_geom("ssb_stem_bwd_apply(in)", gin, 0);
  if (rc) return rc;
  rc = check_geom("ssb_stem_bwd_apply(out)", gout, 0);
  if (rc) return rc;
  SSB_REQUIRE(gp && c0 && bn && dc0 && arg, "ssb_stem_bwd_apply: null pointer");
  SSB_REQUIRE(gin.C == gout.C && gin.B == gout.B, "ssb_stem_bwd_apply: geometry mismatch");
  const size_t smem = (size_t)6 * gin.C * sizeof(float);
  SSB_DISPATCH_DTYPE(dtype, T, {
    const long long total = (long long)gin.B * gin.pitch * (gin.C / Vec<T>::N);
    if (bn->sync_peers) ssb_launch(stem_bwd_apply_kernel<T, true>, dim3(ew_blocks(total)), dim3(BN_THREADS), smem, to_stream(stream), (const T*)gp, (const T*)c0, arg, (T*)dc0, *bn, gin, gout);
    else ssb_launch(stem_bwd_apply_kernel<T, false>, dim3(ew_blocks(total)), dim3(BN_THREADS), smem, to_stream(stream), (const T*)gp, (const T*)c0, arg, (T*)dc0, *bn, gin, gout);
  })
  SSB_LAUNCH_CHECK("ssb_stem_bwd_apply");
  return SSB_OK;
}

}  // extern "C"

SSB_TRACE_DEFINE(bn)
