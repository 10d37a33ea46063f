// Generic (any channel count, fp32 accumulate) Conv1d kernels on CUDA cores:
//  * stem direct conv k7 s2 p3 from the NCL fp32 input (HBM-bound, resnet.py:246-253)
//  * tap-GEMM fprop / dgrad for k in {1,3}, stride in {1,2} over flat padded NLC rows
//  * wgrad (split over rows, fp32 atomics into the reference-layout gradient)
//  * flat storage-dtype copy of the parameter arena (conv weights are kept tap-major, [k][Cin][Cout])
// These are the exact-parity (fp32) path and the fallback for shapes the tcgen05 kernels
// (conv_sm100.cu) do not cover.  Replaces cuDNN fprop/dgrad/wgrad (SURVEY.md 2.2 K1,K4).
#include <cooperative_groups.h>

#include "common.cuh"

// ---------------------------------------------------------------------------------------
// tap-GEMM:  Out[o_mul*m + o_off][n] (+)= sum_tap sum_k A[a_mul*m + a_off[tap]][k] * W[w_tap[tap]][k][n]
// (WT: the tap matrices are stored transposed, W[tap][n][k] -- dgrad reading the [k][Cin][Cout] weights)
// ---------------------------------------------------------------------------------------
struct TapSpec {
  int ntaps;
  int a_off[3];
  int w_tap[3];
};

#define TG_BM 128
#define TG_BN 64
#define TG_BK 16
#define TG_THREADS 256

template <typename T>
__device__ __forceinline__ void load8(const T* p, float* f);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* f) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float* f) {
  Vec<bf16> v;
  v.load(p);
  v.get(f);
}
template <typename T>
__device__ __forceinline__ void load4(const T* p, float* f);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* f) {
  float4 a = *reinterpret_cast<const float4*>(p);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float* f) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float* f);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, const float* f) {
  uint2 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  h[0] = __floats2bfloat162_rn(f[0], f[1]);
  h[1] = __floats2bfloat162_rn(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = raw;
}

template <typename T, bool ACC, bool WT>
__global__ void __launch_bounds__(TG_THREADS)
tap_gemm_kernel(const T* __restrict__ A, const T* __restrict__ W, T* __restrict__ Out, int M, int N, int K,
                int a_rows, int a_mul, TapSpec taps, int o_mul, int o_off, int o_rows, int o_pitch, int o_len) {
  pdl_trigger();
  pdl_wait();
  __shared__ float As[TG_BK][TG_BM + 4];
  __shared__ float Bs[TG_BK][TG_BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TG_BM;
  const int n0 = blockIdx.y * TG_BN;
  const int ty = tid / 16, tx = tid % 16;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int la_row = tid >> 1, la_k = (tid & 1) * 8;  // A loader: row, k offset
  const int lb_k = tid >> 4, lb_n = (tid & 15) * 4;   // B loader
  const int lt_n = tid >> 2, lt_k = (tid & 3) * 4;    // B loader, transposed storage: 4 consecutive k of one n
  for (int tp = 0; tp < taps.ntaps; ++tp) {
    const int m = m0 + la_row;
    const long long arow = (long long)a_mul * m + taps.a_off[tp];
    const bool arow_ok = m < M && arow >= 0 && arow < a_rows;
    const T* wt = W + (size_t)taps.w_tap[tp] * K * N;
    for (int k0 = 0; k0 < K; k0 += TG_BK) {
      float fa[8];
      if (arow_ok && k0 + la_k < K) {
        load8<T>(A + (size_t)arow * K + k0 + la_k, fa);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) fa[i] = 0.f;
      }
      float fb[4];
      if (WT) {
        if (k0 + lt_k < K && n0 + lt_n < N) {
          load4<T>(wt + (size_t)(n0 + lt_n) * K + k0 + lt_k, fb);
        } else {
          fb[0] = fb[1] = fb[2] = fb[3] = 0.f;
        }
      } else if (k0 + lb_k < K && n0 + lb_n < N) {
        load4<T>(wt + (size_t)(k0 + lb_k) * N + n0 + lb_n, fb);
      } else {
        fb[0] = fb[1] = fb[2] = fb[3] = 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) As[la_k + i][la_row] = fa[i];
      if (WT) {
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[lt_k + i][lt_n] = fb[i];
      } else {
        *reinterpret_cast<float4*>(&Bs[lb_k][lb_n]) = make_float4(fb[0], fb[1], fb[2], fb[3]);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TG_BK; ++kk) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
    }
  }
  const int n = n0 + tx * 4;
  if (n >= N) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const long long orow = (long long)o_mul * m + o_off;
    if (orow < 0 || orow >= o_rows) continue;
    T* op = Out + (size_t)orow * N + n;
    float o[4];
    if (row_valid((int)orow, o_pitch, o_len)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = acc[i][j];
      if (ACC) {
        float prev[4];
        load4<T>(op, prev);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] += prev[j];
      }
      store4<T>(op, o);
    } else if (!ACC) {
      o[0] = o[1] = o[2] = o[3] = 0.f;
      store4<T>(op, o);
    }
  }
}

// zero-fill rows of one parity (used when a stride-2 dgrad has no tap for that parity)
template <typename T>
__global__ void zero_parity_rows_kernel(T* out, int rows, int N, int parity) {
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

// ---------------------------------------------------------------------------------------
// wgrad: dW[tap][ci][co] += sum_m X[a_mul*m + a_off[tap]][ci] * dY[m][co]
// ---------------------------------------------------------------------------------------
#define WG_T 64
#define WG_BK 16

template <typename T>
__global__ void __launch_bounds__(256)
wgrad_kernel(const T* __restrict__ X, const T* __restrict__ dY, float* __restrict__ dW, int M, int Cin, int Cout,
             int k, int x_rows, int a_mul, TapSpec taps, int nsplit, int rows_per_split) {
  pdl_trigger();
  pdl_wait();
  __shared__ float Xs[WG_BK][WG_T];
  __shared__ float Ys[WG_BK][WG_T];
  const int tid = threadIdx.x;
  const int ci0 = blockIdx.x * WG_T, co0 = blockIdx.y * WG_T;
  const int tp = blockIdx.z / nsplit, sp = blockIdx.z % nsplit;
  const int ty = tid / 16, tx = tid % 16;
  const int lr = tid >> 4, lc = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int mbeg = sp * rows_per_split;
  const int mend = min(M, mbeg + rows_per_split);
  const int aoff = taps.a_off[tp];
  for (int mm = mbeg; mm < mend; mm += WG_BK) {
    const int m = mm + lr;
    float fx[4] = {0.f, 0.f, 0.f, 0.f}, fy[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      const long long xr = (long long)a_mul * m + aoff;
      if (xr >= 0 && xr < x_rows && ci0 + lc < Cin) load4<T>(X + (size_t)xr * Cin + ci0 + lc, fx);
      if (co0 + lc < Cout) load4<T>(dY + (size_t)m * Cout + co0 + lc, fy);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Xs[lr][lc]) = make_float4(fx[0], fx[1], fx[2], fx[3]);
    *reinterpret_cast<float4*>(&Ys[lr][lc]) = make_float4(fy[0], fy[1], fy[2], fy[3]);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WG_BK; ++r) {
      float4 a = *reinterpret_cast<const float4*>(&Xs[r][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Ys[r][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  const int wt = taps.w_tap[tp];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= Cout) continue;
      atomicAdd(&dW[((size_t)wt * Cin + ci) * Cout + co], acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// storage-dtype copy of the whole parameter arena (same offsets): what the bf16 convs read.
// The weights of the GEMM convs are kept in [k][Cin][Cout] in the master arena itself, so this
// is a flat conversion (no per-tensor repack, no transposed copy).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weight_shadow_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n8) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    Vec<bf16> o;
    o.set(f);
    o.store(dst + 8 * i);
  }
}

// ---------------------------------------------------------------------------------------
// stem: y[b, t, co] = sum_{c,j} w[co][c][j] * x[b][c][2t + j - 3]
// Register tiles: a thread owns 8 consecutive positions x 4 output channels (32 accumulators); per (lead, tap) it
// needs one 16-byte weight load and a sliding 21-sample window of the input that is loaded once per lead.
// ---------------------------------------------------------------------------------------
#define ST_THREADS 256
#define ST_PT 8      // positions per thread

template <typename T>
__global__ void __launch_bounds__(ST_THREADS)
stem_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ y, int Cl, int L,
                     ssb_geom g, int tt) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int Cs = g.C;
  const int cog = Cs / 4;                 // channel groups of 4
  const int XW = 2 * tt + 5;
  float* ws = sm;                         // [Cl*7][Cs]
  float* xs = sm + Cl * 7 * Cs;           // [Cl][XW]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * tt;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < Cs * Cl * 7; idx += ST_THREADS) {
    const int co = idx / (Cl * 7), cj = idx % (Cl * 7);
    ws[cj * Cs + co] = w[idx];
  }
  for (int idx = tid; idx < Cl * XW; idx += ST_THREADS) {
    const int c = idx / XW, i = idx % XW;
    const int l = 2 * t0 - 3 + i;
    xs[idx] = (l >= 0 && l < L) ? x[((size_t)b * Cl + c) * L + l] : 0.f;
  }
  __syncthreads();
  const int cg = tid % cog, tg = tid / cog;
  const int tb = tg * ST_PT;              // first position of this thread inside the tile
  if (tb < tt) {
    float acc[ST_PT][4];
#pragma unroll
    for (int u = 0; u < ST_PT; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
    for (int c = 0; c < Cl; ++c) {
      float xv[2 * ST_PT + 5];
#pragma unroll
      for (int i = 0; i < 2 * ST_PT + 5; ++i) xv[i] = xs[c * XW + 2 * tb + i];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(&ws[(c * 7 + j) * Cs + cg * 4]);
#pragma unroll
        for (int u = 0; u < ST_PT; ++u) {
          const float xx = xv[2 * u + j];
          acc[u][0] = fmaf(wv.x, xx, acc[u][0]);
          acc[u][1] = fmaf(wv.y, xx, acc[u][1]);
          acc[u][2] = fmaf(wv.z, xx, acc[u][2]);
          acc[u][3] = fmaf(wv.w, xx, acc[u][3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < ST_PT; ++u) {
      const int t = t0 + tb + u;
      if (t < g.len) store4<T>(y + ((size_t)b * g.pitch + 1 + t) * Cs + cg * 4, acc[u]);
    }
  }
  // halo / pad rows of this sample stay zero
  if (blockIdx.x == 0)
    for (int c = tid; c < Cs; c += ST_THREADS) y[(size_t)b * g.pitch * Cs + c] = from_f<T>(0.f);
  if (blockIdx.x == gridDim.x - 1) {
    const int npad = g.pitch - g.len - 1;
    for (int idx = tid; idx < npad * Cs; idx += ST_THREADS)
      y[((size_t)b * g.pitch + g.len + 1) * Cs + idx] = from_f<T>(0.f);
  }
}

// dw[co][c][j] += sum_{b,t} dy[b,t,co] * x[b][c][2t+j-3]
// One CTA per SM walks over (sample, 64-position tile) work items.  A thread owns 4 output channels x the 7 taps
// of one lead (28 accumulators, kept in registers across all the tiles the CTA walks) for a slice of the tile's
// positions; the slices are combined through shared memory at the end: one fp32 atomic per output per CTA.
#define SW_THREADS 512
#define SW_TT 64
template <typename T>
__global__ void __launch_bounds__(SW_THREADS)
stem_conv_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int Cl, int L,
                       ssb_geom g, int ntt, int ts) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int Cs = g.C;
  const int cog = Cs / 4;
  const int XW = 2 * SW_TT + 5;
  const int nout = Cl * 7 * Cs;
  float* dys = sm;                   // [TT][Cs]
  float* xs = sm + SW_TT * Cs;       // [Cl][XW]
  float* acc_s = xs + Cl * XW;       // [nout]
  const int tid = threadIdx.x;
  const int nown = cog * Cl;         // (channel group, lead) owners per position slice
  const int own = tid % nown, sl = tid / nown;
  const bool active = sl < ts;
  const int cg = own % cog, c = own / cog;
  float acc[7][4];
#pragma unroll
  for (int j = 0; j < 7; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  for (int o = tid; o < nout; o += SW_THREADS) acc_s[o] = 0.f;
  const int ntiles = ntt * g.B;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / ntt;
    const int t0 = (tile - b * ntt) * SW_TT;
    __syncthreads();   // previous tile's readers are done with dys / xs
    constexpr int V = Vec<T>::N;
    for (int idx = tid; idx < SW_TT * Cs / V; idx += SW_THREADS) {   // 16-byte loads (Cs is a multiple of 8)
      const int e = idx * V;
      const int t = t0 + e / Cs;
      float f[V];
      if (t < g.len) {
        Vec<T> v;
        v.load(dy + ((size_t)b * g.pitch + 1 + t) * Cs + e % Cs);
        v.get(f);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) dys[e + i] = f[i];
    }
    for (int idx = tid; idx < Cl * XW; idx += SW_THREADS) {
      const int cc = idx / XW, i = idx % XW;
      const int l = 2 * t0 - 3 + i;
      xs[idx] = (l >= 0 && l < L) ? x[((size_t)b * Cl + cc) * L + l] : 0.f;
    }
    __syncthreads();
    if (active) {
      const float* xr = xs + c * XW;
      for (int t = sl; t < SW_TT; t += ts) {
        const float4 d = *reinterpret_cast<const float4*>(&dys[t * Cs + cg * 4]);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const float xx = xr[2 * t + j];
          acc[j][0] = fmaf(d.x, xx, acc[j][0]);
          acc[j][1] = fmaf(d.y, xx, acc[j][1]);
          acc[j][2] = fmaf(d.z, xx, acc[j][2]);
          acc[j][3] = fmaf(d.w, xx, acc[j][3]);
        }
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(&acc_s[(c * 7 + j) * Cs + cg * 4 + i], acc[j][i]);   // shared-memory combine of the slices
  }
  __syncthreads();
  for (int o = tid; o < nout; o += SW_THREADS) {
    const int co = o % Cs, cj = o / Cs;
    atomicAdd(&dw[(size_t)co * Cl * 7 + cj], acc_s[o]);
  }
}

// Stem weight gradient for a single lead (the shipped configs): no shared-memory staging of dy -- every thread streams
// 16-byte dy vectors (V channels of one position) straight from L2, four positions in flight, and keeps its
// [7 taps][V] products in registers; positions are spread over the block's lanes and over ~100 blocks.  Combine: warp
// shuffles, per-warp partial rows in shared memory summed by the block, one global atomic per output per block.
// The blocks of a cluster of SD_CLUSTER are then summed through distributed shared memory, so that the global REDs are
// one per output per cluster.  Measured 11.3 us against 21.2 us for the tiled kernel above at the config-2 shape
// (tools/stem_wgrad_timing.py; 15.9 us before the cluster combine).
#define SD_THREADS 256
#define SD_POS_PER_LANE 12
#define SD_UNROLL 4
#define SD_CLUSTER 8
template <typename T, int CL>
__global__ void __launch_bounds__(SD_THREADS)
stem_conv_wgrad_direct_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int L, ssb_geom g,
                              int chunk) {
  pdl_trigger();
  pdl_wait();
  static_assert(CL == 1, "single lead");
  constexpr int V = Vec<T>::N;
  extern __shared__ float part[];           // [warps][7][Cs]
  const int Cs = g.C;
  const int ncg = Cs / V;                   // power of two, <= 32
  const int plane = SD_THREADS / ncg;
  const int cg = threadIdx.x % ncg, lane_p = threadIdx.x / ncg;
  const int nout = 7 * Cs;
  float acc[7][V];
#pragma unroll
  for (int j = 0; j < 7; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[j][i] = 0.f;
  const int P = g.B * g.len;                // (< 2^31: checked by the host)
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, P);
  for (int p = p0 + lane_p; p < p1; p += SD_UNROLL * plane) {
    Vec<T> v[SD_UNROLL];
    float xv[SD_UNROLL][7];
#pragma unroll
    for (int u = 0; u < SD_UNROLL; ++u) {     // issue every load of the batch before the first use
      const int pp = p + u * plane;
      if (pp < p1) {
        const int b = pp / g.len, t = pp - b * g.len;
        v[u].load(dy + ((size_t)b * g.pitch + 1 + t) * Cs + (size_t)cg * V);
        const float* xr = x + (size_t)b * L;
        const int l0 = 2 * t - 3;
        const bool interior = l0 >= 0 && l0 + 6 < L;
#pragma unroll
        for (int j = 0; j < 7; ++j) xv[u][j] = (interior || (l0 + j >= 0 && l0 + j < L)) ? xr[l0 + j] : 0.f;
      } else {
        v[u].zero();
#pragma unroll
        for (int j = 0; j < 7; ++j) xv[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < SD_UNROLL; ++u) {
      float d[V];
      v[u].get(d);
#pragma unroll
      for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[j][i] = fmaf(d[i], xv[u][j], acc[j][i]);
    }
  }
  // lanes of a warp with the same channel group: xor offsets ncg, 2 ncg, ... < 32; then one partial row per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 7; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float a = acc[j][i];
      for (int off = ncg; off < 32; off <<= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      if (lane < ncg) part[(size_t)warp * nout + j * Cs + cg * V + i] = a;
    }
  __syncthreads();
  // block totals, then the SD_CLUSTER blocks of a cluster are summed by its rank-0 block through distributed shared
  // memory: the kernel's tail used to be ~47k global REDs onto 448 addresses (14 cache lines) -- ncu: SMs active
  // 17.8k of 43.4k elapsed cycles, the rest the REDs draining at L2 -- now one RED per output per CLUSTER
  float* tot = part + (size_t)(SD_THREADS / 32) * nout;
  for (int o = threadIdx.x; o < nout; o += SD_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < SD_THREADS / 32; ++w) a += part[(size_t)w * nout + o];
    tot[o] = a;
  }
  namespace cgs = cooperative_groups;
  cgs::cluster_group cluster = cgs::this_cluster();
  cluster.sync();
  if (cluster.block_rank() == 0) {
    const unsigned nr = cluster.num_blocks();
    for (int o = threadIdx.x; o < nout; o += SD_THREADS) {
      float a = 0.f;
      for (unsigned r = 0; r < nr; ++r) a += cluster.map_shared_rank(tot, r)[o];
      const int co = o % Cs, cj = o / Cs;
      if (a != 0.f) atomicAdd(&dw[(size_t)co * 7 + cj], a);
    }
  }
  cluster.sync();      // the peers' shared memory stays alive until rank 0 has read it
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int check_conv_geom(const char* who, const ssb_geom& gi, const ssb_geom& go, int k, int stride) {
  SSB_REQUIRE(k == 1 || k == 3, "%s: kernel size %d not supported (1 or 3)", who, k);
  SSB_REQUIRE(stride == 1 || stride == 2, "%s: stride %d not supported (1 or 2)", who, stride);
  SSB_REQUIRE(gi.B == go.B && gi.B > 0, "%s: batch mismatch", who);
  SSB_REQUIRE(gi.C % 8 == 0 && go.C % 8 == 0 && gi.C > 0 && go.C > 0, "%s: channels must be multiples of 8 (Cin=%d Cout=%d)", who, gi.C, go.C);
  SSB_REQUIRE(gi.pitch == stride * go.pitch, "%s: pitch_in %d != stride*pitch_out %d", who, gi.pitch, stride * go.pitch);
  SSB_REQUIRE(go.len == (gi.len - 1) / stride + 1, "%s: len_out %d inconsistent with len_in %d stride %d", who, go.len, gi.len, stride);
  SSB_REQUIRE(gi.pitch >= gi.len + 2 && go.pitch >= go.len + 2, "%s: pitch too small", who);
  SSB_REQUIRE((long long)gi.B * gi.pitch < (1ll << 31), "%s: too many rows", who);
  return SSB_OK;
}

// taps of the forward conv in flat-row space: in_row = stride*m + a_off[j], m = output row
static TapSpec fwd_taps(int k, int stride) {
  TapSpec t;
  t.ntaps = k;
  for (int j = 0; j < 3; ++j) {
    t.a_off[j] = 0;
    t.w_tap[j] = j < k ? j : 0;
  }
  if (stride == 1) {
    if (k == 3) { t.a_off[0] = -1; t.a_off[1] = 0; t.a_off[2] = 1; }
  } else {
    // in_row = 2*(m-1) + j + (k==1 ? 1 : 0)
    if (k == 3) { t.a_off[0] = -2; t.a_off[1] = -1; t.a_off[2] = 0; }
    else t.a_off[0] = -1;
  }
  return t;
}

int ssb_conv1d_fwd_sm100(const void* x, const void* w, void* y, ssb_geom gin, ssb_geom gout, int k, int stride,
                         double* stats, const ssb_bn* ep_bn, const void* ep_res, int ep_relu, int train_samples,
                         void* y_eval, const ssb_bnf_args* bnf, cudaStream_t st);
int ssb_conv1d_dgrad_sm100(const void* dy, const void* w, void* dx, ssb_geom gin, ssb_geom gout, int k, int stride,
                           int accumulate, const void* red_y, const void* red_x, const ssb_bn* red_bn, const void* red_xr,
                           const ssb_bn* red_bn_r, cudaStream_t st);
int ssb_conv1d_wgrad_sm100(const void* x, const void* dy, float* dw, ssb_geom gin, ssb_geom gout, int k, int stride,
                           cudaStream_t st);

int ssb_conv1d_fwd_bnf_fits_sm100(ssb_geom gin, ssb_geom gout, int k, int stride);
int ssb_stem_conv_fwd_sm100(const float* x, const float* w, void* y, int Cl, int L, ssb_geom g, double* stats, cudaStream_t st);
int ssb_stem_conv_wgrad_sm100(const float* x, const void* dy, float* dw, int Cl, int L, ssb_geom g, cudaStream_t st);

int ssb_simt_prepare() {
  const int big = 96 * 1024;
  cudaError_t e = cudaFuncSetAttribute(stem_conv_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_conv_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_conv_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_conv_wgrad_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e != cudaSuccess) {
    ssb_set_error("ssb_simt_prepare: %s", cudaGetErrorString(e));
    return SSB_ERR_CUDA;
  }
  return SSB_OK;
}

extern "C" {

int ssb_conv1d_fwd(const void* x, const void* w, void* y, ssb_geom gin, ssb_geom gout, int k, int stride, int dtype,
                   int algo, ssb_stream_t stream) {
  return ssb_conv1d_fwd_stats(x, w, y, gin, gout, k, stride, nullptr, dtype, algo, stream);
}

int ssb_conv1d_fwd_stats(const void* x, const void* w, void* y, ssb_geom gin, ssb_geom gout, int k, int stride,
                         double* sums, int dtype, int algo, ssb_stream_t stream) {
  int rc = check_conv_geom("ssb_conv1d_fwd", gin, gout, k, stride);
  if (rc) return rc;
  SSB_REQUIRE(x && y && w, "ssb_conv1d_fwd: null pointer");
  if (algo == SSB_ALGO_TCGEN05) {
    SSB_REQUIRE(dtype == SSB_BF16, "ssb_conv1d_fwd: tcgen05 path needs bf16");
    return ssb_conv1d_fwd_sm100(x, w, y, gin, gout, k, stride, sums, nullptr, nullptr, 0, 0, nullptr, nullptr, to_stream(stream));
  }
  if (sums) {   // generic CUDA-core path: conv, then the statistics pass as its own launch
    rc = ssb_conv1d_fwd_stats(x, w, y, gin, gout, k, stride, nullptr, dtype, algo, stream);
    if (rc) return rc;
    return ssb_bn_stats(y, gout, sums, dtype, stream);
  }
  const int M = gout.B * gout.pitch;
  const TapSpec taps = fwd_taps(k, stride);
  dim3 grid(ceil_div(M, TG_BM), ceil_div(gout.C, TG_BN));
  SSB_DISPATCH_DTYPE(dtype, T, {
    ssb_launch(tap_gemm_kernel<T, false, false>, dim3(grid), dim3(TG_THREADS), 0, to_stream(stream),
        (const T*)x, (const T*)w, (T*)y, M, gout.C, gin.C, gin.B * gin.pitch, stride, taps, 1, 0, M, gout.pitch,
        gout.len);
  })
  SSB_LAUNCH_CHECK("ssb_conv1d_fwd");
  return SSB_OK;
}

int ssb_conv1d_bn_act_fwd(const void* x, const void* w, void* y, ssb_geom gin, ssb_geom gout, int k, int stride,
                          const ssb_bn* bn, const void* res, int relu, int dtype, int algo, ssb_stream_t stream) {
  int rc = check_conv_geom("ssb_conv1d_bn_act_fwd", gin, gout, k, stride);
  if (rc) return rc;
  SSB_REQUIRE(x && y && w && bn, "ssb_conv1d_bn_act_fwd: null pointer");
  if (algo == SSB_ALGO_TCGEN05) {
    SSB_REQUIRE(dtype == SSB_BF16, "ssb_conv1d_bn_act_fwd: tcgen05 path needs bf16");
    return ssb_conv1d_fwd_sm100(x, w, y, gin, gout, k, stride, nullptr, bn, res, relu, 0, nullptr, nullptr, to_stream(stream));
  }
  // generic CUDA-core path: conv, then the eval-mode BN(+residual)(+ReLU) pass in place
  rc = ssb_conv1d_fwd(x, w, y, gin, gout, k, stride, dtype, algo, stream);
  if (rc) return rc;
  return ssb_bn_act_fwd(y, bn, res, nullptr, y, gout, relu, 0, dtype, stream);
}

int ssb_conv1d_fwd_bn_train_fits(ssb_geom gin, ssb_geom gout, int k, int stride, int dtype, int algo) {
  if (algo != SSB_ALGO_TCGEN05 || dtype != SSB_BF16) return 0;
  if (check_conv_geom("ssb_conv1d_fwd_bn_train_fits", gin, gout, k, stride)) return 0;
  return ssb_conv1d_fwd_bnf_fits_sm100(gin, gout, k, stride);
}

int ssb_conv1d_fwd_bn_train(const void* x, const void* w, void* y_raw, void* y_act, ssb_geom gin, ssb_geom gout, int k,
                            int stride, const ssb_bn* bn, const void* res, const ssb_bn* bn_res, int relu, uint32_t* barrier,
                            int dtype, int algo, ssb_stream_t stream) {
  int rc = check_conv_geom("ssb_conv1d_fwd_bn_train", gin, gout, k, stride);
  if (rc) return rc;
  SSB_REQUIRE(x && w && y_raw && y_act && bn && barrier, "ssb_conv1d_fwd_bn_train: null pointer");
  SSB_REQUIRE(!(bn_res && !res), "ssb_conv1d_fwd_bn_train: bn_res given without res");
  SSB_REQUIRE(bn->count_mul <= 1, "ssb_conv1d_fwd_bn_train: not for SyncBN (the statistics exchange sits between the passes)");
  if (algo != SSB_ALGO_TCGEN05 || dtype != SSB_BF16) {
    ssb_set_error("ssb_conv1d_fwd_bn_train: only the bf16 tcgen05 path fuses the BatchNorm pass; use ssb_conv1d_fwd_stats + ssb_bn_act_fwd");
    return SSB_ERR_UNSUPPORTED;
  }
  ssb_bnf_args a = {bn, bn_res, res, y_act, relu, barrier};
  return ssb_conv1d_fwd_sm100(x, w, y_raw, gin, gout, k, stride, nullptr, nullptr, nullptr, 0, 0, nullptr, &a, to_stream(stream));
}

int ssb_conv1d_fwd_dual(const void* x, const void* w, void* y_train, void* y_eval, ssb_geom gin, ssb_geom gout, int k,
                        int stride, int train_samples, double* sums, const ssb_bn* bn_eval, const void* res_eval, int relu,
                        int dtype, int algo, ssb_stream_t stream) {
  int rc = check_conv_geom("ssb_conv1d_fwd_dual", gin, gout, k, stride);
  if (rc) return rc;
  SSB_REQUIRE(x && w && y_train && y_eval && sums && bn_eval, "ssb_conv1d_fwd_dual: null pointer");
  SSB_REQUIRE(train_samples > 0 && train_samples < gout.B, "ssb_conv1d_fwd_dual: train_samples %d not inside (0, B=%d)", train_samples, gout.B);
  if (algo == SSB_ALGO_TCGEN05) {
    SSB_REQUIRE(dtype == SSB_BF16, "ssb_conv1d_fwd_dual: tcgen05 path needs bf16");
    return ssb_conv1d_fwd_sm100(x, w, y_train, gin, gout, k, stride, sums, bn_eval, res_eval, relu, train_samples, y_eval,
                                nullptr, to_stream(stream));
  }
  // generic CUDA-core path: the two row ranges as separate launches (train rows + statistics, eval rows + BN pass)
  const size_t es = dtype == SSB_BF16 ? 2 : 4;
  ssb_geom gi_t = gin, go_t = gout, gi_e = gin, go_e = gout;
  gi_t.B = go_t.B = train_samples;
  gi_e.B = go_e.B = gout.B - train_samples;
  const size_t in_off = (size_t)train_samples * gin.pitch * gin.C * es, out_off = (size_t)train_samples * gout.pitch * gout.C * es;
  rc = ssb_conv1d_fwd_stats(x, w, y_train, gi_t, go_t, k, stride, sums, dtype, algo, stream);
  if (rc) return rc;
  return ssb_conv1d_bn_act_fwd((const char*)x + in_off, w, (char*)y_eval + out_off, gi_e, go_e, k, stride, bn_eval,
                               res_eval ? (const char*)res_eval + out_off : nullptr, relu, dtype, algo, stream);
}

int ssb_conv1d_dgrad(const void* dy, const void* w, void* dx, ssb_geom gin, ssb_geom gout, int k, int stride,
                     int accumulate, int dtype, int algo, ssb_stream_t stream) {
  int rc = check_conv_geom("ssb_conv1d_dgrad", gin, gout, k, stride);
  if (rc) return rc;
  SSB_REQUIRE(dy && dx && w, "ssb_conv1d_dgrad: null pointer");
  if (algo == SSB_ALGO_TCGEN05) {
    SSB_REQUIRE(dtype == SSB_BF16, "ssb_conv1d_dgrad: tcgen05 path needs bf16");
    return ssb_conv1d_dgrad_sm100(dy, w, dx, gin, gout, k, stride, accumulate, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  to_stream(stream));
  }
  cudaStream_t st = to_stream(stream);
  const int rows_in = gin.B * gin.pitch, rows_out = gout.B * gout.pitch;
  // GEMM: M over dx rows (or row pairs), N = Cin, K = Cout; the tap matrices of w are [Cin][Cout] = [N][K] (WT)
  SSB_DISPATCH_DTYPE(dtype, T, {
    if (stride == 1) {
      TapSpec t;
      t.ntaps = k;
      for (int j = 0; j < 3; ++j) { t.a_off[j] = (k == 3) ? 1 - j : 0; t.w_tap[j] = j < k ? j : 0; }
      dim3 grid(ceil_div(rows_in, TG_BM), ceil_div(gin.C, TG_BN));
      if (accumulate)
        ssb_launch(tap_gemm_kernel<T, true, true>, dim3(grid), dim3(TG_THREADS), 0, st, (const T*)dy, (const T*)w, (T*)dx, rows_in, gin.C, gout.C, rows_out, 1, t, 1, 0, rows_in, gin.pitch, gin.len);
      else
        ssb_launch(tap_gemm_kernel<T, false, true>, dim3(grid), dim3(TG_THREADS), 0, st, (const T*)dy, (const T*)w, (T*)dx, rows_in, gin.C, gout.C, rows_out, 1, t, 1, 0, rows_in, gin.pitch, gin.len);
      SSB_LAUNCH_CHECK("ssb_conv1d_dgrad");
    } else {
      // dx row 2q+p: p=0 <- dy[q+1]*W0 + dy[q]*W2 ; p=1 <- dy[q+1]*W1   (k=1: p=1 <- dy[q+1]*W0)
      const int Mq = rows_in / 2;
      dim3 grid(ceil_div(Mq, TG_BM), ceil_div(gin.C, TG_BN));
      for (int p = 0; p < 2; ++p) {
        TapSpec t;
        t.ntaps = 0;
        for (int j = 0; j < 3; ++j) { t.a_off[j] = 0; t.w_tap[j] = 0; }
        if (k == 3) {
          if (p == 0) { t.ntaps = 2; t.a_off[0] = 1; t.w_tap[0] = 0; t.a_off[1] = 0; t.w_tap[1] = 2; }
          else { t.ntaps = 1; t.a_off[0] = 1; t.w_tap[0] = 1; }
        } else if (p == 1) {
          t.ntaps = 1; t.a_off[0] = 1; t.w_tap[0] = 0;
        }
        if (t.ntaps == 0) {
          if (!accumulate) {
            ssb_launch(zero_parity_rows_kernel<T>, dim3(148 * 2), dim3(256), 0, st, (T*)dx, rows_in, gin.C, p);
            SSB_LAUNCH_CHECK("ssb_conv1d_dgrad(zero)");
          }
          continue;
        }
        if (accumulate)
          ssb_launch(tap_gemm_kernel<T, true, true>, dim3(grid), dim3(TG_THREADS), 0, st, (const T*)dy, (const T*)w, (T*)dx, Mq, gin.C, gout.C, rows_out, 1, t, 2, p, rows_in, gin.pitch, gin.len);
        else
          ssb_launch(tap_gemm_kernel<T, false, true>, dim3(grid), dim3(TG_THREADS), 0, st, (const T*)dy, (const T*)w, (T*)dx, Mq, gin.C, gout.C, rows_out, 1, t, 2, p, rows_in, gin.pitch, gin.len);
        SSB_LAUNCH_CHECK("ssb_conv1d_dgrad");
      }
    }
  })
  return SSB_OK;
}

int ssb_conv1d_dgrad_bnred(const void* dy, const void* w, void* dx, ssb_geom gin, ssb_geom gout, int k, int stride,
                           int accumulate, const void* y_act, const void* x_pre, const ssb_bn* bn, const void* x_res,
                           const ssb_bn* bn_res, int dtype, int algo, ssb_stream_t stream) {
  int rc = check_conv_geom("ssb_conv1d_dgrad_bnred", gin, gout, k, stride);
  if (rc) return rc;
  SSB_REQUIRE(dy && dx && w && y_act && x_pre && bn, "ssb_conv1d_dgrad_bnred: null pointer");
  SSB_REQUIRE((x_res != nullptr) == (bn_res != nullptr), "ssb_conv1d_dgrad_bnred: x_res and bn_res go together");
  SSB_REQUIRE(stride == 1, "ssb_conv1d_dgrad_bnred: stride-1 convs only (a stride-2 dgrad writes dx in two launches)");
  if (algo == SSB_ALGO_TCGEN05) {
    SSB_REQUIRE(dtype == SSB_BF16, "ssb_conv1d_dgrad_bnred: tcgen05 path needs bf16");
    return ssb_conv1d_dgrad_sm100(dy, w, dx, gin, gout, k, stride, accumulate, y_act, x_pre, bn, x_res, bn_res, to_stream(stream));
  }
  // generic CUDA-core path: dgrad, then the reduce pass as its own launch
  rc = ssb_conv1d_dgrad(dy, w, dx, gin, gout, k, stride, accumulate, dtype, algo, stream);
  if (rc) return rc;
  return ssb_bn_bwd_reduce(dx, nullptr, y_act, x_pre, bn, x_res, bn_res, gin, dtype, stream);
}

int ssb_conv1d_wgrad(const void* x, const void* dy, float* dw, ssb_geom gin, ssb_geom gout, int k, int stride,
                     int dtype, int algo, ssb_stream_t stream) {
  int rc = check_conv_geom("ssb_conv1d_wgrad", gin, gout, k, stride);
  if (rc) return rc;
  SSB_REQUIRE(x && dy && dw, "ssb_conv1d_wgrad: null pointer");
  if (algo == SSB_ALGO_TCGEN05) {
    SSB_REQUIRE(dtype == SSB_BF16, "ssb_conv1d_wgrad: tcgen05 path needs bf16");
    return ssb_conv1d_wgrad_sm100(x, dy, dw, gin, gout, k, stride, to_stream(stream));
  }
  const int M = gout.B * gout.pitch;
  const TapSpec taps = fwd_taps(k, stride);
  const int tiles = ceil_div(gin.C, WG_T) * ceil_div(gout.C, WG_T) * k;
  int nsplit = ceil_div(148 * 4, tiles);
  const int max_split = ceil_div(M, 4 * WG_BK);
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  int rps = ceil_div(M, nsplit);
  rps = ceil_div(rps, WG_BK) * WG_BK;
  nsplit = ceil_div(M, rps);
  dim3 grid(ceil_div(gin.C, WG_T), ceil_div(gout.C, WG_T), k * nsplit);
  SSB_DISPATCH_DTYPE(dtype, T, {
    ssb_launch(wgrad_kernel<T>, dim3(grid), dim3(256), 0, to_stream(stream), (const T*)x, (const T*)dy, dw, M, gin.C, gout.C, k,
                                                          gin.B * gin.pitch, stride, taps, nsplit, rps);
  })
  SSB_LAUNCH_CHECK("ssb_conv1d_wgrad");
  return SSB_OK;
}

int ssb_weight_shadow(const float* src, void* dst, size_t n, int dtype, ssb_stream_t stream) {
  SSB_REQUIRE(src && dst && n > 0 && n % 8 == 0, "ssb_weight_shadow: bad arguments (n=%zu must be a positive multiple of 8)", n);
  SSB_REQUIRE(dtype == SSB_BF16, "ssb_weight_shadow: only the bf16 storage type needs a copy (fp32 kernels read the master arena)");
  const size_t n8 = n / 8;
  long long blocks = ceil_div_ll((long long)n8, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ssb_launch(weight_shadow_kernel, dim3((int)blocks), dim3(256), 0, to_stream(stream), src, (bf16*)dst, n8);
  SSB_LAUNCH_CHECK("ssb_weight_shadow");
  return SSB_OK;
}

static int check_stem(const char* who, int Cl, int L, const ssb_geom& g) {
  SSB_REQUIRE(Cl >= 1 && Cl <= 16, "%s: num_leads %d out of range [1,16]", who, Cl);
  SSB_REQUIRE(L >= 1 && g.B > 0, "%s: empty input", who);
  SSB_REQUIRE(g.C % 8 == 0 && g.C > 0 && g.C <= 256, "%s: stem channels %d must be a multiple of 8 and <= 256", who, g.C);
  SSB_REQUIRE(g.len == (L - 1) / 2 + 1, "%s: len_out %d != floor((L-1)/2)+1 for L=%d", who, g.len, L);
  SSB_REQUIRE(g.pitch >= g.len + 2, "%s: pitch too small", who);
  return SSB_OK;
}

int ssb_stem_conv_fwd(const float* x, const float* w, void* y, int Cl, int L, ssb_geom g, int dtype,
                      ssb_stream_t stream) {
  int rc = check_stem("ssb_stem_conv_fwd", Cl, L, g);
  if (rc) return rc;
  SSB_REQUIRE(x && w && y, "ssb_stem_conv_fwd: null pointer");
  if (dtype == SSB_BF16) {   // multi-lead stems: implicit GEMM on the tensor cores (conv_sm100.cu: stem_tn_kernel)
    const int tc = ssb_stem_conv_fwd_sm100(x, w, y, Cl, L, g, nullptr, to_stream(stream));
    if (tc <= 0) return tc;
  }
  // tile of positions per block: every thread owns ST_PT positions x 4 channels
  const int cog = g.C / 4;
  int tt = (ST_THREADS / cog) * ST_PT;
  if (tt < ST_PT) tt = ST_PT;
  const size_t smem = ((size_t)Cl * 7 * g.C + (size_t)Cl * (2 * tt + 5)) * sizeof(float);
  SSB_REQUIRE(cog <= ST_THREADS, "ssb_stem_conv_fwd: stem channels %d too wide", g.C);
  SSB_REQUIRE(smem <= 96 * 1024, "ssb_stem_conv_fwd: num_leads x stem_channels too large (%zu B of shared memory)", smem);
  dim3 grid(ceil_div(g.len, tt), g.B);
  SSB_DISPATCH_DTYPE(dtype, T, {
    ssb_launch(stem_conv_fwd_kernel<T>, dim3(grid), dim3(ST_THREADS), smem, to_stream(stream), x, w, (T*)y, Cl, L, g, tt);
  })
  SSB_LAUNCH_CHECK("ssb_stem_conv_fwd");
  return SSB_OK;
}

int ssb_bn_stats(const void* x, ssb_geom g, double* sums, int dtype, ssb_stream_t stream);

int ssb_stem_conv_fwd_stats(const float* x, const float* w, void* y, int Cl, int L, ssb_geom g, double* sums, int dtype,
                            ssb_stream_t stream) {
  int rc = check_stem("ssb_stem_conv_fwd_stats", Cl, L, g);
  if (rc) return rc;
  SSB_REQUIRE(x && w && y && sums, "ssb_stem_conv_fwd_stats: null pointer");
  if (dtype == SSB_BF16) {   // multi-lead stems: tensor-core kernel with the statistics in its epilogue
    const int tc = ssb_stem_conv_fwd_sm100(x, w, y, Cl, L, g, sums, to_stream(stream));
    if (tc <= 0) return tc;
  }
  rc = ssb_stem_conv_fwd(x, w, y, Cl, L, g, dtype, stream);
  if (rc) return rc;
  return ssb_bn_stats(y, g, sums, dtype, stream);
}

int ssb_stem_conv_wgrad(const float* x, const void* dy, float* dw, int Cl, int L, ssb_geom g, int dtype,
                        ssb_stream_t stream) {
  int rc = check_stem("ssb_stem_conv_wgrad", Cl, L, g);
  if (rc) return rc;
  SSB_REQUIRE(x && dy && dw, "ssb_stem_conv_wgrad: null pointer");
  if (dtype == SSB_BF16) {   // multi-lead stems: tensor cores (conv_sm100.cu: stem_wgrad_tn_kernel)
    const int tc = ssb_stem_conv_wgrad_sm100(x, dy, dw, Cl, L, g, to_stream(stream));
    if (tc <= 0) return tc;
  }
  {   // single lead: register-tile kernel streaming dy from L2 (two leads were measured slower than the tiled kernel)
    const int V = dtype == SSB_BF16 ? 8 : 4;
    const int ncg = g.C / V;
    static const bool direct_on = !(getenv("SSB_STEM_WGRAD_DIRECT") && atoi(getenv("SSB_STEM_WGRAD_DIRECT")) == 0);
    if (direct_on && Cl == 1 && (long long)g.B * g.len < (1ll << 30) && g.C % V == 0 && ncg >= 1 && ncg <= 32 && (ncg & (ncg - 1)) == 0 &&
        (size_t)(SD_THREADS / 32 + 1) * 7 * g.C * sizeof(float) <= 48 * 1024) {
      const int plane = SD_THREADS / ncg;
      const long long P = (long long)g.B * g.len;
      static const int ppl = getenv("SSB_STEM_WGRAD_PPL") ? atoi(getenv("SSB_STEM_WGRAD_PPL")) : SD_POS_PER_LANE;
      long long nblk = ceil_div_ll(P, (long long)plane * (ppl > 0 ? ppl : SD_POS_PER_LANE));
      if (nblk > 144) nblk = 144;
      if (nblk < 1) nblk = 1;
      nblk = ceil_div_ll(nblk, SD_CLUSTER) * SD_CLUSTER;          // whole clusters (blocks past the last position idle)
      const int chunk = (int)ceil_div_ll(P, nblk);
      const size_t sm = (size_t)(SD_THREADS / 32 + 1) * 7 * g.C * sizeof(float);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)nblk);
      cfg.blockDim = dim3(SD_THREADS);
      cfg.dynamicSmemBytes = sm;
      cfg.stream = to_stream(stream);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = SD_CLUSTER;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      SSB_DISPATCH_DTYPE(dtype, T, {
        cudaLaunchKernelEx(&cfg, stem_conv_wgrad_direct_kernel<T, 1>, x, (const T*)dy, dw, L, g, chunk);
      })
      SSB_LAUNCH_CHECK("ssb_stem_conv_wgrad");
      return SSB_OK;
    }
  }
  const size_t smem = ((size_t)SW_TT * g.C + (size_t)Cl * (2 * SW_TT + 5) + (size_t)Cl * 7 * g.C) * sizeof(float);
  SSB_REQUIRE(smem <= 96 * 1024, "ssb_stem_conv_wgrad: num_leads x stem_channels too large (%zu B of shared memory)", smem);
  const int nown = (g.C / 4) * Cl;    // (4-channel group, lead) owners; the rest of the block slices the positions
  SSB_REQUIRE(nown <= SW_THREADS, "ssb_stem_conv_wgrad: num_leads x stem_channels / 4 = %d exceeds the block", nown);
  int ts = SW_THREADS / nown;
  if (ts > SW_TT) ts = SW_TT;
  const int ntt = ceil_div(g.len, SW_TT);
  const int ntiles = ntt * g.B;
  dim3 grid(ntiles < 148 ? ntiles : 148);
  SSB_DISPATCH_DTYPE(dtype, T, {
    ssb_launch(stem_conv_wgrad_kernel<T>, dim3(grid), dim3(SW_THREADS), smem, to_stream(stream), x, (const T*)dy, dw, Cl, L, g, ntt, ts);
  })
  SSB_LAUNCH_CHECK("ssb_stem_conv_wgrad");
  return SSB_OK;
}

}  // extern "C"

SSB_TRACE_DEFINE(conv_simt)
