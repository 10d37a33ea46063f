// FCN head tail (dropout + 1x1 classifier), linear upsample, pseudo-labels and the fused
// upsample+softmax+threshold+argmax+mask+loss+gradient kernel.
// Replaces ATen native_dropout / cuDNN 1x1 conv (fcn_head.py:94-96), upsample_linear1d
// (+backward, encoder_decoder.py:102-107) and the ~20 elementwise/reduction kernels of
// fixmatch.py:89-91,105,114-118 / mean_teacher.py:92,106,115-117.
#include "common.cuh"

#define MAX_CLS 8

// ATen area_pixel_compute_source_index, fp32 like torch for float tensors
__device__ __forceinline__ void lerp_index(int t, float scale, int Lin, int align_corners, int& i0, int& i1, float& w0,
                                           float& w1) {
  float src;
  if (align_corners) {
    src = scale * (float)t;
  } else {
    src = scale * ((float)t + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
  }
  i0 = (int)src;
  if (i0 > Lin - 1) i0 = Lin - 1;
  i1 = i0 + (i0 < Lin - 1 ? 1 : 0);
  w1 = src - (float)i0;
  w0 = 1.0f - w1;
}
static inline float lerp_scale(int Lin, int Lout, int align_corners) {
  if (align_corners) return Lout > 1 ? (float)(Lin - 1) / (float)(Lout - 1) : 0.f;
  return (float)Lin / (float)Lout;
}

// ---------------------------------------------------------------------------------------
// head classifier forward: one warp per valid row
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
head_cls_fwd_kernel(const T* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ low, ssb_geom g, int ncls, float p, const uint8_t* __restrict__ dmask,
                    const ssb_step_params* __restrict__ sp) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nrows = g.B * g.len;
  if (warp >= nrows) return;
  const int b = warp / g.len, t = warp - b * g.len;
  const T* ar = a + ((size_t)b * g.pitch + 1 + t) * g.C;
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  uint32_t seed = 0, step = 0;
  const bool use_rng = (p > 0.f) && !dmask && sp;
  if (use_rng) { seed = sp->rng_seed; step = sp->rng_step; }
  float acc[MAX_CLS];
#pragma unroll
  for (int k = 0; k < MAX_CLS; ++k) acc[k] = 0.f;
  for (int c = lane; c < g.C; c += 32) {
    float v = to_f(ar[c]);
    const uint32_t idx = (uint32_t)warp * (uint32_t)g.C + (uint32_t)c;
    if (dmask) v = dmask[idx] ? v * keep_scale : 0.f;
    else if (use_rng) v = dropout_keep(seed, step, idx, p) ? v * keep_scale : 0.f;
#pragma unroll
    for (int k = 0; k < MAX_CLS; ++k)
      if (k < ncls) acc[k] = fmaf(v, w[k * g.C + c], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < MAX_CLS; ++k)
    if (k < ncls) acc[k] = warp_sum(acc[k]);
  if (lane == 0)
    for (int k = 0; k < ncls; ++k) low[(size_t)warp * ncls + k] = acc[k] + bias[k];
}

// head classifier backward: da (flat padded, zero halos), dW[k][c] +=, dbias[k] +=
template <typename T>
__global__ void __launch_bounds__(256)
head_cls_bwd_kernel(const float* __restrict__ dlow, const T* __restrict__ a, const float* __restrict__ w,
                    T* __restrict__ da, float* __restrict__ dw, float* __restrict__ dbias, ssb_geom g, int ncls,
                    float p, const uint8_t* __restrict__ dmask, const ssb_step_params* __restrict__ sp, int rpb) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];  // [ncls][C] dW partials + [ncls] dbias partials
  const int C = g.C;
  float* sW = sm;
  float* sB = sm + ncls * C;
  for (int i = threadIdx.x; i < ncls * C + ncls; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int cb = C < 256 ? C : 256;
  const int nrl = 256 / cb;
  const int rl = threadIdx.x / cb;
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  uint32_t seed = 0, step = 0;
  const bool use_rng = (p > 0.f) && !dmask && sp;
  if (use_rng) { seed = sp->rng_seed; step = sp->rng_step; }
  const int rows = g.B * g.pitch;
  const int r0 = blockIdx.x * rpb, r1 = min(rows, r0 + rpb);
  if (rl < nrl) {
    for (int c = threadIdx.x % cb; c < C; c += cb) {
      float wk[MAX_CLS], accw[MAX_CLS], accb[MAX_CLS];
#pragma unroll
      for (int k = 0; k < MAX_CLS; ++k) {
        wk[k] = k < ncls ? w[k * C + c] : 0.f;
        accw[k] = 0.f;
        accb[k] = 0.f;
      }
      for (int r = r0 + rl; r < r1; r += nrl) {
        const int b = r / g.pitch, pos = r - b * g.pitch;
        float o = 0.f;
        if (pos >= 1 && pos <= g.len) {
          const int cr = b * g.len + pos - 1;  // compact row
          const uint32_t idx = (uint32_t)cr * (uint32_t)C + (uint32_t)c;
          float ks = keep_scale;
          if (dmask) ks = dmask[idx] ? keep_scale : 0.f;
          else if (use_rng) ks = dropout_keep(seed, step, idx, p) ? keep_scale : 0.f;
          const float av = to_f(a[(size_t)r * C + c]) * ks;
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < MAX_CLS; ++k)
            if (k < ncls) {
              const float d = dlow[(size_t)cr * ncls + k];
              s = fmaf(d, wk[k], s);
              accw[k] = fmaf(d, av, accw[k]);
              if (c == 0) accb[k] += d;
            }
          o = s * ks;
        }
        da[(size_t)r * C + c] = from_f<T>(o);
      }
#pragma unroll
      for (int k = 0; k < MAX_CLS; ++k)
        if (k < ncls) {
          atomicAdd(&sW[k * C + c], accw[k]);
          if (c == 0) atomicAdd(&sB[k], accb[k]);
        }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncls * C; i += blockDim.x) atomicAdd(&dw[i], sW[i]);
  for (int i = threadIdx.x; i < ncls; i += blockDim.x) atomicAdd(&dbias[i], sB[i]);
}

// ---------------------------------------------------------------------------------------
// linear upsample, NLC low-res -> NCL full-res (reference logits layout)
// ---------------------------------------------------------------------------------------
__global__ void upsample_fwd_kernel(const float* __restrict__ low, float* __restrict__ out, int B, int Lin, int Lout,
                                    int ncls, float scale, int align_corners) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)B * Lout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / Lout), t = (int)(idx - (long long)b * Lout);
    int i0, i1;
    float w0, w1;
    lerp_index(t, scale, Lin, align_corners, i0, i1, w0, w1);
    const float* p0 = low + ((size_t)b * Lin + i0) * ncls;
    const float* p1 = low + ((size_t)b * Lin + i1) * ncls;
    for (int k = 0; k < ncls; ++k) out[((size_t)b * ncls + k) * Lout + t] = w0 * p0[k] + w1 * p1[k];
  }
}

__global__ void upsample_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dlow, int B, int Lin, int Lout,
                                    int ncls, float scale, int align_corners) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)B * Lin * ncls;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % ncls);
    const int i = (int)((idx / ncls) % Lin);
    const int b = (int)(idx / ((long long)ncls * Lin));
    // candidate range of t whose taps can touch i (generous margin, exact test inside)
    int tlo, thi;
    if (scale > 0.f) {
      tlo = (int)floorf(((float)i - 1.0f) / scale - 0.5f) - 2;
      thi = (int)ceilf(((float)i + 1.5f) / scale - 0.5f) + 2;
    } else {
      tlo = 0;
      thi = Lout - 1;
    }
    tlo = max(tlo, 0);
    thi = min(thi, Lout - 1);
    const float* d = dout + ((size_t)b * ncls + k) * Lout;
    float acc = 0.f;
    for (int t = tlo; t <= thi; ++t) {
      int i0, i1;
      float w0, w1;
      lerp_index(t, scale, Lin, align_corners, i0, i1, w0, w1);
      const float gv = d[t];
      if (i0 == i) acc = fmaf(w0, gv, acc);
      if (i1 == i) acc = fmaf(w1, gv, acc);
    }
    dlow[idx] = acc;
  }
}

// ---------------------------------------------------------------------------------------
// softmax statistics of ncls logits, replicating torch's cunn_SpatialSoftMaxForward
// operation order: max (sequential), sum += exp(z - max) (sequential), p = exp(z-max)/sum
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void softmax_conf_label(const float* z, int ncls, float& conf, int& label) {
  float m = -3.402823466e+38f;  // numeric_limits<float>::lowest()
  for (int k = 0; k < ncls; ++k) m = (m < z[k]) ? z[k] : m;
  float sum = 0.f;
  for (int k = 0; k < ncls; ++k) sum += expf(z[k] - m);
  // max over the soft-max outputs; NaN propagates like torch.max
  float best = expf(z[0] - m) / sum;
  bool nan = best != best;
  for (int k = 1; k < ncls; ++k) {
    const float pk = expf(z[k] - m) / sum;
    if (pk != pk) nan = true;
    if (pk > best) best = pk;
  }
  conf = nan ? __int_as_float(0x7fc00000) : best;
  // torch.argmax: first maximal value; NaN counts as maximal
  int arg = 0;
  float bz = z[0];
  for (int k = 1; k < ncls; ++k) {
    const float v = z[k];
    if (!(bz != bz) && (v > bz || v != v)) { bz = v; arg = k; }
  }
  label = arg;
}

__global__ void pseudo_label_kernel(const float* __restrict__ logits, float thr, float* __restrict__ conf,
                                    int64_t* __restrict__ label, uint8_t* __restrict__ mask, int U, int ncls, int L) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)U * L;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(idx / L), t = (int)(idx - (long long)u * L);
    float z[MAX_CLS];
    for (int k = 0; k < ncls; ++k) z[k] = logits[((size_t)u * ncls + k) * L + t];
    float c;
    int lab;
    softmax_conf_label(z, ncls, c, lab);
    if (conf) conf[idx] = c;
    if (label) label[idx] = (int64_t)lab;
    if (mask) mask[idx] = (c >= thr) ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------
// fused semi-supervised loss.  One block per (sample, chunk of NI low-res positions).
// phase 1: every full-res position t whose taps touch the chunk computes its softmax,
//          loss term and gradient g_t (scaled), staged in shared memory;
// phase 2: thread (i, k) gathers w0*g_t / w1*g_t in t order -> deterministic dlow.
// ---------------------------------------------------------------------------------------
#define SL_NI 8
#define SL_THREADS 256

__global__ void __launch_bounds__(SL_THREADS)
semi_loss_kernel(const float* __restrict__ low_s, const int64_t* __restrict__ target, const float* __restrict__ low_t,
                 float* __restrict__ dlow, double* __restrict__ sums, int Bl, int Bu, int Lin, int L, int ncls, int mode,
                 float thr_arg, const ssb_step_params* __restrict__ sp, float scale, int align_corners,
                 float* __restrict__ conf_out, int64_t* __restrict__ label_out, uint8_t* __restrict__ mask_out,
                 int tcap) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* sG = sm;                                  // [tcap][ncls]
  float* sW1 = sm + (size_t)tcap * ncls;           // [tcap]
  int* sI0 = (int*)(sW1 + tcap);                   // [tcap]
  __shared__ float sRed[3][SL_THREADS / 32];
  const int b = blockIdx.y;
  const int ia = blockIdx.x * SL_NI;
  const int ib = min(Lin, ia + SL_NI);
  const float thr = sp ? sp->conf_thresh : thr_arg;
  int tlo, thi;
  if (scale > 0.f) {
    tlo = (int)floorf(((float)ia - 1.0f) / scale - 0.5f) - 2;
    thi = (int)ceilf(((float)ib + 0.5f) / scale - 0.5f) + 2;
  } else {
    tlo = 0;
    thi = L - 1;
  }
  tlo = max(tlo, 0);
  thi = min(thi, L - 1);
  const int nt = min(thi - tlo + 1, tcap);
  const bool labeled = b < Bl;
  const float cx = (mode == SSB_LOSS_SUP ? 1.0f : 0.5f) / ((float)Bl * (float)L);
  const float cu = Bu > 0 ? 0.5f / ((float)Bu * (float)L) : 0.f;
  // per-block partial sums in fp32 (a few hundred O(1) terms); only the cross-block accumulation is fp64
  float acc_x = 0.f, acc_u = 0.f, acc_m = 0.f;
  for (int tt = threadIdx.x; tt < nt; tt += SL_THREADS) {
    const int t = tlo + tt;
    int i0, i1;
    float w0, w1;
    lerp_index(t, scale, Lin, align_corners, i0, i1, w0, w1);
    sI0[tt] = i0;
    sW1[tt] = w1;
    const bool touches = (i0 >= ia && i0 < ib) || (i1 >= ia && i1 < ib);
    float g[MAX_CLS];
#pragma unroll
    for (int k = 0; k < MAX_CLS; ++k) g[k] = 0.f;
    if (touches) {
      const bool owner = i0 >= ia && i0 < ib;
      const float* p0 = low_s + ((size_t)b * Lin + i0) * ncls;
      const float* p1 = low_s + ((size_t)b * Lin + i1) * ncls;
      float z[MAX_CLS], pr[MAX_CLS];
      float m = -3.402823466e+38f;
      for (int k = 0; k < ncls; ++k) {
        z[k] = w0 * p0[k] + w1 * p1[k];
        m = fmaxf(m, z[k]);
      }
      float sum = 0.f;
      for (int k = 0; k < ncls; ++k) {
        pr[k] = expf(z[k] - m);
        sum += pr[k];
      }
      const float inv = 1.0f / sum;
      const float lse = m + logf(sum);
      if (labeled) {
        const int y = (int)target[(size_t)b * L + t];
        if (y >= 0 && y < ncls) {
          if (owner) acc_x += lse - z[y];
          for (int k = 0; k < ncls; ++k) g[k] = (pr[k] * inv - (k == y ? 1.f : 0.f)) * cx;
        }
      } else if (mode != SSB_LOSS_SUP) {
        const int u = b - Bl;
        const float* q0 = low_t + ((size_t)u * Lin + i0) * ncls;
        const float* q1 = low_t + ((size_t)u * Lin + i1) * ncls;
        float zt[MAX_CLS];
        for (int k = 0; k < ncls; ++k) zt[k] = w0 * q0[k] + w1 * q1[k];
        if (mode == SSB_LOSS_FIXMATCH) {
          float c;
          int lab;
          softmax_conf_label(zt, ncls, c, lab);
          const bool mk = c >= thr;
          if (owner) {
            if (mk) { acc_u += lse - z[lab]; acc_m += 1.0f; }
            const size_t o = (size_t)u * L + t;
            if (conf_out) conf_out[o] = c;
            if (label_out) label_out[o] = (int64_t)lab;
            if (mask_out) mask_out[o] = mk ? 1 : 0;
          }
          if (mk)
            for (int k = 0; k < ncls; ++k) g[k] = (pr[k] * inv - (k == lab ? 1.f : 0.f)) * cu;
        } else {
          float mt = -3.402823466e+38f;
          for (int k = 0; k < ncls; ++k) mt = fmaxf(mt, zt[k]);
          float st = 0.f, q[MAX_CLS];
          for (int k = 0; k < ncls; ++k) { q[k] = expf(zt[k] - mt); st += q[k]; }
          const float it = 1.0f / st;
          bool mk = true;
          if (mode == SSB_LOSS_SOFT_MASKED) {   // loss_u * (conf_u_w >= conf_thresh), conf = softmax(1).max(1)[0] (reco.py:226,249)
            float c;
            int lab;
            softmax_conf_label(zt, ncls, c, lab);
            mk = c >= thr;
            if (owner) {
              if (mk) acc_m += 1.0f;
              const size_t o = (size_t)u * L + t;
              if (conf_out) conf_out[o] = c;
              if (mask_out) mask_out[o] = mk ? 1 : 0;
            }
          }
          float l = 0.f;
          for (int k = 0; k < ncls; ++k) {
            q[k] *= it;
            l += q[k] * (lse - z[k]);
            g[k] = mk ? (pr[k] * inv - q[k]) * cu : 0.f;
          }
          if (owner && mk) acc_u += l;
        }
      }
    }
    for (int k = 0; k < ncls; ++k) sG[(size_t)tt * ncls + k] = g[k];
  }
  __syncthreads();
  // phase 2: deterministic gather
  for (int o = threadIdx.x; o < (ib - ia) * ncls; o += SL_THREADS) {
    const int i = ia + o / ncls, k = o % ncls;
    float acc = 0.f;
    // only positions whose source coordinate lies in (i-1, i+1) can touch low-res index i
    int ta = 0, tb = nt;
    if (scale > 0.f && !align_corners) {
      ta = max(0, (int)floorf(((float)i - 0.5f) / scale - 0.5f) - 2 - tlo);
      tb = min(nt, (int)ceilf(((float)i + 1.5f) / scale - 0.5f) + 3 - tlo);
    }
    for (int tt = ta; tt < tb; ++tt) {
      const int i0 = sI0[tt];
      const int i1 = i0 + (i0 < Lin - 1 ? 1 : 0);
      const float w1 = sW1[tt];
      const float gv = sG[(size_t)tt * ncls + k];
      if (i0 == i) acc = fmaf(1.0f - w1, gv, acc);
      if (i1 == i) acc = fmaf(w1, gv, acc);
    }
    dlow[((size_t)b * Lin + i) * ncls + k] = acc;
  }
  // loss partial sums
  acc_x = warp_sum(acc_x);
  acc_u = warp_sum(acc_u);
  acc_m = warp_sum(acc_m);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sRed[0][wid] = acc_x; sRed[1][wid] = acc_u; sRed[2][wid] = acc_m; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < SL_THREADS / 32; ++i) s += sRed[threadIdx.x][i];
    if (s != 0.f) atomicAdd(&sums[threadIdx.x], (double)s);
  }
}

// ---------------------------------------------------------------------------------------
// evaluation tail (base.py:184-245): linear upsample + softmax + argmax + CE sum + per-sample class counts in one
// pass over the low-res logits.  One block per (chunk of EV_CHUNK positions, sample).
// ---------------------------------------------------------------------------------------
#define EV_THREADS 256
#define EV_CHUNK 1024

__global__ void __launch_bounds__(EV_THREADS)
eval_metrics_kernel(const float* __restrict__ low, const int64_t* __restrict__ target, double* __restrict__ sums,
                    int* __restrict__ counts, float* __restrict__ probs, int64_t* __restrict__ pred, int B, int Lin, int L,
                    int ncls, float scale, int align_corners) {
  pdl_trigger();
  pdl_wait();
  __shared__ int sCnt[3][MAX_CLS];
  __shared__ float sRed[2][EV_THREADS / 32];
  const int b = blockIdx.y;
  if (threadIdx.x < 3 * MAX_CLS) sCnt[threadIdx.x / MAX_CLS][threadIdx.x % MAX_CLS] = 0;
  __syncthreads();
  const int t0 = blockIdx.x * EV_CHUNK;
  const int t1 = min(L, t0 + EV_CHUNK);
  float acc = 0.f, nvalid = 0.f;
  for (int t = t0 + threadIdx.x; t < t1; t += EV_THREADS) {
    int i0, i1;
    float w0, w1;
    lerp_index(t, scale, Lin, align_corners, i0, i1, w0, w1);
    const float* p0 = low + ((size_t)b * Lin + i0) * ncls;
    const float* p1 = low + ((size_t)b * Lin + i1) * ncls;
    float z[MAX_CLS], pr[MAX_CLS];
    float m = -3.402823466e+38f;
    for (int k = 0; k < ncls; ++k) {
      z[k] = w0 * p0[k] + w1 * p1[k];
      m = fmaxf(m, z[k]);
    }
    float sum = 0.f;
    for (int k = 0; k < ncls; ++k) {
      pr[k] = expf(z[k] - m);
      sum += pr[k];
    }
    // prediction = argmax of the soft-max OUTPUTS (base.py:205-209), first maximum
    int arg = 0;
    float best = pr[0] / sum;
    if (probs) probs[((size_t)b * ncls) * L + t] = best;
    for (int k = 1; k < ncls; ++k) {
      const float pk = pr[k] / sum;
      if (probs) probs[((size_t)b * ncls + k) * L + t] = pk;
      if (pk > best) { best = pk; arg = k; }
    }
    if (pred) pred[(size_t)b * L + t] = (int64_t)arg;
    const int y = (int)target[(size_t)b * L + t];
    atomicAdd(&sCnt[1][arg], 1);
    if (y >= 0 && y < ncls) {
      acc += m + logf(sum) - z[y];
      nvalid += 1.f;
      atomicAdd(&sCnt[2][y], 1);
      if (y == arg) atomicAdd(&sCnt[0][y], 1);
    }
  }
  acc = warp_sum(acc);
  nvalid = warp_sum(nvalid);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sRed[0][wid] = acc; sRed[1][wid] = nvalid; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s_ = 0.f;
    for (int i = 0; i < EV_THREADS / 32; ++i) s_ += sRed[threadIdx.x][i];
    if (s_ != 0.f) atomicAdd(&sums[threadIdx.x], (double)s_);
  }
  if (threadIdx.x < 3 * ncls) {
    const int q = threadIdx.x / ncls, c = threadIdx.x % ncls;
    const int v = sCnt[q][c];
    if (v) atomicAdd(&counts[((size_t)b * ncls + c) * 3 + q], v);
  }
}

int ssb_loss_prepare() {
  cudaError_t e = cudaFuncSetAttribute(semi_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  if (e != cudaSuccess) {
    ssb_set_error("ssb_loss_prepare: %s", cudaGetErrorString(e));
    return SSB_ERR_CUDA;
  }
  return SSB_OK;
}

// ---------------------------------------------------------------------------------------
extern "C" {

int ssb_head_cls_fwd(const void* a, const float* w, const float* bias, float* low, ssb_geom g, int ncls, float p,
                     const uint8_t* drop_mask, const ssb_step_params* sp, int dtype, ssb_stream_t stream) {
  SSB_REQUIRE(a && w && bias && low, "ssb_head_cls_fwd: null pointer");
  SSB_REQUIRE(ncls >= 1 && ncls <= MAX_CLS, "ssb_head_cls_fwd: num_classes %d out of range [1,%d]", ncls, MAX_CLS);
  SSB_REQUIRE(g.B > 0 && g.len > 0 && g.C > 0 && g.pitch >= g.len + 2, "ssb_head_cls_fwd: bad geometry");
  SSB_REQUIRE(p >= 0.f && p < 1.f, "ssb_head_cls_fwd: dropout p=%f out of [0,1)", p);
  const int nrows = g.B * g.len;
  SSB_DISPATCH_DTYPE(dtype, T, {
    ssb_launch(head_cls_fwd_kernel<T>, dim3(ceil_div(nrows, 8)), dim3(256), 0, to_stream(stream), (const T*)a, w, bias, low, g, ncls, p, drop_mask, sp);
  })
  SSB_LAUNCH_CHECK("ssb_head_cls_fwd");
  return SSB_OK;
}

int ssb_head_cls_bwd(const float* dlow, const void* a, const float* w, void* da, float* dw, float* dbias, ssb_geom g,
                     int ncls, float p, const uint8_t* drop_mask, const ssb_step_params* sp, int dtype,
                     ssb_stream_t stream) {
  SSB_REQUIRE(dlow && a && w && da && dw && dbias, "ssb_head_cls_bwd: null pointer");
  SSB_REQUIRE(ncls >= 1 && ncls <= MAX_CLS, "ssb_head_cls_bwd: num_classes %d out of range [1,%d]", ncls, MAX_CLS);
  SSB_REQUIRE(g.B > 0 && g.len > 0 && g.C > 0 && g.pitch >= g.len + 2, "ssb_head_cls_bwd: bad geometry");
  const int rows = g.B * g.pitch;
  const int cb = g.C < 256 ? g.C : 256;
  const int nrl = 256 / cb;
  int rpb = ceil_div(rows, 148 * 2);
  if (rpb < nrl * 2) rpb = nrl * 2;
  const size_t smem = ((size_t)ncls * g.C + ncls) * sizeof(float);
  SSB_REQUIRE(smem <= 48 * 1024, "ssb_head_cls_bwd: head channels %d too large", g.C);
  SSB_DISPATCH_DTYPE(dtype, T, {
    ssb_launch(head_cls_bwd_kernel<T>, dim3(ceil_div(rows, rpb)), dim3(256), smem, to_stream(stream), dlow, (const T*)a, w, (T*)da, dw, dbias, g, ncls, p, drop_mask, sp, rpb);
  })
  SSB_LAUNCH_CHECK("ssb_head_cls_bwd");
  return SSB_OK;
}

int ssb_upsample_fwd(const float* low, float* out, int B, int Lin, int Lout, int ncls, int align_corners,
                     ssb_stream_t stream) {
  SSB_REQUIRE(low && out && B > 0 && Lin > 0 && Lout > 0 && ncls > 0, "ssb_upsample_fwd: bad arguments");
  const long long total = (long long)B * Lout;
  int blocks = (int)ceil_div_ll(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ssb_launch(upsample_fwd_kernel, dim3(blocks), dim3(256), 0, to_stream(stream), low, out, B, Lin, Lout, ncls, lerp_scale(Lin, Lout, align_corners), align_corners);
  SSB_LAUNCH_CHECK("ssb_upsample_fwd");
  return SSB_OK;
}

int ssb_upsample_bwd(const float* dout, float* dlow, int B, int Lin, int Lout, int ncls, int align_corners,
                     ssb_stream_t stream) {
  SSB_REQUIRE(dout && dlow && B > 0 && Lin > 0 && Lout > 0 && ncls > 0, "ssb_upsample_bwd: bad arguments");
  const long long total = (long long)B * Lin * ncls;
  int blocks = (int)ceil_div_ll(total, 128);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ssb_launch(upsample_bwd_kernel, dim3(blocks), dim3(128), 0, to_stream(stream), dout, dlow, B, Lin, Lout, ncls, lerp_scale(Lin, Lout, align_corners), align_corners);
  SSB_LAUNCH_CHECK("ssb_upsample_bwd");
  return SSB_OK;
}

int ssb_pseudo_label(const float* logits, float thr, float* conf, int64_t* label, uint8_t* mask, int U, int ncls,
                     int L, ssb_stream_t stream) {
  SSB_REQUIRE(logits && U > 0 && L > 0, "ssb_pseudo_label: bad arguments");
  SSB_REQUIRE(ncls >= 1 && ncls <= MAX_CLS, "ssb_pseudo_label: num_classes %d out of range [1,%d]", ncls, MAX_CLS);
  const long long total = (long long)U * L;
  int blocks = (int)ceil_div_ll(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ssb_launch(pseudo_label_kernel, dim3(blocks), dim3(256), 0, to_stream(stream), logits, thr, conf, label, mask, U, ncls, L);
  SSB_LAUNCH_CHECK("ssb_pseudo_label");
  return SSB_OK;
}

int ssb_eval_metrics(const float* low, const int64_t* target, double* sums, int32_t* counts, float* probs,
                     int64_t* pred, int B, int Lin, int L, int ncls, int align_corners, ssb_stream_t stream) {
  SSB_REQUIRE(low && target && sums && counts, "ssb_eval_metrics: null pointer");
  SSB_REQUIRE(B > 0 && Lin > 0 && L > 0, "ssb_eval_metrics: bad sizes (B=%d Lin=%d L=%d)", B, Lin, L);
  SSB_REQUIRE(ncls >= 1 && ncls <= MAX_CLS, "ssb_eval_metrics: num_classes %d out of range [1,%d]", ncls, MAX_CLS);
  SSB_REQUIRE(B <= 65535, "ssb_eval_metrics: batch %d exceeds the grid limit", B);
  const float scale = lerp_scale(Lin, L, align_corners);
  dim3 grid((L + EV_CHUNK - 1) / EV_CHUNK, B);
  ssb_launch(eval_metrics_kernel, grid, dim3(EV_THREADS), 0, to_stream(stream), low, target, sums, (int*)counts, probs, pred, B,
             Lin, L, ncls, scale, align_corners);
  SSB_LAUNCH_CHECK("ssb_eval_metrics");
  return SSB_OK;
}

int ssb_semi_loss(const float* low_s, const int64_t* target, const float* low_t, float* dlow, double* sums, int Bl,
                  int Bu, int Lin, int L, int ncls, int mode, float thr, const ssb_step_params* sp, int align_corners,
                  float* conf, int64_t* label, uint8_t* mask, ssb_stream_t stream) {
  SSB_REQUIRE(low_s && target && dlow && sums, "ssb_semi_loss: null pointer");
  SSB_REQUIRE(Bl > 0 && Bu >= 0 && Lin > 0 && L > 0, "ssb_semi_loss: bad sizes (Bl=%d Bu=%d Lin=%d L=%d)", Bl, Bu, Lin, L);
  SSB_REQUIRE(ncls >= 1 && ncls <= MAX_CLS, "ssb_semi_loss: num_classes %d out of range [1,%d]", ncls, MAX_CLS);
  SSB_REQUIRE(mode == SSB_LOSS_SUP || mode == SSB_LOSS_FIXMATCH || mode == SSB_LOSS_SOFT || mode == SSB_LOSS_SOFT_MASKED,
              "ssb_semi_loss: bad mode %d", mode);
  SSB_REQUIRE(mode == SSB_LOSS_SUP || (low_t && Bu > 0), "ssb_semi_loss: teacher logits required for mode %d", mode);
  const float scale = lerp_scale(Lin, L, align_corners);
  int tcap;
  if (scale > 0.f) tcap = (int)ceilf((SL_NI + 2.0f) / scale) + 8;
  else tcap = L;
  if (tcap > L) tcap = L;
  const size_t smem = (size_t)tcap * (ncls + 2) * sizeof(float);
  SSB_REQUIRE(smem <= 160 * 1024, "ssb_semi_loss: upsample ratio too large for the staging buffer (%zu bytes)", smem);
  dim3 grid(ceil_div(Lin, SL_NI), Bl + Bu);
  ssb_launch(semi_loss_kernel, dim3(grid), dim3(SL_THREADS), smem, to_stream(stream), low_s, target, low_t, dlow, sums, Bl, Bu, Lin, L, ncls, mode, thr, sp, scale, align_corners, conf, label, mask, tcap);
  SSB_LAUNCH_CHECK("ssb_semi_loss");
  return SSB_OK;
}

}  // extern "C"

SSB_TRACE_DEFINE(head_loss)
