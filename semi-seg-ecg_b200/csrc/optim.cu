// Fused multi-tensor AdamW (+ EMA teacher) over flat parameter arenas, EMA of teacher BN
// buffers and the global gradient norm.  Replaces torch.optim.AdamW's per-tensor loop
// (utils/optimizer.py:22-34), the python EMA loops (mean_teacher.py:139-149; ~384 tiny
// kernels per step) and get_grad_norm_ (misc.py:265-278).  HBM-bound: 28 B/param
// (AdamW) or 36 B/param (AdamW+EMA), 128-bit loads/stores.
#include "common.cuh"

__global__ void __launch_bounds__(256)
adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 float* __restrict__ pe, size_t n4, float beta1, float omb1, float beta2, float omb2, float eps,
                 float wd, const ssb_step_params* __restrict__ sp, const float* __restrict__ gnorm, float max_norm) {
  pdl_trigger();
  pdl_wait();
  const float lr = sp->lr;
  const float step_size = lr * sp->inv_bias1;
  const float isb2 = sp->inv_sqrt_bias2;
  const float decay = 1.0f - lr * wd;
  float gs = sp->grad_scale;
  // torch.nn.utils.clip_grad_norm_ (misc.py:248-250): gradients times min(1, max_norm / (total_norm + 1e-6)); gnorm holds
  // the norm of the arena as stored, i.e. before grad_scale
  if (gnorm) gs *= fminf(1.0f, max_norm / (gnorm[0] * gs + 1e-6f));
  const float d = sp->ema_decay;
  const int first = sp->ema_first;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* e4 = reinterpret_cast<float4*>(pe);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    float P[4] = {pp.x, pp.y, pp.z, pp.w};
    const float G[4] = {gg.x * gs, gg.y * gs, gg.z * gs, gg.w * gs};
    float M[4] = {mm.x, mm.y, mm.z, mm.w};
    float V[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      P[j] *= decay;
      M[j] = beta1 * M[j] + omb1 * G[j];
      V[j] = beta2 * V[j] + omb2 * G[j] * G[j];
      const float denom = sqrtf(V[j]) * isb2 + eps;
      P[j] -= step_size * (M[j] / denom);
    }
    p4[i] = make_float4(P[0], P[1], P[2], P[3]);
    m4[i] = make_float4(M[0], M[1], M[2], M[3]);
    v4[i] = make_float4(V[0], V[1], V[2], V[3]);
    if (pe) {
      float E[4];
      if (first) {
        // teacher still aliases the student: k_old == q_new (mean_teacher.py:285-290)
#pragma unroll
        for (int j = 0; j < 4; ++j) E[j] = P[j] * d + P[j] * (1.0f - d);
      } else {
        const float4 ee = e4[i];
        const float EO[4] = {ee.x, ee.y, ee.z, ee.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) E[j] = EO[j] * d + P[j] * (1.0f - d);
      }
      e4[i] = make_float4(E[0], E[1], E[2], E[3]);
    }
  }
}

__global__ void ema_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n,
                           const ssb_step_params* __restrict__ sp) {
  pdl_trigger();
  pdl_wait();
  const float d = sp->ema_decay;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = dst[i] * d + src[i] * (1.0f - d);
}

__global__ void ema_i64_kernel(float* __restrict__ dst, const int64_t* __restrict__ src, size_t n,
                               const ssb_step_params* __restrict__ sp) {
  pdl_trigger();
  pdl_wait();
  const float d = sp->ema_decay;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = dst[i] * d + (float)src[i] * (1.0f - d);
}

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ ws) {
  pdl_trigger();
  pdl_wait();
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float x = g[i];
    acc += (double)x * (double)x;
  }
  acc = warp_sum_d(acc);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s[i];
    atomicAdd(ws, t);
  }
}
__global__ void grad_norm_final_kernel(const double* ws, float* out) {
  pdl_trigger();
  pdl_wait(); out[0] = (float)sqrt(ws[0]); }

extern "C" {

int ssb_adamw_ema(float* p, const float* g, float* m, float* v, float* p_ema, size_t n, double beta1, double beta2,
                  double eps, double weight_decay, const ssb_step_params* sp, ssb_stream_t stream) {
  return ssb_adamw_ema_clip(p, g, m, v, p_ema, n, beta1, beta2, eps, weight_decay, sp, nullptr, 0.0, stream);
}

int ssb_adamw_ema_clip(float* p, const float* g, float* m, float* v, float* p_ema, size_t n, double beta1, double beta2,
                       double eps, double weight_decay, const ssb_step_params* sp, const float* gnorm, double max_norm,
                       ssb_stream_t stream) {
  SSB_REQUIRE(p && g && m && v && sp, "ssb_adamw_ema: null pointer");
  SSB_REQUIRE(!gnorm || max_norm > 0.0, "ssb_adamw_ema_clip: max_norm must be positive (got %g)", max_norm);
  SSB_REQUIRE(n > 0 && n % 4 == 0, "ssb_adamw_ema: arena length %zu must be a positive multiple of 4", n);
  const size_t n4 = n / 4;
  long long blocks = ceil_div_ll((long long)n4, 256 * 2);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ssb_launch(adamw_ema_kernel, dim3((int)blocks), dim3(256), 0, to_stream(stream), p, g, m, v, p_ema, n4, (float)beta1, (float)(1.0 - beta1), (float)beta2,
                                                                (float)(1.0 - beta2), (float)eps, (float)weight_decay, sp, gnorm, (float)max_norm);
  SSB_LAUNCH_CHECK("ssb_adamw_ema");
  return SSB_OK;
}

int ssb_ema(float* dst, const float* src, size_t n, const ssb_step_params* sp, ssb_stream_t stream) {
  SSB_REQUIRE(dst && src && sp && n > 0, "ssb_ema: bad arguments");
  long long blocks = ceil_div_ll((long long)n, 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  ssb_launch(ema_kernel, dim3((int)blocks), dim3(256), 0, to_stream(stream), dst, src, n, sp);
  SSB_LAUNCH_CHECK("ssb_ema");
  return SSB_OK;
}

int ssb_ema_i64(float* dst, const int64_t* src, size_t n, const ssb_step_params* sp, ssb_stream_t stream) {
  SSB_REQUIRE(dst && src && sp && n > 0, "ssb_ema_i64: bad arguments");
  ssb_launch(ema_i64_kernel, dim3((int)ceil_div_ll((long long)n, 256)), dim3(256), 0, to_stream(stream), dst, src, n, sp);
  SSB_LAUNCH_CHECK("ssb_ema_i64");
  return SSB_OK;
}

int ssb_grad_norm(const float* g, size_t n, double* ws, float* out, ssb_stream_t stream) {
  SSB_REQUIRE(g && ws && out && n > 0, "ssb_grad_norm: bad arguments");
  long long blocks = ceil_div_ll((long long)n, 256 * 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  ssb_launch(grad_sumsq_kernel, dim3((int)blocks), dim3(256), 0, to_stream(stream), g, n, ws);
  SSB_LAUNCH_CHECK("ssb_grad_norm");
  ssb_launch(grad_norm_final_kernel, dim3(1), dim3(1), 0, to_stream(stream), ws, out);
  SSB_LAUNCH_CHECK("ssb_grad_norm(final)");
  return SSB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// SyncBatchNorm statistics exchange over NVLink peer memory (replaces one small NCCL all-reduce per BN
// layer and direction: torch SyncBatchNorm's all_gather of (mean, invstd, count) / all_reduce of the
// backward sums, fixmatch.py:290-291).  Every rank owns a symmetric mailbox; one block per rank
//   1. stores its slice into slot [parity][rank] of every peer's mailbox as self-validating 8-byte words
//      {32 payload bits, exchange number} (two per double) -- no fence, no separate flag round trip,
//   2. polls the words arriving in ITS OWN mailbox until they carry this exchange's number, and
//   3. replaces its slice by the sum of all ranks' slices in rank order (bitwise identical on every rank).
// Two parities: a rank can be at most one exchange ahead of a peer (it cannot finish exchange k+1 before the peer
// has sent k+1, which the peer does only after it has finished reading k), so slot reuse never races.
// ---------------------------------------------------------------------------------------
// (SBX_MAX_WORLD, SBX_HDR_BYTES, sbx_store / sbx_load: common.cuh -- shared with the in-kernel exchange of bn.cu)

__global__ void __launch_bounds__(256) syncbn_exchange_kernel(double* __restrict__ slice, int n, const unsigned long long* __restrict__ peers,
                                                              int world, int rank, int slot_doubles) {
  pdl_trigger();
  pdl_wait();
  __shared__ unsigned int s_epoch;
  char* mine = reinterpret_cast<char*>(peers[rank]);
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(mine);
    s_epoch = *counter + 1;
    *counter = s_epoch;
  }
  __syncthreads();
  const unsigned int epoch = s_epoch;
  const size_t slot_words = (size_t)slot_doubles * 2;                       // 8-byte words per slot
  const size_t par_off = (size_t)(epoch & 1u) * SBX_MAX_WORLD * slot_words;
  const uint2* mail = reinterpret_cast<const uint2*>(mine + SBX_HDR_BYTES) + par_off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double mineval = slice[i];
    const uint32_t lo = (uint32_t)__double2loint(mineval), hi = (uint32_t)__double2hiint(mineval);
    for (int r = 0; r < world; ++r) {                                       // 1. my value into every peer's slot [rank]
      if (r == rank) continue;
      uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<char*>(peers[r]) + SBX_HDR_BYTES) + par_off + (size_t)rank * slot_words + 2 * i;
      sbx_store(dst, lo, epoch);
      sbx_store(dst + 1, hi, epoch);
    }
    double t = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < world; ++r) {                                       // 2./3. the peers' values, summed in rank order
      if (r == rank) { t += mineval; continue; }
      const uint2* src = mail + (size_t)r * slot_words + 2 * i;
      uint2 a, b;
      do {
        a = sbx_load(src);
        b = sbx_load(src + 1);
        if (clock64() - t0 > 40000000000LL) {   // ~20 s: a peer is gone -- record it and go on rather than hang the GPU
          *reinterpret_cast<unsigned int*>(mine + 64) = epoch;
          break;
        }
      } while (a.y != epoch || b.y != epoch);
      t += __hiloint2double((int)b.x, (int)a.x);
    }
    slice[i] = t;
  }
}

extern "C" {

size_t ssb_syncbn_mailbox_bytes(int slot_doubles) {
  return (size_t)SBX_HDR_BYTES + (size_t)2 * SBX_MAX_WORLD * (size_t)slot_doubles * 16;   // two {payload, number} words per double
}

size_t ssb_syncbn_fused_mailbox_bytes(int slot_doubles, int world) {
  return (size_t)SBX_HDR_BYTES + (size_t)world * (size_t)slot_doubles * 16;   // two {payload, tag} words per double
}

int ssb_syncbn_exchange(double* slice, int n, const uint64_t* peers_dev, int world, int rank, int slot_doubles,
                        ssb_stream_t stream) {
  SSB_REQUIRE(slice && peers_dev, "ssb_syncbn_exchange: null pointer");
  SSB_REQUIRE(world >= 2 && world <= SBX_MAX_WORLD && rank >= 0 && rank < world, "ssb_syncbn_exchange: bad world %d / rank %d", world, rank);
  SSB_REQUIRE(n > 0 && n <= slot_doubles, "ssb_syncbn_exchange: slice of %d doubles does not fit the %d-double slots", n, slot_doubles);
  ssb_launch_pro(syncbn_exchange_kernel, dim3(1), dim3(256), 0, to_stream(stream), slice, n,
             reinterpret_cast<const unsigned long long*>(peers_dev), world, rank, slot_doubles);
  SSB_LAUNCH_CHECK("ssb_syncbn_exchange");
  return SSB_OK;
}

}  // extern "C"

SSB_TRACE_DEFINE(optim)
