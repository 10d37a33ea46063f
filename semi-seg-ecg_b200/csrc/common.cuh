// Shared helpers for the libsemiseg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/ssb.h"

typedef __nv_bfloat16 bf16;

void ssb_set_error(const char* fmt, ...);
extern std::atomic<long long> g_ssb_launches;

#define SSB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ssb_set_error(__VA_ARGS__);     \
      return SSB_ERR_INVALID;         \
    }                                 \
  } while (0)

// call after every kernel launch
#define SSB_LAUNCH_CHECK(name)                                              \
  do {                                                                      \
    g_ssb_launches.fetch_add(1, std::memory_order_relaxed);                 \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      ssb_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return SSB_ERR_CUDA;                                                  \
    }                                                                       \
  } while (0)

#define SSB_DISPATCH_DTYPE(dtype, T, ...)              \
  if ((dtype) == SSB_F32) {                            \
    typedef float T;                                   \
    __VA_ARGS__                                        \
  } else if ((dtype) == SSB_BF16) {                    \
    typedef bf16 T;                                    \
    __VA_ARGS__                                        \
  } else {                                             \
    ssb_set_error("unsupported dtype %d", (int)(dtype)); \
    return SSB_ERR_INVALID;                            \
  }

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// A kernel launched with the programmatic-stream-serialization attribute may be scheduled while
// its predecessor in the stream is still running; it blocks at pdl_wait() until all of the
// predecessor's memory operations are visible.  Kernels call pdl_wait() before their first
// global-memory access, so launch latency / CTA scheduling / on-chip prologues (barrier init,
// TMEM allocation, tensor-map prefetch) overlap with the predecessor's tail.
extern int g_ssb_pdl;
// SM count of the device (ssb_prepare(); 148 = B200 until then): every co-residency bound (grid-barrier kernels, persistent
// grids) uses it; the plain `148 * k` caps elsewhere are tuning constants of grid-stride loops, correct for any SM count
extern int g_ssb_num_sms;
// The dependent grid is released only once this one is past its own wait (measured: releasing at kernel
// entry lets a whole chain of parked grids pile up and is slower; profiles/r1b_pdl.md).
#ifndef SSB_TRACE
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory"); }
#else
// Development build (make TRACE=1 -> lib/libsemiseg_b200_trace.so, tools/trace_step.py): block (0,0,0) of every launch
// records %globaltimer at kernel entry and when it gets past its dependency wait, with the source line of the wait --
// the timeline of one step INSIDE the replayed graph (gaps between dependent kernels, overlap of the branches).
struct SsbTraceRec { unsigned long long t_entry, t_start; unsigned int line, grid, block, pad; };
#define SSB_TRACE_CAP 4096
static __device__ SsbTraceRec ssb_trace_buf[SSB_TRACE_CAP];
static __device__ unsigned int ssb_trace_n;
__device__ __forceinline__ unsigned long long ssb_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void pdl_wait_line(unsigned int line) {
  const bool rec = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && threadIdx.x == 0;
  const unsigned long long t0 = rec ? ssb_gtime() : 0ull;
  asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory");
  if (rec) {
    const unsigned int i = atomicAdd(&ssb_trace_n, 1u);
    if (i < SSB_TRACE_CAP) {
      SsbTraceRec r;
      r.t_entry = t0; r.t_start = ssb_gtime(); r.line = line;
      r.grid = gridDim.x * gridDim.y * gridDim.z; r.block = blockDim.x; r.pad = 0;
      ssb_trace_buf[i] = r;
    }
  }
}
#define pdl_wait() pdl_wait_line(__LINE__)
// phase marks inside a kernel (block (0,0,0) only; any one thread may call it): pad = 1, t_start = time of the mark
__device__ __forceinline__ void ssb_mark_line(unsigned int line) {
  if ((blockIdx.x | blockIdx.y | blockIdx.z) != 0) return;
  const unsigned int i = atomicAdd(&ssb_trace_n, 1u);
  if (i < SSB_TRACE_CAP) {
    SsbTraceRec r;
    r.t_entry = 0; r.t_start = ssb_gtime(); r.line = line; r.grid = 0; r.block = threadIdx.x; r.pad = 1;
    ssb_trace_buf[i] = r;
  }
}
#define SSB_MARK() ssb_mark_line(__LINE__)
// one per translation unit: copies (and optionally resets) this unit's records
#define SSB_TRACE_DEFINE(tu)                                                                         \
  extern "C" int ssb_trace_dump_##tu(void* host, int cap, int reset) {                              \
    unsigned int n = 0;                                                                              \
    cudaDeviceSynchronize();                                                                         \
    cudaMemcpyFromSymbol(&n, ssb_trace_n, sizeof(n));                                                \
    if (n > SSB_TRACE_CAP) n = SSB_TRACE_CAP;                                                        \
    if ((int)n > cap) n = cap;                                                                       \
    if (host && n) cudaMemcpyFromSymbol(host, ssb_trace_buf, n * sizeof(SsbTraceRec));               \
    if (reset) { unsigned int z = 0; cudaMemcpyToSymbol(ssb_trace_n, &z, sizeof(z)); }              \
    return (int)n;                                                                                   \
  }
#endif
#ifndef SSB_TRACE_DEFINE
#define SSB_TRACE_DEFINE(tu)
#define SSB_MARK()
#endif
__device__ __forceinline__ void pdl_trigger() {}

// g_ssb_pdl (env SSB_PDL): 0 = plain launches, 1 = every launch carries the attribute, 2 = only the
// kernels with an on-chip prologue worth overlapping (the tcgen05 convs: barrier init, TMEM alloc,
// tensor-map prefetch), launched through ssb_launch_pro.
template <int LEVEL, typename... KArgs, typename... Args>
static inline void ssb_launch_impl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_ssb_pdl == 1 || (g_ssb_pdl == 2 && LEVEL == 2)) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline void ssb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  ssb_launch_impl<1>(kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
template <typename... KArgs, typename... Args>
static inline void ssb_launch_pro(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  ssb_launch_impl<2>(kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}

static inline cudaStream_t to_stream(ssb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- element access ------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vectors: 4 floats or 8 bf16
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ void get(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
  __device__ __forceinline__ void set(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
  __device__ __forceinline__ void zero() { raw = make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void get(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void set(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
  __device__ __forceinline__ void zero() { raw = make_uint4(0u, 0u, 0u, 0u); }
};

// row validity in the flat padded layout
__device__ __forceinline__ bool row_valid(int row, int pitch, int len) {
  int pos = row % pitch;
  return pos >= 1 && pos <= len;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// arguments of the train-mode conv + BatchNorm(+residual)(+ReLU) fusion (conv_sm100.cu, BNF epilogue pass)
struct ssb_bnf_args {
  const ssb_bn* bn;       // BN of the conv output (batch statistics from bn->sums, filled by this launch)
  const ssb_bn* bn_res;   // BN of the residual tensor (its sums are already complete) or NULL
  const void* res;        // residual tensor in the output geometry or NULL
  void* y_act;            // normalised output
  int relu;
  unsigned int* barrier;  // zeroed grid-barrier counter owned by this launch
};

#define BN_EPS 1e-5
#define BN_MOMENTUM 0.1f

// ---- SyncBN exchange inside the consuming kernels (ssb_bn.sync_*; protocol of optim.cu's exchange kernel) ----
// Self-validating 8-byte words {32 payload bits, tag}: a value has arrived when both of its words carry this step's
// tag -- no fence, no flag round trip: one NVLink hop.  A slice is used once per step, so the tag is the step count.
#define SBX_HDR_BYTES 4096          // [0] (legacy exchange counter), [64] error word
#define SBX_MAX_WORLD 16
__device__ __forceinline__ void sbx_store(uint2* p, uint32_t payload, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 sbx_load(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
// Sum over the ranks (in rank order: bitwise identical everywhere) of NV doubles whose local values are mine[0..NV) and
// whose region indices are idx[0..NV).  push: this thread also publishes the local values (exactly one thread per
// value and launch must).  All peer loads of a polling round are in flight together.
template <int NV>
__device__ __forceinline__ void sbx_allsum(const ssb_bn& bn, const unsigned int (&idx)[NV], double (&val)[NV], bool push) {
  const unsigned int epoch = (unsigned int)bn.sync_sp->step;
  const int world = bn.sync_world, rank = bn.sync_rank;
  const unsigned long long* peers = reinterpret_cast<const unsigned long long*>(bn.sync_peers);
  if (push) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const uint32_t lo = (uint32_t)__double2loint(val[v]), hi = (uint32_t)__double2hiint(val[v]);
      for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<char*>(peers[r]) + SBX_HDR_BYTES) +
                     ((size_t)rank * bn.sync_slot + idx[v]) * 2;
        sbx_store(dst, lo, epoch);
        sbx_store(dst + 1, hi, epoch);
      }
    }
  }
  char* mine = reinterpret_cast<char*>(peers[rank]);
  const uint2* mail = reinterpret_cast<const uint2*>(mine + SBX_HDR_BYTES);
  const long long t0 = clock64();
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double t = 0.0;
    for (int r0 = 0; r0 < world; r0 += 4) {          // four ranks per round: their loads are in flight together
      double got[4] = {0.0, 0.0, 0.0, 0.0};
      unsigned int pending = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + j < world && r0 + j != rank) pending |= 1u << j;
      while (pending) {
        uint2 a[4], b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (pending & (1u << j)) {
            const uint2* src = mail + ((size_t)(r0 + j) * bn.sync_slot + idx[v]) * 2;
            a[j] = sbx_load(src);
            b[j] = sbx_load(src + 1);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((pending & (1u << j)) && a[j].y == epoch && b[j].y == epoch) {
            got[j] = __hiloint2double((int)b[j].x, (int)a[j].x);
            pending &= ~(1u << j);
          }
        }
        if (pending && clock64() - t0 > 40000000000LL) {   // ~20 s: a peer is gone -- record it (the host raises) and go on
          *reinterpret_cast<unsigned int*>(mine + 64) = epoch;
          pending = 0u;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + j < world) t += (r0 + j == rank) ? val[v] : got[j];
    }
    val[v] = t;
  }
}

// per-channel affine coefficients: y = x*scale + shift
// SYNC: compiled into the SyncBN instantiations only -- the exchange costs registers, and the elementwise kernels around
// it lose occupancy if every instantiation carries it (measured: +15 us per step at N = 1)
template <bool SYNC = false>
__device__ __forceinline__ void bn_coeffs(const ssb_bn& bn, int c, int C, int train, double inv_n, double n,
                                          bool writer, float& scale, float& shift) {
  float mean, invstd;
  if (train) {
    double st[2] = {__ldcg(&bn.sums[c]), __ldcg(&bn.sums[C + c])};   // (L2 read: the sums may have been completed by other blocks of this launch)
    if (SYNC && bn.sync_peers) {      // SyncBN: totals over the ranks, exchanged here (the writer block publishes this rank's sums)
      const unsigned int ix[2] = {bn.sync_fwd_off + (unsigned int)c, bn.sync_fwd_off + (unsigned int)(C + c)};
      sbx_allsum<2>(bn, ix, st, writer);
    }
    double m = st[0] * inv_n;
    double var = st[1] * inv_n - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = rsqrtf((float)var + (float)BN_EPS);
    invstd = invstd * (1.5f - 0.5f * ((float)var + (float)BN_EPS) * invstd * invstd);   // one Newton step: full fp32 accuracy
    if (writer) {
      bn.mean_invstd[c] = mean;
      bn.mean_invstd[C + c] = invstd;
      double unb = n > 1.0 ? var * (n / (n - 1.0)) : var;
      bn.running_mean[c] = (1.f - BN_MOMENTUM) * bn.running_mean[c] + BN_MOMENTUM * mean;
      bn.running_var[c] = (1.f - BN_MOMENTUM) * bn.running_var[c] + BN_MOMENTUM * (float)unb;
      if (c == 0 && bn.num_batches_tracked) *bn.num_batches_tracked += 1;
    }
  } else {
    mean = bn.running_mean[c];
    invstd = 1.0f / sqrtf(bn.running_var[c] + (float)BN_EPS);
  }
  scale = bn.gamma[c] * invstd;
  shift = bn.beta[c] - mean * scale;
}

// counter-based dropout RNG (keep decision for element idx), shared by fwd and bwd
__device__ __forceinline__ uint32_t ssb_hash3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
__device__ __forceinline__ bool dropout_keep(uint32_t seed, uint32_t step, uint32_t idx, float p) {
  uint32_t h = ssb_hash3(seed, step, idx);
  return (float)(h >> 8) * (1.0f / 16777216.0f) >= p;
}
