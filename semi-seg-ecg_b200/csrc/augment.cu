// GPU-resident augmentation of ECG strips: the per-item numpy/scipy pipeline of the
// reference's DataLoader workers, batched on the device and written straight into the step's input arena.
//   weak   RandomResizeCrop: Fourier resize (scipy.signal.resample) + centre zero-pad + random crop,
//          labels by nearest-neighbour resize                       (utils/transforms.py:93-127)
//   strong RandAugment over AmplitudeScaling / AdaptivePowerlineNoise / RandomPartialWhiteNoise /
//          RandomPartialSineNoise                                    (utils/transforms.py:340-351, 480-546, 647-657)
//   Standardize over (leads, time)                                   (utils/transforms.py:301-310)
// All random DRAWS come from the host (a few scalars per strip, same np.random call order as the
// reference); the bulk noise arrays are either passed in (parity tests: injected draws) or generated
// by a counter-based RNG on the device.
//
// The Fourier resize is evaluated directly (two dense DFT passes with exact integer phase indices into a
// shared-memory twiddle table): sizes are arbitrary integers in [L/2, 2L] (no FFT-friendly factorisation),
// only the L cropped output positions are needed, and at L = 2500 the whole batch is ~1 GFLOP.
#include <stdlib.h>

#include "common.cuh"

#define AUG_THREADS 256

// Both DFT passes walk the phase with a complex rotation (4 FMAs per term) and re-anchor it every
// AUG_RESYNC terms from the EXACT integer phase (k*n mod L) through sincospi, so there is no twiddle table,
// no data-dependent shared-memory traffic, and the rounding drift stays below ~1e-6.
#define AUG_RESYNC 32
#define AUG_FPB 64        // frequencies (or output positions) per block; x 4 segments of the summation range

__device__ __forceinline__ float2 unit_phase(long long num, int den) {   // exp(2*pi*i * (num mod den) / den)
  const int r = (int)(num % den);
  float s, c;
  sincospif(2.0f * (float)r / (float)den, &s, &c);
  return make_float2(c, s);
}

// ---------------------------------------------------------------------------------------------
// forward real DFT of every (strip, lead): spec[bc][k] = sum_n x[n] * exp(-2*pi*i*k*n/L), k <= min(size,L)/2
// grid = (ceil(K1 / 64), B*C); block = 64 frequencies x 4 segments of n
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AUG_THREADS)
aug_spectrum_kernel(const float* __restrict__ x, float2* __restrict__ spec, const int32_t* __restrict__ size, int C, int L,
                    int K1) {
  pdl_wait();
  extern __shared__ float sm[];
  float* xs = sm;                                                   // [L]
  float2* part = reinterpret_cast<float2*>(sm + ((L + 3) & ~3));    // [4][64]
  const int bc = blockIdx.y;
  const int b = bc / C;
  const int kmax = min(size[b], L) / 2;
  if ((int)(blockIdx.x * AUG_FPB) > kmax) return;
  for (int n = threadIdx.x; n < L; n += AUG_THREADS) xs[n] = x[(size_t)bc * L + n];
  __syncthreads();
  const int kl = threadIdx.x & (AUG_FPB - 1), seg = threadIdx.x / AUG_FPB;
  const int k = blockIdx.x * AUG_FPB + kl;
  const int seglen = (L + 3) / 4;
  const int n0 = seg * seglen, n1 = min(L, n0 + seglen);
  float re = 0.f, im = 0.f;
  if (k <= kmax) {
    const float2 w = unit_phase(k, L);            // one step of the rotation (conjugated below)
    for (int nb = n0; nb < n1; nb += AUG_RESYNC) {
      float2 ph = unit_phase((long long)k * nb, L);
      const int ne = min(n1, nb + AUG_RESYNC);
      for (int n = nb; n < ne; ++n) {
        const float a = xs[n];
        re = fmaf(a, ph.x, re);
        im = fmaf(-a, ph.y, im);
        const float c = ph.x * w.x - ph.y * w.y;
        ph.y = fmaf(ph.x, w.y, ph.y * w.x);
        ph.x = c;
      }
    }
  }
  part[seg * AUG_FPB + kl] = make_float2(re, im);
  __syncthreads();
  if (seg == 0 && k <= kmax) {
    const float2 p0 = part[kl], p1 = part[AUG_FPB + kl], p2 = part[2 * AUG_FPB + kl], p3 = part[3 * AUG_FPB + kl];
    spec[(size_t)bc * K1 + k] = make_float2((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y));
  }
}

// ---------------------------------------------------------------------------------------------
// inverse real DFT of length `size` evaluated only at the cropped positions, + pad + crop, + labels
// grid = (ceil(L / 64), B*C); block = 64 output positions x 4 segments of k
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AUG_THREADS)
aug_resize_crop_kernel(const float2* __restrict__ spec, const int64_t* __restrict__ lab_in, float* __restrict__ y,
                       int64_t* __restrict__ lab_out, const int32_t* __restrict__ size_arr,
                       const int32_t* __restrict__ start_arr, int C, int L, int K1) {
  pdl_wait();
  extern __shared__ float sm[];
  const int bc = blockIdx.y;
  const int b = bc / C, c = bc - b * C;
  const int size = size_arr[b], start = start_arr[b];
  const int N = min(size, L);
  const int kmax = N / 2;
  float2* Xs = reinterpret_cast<float2*>(sm);            // [K1]
  float* part = sm + 2 * (size_t)(K1 + 1);               // [4][64]
  for (int k = threadIdx.x; k <= kmax; k += AUG_THREADS) {
    float2 v = spec[(size_t)bc * K1 + k];
    // irfft weights: 1 for DC, 2 for interior bins; the shared Nyquist bin of an even N keeps
    // scipy.signal.resample's fix-up (x2 when shrinking, x0.5 when growing, untouched when equal)
    float w = (k == 0) ? 1.f : 2.f;
    if ((N & 1) == 0 && k == kmax) w = (size < L) ? 2.f : 1.f;
    Xs[k] = make_float2(v.x * w, v.y * w);
  }
  __syncthreads();
  const int jl = threadIdx.x & (AUG_FPB - 1), seg = threadIdx.x / AUG_FPB;
  const int j = blockIdx.x * AUG_FPB + jl;
  const int pad = L - size;
  const int left = pad > 0 ? pad / 2 : 0;
  const int p = start + j - left;                        // position in the resized strip
  const bool inside = j < L && p >= 0 && p < size;
  float acc = 0.f;
  if (inside) {
    const int nk = kmax + 1;
    const int seglen = (nk + 3) / 4;
    const int k0 = seg * seglen, k1 = min(nk, k0 + seglen);
    const float2 w = unit_phase(p, size);
    for (int kb = k0; kb < k1; kb += AUG_RESYNC) {
      float2 ph = unit_phase((long long)kb * p, size);
      const int ke = min(k1, kb + AUG_RESYNC);
      for (int k = kb; k < ke; ++k) {
        const float2 X = Xs[k];
        acc = fmaf(X.x, ph.x, acc);
        acc = fmaf(-X.y, ph.y, acc);
        const float cc = ph.x * w.x - ph.y * w.y;
        ph.y = fmaf(ph.x, w.y, ph.y * w.x);
        ph.x = cc;
      }
    }
  }
  part[seg * AUG_FPB + jl] = acc;
  __syncthreads();
  if (seg != 0 || j >= L) return;
  const float out = inside ? ((part[jl] + part[AUG_FPB + jl]) + (part[2 * AUG_FPB + jl] + part[3 * AUG_FPB + jl])) / (float)L : 0.f;
  y[(size_t)bc * L + j] = out;
  if (lab_in && c == 0) {
    int64_t lab = 0;
    if (inside) {
      // np.linspace(0, L-1, size)[p] in float64, then interp1d(kind='nearest'): half-way points round DOWN
      int src;
      if (size == 1) src = 0;
      else if (p == size - 1) src = L - 1;
      else {
        const double step = (double)(L - 1) / (double)(size - 1);
        const double pos = (double)p * step;
        src = (int)ceil(pos - 0.5);
        src = src < 0 ? 0 : (src > L - 1 ? L - 1 : src);
      }
      lab = lab_in[(size_t)b * L + src];
    }
    lab_out[(size_t)b * L + j] = lab;
  }
}

// ---------------------------------------------------------------------------------------------
// The same two transforms as FFTs (round 2): one block per (strip, lead), everything in shared memory.
// Both are chirp-z transforms -- pass 1 wants the first kmax+1 <= L/2+1 bins of a length-L DFT (L = 2500 = 2^2 5^4 is not
// a power of two), pass 2 evaluates a trigonometric polynomial of <= L/2+1 terms at L consecutive points of a grid of
// ARBITRARY length `size` (int(L * ratio), possibly prime) -- so both go through Bluestein's identity
//     n k = (n^2 + k^2 - (k - n)^2) / 2 :   X[k] = w[k] * sum_n (x[n] w[n]) * conj(w)[k - n],   w[n] = exp(-+ i pi n^2 / den)
// i.e. a linear convolution of length <= L + L/2 + 1 with a chirp, done as a cyclic convolution of P = 4096 (L = 2500)
// or 8192 (L = 5000) points: forward radix-2 DIF (natural in, bit-reversed out) of both sequences, pointwise product,
// inverse DIT (bit-reversed in, natural out) -- no bit-reversal pass.  Chirp phases are reduced exactly in integers
// (n^2 mod 2 den) before sincospi.  ~0.5 MFLOP per strip instead of ~25 MFLOP for the dense evaluation above, which stays
// as the fallback for L > 5460 (P would exceed 8192 points = 128 KB of shared memory) and as the A/B reference
// (SSB_AUG_FFT=0).
// ---------------------------------------------------------------------------------------------
#define AFFT_THREADS 512

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// exp(sign * i * pi * n^2 / den), n < 65536
__device__ __forceinline__ float2 chirp(unsigned int n, unsigned int den, float sign) {
  const unsigned int r = (n * n) % (2u * den);
  float s, c;
  sincospif((float)r / (float)den, &s, &c);
  return make_float2(c, sign * s);
}
// Twiddles are evaluated on the fly (sincospi of an exact dyadic fraction): a shared-memory table read at T[j * P/len]
// is a 16..32-way bank conflict in every stage with len <= P/8 (stride of 128 B or more) -- measured 44 us per launch
// with the table against the dense kernels' 33 us.
__device__ __forceinline__ float2 twiddle(int j, int len) {            // exp(-2 pi i j / len), len a power of two
  float s, c;
  sincospif(2.0f * (float)j / (float)len, &s, &c);
  return make_float2(c, -s);
}
__device__ __forceinline__ void fft_dif_fwd(float2* a, int P) {
  for (int half = P >> 1; half >= 1; half >>= 1) {
    for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
      const int j = t & (half - 1);
      const int i = ((t - j) << 1) + j;
      const float2 u = a[i], v = a[i + half];
      a[i] = make_float2(u.x + v.x, u.y + v.y);
      a[i + half] = cmul(make_float2(u.x - v.x, u.y - v.y), twiddle(j, 2 * half));
    }
    __syncthreads();
  }
}
__device__ __forceinline__ void fft_dit_inv(float2* a, int P) {      // unscaled
  for (int half = 1; half < P; half <<= 1) {
    for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
      const int j = t & (half - 1);
      const int i = ((t - j) << 1) + j;
      const float2 w = twiddle(j, 2 * half);
      const float2 u = a[i], v = cmul(a[i + half], make_float2(w.x, -w.y));
      a[i] = make_float2(u.x + v.x, u.y + v.y);
      a[i + half] = make_float2(u.x - v.x, u.y - v.y);
    }
    __syncthreads();
  }
}
// a <- cyclic convolution of a and h (both length P, in shared memory); h is destroyed
__device__ __forceinline__ void fft_convolve(float2* a, float2* h, int P) {
  __syncthreads();
  fft_dif_fwd(a, P);
  fft_dif_fwd(h, P);
  for (int i = threadIdx.x; i < P; i += blockDim.x) a[i] = cmul(a[i], h[i]);
  __syncthreads();
  fft_dit_inv(a, P);
}

__global__ void __launch_bounds__(AFFT_THREADS)
aug_spectrum_fft_kernel(const float* __restrict__ x, float2* __restrict__ spec, const int32_t* __restrict__ size, int C, int L,
                        int K1, int P) {
  pdl_wait();
  extern __shared__ float sm[];
  float2* a = reinterpret_cast<float2*>(sm);
  float2* h = a + P;
  const int bc = blockIdx.x;
  const int b = bc / C;
  const int kmax = min(size[b], L) / 2;
  const int M = kmax + 1;
  for (int n = threadIdx.x; n < P; n += AFFT_THREADS) {
    float2 av = make_float2(0.f, 0.f), hv = make_float2(0.f, 0.f);
    if (n < L) {
      const float2 w = chirp((unsigned)n, (unsigned)L, -1.f);
      const float xv = x[(size_t)bc * L + n];
      av = make_float2(xv * w.x, xv * w.y);
    }
    if (n < M) hv = chirp((unsigned)n, (unsigned)L, 1.f);
    else if (n > P - L) hv = chirp((unsigned)(P - n), (unsigned)L, 1.f);       // lags -(L-1) .. -1
    a[n] = av;
    h[n] = hv;
  }
  fft_convolve(a, h, P);
  const float invP = 1.0f / (float)P;
  for (int k = threadIdx.x; k <= kmax; k += AFFT_THREADS) {
    const float2 v = cmul(a[k], chirp((unsigned)k, (unsigned)L, -1.f));
    spec[(size_t)bc * K1 + k] = make_float2(v.x * invP, v.y * invP);
  }
}

__global__ void __launch_bounds__(AFFT_THREADS)
aug_resize_crop_fft_kernel(const float2* __restrict__ spec, const int64_t* __restrict__ lab_in, float* __restrict__ y,
                           int64_t* __restrict__ lab_out, const int32_t* __restrict__ size_arr,
                           const int32_t* __restrict__ start_arr, int C, int L, int K1, int P) {
  pdl_wait();
  extern __shared__ float sm[];
  float2* a = reinterpret_cast<float2*>(sm);
  float2* h = a + P;
  const int bc = blockIdx.x;
  const int b = bc / C, c = bc - b * C;
  const int size = size_arr[b], start = start_arr[b];
  const int N = min(size, L);
  const int kmax = N / 2;
  const int pad = L - size;
  const int left = pad > 0 ? pad / 2 : 0;
  const int p0 = start - left;                            // position in the resized strip of output sample 0
  for (int n = threadIdx.x; n < P; n += AFFT_THREADS) {
    float2 av = make_float2(0.f, 0.f), hv = make_float2(0.f, 0.f);
    if (n <= kmax) {
      float2 v = spec[(size_t)bc * K1 + n];
      // irfft weights: 1 for DC, 2 for interior bins; the shared Nyquist bin of an even N keeps
      // scipy.signal.resample's fix-up (x2 when shrinking, x0.5 when growing, untouched when equal)
      float w = (n == 0) ? 1.f : 2.f;
      if ((N & 1) == 0 && n == kmax) w = (size < L) ? 2.f : 1.f;
      long long r = ((long long)n * (long long)p0) % (long long)size;      // exp(2 pi i n p0 / size), exact phase
      if (r < 0) r += size;
      float sn, cs;
      sincospif(2.0f * (float)r / (float)size, &sn, &cs);
      v = cmul(make_float2(v.x * w, v.y * w), make_float2(cs, sn));
      av = cmul(v, chirp((unsigned)n, (unsigned)size, 1.f));
    }
    if (n < L) hv = chirp((unsigned)n, (unsigned)size, -1.f);
    else if (n >= P - kmax) hv = chirp((unsigned)(P - n), (unsigned)size, -1.f);      // lags -kmax .. -1
    a[n] = av;
    h[n] = hv;
  }
  fft_convolve(a, h, P);
  const float scale = 1.0f / ((float)P * (float)L);
  for (int j = threadIdx.x; j < L; j += AFFT_THREADS) {
    const int p = p0 + j;
    const bool inside = p >= 0 && p < size;
    float out = 0.f;
    if (inside) {
      const float2 w = chirp((unsigned)j, (unsigned)size, 1.f);
      out = (a[j].x * w.x - a[j].y * w.y) * scale;
    }
    y[(size_t)bc * L + j] = out;
    if (lab_in && c == 0) {
      int64_t lab = 0;
      if (inside) {
        // np.linspace(0, L-1, size)[p] in float64, then interp1d(kind='nearest'): half-way points round DOWN
        int src;
        if (size == 1) src = 0;
        else if (p == size - 1) src = L - 1;
        else {
          const double step = (double)(L - 1) / (double)(size - 1);
          const double pos = (double)p * step;
          src = (int)ceil(pos - 0.5);
          src = src < 0 ? 0 : (src > L - 1 ? L - 1 : src);
        }
        lab = lab_in[(size_t)b * L + src];
      }
      lab_out[(size_t)b * L + j] = lab;
    }
  }
}

// FFT length of the chirp convolutions (0: use the dense kernels).  One block per (strip, lead) does ~36 radix-2 stage
// passes: ~50 us for a 4096-point problem however few strips there are, while the dense kernels spread a strip over
// 20-40 blocks.  Measured at 16 strips x 1 lead x 2500: dense 33 + 28 us, FFT 52 + 55 us; the dense cost grows with
// strips x leads x L^2 (config 5: 768 strip-leads x 5000 samples = 154 GFLOP), the FFT cost with strips x leads / #SMs.
// SSB_AUG_FFT = 0 | 1 forces a path; default: FFT from 96 strip-leads on (every SM has a block).
static int aug_fft_points(int L, int strips) {
  const char* e = getenv("SSB_AUG_FFT");
  if (e && atoi(e) == 0) return 0;
  if (!(e && atoi(e) == 1) && strips < 96) return 0;
  int P = 64;
  while (P < L + L / 2 + 2) P <<= 1;
  return P <= 8192 ? P : 0;
}

// ---------------------------------------------------------------------------------------------
// strong augmentation + standardise, one block per strip
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float aug_uniform(uint32_t seed, uint32_t stream, uint32_t idx) {
  return ((float)(ssb_hash3(seed, stream, idx) >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
}
__device__ __forceinline__ float aug_normal(uint32_t seed, uint32_t stream, uint32_t idx) {
  const float u1 = aug_uniform(seed, stream, 2 * idx), u2 = aug_uniform(seed, stream, 2 * idx + 1);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// block-wide sum (all threads get the result)
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < AUG_THREADS / 32; ++i) t += red[i];
  return t;
}

// ---- exact order statistics without sorting (numpy.percentile needs two adjacent ones per percentile) ----
// Round 1 sorted every lead with a block-wide bitonic sort (78 compare-exchange stages with a barrier each: ~26 us per
// 2500-sample lead, 130 us per launch -- the largest item of the augmentation).  The k-th smallest value is found by a
// radix select instead: four passes over the samples, 8 key bits per pass, a 256-bin shared-memory histogram and a
// warp scan per pass; the (k+1)-th comes from one more pass (count of values <= the k-th, smallest value above it).
__device__ __forceinline__ unsigned int aug_key(float f) {            // order-preserving float -> uint
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float aug_unkey(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// k-th smallest (0-based) of row[0..n); all threads of the block call it, all get the result.  hist: 256 + 2 words.
__device__ float block_kth(const float* row, int n, int k, unsigned int* hist) {
  unsigned int prefix = 0u, mask = 0u;
  int kk = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += AUG_THREADS) hist[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += AUG_THREADS) {
      const unsigned int key = aug_key(row[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {                       // warp 0: lane l owns bins 8l .. 8l+7
      const int lane = threadIdx.x;
      unsigned int c[8], mine = 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; mine += c[j]; }
      unsigned int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const unsigned int excl = incl - mine;
      if ((unsigned int)kk >= excl && (unsigned int)kk < incl) {     // exactly one lane
        unsigned int cum = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if ((unsigned int)kk >= cum && (unsigned int)kk < cum + c[j]) { hist[256] = (unsigned int)(lane * 8 + j); hist[257] = (unsigned int)kk - cum; }
          cum += c[j];
        }
      }
    }
    __syncthreads();
    prefix |= hist[256] << shift;
    mask |= 0xFFu << shift;
    kk = (int)hist[257];
    __syncthreads();
  }
  return aug_unkey(prefix);
}
// (k+1)-th smallest given the k-th (value vk): vk again if at least k+2 values are <= vk, else the smallest value above vk
__device__ float block_next(const float* row, int n, int k, float vk, float* red) {
  int cnt = 0;
  float mn = INFINITY;
  for (int i = threadIdx.x; i < n; i += AUG_THREADS) {
    const float v = row[i];
    if (v <= vk) ++cnt;
    else mn = fminf(mn, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { red[wid] = (float)cnt; red[AUG_THREADS / 32 + wid] = mn; }
  __syncthreads();
  float tc = 0.f, tm = INFINITY;
  for (int i = 0; i < AUG_THREADS / 32; ++i) { tc += red[i]; tm = fminf(tm, red[AUG_THREADS / 32 + i]); }
  __syncthreads();
  return (int)tc >= k + 2 ? vk : tm;
}
// numpy.percentile(row, q, method='linear')
__device__ float block_percentile(const float* row, int n, double q, unsigned int* hist, float* red) {
  const double h = (double)(n - 1) * q / 100.0;
  const int lo = (int)floor(h);
  const float t = (float)(h - (double)lo);
  const float a = block_kth(row, n, lo, hist);
  const float b = lo + 1 < n ? block_next(row, n, lo, a, red) : a;
  return t >= 0.5f ? b - (b - a) * (1.0f - t) : a + (b - a) * t;
}

__global__ void __launch_bounds__(AUG_THREADS)
aug_strong_standardize_kernel(const float* __restrict__ x, float* __restrict__ y, const ssb_aug_op* __restrict__ ops,
                              int n_ops, const float* __restrict__ scales, const float* __restrict__ white, uint32_t seed,
                              const uint32_t* __restrict__ seed_dev, int C, int L, int fs, float level, int npow2) {
  pdl_wait();
  if (seed_dev) seed += *seed_dev;   // per-replay seed of a captured launch
  extern __shared__ float sm[];
  (void)sm;
  (void)npow2;
  __shared__ float red[AUG_THREADS / 32];
  __shared__ float red2[2 * (AUG_THREADS / 32)];
  __shared__ unsigned int hist[258];
  __shared__ float s_amp;
  const int b = blockIdx.x;
  const size_t base = (size_t)b * C * L;
  const int n = C * L;
  float* yb = y + base;
  for (int i = threadIdx.x; i < n; i += AUG_THREADS) yb[i] = x[base + i];
  __syncthreads();
  for (int oi = 0; oi < n_ops; ++oi) {
    const ssb_aug_op op = ops[(size_t)b * n_ops + oi];
    if (!op.apply) continue;
    if (op.kind == SSB_AUG_AMPLITUDE) {
      const float sigma = level * 0.5f;
      for (int i = threadIdx.x; i < n; i += AUG_THREADS) {
        const float sc = scales ? scales[base + i] : 1.0f + sigma * aug_normal(seed, (uint32_t)(b * 8 + oi), (uint32_t)i);
        yb[i] *= sc;
      }
    } else if (op.kind == SSB_AUG_POWERLINE) {
      for (int c = 0; c < C; ++c) {
        float* row = yb + (size_t)c * L;
        const float p95 = block_percentile(row, L, 95.0, hist, red2);
        const float p05 = block_percentile(row, L, 5.0, hist, red2);
        if (threadIdx.x == 0) s_amp = (p95 - p05) * 0.5f;
        __syncthreads();
        const float amp = s_amp;
        const int f2 = 2 * op.a;                          // phase in half-turns: 2*f*t/fs, reduced exactly
        for (int t = threadIdx.x; t < L; t += AUG_THREADS) {
          const int r = (int)(((long long)f2 * t) % (2 * fs));
          row[t] += amp * sinpif((float)r / (float)fs);
        }
        __syncthreads();
      }
    } else if (op.kind == SSB_AUG_PARTIAL_WHITE) {
      const int cnt = op.a, st = op.b;
      for (int i = threadIdx.x; i < C * cnt; i += AUG_THREADS) {
        const int c = i / cnt, t = i - c * cnt;
        const float nz = white ? white[base + (size_t)c * L + t] : aug_normal(seed, (uint32_t)(b * 8 + oi), (uint32_t)(c * L + t));
        yb[(size_t)c * L + st + t] += level * nz;
      }
    } else if (op.kind == SSB_AUG_PARTIAL_SINE) {
      const int cnt = op.a, st = op.b;
      // amplitude = level, freq = 0.5/level: sin(2*pi*(t/L)/freq) = sinpi(4*level*t/L)
      for (int i = threadIdx.x; i < C * cnt; i += AUG_THREADS) {
        const int c = i / cnt, t = i - c * cnt;
        yb[(size_t)c * L + st + t] += level * sinpif(4.0f * level * (float)t / (float)L);
      }
    }
    __syncthreads();
  }
  // standardise over (C, L): two passes for the variance
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += AUG_THREADS) s += yb[i];
  const float mean = block_sum(s, red) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += AUG_THREADS) {
    const float d = yb[i] - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum(q, red) / (float)n;
  const float sd = sqrtf(var);
  const float inv = sd != 0.f ? 1.0f / sd : 0.f;
  for (int i = threadIdx.x; i < n; i += AUG_THREADS) yb[i] = (yb[i] - mean) * inv;
}

int ssb_aug_prepare() {
  cudaError_t e = cudaFuncSetAttribute(aug_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(aug_resize_crop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(aug_strong_standardize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(aug_spectrum_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(aug_resize_crop_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) {
    ssb_set_error("ssb_aug_prepare: %s", cudaGetErrorString(e));
    return SSB_ERR_CUDA;
  }
  return SSB_OK;
}

extern "C" {

int ssb_aug_spectrum(const float* x, float* spec, const int32_t* size, int B, int C, int L, ssb_stream_t stream) {
  SSB_REQUIRE(x && spec && size, "ssb_aug_spectrum: null pointer");
  SSB_REQUIRE(B > 0 && C > 0 && L >= 4 && L <= 8192, "ssb_aug_spectrum: bad shape (B=%d C=%d L=%d; L in [4, 8192])", B, C, L);
  const int K1 = L / 2 + 1;
  if (const int P = aug_fft_points(L, B * C)) {
    ssb_launch(aug_spectrum_fft_kernel, dim3(B * C), dim3(AFFT_THREADS), (size_t)(2 * P) * sizeof(float2), to_stream(stream), x,
               reinterpret_cast<float2*>(spec), size, C, L, K1, P);
    SSB_LAUNCH_CHECK("ssb_aug_spectrum");
    return SSB_OK;
  }
  const size_t smem = ((size_t)((L + 3) & ~3) + 2 * 4 * AUG_FPB) * sizeof(float);
  ssb_launch(aug_spectrum_kernel, dim3(ceil_div(K1, AUG_FPB), B * C), dim3(AUG_THREADS), smem, to_stream(stream), x,
             reinterpret_cast<float2*>(spec), size, C, L, K1);
  SSB_LAUNCH_CHECK("ssb_aug_spectrum");
  return SSB_OK;
}

int ssb_aug_resize_crop(const float* spec, const int64_t* lab_in, float* y, int64_t* lab_out, const int32_t* size,
                        const int32_t* start, int B, int C, int L, int max_size, ssb_stream_t stream) {
  SSB_REQUIRE(spec && y && size && start, "ssb_aug_resize_crop: null pointer");
  SSB_REQUIRE((lab_in == nullptr) == (lab_out == nullptr), "ssb_aug_resize_crop: lab_in and lab_out go together");
  SSB_REQUIRE(B > 0 && C > 0 && L >= 4 && L <= 8192, "ssb_aug_resize_crop: bad shape (B=%d C=%d L=%d; L in [4, 8192])", B, C, L);
  SSB_REQUIRE(max_size >= 1 && max_size <= 2 * L, "ssb_aug_resize_crop: max_size %d out of [1, 2L]", max_size);
  const int K1 = L / 2 + 1;
  if (const int P = aug_fft_points(L, B * C)) {
    ssb_launch(aug_resize_crop_fft_kernel, dim3(B * C), dim3(AFFT_THREADS), (size_t)(2 * P) * sizeof(float2), to_stream(stream),
               reinterpret_cast<const float2*>(spec), lab_in, y, lab_out, size, start, C, L, K1, P);
    SSB_LAUNCH_CHECK("ssb_aug_resize_crop");
    return SSB_OK;
  }
  const size_t smem = (2 * (size_t)(K1 + 1) + 4 * AUG_FPB) * sizeof(float);
  ssb_launch(aug_resize_crop_kernel, dim3(ceil_div(L, AUG_FPB), B * C), dim3(AUG_THREADS), smem, to_stream(stream),
             reinterpret_cast<const float2*>(spec), lab_in, y, lab_out, size, start, C, L, K1);
  SSB_LAUNCH_CHECK("ssb_aug_resize_crop");
  return SSB_OK;
}

int ssb_aug_strong_standardize(const float* x, float* y, const ssb_aug_op* ops, int n_ops, const float* scales,
                               const float* white, uint32_t seed, const uint32_t* seed_dev, int B, int C, int L, int fs,
                               float level, ssb_stream_t stream) {
  SSB_REQUIRE(x && y, "ssb_aug_strong_standardize: null pointer");
  SSB_REQUIRE(n_ops == 0 || ops, "ssb_aug_strong_standardize: ops table missing");
  SSB_REQUIRE(B > 0 && C > 0 && L >= 2 && L <= 8192 && fs > 0 && n_ops >= 0 && n_ops <= 8,
              "ssb_aug_strong_standardize: bad arguments (B=%d C=%d L=%d fs=%d n_ops=%d)", B, C, L, fs, n_ops);
  int npow2 = 1;
  while (npow2 < L) npow2 <<= 1;
  ssb_launch(aug_strong_standardize_kernel, dim3(B), dim3(AUG_THREADS), 0, to_stream(stream), x, y,
             ops, n_ops, scales, white, seed, seed_dev, C, L, fs, level, npow2);
  SSB_LAUNCH_CHECK("ssb_aug_strong_standardize");
  return SSB_OK;
}

}  // extern "C"

SSB_TRACE_DEFINE(augment)
