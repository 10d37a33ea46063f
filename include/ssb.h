/*
 * ssb.h -- C ABI of libsemiseg_b200.so: the sm_100a kernels behind the SemiSegECG
 * training hot path (1-D ResNet + FCN head + semi-supervised step).
 *
 * The reference (bakqui/semi-seg-ecg) has NO FFI/plugin interface: every GPU op on
 * its hot path is a PyTorch library call.  Each entry point below
 * therefore names the reference call site whose device work it replaces (file:line
 * relative to the reference tree).  Host code (Python, semi-seg-ecg_b200/src) binds
 * these with ctypes; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - plain C: raw device pointers, ints, floats; no torch / C++ types.
 *  - the caller owns every buffer (incl. workspaces); the library never allocates or
 *    frees device memory and never synchronises: all work is enqueued on `stream`
 *    and is CUDA-graph capturable.
 *  - return value: 0 on success, negative ssb_status otherwise; the message is
 *    available (thread-local) from ssb_last_error().  The library never exits.
 *  - dtype: SSB_F32 (exact-parity mode, 1e-5) or SSB_BF16 (2e-2 mode) = storage type
 *    of activations / activation gradients / the weight copy the convs read.  Accumulation is fp32,
 *    BatchNorm statistics fp64, master weights / optimizer state fp32.
 *
 * Activation layout ("flat padded NLC"): a tensor with B samples, C channels and
 * `len` valid positions per sample is stored as rows of C contiguous channels,
 * `pitch` rows per sample: row(b, t) = b*pitch + 1 + t.  Row b*pitch (front halo)
 * and rows b*pitch+len+1 .. (b+1)*pitch-1 (back pad, >= 1 row) are ZERO; every kernel
 * that writes such a tensor writes zeros there.  The halo rows implement the convs'
 * zero padding, so a k=3 conv is three row-shifted GEMMs over the flat row space and
 * M-tiles may span samples.  For a stride-2 layer pitch_in == 2*pitch_out, which makes
 * input row pairs line up with output rows (pair q <-> output row q+1).
 * C must be a multiple of 8.
 */
#ifndef SSB_H_
#define SSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_VERSION 1

typedef void* ssb_stream_t; /* cudaStream_t */

enum ssb_status {
  SSB_OK = 0,
  SSB_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  SSB_ERR_CUDA = -2,      /* CUDA runtime / driver error (launch failed) */
  SSB_ERR_UNSUPPORTED = -3 /* device is not sm_100 / algo not available for shape */
};

enum ssb_dtype { SSB_F32 = 0, SSB_BF16 = 1 };
enum ssb_algo { SSB_ALGO_SIMT = 0, SSB_ALGO_TCGEN05 = 1 };
enum ssb_loss_mode {
  SSB_LOSS_SUP = 0,
  SSB_LOSS_FIXMATCH = 1,     /* hard pseudo-labels, confidence mask (fixmatch.py:105-118; threshold 0: cps / stpp) */
  SSB_LOSS_SOFT = 2,         /* soft targets (mean_teacher.py:115-117) */
  SSB_LOSS_SOFT_MASKED = 3   /* soft targets, confidence mask: the consistency term of ReCo (reco.py:248-250) */
};

/* geometry of one flat padded NLC tensor */
typedef struct ssb_geom {
  int32_t B;      /* samples */
  int32_t pitch;  /* rows per sample (>= len + 2) */
  int32_t len;    /* valid positions per sample */
  int32_t C;      /* channels (multiple of 8) */
} ssb_geom;

/* one BatchNorm1d layer: parameter, buffer, statistic and gradient pointers
 * (nn.BatchNorm1d at resnet.py:41,50,254,290 and fcn_head.py:48). */
typedef struct ssb_bn {
  const float* gamma;      /* [C] weight */
  const float* beta;       /* [C] bias */
  float* running_mean;     /* [C] */
  float* running_var;      /* [C] */
  int64_t* num_batches_tracked; /* scalar, may be NULL */
  double* sums;            /* [2C] batch sum(x), sum(x^2); zeroed by the caller before ssb_bn_stats */
  float* mean_invstd;      /* [2C] batch mean / invstd saved by the train-mode forward */
  double* bwd_sums;        /* [2C] sum(g), sum(g*xhat); zeroed by the caller before the reduce */
  float* dgamma;           /* [C] gradient of weight (written, not accumulated) */
  float* dbeta;            /* [C] gradient of bias */
  int32_t count_mul;       /* SyncBN: world size (sums hold the all-reduced totals); 0/1 = local */
  int32_t pad;
  /* SyncBN statistics exchange INSIDE the consuming kernels (torch SyncBatchNorm's all_gather / all_reduce,
   * fixmatch.py:290-291): when sync_peers != NULL the kernels that turn `sums` / `bwd_sums` into coefficients
   * (ssb_bn_act_fwd, ssb_stem_bn_relu_pool_fwd, ssb_bn_bwd_fused, ssb_bn_bwd_apply, ssb_stem_bwd_apply) publish this
   * rank's local sums into every peer's mailbox over NVLink and sum the ranks' values, in rank order, from their own
   * mailbox -- no separate exchange launch.  Mailbox of a rank: ssb_syncbn_fused_mailbox_bytes(slot, world) bytes of
   * peer-mapped memory, zeroed once; every BN layer owns [sync_fwd_off, +2C) and [sync_bwd_off, +2C) doubles of each
   * rank's region; words are tagged with sync_sp->step (one use of a slice per step). */
  const uint64_t* sync_peers;          /* DEVICE array [sync_world] of the ranks' mailbox base addresses, or NULL */
  const struct ssb_step_params* sync_sp;
  int32_t sync_world, sync_rank;
  uint32_t sync_slot;                  /* doubles per rank region */
  uint32_t sync_fwd_off, sync_bwd_off; /* first double of this layer's forward / backward slice */
  uint32_t pad2;
} ssb_bn;

/* per-step scalars that change between replays of a captured step graph; lives in
 * device memory and is refreshed by one small H2D copy per step */
typedef struct ssb_step_params {
  float lr;            /* utils/lr_sched.py:6-18 */
  float inv_bias1;     /* 1/(1-beta1^t) */
  float inv_sqrt_bias2;/* 1/sqrt(1-beta2^t) */
  float ema_decay;     /* mean_teacher.py:46 */
  int32_t ema_first;   /* 1 while the teacher still aliases the student (mean_teacher.py:285-290) */
  int32_t step;        /* optimizer step count t (1-based) */
  uint32_t rng_seed;   /* dropout RNG key */
  uint32_t rng_step;   /* dropout RNG counter */
  float grad_scale;    /* multiplies gradients before the update (1/world_size after all-reduce sum) */
  float conf_thresh;   /* fixmatch.py:115 */
  float pad[6];
} ssb_step_params;

/* one (possibly skipped) RandAugment op of one strip; the draws are made on the host in the
 * reference's np.random call order (utils/transforms.py:574-583, 647-657) */
enum ssb_aug_kind {
  SSB_AUG_AMPLITUDE = 0,     /* AmplitudeScaling: x *= N(1, level*0.5) elementwise          (:340-351) */
  SSB_AUG_POWERLINE = 1,     /* AdaptivePowerlineNoise: a = frequency in Hz (50 | 60)       (:480-502) */
  SSB_AUG_PARTIAL_WHITE = 2, /* RandomPartialWhiteNoise: a = count, b = start               (:521-546) */
  SSB_AUG_PARTIAL_SINE = 3   /* RandomPartialSineNoise:  a = count, b = start               (:504-509, 529-546) */
};
typedef struct ssb_aug_op {
  int32_t kind;   /* ssb_aug_kind */
  int32_t apply;  /* RandomApply's coin (0: op skipped) */
  int32_t a, b;
} ssb_aug_op;

/* ---- library / device ------------------------------------------------------------ */
int ssb_version(void);
const char* ssb_last_error(void);
/* 0 if the current CUDA device can run these kernels (compute capability 10.x) */
int ssb_device_check(void);
/* number of kernels launched by this library since load (all threads); for bench accounting */
int64_t ssb_launch_count(void);

/* one-time per-device setup (opt-in shared-memory sizes of the large kernels); call once before
 * the first launch / before capturing a CUDA graph */
int ssb_prepare(void);

/* bytes of one rank's mailbox for the in-kernel SyncBN exchange (see ssb_bn.sync_peers) */
size_t ssb_syncbn_fused_mailbox_bytes(int slot_doubles, int world);

int ssb_memset_zero(void* p, size_t bytes, ssb_stream_t stream);

/* ---- stem: Conv1d(C_leads -> Cs, k7, s2, p3, bias=False)  (resnet.py:246-253) ------ */
/* x: [B, Cl, L] fp32 NCL (the reference's input layout); w: [Cs][Cl][7] fp32 master
 * weight; y: flat padded NLC, geometry g (g.len = floor((L-1)/2)+1, g.C = Cs). */
int ssb_stem_conv_fwd(const float* x, const float* w, void* y, int Cl, int L, ssb_geom g,
                      int dtype, ssb_stream_t stream);
/* the same conv with the train-mode BatchNorm statistics of its output (sums[2*Cs] += sum y, sum y^2 of the stored
 * values: what ssb_bn_stats computes) out of the same launch where the tensor-core kernel covers the shape (bf16, Cs in
 * {64, 128}); otherwise the conv followed by the statistics pass */
int ssb_stem_conv_fwd_stats(const float* x, const float* w, void* y, int Cl, int L, ssb_geom g, double* sums,
                            int dtype, ssb_stream_t stream);
/* dw[Cs][Cl][7] (fp32) += sum_{b,t} dy * x   (conv backward w.r.t. weight; the stem has no dgrad) */
int ssb_stem_conv_wgrad(const float* x, const void* dy, float* dw, int Cl, int L, ssb_geom g,
                        int dtype, ssb_stream_t stream);

/* ---- Conv1d k in {1,3}, stride in {1,2}, padding k/2, dilation 1, bias=False --------
 * (resnet.py:32-49, 283-289; fcn_head.py:40-47).  gin/gout: geometries of the conv
 * input / output tensors (same B; gin.pitch == stride*gout.pitch).
 * w: the weight in TAP-MAJOR layout [k][Cin][Cout] (element (co, ci, t) of the reference's
 * [Cout][Cin][k] tensor at ((t*Cin + ci)*Cout + co)), in `dtype`.  This is the ONE weight
 * layout of the library: fprop reads it as an N-contiguous B operand, dgrad as a K-contiguous
 * one, wgrad writes it with 16-byte vector reductions; the host side exposes the reference
 * shape as a strided view.  algo SSB_ALGO_TCGEN05 needs dtype bf16 and Cin, Cout multiples of 64. */
int ssb_conv1d_fwd(const void* x, const void* w, void* y, ssb_geom gin, ssb_geom gout, int k,
                   int stride, int dtype, int algo, ssb_stream_t stream);
/* same conv with the BatchNorm statistics of its output fused into the epilogue:
 * sums[0:Cout] += sum_rows y, sums[Cout:2Cout] += sum_rows y^2 (fp64, of the values as stored);
 * sums == NULL: plain conv.  Replaces the separate statistics pass of native_batch_norm
 * (resnet.py:41,50; fcn_head.py:48) after the conv. */
int ssb_conv1d_fwd_stats(const void* x, const void* w, void* y, ssb_geom gin, ssb_geom gout,
                         int k, int stride, double* sums, int dtype, int algo,
                         ssb_stream_t stream);
/* eval-mode conv + BatchNorm (running statistics) [+ residual] [+ ReLU] in one launch:
 * y = [relu]( bn(conv(x, w)) [+ res] ); res: tensor in the output geometry or NULL.  This is the
 * forward of the pseudo-label / teacher / inference passes (fixmatch.py:87-93, mean_teacher.py:90-92:
 * model.eval() => BatchNorm is a fixed per-channel affine map folded into the conv epilogue). */
int ssb_conv1d_bn_act_fwd(const void* x, const void* w, void* y, ssb_geom gin, ssb_geom gout,
                          int k, int stride, const ssb_bn* bn, const void* res, int relu,
                          int dtype, int algo, ssb_stream_t stream);
/* TRAIN-mode conv + BatchNorm(batch statistics) [+ residual | + bn_res(residual)] [+ ReLU] in one launch: the conv
 * epilogue adds the statistics of its tile to bn->sums, all tiles meet at a grid-wide barrier, then a second pass
 * over the accumulator (still in tensor memory) writes y_act, while y_raw keeps the conv output the backward needs.
 * Also does what ssb_bn_act_fwd(train) does to the BN state (running statistics, saved mean / invstd, counter).
 * Needs every CTA of the launch co-resident: returns SSB_ERR_UNSUPPORTED for shapes with more tiles than that (and
 * for the generic path / SyncBN); callers then use ssb_conv1d_fwd_stats + ssb_bn_act_fwd.  barrier: zeroed uint32. */
/* 1 if ssb_conv1d_fwd_bn_train can run this conv (call after ssb_prepare), else 0 */
int ssb_conv1d_fwd_bn_train_fits(ssb_geom gin, ssb_geom gout, int k, int stride, int dtype, int algo);
int ssb_conv1d_fwd_bn_train(const void* x, const void* w, void* y_raw, void* y_act, ssb_geom gin,
                            ssb_geom gout, int k, int stride, const ssb_bn* bn, const void* res,
                            const ssb_bn* bn_res, int relu, uint32_t* barrier, int dtype, int algo,
                            ssb_stream_t stream);
/* train rows and eval rows of the SAME conv in one launch (FixMatch: the pseudo-label forward uses the
 * student's weights, fixmatch.py:87-102): the first `train_samples` samples of x are train-mode rows
 * (raw output to y_train + statistics into sums, as ssb_conv1d_fwd_stats), the remaining samples are
 * eval-mode rows (y_eval = [relu](bn_eval(conv) [+ res_eval]), as ssb_conv1d_bn_act_fwd).  y_train,
 * y_eval and res_eval are bases of full [B*pitch, Cout] tensors (same row indexing; each launch only
 * touches its own row range of each). */
int ssb_conv1d_fwd_dual(const void* x, const void* w, void* y_train, void* y_eval, ssb_geom gin,
                        ssb_geom gout, int k, int stride, int train_samples, double* sums,
                        const ssb_bn* bn_eval, const void* res_eval, int relu, int dtype, int algo,
                        ssb_stream_t stream);
/* dx = conv_transpose(dy, w) (+ dx if accumulate) */
int ssb_conv1d_dgrad(const void* dy, const void* w, void* dx, ssb_geom gin, ssb_geom gout,
                     int k, int stride, int accumulate, int dtype, int algo, ssb_stream_t stream);
/* dgrad with the BatchNorm-backward REDUCE pass of the gradient it produces fused into the epilogue
 * (stride-1 convs): with g = dx * (y_act > 0) (dx as stored, after the optional accumulate),
 * bn->bwd_sums += (sum g, sum g*xhat(x_pre)); x_res/bn_res: the same for the residual-branch BN.
 * y_act / x_pre / x_res are tensors in the geometry of dx (gin).  Equivalent to ssb_conv1d_dgrad followed
 * by ssb_bn_bwd_reduce(dx, NULL, y_act, x_pre, bn, x_res, bn_res, gin). */
int ssb_conv1d_dgrad_bnred(const void* dy, const void* w, void* dx, ssb_geom gin, ssb_geom gout,
                           int k, int stride, int accumulate, const void* y_act, const void* x_pre,
                           const ssb_bn* bn, const void* x_res, const ssb_bn* bn_res, int dtype,
                           int algo, ssb_stream_t stream);
/* dw[k][Cin][Cout] (fp32, tap-major like w) += x^T dy; the caller zeroes dw once per step */
int ssb_conv1d_wgrad(const void* x, const void* dy, float* dw,
                     ssb_geom gin, ssb_geom gout, int k, int stride,
                     int dtype, int algo, ssb_stream_t stream);
/* bf16 copy of a flat fp32 parameter arena (same element offsets; n multiple of 8): the
 * storage-dtype weights the bf16 kernels read.  Replaces autocast's per-op weight casts. */
int ssb_weight_shadow(const float* src, void* dst, size_t n, int dtype, ssb_stream_t stream);

/* ---- BatchNorm1d (+ReLU, +residual, +MaxPool) -------------------------------------- */
/* sums[0:C] += sum_rows x, sums[C:2C] += sum_rows x^2 (fp64) */
int ssb_bn_stats(const void* x, ssb_geom g, double* sums, int dtype, ssb_stream_t stream);
/* y = [relu]( bn(x) [+ bn_res(res) | + res] ).  train != 0: batch statistics from
 * bn->sums (count B*len), saves mean/invstd, updates running stats (momentum 0.1,
 * unbiased var) and num_batches_tracked; train == 0: running statistics.
 * res may be NULL; bn_res NULL with res != NULL means identity residual.
 * (resnet.py:58-70: bn1+relu, bn2 + identity/downsample + relu; fcn_head.py:48-49) */
int ssb_bn_act_fwd(const void* x, const ssb_bn* bn, const void* res, const ssb_bn* bn_res,
                   void* y, ssb_geom g, int relu, int train, int dtype, ssb_stream_t stream);
/* stem tail: MaxPool1d(3,2,1)(relu(bn(c0)))  (resnet.py:254-257, 354-355).
 * arg (may be NULL): u8 [B*gout.pitch, C], per pooled element the window slot 0..2 of the first
 * maximum (torch's max_pool1d index rule) or 3 where the maximum is <= 0 (ReLU-dead); it is what
 * the backward pass routes gradients by (replaces max_pool1d's int64 indices + the ReLU mask). */
int ssb_stem_bn_relu_pool_fwd(const void* c0, const ssb_bn* bn, void* y, uint8_t* arg, ssb_geom gin,
                              ssb_geom gout, int train, int dtype, ssb_stream_t stream);
/* backward, pass 1: g = (g1 [+ g2]) * (y > 0 if y != NULL);  bn->bwd_sums += (sum g, sum g*xhat);
 * if x_res/bn_res given the same for the residual-branch BN. */
int ssb_bn_bwd_reduce(const void* g1, const void* g2, const void* y, const void* x,
                      const ssb_bn* bn, const void* x_res, const ssb_bn* bn_res,
                      ssb_geom g, int dtype, ssb_stream_t stream);
/* backward, pass 2: dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)); writes dgamma/dbeta;
 * residual branch: dx_res likewise (bn_res) or, for an identity residual, g itself to g_ident. */
int ssb_bn_bwd_apply(const void* g1, const void* g2, const void* y, const void* x,
                     const ssb_bn* bn, void* dx, const void* x_res, const ssb_bn* bn_res,
                     void* dx_res, void* g_ident, ssb_geom g, int dtype, ssb_stream_t stream);
/* both backward passes in ONE launch (reduce -> grid-wide barrier -> apply; g / y / x are read once).  Needs the
 * whole grid co-resident: returns SSB_ERR_UNSUPPORTED for tensors too large for that (callers then use the two
 * calls above).  barrier: a zeroed uint32 owned by this launch (the step engine keeps one per BN layer in the arena
 * it zeroes every step).  rep / rep_res: zeroed scratch of 8 replicas of this layer's (and the residual BN's) [2C]
 * sums, rep_stride doubles apart -- block b accumulates into replica b % 8, so that hundreds of blocks do not queue on
 * the same 2C addresses; the totals are also left in bn->bwd_sums.  Not for SyncBN (the statistics exchange sits
 * between the passes). */
/* 1 if ssb_bn_bwd_fused can run this shape (res_mode: 0 none, 1 identity residual, 2 residual BN), else 0 */
int ssb_bn_bwd_fused_fits(ssb_geom g, int res_mode, int has_y, int dtype);
int ssb_bn_bwd_fused(const void* g1, const void* y, const void* x, const ssb_bn* bn, void* dx,
                     const void* x_res, const ssb_bn* bn_res, void* dx_res, void* g_ident, ssb_geom g,
                     uint32_t* barrier, double* rep, double* rep_res, size_t rep_stride, int dtype,
                     ssb_stream_t stream);
/* stem tail backward (max-pool scatter + relu + bn), same two passes; gp: grad w.r.t. pooled output */
int ssb_stem_bwd_reduce(const void* gp, const void* c0, const uint8_t* arg, const ssb_bn* bn, ssb_geom gin,
                        ssb_geom gout, int dtype, ssb_stream_t stream);
int ssb_stem_bwd_apply(const void* gp, const void* c0, const uint8_t* arg, const ssb_bn* bn, void* dc0,
                       ssb_geom gin, ssb_geom gout, int dtype, ssb_stream_t stream);

/* ---- FCN head tail: Dropout(p) + Conv1d(C -> ncls, 1, bias=True)  (fcn_head.py:83-97) --- */
/* a: flat padded NLC [B*pitch, C]; low: [B, len, ncls] fp32 (compact).
 * drop_mask: optional explicit keep-mask u8 [B, len, C] (tests); else if p > 0 and
 * sp != NULL a counter-based RNG keyed by (sp->rng_seed, sp->rng_step, element). */
int ssb_head_cls_fwd(const void* a, const float* w, const float* bias, float* low, ssb_geom g,
                     int ncls, float p, const uint8_t* drop_mask, const ssb_step_params* sp,
                     int dtype, ssb_stream_t stream);
int ssb_head_cls_bwd(const float* dlow, const void* a, const float* w, void* da, float* dw,
                     float* dbias, ssb_geom g, int ncls, float p, const uint8_t* drop_mask,
                     const ssb_step_params* sp, int dtype, ssb_stream_t stream);

/* ---- linear upsample (F.interpolate, encoder_decoder.py:102-107) ------------------- */
/* low [B, Lin, ncls] fp32 -> out [B, ncls, Lout] fp32 (reference NCL logits) */
int ssb_upsample_fwd(const float* low, float* out, int B, int Lin, int Lout, int ncls,
                     int align_corners, ssb_stream_t stream);
/* dlow [B, Lin, ncls] = gather-reduce of dout [B, ncls, Lout] (deterministic) */
int ssb_upsample_bwd(const float* dout, float* dlow, int B, int Lin, int Lout, int ncls,
                     int align_corners, ssb_stream_t stream);

/* ---- pseudo-labels (fixmatch.py:89-91,115) ----------------------------------------- */
/* logits [U, ncls, L] fp32 NCL -> conf f32 [U,L], label i64 [U,L], mask u8 [U,L];
 * bit-exact with torch softmax(1).max(1)[0] / argmax(1) / (conf >= thr) on CUDA */
int ssb_pseudo_label(const float* logits, float thr, float* conf, int64_t* label, uint8_t* mask,
                     int U, int ncls, int L, ssb_stream_t stream);

/* ---- fused upsample + softmax + pseudo-label + loss + gradient -----------------------
 * (fixmatch.py:87-118, mean_teacher.py:90-117, base.py/encoder_decoder.py:110-111)
 * low_s: student low-res logits [Bl+Bu, Lin, ncls]; target: int64 [Bl, L];
 * low_t: teacher / self-eval low-res logits [Bu, Lin, ncls] (NULL for SSB_LOSS_SUP).
 * dlow: gradient of the total loss w.r.t. low_s (written).  sums (fp64[4], zeroed by the
 * caller) += { sum CE_x, sum masked CE_u or soft CE_u, sum mask (the masked modes), 0 }.
 * Total loss = sum_x/(Bl*L) for SUP, else (sum_x/(Bl*L) + sum_u/(Bu*L))/2.
 * Optional materialised outputs (may be NULL): conf f32 [Bu,L], label i64 [Bu,L], mask u8 [Bu,L].
 * thr is read from sp->conf_thresh when sp != NULL, else from thr. */
int ssb_semi_loss(const float* low_s, const int64_t* target, const float* low_t, float* dlow,
                  double* sums, int Bl, int Bu, int Lin, int L, int ncls, int mode, float thr,
                  const ssb_step_params* sp, int align_corners, float* conf, int64_t* label,
                  uint8_t* mask, ssb_stream_t stream);

/* ---- evaluation tail: upsample + softmax + argmax + CE + class counts  (base.py:198-221) -------------------------
 * low: low-res logits [B, Lin, ncls] fp32 of an eval-mode forward; target int64 [B, L].
 * sums (fp64[2], zeroed by the caller) += { sum of CE over the positions with a label in [0, ncls), their number }.
 * counts (int32 [B, ncls, 3], zeroed by the caller) += per sample and class { |pred == c and target == c|,
 * |pred == c|, |target == c| } -- the intersections and the two marginals every IoU / Dice style metric needs
 * (the reference one-hot-encodes both and hands them to torchmetrics on the CPU).  pred = argmax over the soft-max
 * outputs, first maximum.  Optional outputs (may be NULL): probs f32 [B, ncls, L] (the reference's returned
 * `outputs`), pred i64 [B, L]. */
int ssb_eval_metrics(const float* low, const int64_t* target, double* sums, int32_t* counts, float* probs,
                     int64_t* pred, int B, int Lin, int L, int ncls, int align_corners, ssb_stream_t stream);

/* ---- fused multi-tensor AdamW (+ EMA teacher)  (optimizer.py:22-34, mean_teacher.py:139-149) --- */
/* flat arenas of n floats.  p_ema may be NULL.  Scalars come from sp (device). */
int ssb_adamw_ema(float* p, const float* g, float* m, float* v, float* p_ema, size_t n,
                  double beta1, double beta2, double eps, double weight_decay,
                  const ssb_step_params* sp, ssb_stream_t stream);
/* the same update with gradient clipping (loss_scaler(..., clip_grad=max_norm), misc.py:242-250 ->
 * torch.nn.utils.clip_grad_norm_): gradients are multiplied by min(1, max_norm / (norm + 1e-6)), norm = gnorm[0] *
 * sp->grad_scale with gnorm (DEVICE, from ssb_grad_norm over the WHOLE arena) the norm of the stored gradients.
 * gnorm == NULL: no clipping. */
int ssb_adamw_ema_clip(float* p, const float* g, float* m, float* v, float* p_ema, size_t n,
                       double beta1, double beta2, double eps, double weight_decay,
                       const ssb_step_params* sp, const float* gnorm, double max_norm, ssb_stream_t stream);
/* dst = dst*d + src*(1-d) (teacher BN buffers); d from sp->ema_decay */
int ssb_ema(float* dst, const float* src, size_t n, const ssb_step_params* sp, ssb_stream_t stream);
/* teacher num_batches_tracked quirk: dst_f32[i] = dst_f32[i]*d + (float)src_i64[i]*(1-d) */
int ssb_ema_i64(float* dst, const int64_t* src, size_t n, const ssb_step_params* sp, ssb_stream_t stream);
/* out[0] = sqrt(sum g^2) (misc.py:265-278); ws: fp64[1] zeroed by the caller */
int ssb_grad_norm(const float* g, size_t n, double* ws, float* out, ssb_stream_t stream);

/* ---- SyncBatchNorm statistics exchange over NVLink peer memory (fixmatch.py:290-291) ----------------
 * Replaces the per-layer NCCL collectives of torch's SyncBatchNorm by one small kernel per exchange: every rank
 * stores its slice of the statistic arena into a slot of every peer's mailbox as self-validating 8-byte words
 * {32 payload bits, exchange number}, polls the words arriving in its own mailbox and sums the ranks' slices in
 * rank order (bitwise identical on all ranks): one NVLink hop, no fence, no flag round trip.
 * peers_dev: device array of `world` base addresses of the ranks' mailboxes (a symmetric allocation of
 * ssb_syncbn_mailbox_bytes(slot_doubles) bytes, zero-initialised; mapped into this process, e.g. with
 * torch.distributed._symmetric_memory or CUDA IPC).  All ranks must issue the same sequence of exchanges. */
size_t ssb_syncbn_mailbox_bytes(int slot_doubles);
int ssb_syncbn_exchange(double* slice, int n, const uint64_t* peers_dev, int world, int rank,
                        int slot_doubles, ssb_stream_t stream);

/* ---- GPU-resident augmentation (utils/semi_dataset.py:193-197, 235-242) ---------
 * Strips are [B, C, L] fp32 (the reference's item layout, batched); per-strip draws live in small
 * device arrays filled by the host. */
/* spec[B*C][L/2+1] (complex64) = rfft(x) bins 0 .. min(size[b], L)/2 (the others are not written) */
int ssb_aug_spectrum(const float* x, float* spec, const int32_t* size, int B, int C, int L,
                     ssb_stream_t stream);
/* RandomResizeCrop (utils/transforms.py:93-127) from the spectrum: Fourier resize to size[b]
 * (scipy.signal.resample semantics), centre zero-pad to >= L, crop [start[b], start[b]+L).
 * lab_in/lab_out: optional int64 [B, L] labels, nearest-neighbour resized (interp1d kind='nearest'),
 * padded with 0 and cropped with the same indices.  max_size >= max_b size[b] (<= 2L). */
int ssb_aug_resize_crop(const float* spec, const int64_t* lab_in, float* y, int64_t* lab_out,
                        const int32_t* size, const int32_t* start, int B, int C, int L, int max_size,
                        ssb_stream_t stream);
/* RandAugment ops (ops[B][n_ops], in drawn order) followed by Standardize over (C, L)
 * (utils/transforms.py:301-310); n_ops = 0 is the plain standardise of the weak view.
 * scales / white: optional explicit [B, C, L] draws of AmplitudeScaling's N(1, sigma) factors and
 * of the white noise (injected-draw parity); NULL: counter-based RNG keyed by (seed, strip, element).
 * seed_dev (may be NULL): device word added to `seed` at run time, so a launch captured in a CUDA graph
 * gets a fresh key on every replay.  level = RandAugment level / 10; fs = sampling rate of the powerline
 * op.  y may alias x. */
int ssb_aug_strong_standardize(const float* x, float* y, const ssb_aug_op* ops, int n_ops,
                               const float* scales, const float* white, uint32_t seed,
                               const uint32_t* seed_dev, int B, int C, int L, int fs, float level,
                               ssb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SSB_H_ */
