/*
 * deepcam_b200.h — C ABI of the B200-native DeepCAM training-step kernels.
 *
 * Every entry point replaces one PyTorch/ATen operator call that the reference
 * (azrael417/mlperf-deepcam) makes on its training hot path.  The reference is
 * pure Python; the "FFI" it would bind is therefore ctypes.
 * Citations are file:line in the reference tree:
 *   DX = src/deepCam/architecture/deeplab_xception.py
 *   LS = src/deepCam/utils/losses.py
 *   UT = src/deepCam/utils/utils.py
 *   TR = src/deepCam/train_hdf5_ddp.py
 *
 * Conventions
 *   - plain C types only; no torch types.  All pointers are DEVICE pointers unless
 *     a parameter is documented as host.
 *   - the library never allocates or frees device memory and keeps no pointer
 *     after a call returns.  Workspaces are passed in by the caller.
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*);
 *     the library never synchronises the device and never uses the legacy stream.
 *   - return value: 0 = ok, <0 = invalid argument / unsupported configuration,
 *     >0 = cudaError_t of a failed launch.  dc_last_error_string() describes the
 *     last failure on the calling thread.  No C++ exception crosses the ABI.
 *   - re-entrant and thread-safe (autograd calls backward from its own thread).
 *   - activations are channels-last: logical [N,H,W,C] addressed through dc_view.
 */
#ifndef DEEPCAM_B200_H
#define DEEPCAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DC_ABI_VERSION 2

enum { DC_F32 = 0, DC_BF16 = 1 };

/* Strided 4-D view, logical order (n, h, w, c); strides in ELEMENTS.
 * Channel slices of a concat buffer, parity sub-grids of a transposed
 * convolution and NCHW tensors are all expressed this way. */
typedef struct dc_view {
  void*   ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw, sc;
  int32_t dtype;      /* DC_F32 | DC_BF16 */
  int32_t reserved;
} dc_view;

/* Gather-GEMM descriptor shared by every dense contraction on the path
 * (nn.Conv2d fprop/dgrad DX:145,149,60,74,291,426,430,434,360-366 and
 *  nn.ConvTranspose2d fprop/dgrad DX:352,356,369,374):
 *    out[n,y,x,co] (+)= bias[co] + sum_t sum_ci in[n, y*stride_h+dh[t], x*stride_w+dw[t], ci] * W[wt[t]][ci][co]
 * Out-of-range input coordinates contribute zero (this is the zero padding). */
#define DC_MAX_TAPS 9
typedef struct dc_conv_desc {
  int32_t ntaps;
  int32_t dh[DC_MAX_TAPS];
  int32_t dw[DC_MAX_TAPS];
  int32_t wt[DC_MAX_TAPS];   /* weight-slice index of tap t inside the packed weights */
  int32_t stride_h, stride_w;
  int32_t accumulate;        /* 1: out += result (gradient accumulation at graph forks) */
  int32_t wtaps;             /* number of tap slices in the packed weight tensor */
  /* Two-segment output (dc_conv_gemm_* only; 0 = off): output channels co >= out_csplit are stored at
   * out + out_split_off + (co - out_csplit) * sc instead of out + co * sc.  This is how ONE launch writes two output
   * rows of a stride-2 nn.ConvTranspose2d (DX:374): channels = (row parity, column parity, co), see
   * DC_PACK_NTK_CONVT2.  out_csplit % 4 == 0. */
  int32_t out_csplit;
  int32_t flags;             /* DC_CONV_* bits below (0 = none; the field was `reserved`) */
  int64_t out_split_off;     /* in elements of the output view */
} dc_conv_desc;
/* dc_conv_desc.flags.  DC_CONV_WEIGHTS_STABLE: the packed weights were NOT written by the kernel that immediately precedes this
 * call on the stream (they were re-packed at the start of the captured graph, several launches earlier).  The tcgen05 kernels
 * then fetch their first weight tiles BEFORE griddepcontrol.wait, i.e. while the preceding kernel still drains; only the
 * activation operand waits for it.  Never set it right after a pack launch. */
#define DC_CONV_WEIGHTS_STABLE 1
/* DC_CONV_HALO_PACK: the packed weights are K-DENSE, [Co][slice][Ci] with Ci = the gathered operand's channel count (not rounded
 * up to 64): the layout of the tcgen05 kernel's HALO mode, which stages each tile's input region (halo included) once in shared
 * memory and builds the per-tap operand tiles from it, instead of one TMA box per tap.  Only pass it after
 * dc_conv_gemm_tc_halo_ok() returned 1 for the same descriptor and views. */
#define DC_CONV_HALO_PACK 2

/* ---- library / diagnostics ------------------------------------------------ */
int         dc_abi_version(void);
const char* dc_last_error_string(void);
/* returns 1 when the current device is sm_100 (tcgen05 kernels usable), 0 otherwise, <0 on error */
int         dc_device_supports_tcgen05(void);
/* Programmatic dependent launch between consecutive kernels of a stream (prologue of kernel N+1 overlaps the tail
 * of kernel N; results are identical to plain stream order).  Default on; DEEPCAM_B200_PDL=0 or dc_set_pdl(0) = off. */
int         dc_set_pdl(int on);
int         dc_get_pdl(void);
/* Deterministic mode (what `torch.use_deterministic_algorithms(True)` asks of the cuDNN path the reference runs on,
 * train_hdf5_ddp.py:345-364 being the same step): kernels that split a reduction over blocks and need no workspace switch to their
 * single-writer form (dc_reduce_hw / dc_gap_fwd: one block per channel chunk and image).  The weight-gradient kernels have explicit
 * two-stage entry points below (dc_conv_wgrad_tc_det, dc_conv_wgrad_simt_det, dc_dw_bwd_weight_det) that take a workspace: every
 * split stores its partial result into its own slice and a second launch adds the slices in slice order, so two runs on the same
 * inputs give bit-identical gradients.  The BatchNorm batch sums stay fp64 atomics of fp32 partial sums (order effects ~1e-16,
 * below the fp32 rounding of the statistics derived from them). */
int         dc_set_deterministic(int on);
int         dc_get_deterministic(void);

/* ---- layout / packing ------------------------------------------------------ */
/* Generic strided copy with dtype conversion, dst[n,h,w,c] = src[n,h,w,c] for c < src.c;
 * channels src.c..dst.c-1 of dst are zero-filled (channel padding, e.g. 3 -> 4 logits).
 * Replaces .to(device)/.contiguous()/permute glue around the model (TR:348, DX:441). */
int dc_copy_view(dc_view src, dc_view dst, void* stream);
int dc_fill_zero(void* ptr, size_t bytes, void* stream);
/* Input ingest: dst[p][c] = (src[p][c] - shift[c]) * scale[c] for npix pixels of C channels, src = the raw fp32 [H][W][C]
 * block of a CAM5 sample as stored in the HDF5 file, dst = NHWC bf16 (DC_BF16) or fp32 (DC_F32; bit-identical to the
 * reference's numpy expression).  Replaces the host-side transpose + normalisation of CamDataset.__getitem__
 * (data/cam_hdf5_dataset.py:126-129); shift = minval, scale = 1 / (maxval - minval) from stats.h5 (:97-102). */
int dc_ingest_hwc(const float* src, long long npix, int C, const float* shift, const float* scale, void* dst, int dst_dtype,
                  void* stream);
/* *ptrs[i] += 1 for i < count  (BatchNorm num_batches_tracked, 77 counters, DX:70.. normalizer) */
int dc_i64_increment_many(int64_t* const* ptrs, int count, void* stream);

/* Weight packing from the fp32 master parameter (PyTorch layout [A][B][taps]) into a kernel layout.
 *   src_k_first = 1: src is [k][n][taps]; 0: src is [n][k][taps]
 *   layout DC_PACK_TKN: dst[tap][k][n_pad]      (SIMT gather-GEMM B operand)
 *   layout DC_PACK_NTK: dst[n][tap][k_pad]      (tcgen05 B operand, K-major rows)
 *   layout DC_PACK_TC : dst[tap][c]             (depthwise, k = 1, n = c)
 *   layout DC_PACK_NTK_CONVT2: all four output-parity classes of a k3 s2 p1 op1 nn.ConvTranspose2d (DX:374) as ONE
 *     2x2-tap contraction: src is the parameter [k][n][3][3] (src_k_first = 1), taps = 4 (tap = dh*2+dw, dh,dw in {0,1}),
 *     N_pad = 4*G (G = channels per class, >= N), dst[(a*2+b)*G + n][tap][k_pad] = src[k][n][a+1-2dh][b+1-2dw]
 *     (zero where that kernel index is outside 0..2); y[2i+a, 2j+b, n] = sum_tap sum_k x[i+dh, j+dw, k] * dst[..].
 * padded entries are written as zero. */
enum { DC_PACK_TKN = 0, DC_PACK_NTK = 1, DC_PACK_NTK_CONVT2 = 2 };
int dc_pack_weight(const float* src, int K, int N, int taps, int src_k_first,
                   void* dst, int layout, int K_pad, int N_pad, int dst_dtype, void* stream);
/* Every weight pack of a training step in one launch.  `jobs_dev` is a DEVICE array of njobs descriptors (same
 * meaning as the dc_pack_weight arguments); job i owns blocks [block_start, block_start + n_blocks) of the grid and
 * total_blocks = sum of n_blocks.  taps*K_pad*N_pad of a job must be < 2^31. */
typedef struct dc_pack_job {
  const float* src;
  void*   dst;
  int32_t K, N, taps, src_k_first;
  int32_t layout, K_pad, N_pad, dst_dtype;
  int32_t block_start, n_blocks;
} dc_pack_job;
int dc_pack_weights_multi(const dc_pack_job* jobs_dev, int njobs, int total_blocks, void* stream);
/* Gradient unpack: G is [taps][n][k_stride] fp32 (k < K valid; k_stride >= K allows channel padding);
 * dst is the parameter-layout gradient: dst_k_first=0 -> [n][k][taps], 1 -> [k][n][taps]. */
int dc_unpack_wgrad(const float* G, int K, int N, int taps, int k_stride, int dst_k_first, float* dst, void* stream);

/* ---- dense contractions ---------------------------------------------------- */
/* SIMT fp32-accumulate gather-GEMM (fp32 parity mode, and the small/odd layers in bf16 mode).
 * `w` is DC_PACK_TKN with k = in.c, n_pad = round_up(out.c, 4), same dtype as `in`. */
int dc_conv_gemm_simt(const dc_conv_desc* d, dc_view in, const void* w, const float* bias,
                      dc_view out, void* stream);
/* SIMT weight gradient: G[wt[t]][co][ci] += sum_m in[pix(m,t), ci] * dout[m, co]; G fp32, pre-zeroed by caller. */
int dc_conv_wgrad_simt(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, void* stream);
/* Deterministic two-stage form.  ws: fp32 workspace of at least dc_conv_wgrad_simt_ws_elems() elements (0 = the launch does not
 * split the pixel reduction and ws may be null; -1 = bad arguments), 16-byte aligned, contents irrelevant on entry. */
long long dc_conv_wgrad_simt_ws_elems(const dc_conv_desc* d, dc_view in, dc_view dout);
int dc_conv_wgrad_simt_det(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, float* ws, long long ws_elems, void* stream);

/* tcgen05/TMEM/TMA implicit GEMM, bf16 operands, fp32 accumulation (sm_100a only).
 * `w` is DC_PACK_NTK bf16 with k_pad = round_up(in.c, 64).  in.c % 8 == 0, strides 16-byte aligned. */
int dc_conv_gemm_tc(const dc_conv_desc* d, dc_view in, const void* w, const float* bias,
                    dc_view out, void* stream);
/* Same contraction, plus the BatchNorm batch sums of the result: on return (in stream order) sums[c] += sum over pixels of
 * out[.., c] and sums[out.c + c] += sum of squares, taken from the bf16-rounded values that were stored (`sums` = the
 * zeroed per-layer workspace of dc_bn_ws_bytes; pass it to dc_bn_apply with DC_BN_TRAIN | DC_BN_SUMS_READY).  The sums come
 * out of the GEMM epilogue (no extra pass over `out`); output layouts the TMA-store epilogue cannot write get one
 * statistics pass appended.  d->accumulate must be 0.  Replaces the statistics half of nn.BatchNorm2d after a conv. */
int dc_conv_gemm_tc_bnstats(const dc_conv_desc* d, dc_view in, const void* w, const float* bias,
                            dc_view out, double* sums, void* stream);
/* 1 when dc_conv_gemm_tc / _bnstats / _bn_eval would run this contraction in HALO mode (multi-tap gather, uniform stride 1 | 2,
 * gathered or output channels <= 64, everything fits in shared memory): the caller must then pack the weights K-dense and set
 * DC_CONV_HALO_PACK (for Ci % 64 == 0 the two layouts coincide and the flag is optional). */
int dc_conv_gemm_tc_halo_ok(const dc_conv_desc* d, dc_view in, dc_view out);
/* Eval-mode Conv2d / ConvTranspose2d-parity-class + BatchNorm2d (+ReLU) in ONE tcgen05 launch (DX:144-149, 291-293, 352-372 under
 * net.eval(), TR:428): out = [relu]((conv(x) + bias) * scale + shift) with scale = gamma / sqrt(running_var + eps) and
 * shift = beta - running_mean * scale derived inside the epilogue, applied to the fp32 accumulator before the single bf16
 * rounding - the pre-BatchNorm tensor is never written.  Needs a bf16 output view the TMA-store epilogue can write; returns
 * -2 (nothing launched) otherwise, and the caller runs dc_conv_gemm_tc + dc_bn_apply. */
int dc_conv_gemm_tc_bn_eval(const dc_conv_desc* d, dc_view in, const void* w, const float* bias, dc_view out, const float* gamma,
                            const float* beta, const float* running_mean, const float* running_var, float eps, int relu, void* stream);
/* tcgen05 weight gradient (both operands pixel-major): same contract as dc_conv_wgrad_simt. */
int dc_conv_wgrad_tc(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, void* stream);
/* Deterministic two-stage form (TMA stores of the per-split partial tiles into ws[split][wtaps][Co][Ci] instead of TMA reduce-add
 * into G, then one launch that adds the slices to G in split order); same workspace contract as dc_conv_wgrad_simt_det. */
long long dc_conv_wgrad_tc_ws_elems(const dc_conv_desc* d, dc_view in, dc_view dout);
int dc_conv_wgrad_tc_det(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, float* ws, long long ws_elems, void* stream);

/* ---- depthwise 3x3 (SeparableConv2d_same.conv1 + fixed_padding, DX:45-51, 58-59, 63-64) --------- */
/* out[n,y,x,c] = sum_{kh,kw} in[n, y*stride - dil + kh*dil, x*stride - dil + kw*dil, c] * w[kh*3+kw][c] */
int dc_dw_fwd(dc_view in, const void* w9c, int stride, int dil, dc_view out, void* stream);
int dc_dw_bwd_data(dc_view dout, const void* w9c, int stride, int dil, dc_view din, int accumulate, void* stream);
/* fp32 gradient accumulated with atomics into a pre-zeroed buffer: param_layout = 0 -> G[9][C] (tap-major scratch),
 * param_layout = 1 -> the parameter's own [C][1][3][3] layout (no unpack needed afterwards) */
int dc_dw_bwd_weight(dc_view in, dc_view dout, int stride, int dil, float* G9c, int param_layout, void* stream);
/* Deterministic two-stage form: every block (cluster) stores its 9 x C partial sums into its own slice of ws (fp32,
 * dc_dw_bwd_weight_ws_elems() elements, contents irrelevant on entry), a second launch adds the slices to G9c in slice order. */
long long dc_dw_bwd_weight_ws_elems(dc_view in, dc_view dout, int stride, int dil);
int dc_dw_bwd_weight_det(dc_view in, dc_view dout, int stride, int dil, float* G9c, int param_layout, float* ws, long long ws_elems,
                         void* stream);
/* Backward-data of a stride-1 dilation-1 depthwise conv whose INPUT is a BatchNorm output a = relu(bn(y) [+ residual])
 * (the ReLU -> depthwise chain of every Block, DX:79-97): this call must be the LAST writer of dL/da.  It stores
 * g = (dL/da, plus the already stored part when accumulate) masked by a > 0 into `din`, and adds per channel sum(g) and
 * sum(g*y) to the zeroed BatchNorm backward workspace `bwd_ws`; dc_bn_bwd_apply_reduced then finishes BatchNorm backward
 * in one element-wise pass (no reduction pass, no one-pass barrier kernel).  act = the stored activation a as the mask
 * source (needed when the BatchNorm had a residual), or a null view: the mask is recomputed from y and the coefficients in
 * the forward workspace `fwd_ws`.  relu = 0: no mask.  Returns -2 when the tile does not fit (caller falls back). */
int dc_dw_bwd_data_bnred(dc_view dout, const void* w9c, dc_view din, int accumulate, dc_view y, dc_view act,
                         const void* fwd_ws, void* bwd_ws, int relu, void* stream);

/* ---- BatchNorm2d (+ReLU, +residual add) (normalizer, DX:70,129,283,348,399; relu DX:79,147; add DX:120) ---- */
/* Per-layer BatchNorm workspace, dc_bn_ws_bytes(C) bytes, ZEROED by the caller before dc_bn_stats / dc_bn_bwd_reduce:
 *   double sums[2][C] | float coef[4][C] | uint32 ticket[16].
 * The reduction kernels accumulate fp64 sums with atomics; the last block to finish converts them into fp32
 * per-channel coefficients (forward: scale, shift, mean, invstd; backward: A, B, D with dy = A*g + B*y + D), so the
 * element-wise kernels carry no double-precision prologue and no separate finalize launch exists. */
size_t dc_bn_ws_bytes(int C);
enum {
  DC_BN_RELU       = 1,   /* out = relu(...) */
  DC_BN_TRAIN      = 2,   /* batch statistics from the workspace `sums` */
  DC_BN_IDENTITY   = 4,   /* skip normalisation (pure relu / add) */
  DC_BN_RES_WRITE  = 8,   /* backward: residual gradient is written, not accumulated */
  DC_BN_SUMS_READY = 16,  /* forward apply, train mode: the workspace holds the raw batch sums (left there by
                             dc_conv_gemm_tc_bnstats); dc_bn_apply derives the coefficients itself, stores them for the
                             backward pass and updates the running statistics: no dc_bn_stats launch */
  DC_BN_MASK_FROM_Y = 32  /* backward, train mode, no residual: the ReLU mask (out > 0) is recomputed from y and the forward
                             coefficients kept in the forward workspace; `out` is not read and may be a null view */
};
typedef struct dc_bn_params {
  const float* gamma;  const float* beta;      /* [C] */
  float* running_mean; float* running_var;     /* [C]; updated by dc_bn_stats (may both be NULL) */
  const double* sums;                          /* forward workspace (train mode), see dc_bn_ws_bytes */
  double count;                                /* N*H*W */
  float momentum, eps;
  int32_t flags;
  int32_t reserved;
} dc_bn_params;
/* batch statistics of y into p->sums (zeroed workspace) + coefficients + running-statistics update */
int dc_bn_stats(const dc_bn_params* p, dc_view y, void* stream);
/* out = [relu]( bn(y) [+ residual] ); residual.ptr may be NULL.  Views: channel-contiguous, 16-byte aligned,
 * C % 8 == 0 (bf16) or C % 4 == 0 (fp32). */
int dc_bn_apply(const dc_bn_params* p, dc_view y, dc_view residual, dc_view out, void* stream);
/* The ReLU -> SeparableConv2d_same chain of a Block (DX:79-97) in one launch: out = depthwise3x3(a), a = [relu](bn(y)), stride 1,
 * dilation 1, train-mode BatchNorm whose batch sums are already in the workspace (DC_BN_TRAIN | DC_BN_SUMS_READY, as left by
 * dc_conv_gemm_tc_bnstats).  The kernel finalizes the coefficients like dc_bn_apply (publishing them for the backward pass and
 * updating the running statistics), applies them while the y tile sits in shared memory and never reads a back from memory;
 * `act` (same shape as y, or a null view) receives a, bit-identical to what dc_bn_apply would store.  Returns -2 when the
 * tile does not fit (the caller then runs dc_bn_apply + dc_dw_fwd). */
int dc_dw_fwd_bn(const dc_bn_params* p, dc_view y, const void* w9c, dc_view act, dc_view out, void* stream);
/* backward pass 1: g = dout * (out > 0 if RELU); sums of g and g*y into the zeroed workspace `rws`; the last block
 * writes dgamma/dbeta ([C] fp32, may be NULL) and the coefficients A, B, D. */
int dc_bn_bwd_reduce(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, void* rws, float* dgamma, float* dbeta,
                     void* stream);
/* backward pass 2: dy = A*g + B*y + D (= gamma*invstd*(g - mean(g) - xhat*mean(g*xhat))); optional dres (+)= g */
int dc_bn_bwd_apply(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, const void* rws,
                    dc_view dy, dc_view dres, void* stream);
/* BatchNorm backward for a gradient already masked and reduced by dc_dw_bwd_data_bnred (train mode): coefficients are
 * finalized inside the kernel from the workspace sums; writes dy, dgamma, dbeta and dres (+)= g. */
int dc_bn_bwd_apply_reduced(const dc_bn_params* p, dc_view g, dc_view y, const void* rws, dc_view dy, dc_view dres,
                            float* dgamma, float* dbeta, void* stream);
/* Split BatchNorm backward for L2-resident tensors: dc_bn_bwd_reduce with DC_BN_SUMS_READY in the flags accumulates the sums
 * only; this call finalizes them per block, applies the mask the flags ask for and writes dy, dres, dgamma, dbeta. */
int dc_bn_bwd_apply_finalize(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, const void* rws, dc_view dy,
                             dc_view dres, float* dgamma, float* dbeta, void* stream);
/* One-pass variants for tensors small enough to be held in shared memory across the GPU (dc_bn_onepass_ok): statistics
 * and normalisation (forward) / reduction and gradient (backward) in ONE launch with an inter-block barrier; the grid
 * never exceeds the SM count, so all blocks are co-resident.  Same workspace contract as the two-pass entry points. */
int dc_bn_onepass_ok(int C, long long npix, int dtype, int backward);
int dc_bn_fwd_onepass(const dc_bn_params* p, dc_view y, dc_view residual, dc_view out, void* stream);
int dc_bn_bwd_onepass(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, void* rws, dc_view dy, dc_view dres,
                      float* dgamma, float* dbeta, void* stream);
/* per-channel sum over n,h,w; the first n_out (<= C) sums are written to fp32 out_c (bias gradient of upsample.conv1.6,
 * DX:366; n_out < C when x carries padding channels); ws_c: C doubles of scratch */
int dc_channel_sum(dc_view x, double* ws_c, float* out_c, int n_out, void* stream);

/* ---- image-pooling branch (AdaptiveAvgPool2d DX:425, F.interpolate 1x1 -> HxW DX:450) ---------- */
int dc_gap_fwd(dc_view x, float* mean_nc, void* stream);                       /* [N][C] fp32 */
int dc_broadcast_hw(const float* src_nc, dc_view dst, void* stream);           /* dst[n,h,w,c] = src[n][c] */
int dc_reduce_hw(dc_view x, float* sum_nc, void* stream);                      /* sum over h,w -> [N][C] */
int dc_gap_bwd(const float* dmean_nc, dc_view dx, int accumulate, void* stream); /* dx (+)= dmean/(H*W) */
/* F.interpolate(mode='bilinear', align_corners=True) (InterpolationUpsampler, DX:327-331) between any two sizes and its
 * input gradient (gather form, deterministic).  Channels-last views, C % 4 == 0; out may be fp32 for a bf16 input. */
int dc_bilinear_fwd(dc_view in, dc_view out, void* stream);
int dc_bilinear_bwd(dc_view dout, dc_view din, int accumulate, void* stream);    /* din (+)= resize^T(dout) */

/* ---- weighted cross-entropy "fp_loss" (LS:28-52) ------------------------------------------------ */
/* logits: fp32 view [N,H,W,C] (any strides, e.g. NCHW); target int64 [N*H*W] contiguous;
 * loss_out[0] = mean over N*H*W of w[t]*(-log softmax(logit)[t])  (fp32).  acc: 1 double scratch. */
int dc_wce_fwd(dc_view logits, const int64_t* target, const float* class_w, double* acc,
               float* loss_out, void* stream);
/* dlogits[n,h,w,c] = (*gscale) * w[t]*(softmax_c - [c==t]) / (N*H*W) */
int dc_wce_bwd(dc_view logits, const int64_t* target, const float* class_w, const float* gscale,
               dc_view dlogits, void* stream);

/* ---- IoU metric (compute_score UT:32-60, argmax TR:376/406/458) --------------------------------- */
/* counts[0..C) = tp, [C..2C) = fp, [2C..3C) = fn  (int64, pre-zeroed by the caller or accumulated) */
int dc_iou_counts(const int64_t* pred, const int64_t* gt, int64_t numel, int num_classes,
                  int64_t* counts, void* stream);
/* argmax over c with first-max tie rule (torch.max) fused with the counters; pred_out may be NULL */
int dc_argmax_iou(dc_view logits, const int64_t* gt, int num_classes, int64_t* pred_out,
                  int64_t* counts, void* stream);
/* score = ((iou0 + iou1) + ...)/C in fp32, iou_j = tp/(tp+fp+fn) or 1 when the union is empty */
int dc_iou_finalize(const int64_t* counts, int num_classes, float* score_out, void* stream);

/* ---- gradient buckets (DDP all-reduce payload, TR:227) ------------------------------------------ */
/* x[i] *= s  (1/world_size after an NCCL SUM all-reduce) */
int dc_scale_f32(float* x, size_t count, float s, void* stream);

/* ---- optimizer (optim.Adam / optim.AdamW over net.parameters(), TR:213-220, step at TR:364) ----------- */
/* One launch for all parameters.  jobs_dev: DEVICE array; job i owns blocks [block_start, block_start + n_blocks).
 * Update rule and state of torch.optim.Adam (adamw = 0: L2 weight decay added to the gradient) or AdamW (adamw = 1:
 * decoupled decay); bias_c1 = 1 - beta1^t, bias_c2 = 1 - beta2^t for the step being taken (t >= 1).  Hyper-parameters are
 * doubles: the fp32 scalars of the update (1 - beta, lr / bias_c1, ...) are derived in double like torch does in Python. */
typedef struct dc_adam_job {
  float* p; const float* g; float* m; float* v;
  int64_t numel;
  int32_t block_start, n_blocks;
} dc_adam_job;
int dc_adam_step_multi(const dc_adam_job* jobs_dev, int njobs, int total_blocks, double lr, double beta1, double beta2,
                       double eps, double weight_decay, double bias_c1, double bias_c2, int adamw, void* stream);
/* LAMB step for all parameters (apex.optimizers.FusedLAMB as constructed at TR:217-218; algorithm of apex's
 * multi_tensor_lamb.cu: global gradient-norm clipping, Adam moments, per-tensor trust ratio).  Same job table as
 * dc_adam_step_multi; the update is written over the gradient; norms = device scratch of 1 + 2*njobs doubles. */
int dc_lamb_step_multi(const dc_adam_job* jobs_dev, int njobs, int total_blocks, double lr, double beta1, double beta2,
                       double eps, double weight_decay, double bias_c1, double bias_c2, int adam_w_mode,
                       int grad_averaging, double max_grad_norm, int use_nvlamb, double* norms, void* stream);
/* LARS step for all parameters (BASELINE.json configs[4]; not part of the reference: You et al. 2017, per-tensor trust ratio
 * on SGD with momentum).  Job table as above with m = momentum buffer (v unused); norms = device scratch of 2*njobs doubles. */
int dc_lars_step_multi(const dc_adam_job* jobs_dev, int njobs, int total_blocks, double lr, double momentum,
                       double weight_decay, double trust_coefficient, double eps, double* norms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPCAM_B200_H */
