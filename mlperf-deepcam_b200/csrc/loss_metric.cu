// Weighted cross-entropy "fp_loss" (LS:28-52) and the IoU metric (compute_score UT:32-60, argmax TR:376/406/458).
//
// fp_loss: nn.CrossEntropyLoss(weight, reduction='none') followed by two multiplications with matrices that are
// identically 1 (LS:41 and LS:46 test `eq(preds,k) & ne(preds,k)`, always false) and torch.mean over N*H*W
// (LS:50): loss = 1/(N*H*W) * sum_p w[t_p] * (logsumexp(x_p) - x_p[t_p]).  NOT divided by sum of weights.
// compute_score: integer tp/fp/fn counters per class; bit-exact.
// Both are HBM-bound single passes: algorithmic bytes = logits + targets (+ gradient written once).
#include "common.cuh"
#include <algorithm>

namespace dc {

__device__ __forceinline__ void decode3(long long p, int H, int W, int& n, int& h, int& w) {
  w = (int)(p % W);
  long long t = p / W;
  h = (int)(t % H);
  n = (int)(t / H);
}

__global__ void __launch_bounds__(256) wce_fwd_kernel(View<const float> x, const int64_t* __restrict__ target,
                                                      const float* __restrict__ cw, double* acc) {
  const long long npix = (long long)x.n * x.h * x.w;
  const int C = x.c;
  double local = 0.0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    int n, h, w;
    decode3(p, x.h, x.w, n, h, w);
    const float* px = x.p + n * x.sn + h * x.sh + w * x.sw;
    long long t = target[p];
    if (t == -100) continue;         // ignore_index of nn.CrossEntropyLoss (LS:35 keeps the default): zero loss, still counted in the mean
    if (t < 0 || t >= C) {           // any other label outside [0, C): torch device-asserts; here the loss becomes NaN (loud, no host sync)
      local += (double)__int_as_float(0x7fc00000);
      continue;
    }
    float m = px[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, px[c * x.sc]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(px[c * x.sc] - m);
    float lse = m + logf(s);
    float l = cw[t] * (lse - px[t * x.sc]);
    local += (double)l;
  }
  local = warp_sum(local);
  __shared__ double wsum[8];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) wsum[wid] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += wsum[i];
    atomicAdd(acc, s);
  }
}

__global__ void wce_finalize_kernel(const double* acc, double npix, float* loss_out) { loss_out[0] = (float)(acc[0] / npix); }

template <typename TD>
__global__ void __launch_bounds__(256) wce_bwd_kernel(View<const float> x, const int64_t* __restrict__ target,
                                                      const float* __restrict__ cw, const float* __restrict__ gscale,
                                                      View<TD> dx) {
  const long long npix = (long long)x.n * x.h * x.w;
  const int C = x.c;
  const float gs = (gscale ? gscale[0] : 1.0f) / (float)npix;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    int n, h, w;
    decode3(p, x.h, x.w, n, h, w);
    const float* px = x.p + n * x.sn + h * x.sh + w * x.sw;
    TD* pd = dx.p + n * dx.sn + h * dx.sh + w * dx.sw;
    long long t = target[p];
    const bool valid = (t >= 0 && t < C);
    const bool poison = !valid && t != -100;     // corrupted label (not ignore_index): NaN gradient, like the NaN loss of wce_fwd
    float m = px[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, px[c * x.sc]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(px[c * x.sc] - m);
    float inv = 1.0f / s;
    float wt = valid ? cw[t] * gs : (poison ? __int_as_float(0x7fc00000) : 0.f);
    for (int c = 0; c < dx.c; ++c) {
      float g = 0.f;
      if (c < C) {
        float sm = expf(px[c * x.sc] - m) * inv;
        g = wt * (sm - (c == t ? 1.f : 0.f));
      }
      elem<TD>::st(pd + c * dx.sc, g);
    }
  }
}

// ---- IoU counters ------------------------------------------------------------------------------
// tp[j] = #(pred == gt && gt == j); fp[j] = #(pred != gt && pred == j); fn[j] = #(pred != gt && gt == j)
template <int MAXC>
__device__ __forceinline__ void count_one(long long pr, long long g, int C, unsigned (&tp)[MAXC], unsigned (&fp)[MAXC], unsigned (&fn)[MAXC]) {
  const bool eq = (pr == g);
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    if (j < C) {
      tp[j] += (eq && g == j) ? 1u : 0u;
      fp[j] += (!eq && pr == j) ? 1u : 0u;
      fn[j] += (!eq && g == j) ? 1u : 0u;
    }
  }
}

template <int MAXC>
__device__ __forceinline__ void flush_counts(int C, unsigned (&tp)[MAXC], unsigned (&fp)[MAXC], unsigned (&fn)[MAXC], int64_t* counts) {
  __shared__ unsigned long long sh[3 * MAXC];
  if (threadIdx.x < 3 * MAXC) sh[threadIdx.x] = 0ull;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    if (j < C) {
      unsigned a = __reduce_add_sync(0xffffffffu, tp[j]);
      unsigned b = __reduce_add_sync(0xffffffffu, fp[j]);
      unsigned c = __reduce_add_sync(0xffffffffu, fn[j]);
      if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(&sh[j], (unsigned long long)a);
        if (b) atomicAdd(&sh[MAXC + j], (unsigned long long)b);
        if (c) atomicAdd(&sh[2 * MAXC + j], (unsigned long long)c);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 * MAXC) {
    int kind = threadIdx.x / MAXC, j = threadIdx.x % MAXC;
    if (j < C && sh[threadIdx.x]) atomicAdd(reinterpret_cast<unsigned long long*>(counts) + kind * C + j, sh[threadIdx.x]);
  }
}

template <int MAXC>
__global__ void __launch_bounds__(256) iou_counts_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ gt,
                                                         long long numel, int C, int64_t* counts) {
  unsigned tp[MAXC] = {}, fp[MAXC] = {}, fn[MAXC] = {};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x)
    count_one<MAXC>(pred[i], gt[i], C, tp, fp, fn);
  flush_counts<MAXC>(C, tp, fp, fn, counts);
}

// generic class count: global atomics per element that is not (class-0 true positive heavy) -- slow path, C > 8
__global__ void iou_counts_generic_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ gt, long long numel,
                                          int C, int64_t* counts) {
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    long long pr = pred[i], g = gt[i];
    if (pr == g) { if (g >= 0 && g < C) atomicAdd(cnt + g, 1ull); }
    else {
      if (pr >= 0 && pr < C) atomicAdd(cnt + C + pr, 1ull);
      if (g >= 0 && g < C) atomicAdd(cnt + 2 * C + g, 1ull);
    }
  }
}

template <int MAXC>
__global__ void __launch_bounds__(256) argmax_iou_kernel(View<const float> x, const int64_t* __restrict__ gt, int C,
                                                         int64_t* __restrict__ pred_out, int64_t* counts) {
  unsigned tp[MAXC] = {}, fp[MAXC] = {}, fn[MAXC] = {};
  const long long npix = (long long)x.n * x.h * x.w;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    int n, h, w;
    decode3(p, x.h, x.w, n, h, w);
    const float* px = x.p + n * x.sn + h * x.sh + w * x.sw;
    // torch.max(dim) returns the first maximal index; NaN propagates as the maximum in torch
    float best = px[0];
    int bi = 0;
    for (int c = 1; c < x.c; ++c) {
      float v = px[c * x.sc];
      if (v > best || (v != v && best == best)) { best = v; bi = c; }
    }
    if (pred_out) pred_out[p] = bi;
    if (gt) count_one<MAXC>((long long)bi, gt[p], C, tp, fp, fn);
  }
  if (gt) flush_counts<MAXC>(C, tp, fp, fn, counts);
}

__global__ void iou_finalize_kernel(const int64_t* counts, int C, float* score) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // UT:52-60: iou[j] = tp.float()/union.float() or 1.0; sum(iou) starts from python int 0; / float(num_classes)
  float total = 0.f;
  for (int j = 0; j < C; ++j) {
    long long tp = counts[j], fp = counts[C + j], fn = counts[2 * C + j];
    long long uni = tp + fp + fn;
    float iou = (uni == 0) ? 1.0f : __fdiv_rn((float)tp, (float)uni);
    total = __fadd_rn(total, iou);
  }
  score[0] = __fdiv_rn(total, (float)C);
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_wce_fwd(dc_view logits, const int64_t* target, const float* class_w, double* acc, float* loss_out, void* stream) {
  DC_REQUIRE(view_ok(logits) && logits.dtype == DC_F32, "dc_wce_fwd: logits must be an fp32 view");
  DC_REQUIRE(target && class_w && acc && loss_out, "dc_wce_fwd: null argument");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double), st);
  if (e != cudaSuccess) return dc::fail((int)e, "dc_wce_fwd: %s", cudaGetErrorString(e));
  long long npix = (long long)logits.n * logits.h * logits.w;
  int blocks = (int)std::min<long long>((npix + 255) / 256, (long long)kNumSMs * 8);
  wce_fwd_kernel<<<blocks, 256, 0, st>>>(make_view<const float>(logits), target, class_w, acc);
  wce_finalize_kernel<<<1, 1, 0, st>>>(acc, (double)npix, loss_out);
  return launch_status("dc_wce_fwd");
}

int dc_wce_bwd(dc_view logits, const int64_t* target, const float* class_w, const float* gscale, dc_view dlogits, void* stream) {
  DC_REQUIRE(view_ok(logits) && logits.dtype == DC_F32, "dc_wce_bwd: logits must be an fp32 view");
  DC_REQUIRE(view_ok(dlogits) && dlogits.n == logits.n && dlogits.h == logits.h && dlogits.w == logits.w && dlogits.c >= logits.c,
             "dc_wce_bwd: gradient view mismatch");
  DC_REQUIRE(target && class_w, "dc_wce_bwd: null argument");
  cudaStream_t st = as_stream(stream);
  long long npix = (long long)logits.n * logits.h * logits.w;
  int blocks = (int)std::min<long long>((npix + 255) / 256, (long long)kNumSMs * 8);
  if (dlogits.dtype == DC_F32)
    wce_bwd_kernel<float><<<blocks, 256, 0, st>>>(make_view<const float>(logits), target, class_w, gscale, make_view<float>(dlogits));
  else
    wce_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(make_view<const float>(logits), target, class_w, gscale, make_view<__nv_bfloat16>(dlogits));
  return launch_status("dc_wce_bwd");
}

int dc_iou_counts(const int64_t* pred, const int64_t* gt, int64_t numel, int num_classes, int64_t* counts, void* stream) {
  DC_REQUIRE(pred && gt && counts && numel >= 0 && num_classes > 0, "dc_iou_counts: bad arguments");
  if (numel == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int blocks = (int)std::min<long long>((numel + 255) / 256, (long long)kNumSMs * 8);
  if (num_classes <= 4) iou_counts_kernel<4><<<blocks, 256, 0, st>>>(pred, gt, numel, num_classes, counts);
  else if (num_classes <= 8) iou_counts_kernel<8><<<blocks, 256, 0, st>>>(pred, gt, numel, num_classes, counts);
  else iou_counts_generic_kernel<<<blocks, 256, 0, st>>>(pred, gt, numel, num_classes, counts);
  return launch_status("dc_iou_counts");
}

int dc_argmax_iou(dc_view logits, const int64_t* gt, int num_classes, int64_t* pred_out, int64_t* counts, void* stream) {
  DC_REQUIRE(view_ok(logits) && logits.dtype == DC_F32, "dc_argmax_iou: logits must be an fp32 view");
  DC_REQUIRE(num_classes > 0 && num_classes <= 8, "dc_argmax_iou: num_classes must be in [1, 8]");
  DC_REQUIRE(gt == nullptr || counts != nullptr, "dc_argmax_iou: counts required when gt is given");
  DC_REQUIRE(gt != nullptr || pred_out != nullptr, "dc_argmax_iou: nothing to do");
  cudaStream_t st = as_stream(stream);
  long long npix = (long long)logits.n * logits.h * logits.w;
  int blocks = (int)std::min<long long>((npix + 255) / 256, (long long)kNumSMs * 8);
  if (num_classes <= 4) argmax_iou_kernel<4><<<blocks, 256, 0, st>>>(make_view<const float>(logits), gt, num_classes, pred_out, counts);
  else argmax_iou_kernel<8><<<blocks, 256, 0, st>>>(make_view<const float>(logits), gt, num_classes, pred_out, counts);
  return launch_status("dc_argmax_iou");
}

int dc_iou_finalize(const int64_t* counts, int num_classes, float* score_out, void* stream) {
  DC_REQUIRE(counts && score_out && num_classes > 0, "dc_iou_finalize: bad arguments");
  iou_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(counts, num_classes, score_out);
  return launch_status("dc_iou_finalize");
}

}  // extern "C"
