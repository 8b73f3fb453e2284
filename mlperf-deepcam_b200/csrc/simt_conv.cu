// SIMT gather-GEMM: the fp32-accumulate CUDA-core path for every dense contraction of the network.
// Used (a) for the whole network in fp32 parity mode and (b) in bf16 mode for the layers the tcgen05
// kernel does not take (C_in = 16 first conv DX:145, C_out = 3 last deconv DX:374, the 2-row image-pooling
// 1x1 conv DX:426).  One descriptor (dc_conv_desc) expresses Conv2d fprop, Conv2d dgrad (taps negated),
// strided 1x1, ConvTranspose2d fprop (per output-parity class) and ConvTranspose2d dgrad.
//
//   fprop-like: out[m, co] (+)= bias[co] + sum_t sum_ci in[pix(m,t), ci] * W[wt[t]][ci][co]
//   wgrad     : G[wt[t]][co][ci]  += sum_m in[pix(m,t), ci] * dout[m, co]
// where m enumerates the pixels (n, y, x) of the out/dout view and pix(m,t) = (n, y*sh + dh[t], x*sw + dw[t]).
#include "common.cuh"
#include <algorithm>

namespace dc {

constexpr int BK = 16;

struct OutView {   // output with runtime dtype and generic strides
  void* p;
  int n, h, w, c;
  long long sn, sh, sw, sc;
  int dtype;
};
static inline OutView make_out(const dc_view& v) {
  OutView o;
  o.p = v.ptr; o.n = v.n; o.h = v.h; o.w = v.w; o.c = v.c;
  o.sn = v.sn; o.sh = v.sh; o.sw = v.sw; o.sc = v.sc; o.dtype = v.dtype;
  return o;
}

template <typename T, int BM, int BN>
__global__ void __launch_bounds__(256) conv_gemm_simt_kernel(dc_conv_desc d, View<const T> in, const T* __restrict__ W,
                                                             const float* __restrict__ bias, OutView out, int co_pad) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int LA = (BM * 4 + 255) / 256;          // A vec4 loads per thread per chunk
  constexpr int LB = (BK * BN / 4 + 255) / 256;     // B vec4 loads per thread per chunk
  __shared__ float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int M = out.n * out.h * out.w;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ci = in.c, Co = out.c;

  // rows this thread loads for the A tile
  int a_m[LA], a_kv[LA], a_n[LA], a_y[LA], a_x[LA];
  bool a_ok[LA];
#pragma unroll
  for (int i = 0; i < LA; ++i) {
    int idx = tid + i * 256;
    a_m[i] = idx / 4;
    a_kv[i] = idx % 4;
    int m = m0 + a_m[i];
    a_ok[i] = (idx < BM * 4) && (m < M);
    int mm = a_ok[i] ? m : 0;
    a_x[i] = mm % out.w;
    int t = mm / out.w;
    a_y[i] = t % out.h;
    a_n[i] = t / out.h;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < d.ntaps; ++t) {
    const T* a_ptr[LA];
    bool a_in[LA];
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int ih = a_y[i] * d.stride_h + d.dh[t];
      int iw = a_x[i] * d.stride_w + d.dw[t];
      a_in[i] = a_ok[i] && ih >= 0 && ih < in.h && iw >= 0 && iw < in.w;
      a_ptr[i] = a_in[i] ? in.at(a_n[i], ih, iw) : in.p;
    }
    const T* Wt = W + (size_t)d.wt[t] * Ci * co_pad;
    for (int c0 = 0; c0 < Ci; c0 += BK) {
      // ---- load A (transpose into k-major) ----
#pragma unroll
      for (int i = 0; i < LA; ++i) {
        int idx = tid + i * 256;
        if (idx < BM * 4) {
          int c = c0 + a_kv[i] * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a_in[i] && c < Ci) v = elem<T>::ld4(a_ptr[i] + c);
          As[a_kv[i] * 4 + 0][a_m[i]] = v.x;
          As[a_kv[i] * 4 + 1][a_m[i]] = v.y;
          As[a_kv[i] * 4 + 2][a_m[i]] = v.z;
          As[a_kv[i] * 4 + 3][a_m[i]] = v.w;
        }
      }
      // ---- load B ----
#pragma unroll
      for (int i = 0; i < LB; ++i) {
        int idx = tid + i * 256;
        if (idx < BK * BN / 4) {
          int k = idx / (BN / 4), nv = idx % (BN / 4);
          int c = c0 + k, co = n0 + nv * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < Ci && co < co_pad) v = elem<T>::ld4(Wt + (size_t)c * co_pad + co);
          *reinterpret_cast<float4*>(&Bs[k][nv * 4]) = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
    int x = m % out.w;
    int tt = m / out.w;
    int y = tt % out.h;
    int n = tt / out.h;
    long long base = n * out.sn + y * out.sh + x * out.sw;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int co = n0 + tx * TN + j;
      if (co >= Co) continue;
      float v = acc[i][j];
      if (bias) v += bias[co];
      long long off = base + ((d.out_csplit > 0 && co >= d.out_csplit) ? d.out_split_off + (co - d.out_csplit) * out.sc : co * out.sc);
      if (out.dtype == DC_F32) {
        float* p = reinterpret_cast<float*>(out.p) + off;
        if (d.accumulate) v += *p;
        *p = v;
      } else {
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(out.p) + off;
        if (d.accumulate) v += __bfloat162float(*p);
        *p = __float2bfloat16_rn(v);
      }
    }
  }
}


// Few-row GEMM for the image-pooling branch (DX:425-428: a [N_batch, 2048] x [2048, 256] product and its transpose in
// backward): the tiled kernel above would run it on 4 blocks for 140 us.  32 output columns x 32 k-lanes per block, fixed
// reduction order (no atomics: this branch feeds a BatchNorm over N_batch values, which amplifies any run-to-run noise).
constexpr int kSmallM = 8;
__global__ void __launch_bounds__(1024) small_m_gemm_kernel(const float* __restrict__ in, long long in_row_stride, int M, int K,
                                                            const float* __restrict__ W, int n_pad, const float* __restrict__ bias,
                                                            float* __restrict__ out, long long out_row_stride, int N, int accumulate) {
  __shared__ float red[32][kSmallM][32];
  const int nl = threadIdx.x & 31, kl = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + nl;
  float acc[kSmallM];
#pragma unroll
  for (int m = 0; m < kSmallM; ++m) acc[m] = 0.f;
  if (n < n_pad) {
#pragma unroll 4
    for (int k = kl; k < K; k += 32) {
      const float w = W[(size_t)k * n_pad + n];
#pragma unroll
      for (int m = 0; m < kSmallM; ++m)
        if (m < M) acc[m] = fmaf(in[m * in_row_stride + k], w, acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < kSmallM; ++m) red[kl][m][nl] = acc[m];
  __syncthreads();
  if (kl < M && n < N) {            // warp kl finishes output row m = kl
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) s += red[q][kl][nl];
    if (bias) s += bias[n];
    float* o = out + kl * out_row_stride + n;
    if (accumulate) s += *o;
    *o = s;
  }
}

// wgrad: rows = co (BMC tile, from dout), cols = ci (64 tile, from in), reduction over pixels in chunks of BK.
template <typename T, int BMC>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(dc_conv_desc d, View<const T> in, View<const T> dout,
                                                              float* __restrict__ G, int pix_per_split, int n_ci_tiles, int det_wtaps) {
  // det_wtaps > 0 (dc_conv_wgrad_simt_det): G is a workspace [splits][wtaps][Co][Ci]; pixel split z STORES its partial sums into
  // slice z (one writer per element) and launch_split_reduce adds the slices in split order
  constexpr int BNC = 64;
  constexpr int TM = BMC / 16, TN = BNC / 16;
  __shared__ __align__(16) float As[BK][BMC];   // [pixel][co]
  __shared__ __align__(16) float Bs[BK][BNC];   // [pixel][ci]
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int Ci = in.c, Co = dout.c;
  const int co0 = (blockIdx.x / n_ci_tiles) * BMC;
  const int ci0 = (blockIdx.x % n_ci_tiles) * BNC;
  const int t = blockIdx.y;
  const int M = dout.n * dout.h * dout.w;
  const int m_begin = blockIdx.z * pix_per_split;
  const int m_end = min(M, m_begin + pix_per_split);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // load mapping: A: BK x BMC/4 vec4 ; B: BK x 16 vec4 (= 256, one per thread)
  const int b_k = tid / 16, b_v = tid % 16;
  constexpr int AV = BMC / 4;
  const int a_k = tid / AV, a_v = tid % AV;
  const bool a_thr = tid < BK * AV;

  for (int mc = m_begin; mc < m_end; mc += BK) {
    {  // B tile: in[pix(m,t), ci0 + b_v*4 ..]
      int m = mc + b_k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end) {
        int x = m % dout.w;
        int tt = m / dout.w;
        int y = tt % dout.h;
        int n = tt / dout.h;
        int ih = y * d.stride_h + d.dh[t];
        int iw = x * d.stride_w + d.dw[t];
        int c = ci0 + b_v * 4;
        if (ih >= 0 && ih < in.h && iw >= 0 && iw < in.w && c < Ci) v = elem<T>::ld4(in.at(n, ih, iw) + c);
      }
      *reinterpret_cast<float4*>(&Bs[b_k][b_v * 4]) = v;
    }
    if (a_thr) {  // A tile: dout[m, co0 + a_v*4 ..]
      int m = mc + a_k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end) {
        int x = m % dout.w;
        int tt = m / dout.w;
        int y = tt % dout.h;
        int n = tt / dout.h;
        int c = co0 + a_v * 4;
        if (c < Co) v = elem<T>::ld4(dout.at(n, y, x) + c);
      }
      *reinterpret_cast<float4*>(&As[a_k][a_v * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* Gt = G + (size_t)(det_wtaps > 0 ? (int)blockIdx.z * det_wtaps + d.wt[t] : d.wt[t]) * Co * Ci;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int co = co0 + ty * TM + i;
    if (co >= Co) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int ci = ci0 + tx * TN + j;
      if (ci >= Ci) continue;
      if (det_wtaps > 0) Gt[(size_t)co * Ci + ci] = acc[i][j];
      else atomicAdd(Gt + (size_t)co * Ci + ci, acc[i][j]);
    }
  }
}

static int check_desc(const char* what, const dc_conv_desc* d) {
  DC_REQUIRE(d != nullptr, "%s: null descriptor", what);
  DC_REQUIRE(d->ntaps >= 1 && d->ntaps <= DC_MAX_TAPS, "%s: ntaps=%d out of range", what, d->ntaps);
  DC_REQUIRE(d->stride_h >= 1 && d->stride_w >= 1, "%s: bad stride", what);
  DC_REQUIRE(d->wtaps >= 1 && d->wtaps <= DC_MAX_TAPS, "%s: wtaps=%d out of range", what, d->wtaps);
  for (int t = 0; t < d->ntaps; ++t) DC_REQUIRE(d->wt[t] >= 0 && d->wt[t] < d->wtaps, "%s: wt[%d]=%d out of range", what, t, d->wt[t]);
  DC_REQUIRE(d->out_csplit >= 0 && d->out_csplit % 4 == 0, "%s: out_csplit must be a non-negative multiple of 4", what);
  return 0;
}

template <typename T>
static int conv_gemm_simt_t(const dc_conv_desc* d, const dc_view& in, const void* w, const float* bias, const dc_view& out, cudaStream_t st) {
  const int M = out.n * out.h * out.w;
  const int co_pad = (out.c + 3) & ~3;
  OutView o = make_out(out);
  View<const T> iv = make_view<const T>(in);
  if (out.c <= 16) {
    dim3 grid(ceil_div(M, 128), ceil_div(out.c, 16));
    conv_gemm_simt_kernel<T, 128, 16><<<grid, 256, 0, st>>>(*d, iv, (const T*)w, bias, o, co_pad);
  } else if (out.c <= 32) {
    dim3 grid(ceil_div(M, 128), ceil_div(out.c, 32));
    conv_gemm_simt_kernel<T, 128, 32><<<grid, 256, 0, st>>>(*d, iv, (const T*)w, bias, o, co_pad);
  } else {
    dim3 grid(ceil_div(M, 64), ceil_div(out.c, 64));
    conv_gemm_simt_kernel<T, 64, 64><<<grid, 256, 0, st>>>(*d, iv, (const T*)w, bias, o, co_pad);
  }
  return launch_status("dc_conv_gemm_simt");
}

static void wgrad_simt_plan(const dc_conv_desc* d, const dc_view& in, const dc_view& dout, int& bmc, int& pix_per_split, int& splits) {
  const int M = dout.n * dout.h * dout.w;
  bmc = dout.c <= 16 ? 16 : 64;
  const int tiles = ceil_div(dout.c, bmc) * ceil_div(in.c, 64) * d->ntaps;
  // split the pixel reduction so that the grid has ~4 waves of blocks
  splits = std::max(1, std::min(ceil_div(kNumSMs * 4, tiles), ceil_div(M, 256)));
  pix_per_split = ceil_div(ceil_div(M, splits), BK) * BK;
  splits = ceil_div(M, pix_per_split);
}

template <typename T>
static int conv_wgrad_simt_t(const dc_conv_desc* d, const dc_view& in, const dc_view& dout, float* G, float* ws, long long ws_elems,
                             cudaStream_t st) {
  int bmc, pix_per_split, splits;
  wgrad_simt_plan(d, in, dout, bmc, pix_per_split, splits);
  const int n_co_tiles = ceil_div(dout.c, bmc), n_ci_tiles = ceil_div(in.c, 64);
  const long long per_tap = (long long)dout.c * in.c, slice = per_tap * d->wtaps;
  const bool det = ws != nullptr && splits > 1;
  if (det) {
    DC_REQUIRE(ws_elems >= slice * splits, "dc_conv_wgrad_simt_det: workspace of %lld floats required, %lld given", slice * splits, ws_elems);
    for (int a = 0; a < d->ntaps; ++a)
      for (int b = a + 1; b < d->ntaps; ++b)
        DC_REQUIRE(d->wt[a] != d->wt[b], "dc_conv_wgrad_simt_det: two taps of one launch share weight tap %d", d->wt[a]);
  }
  float* target = det ? ws : G;
  const int det_wtaps = det ? d->wtaps : 0;
  dim3 grid(n_co_tiles * n_ci_tiles, d->ntaps, splits);
  if (bmc == 16)
    conv_wgrad_simt_kernel<T, 16><<<grid, 256, 0, st>>>(*d, make_view<const T>(in), make_view<const T>(dout), target, pix_per_split, n_ci_tiles,
                                                        det_wtaps);
  else
    conv_wgrad_simt_kernel<T, 64><<<grid, 256, 0, st>>>(*d, make_view<const T>(in), make_view<const T>(dout), target, pix_per_split, n_ci_tiles,
                                                        det_wtaps);
  if (int r = launch_status("dc_conv_wgrad_simt")) return r;
  if (!det) return 0;
  SplitReduceTaps taps;
  taps.n = d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) taps.wt[t] = d->wt[t];
  return launch_split_reduce(ws, splits, slice, taps, per_tap, G, st);
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_conv_gemm_simt(const dc_conv_desc* d, dc_view in, const void* w, const float* bias, dc_view out, void* stream) {
  if (int r = check_desc("dc_conv_gemm_simt", d)) return r;
  DC_REQUIRE(view_ok(in) && view_vec4(in), "dc_conv_gemm_simt: input view must be channel-contiguous with C %% 4 == 0");
  DC_REQUIRE(view_ok(out) && out.n == in.n, "dc_conv_gemm_simt: bad output view");
  DC_REQUIRE(w != nullptr && (reinterpret_cast<uintptr_t>(w) % 16) == 0, "dc_conv_gemm_simt: weights must be 16-byte aligned");
  DC_REQUIRE((long long)out.n * out.h * out.w < (1ll << 31), "dc_conv_gemm_simt: too many pixels");
  cudaStream_t st = as_stream(stream);
  // few-row fast path: 1x1, no gather offset, fp32 in and out, pixel-linear rows (the image-pooling conv and its dgrad)
  const long long M = (long long)out.n * out.h * out.w;
  if (M <= kSmallM && d->out_csplit == 0 && d->ntaps == 1 && d->dh[0] == 0 && d->dw[0] == 0 && d->stride_h == 1 && d->stride_w == 1 && d->wt[0] == 0 &&
      in.dtype == DC_F32 && out.dtype == DC_F32 && in.sc == 1 && out.sc == 1 && in.h == out.h && in.w == out.w &&
      in.sh == (long long)in.w * in.sw && in.sn == (long long)in.h * in.sh && out.sh == (long long)out.w * out.sw &&
      out.sn == (long long)out.h * out.sh) {
    const int K = in.c, N = out.c, n_pad = (N + 3) / 4 * 4;
    small_m_gemm_kernel<<<ceil_div(n_pad, 32), 1024, 0, st>>>(reinterpret_cast<const float*>(in.ptr), in.sw, (int)M, K,
                                                             reinterpret_cast<const float*>(w), n_pad, bias,
                                                             reinterpret_cast<float*>(out.ptr), out.sw, N, d->accumulate);
    return launch_status("dc_conv_gemm_simt");
  }
  return in.dtype == DC_F32 ? conv_gemm_simt_t<float>(d, in, w, bias, out, st) : conv_gemm_simt_t<__nv_bfloat16>(d, in, w, bias, out, st);
}

static int conv_wgrad_simt_impl(const char* what, const dc_conv_desc* d, dc_view in, dc_view dout, float* G, float* ws, long long ws_elems,
                                void* stream) {
  if (int r = check_desc(what, d)) return r;
  DC_REQUIRE(view_ok(in) && view_vec4(in) && view_ok(dout) && view_vec4(dout), "%s: views must be channel-contiguous, C %% 4 == 0", what);
  DC_REQUIRE(in.dtype == dout.dtype && in.n == dout.n, "%s: dtype/batch mismatch", what);
  DC_REQUIRE(G != nullptr, "%s: null gradient", what);
  cudaStream_t st = as_stream(stream);
  return in.dtype == DC_F32 ? conv_wgrad_simt_t<float>(d, in, dout, G, ws, ws_elems, st)
                            : conv_wgrad_simt_t<__nv_bfloat16>(d, in, dout, G, ws, ws_elems, st);
}

int dc_conv_wgrad_simt(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, void* stream) {
  return conv_wgrad_simt_impl("dc_conv_wgrad_simt", d, in, dout, G, nullptr, 0, stream);
}

long long dc_conv_wgrad_simt_ws_elems(const dc_conv_desc* d, dc_view in, dc_view dout) {
  if (d == nullptr || d->ntaps < 1 || d->ntaps > DC_MAX_TAPS || !view_ok(in) || !view_ok(dout)) return -1;
  int bmc, pps, splits;
  wgrad_simt_plan(d, in, dout, bmc, pps, splits);
  return splits > 1 ? (long long)splits * d->wtaps * dout.c * in.c : 0;
}

/* deterministic form: partial sums per pixel split in ws, added to G in split order by a second launch */
int dc_conv_wgrad_simt_det(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, float* ws, long long ws_elems, void* stream) {
  DC_REQUIRE(ws != nullptr || dc_conv_wgrad_simt_ws_elems(d, in, dout) == 0, "dc_conv_wgrad_simt_det: workspace required (dc_conv_wgrad_simt_ws_elems)");
  return conv_wgrad_simt_impl("dc_conv_wgrad_simt_det", d, in, dout, G, ws, ws_elems, stream);
}

}  // extern "C"
