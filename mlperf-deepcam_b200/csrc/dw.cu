// Depthwise 3x3 convolution of SeparableConv2d_same (DX:54-66): fixed_padding (DX:45-51) is folded into
// the index arithmetic (implicit halo, no padded copy), groups = C, no bias, stride 1|2, dilation 1|2.
//   fwd : out[n,y,x,c]  = sum_{kh,kw} in[n, y*s - d + kh*d, x*s - d + kw*d, c] * w[kh*3+kw][c]
//   bwdD: din[n,h,w,c]  = sum_{kh,kw} dout[n, (h + d - kh*d)/s, (w + d - kw*d)/s, c] * w[kh*3+kw][c]   (when divisible)
//   bwdW: G[kh*3+kw][c] = sum_{n,y,x} in[n, y*s - d + kh*d, x*s - d + kw*d, c] * dout[n,y,x,c]
// HBM-bound (AI ~ 4 FLOP/B): algorithmic bytes = in + out (+ 9*C weights); neighbours are served from L1/L2.
#include "common.cuh"
#include <algorithm>

namespace dc {

__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
  a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y); a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
}

// One thread = one output pixel x 4 channels.  Threads of a warp cover consecutive channel vectors of the
// same pixel, consecutive warps cover consecutive pixels along w (so the 3x3 halo hits L1).
template <typename T>
__global__ void __launch_bounds__(256) dw_fwd_kernel(View<const T> in, const T* __restrict__ w9c, int s, int d, View<T> out) {
  const int cv = out.c >> 2;
  const long long total = (long long)out.n * out.h * out.w * cv;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < total;
       item += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(item % cv);
    int pix = (int)(item / cv);
    int x = pix % out.w;
    int t = pix / out.w;
    int y = t % out.h;
    int n = t / out.h;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int ih = y * s - d + kh * d;
      if (ih < 0 || ih >= in.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int iw = x * s - d + kw * d;
        if (iw < 0 || iw >= in.w) continue;
        float4 v = elem<T>::ld4(in.at(n, ih, iw) + c4 * 4);
        float4 wv = elem<T>::ld4(w9c + (kh * 3 + kw) * out.c + c4 * 4);
        fma4(acc, v, wv);
      }
    }
    elem<T>::st4(out.at(n, y, x) + c4 * 4, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dw_bwd_data_kernel(View<const T> dout, const T* __restrict__ w9c, int s, int d,
                                                          View<T> din, int accumulate) {
  const int cv = din.c >> 2;
  const long long total = (long long)din.n * din.h * din.w * cv;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < total;
       item += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(item % cv);
    int pix = (int)(item / cv);
    int x = pix % din.w;
    int t = pix / din.w;
    int y = t % din.h;
    int n = t / din.h;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int ty = y + d - kh * d;
      if (ty < 0 || (ty % s) != 0) continue;
      int oy = ty / s;
      if (oy >= dout.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int tx = x + d - kw * d;
        if (tx < 0 || (tx % s) != 0) continue;
        int ox = tx / s;
        if (ox >= dout.w) continue;
        float4 v = elem<T>::ld4(dout.at(n, oy, ox) + c4 * 4);
        float4 wv = elem<T>::ld4(w9c + (kh * 3 + kw) * din.c + c4 * 4);
        fma4(acc, v, wv);
      }
    }
    T* dp = din.at(n, y, x) + c4 * 4;
    if (accumulate) {
      float4 old = elem<T>::ld4(dp);
      acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
    }
    elem<T>::st4(dp, acc);
  }
}

// Weight gradient: channel lanes x pixel lanes (see bn.cu), 9 taps x 4 channels of fp32 partials per thread,
// block reduction through shared memory, one fp32 atomicAdd per (tap, channel) per block.
template <typename T>
__global__ void __launch_bounds__(256) dw_bwd_weight_kernel(View<const T> in, View<const T> dout, int s, int d,
                                                            float* __restrict__ G, int cvb, int rows) {
  extern __shared__ float redf[];   // [rows][cvb*4*9]
  const int tx = threadIdx.x % cvb, ty = threadIdx.x / cvb;
  const int c4 = blockIdx.y * cvb + tx;
  const int C = dout.c;
  const bool ok = (ty < rows) && (c4 * 4 < C);
  const int npix = dout.n * dout.h * dout.w;
  float4 acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    for (int pix = blockIdx.x * rows + ty; pix < npix; pix += gridDim.x * rows) {
      int x = pix % dout.w;
      int t = pix / dout.w;
      int y = t % dout.h;
      int n = t / dout.h;
      float4 g = elem<T>::ld4(dout.at(n, y, x) + c4 * 4);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        int ih = y * s - d + kh * d;
        if (ih < 0 || ih >= in.h) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          int iw = x * s - d + kw * d;
          if (iw < 0 || iw >= in.w) continue;
          float4 v = elem<T>::ld4(in.at(n, ih, iw) + c4 * 4);
          fma4(acc[kh * 3 + kw], v, g);
        }
      }
    }
  }
  const int per_row = cvb * 36;
  if (ty < rows) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float* r = redf + ty * per_row + (k * cvb + tx) * 4;
      r[0] = ok ? acc[k].x : 0.f; r[1] = ok ? acc[k].y : 0.f; r[2] = ok ? acc[k].z : 0.f; r[3] = ok ? acc[k].w : 0.f;
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < per_row; col += blockDim.x) {
    float sum = 0.f;
    for (int r = 0; r < rows; ++r) sum += redf[r * per_row + col];
    int k = col / (cvb * 4);
    int rem = col - k * cvb * 4;
    int c = (blockIdx.y * cvb + (rem >> 2)) * 4 + (rem & 3);
    if (c < C) atomicAdd(G + k * C + c, sum);
  }
}

template <typename T>
static int dw_fwd_t(const dc_view& in, const void* w, int s, int d, const dc_view& out, cudaStream_t st) {
  long long total = (long long)out.n * out.h * out.w * (out.c / 4);
  int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 32);
  dw_fwd_kernel<T><<<blocks, 256, 0, st>>>(make_view<const T>(in), (const T*)w, s, d, make_view<T>(out));
  return launch_status("dc_dw_fwd");
}
template <typename T>
static int dw_bwd_data_t(const dc_view& dout, const void* w, int s, int d, const dc_view& din, int acc, cudaStream_t st) {
  long long total = (long long)din.n * din.h * din.w * (din.c / 4);
  int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 32);
  dw_bwd_data_kernel<T><<<blocks, 256, 0, st>>>(make_view<const T>(dout), (const T*)w, s, d, make_view<T>(din), acc);
  return launch_status("dc_dw_bwd_data");
}
template <typename T>
static int dw_bwd_weight_t(const dc_view& in, const dc_view& dout, int s, int d, float* G, cudaStream_t st) {
  int cv = dout.c / 4;
  int cvb = std::min(cv, 32);
  int rows = 256 / cvb;
  int gy = ceil_div(cv, cvb);
  long long npix = (long long)dout.n * dout.h * dout.w;
  long long gx_need = (npix + rows - 1) / rows;
  int gx_cap = std::max(1, (kNumSMs * 4) / gy);
  dim3 grid((unsigned)std::min<long long>(gx_need, gx_cap), gy, 1);
  size_t smem = (size_t)rows * cvb * 36 * sizeof(float);
  dw_bwd_weight_kernel<T><<<grid, 256, smem, st>>>(make_view<const T>(in), make_view<const T>(dout), s, d, G, cvb, rows);
  return launch_status("dc_dw_bwd_weight");
}

static int check_dw(const char* what, const dc_view& big, const dc_view& small, int s, int d) {
  DC_REQUIRE(view_ok(big) && view_ok(small) && view_vec4(big) && view_vec4(small), "%s: views must be channel-contiguous, C %% 4 == 0", what);
  DC_REQUIRE(big.dtype == small.dtype && big.c == small.c && big.n == small.n, "%s: dtype/channel/batch mismatch", what);
  DC_REQUIRE((s == 1 || s == 2) && d >= 1, "%s: stride must be 1 or 2, dilation >= 1", what);
  // fixed_padding pads d on every side: H_out = floor((H + 2d - (2d+1)) / s) + 1 = floor((H-1)/s) + 1
  DC_REQUIRE(small.h == (big.h - 1) / s + 1 && small.w == (big.w - 1) / s + 1, "%s: output size mismatch (%dx%d -> %dx%d, stride %d)",
             what, big.h, big.w, small.h, small.w, s);
  return 0;
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_dw_fwd(dc_view in, const void* w9c, int stride, int dil, dc_view out, void* stream) {
  if (int r = check_dw("dc_dw_fwd", in, out, stride, dil)) return r;
  DC_REQUIRE(w9c != nullptr, "dc_dw_fwd: null weights");
  cudaStream_t st = as_stream(stream);
  return in.dtype == DC_F32 ? dw_fwd_t<float>(in, w9c, stride, dil, out, st) : dw_fwd_t<__nv_bfloat16>(in, w9c, stride, dil, out, st);
}

int dc_dw_bwd_data(dc_view dout, const void* w9c, int stride, int dil, dc_view din, int accumulate, void* stream) {
  if (int r = check_dw("dc_dw_bwd_data", din, dout, stride, dil)) return r;
  DC_REQUIRE(w9c != nullptr, "dc_dw_bwd_data: null weights");
  cudaStream_t st = as_stream(stream);
  return din.dtype == DC_F32 ? dw_bwd_data_t<float>(dout, w9c, stride, dil, din, accumulate, st)
                             : dw_bwd_data_t<__nv_bfloat16>(dout, w9c, stride, dil, din, accumulate, st);
}

int dc_dw_bwd_weight(dc_view in, dc_view dout, int stride, int dil, float* G9c, void* stream) {
  if (int r = check_dw("dc_dw_bwd_weight", in, dout, stride, dil)) return r;
  DC_REQUIRE(G9c != nullptr, "dc_dw_bwd_weight: null gradient");
  cudaStream_t st = as_stream(stream);
  return in.dtype == DC_F32 ? dw_bwd_weight_t<float>(in, dout, stride, dil, G9c, st)
                            : dw_bwd_weight_t<__nv_bfloat16>(in, dout, stride, dil, G9c, st);
}

}  // extern "C"
