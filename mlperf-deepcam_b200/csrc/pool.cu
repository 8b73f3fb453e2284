// Image-pooling branch of the ASPP head: AdaptiveAvgPool2d((1,1)) (DX:425) and the bilinear
// align_corners=True resize of a 1x1 map to HxW (DX:450), which is a broadcast; general bilinear align_corners=True
// resize and its gradient for the InterpolationUpsampler decoder variant (DX:327-331).  The pooled vectors
// stay in fp32 (BatchNorm over two values is a sign function, keep the branch in fp32).
#include "common.cuh"
#include <algorithm>

namespace dc {

// grid = (pixel chunks, channel chunks, n); block = 256 = cvb channel-vector lanes x rows pixel lanes
template <typename T>
__global__ void __launch_bounds__(256) reduce_hw_kernel(View<const T> x, float* __restrict__ out_nc, float scale, int cvb, int rows) {
  extern __shared__ float red[];  // [rows][cvb*4]
  const int tx = threadIdx.x % cvb, ty = threadIdx.x / cvb;
  const int c4 = blockIdx.y * cvb + tx;
  const int n = blockIdx.z;
  const bool ok = (ty < rows) && (c4 * 4 < x.c);
  const int hw = x.h * x.w;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    for (int p = blockIdx.x * rows + ty; p < hw; p += gridDim.x * rows) {
      int h = p / x.w, w = p - h * x.w;
      float4 v = elem<T>::ld4(x.at(n, h, w) + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (ty < rows) {
    float* r = red + (ty * cvb + tx) * 4;
    r[0] = ok ? acc.x : 0.f; r[1] = ok ? acc.y : 0.f; r[2] = ok ? acc.z : 0.f; r[3] = ok ? acc.w : 0.f;
  }
  __syncthreads();
  for (int col = threadIdx.x; col < cvb * 4; col += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += red[r * cvb * 4 + col];
    int c = blockIdx.y * cvb * 4 + col;
    if (c < x.c) atomicAdd(out_nc + (size_t)n * x.c + c, s * scale);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) broadcast_hw_kernel(const float* __restrict__ src_nc, View<T> dst) {
  const int cv = dst.c >> 2;
  const long long total = (long long)dst.n * dst.h * dst.w * cv;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < total;
       item += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(item % cv);
    int pix = (int)(item / cv);
    int w = pix % dst.w;
    int t = pix / dst.w;
    int h = t % dst.h;
    int n = t / dst.h;
    float4 v = *reinterpret_cast<const float4*>(src_nc + (size_t)n * dst.c + c4 * 4);
    elem<T>::st4(dst.at(n, h, w) + c4 * 4, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gap_bwd_kernel(const float* __restrict__ dmean_nc, View<T> dx, float inv_hw, int accumulate) {
  const int cv = dx.c >> 2;
  const long long total = (long long)dx.n * dx.h * dx.w * cv;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < total;
       item += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(item % cv);
    int pix = (int)(item / cv);
    int w = pix % dx.w;
    int t = pix / dx.w;
    int h = t % dx.h;
    int n = t / dx.h;
    float4 v = *reinterpret_cast<const float4*>(dmean_nc + (size_t)n * dx.c + c4 * 4);
    v.x *= inv_hw; v.y *= inv_hw; v.z *= inv_hw; v.w *= inv_hw;
    T* p = dx.at(n, h, w) + c4 * 4;
    if (accumulate) {
      float4 o = elem<T>::ld4(p);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    elem<T>::st4(p, v);
  }
}

// ---- bilinear resize, align_corners=True (F.interpolate of InterpolationUpsampler, DX:327-331) --------------------------
// Source coordinate of output index o: scale * o with scale = (in - 1) / (out - 1) in fp32 (0 when out == 1), lower
// neighbour i0 = min(int(src), in - 1), upper i1 = min(i0 + 1, in - 1), weight of the upper one = src - i0.
struct LinCoord { int i0, i1; float l1; };
__device__ __forceinline__ LinCoord lin_coord(int o, float scale, int in_size) {
  LinCoord r;
  const float s = scale * (float)o;
  r.i0 = min((int)s, in_size - 1);
  r.i1 = min(r.i0 + 1, in_size - 1);
  r.l1 = fminf(fmaxf(s - (float)r.i0, 0.f), 1.f);
  return r;
}
static inline float lin_scale(int in_size, int out_size) { return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f; }

// one thread per (output pixel, 4 channels): 4 coalesced 8/16-byte reads, one write
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(View<const TI> in, View<TO> out, float sy, float sx) {
  const int cv = out.c >> 2;
  const long long total = (long long)out.n * out.h * out.w * cv;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < total;
       item += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(item % cv);
    const int pix = (int)(item / cv);
    const int ox = pix % out.w;
    const int t = pix / out.w;
    const int oy = t % out.h;
    const int n = t / out.h;
    const LinCoord y = lin_coord(oy, sy, in.h), x = lin_coord(ox, sx, in.w);
    const float4 v00 = elem<TI>::ld4(in.at(n, y.i0, x.i0) + c4 * 4), v01 = elem<TI>::ld4(in.at(n, y.i0, x.i1) + c4 * 4);
    const float4 v10 = elem<TI>::ld4(in.at(n, y.i1, x.i0) + c4 * 4), v11 = elem<TI>::ld4(in.at(n, y.i1, x.i1) + c4 * 4);
    const float ly0 = 1.f - y.l1, lx0 = 1.f - x.l1;
    float4 r;
    r.x = ly0 * (lx0 * v00.x + x.l1 * v01.x) + y.l1 * (lx0 * v10.x + x.l1 * v11.x);
    r.y = ly0 * (lx0 * v00.y + x.l1 * v01.y) + y.l1 * (lx0 * v10.y + x.l1 * v11.y);
    r.z = ly0 * (lx0 * v00.z + x.l1 * v01.z) + y.l1 * (lx0 * v10.z + x.l1 * v11.z);
    r.w = ly0 * (lx0 * v00.w + x.l1 * v01.w) + y.l1 * (lx0 * v10.w + x.l1 * v11.w);
    elem<TO>::st4(out.at(n, oy, ox) + c4 * 4, r);
  }
}

// Weight with which output index o reads input index i (both neighbours may coincide at the last row/column).
__device__ __forceinline__ float lin_weight(int o, int i, float scale, int in_size) {
  const LinCoord c = lin_coord(o, scale, in_size);
  return (c.i0 == i ? 1.f - c.l1 : 0.f) + (c.i1 == i ? c.l1 : 0.f);
}
// Conservative range of output indices that can read input index i.
__device__ __forceinline__ void lin_support(int i, float scale, int out_size, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out_size - 1; return; }
  const float inv = 1.f / scale;
  lo = max(0, (int)floorf((float)(i - 1) * inv) - 1);
  hi = min(out_size - 1, (int)ceilf((float)(i + 1) * inv) + 1);
}

// Gather form of the gradient (deterministic, no atomics): one thread per (input pixel, 4 channels) sums every output pixel
// whose stencil contains it.  For the x4 upsampling of the decoder that is at most 8 x 8 reads, all L2 hits.
template <typename T>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(View<const T> dout, View<T> din, float sy, float sx, int accumulate) {
  const int cv = din.c >> 2;
  const long long total = (long long)din.n * din.h * din.w * cv;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < total;
       item += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(item % cv);
    const int pix = (int)(item / cv);
    const int ix = pix % din.w;
    const int t = pix / din.w;
    const int iy = t % din.h;
    const int n = t / din.h;
    int ylo, yhi, xlo, xhi;
    lin_support(iy, sy, dout.h, ylo, yhi);
    lin_support(ix, sx, dout.w, xlo, xhi);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oy = ylo; oy <= yhi; ++oy) {
      const float wy = lin_weight(oy, iy, sy, din.h);
      if (wy == 0.f) continue;
      for (int ox = xlo; ox <= xhi; ++ox) {
        const float wgt = wy * lin_weight(ox, ix, sx, din.w);
        if (wgt == 0.f) continue;
        const float4 g = elem<T>::ld4(dout.at(n, oy, ox) + c4 * 4);
        acc.x += wgt * g.x; acc.y += wgt * g.y; acc.z += wgt * g.z; acc.w += wgt * g.w;
      }
    }
    T* p = din.at(n, iy, ix) + c4 * 4;
    if (accumulate) {
      const float4 o = elem<T>::ld4(p);
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    elem<T>::st4(p, acc);
  }
}

template <typename T>
static int reduce_hw_t(const dc_view& x, float* out, float scale, cudaStream_t st, const char* what) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)x.n * x.c, st);
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  int cv = x.c / 4, cvb = std::min(cv, 32), rows = 256 / cvb, gy = ceil_div(cv, cvb);
  int hw = x.h * x.w;
  int gx = std::max(1, std::min(ceil_div(hw, rows * 4), std::max(1, (kNumSMs * 4) / (gy * x.n))));
  if (deterministic()) gx = 1;      // dc_set_deterministic: one block per (channel chunk, image) - a single add onto the zeroed output
  dim3 grid(gx, gy, x.n);
  reduce_hw_kernel<T><<<grid, 256, (size_t)rows * cvb * 4 * sizeof(float), st>>>(make_view<const T>(x), out, scale, cvb, rows);
  return launch_status(what);
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_gap_fwd(dc_view x, float* mean_nc, void* stream) {
  DC_REQUIRE(view_ok(x) && view_vec4(x) && mean_nc, "dc_gap_fwd: bad arguments");
  float scale = 1.0f / (float)(x.h * x.w);
  cudaStream_t st = as_stream(stream);
  return x.dtype == DC_F32 ? reduce_hw_t<float>(x, mean_nc, scale, st, "dc_gap_fwd")
                           : reduce_hw_t<__nv_bfloat16>(x, mean_nc, scale, st, "dc_gap_fwd");
}

int dc_reduce_hw(dc_view x, float* sum_nc, void* stream) {
  DC_REQUIRE(view_ok(x) && view_vec4(x) && sum_nc, "dc_reduce_hw: bad arguments");
  cudaStream_t st = as_stream(stream);
  return x.dtype == DC_F32 ? reduce_hw_t<float>(x, sum_nc, 1.0f, st, "dc_reduce_hw")
                           : reduce_hw_t<__nv_bfloat16>(x, sum_nc, 1.0f, st, "dc_reduce_hw");
}

int dc_broadcast_hw(const float* src_nc, dc_view dst, void* stream) {
  DC_REQUIRE(view_ok(dst) && view_vec4(dst) && src_nc, "dc_broadcast_hw: bad arguments");
  DC_REQUIRE((reinterpret_cast<uintptr_t>(src_nc) % 16) == 0, "dc_broadcast_hw: src must be 16-byte aligned");
  long long total = (long long)dst.n * dst.h * dst.w * (dst.c / 4);
  int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  cudaStream_t st = as_stream(stream);
  if (dst.dtype == DC_F32) broadcast_hw_kernel<float><<<blocks, 256, 0, st>>>(src_nc, make_view<float>(dst));
  else broadcast_hw_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(src_nc, make_view<__nv_bfloat16>(dst));
  return launch_status("dc_broadcast_hw");
}

int dc_gap_bwd(const float* dmean_nc, dc_view dx, int accumulate, void* stream) {
  DC_REQUIRE(view_ok(dx) && view_vec4(dx) && dmean_nc, "dc_gap_bwd: bad arguments");
  DC_REQUIRE((reinterpret_cast<uintptr_t>(dmean_nc) % 16) == 0, "dc_gap_bwd: dmean must be 16-byte aligned");
  long long total = (long long)dx.n * dx.h * dx.w * (dx.c / 4);
  int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  float inv_hw = 1.0f / (float)(dx.h * dx.w);
  cudaStream_t st = as_stream(stream);
  if (dx.dtype == DC_F32) gap_bwd_kernel<float><<<blocks, 256, 0, st>>>(dmean_nc, make_view<float>(dx), inv_hw, accumulate);
  else gap_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(dmean_nc, make_view<__nv_bfloat16>(dx), inv_hw, accumulate);
  return launch_status("dc_gap_bwd");
}

int dc_bilinear_fwd(dc_view in, dc_view out, void* stream) {
  DC_REQUIRE(view_ok(in) && view_vec4(in) && view_ok(out) && view_vec4(out), "dc_bilinear_fwd: views must be channel-contiguous with C %% 4 == 0");
  DC_REQUIRE(in.n == out.n && in.c == out.c, "dc_bilinear_fwd: batch and channel counts must match");
  DC_REQUIRE(in.dtype == out.dtype || (in.dtype == DC_BF16 && out.dtype == DC_F32), "dc_bilinear_fwd: output must have the input's type or be fp32");
  const long long total = (long long)out.n * out.h * out.w * (out.c / 4);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  const float sy = lin_scale(in.h, out.h), sx = lin_scale(in.w, out.w);
  cudaStream_t st = as_stream(stream);
  if (in.dtype == DC_F32) bilinear_fwd_kernel<float, float><<<blocks, 256, 0, st>>>(make_view<const float>(in), make_view<float>(out), sy, sx);
  else if (out.dtype == DC_F32)
    bilinear_fwd_kernel<__nv_bfloat16, float><<<blocks, 256, 0, st>>>(make_view<const __nv_bfloat16>(in), make_view<float>(out), sy, sx);
  else
    bilinear_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>(make_view<const __nv_bfloat16>(in), make_view<__nv_bfloat16>(out), sy, sx);
  return launch_status("dc_bilinear_fwd");
}

int dc_bilinear_bwd(dc_view dout, dc_view din, int accumulate, void* stream) {
  DC_REQUIRE(view_ok(dout) && view_vec4(dout) && view_ok(din) && view_vec4(din), "dc_bilinear_bwd: views must be channel-contiguous with C %% 4 == 0");
  DC_REQUIRE(din.n == dout.n && din.c == dout.c && din.dtype == dout.dtype, "dc_bilinear_bwd: batch, channels and type must match");
  const long long total = (long long)din.n * din.h * din.w * (din.c / 4);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  const float sy = lin_scale(din.h, dout.h), sx = lin_scale(din.w, dout.w);
  cudaStream_t st = as_stream(stream);
  if (din.dtype == DC_F32) bilinear_bwd_kernel<float><<<blocks, 256, 0, st>>>(make_view<const float>(dout), make_view<float>(din), sy, sx, accumulate);
  else bilinear_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(make_view<const __nv_bfloat16>(dout), make_view<__nv_bfloat16>(din), sy, sx, accumulate);
  return launch_status("dc_bilinear_bwd");
}

}  // extern "C"
