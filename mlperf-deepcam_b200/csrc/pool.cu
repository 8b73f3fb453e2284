// Image-pooling branch of the ASPP head: AdaptiveAvgPool2d((1,1)) (DX:425) and the bilinear
// align_corners=True resize of a 1x1 map to HxW (DX:450), which is a broadcast.  The pooled vectors
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

template <typename T>
static int reduce_hw_t(const dc_view& x, float* out, float scale, cudaStream_t st, const char* what) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)x.n * x.c, st);
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  int cv = x.c / 4, cvb = std::min(cv, 32), rows = 256 / cvb, gy = ceil_div(cv, cvb);
  int hw = x.h * x.w;
  int gx = std::max(1, std::min(ceil_div(hw, rows * 4), std::max(1, (kNumSMs * 4) / (gy * x.n))));
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

}  // extern "C"
