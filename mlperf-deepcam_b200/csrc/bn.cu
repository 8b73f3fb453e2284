// BatchNorm2d training/eval kernels fused with ReLU and the residual add.
// Reference semantics: torch.nn.BatchNorm2d as `normalizer` (DX:70,129,283,348,399; eps 1e-5, momentum 0.1),
// nn.ReLU(inplace=True) (DX:79,147) and the in-place residual add `x += skip` (DX:120).
//
// Thread mapping shared by all kernels here ("channel lanes x pixel lanes"):
//   a block owns `cvb` consecutive 4-channel vectors (<= 32, i.e. <= 128 channels) and `rows = 256/cvb`
//   pixel lanes; each thread keeps its channel vector fixed and strides over pixels, so a warp reads
//   up to 256 contiguous bytes per pixel and per-channel coefficients are computed once per thread.
// HBM-bound: algorithmic bytes = each tensor read or written exactly once.
#include "common.cuh"
#include <algorithm>

namespace dc {

struct ChanGrid {
  int cv, cvb, rows;
  dim3 grid;
};
static inline ChanGrid chan_grid(int C, long long npix) {
  ChanGrid g;
  g.cv = C / 4;
  g.cvb = std::min(g.cv, 32);
  g.rows = 256 / g.cvb;
  int gy = ceil_div(g.cv, g.cvb);
  long long gx_need = (npix + g.rows - 1) / g.rows;
  int gx_cap = std::max(1, (kNumSMs * 8) / gy);
  g.grid = dim3((unsigned)std::min<long long>(gx_need, gx_cap), gy, 1);
  return g;
}

__device__ __forceinline__ void decode_pix(int p, int H, int W, int& n, int& h, int& w) {
  w = p % W;
  int t = p / W;
  h = t % H;
  n = t / H;
}

struct Coef { float mean, invstd, scale, shift; };
__device__ __forceinline__ Coef bn_coef(const dc_bn_params& p, int C, int c) {
  Coef k;
  if (p.flags & DC_BN_IDENTITY) { k.mean = 0.f; k.invstd = 1.f; k.scale = 1.f; k.shift = 0.f; return k; }
  double m, var;
  if (p.flags & DC_BN_TRAIN) {
    m = p.sums[c] / p.count;
    var = p.sums[C + c] / p.count - m * m;
    if (var < 0.0) var = 0.0;
  } else {
    m = (double)p.running_mean[c];
    var = (double)p.running_var[c];
  }
  double inv = 1.0 / sqrt(var + (double)p.eps);
  k.mean = (float)m;
  k.invstd = (float)inv;
  k.scale = p.gamma[c] * k.invstd;
  k.shift = p.beta[c] - k.mean * k.scale;
  return k;
}

// Reduce `nacc` per-thread fp32 partials (per channel of the thread's vector) across the block's pixel
// lanes and add them to double accumulators in global memory: dst[a*C + c].
template <int NACC>
__device__ __forceinline__ void block_reduce_to_global(float (&acc)[NACC][4], double* dst, int C, int c4, bool lane_ok,
                                                       int cvb, int rows, int tx, int ty) {
  extern __shared__ double red[];   // [rows][cvb*4*NACC]
  const int per_row = cvb * 4 * NACC;
  if (ty < rows) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[ty * per_row + (a * cvb + tx) * 4 + j] = lane_ok ? (double)acc[a][j] : 0.0;
  }
  __syncthreads();
  // threads 0..per_row-1 each own one (a, tx, j) column
  for (int col = threadIdx.x; col < per_row; col += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += red[r * per_row + col];
    int a = col / (cvb * 4);
    int rem = col - a * cvb * 4;
    int ltx = rem >> 2, j = rem & 3;
    int c = (blockIdx.y * cvb + ltx) * 4 + j;
    if (c < C) atomicAdd(dst + (size_t)a * C + c, s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(View<const T> y, double* sums, int cvb, int rows) {
  const int tx = threadIdx.x % cvb, ty = threadIdx.x / cvb;
  const int c4 = blockIdx.y * cvb + tx;
  const bool ok = (ty < rows) && (c4 * 4 < y.c);
  const int npix = y.n * y.h * y.w;
  float acc[2][4] = {};
  if (ok) {
    for (int p = blockIdx.x * rows + ty; p < npix; p += gridDim.x * rows) {
      int n, h, w;
      decode_pix(p, y.h, y.w, n, h, w);
      float4 v = elem<T>::ld4(y.at(n, h, w) + c4 * 4);
      acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
      acc[1][0] += v.x * v.x; acc[1][1] += v.y * v.y; acc[1][2] += v.z * v.z; acc[1][3] += v.w * v.w;
    }
  }
  block_reduce_to_global<2>(acc, sums, y.c, c4, ok, cvb, rows, tx, ty);
}

template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(dc_bn_params p, View<const T> y, View<const T> res, View<T> out,
                                                       int cvb, int rows) {
  const int tx = threadIdx.x % cvb, ty = threadIdx.x / cvb;
  const int c4 = blockIdx.y * cvb + tx;
  const int C = y.c;
  if (ty >= rows || c4 * 4 >= C) return;
  const int npix = y.n * y.h * y.w;
  Coef k[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) k[j] = bn_coef(p, C, c4 * 4 + j);
  if ((p.flags & DC_BN_TRAIN) && !(p.flags & DC_BN_IDENTITY) && blockIdx.x == 0 && ty == 0 && p.running_mean != nullptr) {
    // running statistics: torch uses the unbiased variance for the running estimate
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = c4 * 4 + j;
      double m = p.sums[c] / p.count;
      double var = p.sums[C + c] / p.count - m * m;
      if (var < 0.0) var = 0.0;
      double unb = var * (p.count / (p.count - 1.0));
      p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * (float)m;
      p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unb;
    }
  }
  const bool relu = (p.flags & DC_BN_RELU) != 0;
  const bool has_res = res.p != nullptr;
  for (int pix = blockIdx.x * rows + ty; pix < npix; pix += gridDim.x * rows) {
    int n, h, w;
    decode_pix(pix, y.h, y.w, n, h, w);
    float4 v = elem<T>::ld4(y.at(n, h, w) + c4 * 4);
    float4 o;
    o.x = fmaf(v.x, k[0].scale, k[0].shift);
    o.y = fmaf(v.y, k[1].scale, k[1].shift);
    o.z = fmaf(v.z, k[2].scale, k[2].shift);
    o.w = fmaf(v.w, k[3].scale, k[3].shift);
    if (has_res) {
      float4 r = elem<T>::ld4(res.at(n, h, w) + c4 * 4);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    elem<T>::st4(out.at(n, h, w) + c4 * 4, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(dc_bn_params p, View<const T> dout, View<const T> out,
                                                            View<const T> y, double* rsums, int cvb, int rows) {
  const int tx = threadIdx.x % cvb, ty = threadIdx.x / cvb;
  const int c4 = blockIdx.y * cvb + tx;
  const bool ok = (ty < rows) && (c4 * 4 < y.c);
  const int npix = y.n * y.h * y.w;
  const bool relu = (p.flags & DC_BN_RELU) != 0;
  float acc[2][4] = {};
  if (ok) {
    for (int pix = blockIdx.x * rows + ty; pix < npix; pix += gridDim.x * rows) {
      int n, h, w;
      decode_pix(pix, y.h, y.w, n, h, w);
      float4 g = elem<T>::ld4(dout.at(n, h, w) + c4 * 4);
      if (relu) {
        float4 o = elem<T>::ld4(out.at(n, h, w) + c4 * 4);
        g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
        g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
      }
      float4 v = elem<T>::ld4(y.at(n, h, w) + c4 * 4);
      acc[0][0] += g.x; acc[0][1] += g.y; acc[0][2] += g.z; acc[0][3] += g.w;
      acc[1][0] += g.x * v.x; acc[1][1] += g.y * v.y; acc[1][2] += g.z * v.z; acc[1][3] += g.w * v.w;
    }
  }
  block_reduce_to_global<2>(acc, rsums, y.c, c4, ok, cvb, rows, tx, ty);
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(dc_bn_params p, View<const T> dout, View<const T> out,
                                                           View<const T> y, const double* rsums, View<T> dy, View<T> dres,
                                                           float* dgamma, float* dbeta, int cvb, int rows) {
  const int tx = threadIdx.x % cvb, ty = threadIdx.x / cvb;
  const int c4 = blockIdx.y * cvb + tx;
  const int C = dout.c;
  if (ty >= rows || c4 * 4 >= C) return;
  const int npix = dout.n * dout.h * dout.w;
  const bool relu = (p.flags & DC_BN_RELU) != 0;
  const bool ident = (p.flags & DC_BN_IDENTITY) != 0;
  const bool has_res = dres.p != nullptr;
  const bool res_write = (p.flags & DC_BN_RES_WRITE) != 0;
  const bool has_dy = dy.p != nullptr;
  Coef k[4];
  float mg[4], mgx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = c4 * 4 + j;
    k[j] = bn_coef(p, C, c);
    if (!ident) {
      double sg = rsums[c], sgy = rsums[C + c];
      double sgx = (double)k[j].invstd * (sgy - (double)k[j].mean * sg);   // sum g*xhat
      if (p.flags & DC_BN_TRAIN) {
        mg[j] = (float)(sg / p.count);
        mgx[j] = (float)(sgx / p.count);
      } else {            // eval-mode BN inside a training graph (freeze_bn, DX:467): statistics are constants
        mg[j] = 0.f; mgx[j] = 0.f;
      }
      if (blockIdx.x == 0 && ty == 0) {
        if (dgamma) dgamma[c] = (float)sgx;
        if (dbeta) dbeta[c] = (float)sg;
      }
    } else { mg[j] = 0.f; mgx[j] = 0.f; }
  }
  for (int pix = blockIdx.x * rows + ty; pix < npix; pix += gridDim.x * rows) {
    int n, h, w;
    decode_pix(pix, dout.h, dout.w, n, h, w);
    float4 g = elem<T>::ld4(dout.at(n, h, w) + c4 * 4);
    if (relu) {
      float4 o = elem<T>::ld4(out.at(n, h, w) + c4 * 4);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
      g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    if (has_res) {
      T* rp = dres.at(n, h, w) + c4 * 4;
      float4 r = g;
      if (!res_write) {
        float4 old = elem<T>::ld4(rp);
        r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
      }
      elem<T>::st4(rp, r);
    }
    if (has_dy) {
      float4 d;
      if (ident) {
        d = g;
      } else {
        float4 v = elem<T>::ld4(y.at(n, h, w) + c4 * 4);
        d.x = k[0].scale * (g.x - mg[0] - (v.x - k[0].mean) * k[0].invstd * mgx[0]);
        d.y = k[1].scale * (g.y - mg[1] - (v.y - k[1].mean) * k[1].invstd * mgx[1]);
        d.z = k[2].scale * (g.z - mg[2] - (v.z - k[2].mean) * k[2].invstd * mgx[2]);
        d.w = k[3].scale * (g.w - mg[3] - (v.w - k[3].mean) * k[3].invstd * mgx[3]);
      }
      elem<T>::st4(dy.at(n, h, w) + c4 * 4, d);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(View<const T> x, double* sums, int cvb, int rows) {
  const int tx = threadIdx.x % cvb, ty = threadIdx.x / cvb;
  const int c4 = blockIdx.y * cvb + tx;
  const bool ok = (ty < rows) && (c4 * 4 < x.c);
  const int npix = x.n * x.h * x.w;
  float acc[1][4] = {};
  if (ok) {
    for (int p = blockIdx.x * rows + ty; p < npix; p += gridDim.x * rows) {
      int n, h, w;
      decode_pix(p, x.h, x.w, n, h, w);
      float4 v = elem<T>::ld4(x.at(n, h, w) + c4 * 4);
      acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
    }
  }
  block_reduce_to_global<1>(acc, sums, x.c, c4, ok, cvb, rows, tx, ty);
}

__global__ void double_to_float_kernel(const double* s, float* d, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (float)s[i];
}

static inline size_t red_smem(const ChanGrid& g, int nacc) { return (size_t)g.rows * g.cvb * 4 * nacc * sizeof(double); }

template <typename T>
static int bn_stats_t(const dc_view& y, double* sums, cudaStream_t st) {
  ChanGrid g = chan_grid(y.c, (long long)y.n * y.h * y.w);
  bn_stats_kernel<T><<<g.grid, 256, red_smem(g, 2), st>>>(make_view<const T>(y), sums, g.cvb, g.rows);
  return launch_status("dc_bn_stats");
}
template <typename T>
static int bn_apply_t(const dc_bn_params& p, const dc_view& y, const dc_view& res, const dc_view& out, cudaStream_t st) {
  ChanGrid g = chan_grid(y.c, (long long)y.n * y.h * y.w);
  View<const T> r = make_view<const T>(res);
  bn_apply_kernel<T><<<g.grid, 256, 0, st>>>(p, make_view<const T>(y), r, make_view<T>(out), g.cvb, g.rows);
  return launch_status("dc_bn_apply");
}
template <typename T>
static int bn_bwd_reduce_t(const dc_bn_params& p, const dc_view& dout, const dc_view& out, const dc_view& y, double* rs, cudaStream_t st) {
  ChanGrid g = chan_grid(y.c, (long long)y.n * y.h * y.w);
  bn_bwd_reduce_kernel<T><<<g.grid, 256, red_smem(g, 2), st>>>(p, make_view<const T>(dout), make_view<const T>(out),
                                                               make_view<const T>(y), rs, g.cvb, g.rows);
  return launch_status("dc_bn_bwd_reduce");
}
template <typename T>
static int bn_bwd_apply_t(const dc_bn_params& p, const dc_view& dout, const dc_view& out, const dc_view& y, const double* rs,
                          const dc_view& dy, const dc_view& dres, float* dgamma, float* dbeta, cudaStream_t st) {
  ChanGrid g = chan_grid(dout.c, (long long)dout.n * dout.h * dout.w);
  bn_bwd_apply_kernel<T><<<g.grid, 256, 0, st>>>(p, make_view<const T>(dout), make_view<const T>(out), make_view<const T>(y), rs,
                                                 make_view<T>(dy), make_view<T>(dres), dgamma, dbeta, g.cvb, g.rows);
  return launch_status("dc_bn_bwd_apply");
}

}  // namespace dc

using namespace dc;

static bool opt_view_ok(const dc_view& v, const dc_view& like) {
  if (v.ptr == nullptr) return true;
  return view_ok(v) && view_vec4(v) && same_shape(v, like) && v.dtype == like.dtype;
}

extern "C" {

int dc_bn_stats(dc_view y, double* sums, void* stream) {
  DC_REQUIRE(view_ok(y) && view_vec4(y), "dc_bn_stats: view must be channel-contiguous with C %% 4 == 0");
  DC_REQUIRE(sums != nullptr, "dc_bn_stats: null sums");
  cudaStream_t st = as_stream(stream);
  return y.dtype == DC_F32 ? bn_stats_t<float>(y, sums, st) : bn_stats_t<__nv_bfloat16>(y, sums, st);
}

int dc_bn_apply(const dc_bn_params* p, dc_view y, dc_view residual, dc_view out, void* stream) {
  DC_REQUIRE(p != nullptr, "dc_bn_apply: null params");
  DC_REQUIRE(view_ok(y) && view_vec4(y), "dc_bn_apply: bad y view");
  DC_REQUIRE(view_ok(out) && opt_view_ok(out, y), "dc_bn_apply: bad out view");
  DC_REQUIRE(opt_view_ok(residual, y), "dc_bn_apply: bad residual view");
  if (!(p->flags & DC_BN_IDENTITY)) {
    DC_REQUIRE(p->gamma && p->beta, "dc_bn_apply: gamma/beta required");
    if (p->flags & DC_BN_TRAIN) {
      DC_REQUIRE(p->sums != nullptr, "dc_bn_apply: batch statistics required in train mode");
      DC_REQUIRE(p->count > 1.0, "dc_bn_apply: Expected more than 1 value per channel when training (count=%g)", p->count);
    } else {
      DC_REQUIRE(p->running_mean && p->running_var, "dc_bn_apply: running statistics required in eval mode");
    }
  }
  cudaStream_t st = as_stream(stream);
  return y.dtype == DC_F32 ? bn_apply_t<float>(*p, y, residual, out, st) : bn_apply_t<__nv_bfloat16>(*p, y, residual, out, st);
}

int dc_bn_bwd_reduce(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, double* rsums, void* stream) {
  DC_REQUIRE(p != nullptr && rsums != nullptr, "dc_bn_bwd_reduce: null argument");
  DC_REQUIRE(view_ok(y) && view_vec4(y) && opt_view_ok(dout, y) && view_ok(dout), "dc_bn_bwd_reduce: bad views");
  if (p->flags & DC_BN_RELU) DC_REQUIRE(view_ok(out) && opt_view_ok(out, y), "dc_bn_bwd_reduce: out view required for ReLU mask");
  cudaStream_t st = as_stream(stream);
  return y.dtype == DC_F32 ? bn_bwd_reduce_t<float>(*p, dout, out, y, rsums, st)
                           : bn_bwd_reduce_t<__nv_bfloat16>(*p, dout, out, y, rsums, st);
}

int dc_bn_bwd_apply(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, const double* rsums, dc_view dy,
                    dc_view dres, float* dgamma, float* dbeta, void* stream) {
  DC_REQUIRE(p != nullptr, "dc_bn_bwd_apply: null params");
  DC_REQUIRE(view_ok(dout) && view_vec4(dout), "dc_bn_bwd_apply: bad dout view");
  DC_REQUIRE(opt_view_ok(dy, dout) && opt_view_ok(dres, dout), "dc_bn_bwd_apply: bad dy/dres view");
  if (p->flags & DC_BN_RELU) DC_REQUIRE(view_ok(out) && opt_view_ok(out, dout), "dc_bn_bwd_apply: out view required for ReLU mask");
  if (!(p->flags & DC_BN_IDENTITY)) {
    DC_REQUIRE(rsums != nullptr && view_ok(y) && opt_view_ok(y, dout), "dc_bn_bwd_apply: y and rsums required");
    DC_REQUIRE(p->gamma && p->beta, "dc_bn_bwd_apply: gamma/beta required");
  }
  cudaStream_t st = as_stream(stream);
  return dout.dtype == DC_F32 ? bn_bwd_apply_t<float>(*p, dout, out, y, rsums, dy, dres, dgamma, dbeta, st)
                              : bn_bwd_apply_t<__nv_bfloat16>(*p, dout, out, y, rsums, dy, dres, dgamma, dbeta, st);
}

/* channel sum with a caller-provided double workspace of C elements */
int dc_channel_sum(dc_view x, double* ws_c, float* out_c, void* stream) {
  DC_REQUIRE(view_ok(x) && view_vec4(x) && out_c != nullptr && ws_c != nullptr, "dc_channel_sum: bad arguments");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(ws_c, 0, sizeof(double) * x.c, st);
  if (e != cudaSuccess) return dc::fail((int)e, "dc_channel_sum: %s", cudaGetErrorString(e));
  ChanGrid g = chan_grid(x.c, (long long)x.n * x.h * x.w);
  if (x.dtype == DC_F32)
    channel_sum_kernel<float><<<g.grid, 256, red_smem(g, 1), st>>>(make_view<const float>(x), ws_c, g.cvb, g.rows);
  else
    channel_sum_kernel<__nv_bfloat16><<<g.grid, 256, red_smem(g, 1), st>>>(make_view<const __nv_bfloat16>(x), ws_c, g.cvb, g.rows);
  double_to_float_kernel<<<ceil_div(x.c, 256), 256, 0, st>>>(ws_c, out_c, x.c);
  return launch_status("dc_channel_sum");
}

}  // extern "C"
