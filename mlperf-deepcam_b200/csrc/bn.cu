// BatchNorm2d training/eval kernels fused with ReLU and the residual add.
// Reference semantics: torch.nn.BatchNorm2d as `normalizer` (DX:70,129,283,348,399; eps 1e-5, momentum 0.1),
// nn.ReLU(inplace=True) (DX:79,147) and the in-place residual add `x += skip` (DX:120).
//
// Thread mapping shared by all kernels ("channel lanes x pixel lanes"):
//   a thread owns one vector of V channels (16 bytes: V = 8 bf16 or 4 fp32) and strides over pixels; the 32 lanes
//   of a warp cover 32 consecutive channel vectors of one pixel (512 contiguous bytes), or, when a pixel has fewer
//   than 32 vectors, several consecutive pixels.  Per-channel coefficients are loaded once per thread.
// Statistics: per-thread fp32 partials -> block reduction -> fp64 atomics into the per-layer workspace; the LAST
// block to finish (atomic ticket) turns the sums into fp32 per-channel coefficients, so the apply kernels carry
// no double-precision prologue and no extra "finalize" launch is needed.
//   workspace (dc_bn_ws_bytes(C) bytes, zeroed by the caller):  double sums[2][C] | float coef[4][C] | uint32 ticket[16]
//     forward : coef = scale, shift, mean, invstd            (out = y*scale + shift)
//     backward: coef = A, B, D                               (dy  = A*g + B*y + D,  g = dout masked by ReLU)
// HBM-bound: algorithmic bytes = each tensor read or written exactly once.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace dc {

// ---- 16-byte channel vectors ---------------------------------------------------------------------------
template <typename T> struct vec16;
template <> struct vec16<float> {
  static constexpr int V = 4;
  __device__ static __forceinline__ uint4 ldraw(const float* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static __forceinline__ void unpack(const uint4& t, float (&f)[4]) {
    f[0] = __uint_as_float(t.x); f[1] = __uint_as_float(t.y); f[2] = __uint_as_float(t.z); f[3] = __uint_as_float(t.w);
  }
  __device__ static __forceinline__ void ld(const float* p, float (&f)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  }
  __device__ static __forceinline__ void st(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct vec16<__nv_bfloat16> {
  static constexpr int V = 8;
  __device__ static __forceinline__ uint4 ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static __forceinline__ void unpack(const uint4& t, float (&f)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w[j] << 16); f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
  }
  __device__ static __forceinline__ void ld(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w[j] << 16); f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
  }
  __device__ static __forceinline__ void st(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// ---- lane mapping ----------------------------------------------------------------------------------------
struct LaneMap {
  int cv;        // channel vectors per pixel
  int cvp;       // lanes per pixel inside a warp (power of two <= 32)
  int ppw;       // pixels per warp = 32 / cvp
  int ppb;       // pixels per block = 8 warps * ppw
  int gy;        // channel blocks
};
static inline LaneMap lane_map(int C, int V, int warps = 8) {
  LaneMap m;
  m.cv = C / V;
  m.cvp = 1;
  while (m.cvp < 32 && m.cvp < m.cv) m.cvp <<= 1;
  m.ppw = 32 / m.cvp;
  m.ppb = warps * m.ppw;
  m.gy = ceil_div(m.cv, m.cvp);
  return m;
}

// strided pixel addressing; `lin` = the view is pixel-linear (offset = pixel * sw), true for dense NHWC tensors and
// channel slices of concat buffers
template <typename T>
struct PixView {
  T* p;
  int h, w;
  long long sn, sh, sw;
  int lin;
  __device__ __forceinline__ T* at(long long pix) const {
    if (lin) return p + pix * sw;
    int x = (int)(pix % w);
    long long t = pix / w;
    int y = (int)(t % h);
    long long n = t / h;
    return p + n * sn + y * sh + x * sw;
  }
};
template <typename T>
static inline PixView<T> pix_view(const dc_view& v) {
  PixView<T> r;
  r.p = reinterpret_cast<T*>(v.ptr);
  r.h = v.h; r.w = v.w; r.sn = v.sn; r.sh = v.sh; r.sw = v.sw;
  r.lin = (v.ptr != nullptr && v.sh == (long long)v.w * v.sw && v.sn == (long long)v.h * v.sh) ? 1 : 0;
  return r;
}

constexpr int kBnThreads = 256;
constexpr int kUnroll = 4;


// returns true in every thread of the block that took the last ticket
__device__ __forceinline__ bool last_block(unsigned* ticket) {
  __shared__ unsigned s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    s_last = (atomicAdd(ticket, 1u) == total - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0u;
}

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(dc_bn_params p, PixView<const T> y, int C, long long npix, LaneMap m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_sync();
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int cvi = blockIdx.y * m.cvp + cvl;
  const bool ok = cvi < m.cv;
  float acc[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const long long stride = (long long)gridDim.x * m.ppb;
  long long pix = (long long)blockIdx.x * m.ppb + warp * m.ppw + psub;
  if (ok) {
    for (; pix + (kUnroll - 1) * stride < npix; pix += kUnroll * stride) {
      uint4 raw[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) raw[u] = vec16<T>::ldraw(y.at(pix + u * stride) + cvi * V);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        float v[V];
        vec16<T>::unpack(raw[u], v);
#pragma unroll
        for (int j = 0; j < V; ++j) { acc[0][j] += v[j]; acc[1][j] = fmaf(v[j], v[j], acc[1][j]); }
      }
    }
    for (; pix < npix; pix += stride) {
      float v[V];
      vec16<T>::ld(y.at(pix) + cvi * V, v);
#pragma unroll
      for (int j = 0; j < V; ++j) { acc[0][j] += v[j]; acc[1][j] = fmaf(v[j], v[j], acc[1][j]); }
    }
  }
  BnWs ws = bn_ws(const_cast<double*>(p.sums), C);
  reduce_to_ws<2, V>(acc, ws.sums, C, m.cvp, blockIdx.y * m.cvp, min(m.cvp, m.cv - blockIdx.y * m.cvp));
  if (p.flags & DC_BN_SUMS_READY) return;          // sums-only use (bn_accumulate_sums): the consumer finalizes
  if (!last_block(ws.ticket)) return;
  // ---- finalize: batch statistics -> fp32 coefficients, running statistics (torch uses the unbiased variance there).
  // mean and variance in double (cancellation), 1/sqrt in fp32 with one Newton step (the apply path is fp32 anyway).
  const double inv_count = 1.0 / p.count;
  const double unbias = p.count / (p.count - 1.0);
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    const double s = __ldcg(ws.sums + c), q = __ldcg(ws.sums + C + c);
    const double mean = s * inv_count;
    double var = q * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = inv_sqrt_f32((float)(var + (double)p.eps));
    const float scale = p.gamma[c] * inv;
    ws.coef[c] = scale;
    ws.coef[C + c] = p.beta[c] - (float)mean * scale;
    ws.coef[2 * C + c] = (float)mean;
    ws.coef[3 * C + c] = inv;
    if (p.running_mean != nullptr) {
      p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * (float)mean;
      p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)(var * unbias);
    }
  }
}

// per-thread forward coefficients of V channels
template <int V>
__device__ __forceinline__ void load_fwd_coef(const dc_bn_params& p, int C, int c0, float (&scale)[V], float (&shift)[V],
                                              float (&mean)[V], float (&invstd)[V]) {
  if (p.flags & DC_BN_IDENTITY) {
#pragma unroll
    for (int j = 0; j < V; ++j) { scale[j] = 1.f; shift[j] = 0.f; mean[j] = 0.f; invstd[j] = 1.f; }
  } else if (p.flags & DC_BN_TRAIN) {
    const BnWs ws = bn_ws(const_cast<double*>(p.sums), C);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      scale[j] = ws.coef[c0 + j]; shift[j] = ws.coef[C + c0 + j];
      mean[j] = ws.coef[2 * C + c0 + j]; invstd[j] = ws.coef[3 * C + c0 + j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float inv = inv_sqrt_f32(p.running_var[c0 + j] + p.eps);
      mean[j] = p.running_mean[c0 + j];
      invstd[j] = inv;
      scale[j] = p.gamma[c0 + j] * inv;
      shift[j] = p.beta[c0 + j] - mean[j] * scale[j];
    }
  }
}

template <typename T, int V, int kUnroll, bool HAS_RES>
__global__ void __launch_bounds__(kBnThreads, 3) bn_apply_kernel(dc_bn_params p, PixView<const T> y, PixView<const T> res, PixView<T> out,
                                                              int C, long long npix, LaneMap m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_sync();
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int cvi = blockIdx.y * m.cvp + cvl;
  const int c0 = cvi * V;
  float scale[V], shift[V];
  if ((p.flags & DC_BN_TRAIN) && (p.flags & DC_BN_SUMS_READY) && !(p.flags & DC_BN_IDENTITY)) {
    // The producer of y (GEMM epilogue) left the raw batch sums in the workspace.  The block finalizes its <= 256 channels
    // cooperatively (thread t -> one channel, same arithmetic as the last block of bn_stats_kernel) into shared memory;
    // block column 0 also publishes the coefficients for the backward pass and updates the running statistics.
    __shared__ float s_coef[2][32 * V];
    const int nch = m.cvp * V;
    if ((int)threadIdx.x < nch) {
      const int c = blockIdx.y * nch + threadIdx.x;
      float sc = 0.f, sh = 0.f;
      if (c < C) {
        const BnWs ws = bn_ws(const_cast<double*>(p.sums), C);
        const double inv_count = 1.0 / p.count;
        const double mu = ws.sums[c] * inv_count;
        double var = ws.sums[C + c] * inv_count - mu * mu;
        if (var < 0.0) var = 0.0;
        const float inv = inv_sqrt_f32((float)(var + (double)p.eps));
        sc = p.gamma[c] * inv;
        sh = p.beta[c] - (float)mu * sc;
        if (blockIdx.x == 0) {
          ws.coef[c] = sc;
          ws.coef[C + c] = sh;
          ws.coef[2 * C + c] = (float)mu;
          ws.coef[3 * C + c] = inv;
          if (p.running_mean != nullptr) {
            const double unbias = p.count / (p.count - 1.0);
            p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * (float)mu;
            p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)(var * unbias);
          }
        }
      }
      s_coef[0][threadIdx.x] = sc;
      s_coef[1][threadIdx.x] = sh;
    }
    __syncthreads();
    if (cvi >= m.cv) return;
#pragma unroll
    for (int j = 0; j < V; ++j) { scale[j] = s_coef[0][cvl * V + j]; shift[j] = s_coef[1][cvl * V + j]; }
  } else {
    if (cvi >= m.cv) return;
    float mean[V], invstd[V];
    load_fwd_coef<V>(p, C, c0, scale, shift, mean, invstd);
  }
  const bool relu = (p.flags & DC_BN_RELU) != 0;
  constexpr bool has_res = HAS_RES;
  const long long stride = (long long)gridDim.x * m.ppb;
  long long pix = (long long)blockIdx.x * m.ppb + warp * m.ppw + psub;
  for (; pix < npix; pix += kUnroll * stride) {
    uint4 vraw[kUnroll], rraw[HAS_RES ? kUnroll : 1];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (pix + u * stride < npix) {
        vraw[u] = vec16<T>::ldraw(y.at(pix + u * stride) + c0);
        if (has_res) rraw[HAS_RES ? u : 0] = vec16<T>::ldraw(res.at(pix + u * stride) + c0);
      }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (pix + u * stride < npix) {
        float v[V], r[V], o[V];
        vec16<T>::unpack(vraw[u], v);
        if (has_res) vec16<T>::unpack(rraw[HAS_RES ? u : 0], r);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          o[j] = fmaf(v[j], scale[j], shift[j]);
          if (has_res) o[j] += r[j];
          if (relu) o[j] = fmaxf(o[j], 0.f);
        }
        vec16<T>::st(out.at(pix + u * stride) + c0, o);
      }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(dc_bn_params p, PixView<const T> dout, PixView<const T> out,
                                                                   PixView<const T> y, void* rws_raw, float* dgamma, float* dbeta,
                                                                   int C, long long npix, LaneMap m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_sync();
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int cvi = blockIdx.y * m.cvp + cvl;
  const bool ok = cvi < m.cv;
  const int c0 = cvi * V;
  const bool relu = (p.flags & DC_BN_RELU) != 0;
  // DC_BN_MASK_FROM_Y: out = relu(fma(y, scale, shift)) with the forward coefficients still in the forward workspace, so
  // the ReLU decision is recomputed from y (identical) and `out` is not read: one input stream less
  const bool mask_y = relu && (p.flags & DC_BN_MASK_FROM_Y) != 0;
  const bool load_out = relu && !mask_y;
  float fsc[V], fsh[V];
  if (mask_y && ok) {
    const BnWs fws = bn_ws(const_cast<double*>(p.sums), C);
#pragma unroll
    for (int j = 0; j < V; ++j) { fsc[j] = fws.coef[c0 + j]; fsh[j] = fws.coef[C + c0 + j]; }
  }
  float acc[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const long long stride = (long long)gridDim.x * m.ppb;
  long long pix = (long long)blockIdx.x * m.ppb + warp * m.ppw + psub;
  if (ok) {
    for (; pix < npix; pix += kUnroll * stride) {
      uint4 graw[kUnroll], oraw[kUnroll], vraw[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (pix + u * stride < npix) {
          graw[u] = vec16<T>::ldraw(dout.at(pix + u * stride) + c0);
          if (load_out) oraw[u] = vec16<T>::ldraw(out.at(pix + u * stride) + c0);
          vraw[u] = vec16<T>::ldraw(y.at(pix + u * stride) + c0);
        }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (pix + u * stride < npix) {
          float g[V], o[V], v[V];
          vec16<T>::unpack(graw[u], g);
          vec16<T>::unpack(vraw[u], v);
          if (load_out) vec16<T>::unpack(oraw[u], o);
          else if (mask_y) {
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] = fmaf(v[j], fsc[j], fsh[j]);
          }
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float gg = g[j];
            if (relu) gg = o[j] > 0.f ? gg : 0.f;
            acc[0][j] += gg;
            acc[1][j] = fmaf(gg, v[j], acc[1][j]);
          }
        }
    }
  }
  BnWs rws = bn_ws(rws_raw, C);
  reduce_to_ws<2, V>(acc, rws.sums, C, m.cvp, blockIdx.y * m.cvp, min(m.cvp, m.cv - blockIdx.y * m.cvp));
  if (p.flags & DC_BN_SUMS_READY) return;          // sums only: dc_bn_bwd_apply_finalize derives the coefficients itself
  if (!last_block(rws.ticket)) return;
  // ---- finalize: dgamma, dbeta and the per-channel coefficients of dy = A*g + B*y + D
  const bool train = (p.flags & DC_BN_TRAIN) != 0;
  const double inv_count = 1.0 / p.count;
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    double mean, inv;
    if (train) {                 // the forward workspace holds exactly the coefficients the forward pass applied
      const BnWs fws = bn_ws(const_cast<double*>(p.sums), C);
      mean = p.sums[c] * inv_count;
      inv = (double)fws.coef[3 * C + c];
    } else {
      mean = (double)p.running_mean[c];
      inv = (double)inv_sqrt_f32(p.running_var[c] + p.eps);
    }
    const double sg = __ldcg(rws.sums + c), sgy = __ldcg(rws.sums + C + c);
    const double sgx = inv * (sgy - mean * sg);                  // sum g * xhat
    if (dgamma) dgamma[c] = (float)sgx;
    if (dbeta) dbeta[c] = (float)sg;
    const double scale = (double)p.gamma[c] * inv;
    double A = scale, B = 0.0, D = 0.0;
    if (train) {          // eval-mode BN inside a training graph (freeze_bn, DX:467): statistics are constants
      const double mg = sg * inv_count, mgx = sgx * inv_count;
      B = -scale * inv * mgx;
      D = -scale * mg + scale * mean * inv * mgx;
    }
    rws.coef[c] = (float)A;
    rws.coef[C + c] = (float)B;
    rws.coef[2 * C + c] = (float)D;
  }
}

template <typename T, int V, int kUnroll, bool MASK_Y>
__global__ void __launch_bounds__(kBnThreads, 2) bn_bwd_apply_kernel(dc_bn_params p, PixView<const T> dout, PixView<const T> out,
                                                                  PixView<const T> y, const void* rws_raw, PixView<T> dy, PixView<T> dres,
                                                                  float* dgamma, float* dbeta, int C, long long npix, LaneMap m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_sync();
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int cvi = blockIdx.y * m.cvp + cvl;
  const int c0 = cvi * V;
  const bool relu = (p.flags & DC_BN_RELU) != 0;
  const bool ident = (p.flags & DC_BN_IDENTITY) != 0;
  const bool reduced = !ident && (p.flags & DC_BN_SUMS_READY) != 0;
  __shared__ float s_abd[3][32 * V];
  if (reduced) {
    // dc_bn_bwd_apply_reduced (train mode): the producer of dout (dc_dw_bwd_data_bnred) left sum(g) and sum(g*y) in the
    // backward workspace; the block finalizes its channels cooperatively (thread t -> one channel, the arithmetic of the
    // last block of bn_bwd_reduce_kernel); block column 0 also writes the affine gradients
    const int nch = m.cvp * V;
    if ((int)threadIdx.x < nch) {
      const int c = blockIdx.y * nch + threadIdx.x;
      float a = 0.f, b = 0.f, d = 0.f;
      if (c < C) {
        const BnWs rws = bn_ws(const_cast<void*>(rws_raw), C);
        const BnWs fws = bn_ws(const_cast<double*>(p.sums), C);
        const double inv_count = 1.0 / p.count;
        const double mean = p.sums[c] * inv_count;
        const double inv = (double)fws.coef[3 * C + c];
        const double sg = rws.sums[c], sgy = rws.sums[C + c];
        const double sgx = inv * (sgy - mean * sg);
        const double scale = (double)p.gamma[c] * inv;
        const double mg = sg * inv_count, mgx = sgx * inv_count;
        a = (float)scale;
        b = (float)(-scale * inv * mgx);
        d = (float)(-scale * mg + scale * mean * inv * mgx);
        if (blockIdx.x == 0) {
          if (dgamma) dgamma[c] = (float)sgx;
          if (dbeta) dbeta[c] = (float)sg;
        }
      }
      s_abd[0][threadIdx.x] = a; s_abd[1][threadIdx.x] = b; s_abd[2][threadIdx.x] = d;
    }
    __syncthreads();
  }
  if (cvi >= m.cv) return;
  const bool has_res = dres.p != nullptr;
  const bool res_write = (p.flags & DC_BN_RES_WRITE) != 0;
  const bool has_dy = dy.p != nullptr;
  const bool need_y = has_dy && !ident;
  constexpr bool mask_y = MASK_Y;                  // host: relu && DC_BN_MASK_FROM_Y (no residual); see bn_bwd_reduce_kernel
  const bool load_out = relu && !mask_y;
  float fsc[MASK_Y ? V : 1], fsh[MASK_Y ? V : 1];
  if (mask_y && relu) {
    const BnWs fws = bn_ws(const_cast<double*>(p.sums), C);
#pragma unroll
    for (int j = 0; j < V; ++j) { fsc[j] = fws.coef[c0 + j]; fsh[j] = fws.coef[C + c0 + j]; }
  }
  float A[V], B[V], D[V];
  if (reduced) {
#pragma unroll
    for (int j = 0; j < V; ++j) { A[j] = s_abd[0][cvl * V + j]; B[j] = s_abd[1][cvl * V + j]; D[j] = s_abd[2][cvl * V + j]; }
  } else if (need_y) {
    const BnWs rws = bn_ws(const_cast<void*>(rws_raw), C);
#pragma unroll
    for (int j = 0; j < V; ++j) { A[j] = rws.coef[c0 + j]; B[j] = rws.coef[C + c0 + j]; D[j] = rws.coef[2 * C + c0 + j]; }
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) { A[j] = 1.f; B[j] = 0.f; D[j] = 0.f; }
  }
  const long long stride = (long long)gridDim.x * m.ppb;
  long long pix = (long long)blockIdx.x * m.ppb + warp * m.ppw + psub;
  for (; pix < npix; pix += kUnroll * stride) {
    constexpr int KO = MASK_Y ? 1 : kUnroll;      // `out` and the residual gradient are not read in mask-from-y mode
    uint4 graw[kUnroll], oraw[KO], vraw[kUnroll], rraw[KO];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (pix + u * stride < npix) {
        graw[u] = vec16<T>::ldraw(dout.at(pix + u * stride) + c0);
        if (load_out) oraw[MASK_Y ? 0 : u] = vec16<T>::ldraw(out.at(pix + u * stride) + c0);
        if (need_y || mask_y) vraw[u] = vec16<T>::ldraw(y.at(pix + u * stride) + c0);
        if (!MASK_Y && has_res && !res_write) rraw[MASK_Y ? 0 : u] = vec16<T>::ldraw(dres.at(pix + u * stride) + c0);
      }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (pix + u * stride < npix) {
        float gg[V], o[V], v[V], r[V];
        vec16<T>::unpack(graw[u], gg);
        if (need_y || mask_y) vec16<T>::unpack(vraw[u], v);
        if (relu) {
          if (load_out) vec16<T>::unpack(oraw[MASK_Y ? 0 : u], o);
          else {
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] = fmaf(v[j], fsc[MASK_Y ? j : 0], fsh[MASK_Y ? j : 0]);
          }
#pragma unroll
          for (int j = 0; j < V; ++j) gg[j] = o[j] > 0.f ? gg[j] : 0.f;
        }
        if (!MASK_Y && has_res) {
          float rr[V];
          if (!res_write) vec16<T>::unpack(rraw[MASK_Y ? 0 : u], r);
#pragma unroll
          for (int j = 0; j < V; ++j) rr[j] = res_write ? gg[j] : gg[j] + r[j];
          vec16<T>::st(dres.at(pix + u * stride) + c0, rr);
        }
        if (has_dy) {
          float d[V];
#pragma unroll
          for (int j = 0; j < V; ++j) d[j] = need_y ? fmaf(A[j], gg[j], fmaf(B[j], v[j], D[j])) : gg[j];
          vec16<T>::st(dy.at(pix + u * stride) + c0, d);
        }
      }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads) channel_sum_kernel(PixView<const T> x, double* sums, int C, long long npix, LaneMap m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_sync();
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int cvi = blockIdx.y * m.cvp + cvl;
  const bool ok = cvi < m.cv;
  float acc[1][V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[0][j] = 0.f;
  const long long stride = (long long)gridDim.x * m.ppb;
  if (ok) {
    for (long long pix = (long long)blockIdx.x * m.ppb + warp * m.ppw + psub; pix < npix; pix += stride) {
      float v[V];
      vec16<T>::ld(x.at(pix) + cvi * V, v);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[0][j] += v[j];
    }
  }
  reduce_to_ws<1, V>(acc, sums, C, m.cvp, blockIdx.y * m.cvp, min(m.cvp, m.cv - blockIdx.y * m.cvp));
}

__global__ void double_to_float_kernel(const double* s, float* d, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (float)s[i];
}

// ---- host side ---------------------------------------------------------------------------------------------
// grid.x: enough blocks that every thread handles about `items` pixels, capped at 16 blocks per SM (grid-stride beyond)
constexpr int kApplyUnroll = 8;       // forward apply: one read stream -> eight 16-byte loads in flight per thread
constexpr int kBwdApplyUnroll = 4;    // backward apply: three read streams
// blocks-per-SM caps of the grid-stride kernels: resident-sized grids amortize the per-block prologue (coefficient loads,
// finalize) and were measured ~1.4x faster on the 100+ MB tensors than one block per 4-8 pixels.  Tunables for sweeps:
// DEEPCAM_B200_BN_CAP_{APPLY,BWD_APPLY,REDUCE}.
static int bn_cap(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : 0;
  return v > 0 ? v : dflt;
}
static inline dim3 bn_grid(const LaneMap& m, long long npix, int items, int blocks_per_sm_cap = 16) {
  long long gx = ceil_div64(npix, (long long)m.ppb * items);
  long long cap = std::max<long long>(1, (long long)kNumSMs * blocks_per_sm_cap / m.gy);
  return dim3((unsigned)std::max<long long>(1, std::min(gx, cap)), (unsigned)m.gy, 1);
}
template <int NACC, int V> static inline size_t red_smem() { return (size_t)8 * 32 * NACC * V * sizeof(float); }

// 16-byte channel vectors need: unit channel stride, C % V == 0, 16-byte aligned base and pixel strides
template <typename T>
static bool vec_ok(const dc_view& v) {
  const int V = vec16<T>::V;
  return v.sc == 1 && (v.c % V == 0) && (v.sn % V == 0) && (v.sh % V == 0) && (v.sw % V == 0) &&
         ((reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0);
}

template <typename T>
static int bn_stats_t(const dc_bn_params& p, const dc_view& y, cudaStream_t st) {
  constexpr int V = vec16<T>::V;
  const long long npix = (long long)y.n * y.h * y.w;
  LaneMap m = lane_map(y.c, V);
  dim3 grid = bn_grid(m, npix, 4 * kUnroll);
  launch_k(bn_stats_kernel<T, V>, grid, dim3(kBnThreads), red_smem<2, V>(), st, p, pix_view<const T>(y), y.c, npix, m);
  return launch_status("dc_bn_stats");
}
template <typename T>
static int bn_apply_t(const dc_bn_params& p, const dc_view& y, const dc_view& res, const dc_view& out, cudaStream_t st) {
  constexpr int V = vec16<T>::V;
  const long long npix = (long long)y.n * y.h * y.w;
  LaneMap m = lane_map(y.c, V);
  // element-wise: two pixels per thread and as many blocks as that needs (small register footprint, 3 blocks per SM)
  // with DC_BN_SUMS_READY every block first finalizes its channels: cap the grid at 3 resident blocks per SM (grid-stride
  // loop) so that this prologue is paid once per resident block instead of once per 8 pixels
  static const int cap = bn_cap("DEEPCAM_B200_BN_CAP_APPLY", 3);
  if (res.ptr != nullptr) {
    dim3 grid = bn_grid(m, npix, kApplyUnroll / 2, cap);
    launch_k(bn_apply_kernel<T, V, kApplyUnroll / 2, true>, grid, dim3(kBnThreads), (size_t)0, st, p, pix_view<const T>(y), pix_view<const T>(res), pix_view<T>(out), y.c, npix, m);
  } else {
    dim3 grid = bn_grid(m, npix, kApplyUnroll, cap);
    launch_k(bn_apply_kernel<T, V, kApplyUnroll, false>, grid, dim3(kBnThreads), (size_t)0, st, p, pix_view<const T>(y), pix_view<const T>(res), pix_view<T>(out), y.c, npix, m);
  }
  return launch_status("dc_bn_apply");
}
template <typename T>
static int bn_bwd_reduce_t(const dc_bn_params& p, const dc_view& dout, const dc_view& out, const dc_view& y, void* rws,
                           float* dgamma, float* dbeta, cudaStream_t st) {
  constexpr int V = vec16<T>::V;
  const long long npix = (long long)y.n * y.h * y.w;
  LaneMap m = lane_map(y.c, V);
  static const int cap = bn_cap("DEEPCAM_B200_BN_CAP_REDUCE", 2);
  static const int cap_small = bn_cap("DEEPCAM_B200_BN_CAP_REDUCE_SMALL", 2), items_small = bn_cap("DEEPCAM_B200_BN_ITEMS_REDUCE_SMALL", 8);
  // L2-resident tensors (<= 24 MB) are latency-bound: fewer pixels per thread, more resident blocks (tunable)
  const bool small = (long long)npix * y.c * (long long)sizeof(T) <= (24ll << 20);
  dim3 grid = small ? bn_grid(m, npix, items_small, cap_small) : bn_grid(m, npix, 4 * kUnroll, cap);
  launch_k(bn_bwd_reduce_kernel<T, V>, grid, dim3(kBnThreads), red_smem<2, V>(), st, p, pix_view<const T>(dout), pix_view<const T>(out),
                                                                         pix_view<const T>(y), rws, dgamma, dbeta, y.c, npix, m);
  return launch_status("dc_bn_bwd_reduce");
}
template <typename T>
static int bn_bwd_apply_t(const dc_bn_params& p, const dc_view& dout, const dc_view& out, const dc_view& y, const void* rws,
                          const dc_view& dy, const dc_view& dres, cudaStream_t st, float* dgamma = nullptr, float* dbeta = nullptr) {
  constexpr int V = vec16<T>::V;
  const long long npix = (long long)dout.n * dout.h * dout.w;
  LaneMap m = lane_map(dout.c, V);
  const bool reduced = (p.flags & DC_BN_SUMS_READY) != 0;
  static const int cap = bn_cap("DEEPCAM_B200_BN_CAP_BWD_APPLY", 2);
  dim3 grid = bn_grid(m, npix, kBwdApplyUnroll, cap);
  // lean instantiation (no `out`, no residual-gradient registers): mask-from-y mode, and reduced mode without a residual
  if (((p.flags & DC_BN_RELU) && (p.flags & DC_BN_MASK_FROM_Y)) || (reduced && dres.ptr == nullptr))
    launch_k(bn_bwd_apply_kernel<T, V, kBwdApplyUnroll, true>, grid, dim3(kBnThreads), (size_t)0, st, p, pix_view<const T>(dout), pix_view<const T>(out), pix_view<const T>(y), rws,
                                                         pix_view<T>(dy), pix_view<T>(dres), dgamma, dbeta, dout.c, npix, m);
  else
    launch_k(bn_bwd_apply_kernel<T, V, kBwdApplyUnroll, false>, grid, dim3(kBnThreads), (size_t)0, st, p, pix_view<const T>(dout), pix_view<const T>(out), pix_view<const T>(y), rws,
                                                         pix_view<T>(dy), pix_view<T>(dres), dgamma, dbeta, dout.c, npix, m);
  return launch_status("dc_bn_bwd_apply");
}


// ---- one-pass kernels for tensors that fit on chip -----------------------------------------------------------------
// When a BatchNorm input is small enough to be held in shared memory across the whole GPU (the 50 middle-flow layers:
// 2 x 48 x 72 x 728 bf16 = 10 MB over 147 SMs), statistics and normalisation run in ONE launch: every block loads its
// slice once, keeps it in shared memory, publishes its partial sums, waits at an inter-block barrier (all blocks are
// co-resident: grid <= number of SMs, one block per SM) and then applies the coefficients to the data it holds.  Compared
// with the stats + apply pair this saves one launch, the serial finalize tail and one full read of the tensor.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void grid_barrier(unsigned* ticket, unsigned expected) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(ticket, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ticket) : "memory");
      if (seen < expected) __nanosleep(64);
    } while (seen < expected);
  }
  __syncthreads();
}

constexpr int kOnePassWarps = 8;
constexpr int kOnePassThreads = kOnePassWarps * 32;
struct OnePassMap {
  LaneMap m;
  int gx;        // pixel blocks (grid.x); grid.y = m.gy channel blocks; gx * gy <= SM count
  int K;         // vectors held per thread
};

template <typename T, int V>
__global__ void __launch_bounds__(kOnePassThreads, 1) bn_fwd_onepass_kernel(dc_bn_params p, PixView<const T> y, PixView<const T> res,
                                                                       PixView<T> out, int C, long long npix, OnePassMap om) {
  extern __shared__ float red[];                                    // [0, 16 KB): block reduction; then K*256 held vectors
  uint4* hold = reinterpret_cast<uint4*>(reinterpret_cast<char*>(red) + kOnePassWarps * 32 * 2 * V * sizeof(float));
  const LaneMap& m = om.m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_wait();
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int cvi = blockIdx.y * m.cvp + cvl;
  const bool ok = cvi < m.cv;
  const int c0 = cvi * V;
  float acc[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const long long stride = (long long)om.gx * m.ppb;
  const long long pix0 = (long long)blockIdx.x * m.ppb + warp * m.ppw + psub;
  if (ok) {
    // the block's whole slice goes global -> shared memory with 16-byte cp.async copies that are all in flight at once
    // (K per thread, no register staging): one L2/HBM round trip instead of K/6 dependent ones
    const uint32_t hold_s = (uint32_t)__cvta_generic_to_shared(hold);
    for (int k = 0; k < om.K; ++k) {
      const long long pix = pix0 + k * stride;
      if (pix < npix) cp_async16(hold_s + (uint32_t)(k * kOnePassThreads + threadIdx.x) * 16u, y.at(pix) + c0);
    }
    cp_async_commit_wait_all();
    for (int k = 0; k < om.K; ++k) {
      const long long pix = pix0 + k * stride;
      if (pix < npix) {
        float v[V];
        vec16<T>::unpack(hold[k * kOnePassThreads + threadIdx.x], v);
#pragma unroll
        for (int j = 0; j < V; ++j) { acc[0][j] += v[j]; acc[1][j] = fmaf(v[j], v[j], acc[1][j]); }
      }
    }
  }
  BnWs ws = bn_ws(const_cast<double*>(p.sums), C);
  reduce_to_ws<2, V, kOnePassWarps>(acc, ws.sums, C, m.cvp, blockIdx.y * m.cvp, min(m.cvp, m.cv - blockIdx.y * m.cvp));
  grid_barrier(ws.ticket + blockIdx.y, (unsigned)om.gx);
  pdl_trigger();          // only now: every block of this grid is resident (see common.cuh)
  if (!ok) return;
  // every thread derives the coefficients of its own channels (mean / variance in double, 1/sqrt in fp32)
  const double inv_count = 1.0 / p.count;
  float scale[V], shift[V];
  const bool writer = (blockIdx.x == 0 && warp == 0 && psub == 0);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int c = c0 + j;
    const double sm = __ldcg(ws.sums + c), sq = __ldcg(ws.sums + C + c);
    const double mean = sm * inv_count;
    double var = sq * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = inv_sqrt_f32((float)(var + (double)p.eps));
    scale[j] = p.gamma[c] * inv;
    shift[j] = p.beta[c] - (float)mean * scale[j];
    if (writer) {
      ws.coef[c] = scale[j];
      ws.coef[C + c] = shift[j];
      ws.coef[2 * C + c] = (float)mean;
      ws.coef[3 * C + c] = inv;
      if (p.running_mean != nullptr) {
        const double unbias = p.count / (p.count - 1.0);
        p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * (float)mean;
        p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)(var * unbias);
      }
    }
  }
  const bool relu = (p.flags & DC_BN_RELU) != 0;
  const bool has_res = res.p != nullptr;
  for (int k = 0; k < om.K; ++k) {
    const long long pix = pix0 + k * stride;
    if (pix >= npix) break;
    float v[V], r[V], o[V];
    vec16<T>::unpack(hold[k * kOnePassThreads + threadIdx.x], v);
    if (has_res) vec16<T>::unpack(vec16<T>::ldraw(res.at(pix) + c0), r);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      o[j] = fmaf(v[j], scale[j], shift[j]);
      if (has_res) o[j] += r[j];
      if (relu) o[j] = fmaxf(o[j], 0.f);
    }
    vec16<T>::st(out.at(pix) + c0, o);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kOnePassThreads, 1) bn_bwd_onepass_kernel(dc_bn_params p, PixView<const T> dout, PixView<const T> out,
                                                                       PixView<const T> y, void* rws_raw, PixView<T> dy, PixView<T> dres,
                                                                       float* dgamma, float* dbeta, int C, long long npix, OnePassMap om) {
  extern __shared__ float red[];
  uint4* hold = reinterpret_cast<uint4*>(reinterpret_cast<char*>(red) + kOnePassWarps * 32 * 2 * V * sizeof(float));   // [K][256][2]: g, y
  const LaneMap& m = om.m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_wait();
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int cvi = blockIdx.y * m.cvp + cvl;
  const bool ok = cvi < m.cv;
  const int c0 = cvi * V;
  const bool relu = (p.flags & DC_BN_RELU) != 0;
  float acc[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const long long stride = (long long)om.gx * m.ppb;
  const long long pix0 = (long long)blockIdx.x * m.ppb + warp * m.ppw + psub;
  const bool train = (p.flags & DC_BN_TRAIN) != 0;
  // ReLU mask = (out > 0).  Without a residual, out = relu(fma(y, scale, shift)) with the forward coefficients still in the
  // forward workspace, so the same decision is recomputed from y and `out` is not read at all (one input stream less).
  const bool mask_from_y = relu && train && (dres.p == nullptr || (p.flags & DC_BN_MASK_FROM_Y) != 0);
  if (ok) {
    // gradient and pre-BN activation of the block's slice: global -> shared memory, every 16-byte copy in flight at once
    const uint32_t hold_s = (uint32_t)__cvta_generic_to_shared(hold);
    for (int k = 0; k < om.K; ++k) {
      const long long pix = pix0 + k * stride;
      if (pix < npix) {
        const uint32_t d = hold_s + (uint32_t)(k * kOnePassThreads + threadIdx.x) * 32u;
        cp_async16(d, dout.at(pix) + c0);
        cp_async16(d + 16u, y.at(pix) + c0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    float fsc[V], fsh[V];
    if (mask_from_y) {
      const BnWs fws = bn_ws(const_cast<double*>(p.sums), C);
#pragma unroll
      for (int j = 0; j < V; ++j) { fsc[j] = fws.coef[c0 + j]; fsh[j] = fws.coef[C + c0 + j]; }
    }
    constexpr int U = 6;
    const bool need_out = relu && !mask_from_y;
    for (int k0 = 0; k0 < om.K; k0 += U) {
      uint4 oraws[U];
      if (need_out) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long pix = pix0 + (k0 + u) * stride;
          if (k0 + u < om.K && pix < npix) oraws[u] = vec16<T>::ldraw(out.at(pix) + c0);
        }
      }
      if (k0 == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u;
        const long long pix = pix0 + k * stride;
        if (k < om.K && pix < npix) {
          const uint4 graw = hold[(k * kOnePassThreads + threadIdx.x) * 2];
          const uint4 yraw = hold[(k * kOnePassThreads + threadIdx.x) * 2 + 1];
          float g[V], v[V];
          vec16<T>::unpack(graw, g);
          vec16<T>::unpack(yraw, v);
          if (relu) {
            float o[V];
            if (mask_from_y) {
#pragma unroll
              for (int j = 0; j < V; ++j) o[j] = fmaf(v[j], fsc[j], fsh[j]);
            } else {
              vec16<T>::unpack(oraws[u], o);
            }
#pragma unroll
            for (int j = 0; j < V; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < V; ++j) { acc[0][j] += g[j]; acc[1][j] = fmaf(g[j], v[j], acc[1][j]); }
          if (relu) {
            // write the masked gradient back (bf16: the mask only zeroes lanes, so re-packing loses nothing)
            uint4 gm;
            if (sizeof(T) == 4) gm = make_uint4(__float_as_uint(g[0]), __float_as_uint(g[1]), __float_as_uint(g[2]), __float_as_uint(g[3]));
            else {
              uint32_t w[4] = {graw.x, graw.y, graw.z, graw.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (g[(2 * q) % V] == 0.f) w[q] &= 0xffff0000u;
                if (g[(2 * q + 1) % V] == 0.f) w[q] &= 0x0000ffffu;
              }
              gm = make_uint4(w[0], w[1], w[2], w[3]);
            }
            hold[(k * kOnePassThreads + threadIdx.x) * 2] = gm;
          }
        }
      }
    }
  }
  BnWs rws = bn_ws(rws_raw, C);
  reduce_to_ws<2, V, kOnePassWarps>(acc, rws.sums, C, m.cvp, blockIdx.y * m.cvp, min(m.cvp, m.cv - blockIdx.y * m.cvp));
  grid_barrier(rws.ticket + blockIdx.y, (unsigned)om.gx);
  pdl_trigger();
  if (!ok) return;
  const double inv_count = 1.0 / p.count;
  const bool writer = (blockIdx.x == 0 && warp == 0 && psub == 0);
  float A[V], B[V], D[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int c = c0 + j;
    double mean, inv;
    if (train) {
      const BnWs fws = bn_ws(const_cast<double*>(p.sums), C);
      mean = p.sums[c] * inv_count;
      inv = (double)fws.coef[3 * C + c];
    } else {
      mean = (double)p.running_mean[c];
      inv = (double)inv_sqrt_f32(p.running_var[c] + p.eps);
    }
    const double sg = __ldcg(rws.sums + c), sgy = __ldcg(rws.sums + C + c);
    const double sgx = inv * (sgy - mean * sg);
    const double scale = (double)p.gamma[c] * inv;
    double a = scale, b = 0.0, d = 0.0;
    if (train) {
      const double mg = sg * inv_count, mgx = sgx * inv_count;
      b = -scale * inv * mgx;
      d = -scale * mg + scale * mean * inv * mgx;
    }
    A[j] = (float)a; B[j] = (float)b; D[j] = (float)d;
    if (writer) {
      if (dgamma) dgamma[c] = (float)sgx;
      if (dbeta) dbeta[c] = (float)sg;
      rws.coef[c] = A[j]; rws.coef[C + c] = B[j]; rws.coef[2 * C + c] = D[j];
    }
  }
  const bool has_res = dres.p != nullptr;
  const bool res_write = (p.flags & DC_BN_RES_WRITE) != 0;
  const bool has_dy = dy.p != nullptr;
  for (int k = 0; k < om.K; ++k) {
    const long long pix = pix0 + k * stride;
    if (pix >= npix) break;
    float g[V], v[V];
    vec16<T>::unpack(hold[(k * kOnePassThreads + threadIdx.x) * 2], g);
    vec16<T>::unpack(hold[(k * kOnePassThreads + threadIdx.x) * 2 + 1], v);
    if (has_res) {
      float rr[V];
      if (!res_write) {
        float r[V];
        vec16<T>::unpack(vec16<T>::ldraw(dres.at(pix) + c0), r);
#pragma unroll
        for (int j = 0; j < V; ++j) rr[j] = g[j] + r[j];
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) rr[j] = g[j];
      }
      vec16<T>::st(dres.at(pix) + c0, rr);
    }
    if (has_dy) {
      float d[V];
#pragma unroll
      for (int j = 0; j < V; ++j) d[j] = fmaf(A[j], g[j], fmaf(B[j], v[j], D[j]));
      vec16<T>::st(dy.at(pix) + c0, d);
    }
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    else n = kNumSMs;
  }
  return n;
}
// per_vec = 16-byte vectors held per (thread, step): 1 forward, 2 backward
static bool onepass_map(int C, int V, long long npix, int per_vec, OnePassMap& om) {
  om.m = lane_map(C, V, kOnePassWarps);
  if (om.m.gy > 16) return false;
  om.gx = sm_count() / om.m.gy;
  if (om.gx < 1) return false;
  const long long per_round = (long long)om.gx * om.m.ppb;
  const long long K = ceil_div64(npix, per_round);
  const long long smem = (long long)kOnePassWarps * 32 * 2 * V * 4 + K * kOnePassThreads * 16 * per_vec;
  if (K < 1 || K > 64 || smem > 200 * 1024) return false;
  om.K = (int)K;
  return true;
}
template <typename KernelT>
static int set_smem_attr(KernelT kernel, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
  return 0;
}

template <typename T>
static int bn_fwd_onepass_t(const dc_bn_params& p, const dc_view& y, const dc_view& res, const dc_view& out, cudaStream_t st) {
  constexpr int V = vec16<T>::V;
  const long long npix = (long long)y.n * y.h * y.w;
  OnePassMap om;
  if (!onepass_map(y.c, V, npix, 1, om)) return fail(-2, "dc_bn_fwd_onepass: tensor does not fit on chip");
  if (int r = set_smem_attr(bn_fwd_onepass_kernel<T, V>, "dc_bn_fwd_onepass")) return r;
  const size_t smem = (size_t)kOnePassWarps * 32 * 2 * V * 4 + (size_t)om.K * kOnePassThreads * 16;
  launch_k(bn_fwd_onepass_kernel<T, V>, dim3(om.gx, om.m.gy, 1), dim3(kOnePassThreads), smem, st, p, pix_view<const T>(y), pix_view<const T>(res),
                                                                                 pix_view<T>(out), y.c, npix, om);
  return launch_status("dc_bn_fwd_onepass");
}
template <typename T>
static int bn_bwd_onepass_t(const dc_bn_params& p, const dc_view& dout, const dc_view& out, const dc_view& y, void* rws, const dc_view& dy,
                            const dc_view& dres, float* dgamma, float* dbeta, cudaStream_t st) {
  constexpr int V = vec16<T>::V;
  const long long npix = (long long)y.n * y.h * y.w;
  OnePassMap om;
  if (!onepass_map(y.c, V, npix, 2, om)) return fail(-2, "dc_bn_bwd_onepass: tensor does not fit on chip");
  if (int r = set_smem_attr(bn_bwd_onepass_kernel<T, V>, "dc_bn_bwd_onepass")) return r;
  const size_t smem = (size_t)kOnePassWarps * 32 * 2 * V * 4 + (size_t)om.K * kOnePassThreads * 32;
  launch_k(bn_bwd_onepass_kernel<T, V>, dim3(om.gx, om.m.gy, 1), dim3(kOnePassThreads), smem, st, p, pix_view<const T>(dout), pix_view<const T>(out),
                                                                                 pix_view<const T>(y), rws, pix_view<T>(dy), pix_view<T>(dres),
                                                                                 dgamma, dbeta, y.c, npix, om);
  return launch_status("dc_bn_bwd_onepass");
}

int bn_accumulate_sums(const dc_view& y, double* sums, cudaStream_t st) {
  dc_bn_params p = {};
  p.sums = sums;
  p.count = (double)y.n * y.h * y.w;
  p.flags = DC_BN_TRAIN | DC_BN_SUMS_READY;
  const bool ok = view_ok(y) && y.sc == 1 && (y.dtype == DC_F32 ? vec_ok<float>(y) : vec_ok<__nv_bfloat16>(y));
  if (!ok) return fail(-1, "bn_accumulate_sums: view must be channel-contiguous and 16-byte aligned");
  return y.dtype == DC_F32 ? bn_stats_t<float>(p, y, st) : bn_stats_t<__nv_bfloat16>(p, y, st);
}

}  // namespace dc

using namespace dc;

template <typename T>
static bool opt_ok(const dc_view& v, const dc_view& like) {
  if (v.ptr == nullptr) return true;
  return view_ok(v) && vec_ok<T>(v) && same_shape(v, like) && v.dtype == like.dtype;
}
#define DC_BN_DISPATCH(view_, expr_f32, expr_bf16) ((view_).dtype == DC_F32 ? (expr_f32) : (expr_bf16))

static bool views_ok(const dc_view& main, const dc_view* opts, int nopt) {
  if (!view_ok(main)) return false;
  if (main.dtype == DC_F32) {
    if (!vec_ok<float>(main)) return false;
    for (int i = 0; i < nopt; ++i) if (!opt_ok<float>(opts[i], main)) return false;
  } else {
    if (!vec_ok<__nv_bfloat16>(main)) return false;
    for (int i = 0; i < nopt; ++i) if (!opt_ok<__nv_bfloat16>(opts[i], main)) return false;
  }
  return true;
}

extern "C" {

size_t dc_bn_ws_bytes(int C) { return (size_t)32 * C + 64; }

int dc_bn_stats(const dc_bn_params* p, dc_view y, void* stream) {
  DC_REQUIRE(p != nullptr && p->sums != nullptr, "dc_bn_stats: null params / workspace");
  DC_REQUIRE(views_ok(y, nullptr, 0), "dc_bn_stats: view must be channel-contiguous, 16-byte aligned, C %% 8 == 0 (bf16) / C %% 4 == 0 (fp32)");
  DC_REQUIRE(p->gamma && p->beta, "dc_bn_stats: gamma/beta required");
  DC_REQUIRE(p->count > 1.0, "dc_bn_stats: Expected more than 1 value per channel when training (count=%g)", p->count);
  DC_REQUIRE((p->running_mean == nullptr) == (p->running_var == nullptr), "dc_bn_stats: running_mean and running_var go together");
  cudaStream_t st = as_stream(stream);
  return y.dtype == DC_F32 ? bn_stats_t<float>(*p, y, st) : bn_stats_t<__nv_bfloat16>(*p, y, st);
}

int dc_bn_apply(const dc_bn_params* p, dc_view y, dc_view residual, dc_view out, void* stream) {
  DC_REQUIRE(p != nullptr, "dc_bn_apply: null params");
  const dc_view opts[2] = {out, residual};
  DC_REQUIRE(view_ok(out) && views_ok(y, opts, 2), "dc_bn_apply: views must be channel-contiguous, 16-byte aligned and of one shape/dtype");
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

int dc_bn_bwd_reduce(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, void* rws, float* dgamma, float* dbeta,
                     void* stream) {
  DC_REQUIRE(p != nullptr && rws != nullptr, "dc_bn_bwd_reduce: null argument");
  const dc_view opts[2] = {dout, out};
  DC_REQUIRE(view_ok(dout) && views_ok(y, opts, 2), "dc_bn_bwd_reduce: bad views");
  if ((p->flags & DC_BN_RELU) && !(p->flags & DC_BN_MASK_FROM_Y)) DC_REQUIRE(view_ok(out), "dc_bn_bwd_reduce: out view required for ReLU mask");
  if (p->flags & DC_BN_MASK_FROM_Y) DC_REQUIRE((p->flags & DC_BN_TRAIN) && p->sums != nullptr, "dc_bn_bwd_reduce: DC_BN_MASK_FROM_Y needs the train-mode forward workspace");
  DC_REQUIRE(p->gamma != nullptr, "dc_bn_bwd_reduce: gamma required");
  if (p->flags & DC_BN_TRAIN) DC_REQUIRE(p->sums != nullptr, "dc_bn_bwd_reduce: forward workspace required in train mode");
  else DC_REQUIRE(p->running_mean && p->running_var, "dc_bn_bwd_reduce: running statistics required in eval mode");
  cudaStream_t st = as_stream(stream);
  return y.dtype == DC_F32 ? bn_bwd_reduce_t<float>(*p, dout, out, y, rws, dgamma, dbeta, st)
                           : bn_bwd_reduce_t<__nv_bfloat16>(*p, dout, out, y, rws, dgamma, dbeta, st);
}

int dc_bn_bwd_apply(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, const void* rws, dc_view dy, dc_view dres,
                    void* stream) {
  DC_REQUIRE(p != nullptr, "dc_bn_bwd_apply: null params");
  const dc_view opts[4] = {dy, dres, out, y};
  DC_REQUIRE(views_ok(dout, opts, 4), "dc_bn_bwd_apply: views must be channel-contiguous, 16-byte aligned and of one shape/dtype");
  if ((p->flags & DC_BN_RELU) && !(p->flags & DC_BN_MASK_FROM_Y)) DC_REQUIRE(view_ok(out), "dc_bn_bwd_apply: out view required for ReLU mask");
  if (p->flags & DC_BN_MASK_FROM_Y)
    DC_REQUIRE((p->flags & DC_BN_TRAIN) && p->sums != nullptr && view_ok(y) && dres.ptr == nullptr,
               "dc_bn_bwd_apply: DC_BN_MASK_FROM_Y needs the train-mode forward workspace, y, and no residual");
  if (!(p->flags & DC_BN_IDENTITY) && dy.ptr != nullptr)
    DC_REQUIRE(rws != nullptr && view_ok(y), "dc_bn_bwd_apply: y and the backward workspace are required");
  cudaStream_t st = as_stream(stream);
  return dout.dtype == DC_F32 ? bn_bwd_apply_t<float>(*p, dout, out, y, rws, dy, dres, st)
                              : bn_bwd_apply_t<__nv_bfloat16>(*p, dout, out, y, rws, dy, dres, st);
}

/* BatchNorm backward when the gradient has already been masked and reduced by its producer (dc_dw_bwd_data_bnred):
 * dy = A*g + B*y + D with the coefficients finalized inside the kernel from the workspace sums; dgamma/dbeta written;
 * dres (+)= g.  Train mode only; p->sums = forward workspace, rws = backward workspace holding sum(g), sum(g*y). */
int dc_bn_bwd_apply_reduced(const dc_bn_params* p, dc_view g, dc_view y, const void* rws, dc_view dy, dc_view dres, float* dgamma,
                            float* dbeta, void* stream) {
  DC_REQUIRE(p != nullptr && rws != nullptr, "dc_bn_bwd_apply_reduced: null argument");
  DC_REQUIRE((p->flags & DC_BN_TRAIN) && p->sums != nullptr && p->gamma != nullptr && !(p->flags & DC_BN_IDENTITY),
             "dc_bn_bwd_apply_reduced: train-mode BatchNorm with its forward workspace required");
  const dc_view opts[3] = {dy, dres, y};
  DC_REQUIRE(views_ok(g, opts, 3) && view_ok(y), "dc_bn_bwd_apply_reduced: views must be channel-contiguous, 16-byte aligned and of one shape/dtype");
  dc_bn_params q = *p;
  q.flags = (q.flags & ~(DC_BN_RELU | DC_BN_MASK_FROM_Y)) | DC_BN_SUMS_READY;       // g is already masked
  dc_view none = {};
  cudaStream_t st = as_stream(stream);
  return g.dtype == DC_F32 ? bn_bwd_apply_t<float>(q, g, none, y, rws, dy, dres, st, dgamma, dbeta)
                           : bn_bwd_apply_t<__nv_bfloat16>(q, g, none, y, rws, dy, dres, st, dgamma, dbeta);
}

/* Second half of the split BatchNorm backward for small tensors: dc_bn_bwd_reduce was called with DC_BN_SUMS_READY (sums
 * only, no last-block finalize); this call finalizes the coefficients per block (shared memory), applies the ReLU mask as the
 * flags say (out, or DC_BN_MASK_FROM_Y) and writes dy / dres / dgamma / dbeta.  Train mode. */
int dc_bn_bwd_apply_finalize(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, const void* rws, dc_view dy, dc_view dres,
                             float* dgamma, float* dbeta, void* stream) {
  DC_REQUIRE(p != nullptr && rws != nullptr, "dc_bn_bwd_apply_finalize: null argument");
  DC_REQUIRE((p->flags & DC_BN_TRAIN) && p->sums != nullptr && p->gamma != nullptr && !(p->flags & DC_BN_IDENTITY),
             "dc_bn_bwd_apply_finalize: train-mode BatchNorm with its forward workspace required");
  const dc_view opts[4] = {dy, dres, out, y};
  DC_REQUIRE(views_ok(dout, opts, 4) && view_ok(y), "dc_bn_bwd_apply_finalize: views must be channel-contiguous, 16-byte aligned and of one shape/dtype");
  if ((p->flags & DC_BN_RELU) && !(p->flags & DC_BN_MASK_FROM_Y)) DC_REQUIRE(view_ok(out), "dc_bn_bwd_apply_finalize: out view required for ReLU mask");
  if (p->flags & DC_BN_MASK_FROM_Y) DC_REQUIRE(dres.ptr == nullptr, "dc_bn_bwd_apply_finalize: DC_BN_MASK_FROM_Y excludes a residual");
  dc_bn_params q = *p;
  q.flags |= DC_BN_SUMS_READY;
  cudaStream_t st = as_stream(stream);
  return dout.dtype == DC_F32 ? bn_bwd_apply_t<float>(q, dout, out, y, rws, dy, dres, st, dgamma, dbeta)
                              : bn_bwd_apply_t<__nv_bfloat16>(q, dout, out, y, rws, dy, dres, st, dgamma, dbeta);
}

/* 1 when the one-pass kernels can handle a [npix, C] tensor of this dtype (everything held on chip), else 0 */
int dc_bn_onepass_ok(int C, long long npix, int dtype, int backward) {
  OnePassMap om;
  const int V = dtype == DC_F32 ? 4 : 8;
  if (C % V) return 0;
  return onepass_map(C, V, npix, backward ? 2 : 1, om) ? 1 : 0;
}

/* statistics + coefficients + running statistics + out = [relu](bn(y) [+ residual]) in one launch (train mode) */
int dc_bn_fwd_onepass(const dc_bn_params* p, dc_view y, dc_view residual, dc_view out, void* stream) {
  DC_REQUIRE(p != nullptr && p->sums != nullptr && (p->flags & DC_BN_TRAIN) && !(p->flags & DC_BN_IDENTITY),
             "dc_bn_fwd_onepass: train-mode parameters with a workspace are required");
  const dc_view opts[2] = {out, residual};
  DC_REQUIRE(view_ok(out) && views_ok(y, opts, 2), "dc_bn_fwd_onepass: views must be channel-contiguous, 16-byte aligned and of one shape/dtype");
  DC_REQUIRE(p->gamma && p->beta, "dc_bn_fwd_onepass: gamma/beta required");
  DC_REQUIRE(p->count > 1.0, "dc_bn_fwd_onepass: Expected more than 1 value per channel when training (count=%g)", p->count);
  cudaStream_t st = as_stream(stream);
  return y.dtype == DC_F32 ? bn_fwd_onepass_t<float>(*p, y, residual, out, st) : bn_fwd_onepass_t<__nv_bfloat16>(*p, y, residual, out, st);
}

/* dc_bn_bwd_reduce + dc_bn_bwd_apply in one launch */
int dc_bn_bwd_onepass(const dc_bn_params* p, dc_view dout, dc_view out, dc_view y, void* rws, dc_view dy, dc_view dres,
                      float* dgamma, float* dbeta, void* stream) {
  DC_REQUIRE(p != nullptr && rws != nullptr && !(p->flags & DC_BN_IDENTITY), "dc_bn_bwd_onepass: null argument");
  const dc_view opts[4] = {dout, out, dy, dres};
  DC_REQUIRE(view_ok(dout) && views_ok(y, opts, 4), "dc_bn_bwd_onepass: bad views");
  if ((p->flags & DC_BN_RELU) && !((p->flags & DC_BN_TRAIN) && dres.ptr == nullptr))
    DC_REQUIRE(view_ok(out), "dc_bn_bwd_onepass: out view required for ReLU mask");
  DC_REQUIRE(p->gamma != nullptr, "dc_bn_bwd_onepass: gamma required");
  if (p->flags & DC_BN_TRAIN) DC_REQUIRE(p->sums != nullptr, "dc_bn_bwd_onepass: forward workspace required in train mode");
  else DC_REQUIRE(p->running_mean && p->running_var, "dc_bn_bwd_onepass: running statistics required in eval mode");
  cudaStream_t st = as_stream(stream);
  return y.dtype == DC_F32 ? bn_bwd_onepass_t<float>(*p, dout, out, y, rws, dy, dres, dgamma, dbeta, st)
                           : bn_bwd_onepass_t<__nv_bfloat16>(*p, dout, out, y, rws, dy, dres, dgamma, dbeta, st);
}

/* channel sum with a caller-provided double workspace of C elements */
int dc_channel_sum(dc_view x, double* ws_c, float* out_c, int n_out, void* stream) {
  DC_REQUIRE(views_ok(x, nullptr, 0) && out_c != nullptr && ws_c != nullptr, "dc_channel_sum: bad arguments");
  DC_REQUIRE(n_out >= 1 && n_out <= x.c, "dc_channel_sum: n_out must be in 1..C");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(ws_c, 0, sizeof(double) * x.c, st);
  if (e != cudaSuccess) return dc::fail((int)e, "dc_channel_sum: %s", cudaGetErrorString(e));
  const long long npix = (long long)x.n * x.h * x.w;
  if (x.dtype == DC_F32) {
    LaneMap m = lane_map(x.c, 4);
    launch_k(channel_sum_kernel<float, 4>, bn_grid(m, npix, 8), dim3(kBnThreads), red_smem<1, 4>(), st, pix_view<const float>(x), ws_c, x.c, npix, m);
  } else {
    LaneMap m = lane_map(x.c, 8);
    launch_k(channel_sum_kernel<__nv_bfloat16, 8>, bn_grid(m, npix, 8), dim3(kBnThreads), red_smem<1, 8>(), st, pix_view<const __nv_bfloat16>(x), ws_c, x.c, npix, m);
  }
  double_to_float_kernel<<<ceil_div(n_out, 256), 256, 0, st>>>(ws_c, out_c, n_out);
  return launch_status("dc_channel_sum");
}

}  // extern "C"
