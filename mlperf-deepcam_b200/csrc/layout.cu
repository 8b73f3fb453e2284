// Layout / packing kernels: strided copies with dtype conversion (NCHW fp32 <-> NHWC bf16/fp32),
// weight packing from the fp32 master parameters, gradient unpacking, small utilities.
// These replace the .to()/.contiguous()/permute glue around the reference model (TR:348-349, DX:441).
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace dc {

static thread_local char g_err[512] = "ok";
char* err_buf() { return g_err; }
// PDL switch: on by default, DEEPCAM_B200_PDL=0 in the environment (read once) or dc_set_pdl(0) turns it off.
static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("DEEPCAM_B200_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- generic strided copy ----------------------------------------------------------------------
// Path A: both sides channel-contiguous -> vector of 4 channels per thread.
template <typename TS, typename TD>
__global__ void copy_vec4_kernel(View<const TS> s, View<TD> d, long long npix) {
  pdl_sync();
  const int cv = d.c >> 2;
  const int scv = s.c >> 2;
  long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = npix * cv;
  for (; item < total; item += (long long)gridDim.x * blockDim.x) {
    long long pix = item / cv;
    int c4 = (int)(item - pix * cv);
    int w = (int)(pix % d.w);
    long long t = pix / d.w;
    int h = (int)(t % d.h);
    int n = (int)(t / d.h);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < scv) v = elem<TS>::ld4(s.at(n, h, w) + c4 * 4);
    elem<TD>::st4(d.at(n, h, w) + c4 * 4, v);
  }
}

// Path B: tiled transpose between a w-contiguous side and a c-contiguous side (NCHW <-> NHWC).
// tile = 32 (w) x 32 (c); block = (32, 8)
template <typename TS, typename TD, bool SRC_W_CONTIG>
__global__ void copy_transpose_kernel(View<const TS> s, View<TD> d) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int wt = blockIdx.x * 32;
  const int ct = blockIdx.y * 32;
  const int nh = blockIdx.z;
  const int n = nh / d.h, h = nh % d.h;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (SRC_W_CONTIG) {
    // read: tx along w, ty along c
    for (int j = ty; j < 32; j += 8) {
      int c = ct + j, w = wt + tx;
      float v = 0.f;
      if (c < s.c && w < s.w) v = elem<TS>::ld(s.p + n * s.sn + h * s.sh + w * s.sw + c * s.sc);
      tile[j][tx] = v;   // [c][w]
    }
    __syncthreads();
    // write: tx along c, ty along w
    for (int j = ty; j < 32; j += 8) {
      int w = wt + j, c = ct + tx;
      if (c < d.c && w < d.w) elem<TD>::st(d.p + n * d.sn + h * d.sh + w * d.sw + c * d.sc, tile[tx][j]);
    }
  } else {
    // source is c-contiguous: read tx along c, ty along w
    for (int j = ty; j < 32; j += 8) {
      int w = wt + j, c = ct + tx;
      float v = 0.f;
      if (c < s.c && w < s.w) v = elem<TS>::ld(s.p + n * s.sn + h * s.sh + w * s.sw + c * s.sc);
      tile[j][tx] = v;   // [w][c]
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      int c = ct + j, w = wt + tx;
      if (c < d.c && w < d.w) elem<TD>::st(d.p + n * d.sn + h * d.sh + w * d.sw + c * d.sc, tile[tx][j]);
    }
  }
}

// Path B2: few channels (the 16-channel network input, the 3(+pad) logits and their gradient).  One thread per pixel: the
// planar side is read/written plane by plane (a warp covers 32 consecutive w = 128 contiguous bytes per plane), the pixel
// side as whole 16-byte vectors; no shared memory, CP independent loads in flight per thread.
template <typename T, int CP> struct pixvec;
template <int CP> struct pixvec<float, CP> {
  __device__ static __forceinline__ void store(float* p, const float (&v)[CP]) {
#pragma unroll
    for (int q = 0; q < CP / 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  __device__ static __forceinline__ void load(const float* p, float (&v)[CP]) {
#pragma unroll
    for (int q = 0; q < CP / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  }
};
template <int CP> struct pixvec<__nv_bfloat16, CP> {
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[CP]) {
#pragma unroll
    for (int q = 0; q < CP / 8; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 b = __floats2bfloat162_rn(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&b);
      }
      *reinterpret_cast<uint4*>(p + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[CP]) {
#pragma unroll
    for (int q = 0; q < CP / 8; ++q) {
      const uint4 t = *reinterpret_cast<const uint4*>(p + 8 * q);
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[8 * q + 2 * j] = __uint_as_float(w[j] << 16); v[8 * q + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
    }
  }
};

template <typename TS, typename TD, int CP>
__global__ void __launch_bounds__(256) copy_planar_to_pixel_kernel(View<const TS> s, View<TD> d, long long npix) {
  pdl_sync();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int w = (int)(pix % d.w);
  const long long t = pix / d.w;
  const int h = (int)(t % d.h), n = (int)(t / d.h);
  const TS* sp = s.p + n * s.sn + h * s.sh + w;               // s.sw == 1
  float v[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) v[c] = (c < s.c) ? elem<TS>::ld(sp + c * s.sc) : 0.f;
  pixvec<TD, CP>::store(d.at(n, h, w), v);
}
template <typename TS, typename TD, int CP>
__global__ void __launch_bounds__(256) copy_pixel_to_planar_kernel(View<const TS> s, View<TD> d, long long npix) {
  pdl_sync();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int w = (int)(pix % d.w);
  const long long t = pix / d.w;
  const int h = (int)(t % d.h), n = (int)(t / d.h);
  float v[CP];
  pixvec<TS, CP>::load(s.at(n, h, w), v);
  TD* dp = d.p + n * d.sn + h * d.sh + w;                     // d.sw == 1
#pragma unroll
  for (int c = 0; c < CP; ++c)
    if (c < d.c) elem<TD>::st(dp + c * d.sc, c < s.c ? v[c] : 0.f);
}
// the pixel side of a view: CP channels per pixel, unit channel stride, every pixel 16-byte aligned
static inline bool pixel_side_ok(const dc_view& v, int CP) {
  const size_t vb = (size_t)CP * dtype_size(v.dtype);
  const size_t es = dtype_size(v.dtype);
  return v.sc == 1 && vb % 16 == 0 && (v.sw * es) % 16 == 0 && (v.sh * es) % 16 == 0 && (v.sn * es) % 16 == 0 &&
         (reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0;
}

// Path C: fully generic scalar copy.
template <typename TS, typename TD>
__global__ void copy_scalar_kernel(View<const TS> s, View<TD> d, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % d.c);
    long long t = i / d.c;
    int w = (int)(t % d.w); t /= d.w;
    int h = (int)(t % d.h);
    int n = (int)(t / d.h);
    float v = 0.f;
    if (c < s.c) v = elem<TS>::ld(s.p + n * s.sn + h * s.sh + w * s.sw + c * s.sc);
    elem<TD>::st(d.p + n * d.sn + h * d.sh + w * d.sw + c * d.sc, v);
  }
}

template <typename TS, typename TD>
static int copy_dispatch(const dc_view& src, const dc_view& dst, cudaStream_t st) {
  View<const TS> s = make_view<const TS>(src);
  View<TD> d = make_view<TD>(dst);
  const long long npix = (long long)dst.n * dst.h * dst.w;
  if (view_vec4(src) && view_vec4(dst)) {
    long long total = npix * (dst.c / 4);
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
    launch_k(copy_vec4_kernel<TS, TD>, dim3(blocks), dim3(256), (size_t)0, st, s, d, npix);
  } else if (src.sw == 1 && src.c <= dst.c && (dst.c == 8 || dst.c == 16) && pixel_side_ok(dst, dst.c)) {
    const unsigned blocks = (unsigned)ceil_div64(npix, 256);
    if (dst.c == 8) launch_k(copy_planar_to_pixel_kernel<TS, TD, 8>, dim3(blocks), dim3(256), (size_t)0, st, s, d, npix);
    else launch_k(copy_planar_to_pixel_kernel<TS, TD, 16>, dim3(blocks), dim3(256), (size_t)0, st, s, d, npix);
  } else if (dst.sw == 1 && src.c <= src.sw && (src.sw == 8 || src.sw == 16) && pixel_side_ok(src, (int)src.sw) &&
             (reinterpret_cast<uintptr_t>(src.ptr) % ((size_t)src.sw * dtype_size(src.dtype))) == 0 && src.sh % src.sw == 0 &&
             src.sn % src.sw == 0) {
    // the source pixel is read as its whole pitch (8 or 16 elements, e.g. 3 logits padded to 8); pad lanes are ignored
    const unsigned blocks = (unsigned)ceil_div64(npix, 256);
    if (src.sw == 8) launch_k(copy_pixel_to_planar_kernel<TS, TD, 8>, dim3(blocks), dim3(256), (size_t)0, st, s, d, npix);
    else launch_k(copy_pixel_to_planar_kernel<TS, TD, 16>, dim3(blocks), dim3(256), (size_t)0, st, s, d, npix);
  } else if (src.sw == 1 && dst.sc == 1 && src.c <= dst.c) {
    dim3 grid(ceil_div(dst.w, 32), ceil_div(dst.c, 32), dst.n * dst.h);
    launch_k(copy_transpose_kernel<TS, TD, true>, grid, dim3(32, 8), (size_t)0, st, s, d);
  } else if (src.sc == 1 && dst.sw == 1 && src.c >= dst.c) {
    dim3 grid(ceil_div(dst.w, 32), ceil_div(dst.c, 32), dst.n * dst.h);
    launch_k(copy_transpose_kernel<TS, TD, false>, grid, dim3(32, 8), (size_t)0, st, s, d);
  } else {
    long long total = npix * dst.c;
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
    copy_scalar_kernel<TS, TD><<<blocks, 256, 0, st>>>(s, d, total);
  }
  return launch_status("dc_copy_view");
}

// ---- small utilities ---------------------------------------------------------------------------
// Input ingest: the CAM5 files store a sample as [H][W][C] fp32; the reference Dataset transposes it to CHW on
// the host and computes scale * (data - shift) per channel (DS:126-129).  Here the raw HWC block is normalised on the
// device in the layout it already has: dst[p][c] = (src[p][c] - shift[c]) * scale[c], stored as NHWC bf16 (or fp32, where
// it is bit-identical to the reference: separate subtract and multiply, no FMA contraction).
template <typename TD>
__global__ void __launch_bounds__(256) ingest_hwc_kernel(const float* __restrict__ src, const float* __restrict__ shift,
                                                         const float* __restrict__ scale, TD* __restrict__ dst, long long nvec, int cv) {
  pdl_sync();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const int c4 = (int)(i % cv) * 4;
    const float4 x = *reinterpret_cast<const float4*>(src + i * 4);
    const float4 sh = *reinterpret_cast<const float4*>(shift + c4);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c4);
    float4 o;
    o.x = __fmul_rn(__fsub_rn(x.x, sh.x), sc.x);
    o.y = __fmul_rn(__fsub_rn(x.y, sh.y), sc.y);
    o.z = __fmul_rn(__fsub_rn(x.z, sh.z), sc.z);
    o.w = __fmul_rn(__fsub_rn(x.w, sh.w), sc.w);
    elem<TD>::st4(dst + i * 4, o);
  }
}

__global__ void i64_increment_kernel(int64_t* const* ptrs, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) *ptrs[i] += 1;
}

__global__ void scale_f32_kernel(float* x, size_t count, float s) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t n4 = count / 4;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (size_t j = i; j < n4; j += stride) {
    float4 v = x4[j];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    x4[j] = v;
  }
  for (size_t j = n4 * 4 + i; j < count; j += stride) x[j] *= s;
}

// ---- weight packing ----------------------------------------------------------------------------
// DC_PACK_NTK_CONVT2: destination element (row nn, tap t, k) of the fused stride-2 transposed-convolution pack -> source
// index in the [k][n][3][3] parameter, or -1 (zero).  nn = (a*2+b)*G + n, t = dh*2+dw, kernel index = (a+1-2dh, b+1-2dw).
__device__ __forceinline__ long long convt2_src_index(int k, int nn, int t, int K, int N, int N_pad) {
  const int G = N_pad >> 2;
  const int cls = nn / G, n = nn - cls * G;
  const int kh = (cls >> 1) + 1 - 2 * (t >> 1), kw = (cls & 1) + 1 - 2 * (t & 1);
  if (k >= K || n >= N || kh < 0 || kh > 2 || kw < 0 || kw > 2) return -1;
  return ((long long)k * N + n) * 9 + kh * 3 + kw;
}

// one thread per destination element; destinations are small (<= 4.7 M elements per layer).
template <typename TD>
__global__ void pack_weight_kernel(const float* __restrict__ src, int K, int N, int taps, int src_k_first,
                                   TD* __restrict__ dst, int layout, int K_pad, int N_pad) {
  long long total = (long long)taps * K_pad * N_pad;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t, k, n;
    if (layout == DC_PACK_TKN) {          // [tap][k][n_pad]
      n = (int)(i % N_pad);
      long long r = i / N_pad;
      k = (int)(r % K_pad);
      t = (int)(r / K_pad);
    } else {                               // [n][tap][k_pad]
      k = (int)(i % K_pad);
      long long r = i / K_pad;
      t = (int)(r % taps);
      n = (int)(r / taps);
    }
    float v = 0.f;
    if (layout == DC_PACK_NTK_CONVT2) {
      const long long si = convt2_src_index(k, n, t, K, N, N_pad);
      if (si >= 0) v = src[si];
    } else if (k < K && n < N) {
      long long si = src_k_first ? ((long long)k * N + n) * taps + t : ((long long)n * K + k) * taps + t;
      v = src[si];
    }
    elem<TD>::st(dst + i, v);
  }
}

// All weight packs of a step in ONE launch: a device table of jobs, each owning a contiguous range of blocks.
template <typename TD>
__device__ __forceinline__ void pack_job_elems(const dc_pack_job& j, int local_block, int tid) {
  const int total = j.taps * j.K_pad * j.N_pad;
  const float* __restrict__ src = j.src;
  TD* __restrict__ dst = reinterpret_cast<TD*>(j.dst);
  for (int i = local_block * 256 + tid; i < total; i += j.n_blocks * 256) {
    int t, k, n;
    if (j.layout == DC_PACK_TKN) {
      n = i % j.N_pad;
      const int r = i / j.N_pad;
      k = r % j.K_pad;
      t = r / j.K_pad;
    } else {
      k = i % j.K_pad;
      const int r = i / j.K_pad;
      t = r % j.taps;
      n = r / j.taps;
    }
    float v = 0.f;
    if (j.layout == DC_PACK_NTK_CONVT2) {
      const long long si = convt2_src_index(k, n, t, j.K, j.N, j.N_pad);
      if (si >= 0) v = src[si];
    } else if (k < j.K && n < j.N) {
      const long long si = j.src_k_first ? ((long long)k * j.N + n) * j.taps + t : ((long long)n * j.K + k) * j.taps + t;
      v = src[si];
    }
    elem<TD>::st(dst + i, v);
  }
}

// Packs with more than one tap, or with swapped roles, go through shared memory so that BOTH sides are coalesced; the
// element-wise path reads one 4-byte word per 36-byte stride (3x3 weights) or per row (transposes), i.e. one useful
// word per 32-byte sector.
//
// (a) source [n][k][taps] -> destination [n][tap][k_pad] (fprop role of a 3x3 Conv2d, dgrad role of a ConvTranspose2d):
//     a work item is one n and 256 consecutive k: 256*taps contiguous source floats in, `taps` runs of 256 elements out.
// The tap count is a template parameter (1 and 9 cover the network; 0 = run-time value) so that the index arithmetic is
// shifts and constant divisions: with run-time divisors the pack was ALU-bound (~100 instructions per element).
constexpr int kPackMaxTaps = 9;
template <typename TD, int TAPS>
__device__ __forceinline__ void pack_job_taps(const dc_pack_job& j, int local_block, int tid, float* sm) {
  const int taps = TAPS ? TAPS : j.taps;
  const int kchunks = (j.K_pad + 255) >> 8;
  const int nitems = j.N_pad * kchunks;
  const float* __restrict__ src = j.src;
  TD* __restrict__ dst = reinterpret_cast<TD*>(j.dst);
  for (int it = local_block; it < nitems; it += j.n_blocks) {
    const int n = it / kchunks, k0 = (it - n * kchunks) << 8;
    const int kcount = min(256, j.K - k0);                         // valid k of this chunk (<= 0: pure padding)
    const int nfl = (n < j.N && kcount > 0) ? kcount * taps : 0;
    const float* sp = src + ((long long)n * j.K + k0) * taps;
    for (int e = tid; e < nfl; e += 256) sm[e] = sp[e];
    __syncthreads();
    const int kk = k0 + tid;
    if (kk < j.K_pad) {
#pragma unroll
      for (int t = 0; t < (TAPS ? TAPS : kPackMaxTaps); ++t) {
        if (t >= taps) break;
        const float v = (tid * taps + t < nfl) ? sm[tid * taps + t] : 0.f;
        elem<TD>::st(dst + ((long long)n * taps + t) * j.K_pad + kk, v);
      }
    }
    __syncthreads();
  }
}

// (a') taps == 1, same role: a plain fp32 -> storage-type conversion of [n][k] rows into [n][k_pad] rows, four k per thread
template <typename TD>
__device__ __forceinline__ void pack_job_rows(const dc_pack_job& j, int local_block, int tid) {
  const int kq = (j.K_pad + 3) >> 2;                     // quads per destination row (K_pad % 4 == 0 for NTK packs)
  const int qchunks = (kq + 255) >> 8;
  const int nitems = j.N_pad * qchunks;
  const bool vec = (j.K % 4 == 0) && ((reinterpret_cast<uintptr_t>(j.src) & 15) == 0);
  TD* __restrict__ dst = reinterpret_cast<TD*>(j.dst);
  // four items per pass: their four 16-byte loads are in flight together (one load per thread and pass left the kernel latency-bound)
  for (int it0 = local_block; it0 < nitems; it0 += 4 * j.n_blocks) {
    float4 v[4];
    long long doff[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = it0 + u * j.n_blocks;
      doff[u] = -1;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (it >= nitems) continue;
      const int n = it / qchunks, q = ((it - n * qchunks) << 8) + tid;
      if (q >= kq) continue;
      const int k = q << 2;
      doff[u] = (long long)n * j.K_pad + k;
      if (n < j.N) {
        const float* sp = j.src + (long long)n * j.K + k;
        if (vec && k + 4 <= j.K) v[u] = *reinterpret_cast<const float4*>(sp);
        else {
          if (k < j.K) v[u].x = sp[0];
          if (k + 1 < j.K) v[u].y = sp[1];
          if (k + 2 < j.K) v[u].z = sp[2];
          if (k + 3 < j.K) v[u].w = sp[3];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (doff[u] >= 0) elem<TD>::st4(dst + doff[u], v[u]);
  }
}

// (b) source [k][n][taps] -> destination [n][tap][k_pad] (dgrad role of a Conv2d, fprop role of a ConvTranspose2d): a work
//     item is 64 k x TN n (TN*taps <= 72 floats per k row, contiguous in the source); the destination gets runs of 64 k.
// 8 consecutive destination elements (k run) as one 16-byte (bf16) / two 16-byte (fp32) stores
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
    w[q] = *reinterpret_cast<uint32_t*>(&b);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename TD, int TAPS>
__device__ __forceinline__ void pack_job_transpose(const dc_pack_job& j, int local_block, int tid, float* sm) {
  const int taps = TAPS ? TAPS : j.taps;
  const int TN = TAPS == 1 ? 32 : (TAPS == 9 ? 8 : (taps == 1 ? 32 : (taps <= 2 ? 16 : 8)));
  const int RW = TN * taps;                      // floats per k row of the tile (32 / 72 when TAPS is 1 / 9)
  const int pitch = RW | 1;                      // odd pitch: conflict-free column reads
  const int kt = (j.K_pad + 63) >> 6, ntl = (j.N_pad + TN - 1) / TN;
  const int ntiles = kt * ntl;
  const float* __restrict__ src = j.src;
  TD* __restrict__ dst = reinterpret_cast<TD*>(j.dst);
  const long long krow = (long long)j.N * taps;
  // 16-byte source loads need 4-float alignment of every k row segment; 16-byte destination stores need K_pad % 8 == 0
  const bool vec_src = (RW % 4 == 0) && (krow % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const bool vec_dst = (j.K_pad % 8 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  const int RW4 = RW >> 2;
  for (int tl = local_block; tl < ntiles; tl += j.n_blocks) {
    const int ktile = tl / ntl;
    const int k0 = ktile << 6, n0 = (tl - ktile * ntl) * TN;
    const int rw_valid = max(0, min(RW, (j.N - n0) * taps));
    const float* sp = src + ((long long)k0 * j.N + n0) * taps;
    if (vec_src && rw_valid == RW && ((n0 * taps) & 3) == 0) {
      // all of a thread's 16-byte loads of the tile (2 for one tap, 5 for nine) are issued before the first shared-memory store
      float4 v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int e = tid + u * 256;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < 64 * RW4) {
          const int k = e / RW4, r = (e - k * RW4) << 2;
          if (k0 + k < j.K) v[u] = *reinterpret_cast<const float4*>(sp + k * krow + r);
        }
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int e = tid + u * 256;
        if (e < 64 * RW4) {
          const int k = e / RW4, r = (e - k * RW4) << 2;
          float* d = sm + k * pitch + r;
          d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
        }
      }
    } else {
      for (int e = tid; e < 64 * RW; e += 256) {
        const int k = e / RW, r = e - k * RW;
        float v = 0.f;
        if (k0 + k < j.K && r < rw_valid) v = sp[k * krow + r];
        sm[k * pitch + r] = v;
      }
    }
    __syncthreads();
    if (vec_dst) {
      // 8 k per thread: 8 rows (n, tap) x 8 k-groups per pass of 64 threads
      for (int e = tid; e < 8 * RW; e += 256) {
        const int kg = e & 7, row = e >> 3;       // row = n_local * taps + t
        const int nl = row / taps, t = row - nl * taps;
        const int kk = kg << 3;
        if (n0 + nl < j.N_pad && k0 + kk < j.K_pad) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = sm[(kk + q) * pitch + row];
          st8(dst + ((long long)(n0 + nl) * taps + t) * j.K_pad + k0 + kk, v);
        }
      }
    } else {
      for (int e = tid; e < 64 * RW; e += 256) {
        const int kk = e & 63, row = e >> 6;
        const int nl = row / taps, t = row - nl * taps;
        if (n0 + nl < j.N_pad && k0 + kk < j.K_pad)
          elem<TD>::st(dst + ((long long)(n0 + nl) * taps + t) * j.K_pad + k0 + kk, sm[kk * pitch + row]);
      }
    }
    __syncthreads();
  }
}

template <typename TD>
__device__ __forceinline__ void pack_job_ntk(const dc_pack_job& j, int lb, int tid, float* sm) {
  if (j.src_k_first) {
    if (j.taps == 1) pack_job_transpose<TD, 1>(j, lb, tid, sm);
    else if (j.taps == 9) pack_job_transpose<TD, 9>(j, lb, tid, sm);
    else pack_job_transpose<TD, 0>(j, lb, tid, sm);
  } else {
    if (j.taps == 1) pack_job_rows<TD>(j, lb, tid);
    else if (j.taps == 9) pack_job_taps<TD, 9>(j, lb, tid, sm);
    else pack_job_taps<TD, 0>(j, lb, tid, sm);
  }
}

__global__ void __launch_bounds__(256, 6) pack_multi_kernel(const dc_pack_job* __restrict__ jobs, int njobs) {
  pdl_sync();
  // job of this block = last job whose block_start <= blockIdx.x.  The table's block_start column is first copied to shared memory
  // with ONE round of parallel loads: the binary search straight on global memory was 8 dependent L2 round trips per block, ~2.9 of
  // the 7.6 us an average block lived (ncu source page, round 2: 13 % of the samples on those nine instructions)
  __shared__ int s_start[512];
  const bool staged = njobs <= 512;
  if (staged) {
    for (int i = threadIdx.x; i < njobs; i += 256) s_start[i] = jobs[i].block_start;
    __syncthreads();
  }
  int lo = 0, hi = njobs - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((staged ? s_start[mid] : jobs[mid].block_start) <= b) lo = mid; else hi = mid - 1;
  }
  const dc_pack_job j = jobs[lo];
  __shared__ float sm[64 * 73];                 // 18.7 KB: 64 x (72 | 1) transpose tile, or 256 x 9 tap chunk
  const int lb = b - j.block_start;
  if (j.layout == DC_PACK_NTK && j.taps <= kPackMaxTaps && (j.K_pad & 3) == 0) {
    if (j.dst_dtype == DC_F32) pack_job_ntk<float>(j, lb, threadIdx.x, sm);
    else pack_job_ntk<__nv_bfloat16>(j, lb, threadIdx.x, sm);
    return;
  }
  if (j.dst_dtype == DC_F32) pack_job_elems<float>(j, lb, threadIdx.x);
  else pack_job_elems<__nv_bfloat16>(j, lb, threadIdx.x);
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ G, int K, int N, int taps, int k_stride, int dst_k_first,
                                    float* __restrict__ dst) {
  pdl_sync();
  long long total = (long long)taps * K * N;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes dst
    int t = (int)(i % taps);
    long long r = i / taps;
    int k, n;
    if (dst_k_first) { n = (int)(r % N); k = (int)(r / N); }
    else             { k = (int)(r % K); n = (int)(r / K); }
    dst[i] = G[((long long)t * N + n) * k_stride + k];
  }
}

// deterministic second stage of the split weight-gradient reductions (common.cuh)
template <int VEC>
__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ ws, int nslices, long long slice_stride, SplitReduceTaps taps,
                                                           long long per_tap, float* __restrict__ dst) {
  pdl_sync();
  const long long off = (long long)taps.wt[blockIdx.y] * per_tap;
  const long long n = per_tap / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (VEC == 4) {
      const float4* src = reinterpret_cast<const float4*>(ws + off) + i;
      float4 a = *src;
      for (int z = 1; z < nslices; ++z) {
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (long long)z * slice_stride);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      float4* d = reinterpret_cast<float4*>(dst + off) + i;
      float4 o = *d;
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      *d = o;
    } else {
      float a = ws[off + i];
      for (int z = 1; z < nslices; ++z) a += ws[(long long)z * slice_stride + off + i];
      dst[off + i] += a;
    }
  }
}

int launch_split_reduce(const float* ws, int nslices, long long slice_stride, const SplitReduceTaps& taps, long long per_tap, float* dst,
                        cudaStream_t st) {
  const bool vec = per_tap % 4 == 0 && slice_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(ws) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst) % 16) == 0;
  const long long n = vec ? per_tap / 4 : per_tap;
  const int bx = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)kNumSMs * 8 / std::max(1, taps.n)));
  dim3 grid((unsigned)bx, (unsigned)taps.n);
  if (vec) launch_k(split_reduce_kernel<4>, grid, dim3(256), (size_t)0, st, ws, nslices, slice_stride, taps, per_tap, dst);
  else launch_k(split_reduce_kernel<1>, grid, dim3(256), (size_t)0, st, ws, nslices, slice_stride, taps, per_tap, dst);
  return launch_status("dc_split_reduce");
}

static int g_deterministic = 0;
bool deterministic() { return g_deterministic != 0; }

}  // namespace dc

using namespace dc;

extern "C" {

int dc_abi_version(void) { return DC_ABI_VERSION; }
const char* dc_last_error_string(void) { return dc::err_buf(); }
int dc_set_pdl(int on) { dc::set_pdl(on); return 0; }
int dc_get_pdl(void) { return dc::pdl_enabled() ? 1 : 0; }
int dc_set_deterministic(int on) { dc::g_deterministic = on ? 1 : 0; return 0; }
int dc_get_deterministic(void) { return dc::g_deterministic; }

int dc_device_supports_tcgen05(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { dc::fail(-1, "cudaGetDevice: %s", cudaGetErrorString(e)); return -1; }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return (major == 10 && minor == 0) ? 1 : 0;
}

int dc_copy_view(dc_view src, dc_view dst, void* stream) {
  DC_REQUIRE(view_ok(src) && view_ok(dst), "dc_copy_view: bad view");
  DC_REQUIRE(src.n == dst.n && src.h == dst.h && src.w == dst.w, "dc_copy_view: shape mismatch");
  cudaStream_t st = as_stream(stream);
  if (src.dtype == DC_F32 && dst.dtype == DC_F32) return copy_dispatch<float, float>(src, dst, st);
  if (src.dtype == DC_F32 && dst.dtype == DC_BF16) return copy_dispatch<float, __nv_bfloat16>(src, dst, st);
  if (src.dtype == DC_BF16 && dst.dtype == DC_F32) return copy_dispatch<__nv_bfloat16, float>(src, dst, st);
  return copy_dispatch<__nv_bfloat16, __nv_bfloat16>(src, dst, st);
}

int dc_fill_zero(void* ptr, size_t bytes, void* stream) {
  DC_REQUIRE(ptr != nullptr || bytes == 0, "dc_fill_zero: null pointer");
  if (bytes == 0) return 0;
  cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, as_stream(stream));
  if (e != cudaSuccess) return dc::fail((int)e, "dc_fill_zero: %s", cudaGetErrorString(e));
  return 0;
}

int dc_ingest_hwc(const float* src, long long npix, int C, const float* shift, const float* scale, void* dst, int dst_dtype,
                  void* stream) {
  DC_REQUIRE(src && shift && scale && dst && npix > 0 && C > 0 && C % 4 == 0, "dc_ingest_hwc: bad arguments (C %% 4 == 0 required)");
  DC_REQUIRE(dst_dtype == DC_F32 || dst_dtype == DC_BF16, "dc_ingest_hwc: dst dtype must be fp32 or bf16");
  DC_REQUIRE((reinterpret_cast<uintptr_t>(src) % 16) == 0 && (reinterpret_cast<uintptr_t>(dst) % 16) == 0 &&
             (reinterpret_cast<uintptr_t>(shift) % 16) == 0 && (reinterpret_cast<uintptr_t>(scale) % 16) == 0,
             "dc_ingest_hwc: pointers must be 16-byte aligned");
  const long long nvec = npix * (C / 4);
  const unsigned blocks = (unsigned)std::min<long long>((nvec + 255) / 256, (long long)kNumSMs * 8);
  if (dst_dtype == DC_F32)
    launch_k(ingest_hwc_kernel<float>, dim3(blocks), dim3(256), (size_t)0, as_stream(stream), src, shift, scale, (float*)dst, nvec, C / 4);
  else
    launch_k(ingest_hwc_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), (size_t)0, as_stream(stream), src, shift, scale, (__nv_bfloat16*)dst,
             nvec, C / 4);
  return launch_status("dc_ingest_hwc");
}

int dc_i64_increment_many(int64_t* const* ptrs, int count, void* stream) {
  DC_REQUIRE(ptrs != nullptr && count >= 0, "dc_i64_increment_many: bad arguments");
  if (count == 0) return 0;
  i64_increment_kernel<<<ceil_div(count, 128), 128, 0, as_stream(stream)>>>(ptrs, count);
  return launch_status("dc_i64_increment_many");
}

int dc_scale_f32(float* x, size_t count, float s, void* stream) {
  DC_REQUIRE(x != nullptr, "dc_scale_f32: null pointer");
  DC_REQUIRE((reinterpret_cast<uintptr_t>(x) % 16) == 0, "dc_scale_f32: pointer must be 16-byte aligned");
  if (count == 0) return 0;
  int blocks = (int)std::min<size_t>((count / 4 + 255) / 256 + 1, (size_t)kNumSMs * 8);
  scale_f32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, count, s);
  return launch_status("dc_scale_f32");
}

int dc_pack_weight(const float* src, int K, int N, int taps, int src_k_first, void* dst, int layout,
                   int K_pad, int N_pad, int dst_dtype, void* stream) {
  DC_REQUIRE(src && dst && K > 0 && N > 0 && taps > 0 && K_pad >= K && N_pad >= N, "dc_pack_weight: bad arguments");
  DC_REQUIRE(layout == DC_PACK_TKN || layout == DC_PACK_NTK || layout == DC_PACK_NTK_CONVT2, "dc_pack_weight: unknown layout %d", layout);
  DC_REQUIRE(layout != DC_PACK_NTK_CONVT2 || (taps == 4 && src_k_first == 1 && N_pad % 4 == 0 && N_pad / 4 >= N),
             "dc_pack_weight: DC_PACK_NTK_CONVT2 needs taps = 4, src_k_first = 1, N_pad = 4 * (channels per class >= N)");
  long long total = (long long)taps * K_pad * N_pad;
  int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 8);
  if (dst_dtype == DC_F32)
    pack_weight_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(src, K, N, taps, src_k_first, (float*)dst, layout, K_pad, N_pad);
  else if (dst_dtype == DC_BF16)
    pack_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(src, K, N, taps, src_k_first, (__nv_bfloat16*)dst, layout, K_pad, N_pad);
  else
    return dc::fail(-1, "dc_pack_weight: bad dtype %d", dst_dtype);
  return launch_status("dc_pack_weight");
}

int dc_pack_weights_multi(const dc_pack_job* jobs_dev, int njobs, int total_blocks, void* stream) {
  DC_REQUIRE(jobs_dev != nullptr && njobs > 0 && total_blocks > 0, "dc_pack_weights_multi: bad arguments");
  launch_k(pack_multi_kernel, dim3(total_blocks), dim3(256), (size_t)0, as_stream(stream), jobs_dev, njobs);
  return launch_status("dc_pack_weights_multi");
}

int dc_unpack_wgrad(const float* G, int K, int N, int taps, int k_stride, int dst_k_first, float* dst, void* stream) {
  DC_REQUIRE(G && dst && K > 0 && N > 0 && taps > 0 && k_stride >= K, "dc_unpack_wgrad: bad arguments");
  long long total = (long long)taps * K * N;
  int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 8);
  launch_k(unpack_wgrad_kernel, dim3(blocks), dim3(256), (size_t)0, as_stream(stream), G, K, N, taps, k_stride, dst_k_first, dst);
  return launch_status("dc_unpack_wgrad");
}

}  // extern "C"
