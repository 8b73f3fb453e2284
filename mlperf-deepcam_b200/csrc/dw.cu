// Depthwise 3x3 convolution of SeparableConv2d_same (DX:54-66): fixed_padding (DX:45-51) is folded into
// the index arithmetic (implicit halo, no padded copy), groups = C, no bias, stride 1|2, dilation 1|2.
//   fwd : out[n,y,x,c]  = sum_{kh,kw} in[n, y*s - d + kh*d, x*s - d + kw*d, c] * w[kh*3+kw][c]
//   bwdD: din[n,h,w,c]  = sum_{kh,kw} dout[n, (h + d - kh*d)/s, (w + d - kw*d)/s, c] * w[kh*3+kw][c]   (when divisible)
//   bwdW: G[kh*3+kw][c] = sum_{n,y,x} in[n, y*s - d + kh*d, x*s - d + kw*d, c] * dout[n,y,x,c]
// HBM-bound (AI ~ 4 FLOP/B): algorithmic bytes = in + out (+ 9*C weights).
//
// Thread mapping: a thread owns one 16-byte channel vector (V = 8 bf16 / 4 fp32) of one pixel COLUMN and walks down
// the rows of a strip; the 32 lanes of a warp cover 32 consecutive channel vectors of a pixel (512 contiguous bytes)
// or several consecutive pixels when a pixel has fewer vectors; the 8 warps of a block sit on consecutive pixels,
// so the x-1 / x+1 taps are L1 hits.  The stride-1 dilation-1 kernels (59 of the 63 layers, and every backward-data
// of a stride-1 layer = the same kernel with the 3x3 filter flipped) are input-stationary: each input row is loaded
// and unpacked once per strip and scattered into the three live output rows with packed fp32 FMAs (FFMA2).
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace dc {

// ---- TMA tile fill of the staged kernels -----------------------------------------------------------------------------------
// The staged kernels used to fill their shared-memory tiles with one 16-byte cp.async per thread and vector: 43-61 instructions
// per vector for address + bounds arithmetic, 46 % of all instructions of the weight-gradient kernel on the 113 MB tensors (ncu
// source page, round 2).  One un-swizzled 4-D TMA box per tile ({cvp*V channels, tile width, tile rows, 1 image} of the NHWC
// tensor, out-of-range pixels and channels zero-filled = fixed_padding) is issued by ONE thread and lands in exactly the
// [row][pixel][channel vector] layout the row walk reads.
__device__ __forceinline__ void dw_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void dw_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dw_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void dw_tma_load_5d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void dw_tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// tile box of virtual image vn: rank-4 map {C, W, H, N} for plain views; rank-5 map {C, W/d, H/d, d (px), N*H (n*H + py)} for the
// parity sub-grids of a dilated layer (DwView::dsub = d, hsub = sub-grid height)
__device__ __forceinline__ void dw_tma_load_tile(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int x, int y, int vn, int dsub, int hsub) {
  if (dsub == 1) {
    dw_tma_load_4d(map, bar, dst, c0, x, y, vn);
  } else {
    const int px = vn % dsub, t = vn / dsub;
    const int py = t % dsub, n = t / dsub;
    dw_tma_load_5d(map, bar, dst, c0, x, y, px, n * hsub * dsub + py);
  }
}

// ---- 16-byte channel vectors -------------------------------------------------------------------------------
template <typename T> struct dwvec;
template <> struct dwvec<float> {
  static constexpr int V = 4;
  __device__ static __forceinline__ void unpack(const uint4& t, float (&f)[4]) {
    f[0] = __uint_as_float(t.x); f[1] = __uint_as_float(t.y); f[2] = __uint_as_float(t.z); f[3] = __uint_as_float(t.w);
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <> struct dwvec<__nv_bfloat16> {
  static constexpr int V = 8;
  __device__ static __forceinline__ void unpack(const uint4& t, float (&f)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w[j] << 16); f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&b);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
// channel pairs as float2 (operands of the packed fp32 FMA, FFMA2 on sm_100)
template <typename T> struct dwpair;
template <> struct dwpair<float> {
  static constexpr int VP = 2;
  __device__ static __forceinline__ void unpack(const uint4& t, float2 (&f)[2]) {
    f[0] = make_float2(__uint_as_float(t.x), __uint_as_float(t.y));
    f[1] = make_float2(__uint_as_float(t.z), __uint_as_float(t.w));
  }
  __device__ static __forceinline__ uint4 pack(const float2 (&f)[2]) {
    return make_uint4(__float_as_uint(f[0].x), __float_as_uint(f[0].y), __float_as_uint(f[1].x), __float_as_uint(f[1].y));
  }
};
template <> struct dwpair<__nv_bfloat16> {
  static constexpr int VP = 4;
  __device__ static __forceinline__ void unpack(const uint4& t, float2 (&f)[4]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] = make_float2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u));
  }
  __device__ static __forceinline__ uint4 pack(const float2 (&f)[4]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 b = __floats2bfloat162_rn(f[j].x, f[j].y);
      w[j] = *reinterpret_cast<uint32_t*>(&b);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
__device__ __forceinline__ float2 fma2(const float2& a, const float2& b, const float2& c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(const float2& a, const float2& b) { return __fmul2_rn(a, b); }

__device__ __forceinline__ uint4 ld16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

struct DwMap {
  int cv, cvp, ppw, ppb, gy;   // channel vectors / lanes per pixel / pixels per warp / pixels per block / channel blocks
  int rs, nstrips;             // rows per strip, strips per image
};
static inline DwMap dw_map(int C, int V, int rows, int cols, int n_img, int target_blocks, int min_rs) {
  DwMap m;
  m.cv = C / V;
  m.cvp = 1;
  while (m.cvp < 32 && m.cvp < m.cv) m.cvp <<= 1;
  m.ppw = 32 / m.cvp;
  m.ppb = 8 * m.ppw;
  m.gy = ceil_div(m.cv, m.cvp);
  const long long per_strip = (long long)ceil_div(cols, m.ppb) * m.gy * n_img;
  int ns = (int)std::max<long long>(1, target_blocks / std::max<long long>(1, per_strip));
  ns = std::min(ns, std::max(1, rows / min_rs));
  m.rs = ceil_div(rows, ns);
  m.nstrips = ceil_div(rows, m.rs);
  return m;
}

// dsub > 1: the view addresses the dsub x dsub PARITY SUB-GRIDS of the tensor as separate images (virtual image index
// vn = (n * dsub + py) * dsub + px, h / w = sub-grid size, sh / sw = dsub x the tensor's strides).  A stride-1 depthwise
// convolution with dilation d never mixes the parity classes modulo d, and the zero padding of d pixels is a padding of one
// sub-grid pixel, so a dilation-d layer IS d*d independent dilation-1 layers: every s1d1 kernel below serves dilation 2
// (exit flow, DX:176-186) and 4 (os=8) through such a view, in one launch.
template <typename T>
struct DwView {
  T* p;
  int h, w;
  long long sn, sh, sw;
  int dsub;
  __device__ __forceinline__ long long img(int vn) const {
    if (dsub == 1) return (long long)vn * sn;
    const int px = vn % dsub, t = vn / dsub;
    const int py = t % dsub, n = t / dsub;
    return (long long)n * sn + (long long)py * (sh / dsub) + (long long)px * (sw / dsub);
  }
};
template <typename T>
static inline DwView<T> dw_view(const dc_view& v, int dsub = 1) {
  DwView<T> r;
  r.p = reinterpret_cast<T*>(v.ptr);
  r.h = v.h / dsub; r.w = v.w / dsub; r.sn = v.sn; r.sh = v.sh * dsub; r.sw = v.sw * dsub;
  r.dsub = dsub;
  return r;
}

constexpr int kDwThreads = 256;

// common thread decode
struct DwLane {
  int cvi, x, n, y0, y1;
  bool ok;
};
__device__ __forceinline__ DwLane dw_lane(const DwMap& m, int rows, int cols) {
  DwLane l;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  l.cvi = blockIdx.y * m.cvp + cvl;
  l.x = (blockIdx.x * 8 + warp) * m.ppw + psub;
  l.n = blockIdx.z / m.nstrips;
  const int strip = blockIdx.z - l.n * m.nstrips;
  l.y0 = strip * m.rs;
  l.y1 = min(rows, l.y0 + m.rs);
  l.ok = (l.cvi < m.cv) && (l.x < cols);
  return l;
}

template <typename T, int V>
__device__ __forceinline__ void load_weights(const T* __restrict__ w9c, int C, int c0, bool flip, float (&wv)[9][V]) {
#pragma unroll
  for (int k = 0; k < 9; ++k) dwvec<T>::unpack(ld16(w9c + (size_t)(flip ? 8 - k : k) * C + c0), wv[k]);
}

// ---- stride 1, dilation 1 (forward, and backward-data with flip = 1): input-stationary row walk ----------------------
// Every input row is loaded and unpacked ONCE; its three vectors update the three live output rows (r+1, r, r-1) with
// packed fp32 FMAs, then output row r-1 is complete and stored.  No im2col, no padded copy, ~70 instructions per
// 8-channel output instead of ~165 for the gather form, so the kernel stays HBM-bound.
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) dw_s1d1_kernel(DwView<const T> in, const T* __restrict__ w9c, DwView<T> out, int C,
                                                             DwMap m, int flip, int accumulate) {
  constexpr int VP = V / 2;
  const DwLane l = dw_lane(m, out.h, out.w);
  pdl_sync();
  if (!l.ok) return;
  const int c0 = l.cvi * V;
  float2 wv[9][VP];
#pragma unroll
  for (int k = 0; k < 9; ++k) dwpair<T>::unpack(ld16(w9c + (size_t)(flip ? 8 - k : k) * C + c0), wv[k]);
  const int H = in.h, W = in.w;
  const T* base = in.p + in.img(l.n) + (long long)l.x * in.sw + c0;
  T* obase = out.p + out.img(l.n) + (long long)l.x * out.sw + c0;
  const bool xl = l.x >= 1, xr = l.x + 1 < W;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  auto ldrow = [&](int r, uint4 (&v)[3]) {
    if (r < 0 || r >= H) { v[0] = zero; v[1] = zero; v[2] = zero; return; }
    const T* rp = base + (long long)r * in.sh;
    v[1] = ld16(rp);
    v[0] = xl ? ld16(rp - in.sw) : zero;
    v[2] = xr ? ld16(rp + in.sw) : zero;
  };
  // A = output row r-1 (receives filter row 2 and is finished), B = output row r (filter row 1), Cn = output row r+1
  // (filter row 0, first contribution)
  auto step = [&](int r, const uint4 (&v)[3], float2 (&A)[VP], float2 (&B)[VP], float2 (&Cn)[VP]) {
    float2 f[3][VP];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) dwpair<T>::unpack(v[kw], f[kw]);
#pragma unroll
    for (int j = 0; j < VP; ++j) {
      Cn[j] = mul2(f[0][j], wv[0][j]);
      Cn[j] = fma2(f[1][j], wv[1][j], Cn[j]);
      Cn[j] = fma2(f[2][j], wv[2][j], Cn[j]);
      B[j] = fma2(f[0][j], wv[3][j], B[j]);
      B[j] = fma2(f[1][j], wv[4][j], B[j]);
      B[j] = fma2(f[2][j], wv[5][j], B[j]);
      A[j] = fma2(f[0][j], wv[6][j], A[j]);
      A[j] = fma2(f[1][j], wv[7][j], A[j]);
      A[j] = fma2(f[2][j], wv[8][j], A[j]);
    }
    const int y = r - 1;
    if (y >= l.y0 && y < l.y1) {
      T* op = obase + (long long)y * out.sh;
      if (accumulate) {
        float2 old[VP];
        dwpair<T>::unpack(ld16(op), old);
#pragma unroll
        for (int j = 0; j < VP; ++j) { A[j].x += old[j].x; A[j].y += old[j].y; }
      }
      st16(op, dwpair<T>::pack(A));
    }
  };
  float2 a0[VP], a1[VP], a2[VP];
#pragma unroll
  for (int j = 0; j < VP; ++j) { a0[j] = make_float2(0.f, 0.f); a1[j] = a0[j]; a2[j] = a0[j]; }
  uint4 cur[3], nxt[3];
  int r = l.y0 - 1;
  ldrow(r, cur);
  while (true) {
    ldrow(r + 1, nxt);
    step(r, cur, a0, a1, a2);
    if (++r > l.y1) break;
    ldrow(r + 1, cur);
    step(r, nxt, a1, a2, a0);
    if (++r > l.y1) break;
    ldrow(r + 1, nxt);
    step(r, cur, a2, a0, a1);
    if (++r > l.y1) break;
    ldrow(r + 1, cur);
    step(r, nxt, a0, a1, a2);
    if (++r > l.y1) break;
    ldrow(r + 1, nxt);
    step(r, cur, a1, a2, a0);
    if (++r > l.y1) break;
    ldrow(r + 1, cur);
    step(r, nxt, a2, a0, a1);
    if (++r > l.y1) break;
    // after six steps the accumulator roles are back to (a0, a1, a2) and `cur` holds row r again
  }
}

// ---- stride 1, dilation 1, shared-memory staged ---------------------------------------------------------------------------
// Same arithmetic as dw_s1d1_kernel, but the block first pulls its whole input tile ((rs+2) rows x (ppb+2) pixels x cvp
// channel vectors, <= 48 KB) into shared memory with 16-byte cp.async copies that are ALL in flight at once (zero-filled
// outside the image = fixed_padding), and only then walks the rows out of shared memory.  The register-pipelined kernel
// above has one row (3 loads) in flight per thread, i.e. a chain of rs+2 dependent L2/HBM round trips; for the 10 MB
// middle-flow tensors that chain, not bandwidth, was the kernel's duration.
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <typename T, int V, bool kTma>
__global__ void __launch_bounds__(kDwThreads) dw_s1d1_tile_kernel(DwView<const T> in, const T* __restrict__ w9c, DwView<T> out, int C,
                                                                  DwMap m, int flip, int accumulate, const __grid_constant__ CUtensorMap in_map) {
  constexpr int VP = V / 2;
  extern __shared__ __align__(128) uint4 dw_tile[];   // [rs + 2][ppb + 2][cvp]
  __shared__ __align__(8) uint64_t tma_bar;
  const DwLane l = dw_lane(m, out.h, out.w);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int TW = m.ppb + 2;
  const int nrows = (l.y1 - l.y0) + 2;                // input rows y0-1 .. y1
  const int x_base = blockIdx.x * m.ppb - 1, y_base = l.y0 - 1;
  const int cv0 = blockIdx.y * m.cvp;
  const int H = in.h, W = in.w;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(dw_tile);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&tma_bar);
  if (kTma) {
    if (threadIdx.x == 0) dw_mbar_init(bar_s, 1);
    __syncthreads();
  }
  pdl_sync();
  if (kTma) {
    // one box = the whole tile, halo included; rows / pixels / channels outside the tensor arrive as zeros
    if (threadIdx.x == 0) {
      dw_mbar_expect_tx(bar_s, (uint32_t)((m.rs + 2) * TW * m.cvp * 16));
      dw_tma_load_tile(&in_map, bar_s, tile_s, cv0 * V, x_base, y_base, l.n, in.dsub, in.h);
    }
  } else {
    const int nvec = nrows * TW * m.cvp;
    const int cshift = 31 - __clz(m.cvp);
    const T* nbase = in.p + in.img(l.n);
    // (ty, tx) of this thread's vectors advance by a fixed pixel step < TW per iteration: tracked incrementally, the loop
    // carries no integer division (it used to be a quarter of the kernel's instructions)
    const int cl = threadIdx.x & (m.cvp - 1), cvi = cv0 + cl;
    const int pstep = kDwThreads >> cshift;
    int ty = (threadIdx.x >> cshift) / TW, tx = (threadIdx.x >> cshift) - ty * TW;
    for (int i = threadIdx.x; i < nvec; i += kDwThreads) {
      const int gy = y_base + ty, gx = x_base + tx;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W && cvi < m.cv;
      const T* src = ok ? nbase + (long long)gy * in.sh + (long long)gx * in.sw + cvi * V : in.p;
      cp_async16_zfill(tile_s + (uint32_t)i * 16u, src, ok);
      tx += pstep;
      if (tx >= TW) { tx -= TW; ++ty; }
    }
  }
  float2 wv[9][VP];
  if (l.ok) {
    const int c0 = l.cvi * V;
#pragma unroll
    for (int k = 0; k < 9; ++k) dwpair<T>::unpack(ld16(w9c + (size_t)(flip ? 8 - k : k) * C + c0), wv[k]);
  }
  if (kTma) {
    dw_mbar_wait(bar_s, 0);
  } else {
    cp_async_commit_wait_all();
    __syncthreads();
  }
  if (!l.ok) return;
  T* obase = out.p + out.img(l.n) + (long long)l.x * out.sw + l.cvi * V;
  const uint4* tp = dw_tile + ((warp * m.ppw + psub) * m.cvp + cvl);      // tile column of x-1, row 0
  const int rstride = TW * m.cvp;
  auto step = [&](int ty, float2 (&A)[VP], float2 (&B)[VP], float2 (&Cn)[VP]) {
    const uint4* rp = tp + ty * rstride;
    float2 f[3][VP];
    dwpair<T>::unpack(rp[0], f[0]);
    dwpair<T>::unpack(rp[m.cvp], f[1]);
    dwpair<T>::unpack(rp[2 * m.cvp], f[2]);
#pragma unroll
    for (int j = 0; j < VP; ++j) {
      Cn[j] = mul2(f[0][j], wv[0][j]);
      Cn[j] = fma2(f[1][j], wv[1][j], Cn[j]);
      Cn[j] = fma2(f[2][j], wv[2][j], Cn[j]);
      B[j] = fma2(f[0][j], wv[3][j], B[j]);
      B[j] = fma2(f[1][j], wv[4][j], B[j]);
      B[j] = fma2(f[2][j], wv[5][j], B[j]);
      A[j] = fma2(f[0][j], wv[6][j], A[j]);
      A[j] = fma2(f[1][j], wv[7][j], A[j]);
      A[j] = fma2(f[2][j], wv[8][j], A[j]);
    }
    const int y = y_base + ty - 1;                     // output row completed by input row y_base + ty
    if (y >= l.y0 && y < l.y1) {
      T* op = obase + (long long)y * out.sh;
      if (accumulate) {
        float2 old[VP];
        dwpair<T>::unpack(ld16(op), old);
#pragma unroll
        for (int j = 0; j < VP; ++j) { A[j].x += old[j].x; A[j].y += old[j].y; }
      }
      st16(op, dwpair<T>::pack(A));
    }
  };
  float2 a0[VP], a1[VP], a2[VP];
#pragma unroll
  for (int j = 0; j < VP; ++j) { a0[j] = make_float2(0.f, 0.f); a1[j] = a0[j]; a2[j] = a0[j]; }
  int ty = 0;
  while (true) {
    step(ty, a0, a1, a2);
    if (++ty >= nrows) break;
    step(ty, a1, a2, a0);
    if (++ty >= nrows) break;
    step(ty, a2, a0, a1);
    if (++ty >= nrows) break;
  }
}

// ---- stride 1, dilation 1 forward with the PRECEDING BatchNorm (+ReLU) applied on load ------------------------------
// The ReLU -> SeparableConv2d_same chain of every Block (DX:79-97) reads a = relu(bn(y)), where y is the previous pointwise
// convolution's output whose batch sums came out of the GEMM epilogue (DC_BN_SUMS_READY).  Instead of one bn_apply launch
// that writes a and one depthwise launch that reads it back, this kernel stages the raw y tile (halo included) in shared
// memory, finalizes the block's <= 256 channels from the sums exactly as bn_apply_kernel does (same arithmetic, so the
// activation is bit-identical; block (0, y, 0) publishes the coefficients for the backward pass and updates the running
// statistics), applies scale/shift/ReLU in place - the zero padding of fixed_padding stays zero, it pads a, not y - writes
// the interior of a (backward needs it: depthwise weight gradient, residual-free mask recompute does not) and then runs the
// same input-stationary row walk as dw_s1d1_tile_kernel.
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) dw_s1d1_tile_bn_kernel(dc_bn_params p, DwView<const T> in, const T* __restrict__ w9c,
                                                                     DwView<T> act, DwView<T> out, int C, DwMap m) {
  constexpr int VP = V / 2;
  extern __shared__ uint4 dw_tile[];                  // [rs + 2][ppb + 2][cvp] | float scale[cvp * V] | float shift[cvp * V]
  const DwLane l = dw_lane(m, out.h, out.w);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int TW = m.ppb + 2;
  const int nrows = (l.y1 - l.y0) + 2;                // input rows y0-1 .. y1
  const int x_base = blockIdx.x * m.ppb - 1, y_base = l.y0 - 1;
  const int cv0 = blockIdx.y * m.cvp;
  const int H = in.h, W = in.w;
  const int nvec_max = (m.rs + 2) * TW * m.cvp;
  float* s_scale = reinterpret_cast<float*>(dw_tile + nvec_max);
  float* s_shift = s_scale + m.cvp * V;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(dw_tile);
  const int nvec = nrows * TW * m.cvp;
  const int cshift = 31 - __clz(m.cvp);
  pdl_sync();
  {
    const T* nbase = in.p + in.img(l.n);
    // (ty, tx) of this thread's vectors advance by a fixed pixel step < TW per iteration: tracked incrementally, the loop
    // carries no integer division (it used to be a quarter of the kernel's instructions)
    const int cl = threadIdx.x & (m.cvp - 1), cvi = cv0 + cl;
    const int pstep = kDwThreads >> cshift;
    int ty = (threadIdx.x >> cshift) / TW, tx = (threadIdx.x >> cshift) - ty * TW;
    for (int i = threadIdx.x; i < nvec; i += kDwThreads) {
      const int gy = y_base + ty, gx = x_base + tx;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W && cvi < m.cv;
      const T* src = ok ? nbase + (long long)gy * in.sh + (long long)gx * in.sw + cvi * V : in.p;
      cp_async16_zfill(tile_s + (uint32_t)i * 16u, src, ok);
      tx += pstep;
      if (tx >= TW) { tx -= TW; ++ty; }
    }
  }
  {
    // per-channel coefficients of this block's channel group while the tile is in flight (thread t -> one channel)
    const int nch = m.cvp * V;
    if ((int)threadIdx.x < nch) {
      const int c = cv0 * V + threadIdx.x;
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
        if (blockIdx.x == 0 && blockIdx.z == 0) {
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
      s_scale[threadIdx.x] = sc;
      s_shift[threadIdx.x] = sh;
    }
  }
  float2 wv[9][VP];
  if (l.ok) {
    const int c0 = l.cvi * V;
#pragma unroll
    for (int k = 0; k < 9; ++k) dwpair<T>::unpack(ld16(w9c + (size_t)k * C + c0), wv[k]);
  }
  cp_async_commit_wait_all();
  __syncthreads();
  {
    // a = [relu](y * scale + shift) in place; pixels outside the image stay zero (they pad a); interior pixels are stored
    const bool relu = (p.flags & DC_BN_RELU) != 0;
    T* abase = act.p ? act.p + act.img(l.n) : nullptr;
    const int cl = threadIdx.x & (m.cvp - 1), cvi = cv0 + cl;
    const int pstep = kDwThreads >> cshift;
    int ty = (threadIdx.x >> cshift) / TW, tx = (threadIdx.x >> cshift) - ty * TW;
    float2 sc2[VP], sh2[VP];
#pragma unroll
    for (int j = 0; j < VP; ++j) {
      sc2[j] = make_float2(s_scale[cl * V + 2 * j], s_scale[cl * V + 2 * j + 1]);
      sh2[j] = make_float2(s_shift[cl * V + 2 * j], s_shift[cl * V + 2 * j + 1]);
    }
    for (int i = threadIdx.x; i < nvec; i += kDwThreads, tx += pstep) {
      if (tx >= TW) { tx -= TW; ++ty; }
      const int gy = y_base + ty, gx = x_base + tx;
      if (!(gy >= 0 && gy < H && gx >= 0 && gx < W && cvi < m.cv)) continue;
      float2 f[VP];
      dwpair<T>::unpack(dw_tile[i], f);
#pragma unroll
      for (int j = 0; j < VP; ++j) {
        f[j] = fma2(f[j], sc2[j], sh2[j]);
        if (relu) { f[j].x = fmaxf(f[j].x, 0.f); f[j].y = fmaxf(f[j].y, 0.f); }
      }
      const uint4 o = dwpair<T>::pack(f);
      dw_tile[i] = o;
      if (abase != nullptr && ty >= 1 && ty <= nrows - 2 && tx >= 1 && tx <= m.ppb)
        st16(abase + (long long)gy * act.sh + (long long)gx * act.sw + cvi * V, o);
    }
  }
  __syncthreads();
  if (!l.ok) return;
  T* obase = out.p + out.img(l.n) + (long long)l.x * out.sw + l.cvi * V;
  const uint4* tp = dw_tile + ((warp * m.ppw + psub) * m.cvp + cvl);      // tile column of x-1, row 0
  const int rstride = TW * m.cvp;
  auto step = [&](int ty, float2 (&A)[VP], float2 (&B)[VP], float2 (&Cn)[VP]) {
    const uint4* rp = tp + ty * rstride;
    float2 f[3][VP];
    dwpair<T>::unpack(rp[0], f[0]);
    dwpair<T>::unpack(rp[m.cvp], f[1]);
    dwpair<T>::unpack(rp[2 * m.cvp], f[2]);
#pragma unroll
    for (int j = 0; j < VP; ++j) {
      Cn[j] = mul2(f[0][j], wv[0][j]);
      Cn[j] = fma2(f[1][j], wv[1][j], Cn[j]);
      Cn[j] = fma2(f[2][j], wv[2][j], Cn[j]);
      B[j] = fma2(f[0][j], wv[3][j], B[j]);
      B[j] = fma2(f[1][j], wv[4][j], B[j]);
      B[j] = fma2(f[2][j], wv[5][j], B[j]);
      A[j] = fma2(f[0][j], wv[6][j], A[j]);
      A[j] = fma2(f[1][j], wv[7][j], A[j]);
      A[j] = fma2(f[2][j], wv[8][j], A[j]);
    }
    const int y = y_base + ty - 1;                     // output row completed by input row y_base + ty
    if (y >= l.y0 && y < l.y1) st16(obase + (long long)y * out.sh, dwpair<T>::pack(A));
  };
  float2 a0[VP], a1[VP], a2[VP];
#pragma unroll
  for (int j = 0; j < VP; ++j) { a0[j] = make_float2(0.f, 0.f); a1[j] = a0[j]; a2[j] = a0[j]; }
  int ty = 0;
  while (true) {
    step(ty, a0, a1, a2);
    if (++ty >= nrows) break;
    step(ty, a1, a2, a0);
    if (++ty >= nrows) break;
    step(ty, a2, a0, a1);
    if (++ty >= nrows) break;
  }
}

// ---- stride 1, dilation 1 backward-data FUSED with the ReLU mask and the BatchNorm backward reduction of the layer below ---
// The depthwise conv's input a = relu(bn(y) [+ res]) is a BatchNorm output.  This kernel is the last writer of dL/da: it
// finishes the gradient (adds what other consumers already stored when `accumulate`), rounds it to the storage type, masks
// it with a > 0 and stores it, and accumulates per channel  sum(g)  and  sum(g * y)  into the BatchNorm backward workspace.
// BatchNorm backward is then one element-wise launch (dc_bn_bwd_apply_reduced) instead of reduce + apply or the one-pass
// barrier kernel, and neither `a` nor the gradient is re-read for the reduction.
//   HAS_X = false: no residual, the mask is recomputed from y and the forward coefficients (like DC_BN_MASK_FROM_Y);
//   HAS_X = true : the stored activation a is the mask source (residual blocks); the previously stored gradient is
//                  staged through shared memory as well.
// Weights stay packed in registers (unpacked per use) so that the kernel fits 2 blocks per SM next to the extra state.
template <typename T, int V, bool HAS_X>
__global__ void __launch_bounds__(kDwThreads, 2) dw_s1d1_tile_bnred_kernel(DwView<const T> in, const T* __restrict__ w9c, DwView<T> out,
                                                                           int C, DwMap m, int accumulate, DwView<const T> yv,
                                                                           DwView<const T> xv, const float* __restrict__ fcoef,
                                                                           double* __restrict__ sums, int relu) {
  constexpr int VP = V / 2;
  extern __shared__ uint4 dw_tile[];                  // [rs + 2][ppb + 2][cvp] | y [rs][ppb][cvp] | (HAS_X) x, old [rs][ppb][cvp]
  const DwLane l = dw_lane(m, out.h, out.w);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int TW = m.ppb + 2;
  const int nrows = (l.y1 - l.y0) + 2;
  const int x_base = blockIdx.x * m.ppb - 1, y_base = l.y0 - 1;
  const int cv0 = blockIdx.y * m.cvp;
  const int H = in.h, W = in.w;
  const int in_vecs = (m.rs + 2) * TW * m.cvp;
  const int int_vecs = m.rs * m.ppb * m.cvp;          // interior tile
  uint4* ty_s = dw_tile + in_vecs;
  uint4* tx_s = ty_s + int_vecs;
  uint4* to_s = tx_s + int_vecs;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(dw_tile);
  const int cshift = 31 - __clz(m.cvp);
  pdl_sync();
  {
    const T* nbase = in.p + in.img(l.n);
    for (int i = threadIdx.x; i < nrows * TW * m.cvp; i += kDwThreads) {
      const int cl = i & (m.cvp - 1);
      const int pix = i >> cshift;
      const int ty = pix / TW, tx = pix - ty * TW;
      const int gy = y_base + ty, gx = x_base + tx, cvi = cv0 + cl;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W && cvi < m.cv;
      const T* src = ok ? nbase + (long long)gy * in.sh + (long long)gx * in.sw + cvi * V : in.p;
      cp_async16_zfill(tile_s + (uint32_t)i * 16u, src, ok);
    }
    const int irows = l.y1 - l.y0;
    const uint32_t ty_a = (uint32_t)__cvta_generic_to_shared(ty_s), tx_a = (uint32_t)__cvta_generic_to_shared(tx_s);
    const uint32_t to_a = (uint32_t)__cvta_generic_to_shared(to_s);
    for (int i = threadIdx.x; i < irows * m.ppb * m.cvp; i += kDwThreads) {
      const int cl = i & (m.cvp - 1);
      const int pix = i >> cshift;
      const int ty = pix / m.ppb, tx = pix - ty * m.ppb;
      const int gy = l.y0 + ty, gx = blockIdx.x * m.ppb + tx, cvi = cv0 + cl;
      const bool ok = gx < out.w && cvi < m.cv;
      const long long off = (long long)gy * yv.sh + (long long)gx * yv.sw + cvi * V;
      cp_async16_zfill(ty_a + (uint32_t)i * 16u, ok ? yv.p + yv.img(l.n) + off : yv.p, ok);
      if (HAS_X) {
        cp_async16_zfill(tx_a + (uint32_t)i * 16u,
                         ok ? xv.p + xv.img(l.n) + (long long)gy * xv.sh + (long long)gx * xv.sw + cvi * V : xv.p, ok);
        if (accumulate)
          cp_async16_zfill(to_a + (uint32_t)i * 16u,
                           ok ? out.p + out.img(l.n) + (long long)gy * out.sh + (long long)gx * out.sw + cvi * V : out.p, ok);
      }
    }
  }
  uint4 wraw[9];
  // forward scale / shift of this block's channels (mask recomputation) live in shared memory behind the tiles: [2][cvp * V]
  float* fco_s = reinterpret_cast<float*>(HAS_X ? to_s + int_vecs : tx_s);
  if (!HAS_X && relu) {
    for (int i = threadIdx.x; i < 2 * m.cvp * V; i += kDwThreads) {
      const int which = i / (m.cvp * V), c = cv0 * V + (i - which * m.cvp * V);
      fco_s[i] = c < C ? fcoef[which * C + c] : 0.f;
    }
  }
  if (l.ok) {
    const int c0 = l.cvi * V;
#pragma unroll
    for (int k = 0; k < 9; ++k) wraw[k] = ld16(w9c + (size_t)(8 - k) * C + c0);          // flipped filter
  }
  cp_async_commit_wait_all();
  __syncthreads();
  float acc[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  if (l.ok) {
    T* obase = out.p + out.img(l.n) + (long long)l.x * out.sw + l.cvi * V;
    const int pcol = warp * m.ppw + psub;                                   // pixel column inside the tile (interior index)
    const uint4* tp = dw_tile + (pcol * m.cvp + cvl);                       // halo tile column of x-1, row 0
    const int rstride = TW * m.cvp;
    auto step = [&](int ty, float2 (&A)[VP], float2 (&B)[VP], float2 (&Cn)[VP]) {
      const uint4* rp = tp + ty * rstride;
      float2 f[3][VP];
      dwpair<T>::unpack(rp[0], f[0]);
      dwpair<T>::unpack(rp[m.cvp], f[1]);
      dwpair<T>::unpack(rp[2 * m.cvp], f[2]);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float2 w0[VP], w1[VP], w2[VP];
        dwpair<T>::unpack(wraw[kw], w0);
        dwpair<T>::unpack(wraw[3 + kw], w1);
        dwpair<T>::unpack(wraw[6 + kw], w2);
#pragma unroll
        for (int j = 0; j < VP; ++j) {
          Cn[j] = kw == 0 ? mul2(f[0][j], w0[j]) : fma2(f[kw][j], w0[j], Cn[j]);
          B[j] = fma2(f[kw][j], w1[j], B[j]);
          A[j] = fma2(f[kw][j], w2[j], A[j]);
        }
      }
      const int y = y_base + ty - 1;                     // output row completed by input row y_base + ty
      if (y >= l.y0 && y < l.y1) {
        const int ii = ((y - l.y0) * m.ppb + pcol) * m.cvp + cvl;
        if (HAS_X && accumulate) {
          float2 old[VP];
          dwpair<T>::unpack(to_s[ii], old);
#pragma unroll
          for (int j = 0; j < VP; ++j) { A[j].x += old[j].x; A[j].y += old[j].y; }
        } else if (accumulate) {
          float2 old[VP];
          dwpair<T>::unpack(ld16(obase + (long long)y * out.sh), old);
#pragma unroll
          for (int j = 0; j < VP; ++j) { A[j].x += old[j].x; A[j].y += old[j].y; }
        }
        float2 g[VP], yy[VP];
        dwpair<T>::unpack(dwpair<T>::pack(A), g);        // the gradient as it is stored (rounded to T)
        dwpair<T>::unpack(ty_s[ii], yy);
        if (relu) {
          if (HAS_X) {
            float2 xx[VP];
            dwpair<T>::unpack(tx_s[ii], xx);
#pragma unroll
            for (int j = 0; j < VP; ++j) { g[j].x = xx[j].x > 0.f ? g[j].x : 0.f; g[j].y = xx[j].y > 0.f ? g[j].y : 0.f; }
          } else {
            const float2* sc2 = reinterpret_cast<const float2*>(fco_s + cvl * V);
            const float2* sh2 = reinterpret_cast<const float2*>(fco_s + m.cvp * V + cvl * V);
#pragma unroll
            for (int j = 0; j < VP; ++j) {
              const float2 o = fma2(yy[j], sc2[j], sh2[j]);
              g[j].x = o.x > 0.f ? g[j].x : 0.f;
              g[j].y = o.y > 0.f ? g[j].y : 0.f;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < VP; ++j) {
          acc[0][2 * j] += g[j].x; acc[0][2 * j + 1] += g[j].y;
          acc[1][2 * j] = fmaf(g[j].x, yy[j].x, acc[1][2 * j]);
          acc[1][2 * j + 1] = fmaf(g[j].y, yy[j].y, acc[1][2 * j + 1]);
        }
        st16(obase + (long long)y * out.sh, dwpair<T>::pack(g));
      }
    };
    float2 a0[VP], a1[VP], a2[VP];
#pragma unroll
    for (int j = 0; j < VP; ++j) { a0[j] = make_float2(0.f, 0.f); a1[j] = a0[j]; a2[j] = a0[j]; }
    int ty = 0;
    while (true) {
      step(ty, a0, a1, a2);
      if (++ty >= nrows) break;
      step(ty, a1, a2, a0);
      if (++ty >= nrows) break;
      step(ty, a2, a0, a1);
      if (++ty >= nrows) break;
    }
  }
  __syncthreads();                                       // the tiles are dead: their memory becomes the reduction scratch
  reduce_to_ws<2, V>(acc, sums, C, m.cvp, cv0, min(m.cvp, m.cv - cv0));
}

// ---- generic stride / dilation forward (and stride-1 backward-data with flip): nine direct taps ----------------------
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) dw_direct_kernel(DwView<const T> in, const T* __restrict__ w9c, DwView<T> out, int C,
                                                               DwMap m, int s, int d, int flip, int accumulate) {
  const DwLane l = dw_lane(m, out.h, out.w);
  pdl_sync();
  if (!l.ok) return;
  const int c0 = l.cvi * V;
  float wv[9][V];
  load_weights<T, V>(w9c, C, c0, flip != 0, wv);
  const T* base = in.p + in.img(l.n) + c0;
  T* obase = out.p + out.img(l.n) + (long long)l.x * out.sw + c0;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  int ix[3];
  bool vx[3];
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) { ix[kw] = l.x * s - d + kw * d; vx[kw] = ix[kw] >= 0 && ix[kw] < in.w; }
  for (int y = l.y0; y < l.y1; ++y) {
    uint4 t[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = y * s - d + kh * d;
      const bool vy = ih >= 0 && ih < in.h;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
        t[kh * 3 + kw] = (vy && vx[kw]) ? ld16(base + (long long)ih * in.sh + (long long)ix[kw] * in.sw) : zero;
    }
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float f[V];
      dwvec<T>::unpack(t[k], f);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = fmaf(f[j], wv[k][j], acc[j]);
    }
    T* op = obase + (long long)y * out.sh;
    if (accumulate) {
      float old[V];
      dwvec<T>::unpack(ld16(op), old);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += old[j];
    }
    st16(op, dwvec<T>::pack(acc));
  }
}

// ---- backward-data of a stride-2 layer: gather over the taps whose parity matches ----------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) dw_bwd_data_strided_kernel(DwView<const T> dout, const T* __restrict__ w9c, DwView<T> din,
                                                                         int C, DwMap m, int s, int d, int accumulate) {
  const DwLane l = dw_lane(m, din.h, din.w);
  pdl_sync();
  if (!l.ok) return;
  const int c0 = l.cvi * V;
  float wv[9][V];
  load_weights<T, V>(w9c, C, c0, false, wv);
  const T* base = dout.p + dout.img(l.n) + c0;
  T* obase = din.p + din.img(l.n) + (long long)l.x * din.sw + c0;
  int ox[3];
  bool vx[3];
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
    const int tx = l.x + d - kw * d;
    ox[kw] = tx / s;
    vx[kw] = tx >= 0 && (tx % s) == 0 && ox[kw] < dout.w;
  }
  for (int y = l.y0; y < l.y1; ++y) {
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ty = y + d - kh * d;
      if (ty < 0 || (ty % s) != 0) continue;
      const int oy = ty / s;
      if (oy >= dout.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        if (!vx[kw]) continue;
        float f[V];
        dwvec<T>::unpack(ld16(base + (long long)oy * dout.sh + (long long)ox[kw] * dout.sw), f);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = fmaf(f[j], wv[kh * 3 + kw][j], acc[j]);
      }
    }
    T* op = obase + (long long)y * din.sh;
    if (accumulate) {
      float old[V];
      dwvec<T>::unpack(ld16(op), old);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += old[j];
    }
    st16(op, dwvec<T>::pack(acc));
  }
}

// ---- weight gradient -------------------------------------------------------------------------------------------------
// Per-thread 9 x V fp32 partial sums over its column strip; reduction across the pixel lanes of the block in three passes
// (one filter row each) through shared memory, then one fp32 atomicAdd per (tap, channel) per block.
// Gout[(kh*3+kw) * tap_stride + c * c_stride]: tap-major scratch (tap_stride = C, c_stride = 1) or the parameter's own
// [C][1][3][3] layout (tap_stride = 1, c_stride = 9), which needs no unpack launch afterwards.
__device__ int g_dww_noatomic = 0;      // experiment knob (DEEPCAM_B200_DWW_NOATOMIC=1, results are garbage): time without the atomics
template <int V>
__device__ __forceinline__ void dw_reduce_G(float (&G)[9][V], float* __restrict__ Gout, int C, const DwMap& m, int cvi_base,
                                            int tap_stride, int c_stride, int det) {
  // det (dc_dw_bwd_weight_det): Gout is a workspace [slices][9][C]; this block STORES its sums into slice (blockIdx.x, blockIdx.z) -
  // one writer per element - and dww_slice_reduce_kernel adds the slices in order
  if (det) Gout += ((size_t)blockIdx.x + (size_t)gridDim.x * blockIdx.z) * 9 * C;
  extern __shared__ float red[];                       // [8 warps][32 lanes][3*V]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = m.cvp; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) G[k][j] += __shfl_xor_sync(0xffffffffu, G[k][j], o);
  }
  constexpr int PER = 3 * V;
  const int cv_count = min(m.cvp, m.cv - cvi_base);
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    if (kh) __syncthreads();
    if (lane < m.cvp) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int j = 0; j < V; ++j) red[(warp * 32 + lane) * PER + kw * V + j] = G[kh * 3 + kw][j];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < m.cvp * PER; col += kDwThreads) {
      const int ln = col / PER, r = col - ln * PER;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[(w * 32 + ln) * PER + r];
      const int kw = r / V, j = r - kw * V;
      if (ln < cv_count && !g_dww_noatomic) {
        float* dst = Gout + (size_t)(kh * 3 + kw) * tap_stride + (size_t)((cvi_base + ln) * V + j) * c_stride;
        if (det) *dst = s; else atomicAdd(dst, s);
      }
    }
  }
}

// input-stationary like dw_s1d1_kernel: input row r meets dout rows r+1 (filter row 0), r (row 1), r-1 (row 2)
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) dw_bwd_weight_s1d1_kernel(DwView<const T> in, DwView<const T> dout, float* __restrict__ Gout,
                                                                        int C, DwMap m, int tap_stride, int c_stride, int det) {
  constexpr int VP = V / 2;
  const DwLane l = dw_lane(m, dout.h, dout.w);
  pdl_sync();
  float2 G2[9][VP];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < VP; ++j) G2[k][j] = make_float2(0.f, 0.f);
  if (l.ok) {
    const int c0 = l.cvi * V;
    const int H = in.h, W = in.w;
    const T* base = in.p + in.img(l.n) + (long long)l.x * in.sw + c0;
    const T* gbase = dout.p + dout.img(l.n) + (long long)l.x * dout.sw + c0;
    const bool xl = l.x >= 1, xr = l.x + 1 < W;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    auto ldrow = [&](int r, uint4 (&v)[4]) {          // v[0..2]: input row r at x-1, x, x+1; v[3]: dout row r+1
      if (r < 0 || r >= H) { v[0] = zero; v[1] = zero; v[2] = zero; }
      else {
        const T* rp = base + (long long)r * in.sh;
        v[1] = ld16(rp);
        v[0] = xl ? ld16(rp - in.sw) : zero;
        v[2] = xr ? ld16(rp + in.sw) : zero;
      }
      const int y = r + 1;
      v[3] = (y >= l.y0 && y < l.y1) ? ld16(gbase + (long long)y * dout.sh) : zero;
    };
    // gM = dout row r-1, gC = dout row r, gP = dout row r+1 (unpacked here)
    auto step = [&](const uint4 (&v)[4], const float2 (&gM)[VP], const float2 (&gC)[VP], float2 (&gP)[VP]) {
      float2 f[3][VP];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) dwpair<T>::unpack(v[kw], f[kw]);
      dwpair<T>::unpack(v[3], gP);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int j = 0; j < VP; ++j) {
          G2[kw][j] = fma2(f[kw][j], gP[j], G2[kw][j]);
          G2[3 + kw][j] = fma2(f[kw][j], gC[j], G2[3 + kw][j]);
          G2[6 + kw][j] = fma2(f[kw][j], gM[j], G2[6 + kw][j]);
        }
    };
    float2 g0[VP], g1[VP], g2[VP];
#pragma unroll
    for (int j = 0; j < VP; ++j) { g0[j] = make_float2(0.f, 0.f); g1[j] = g0[j]; g2[j] = g0[j]; }
    uint4 cur[4], nxt[4];
    int r = l.y0 - 1;
    ldrow(r, cur);
    while (true) {
      ldrow(r + 1, nxt);
      step(cur, g0, g1, g2);
      if (++r > l.y1) break;
      ldrow(r + 1, cur);
      step(nxt, g1, g2, g0);
      if (++r > l.y1) break;
      ldrow(r + 1, nxt);
      step(cur, g2, g0, g1);
      if (++r > l.y1) break;
      ldrow(r + 1, cur);
      step(nxt, g0, g1, g2);
      if (++r > l.y1) break;
      ldrow(r + 1, nxt);
      step(cur, g1, g2, g0);
      if (++r > l.y1) break;
      ldrow(r + 1, cur);
      step(nxt, g2, g0, g1);
      if (++r > l.y1) break;
    }
  }
  float G[9][V];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < VP; ++j) { G[k][2 * j] = G2[k][j].x; G[k][2 * j + 1] = G2[k][j].y; }
  dw_reduce_G<V>(G, Gout, C, m, blockIdx.y * m.cvp, tap_stride, c_stride, det);
}

// ---- stride 2, dilation 1 backward-data, staged (the last separable unit of blocks 1-3, DX:99-101) ------------------------------
// out[y,x] = sum in[2y-1+kh, 2x-1+kw] * w[kh][kw], so every dout pixel (y, x) owns the 2 x 2 input quad (2y.., 2x..):
//   din[2y  , 2x  ] = d[y,x] w11
//   din[2y  , 2x+1] = d[y,x] w12 + d[y,x+1] w10
//   din[2y+1, 2x  ] = d[y,x] w21 + d[y+1,x] w01
//   din[2y+1, 2x+1] = d[y,x] w22 + d[y,x+1] w20 + d[y+1,x] w02 + d[y+1,x+1] w00
// A block stages its dout tile (one halo row below, one halo column right, zero outside) with cp.async copies that are all in
// flight at once; a thread (channel vector, dout column) walks down the rows keeping d[y][x], d[y][x+1] in registers and writes
// the quad's four 16-byte vectors.  The gather-form dw_bwd_data_strided_kernel it replaces ran at 1.1 TB/s on the 113 MB
// gradient of block1 (124 us): per input pixel it tested nine taps for parity and issued dependent global loads.
template <typename T, int V, bool kTma>
__global__ void __launch_bounds__(kDwThreads, 2) dw_bwd_data_s2_tile_kernel(DwView<const T> dout, const T* __restrict__ w9c, DwView<T> din,
                                                                           int C, DwMap m, int accumulate,
                                                                           const __grid_constant__ CUtensorMap g_map) {
  constexpr int VP = V / 2;
  extern __shared__ __align__(128) uint4 dw_tile[];   // [rs + 1][ppb + 1][cvp]
  __shared__ __align__(8) uint64_t tma_bar;
  const DwLane l = dw_lane(m, dout.h, dout.w);        // rows / columns of dout
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int TW = m.ppb + 1;
  const int nrows = (l.y1 - l.y0) + 1;                // dout rows y0 .. y1 (y1 = halo)
  const int x0 = blockIdx.x * m.ppb;
  const int cv0 = blockIdx.y * m.cvp;
  const int Ho = dout.h, Wo = dout.w;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(dw_tile);
  const int cshift = 31 - __clz(m.cvp);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&tma_bar);
  if (kTma) {
    if (threadIdx.x == 0) dw_mbar_init(bar_s, 1);
    __syncthreads();
  }
  pdl_sync();
  if (kTma) {
    if (threadIdx.x == 0) {
      dw_mbar_expect_tx(bar_s, (uint32_t)((m.rs + 1) * TW * m.cvp * 16));
      dw_tma_load_4d(&g_map, bar_s, tile_s, cv0 * V, x0, l.y0, l.n);
    }
  } else {
    const T* nbase = dout.p + dout.img(l.n);
    const int nvec = nrows * TW * m.cvp;
    const int cl = threadIdx.x & (m.cvp - 1), cvi = cv0 + cl;
    const int pstep = kDwThreads >> cshift;
    int ty = (threadIdx.x >> cshift) / TW, tx = (threadIdx.x >> cshift) - ty * TW;
    for (int i = threadIdx.x; i < nvec; i += kDwThreads) {
      const int gy = l.y0 + ty, gx = x0 + tx;
      const bool ok = gy < Ho && gx < Wo && cvi < m.cv;
      cp_async16_zfill(tile_s + (uint32_t)i * 16u, ok ? nbase + (long long)gy * dout.sh + (long long)gx * dout.sw + cvi * V : dout.p, ok);
      tx += pstep;
      while (tx >= TW) { tx -= TW; ++ty; }
    }
  }
  float2 wv[9][VP];
  if (l.ok) {
    const int c0 = l.cvi * V;
#pragma unroll
    for (int k = 0; k < 9; ++k) dwpair<T>::unpack(ld16(w9c + (size_t)k * C + c0), wv[k]);
  }
  if (kTma) {
    dw_mbar_wait(bar_s, 0);
  } else {
    cp_async_commit_wait_all();
    __syncthreads();
  }
  if (!l.ok) return;
  const int H = din.h, W = din.w;
  const int ix = 2 * l.x;
  T* ibase = din.p + din.img(l.n) + (long long)ix * din.sw + l.cvi * V;
  const bool x1ok = ix + 1 < W;
  const uint4* tp = dw_tile + ((warp * m.ppw + psub) * m.cvp + cvl);      // tile column of x, row 0
  const int rstride = TW * m.cvp;
  float2 a[VP], b[VP];                                // d[y][x], d[y][x+1]
  dwpair<T>::unpack(tp[0], a);
  dwpair<T>::unpack(tp[m.cvp], b);
  auto put = [&](T* dst, float2 (&v)[VP]) {
    if (accumulate) {
      float2 old[VP];
      dwpair<T>::unpack(ld16(dst), old);
#pragma unroll
      for (int j = 0; j < VP; ++j) { v[j].x += old[j].x; v[j].y += old[j].y; }
    }
    st16(dst, dwpair<T>::pack(v));
  };
  for (int ty = 0; ty + 1 < nrows; ++ty) {
    float2 c[VP], d[VP];                              // d[y+1][x], d[y+1][x+1]
    dwpair<T>::unpack(tp[(ty + 1) * rstride], c);
    dwpair<T>::unpack(tp[(ty + 1) * rstride + m.cvp], d);
    const int iy = 2 * (l.y0 + ty);
    T* r0 = ibase + (long long)iy * din.sh;
    float2 o[VP];
#pragma unroll
    for (int j = 0; j < VP; ++j) o[j] = mul2(a[j], wv[4][j]);
    put(r0, o);
    if (x1ok) {
#pragma unroll
      for (int j = 0; j < VP; ++j) o[j] = fma2(b[j], wv[3][j], mul2(a[j], wv[5][j]));
      put(r0 + din.sw, o);
    }
    if (iy + 1 < H) {
      T* r1 = r0 + din.sh;
#pragma unroll
      for (int j = 0; j < VP; ++j) o[j] = fma2(c[j], wv[1][j], mul2(a[j], wv[7][j]));
      put(r1, o);
      if (x1ok) {
#pragma unroll
        for (int j = 0; j < VP; ++j) o[j] = fma2(d[j], wv[0][j], fma2(c[j], wv[2][j], fma2(b[j], wv[6][j], mul2(a[j], wv[8][j]))));
        put(r1 + din.sw, o);
      }
    }
#pragma unroll
    for (int j = 0; j < VP; ++j) { a[j] = c[j]; b[j] = d[j]; }
  }
}

// ---- stride 2 forward, staged (TMA) ---------------------------------------------------------------------------------------
// The three stride-2 layers of the entry flow (DX:151-159, block1-3 rep[-1]) used the gather kernel: nine global loads per output
// vector, 61 us per step for ~40 us of HBM time.  Here one TMA box brings the (2 rs + 1) x (2 ppb + 1) input pixels of a tile of
// rs x ppb outputs into shared memory (zero fill outside the image = fixed_padding) and every thread walks its output column down
// the strip, nine 16-byte shared-memory loads per output.
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads, 2) dw_fwd_s2_tile_kernel(const T* __restrict__ w9c, DwView<T> out, int C, DwMap m,
                                                                       const __grid_constant__ CUtensorMap in_map) {
  constexpr int VP = V / 2;
  extern __shared__ __align__(128) uint4 dw_tile[];   // [2 rs + 1][2 ppb + 1][cvp]
  __shared__ __align__(8) uint64_t tma_bar;
  const DwLane l = dw_lane(m, out.h, out.w);          // rows / columns of out
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  const int TW = 2 * m.ppb + 1;
  const int cv0 = blockIdx.y * m.cvp;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(dw_tile);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&tma_bar);
  if (threadIdx.x == 0) dw_mbar_init(bar_s, 1);
  __syncthreads();
  pdl_sync();
  if (threadIdx.x == 0) {
    dw_mbar_expect_tx(bar_s, (uint32_t)((2 * m.rs + 1) * TW * m.cvp * 16));
    dw_tma_load_4d(&in_map, bar_s, tile_s, cv0 * V, 2 * (int)blockIdx.x * m.ppb - 1, 2 * l.y0 - 1, l.n);
  }
  float2 wv[9][VP];
  if (l.ok) {
    const int c0 = l.cvi * V;
#pragma unroll
    for (int k = 0; k < 9; ++k) dwpair<T>::unpack(ld16(w9c + (size_t)k * C + c0), wv[k]);
  }
  dw_mbar_wait(bar_s, 0);
  if (!l.ok) return;
  T* obase = out.p + out.img(l.n) + (long long)l.x * out.sw + l.cvi * V;
  const uint4* tp = dw_tile + (2 * (warp * m.ppw + psub) * m.cvp + cvl);      // tile column of input x = 2 * x_out - 1, row 0
  const int rstride = TW * m.cvp;
  float2 f0[3][VP];                                   // input row 2 ty (filter row 0 of output row ty): carried over from row ty - 1
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) dwpair<T>::unpack(tp[kw * m.cvp], f0[kw]);
  for (int ty = 0; ty < l.y1 - l.y0; ++ty) {
    const uint4* r1 = tp + (2 * ty + 1) * rstride;
    float2 f1[3][VP], f2[3][VP], acc[VP];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) { dwpair<T>::unpack(r1[kw * m.cvp], f1[kw]); dwpair<T>::unpack(r1[rstride + kw * m.cvp], f2[kw]); }
#pragma unroll
    for (int j = 0; j < VP; ++j) {
      acc[j] = mul2(f0[0][j], wv[0][j]);
      acc[j] = fma2(f0[1][j], wv[1][j], acc[j]);
      acc[j] = fma2(f0[2][j], wv[2][j], acc[j]);
      acc[j] = fma2(f1[0][j], wv[3][j], acc[j]);
      acc[j] = fma2(f1[1][j], wv[4][j], acc[j]);
      acc[j] = fma2(f1[2][j], wv[5][j], acc[j]);
      acc[j] = fma2(f2[0][j], wv[6][j], acc[j]);
      acc[j] = fma2(f2[1][j], wv[7][j], acc[j]);
      acc[j] = fma2(f2[2][j], wv[8][j], acc[j]);
    }
    st16(obase + (long long)(l.y0 + ty) * out.sh, dwpair<T>::pack(acc));
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int j = 0; j < VP; ++j) f0[kw][j] = f2[kw][j];
  }
}

// ---- stride-1 weight gradient, staged + cluster-reduced ------------------------------------------------------------------
// dw_bwd_weight_s1d1_kernel above walks its strip with one row of global loads in flight per thread (a chain of rs + 2
// dependent L2 round trips: 11.5 of its 15.3 us on the 10 MB middle-flow tensors, tools/kbench.py) and every block then issues
// 9 * channels atomics (the other 3.8 us).  Here a block pulls its input tile (halo included) and its dout tile into shared
// memory with cp.async copies that are all in flight at once, walks them input-stationary from there, reduces its eight pixel
// columns through shared memory, and the blocks of one thread-block CLUSTER (the strips / images / parity sub-grids that share
// an x block and a channel group: consecutive blockIdx.z) add their partial sums through distributed shared memory
// (mapa + ld.shared::cluster), so that only one block's worth of atomics per cluster reaches the L2.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_saddr, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

template <typename T, int V, bool kTma, int S>
__global__ void __launch_bounds__(kDwThreads, 2) dw_bwd_weight_tile_kernel(DwView<const T> in, DwView<const T> dout, float* __restrict__ Gout,
                                                                          int C, DwMap m, int tap_stride, int c_stride, int spc, int det,
                                                                          const __grid_constant__ CUtensorMap in_map,
                                                                          const __grid_constant__ CUtensorMap g_map) {
  constexpr int VP = V / 2;
  extern __shared__ __align__(128) uint4 dww_smem[];  // in tile [rs + 2][ppb + 2][cvp] | dout tile [rs][ppb][cvp]   (then reused)
  __shared__ __align__(8) uint64_t tma_bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvl = lane & (m.cvp - 1), psub = lane / m.cvp;
  // S = 2 (TMA fill only): stride-2 layers - the input tile holds the 2 rs + 1 rows x 2 ppb + 1 pixels the rs x ppb dout pixels
  // gather from, and the row walk is dout-stationary (nine shared-memory loads per dout pixel)
  const int TW = S == 2 ? 2 * m.ppb + 1 : m.ppb + 2;
  const int TR = S == 2 ? 2 * m.rs + 1 : m.rs + 2;
  // blockIdx.z = (virtual image, chunk of `spc` consecutive strips): a block walks its strips one after the other and keeps the
  // 9 x V partial sums in registers, so the block reduction + cluster reduction + atomics are paid once per `spc` strips (the
  // 113 MB entry-flow tensors would otherwise pay them 2808 times for 160-pixel tiles)
  const int nchunks = (m.nstrips + spc - 1) / spc;
  const int vn = blockIdx.z / nchunks, chunk = blockIdx.z - vn * nchunks;
  const int x0 = blockIdx.x * m.ppb, x_base = S * x0 - 1;
  const int cv0 = blockIdx.y * m.cvp;
  const int H = in.h, W = in.w;
  uint4* in_t = dww_smem;
  uint4* g_t = dww_smem + ((TR * TW * m.cvp + 7) & ~7);               // 128-byte aligned (TMA destination)
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&tma_bar);
  uint32_t tma_phase = 0;
  if (kTma) {
    if (threadIdx.x == 0) dw_mbar_init(bar_s, 1);
    __syncthreads();
  }
  const int cshift = 31 - __clz(m.cvp);
  const int cl = threadIdx.x & (m.cvp - 1), cvi_f = cv0 + cl;
  const int pstep = kDwThreads >> cshift;
  float2 G2[9][VP];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < VP; ++j) G2[k][j] = make_float2(0.f, 0.f);
  pdl_sync();
  const int strip_end = min(m.nstrips, (chunk + 1) * spc);
  for (int strip = chunk * spc; strip < strip_end; ++strip) {
    const int y0 = strip * m.rs, y1 = min(dout.h, y0 + m.rs);
    const int orows = y1 - y0, nrows = orows + 2;     // dout rows y0 .. y1-1, input rows y0-1 .. y1
    const int y_base = S * y0 - 1;
    if (strip != chunk * spc) __syncthreads();        // the previous strip's walk is over: its tiles may be overwritten
    if (kTma) {
      // two boxes: input rows y0-1 .. y0+rs with a one-pixel halo left and right, dout rows y0 .. y0+rs-1; everything outside the
      // tensors (and the rows past a short last strip, which lie outside the image) arrives as zeros
      if (threadIdx.x == 0) {
        dw_mbar_expect_tx(bar_s, (uint32_t)((TR * TW + m.rs * m.ppb) * m.cvp * 16));
        dw_tma_load_tile(&in_map, bar_s, (uint32_t)__cvta_generic_to_shared(in_t), cv0 * V, x_base, y_base, vn, in.dsub, in.h);
        dw_tma_load_tile(&g_map, bar_s, (uint32_t)__cvta_generic_to_shared(g_t), cv0 * V, x0, y0, vn, dout.dsub, dout.h);
      }
      dw_mbar_wait(bar_s, tma_phase);
      tma_phase ^= 1u;
    } else {
      const uint32_t in_s = (uint32_t)__cvta_generic_to_shared(in_t);
      const T* nbase = in.p + in.img(vn);
      const int nvec = nrows * TW * m.cvp;
      int ty = (threadIdx.x >> cshift) / TW, tx = (threadIdx.x >> cshift) - ty * TW;
      for (int i = threadIdx.x; i < nvec; i += kDwThreads) {
        const int gy = y_base + ty, gx = x_base + tx;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W && cvi_f < m.cv;
        cp_async16_zfill(in_s + (uint32_t)i * 16u, ok ? nbase + (long long)gy * in.sh + (long long)gx * in.sw + cvi_f * V : in.p, ok);
        tx += pstep;
        if (tx >= TW) { tx -= TW; ++ty; }
      }
      const uint32_t g_s = (uint32_t)__cvta_generic_to_shared(g_t);
      const T* gbase = dout.p + dout.img(vn);
      const int gvec = orows * m.ppb * m.cvp;
      ty = (threadIdx.x >> cshift) / m.ppb; tx = (threadIdx.x >> cshift) - ty * m.ppb;
      for (int i = threadIdx.x; i < gvec; i += kDwThreads) {
        const int gy = y0 + ty, gx = x0 + tx;
        const bool ok = gx < dout.w && cvi_f < m.cv;
        cp_async16_zfill(g_s + (uint32_t)i * 16u, ok ? gbase + (long long)gy * dout.sh + (long long)gx * dout.sw + cvi_f * V : dout.p, ok);
        tx += pstep;
        while (tx >= m.ppb) { tx -= m.ppb; ++ty; }
      }
      cp_async_commit_wait_all();
      __syncthreads();
    }
    if (S == 2) {
      const int col = warp * m.ppw + psub;
      const uint4* ip = in_t + (2 * col * m.cvp + cvl);                   // tile column of input x = 2 * x_out - 1, row 0
      const uint4* gp = g_t + (col * m.cvp + cvl);
      const int irs = TW * m.cvp, grs = m.ppb * m.cvp;
      for (int ty = 0; ty < orows; ++ty) {
        float2 g[VP];
        dwpair<T>::unpack(gp[ty * grs], g);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const uint4* rp = ip + (2 * ty + kh) * irs;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            float2 f[VP];
            dwpair<T>::unpack(rp[kw * m.cvp], f);
#pragma unroll
            for (int j = 0; j < VP; ++j) G2[kh * 3 + kw][j] = fma2(f[j], g[j], G2[kh * 3 + kw][j]);
          }
        }
      }
    } else {
      // thread = (channel vector cvl, pixel column col of the block); input-stationary walk over the staged rows:
      // input row r meets dout rows r+1 (filter row 0), r (row 1), r-1 (row 2)
      const int col = warp * m.ppw + psub;
      const uint4* ip = in_t + (col * m.cvp + cvl);                       // tile column of x-1, row 0
      const uint4* gp = g_t + (col * m.cvp + cvl);
      const int irs = TW * m.cvp, grs = m.ppb * m.cvp;
      auto step = [&](int ty, const float2 (&gM)[VP], const float2 (&gC)[VP], float2 (&gP)[VP]) {
        // ty = input tile row (input row y0-1+ty); gP <- dout row y0+ty (tile row ty), zero past the strip
        float2 f[3][VP];
        const uint4* rp = ip + ty * irs;
        dwpair<T>::unpack(rp[0], f[0]);
        dwpair<T>::unpack(rp[m.cvp], f[1]);
        dwpair<T>::unpack(rp[2 * m.cvp], f[2]);
        if (ty < orows) dwpair<T>::unpack(gp[ty * grs], gP);
        else {
#pragma unroll
          for (int j = 0; j < VP; ++j) gP[j] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int j = 0; j < VP; ++j) {
            G2[kw][j] = fma2(f[kw][j], gP[j], G2[kw][j]);
            G2[3 + kw][j] = fma2(f[kw][j], gC[j], G2[3 + kw][j]);
            G2[6 + kw][j] = fma2(f[kw][j], gM[j], G2[6 + kw][j]);
          }
      };
      float2 g0[VP], g1[VP], g2[VP];
#pragma unroll
      for (int j = 0; j < VP; ++j) { g0[j] = make_float2(0.f, 0.f); g1[j] = g0[j]; g2[j] = g0[j]; }
      int ty = 0;
      while (true) {
        step(ty, g0, g1, g2);
        if (++ty >= nrows) break;
        step(ty, g1, g2, g0);
        if (++ty >= nrows) break;
        step(ty, g2, g0, g1);
        if (++ty >= nrows) break;
      }
    }
  }
  __syncthreads();                                    // tiles are dead: their memory becomes the reduction scratch
  // ---- block reduction over the 8 pixel columns (warps) [and the pixel lanes of a warp when a pixel has < 32 vectors] ----
  float* red = reinterpret_cast<float*>(dww_smem);     // [8 warps][32 lanes][3 * V]
  float* part = red + 8 * 32 * 3 * V;                  // [9][cvp * V] block partial sums
  float G[9][V];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < VP; ++j) { G[k][2 * j] = G2[k][j].x; G[k][2 * j + 1] = G2[k][j].y; }
  for (int o = m.cvp; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) G[k][j] += __shfl_xor_sync(0xffffffffu, G[k][j], o);
  }
  constexpr int PER = 3 * V;
  const int nch = m.cvp * V;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    if (kh) __syncthreads();
    if (lane < m.cvp) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int j = 0; j < V; ++j) red[(warp * 32 + lane) * PER + kw * V + j] = G[kh * 3 + kw][j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < m.cvp * PER; c += kDwThreads) {
      const int ln = c / PER, r = c - ln * PER;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += red[(w * 32 + ln) * PER + r];
      const int kw = r / V, j = r - kw * V;
      part[(kh * 3 + kw) * nch + ln * V + j] = sum;
    }
  }
  // ---- cluster reduction through distributed shared memory, then one block's worth of atomics per cluster ----
  cluster_sync_all();                                  // every block's `part` is complete and visible cluster-wide
  {
    const uint32_t rank = cluster_ctarank(), nrank = cluster_nctarank();
    // det: the cluster STORES its sums into workspace slice (blockIdx.x, cluster index along z), every element written by one rank
    if (det) Gout += ((size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.z / nrank)) * 9 * C;
    const int total = 9 * nch;
    const int chunk = (total + (int)nrank - 1) / (int)nrank;
    const int lo = (int)rank * chunk, hi = min(total, lo + chunk);
    const uint32_t part_s = (uint32_t)__cvta_generic_to_shared(part);
    const int cv_count = min(m.cvp, m.cv - cv0);
    for (int idx = lo + threadIdx.x; idx < hi; idx += kDwThreads) {
      float sum = 0.f;
      for (uint32_t r = 0; r < nrank; ++r) sum += ld_dsmem_f32(part_s + (uint32_t)idx * 4u, r);
      const int k = idx / nch, c = idx - k * nch;
      if (c < cv_count * V && !g_dww_noatomic) {
        float* dst = Gout + (size_t)k * tap_stride + (size_t)(cv0 * V + c) * c_stride;
        if (det) *dst = sum; else atomicAdd(dst, sum);
      }
    }
  }
  cluster_sync_all();                                  // no block may exit (and free its shared memory) while a peer still reads it
}

template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) dw_bwd_weight_direct_kernel(DwView<const T> in, DwView<const T> dout, float* __restrict__ Gout,
                                                                          int C, DwMap m, int s, int d, int tap_stride, int c_stride, int det) {
  const DwLane l = dw_lane(m, dout.h, dout.w);
  pdl_sync();
  float G[9][V];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < V; ++j) G[k][j] = 0.f;
  if (l.ok) {
    const int c0 = l.cvi * V;
    const T* base = in.p + in.img(l.n) + c0;
    const T* gbase = dout.p + dout.img(l.n) + (long long)l.x * dout.sw + c0;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    int ix[3];
    bool vx[3];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) { ix[kw] = l.x * s - d + kw * d; vx[kw] = ix[kw] >= 0 && ix[kw] < in.w; }
    for (int y = l.y0; y < l.y1; ++y) {
      uint4 t[9];
      const uint4 graw = ld16(gbase + (long long)y * dout.sh);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = y * s - d + kh * d;
        const bool vy = ih >= 0 && ih < in.h;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
          t[kh * 3 + kw] = (vy && vx[kw]) ? ld16(base + (long long)ih * in.sh + (long long)ix[kw] * in.sw) : zero;
      }
      float g[V];
      dwvec<T>::unpack(graw, g);
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float f[V];
        dwvec<T>::unpack(t[k], f);
#pragma unroll
        for (int j = 0; j < V; ++j) G[k][j] = fmaf(f[j], g[j], G[k][j]);
      }
    }
  }
  dw_reduce_G<V>(G, Gout, C, m, blockIdx.y * m.cvp, tap_stride, c_stride, det);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
template <typename T>
static bool dw_vec_ok(const dc_view& v) {
  const int V = dwvec<T>::V;
  return v.sc == 1 && (v.c % V == 0) && (v.sn % V == 0) && (v.sh % V == 0) && (v.sw % V == 0) &&
         ((reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0);
}
static inline dim3 dw_grid(const DwMap& m, int cols, int n_img) {
  return dim3((unsigned)ceil_div(cols, m.ppb), (unsigned)m.gy, (unsigned)(n_img * m.nstrips));
}

// DEEPCAM_B200_DW_TILE=0 selects the register-pipelined s1d1 kernels (for A/B measurements); default: staged tiles
static bool dw_tile_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DEEPCAM_B200_DW_TILE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
// rows per staged tile: measured on B200 (tools/kbench.py): 12 for the 48-row middle-flow tensors (9.6 us), 16 for the large
// entry-flow / decoder tensors (3.7 TB/s vs 2.7 TB/s for the register-pipelined kernel); DEEPCAM_B200_DW_TILE_ROWS overrides
static int dw_tile_rows(int H) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DEEPCAM_B200_DW_TILE_ROWS"); v = e ? atoi(e) : 0; if (v < 0 || v > 32) v = 0; }
  return v > 0 ? v : (H >= 96 ? 16 : 12);
}
// dilation d of a stride-1 layer as a parity split (see DwView): usable when both extents are multiples of d
static inline int dw_dsub(int s, int d, int h, int w) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("DEEPCAM_B200_DW_PARITY_SPLIT"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (s != 1) return 0;                                                // 0: not a stride-1 layer / not expressible: direct kernels
  if (d == 1) return 1;
  return (enabled && d <= 4 && h % d == 0 && w % d == 0) ? d : 0;
}

// ---- host side of the TMA tile fill ----------------------------------------------------------------------------------------
typedef CUresult (*PFN_dwEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_dwEncodeTiled dw_get_encode() {
  static PFN_dwEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_dwEncodeTiled>(f);
  });
  return fn;
}
// DEEPCAM_B200_DW_TMA=0: cp.async tile fills (A/B measurements)
static bool dw_tma_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DEEPCAM_B200_DW_TMA"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
// un-swizzled 4-D map {C, W, H, N} of an NHWC view with box {box_c, box_w, box_h, 1}; false = not expressible (the caller keeps
// the cp.async fill): box extents > 256, strides that are not multiples of 16 bytes, encode failure
static bool dw_encode_tile_map(CUtensorMap* map, const dc_view& v, int box_c, int box_w, int box_h, int dsub = 1) {
  if (!dw_tma_enabled()) return false;
  if (dsub > 1) {
    // parity sub-grids of a dilated layer: {C, W/d (stride d*sw), H/d (stride d*sh), px (stride sw), n*H + py (stride sh)}; the last
    // dimension folds image and row parity into one index, which needs densely stacked images (sn == H * sh)
    PFN_dwEncodeTiled enc5 = dw_get_encode();
    if (enc5 == nullptr) return false;
    const int es5 = v.dtype == DC_F32 ? 4 : 2;
    if (box_c > 256 || box_w > 256 || box_h > 256 || (box_c * es5) % 16) return false;
    if (v.sc != 1 || v.w % dsub || v.h % dsub || v.sn != (long long)v.h * v.sh || (v.sw * es5) % 16 || (v.sh * es5) % 16 ||
        (reinterpret_cast<uintptr_t>(v.ptr) % 16))
      return false;
    cuuint64_t dims[5] = {(cuuint64_t)v.c, (cuuint64_t)(v.w / dsub), (cuuint64_t)(v.h / dsub), (cuuint64_t)dsub, (cuuint64_t)v.n * v.h};
    cuuint64_t strides[4] = {(cuuint64_t)v.sw * dsub * es5, (cuuint64_t)v.sh * dsub * es5, (cuuint64_t)v.sw * es5, (cuuint64_t)v.sh * es5};
    cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc5(map, v.dtype == DC_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(v.ptr), dims,
                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
  }
  PFN_dwEncodeTiled enc = dw_get_encode();
  if (enc == nullptr) return false;
  const int es = v.dtype == DC_F32 ? 4 : 2;
  if (box_c > 256 || box_w > 256 || box_h > 256 || box_c < 1 || box_w < 1 || box_h < 1 || (box_c * es) % 16) return false;
  if (v.sc != 1 || (v.sw * es) % 16 || (v.sh * es) % 16 || (v.sn * es) % 16 || (reinterpret_cast<uintptr_t>(v.ptr) % 16)) return false;
  cuuint64_t dims[4] = {(cuuint64_t)v.c, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * es, (cuuint64_t)v.sh * es, (cuuint64_t)std::max<long long>(v.sn, 1) * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, v.dtype == DC_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <typename T, int V>
static bool dw_s1d1_tile_launch(const dc_view& in, const void* w, const dc_view& out, int flip, int acc, cudaStream_t st, int dsub = 1) {
  if (!dw_tile_enabled()) return false;
  const int oh = out.h / dsub, ow = out.w / dsub, nimg = out.n * dsub * dsub;
  DwMap m = dw_map(out.c, V, oh, ow, nimg, 1 << 30, dw_tile_rows(oh));
  dim3 grid = dw_grid(m, ow, nimg);
  const size_t smem = (size_t)(m.rs + 2) * (m.ppb + 2) * m.cvp * 16;
  if (smem > 200 * 1024) return false;                                         // odd row counts: register-pipelined kernel
  static bool attr_set[2] = {false, false};
  CUtensorMap map;
  const bool tma = dw_encode_tile_map(&map, in, m.cvp * V, m.ppb + 2, m.rs + 2, dsub);
  if (!attr_set[tma]) {
    cudaError_t e = tma ? cudaFuncSetAttribute(dw_s1d1_tile_kernel<T, V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                        : cudaFuncSetAttribute(dw_s1d1_tile_kernel<T, V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return false;
    attr_set[tma] = true;
  }
  if (tma)
    launch_k(dw_s1d1_tile_kernel<T, V, true>, grid, dim3(kDwThreads), smem, st, dw_view<const T>(in, dsub), (const T*)w, dw_view<T>(out, dsub), out.c,
             m, flip, acc, map);
  else {
    memset(&map, 0, sizeof(map));
    launch_k(dw_s1d1_tile_kernel<T, V, false>, grid, dim3(kDwThreads), smem, st, dw_view<const T>(in, dsub), (const T*)w, dw_view<T>(out, dsub), out.c,
             m, flip, acc, map);
  }
  return true;
}

template <typename T, int V>
static int dw_fwd_bn_t(const dc_bn_params& p, const dc_view& in, const void* w, const dc_view& act, const dc_view& out, cudaStream_t st) {
  DwMap m = dw_map(out.c, V, out.h, out.w, out.n, 1 << 30, dw_tile_rows(out.h));
  dim3 grid = dw_grid(m, out.w, out.n);
  const size_t smem = (size_t)(m.rs + 2) * (m.ppb + 2) * m.cvp * 16 + (size_t)2 * m.cvp * V * sizeof(float);
  if (smem > 200 * 1024) return fail(-2, "dc_dw_fwd_bn: tile does not fit (rows per strip %d)", m.rs);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dw_s1d1_tile_bn_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail((int)e, "dc_dw_fwd_bn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  DwView<T> av = act.ptr ? dw_view<T>(act) : DwView<T>{nullptr, 0, 0, 0, 0, 0, 1};
  launch_k(dw_s1d1_tile_bn_kernel<T, V>, grid, dim3(kDwThreads), smem, st, p, dw_view<const T>(in), (const T*)w, av, dw_view<T>(out), out.c, m);
  return launch_status("dc_dw_fwd_bn");
}

// rows per tile of the fused kernel: (rows + 2) halo rows + rows of y (+ rows of a and of the old gradient) must leave room
// for 2 blocks per SM
template <typename T, int V>
static int dw_bwd_data_bnred_t(const dc_view& dout, const void* w, const dc_view& din, int acc, const dc_view& y, const dc_view& xact,
                               const float* fcoef, double* sums, int relu, cudaStream_t st) {
  const bool has_x = xact.ptr != nullptr;
  // exact strip height: 10 rows (5 with the two extra staged tiles) keep two blocks per SM resident, and the 48-row
  // middle-flow tensors then need 270 blocks = one wave
  DwMap m = dw_map(din.c, V, din.h, din.w, din.n, 1 << 30, 1);
  m.rs = std::min(din.h, has_x ? 5 : 10);
  m.nstrips = ceil_div(din.h, m.rs);
  dim3 grid = dw_grid(m, din.w, din.n);
  const size_t in_b = (size_t)(m.rs + 2) * (m.ppb + 2) * m.cvp * 16, int_b = (size_t)m.rs * m.ppb * m.cvp * 16;
  const size_t smem = std::max(in_b + int_b * (has_x ? 3 : 1) + (size_t)2 * m.cvp * V * sizeof(float), (size_t)8 * 32 * 2 * V * sizeof(float));
  if (smem > 110 * 1024) return fail(-2, "dc_dw_bwd_data_bnred: tile does not fit (rows per strip %d)", m.rs);
  static bool attr_set[2] = {false, false};
  if (!attr_set[has_x]) {
    cudaError_t e = has_x ? cudaFuncSetAttribute(dw_s1d1_tile_bnred_kernel<T, V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)
                          : cudaFuncSetAttribute(dw_s1d1_tile_bnred_kernel<T, V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) return fail((int)e, "dc_dw_bwd_data_bnred: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[has_x] = true;
  }
  if (has_x)
    launch_k(dw_s1d1_tile_bnred_kernel<T, V, true>, grid, dim3(kDwThreads), smem, st, dw_view<const T>(dout), (const T*)w, dw_view<T>(din),
             din.c, m, acc, dw_view<const T>(y), dw_view<const T>(xact), fcoef, sums, relu);
  else
    launch_k(dw_s1d1_tile_bnred_kernel<T, V, false>, grid, dim3(kDwThreads), smem, st, dw_view<const T>(dout), (const T*)w, dw_view<T>(din),
             din.c, m, acc, dw_view<const T>(y), dw_view<const T>(y), fcoef, sums, relu);
  return launch_status("dc_dw_bwd_data_bnred");
}

template <typename T>
static int dw_fwd_t(const dc_view& in, const void* w, int s, int d, const dc_view& out, cudaStream_t st) {
  constexpr int V = dwvec<T>::V;
  const int dsub = dw_dsub(s, d, out.h, out.w);
  if (dsub >= 1 && dw_s1d1_tile_launch<T, V>(in, w, out, 0, 0, st, dsub)) return launch_status("dc_dw_fwd");
  static int s2_tile = -1;        // DEEPCAM_B200_DW_S2_TILE=0: gather-form kernels (A/B measurements)
  if (s2_tile < 0) { const char* e = getenv("DEEPCAM_B200_DW_S2_TILE"); s2_tile = (e && e[0] == '0') ? 0 : 1; }
  if (s == 2 && d == 1 && s2_tile && dw_tile_enabled()) {
    DwMap mt = dw_map(out.c, V, out.h, out.w, out.n, 1 << 30, 1);
    mt.nstrips = ceil_div(out.h, 4);                   // strips of <= 4 output rows: 9 input rows x (2 ppb + 1) pixels stay below 110 KB
    mt.rs = ceil_div(out.h, mt.nstrips);
    mt.nstrips = ceil_div(out.h, mt.rs);
    const size_t smem = (size_t)(2 * mt.rs + 1) * (2 * mt.ppb + 1) * mt.cvp * 16;
    CUtensorMap map;
    if (smem <= 110 * 1024 && dw_encode_tile_map(&map, in, mt.cvp * V, 2 * mt.ppb + 1, 2 * mt.rs + 1)) {
      static bool attr_set = false;
      if (!attr_set && cudaFuncSetAttribute(dw_fwd_s2_tile_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) == cudaSuccess)
        attr_set = true;
      if (attr_set) {
        launch_k(dw_fwd_s2_tile_kernel<T, V>, dw_grid(mt, out.w, out.n), dim3(kDwThreads), smem, st, (const T*)w, dw_view<T>(out), out.c, mt, map);
        return launch_status("dc_dw_fwd");
      }
    }
  }
  DwMap m = dw_map(out.c, V, out.h, out.w, out.n, kNumSMs * 4, 6);
  dim3 grid = dw_grid(m, out.w, out.n);
  if (s == 1 && d == 1)
    launch_k(dw_s1d1_kernel<T, V>, grid, dim3(kDwThreads), (size_t)0, st, dw_view<const T>(in), (const T*)w, dw_view<T>(out), out.c, m, 0, 0);
  else
    launch_k(dw_direct_kernel<T, V>, grid, dim3(kDwThreads), (size_t)0, st, dw_view<const T>(in), (const T*)w, dw_view<T>(out), out.c, m, s, d, 0, 0);
  return launch_status("dc_dw_fwd");
}
template <typename T>
static int dw_bwd_data_t(const dc_view& dout, const void* w, int s, int d, const dc_view& din, int acc, cudaStream_t st) {
  constexpr int V = dwvec<T>::V;
  const int dsub = dw_dsub(s, d, din.h, din.w);
  if (dsub >= 1 && dw_s1d1_tile_launch<T, V>(dout, w, din, 1, acc, st, dsub)) return launch_status("dc_dw_bwd_data");
  DwMap m = dw_map(din.c, V, din.h, din.w, din.n, kNumSMs * 4, 6);
  dim3 grid = dw_grid(m, din.w, din.n);
  if (s == 1 && d == 1)          // full correlation with the flipped filter
    launch_k(dw_s1d1_kernel<T, V>, grid, dim3(kDwThreads), (size_t)0, st, dw_view<const T>(dout), (const T*)w, dw_view<T>(din), din.c, m, 1, acc);
  else if (s == 1)
    launch_k(dw_direct_kernel<T, V>, grid, dim3(kDwThreads), (size_t)0, st, dw_view<const T>(dout), (const T*)w, dw_view<T>(din), din.c, m, 1, d, 1, acc);
  else {
    static int s2_tile = -1;      // DEEPCAM_B200_DW_S2_TILE=0: gather-form kernel (A/B measurements)
    if (s2_tile < 0) { const char* e = getenv("DEEPCAM_B200_DW_S2_TILE"); s2_tile = (e && e[0] == '0') ? 0 : 1; }
    if (s == 2 && d == 1 && s2_tile && dout.h == (din.h - 1) / 2 + 1 && dout.w == (din.w - 1) / 2 + 1) {
      DwMap mt = dw_map(dout.c, V, dout.h, dout.w, dout.n, 1 << 30, 8);       // strips of >= 8 dout rows
      mt.nstrips = ceil_div(dout.h, 16);
      mt.rs = ceil_div(dout.h, mt.nstrips);
      mt.nstrips = ceil_div(dout.h, mt.rs);
      const size_t smem = (size_t)(mt.rs + 1) * (mt.ppb + 1) * mt.cvp * 16;
      static bool attr_done[2] = {false, false};
      CUtensorMap g_map;
      const bool tma = smem <= 110 * 1024 && dw_encode_tile_map(&g_map, dout, mt.cvp * V, mt.ppb + 1, mt.rs + 1);
      if (!tma) memset(&g_map, 0, sizeof(g_map));
      if (!attr_done[tma]) {
        cudaError_t e = tma ? cudaFuncSetAttribute(dw_bwd_data_s2_tile_kernel<T, V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)
                            : cudaFuncSetAttribute(dw_bwd_data_s2_tile_kernel<T, V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        if (e == cudaSuccess) attr_done[tma] = true;
      }
      if (attr_done[tma] && smem <= 110 * 1024) {
        if (tma)
          launch_k(dw_bwd_data_s2_tile_kernel<T, V, true>, dw_grid(mt, dout.w, dout.n), dim3(kDwThreads), smem, st, dw_view<const T>(dout), (const T*)w,
                   dw_view<T>(din), din.c, mt, acc, g_map);
        else
          launch_k(dw_bwd_data_s2_tile_kernel<T, V, false>, dw_grid(mt, dout.w, dout.n), dim3(kDwThreads), smem, st, dw_view<const T>(dout), (const T*)w,
                   dw_view<T>(din), din.c, mt, acc, g_map);
        return launch_status("dc_dw_bwd_data");
      }
    }
    launch_k(dw_bwd_data_strided_kernel<T, V>, grid, dim3(kDwThreads), (size_t)0, st, dw_view<const T>(dout), (const T*)w, dw_view<T>(din), din.c, m, s, d, acc);
  }
  return launch_status("dc_dw_bwd_data");
}
// second stage of the deterministic depthwise weight gradient: G[k * tap_stride + c * c_stride] += sum over slices (in order)
__global__ void __launch_bounds__(256) dww_slice_reduce_kernel(const float* __restrict__ ws, int nslices, int C, float* __restrict__ G,
                                                               int tap_stride, int c_stride) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // i = k * C + c
  if (i >= 9 * C) return;
  float a = 0.f;
  for (int z = 0; z < nslices; ++z) a += ws[(size_t)z * 9 * C + i];
  const int k = i / C, c = i - k * C;
  G[(size_t)k * tap_stride + (size_t)c * c_stride] += a;
}

struct DwwPlan {
  bool tile;
  bool s2;                     // stride-2 staged form (TMA fill only)
  int dsub, gn, gw, spc;
  DwMap m;
  dim3 grid;
  size_t smem;
};
template <typename T>
static DwwPlan dww_plan(const dc_view& in, const dc_view& dout, int s, int d) {
  constexpr int V = dwvec<T>::V;
  // (a shared-memory staged variant with one tile per block was measured slower: the 9*C atomics per block dominate; the tile kernel
  //  below reduces across a cluster first, the register-pipelined kernels keep long strips per block and reduce once)
  static int tb_mult = -1, min_rows = -1;   // sweep knobs: DEEPCAM_B200_DWW_BLOCKS_PER_SM, DEEPCAM_B200_DWW_MIN_ROWS
  if (tb_mult < 0) { const char* e = getenv("DEEPCAM_B200_DWW_BLOCKS_PER_SM"); tb_mult = e ? std::max(1, atoi(e)) : 2; }
  if (min_rows < 0) { const char* e = getenv("DEEPCAM_B200_DWW_MIN_ROWS"); min_rows = e ? std::max(1, atoi(e)) : 12; }
  static int tile_on = -1;        // DEEPCAM_B200_DWW_TILE=0: the register-pipelined kernel (A/B measurements)
  if (tile_on < 0) { const char* e = getenv("DEEPCAM_B200_DWW_TILE"); tile_on = (e && e[0] == '0') ? 0 : 1; }
  DwwPlan pl;
  pl.dsub = dw_dsub(s, d, dout.h, dout.w);
  const int gh = pl.dsub >= 1 ? dout.h / pl.dsub : dout.h;
  pl.gw = pl.dsub >= 1 ? dout.w / pl.dsub : dout.w;
  pl.gn = pl.dsub >= 1 ? dout.n * pl.dsub * pl.dsub : dout.n;
  pl.tile = false;
  pl.s2 = false;
  pl.spc = 1;
  static int s2_on = -1;          // DEEPCAM_B200_DW_S2_TILE=0: gather-form kernels for the stride-2 layers
  if (s2_on < 0) { const char* e = getenv("DEEPCAM_B200_DW_S2_TILE"); s2_on = (e && e[0] == '0') ? 0 : 1; }
  if (s == 2 && d == 1 && tile_on && s2_on && dw_tma_enabled()) {
    // strips of <= 4 dout rows: 9 input rows x (2 ppb + 1) pixels + the dout tile stay below 110 KB
    DwMap m = dw_map(dout.c, V, dout.h, dout.w, dout.n, 1 << 30, 1);
    m.nstrips = ceil_div(dout.h, 4);
    m.rs = ceil_div(dout.h, m.nstrips);
    m.nstrips = ceil_div(dout.h, m.rs);
    const size_t in_vecs = ((size_t)(2 * m.rs + 1) * (2 * m.ppb + 1) * m.cvp + 7) & ~(size_t)7;
    const size_t tiles = (in_vecs + (size_t)m.rs * m.ppb * m.cvp) * 16;
    const size_t scratch = ((size_t)8 * 32 * 3 * V + (size_t)9 * m.cvp * V) * sizeof(float);
    const size_t smem = std::max(tiles, scratch);
    if (smem <= 110 * 1024) {
      const long long nblk1 = (long long)ceil_div(dout.w, m.ppb) * m.gy * dout.n * m.nstrips;
      pl.spc = (int)std::min<long long>(m.nstrips, std::max<long long>(1, nblk1 / 592));
      const int nchunks = ceil_div(m.nstrips, pl.spc);
      pl.grid = dim3((unsigned)ceil_div(dout.w, m.ppb), (unsigned)m.gy, (unsigned)(dout.n * nchunks));
      pl.m = m;
      pl.smem = smem;
      pl.tile = true;
      pl.s2 = true;
      pl.dsub = 1;               // plain views
      pl.gw = dout.w; pl.gn = dout.n;
      return pl;
    }
  }
  if (pl.dsub >= 1 && tile_on) {
    // strips of <= 10 rows: input + dout tiles stay below 110 KB, two blocks per SM; the 48-row middle-flow tensors then make
    // 270 blocks = one wave, in 54 clusters of 5 strips
    DwMap m = dw_map(dout.c, V, gh, pl.gw, pl.gn, 1 << 30, 1);
    m.nstrips = ceil_div(gh, 10);
    m.rs = ceil_div(gh, m.nstrips);
    m.nstrips = ceil_div(gh, m.rs);
    const size_t in_vecs = ((size_t)(m.rs + 2) * (m.ppb + 2) * m.cvp + 7) & ~(size_t)7;      // dout tile starts 128-byte aligned
    const size_t tiles = (in_vecs + (size_t)m.rs * m.ppb * m.cvp) * 16;
    const size_t scratch = ((size_t)8 * 32 * 3 * V + (size_t)9 * m.cvp * V) * sizeof(float);
    const size_t smem = std::max(tiles, scratch);
    if (smem <= 110 * 1024) {
      // strips per block: one for the 48-row tensors (270 blocks = one wave); the large entry-flow / decoder tensors get
      // several strips per block so that ~600 blocks walk the tensor and reduce once each
      const long long nblk1 = (long long)ceil_div(pl.gw, m.ppb) * m.gy * pl.gn * m.nstrips;
      pl.spc = (int)std::min<long long>(m.nstrips, std::max<long long>(1, nblk1 / 592));
      const int nchunks = ceil_div(m.nstrips, pl.spc);
      pl.grid = dim3((unsigned)ceil_div(pl.gw, m.ppb), (unsigned)m.gy, (unsigned)(pl.gn * nchunks));
      pl.m = m;
      pl.smem = smem;
      pl.tile = true;
      return pl;
    }
  }
  pl.m = dw_map(dout.c, V, gh, pl.gw, pl.gn, kNumSMs * tb_mult, min_rows);
  pl.grid = dw_grid(pl.m, pl.gw, pl.gn);
  pl.smem = (size_t)8 * 32 * 3 * V * sizeof(float);
  return pl;
}

/* ws != null: deterministic two-stage form (dc_dw_bwd_weight_det) */
template <typename T>
static int dw_bwd_weight_t(const dc_view& in, const dc_view& dout, int s, int d, float* G, int param_layout, float* ws, long long ws_elems,
                           cudaStream_t st) {
  const int g_tap_stride = param_layout ? 1 : dout.c, g_c_stride = param_layout ? 9 : 1;
  constexpr int V = dwvec<T>::V;
  {
    static int noat = -1;
    if (noat < 0) { const char* e = getenv("DEEPCAM_B200_DWW_NOATOMIC"); noat = (e && e[0] == '1') ? 1 : 0; if (noat) cudaMemcpyToSymbol(g_dww_noatomic, &noat, sizeof(int)); }
  }
  DwwPlan pl = dww_plan<T>(in, dout, s, d);
  const int det = ws != nullptr ? 1 : 0;
  if (det)
    DC_REQUIRE(ws_elems >= (long long)pl.grid.x * pl.grid.z * 9 * dout.c, "dc_dw_bwd_weight_det: workspace of %lld floats required, %lld given",
               (long long)pl.grid.x * pl.grid.z * 9 * dout.c, ws_elems);
  float* target = det ? ws : G;
  const int tap_stride = det ? dout.c : g_tap_stride, c_stride = det ? 1 : g_c_stride;
  int nslices = (int)(pl.grid.x * pl.grid.z);
  bool launched = false;
  if (pl.tile) {
    static bool attr_done[3] = {false, false, false};
    CUtensorMap in_map, g_map;
    const bool tma = pl.s2 ? (dw_encode_tile_map(&in_map, in, pl.m.cvp * V, 2 * pl.m.ppb + 1, 2 * pl.m.rs + 1) &&
                              dw_encode_tile_map(&g_map, dout, pl.m.cvp * V, pl.m.ppb, pl.m.rs))
                           : (dw_encode_tile_map(&in_map, in, pl.m.cvp * V, pl.m.ppb + 2, pl.m.rs + 2, pl.dsub) &&
                              dw_encode_tile_map(&g_map, dout, pl.m.cvp * V, pl.m.ppb, pl.m.rs, pl.dsub));
    if (!tma) { memset(&in_map, 0, sizeof(in_map)); memset(&g_map, 0, sizeof(g_map)); }
    const int variant = pl.s2 ? 2 : (tma ? 1 : 0);
    if (!attr_done[variant] && !(pl.s2 && !tma)) {
      cudaError_t e = variant == 2 ? cudaFuncSetAttribute(dw_bwd_weight_tile_kernel<T, V, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)
                      : variant == 1 ? cudaFuncSetAttribute(dw_bwd_weight_tile_kernel<T, V, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)
                                     : cudaFuncSetAttribute(dw_bwd_weight_tile_kernel<T, V, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
      if (e == cudaSuccess) attr_done[variant] = true;
    }
    const bool attr_set = attr_done[variant] && !(pl.s2 && !tma);     // the stride-2 form exists with the TMA fill only
    if (attr_set) {
      // cluster = consecutive blockIdx.z (strips / images / parity sub-grids of one (x block, channel group)): largest divisor <= 8
      static int cz_max = -1;         // DEEPCAM_B200_DWW_CLUSTER: largest cluster size tried (1 = no cluster reduction)
      if (cz_max < 0) { const char* e = getenv("DEEPCAM_B200_DWW_CLUSTER"); cz_max = e ? std::max(1, std::min(8, atoi(e))) : 8; }
      for (int cz = cz_max; cz >= 1 && !launched; --cz) {
        if (pl.grid.z % cz) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = pl.grid; cfg.blockDim = dim3(kDwThreads); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = cz;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        cudaError_t e =
            variant == 2 ? cudaLaunchKernelEx(&cfg, dw_bwd_weight_tile_kernel<T, V, true, 2>, dw_view<const T>(in, pl.dsub), dw_view<const T>(dout, pl.dsub),
                                              target, dout.c, pl.m, tap_stride, c_stride, pl.spc, det, in_map, g_map)
            : variant == 1 ? cudaLaunchKernelEx(&cfg, dw_bwd_weight_tile_kernel<T, V, true, 1>, dw_view<const T>(in, pl.dsub), dw_view<const T>(dout, pl.dsub),
                                                target, dout.c, pl.m, tap_stride, c_stride, pl.spc, det, in_map, g_map)
                           : cudaLaunchKernelEx(&cfg, dw_bwd_weight_tile_kernel<T, V, false, 1>, dw_view<const T>(in, pl.dsub), dw_view<const T>(dout, pl.dsub),
                                                target, dout.c, pl.m, tap_stride, c_stride, pl.spc, det, in_map, g_map);
        if (e == cudaSuccess) {
          launched = true;
          nslices = (int)(pl.grid.x * (pl.grid.z / cz));
        } else {
          cudaGetLastError();                          // this cluster size cannot be scheduled here: try the next divisor
        }
      }
    }
    if (!launched) {                                   // the staged kernel cannot run here: register-pipelined kernel, its own grid
      pl.m = dw_map(dout.c, V, pl.dsub >= 1 ? dout.h / pl.dsub : dout.h, pl.gw, pl.gn, kNumSMs * 2, 12);
      pl.grid = dw_grid(pl.m, pl.gw, pl.gn);
      pl.smem = (size_t)8 * 32 * 3 * V * sizeof(float);
      nslices = (int)(pl.grid.x * pl.grid.z);
      if (det)
        DC_REQUIRE(ws_elems >= (long long)nslices * 9 * dout.c, "dc_dw_bwd_weight_det: workspace too small for the fallback kernel (%lld floats)",
                   (long long)nslices * 9 * dout.c);
    }
  }
  if (!launched) {
    if (pl.dsub >= 1 && !pl.s2)
      launch_k(dw_bwd_weight_s1d1_kernel<T, V>, pl.grid, dim3(kDwThreads), pl.smem, st, dw_view<const T>(in, pl.dsub), dw_view<const T>(dout, pl.dsub),
               target, dout.c, pl.m, tap_stride, c_stride, det);
    else
      launch_k(dw_bwd_weight_direct_kernel<T, V>, pl.grid, dim3(kDwThreads), pl.smem, st, dw_view<const T>(in), dw_view<const T>(dout), target, dout.c,
               pl.m, s, d, tap_stride, c_stride, det);
  }
  if (int r = launch_status("dc_dw_bwd_weight")) return r;
  if (!det) return 0;
  launch_k(dww_slice_reduce_kernel, dim3((unsigned)ceil_div(9 * dout.c, 256)), dim3(256), (size_t)0, st, (const float*)ws, nslices, dout.c, G,
           g_tap_stride, g_c_stride);
  return launch_status("dc_dw_bwd_weight_det");
}

static int check_dw(const char* what, const dc_view& big, const dc_view& small, int s, int d) {
  DC_REQUIRE(view_ok(big) && view_ok(small), "%s: bad views", what);
  DC_REQUIRE(big.dtype == small.dtype && big.c == small.c && big.n == small.n, "%s: dtype/channel/batch mismatch", what);
  const bool vec = big.dtype == DC_F32 ? (dw_vec_ok<float>(big) && dw_vec_ok<float>(small))
                                       : (dw_vec_ok<__nv_bfloat16>(big) && dw_vec_ok<__nv_bfloat16>(small));
  DC_REQUIRE(vec, "%s: views must be channel-contiguous and 16-byte aligned with C %% 8 == 0 (bf16) / C %% 4 == 0 (fp32)", what);
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
  DC_REQUIRE(w9c != nullptr && (reinterpret_cast<uintptr_t>(w9c) % 16) == 0, "dc_dw_fwd: weights must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  return in.dtype == DC_F32 ? dw_fwd_t<float>(in, w9c, stride, dil, out, st) : dw_fwd_t<__nv_bfloat16>(in, w9c, stride, dil, out, st);
}

int dc_dw_fwd_bn(const dc_bn_params* p, dc_view y, const void* w9c, dc_view act, dc_view out, void* stream) {
  DC_REQUIRE(p != nullptr && p->sums != nullptr && p->gamma != nullptr && p->beta != nullptr, "dc_dw_fwd_bn: null BatchNorm argument");
  DC_REQUIRE((p->flags & DC_BN_TRAIN) && (p->flags & DC_BN_SUMS_READY) && !(p->flags & DC_BN_IDENTITY),
             "dc_dw_fwd_bn: train-mode BatchNorm whose batch sums are already in the workspace (DC_BN_SUMS_READY) required");
  DC_REQUIRE(p->count > 1.0, "dc_dw_fwd_bn: more than one value per channel required");
  if (int r = check_dw("dc_dw_fwd_bn", y, out, 1, 1)) return r;
  if (act.ptr != nullptr) {
    if (int r = check_dw("dc_dw_fwd_bn(act)", y, act, 1, 1)) return r;
  }
  DC_REQUIRE(w9c != nullptr && (reinterpret_cast<uintptr_t>(w9c) % 16) == 0, "dc_dw_fwd_bn: weights must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  return y.dtype == DC_F32 ? dw_fwd_bn_t<float, 4>(*p, y, w9c, act, out, st) : dw_fwd_bn_t<__nv_bfloat16, 8>(*p, y, w9c, act, out, st);
}

int dc_dw_bwd_data(dc_view dout, const void* w9c, int stride, int dil, dc_view din, int accumulate, void* stream) {
  if (int r = check_dw("dc_dw_bwd_data", din, dout, stride, dil)) return r;
  DC_REQUIRE(w9c != nullptr && (reinterpret_cast<uintptr_t>(w9c) % 16) == 0, "dc_dw_bwd_data: weights must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  return din.dtype == DC_F32 ? dw_bwd_data_t<float>(dout, w9c, stride, dil, din, accumulate, st)
                             : dw_bwd_data_t<__nv_bfloat16>(dout, w9c, stride, dil, din, accumulate, st);
}

int dc_dw_bwd_data_bnred(dc_view dout, const void* w9c, dc_view din, int accumulate, dc_view y, dc_view act, const void* fwd_ws,
                         void* bwd_ws, int relu, void* stream) {
  if (int r = check_dw("dc_dw_bwd_data_bnred", din, dout, 1, 1)) return r;
  DC_REQUIRE(w9c != nullptr && (reinterpret_cast<uintptr_t>(w9c) % 16) == 0, "dc_dw_bwd_data_bnred: weights must be 16-byte aligned");
  DC_REQUIRE(bwd_ws != nullptr && view_ok(y) && same_shape(y, din) && y.dtype == din.dtype, "dc_dw_bwd_data_bnred: y must match the gradient");
  const bool y_ok = din.dtype == DC_F32 ? dw_vec_ok<float>(y) : dw_vec_ok<__nv_bfloat16>(y);
  DC_REQUIRE(y_ok, "dc_dw_bwd_data_bnred: y must be channel-contiguous and 16-byte aligned");
  if (act.ptr != nullptr) {
    const bool a_ok = view_ok(act) && same_shape(act, din) && act.dtype == din.dtype &&
                      (din.dtype == DC_F32 ? dw_vec_ok<float>(act) : dw_vec_ok<__nv_bfloat16>(act));
    DC_REQUIRE(a_ok, "dc_dw_bwd_data_bnred: activation view must match the gradient");
  } else if (relu) {
    DC_REQUIRE(fwd_ws != nullptr, "dc_dw_bwd_data_bnred: the forward BatchNorm workspace is required to recompute the ReLU mask");
  }
  // workspace layout (bn.cu): double sums[2][C] | float coef[4][C] | ...
  const float* fcoef = fwd_ws ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(fwd_ws) + (size_t)16 * din.c) : nullptr;
  double* sums = reinterpret_cast<double*>(bwd_ws);
  cudaStream_t st = as_stream(stream);
  return din.dtype == DC_F32 ? dw_bwd_data_bnred_t<float, 4>(dout, w9c, din, accumulate, y, act, fcoef, sums, relu, st)
                             : dw_bwd_data_bnred_t<__nv_bfloat16, 8>(dout, w9c, din, accumulate, y, act, fcoef, sums, relu, st);
}

int dc_dw_bwd_weight(dc_view in, dc_view dout, int stride, int dil, float* G9c, int param_layout, void* stream) {
  if (int r = check_dw("dc_dw_bwd_weight", in, dout, stride, dil)) return r;
  DC_REQUIRE(G9c != nullptr, "dc_dw_bwd_weight: null gradient");
  cudaStream_t st = as_stream(stream);
  return in.dtype == DC_F32 ? dw_bwd_weight_t<float>(in, dout, stride, dil, G9c, param_layout, nullptr, 0, st)
                            : dw_bwd_weight_t<__nv_bfloat16>(in, dout, stride, dil, G9c, param_layout, nullptr, 0, st);
}

long long dc_dw_bwd_weight_ws_elems(dc_view in, dc_view dout, int stride, int dil) {
  if (check_dw("dc_dw_bwd_weight_ws_elems", in, dout, stride, dil)) return -1;
  const DwwPlan pl = in.dtype == DC_F32 ? dww_plan<float>(in, dout, stride, dil) : dww_plan<__nv_bfloat16>(in, dout, stride, dil);
  long long n = (long long)pl.grid.x * pl.grid.z;
  if (pl.tile) {                                       // the register-pipelined fallback of the staged kernel has its own grid
    const int V = in.dtype == DC_F32 ? 4 : 8;
    const DwMap m = dw_map(dout.c, V, dout.h / pl.dsub, pl.gw, pl.gn, kNumSMs * 2, 12);
    const dim3 g = dw_grid(m, pl.gw, pl.gn);
    n = std::max(n, (long long)g.x * g.z);
  }
  return n * 9 * dout.c;
}

/* deterministic form: every block (cluster) stores its partial sums into its own workspace slice, a second launch adds the slices to
   G9c in slice order.  G9c is accumulated into, as in dc_dw_bwd_weight. */
int dc_dw_bwd_weight_det(dc_view in, dc_view dout, int stride, int dil, float* G9c, int param_layout, float* ws, long long ws_elems,
                         void* stream) {
  if (int r = check_dw("dc_dw_bwd_weight_det", in, dout, stride, dil)) return r;
  DC_REQUIRE(G9c != nullptr && ws != nullptr, "dc_dw_bwd_weight_det: null gradient or workspace");
  cudaStream_t st = as_stream(stream);
  return in.dtype == DC_F32 ? dw_bwd_weight_t<float>(in, dout, stride, dil, G9c, param_layout, ws, ws_elems, st)
                            : dw_bwd_weight_t<__nv_bfloat16>(in, dout, stride, dil, G9c, param_layout, ws, ws_elems, st);
}

}  // extern "C"
