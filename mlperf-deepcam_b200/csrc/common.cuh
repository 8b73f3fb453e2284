// Shared device/host helpers for the deepcam_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/deepcam_b200.h"

namespace dc {

// ---- error reporting ---------------------------------------------------------------------------
char* err_buf();                       // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);

#define DC_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return ::dc::fail(-1, __VA_ARGS__);            \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// A step is ~770 short kernels on one stream.  Launched through launch_k(), kernel N+1 may be scheduled while
// kernel N drains: its blocks run their prologue (barrier init, TMEM allocation, tensor-map prefetch, index
// arithmetic) and then block in pdl_wait() until kernel N has completed and its memory is visible.  Rules that keep
// this equivalent to plain stream order:
//   * every kernel launched through launch_k() executes pdl_wait() before its first global-memory access
//     (reads AND writes: the predecessor may still be reading what this kernel overwrites);
//   * pdl_trigger() comes after pdl_wait(), so at most one successor is ever queued behind a running kernel;
//   * kernels with an inter-block barrier trigger only after the barrier (all their blocks are resident by then:
//     a queued successor can never take the slot of a block the barrier is waiting for).
// Both instructions are no-ops when the kernel was launched without the attribute (dc_set_pdl(0), or <<<>>>).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // errors surface through launch_status()
}

// ---- element access ----------------------------------------------------------------------------
template <typename T> struct elem;
template <> struct elem<float> {
  static constexpr int dtype = DC_F32;
  __device__ static __forceinline__ float ld(const float* p) { return *p; }
  __device__ static __forceinline__ void st(float* p, float v) { *p = v; }
  // 4 contiguous elements, 16-byte aligned
  __device__ static __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct elem<__nv_bfloat16> {
  static constexpr int dtype = DC_BF16;
  __device__ static __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  // 4 contiguous elements, 8-byte aligned
  __device__ static __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float4 r;
    r.x = __uint_as_float(u.x << 16);
    r.y = __uint_as_float(u.x & 0xffff0000u);
    r.z = __uint_as_float(u.y << 16);
    r.w = __uint_as_float(u.y & 0xffff0000u);
    return r;
  }
  __device__ static __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// Round a float the way the storage type would (identity for fp32).
template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

// Device-side view with typed pointer.
template <typename T>
struct View {
  T* p;
  int n, h, w, c;
  long long sn, sh, sw, sc;
  __device__ __forceinline__ T* at(int in_, int ih, int iw) const {
    return p + in_ * sn + ih * sh + iw * sw;
  }
};
template <typename T>
static inline View<T> make_view(const dc_view& v) {
  View<T> r;
  r.p = reinterpret_cast<T*>(v.ptr);
  r.n = v.n; r.h = v.h; r.w = v.w; r.c = v.c;
  r.sn = v.sn; r.sh = v.sh; r.sw = v.sw; r.sc = v.sc;
  return r;
}

static inline size_t dtype_size(int dt) { return dt == DC_BF16 ? 2 : 4; }
static inline bool view_ok(const dc_view& v) {
  return v.ptr != nullptr && v.n > 0 && v.h > 0 && v.w > 0 && v.c > 0 && (v.dtype == DC_F32 || v.dtype == DC_BF16);
}
// channel-vectorisable: unit channel stride, C % 4 == 0, base and strides aligned to 4 elements
static inline bool view_vec4(const dc_view& v) {
  size_t es = dtype_size(v.dtype);
  return v.sc == 1 && (v.c % 4 == 0) && (v.sn % 4 == 0) && (v.sh % 4 == 0) && (v.sw % 4 == 0) &&
         ((reinterpret_cast<uintptr_t>(v.ptr) % (4 * es)) == 0);
}
static inline bool same_shape(const dc_view& a, const dc_view& b) {
  return a.n == b.n && a.h == b.h && a.w == b.w && a.c == b.c;
}

// bn.cu: sums[0][c] += sum of y[.., c], sums[1][c] += sum of squares (no finalize); used when a producer cannot do it itself
int bn_accumulate_sums(const dc_view& y, double* sums, cudaStream_t st);

// ---- BatchNorm workspace (bn.cu; also read by the BatchNorm-fused depthwise kernels in dw.cu) ----------------
// dc_bn_ws_bytes(C) bytes: double sums[2][C] | float coef[4][C] | uint32 ticket[16]
struct BnWs {
  double* sums;     // [2][C]
  float* coef;      // [4][C]
  unsigned* ticket;
};
__host__ __device__ static inline BnWs bn_ws(void* ws, int C) {
  BnWs w;
  w.sums = reinterpret_cast<double*>(ws);
  w.coef = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + (size_t)16 * C);
  w.ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + (size_t)32 * C);
  return w;
}

// 1/sqrt(v) in fp32: hardware approximation + one Newton-Raphson step (~1 ulp)
__device__ __forceinline__ float inv_sqrt_f32(float v) {
  float r = rsqrtf(v);
  return r * (1.5f - 0.5f * v * r * r);
}

// ---- warp / block reductions -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block reduction of NACC x V per-thread partials over the pixel lanes of a block, then fp64 atomics.
// smem: float red[NW warps][32 lanes][NACC*V]
// (shared by bn.cu and the BatchNorm-reducing depthwise backward in dw.cu; lane = psub * cvp + channel-vector lane)
template <int NACC, int V, int NW = 8>
__device__ __forceinline__ void reduce_to_ws(float (&acc)[NACC][V], double* dst, int C, int cvp, int cv_base, int cv_count) {
  extern __shared__ float red[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // lanes of a warp that share a channel vector (different pixels): butterfly over the pixel-sub index
  for (int o = cvp; o < 32; o <<= 1) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[a][j] += __shfl_xor_sync(0xffffffffu, acc[a][j], o);
  }
  constexpr int PER = NACC * V;
  if (lane < cvp) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < V; ++j) red[(warp * 32 + lane) * PER + a * V + j] = acc[a][j];
  }
  __syncthreads();
  for (int col = threadIdx.x; col < cvp * PER; col += NW * 32) {
    const int l = col / PER, r = col - l * PER;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[(w * 32 + l) * PER + r];
    const int a = r / V, j = r - a * V;
    const int cvi = cv_base + l;
    if (l < cv_count) atomicAdd(dst + (size_t)a * C + cvi * V + j, (double)s);
  }
}


// ---- deterministic two-stage reductions (dc_*_det entry points, dc_set_deterministic) --------------------------------------------
// A split reduction stores the partial result of split z into slice z of a workspace (one writer per element); this launch then
// adds the slices to the destination in slice order: dst[wt * per_tap + i] += sum_z ws[z * slice_stride + wt * per_tap + i] for the
// weight taps of the call.  Same summation order on every run, whatever order the splits finished in.
struct SplitReduceTaps { int n; int wt[DC_MAX_TAPS]; };
int launch_split_reduce(const float* ws, int nslices, long long slice_stride, const SplitReduceTaps& taps, long long per_tap, float* dst,
                        cudaStream_t st);
bool deterministic();          // dc_set_deterministic(): kernels that need no workspace pick their single-writer form

}  // namespace dc
