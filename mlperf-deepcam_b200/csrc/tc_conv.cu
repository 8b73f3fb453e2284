// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (bf16 operands, fp32 accumulation).
//
// fprop-like kernel (Conv2d fprop + dgrad, ConvTranspose2d fprop per parity class + dgrad):
//   D[m, co] = sum_t sum_ci A_t[m, ci] * B[co, t, ci]
//   * M tile = 128 output pixels arranged as a TH x TW rectangle of one image; the A operand of tap t is
//     one 4-D TMA box {64 ch, TW, TH, 1} of the NHWC activation tensor at the tap's offset.  Out-of-range
//     coordinates are zero-filled by TMA, which implements the zero padding (no im2col, no F.pad copy).
//   * strided gathers (1x1 stride-2 skips, ConvTranspose dgrad) use up to four "parity" tensor maps
//     (sub-grids with doubled strides), so every box is unit-stride.
//   * B = packed weights [Co][taps][Ci_pad64] (K-major), one 2-D TMA box {64, BN}.
//   * both operands land in shared memory in the 128-byte-swizzled K-major UMMA canonical layout;
//     one elected thread issues tcgen05.mma (M=128, N=BN, K=16) accumulating in TMEM.
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
//     (tcgen05.ld -> bias -> bf16 -> global, generic output strides so concat slices / parity sub-grids work); the
//     persistent kernel adds a second epilogue group (warps 6..9) that drains alternate 64-column chunks.
// wgrad kernel: D[co, ci] = sum_pixels dY[p, co] * X_t[p, ci]; both operands are pixel-major in memory,
//   i.e. MN-major UMMA operands (a_major = b_major = 1), reduction (pixels) split across CTAs, fp32 atomics.
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <mutex>
#include <stdlib.h>

namespace dc {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
//   [46,48) version = 1 (Blackwell) | [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10,
// a_major @15, b_major @16 (0 = K-major, 1 = MN-major), N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------------
struct TcMaps {
  CUtensorMap a[4];   // activation (gathered operand) parity maps
  CUtensorMap b;      // fprop: packed weights; wgrad: dY
  CUtensorMap c;      // fprop (persistent kernel): output view for the TMA-store epilogue
};

struct TcOut {
  void* p;
  int n, h, w, c;
  long long sn, sh, sw, sc;
  int dtype;
  int csplit;            // dc_conv_desc.out_csplit / out_split_off: channels >= csplit go to a second segment (0 = off)
  long long split_off;
};

__device__ __forceinline__ long long out_col_off(const TcOut& o, int co) {
  return (o.csplit > 0 && co >= o.csplit) ? o.split_off + (long long)(co - o.csplit) * o.sc : (long long)co * o.sc;
}

struct TcFpropParams {
  int ntaps;
  int map_id[DC_MAX_TAPS];
  int qh[DC_MAX_TAPS], qw[DC_MAX_TAPS];   // box offset (in the tap's parity sub-grid) relative to the output tile origin
  int wt[DC_MAX_TAPS];
  int kblocks;         // ceil(Ci / 64)
  int TH, TW, tiles_x, tiles_y;
  int accumulate;
  int out_vec_ok;      // output rows are 16-byte aligned bf16 with unit channel stride
  const float* bias;
  TcOut out;
  double* stats;       // BatchNorm workspace sums[2][stats_C] of the layer that normalises `out` (or null): the epilogue adds the
  int stats_C;         // per-channel sum and sum of squares of the bf16-rounded outputs it stores (TMA-store path only)
  // eval-mode BatchNorm (+ReLU) folded into the epilogue (TMA-store path only; aff_gamma null = off):
  //   out = [relu]((acc + bias) * scale + shift), scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale
  int b_early;         // DC_CONV_WEIGHTS_STABLE: the first weight tiles may be fetched before griddepcontrol.wait
  const float* aff_gamma;
  const float* aff_beta;
  const float* aff_mean;
  const float* aff_var;
  float aff_eps;
  int aff_relu;
};

// Build with -DDC_TC_TRACE (tools/tc_trace.py) to record SM-clock timestamps of the persistent kernel's phases per CTA:
// slot 0 entry, 1 after griddepcontrol.wait, 2 first operand stage landed, 3 last MMA issued, 4 accumulator complete (seen by
// the epilogue), 5 epilogue done, 6 exit, 7 globaltimer at entry,
// 8 first chunk converted + staged, 9 first chunk's TMA store issued, 10 first chunk's statistics done.  The product build compiles the macro away.
#ifdef DC_TC_TRACE
__device__ unsigned long long g_tc_trace[148 * 16];
// (the "memory" clobber keeps the clock read on its side of barriers; a plain clock64() was hoisted above the final __syncthreads)
#define TC_TRACE(slot) do { unsigned long long tc_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tc_) :: "memory"); \
                            g_tc_trace[blockIdx.x * 16 + (slot)] = tc_; } while (0)
// per-TILE timeline of CTA 0 in halo mode (tools/halo_trace.py): [tile < 32][slot]: 0 halo load issued, 1 producer warp 0 saw the
// halo, 2 producer warp 0 done with the tile, 3 MMA lane owns the accumulator, 4 MMA lane committed the tile, 5 epilogue saw the
// accumulator, 6 epilogue released it, 7 epilogue done with the tile
__device__ unsigned long long g_halo_trace[32 * 8 + 8];      // + 8: phases of one k block of producer warp 0 in tile 6
#define HALO_TRACE2(i, slot) do { if (blockIdx.x == 0 && (i) == 6 && pw == ((i) * cfg.nkb) % 4 && lane == 0) { unsigned long long tc_; \
                                  asm volatile("mov.u64 %0, %%clock64;" : "=l"(tc_) :: "memory"); g_halo_trace[32 * 8 + (slot)] = tc_; } } while (0)
#define HALO_TRACE(i, slot) do { if (blockIdx.x == 0 && (i) < 32) { unsigned long long tc_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tc_) :: "memory"); \
                                 g_halo_trace[(i) * 8 + (slot)] = tc_; } } while (0)
#else
#define TC_TRACE(slot) do { } while (0)
#define HALO_TRACE(i, slot) do { } while (0)
#define HALO_TRACE2(i, slot) do { } while (0)
#endif

constexpr int kTcThreads = 192;
constexpr int kABytes = 128 * 128;   // 128 rows x 64 bf16

// Runtime tile configuration of the fprop-like kernel.
struct TcFpropCfg {
  int BN;            // N tile (multiple of 16, 16..256) = UMMA N
  int stages;        // smem pipeline depth
  int stage_bytes;   // 16384 (A) + BN*128 (B); multiple of 1024
  int tmem_cols;     // power of two >= max(32, BN)
  int smem_bytes;
};

__global__ void __launch_bounds__(kTcThreads) conv_gemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcFpropParams p,
                                                                  const TcFpropCfg cfg) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int STAGES = cfg.stages;
  const int BN = cfg.BN;
  const uint32_t bar_base = smem_base + STAGES * cfg.stage_bytes;
  // barriers: full[STAGES], empty[STAGES], tmem_full, then tmem base pointer slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  int mt = blockIdx.x;
  const int tile_x = mt % p.tiles_x; mt /= p.tiles_x;
  const int tile_y = mt % p.tiles_y;
  const int img = mt / p.tiles_y;
  const int x0 = tile_x * p.TW, y0 = tile_y * p.TH;
  const int n0 = blockIdx.y * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, cfg.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.a[p.map_id[0]]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_sync();     // everything above overlapped the previous kernel's tail; global memory is touched only below

  const int total_k = p.ntaps * p.kblocks;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        const CUtensorMap* am = &maps.a[p.map_id[t]];
        const int bx = x0 + p.qw[t], by = y0 + p.qh[t];
        const int kb0 = p.wt[t] * p.kblocks;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), cfg.stage_bytes);
          const uint32_t sa = smem_base + s * cfg.stage_bytes;
          tma_load_4d(am, full_bar(s), sa, kb * 64, bx, by, img);
          tma_load_2d(&maps.b, full_bar(s), sa + kABytes, (kb0 + kb) * 64, n0);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, BN, 0, 0);
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < total_k; ++it) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * cfg.stage_bytes;
        const uint64_t da = make_smem_desc(sa, 16, 1024);
        const uint64_t db = make_smem_desc(sa + kABytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (>>4) address field
          umma_bf16(tmem_base, da + 2u * k, db + 2u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ---- epilogue: warps 2..5; warp w may only touch TMEM lanes 32*(w%4) .. +31 ----
    // Phase 1: TMEM -> registers -> (+bias) -> per-warp staging rows in shared memory (the pipeline stages are
    //          free once tmem_full has fired).  Phase 2: lanes walk along the channels of one pixel row at a time,
    //          so every global store / load-add-store instruction covers one contiguous run of the output row.
    const int lg = warp & 3;
    const int row = lg * 32 + lane;             // row of the 128 x BN accumulator = pixel of the tile
    const int ty = row / p.TW, tx = row - ty * p.TW;
    const int oy = y0 + ty, ox = x0 + tx;
    const bool pix_ok = (oy < p.out.h) && (ox < p.out.w);
    const long long base = (long long)img * p.out.sn + (long long)oy * p.out.sh + (long long)ox * p.out.sw;
    const int ncols = min(BN, p.out.c - n0);    // valid columns of this tile
    const bool stage_f32 = (p.accumulate != 0) || (p.out.dtype == DC_F32);
    const int es = stage_f32 ? 4 : 2;
    // bytes per staged row: whole 32-column chunks; (pitch/16) is odd -> conflict-free 16-byte row accesses
    const int pitch = ((BN + 31) & ~31) * es + 16;
    uint8_t* stg = smem_gen + (size_t)lg * 32 * pitch;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < ncols) v[j] = __float_as_uint(__uint_as_float(v[j]) + p.bias[n0 + c0 + j]);
      }
      uint8_t* rp = stg + (size_t)lane * pitch + (size_t)c0 * es;
      if (stage_f32) {
#pragma unroll
        for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(rp + q * 16) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 r;
          __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
          __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
          __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
          __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
          r.x = *reinterpret_cast<uint32_t*>(&b0); r.y = *reinterpret_cast<uint32_t*>(&b1);
          r.z = *reinterpret_cast<uint32_t*>(&b2); r.w = *reinterpret_cast<uint32_t*>(&b3);
          *reinterpret_cast<uint4*>(rp + q * 16) = r;
        }
      }
    }
    __syncwarp();
    const unsigned okmask = __ballot_sync(0xffffffffu, pix_ok);
    const int V = stage_f32 ? 4 : 8;            // staged elements per 16-byte lane access
    // lanes_per_row lanes walk along the channels of one pixel row; 32/lanes_per_row rows per warp instruction
    int lpr = 1;
    while (lpr < 32 && lpr * V < ncols) lpr <<= 1;
    const int rpi = 32 / lpr;
    const int my_sub = lane / lpr, my_l = lane - my_sub * lpr;
    const bool raw_copy = !stage_f32 && p.out_vec_ok;     // bf16 staged, bf16 vector output: move 16 bytes as they are
#pragma unroll 2
    for (int r0 = 0; r0 < 32; r0 += rpi) {
      const int r = r0 + my_sub;
      const long long rbase = __shfl_sync(0xffffffffu, base, r);
      const bool ok = (okmask >> r) & 1u;
      const uint8_t* rp = stg + (size_t)r * pitch;
      for (int col = my_l * V; col < ncols; col += lpr * V) {
        if (!ok) continue;
        const int co = n0 + col;
        if (raw_copy && col + 8 <= ncols) {
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out.p) + rbase + co) =
              *reinterpret_cast<const uint4*>(rp + (size_t)col * 2);
          continue;
        }
        float f[8];
        if (stage_f32) {
          const float4 t = *reinterpret_cast<const float4*>(rp + (size_t)col * 4);
          f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
        } else {
          const uint4 t = *reinterpret_cast<const uint4*>(rp + (size_t)col * 2);
          const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w4[j] << 16); f[2 * j + 1] = __uint_as_float(w4[j] & 0xffff0000u); }
        }
        if (stage_f32 && p.out_vec_ok && col + 4 <= ncols) {
          // accumulate into a bf16 row: staged fp32, 4 channels (8 bytes of output) per lane
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out.p) + rbase + co;
          if (p.accumulate) {
            const uint2 old = *reinterpret_cast<const uint2*>(op);
            f[0] += __uint_as_float(old.x << 16); f[1] += __uint_as_float(old.x & 0xffff0000u);
            f[2] += __uint_as_float(old.y << 16); f[3] += __uint_as_float(old.y & 0xffff0000u);
          }
          __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
          uint2 o;
          o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
          *reinterpret_cast<uint2*>(op) = o;
        } else if (p.out.dtype == DC_F32 && p.out.sc == 1 && col + 4 <= ncols && ((rbase + out_col_off(p.out, co)) & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(p.out.p) & 15) == 0)) {
          float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out.p) + rbase + out_col_off(p.out, co));
          float4 o = make_float4(f[0], f[1], f[2], f[3]);
          if (p.accumulate) { const float4 old = *q; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
          *q = o;
        } else {
          for (int j = 0; j < V && col + j < ncols; ++j) {
            const long long off = rbase + out_col_off(p.out, co + j);
            float val = f[j];
            if (p.out.dtype == DC_F32) {
              float* q = reinterpret_cast<float*>(p.out.p) + off;
              if (p.accumulate) val += *q;
              *q = val;
            } else {
              __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(p.out.p) + off;
              if (p.accumulate) val += __bfloat162float(*q);
              *q = __float2bfloat16_rn(val);
            }
          }
        }
      }
    }
  }
  __syncwarp();        // warps 0/1: the single-lane role (producer / MMA issuer) rejoins its warp, so every warp arrives at the
                       // teardown barrier exactly once and the TMEM release below is ordered after the epilogue's last read
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, cfg.tmem_cols);
}


// ------------------------------------------------------------------------------------------------
// fprop-like kernel, second generation: PERSISTENT CTAs (one per SM) walking a static tile list, two TMEM
// accumulators so the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1, epilogue staged in 64-column
// chunks through a dedicated shared-memory buffer, and a B-RESIDENT mode for layers whose whole packed weight tile
// fits in shared memory (1x1 convolutions with K*N <= 64K elements: the weights are loaded once per CTA instead of
// once per tile, which removes ~2/3 of the L2 traffic of those HBM-bound layers).
// ------------------------------------------------------------------------------------------------
struct TcV2Cfg {
  int BN;              // N tile (multiple of 16, <= 256)
  int stages;          // pipeline stages
  int stage_bytes;     // 16384 (A) [+ BN*128 (B) when streaming]
  int b_resident;      // whole weight tile resident in shared memory (requires n_ntiles == 1)
  int bres_bytes;      // wtaps * kblocks * BN * 128 when resident
  int acc_stride;      // TMEM columns between the two accumulators (power of two >= BN)
  int tmem_cols;       // 2 * acc_stride
  int n_mtiles, n_ntiles, total_tiles;
  int tma_store;       // epilogue writes 128 x 64 bf16 chunks with TMA (dense bf16 output views)
  int bm2;             // 256-row tiles: two M tiles (two accumulators) share every B tile -> 30 % fewer L2->SM bytes per FLOP;
                       // single accumulator set (no epilogue/main-loop overlap), used when one round covers the problem
  int real_mtiles;     // number of 128-row M tiles (n_mtiles counts 256-row super tiles when bm2)
  int wide;            // N tile of 257..512 columns: two MMAs per k step (bn_sub0 + the rest) into ONE accumulator of BN TMEM
  int bn_sub0;         // columns, B tile fetched as two TMA boxes of b_box_rows rows; used when it turns a two-round problem
  int b_box_rows;      // (e.g. 54 M tiles x 728 columns) into one round over fewer, fatter CTAs
  // HALO mode (conv_gemm_tc2_kernel<true>): few-channel / few-output layers whose per-tap TMA boxes would re-read the input once
  // per tap through the L2->SM path (conv1, conv2 and its dgrad, last_deconv fprop + dgrad: 9 x resp. 4 x their HBM bytes).
  // The input region of a tile (halo included) is loaded ONCE per tile by one un-swizzled TMA box, four producer warps copy it
  // tap by tap into the 128B-swizzled K-major A stages (im2col in shared memory, K = tap-major dense: k = slice * Ci + ci), the
  // whole weight tile is resident, and the MMA / epilogue side is the same as in the streaming mode.
  int halo;            // 1 = this mode
  int halo_w, halo_h;  // pixels of the staged input region
  int halo_bytes;      // halo_w * halo_h * Ci * 2
  int halo_stride;     // bytes between the halo buffers (multiple of 1024)
  int halo_bufs;       // 1 or 2 (double-buffered across tiles when shared memory allows)
  int halo_ci;         // gathered channels
  int halo_s;          // gather stride (1 | 2)
  int halo_x0, halo_y0; // smallest tap offset (dw, dh): input coordinate of the region's first pixel relative to stride * tile origin
  int halo_nbox;        // the region is fetched as halo_nbox boxes of {256 flattened (w, c) elements, halo_h rows}: W and C are one
                        // contiguous dimension of a dense NHWC row, so a TMA "row" is 512 bytes instead of one Ci-wide pixel (the
                        // TMA unit is row-rate bound: 561 32-byte rows per tile for conv1 took 1.7 us, 3 x 17 rows of 512 B do not)
  int nkb;             // K blocks of 64 = ceil(wtaps * Ci / 64)
  int ktot;            // wtaps * Ci
  int dbg;             // experiments only (DEEPCAM_B200_TC_HALO_DBG, results are garbage): 1 = producers skip the copy, 2 = skip the MMAs
  int smem_bytes;
};
constexpr int kV2StagePitchF32 = 64 * 4 + 16;     // staged row: 64 fp32 columns + 16 B (odd multiple of 16 B: conflict-free)
// generic epilogue: 4 warps x 32 staged rows; TMA-store epilogue: one 16 KB chunk buffer + 2 KB of statistics scratch per
// epilogue group
constexpr int kV2StagingBytes = 2 * 16384 + 2 * 2048;
static_assert(kV2StagingBytes >= 4 * 32 * kV2StagePitchF32, "staging must also hold the generic epilogue's rows");
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue group A, warps 6..9 = epilogue group B.  With one warp
// per scheduler the epilogue is latency-bound (tools/tc_trace.py: 0.7 us per 128 x 64 chunk, 4.3 of the 11 us of a 728 x 728
// layer); the two groups drain alternate 64-column chunks of the same accumulator concurrently.
constexpr int kTc2Threads = 320;
constexpr int kTc2HaloThreads = 448;          // + warps 10..13: im2col producers of the halo mode
constexpr int kHaloTH = 8, kHaloTW = 16;      // halo-mode tile: 128 output pixels as 8 rows x 16 columns (1.41x halo overhead for a 3x3 stride-1 gather)
__device__ __forceinline__ void epi_group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <bool kHalo>
__global__ void __launch_bounds__(kHalo ? kTc2HaloThreads : kTc2Threads, 1) conv_gemm_tc2_kernel(const __grid_constant__ TcMaps maps,
                                                                                                const TcFpropParams p, const TcV2Cfg cfg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int STAGES = cfg.stages;
  const int BN = cfg.BN;
  const uint32_t bres_base = smem_base + STAGES * cfg.stage_bytes;
  const uint32_t halo_region = kHalo ? (uint32_t)(cfg.halo_bufs * cfg.halo_stride) : 0u;
  const uint32_t halo_off = (uint32_t)(STAGES * cfg.stage_bytes + cfg.bres_bytes);          // halo buffers follow the resident weights
  // wide mode runs exactly one tile per CTA: the epilogue staging aliases pipeline stage 0 (all MMAs have completed, hence
  // all stages have been consumed, before the first accumulator read), which buys a third pipeline stage
  const uint32_t stg_off = cfg.wide ? 0u : (uint32_t)(STAGES * cfg.stage_bytes + cfg.bres_bytes) + halo_region;
  const uint32_t bar_base = smem_base + STAGES * cfg.stage_bytes + cfg.bres_bytes + halo_region + (cfg.wide ? 0u : (uint32_t)kV2StagingBytes);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  const uint32_t bres_bar = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 5);
  auto hfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 6 + b); };      // up to 4 halo buffers
  auto hempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 10 + b); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef DC_TC_TRACE
  if (threadIdx.x == 0) {
    TC_TRACE(0);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    g_tc_trace[blockIdx.x * 16 + 7] = gt;
  }
#endif

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }   // (halo: ONE producer warp builds a stage)
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), cfg.bm2 ? 16 : 8); }   // every epilogue warp of both groups arrives once per (sub-)tile
    mbar_init(bres_bar, 1);
    if (kHalo) for (int b = 0; b < 4; ++b) { mbar_init(hfull_bar(b), 1); mbar_init(hempty_bar(b), 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, cfg.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.a[p.map_id[0]]);
    if (cfg.tma_store) tma_prefetch_desc(&maps.c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const int b_block_bytes = BN * 128;
  // Weights that the preceding kernel did not write (DC_CONV_WEIGHTS_STABLE) are fetched BEFORE griddepcontrol.wait: the
  // resident weight tile, or the B halves of the first pipeline stages of this CTA's first tile, are in flight while the
  // previous kernel drains; only the activation operand (and every store) waits.  early_b = stages whose B part is issued here.
  int early_b = 0;
  if (p.b_early && warp == 0 && lane == 0 && blockIdx.x < (unsigned)cfg.total_tiles) {
    if (cfg.b_resident) {
      mbar_expect_tx(bres_bar, cfg.bres_bytes);
      const int nblk = cfg.bres_bytes / b_block_bytes;
      for (int j = 0; j < nblk; ++j) tma_load_2d(&maps.b, bres_bar, bres_base + j * b_block_bytes, j * 64, 0);
      early_b = 1;
    } else {
      const int MT0 = cfg.bm2 ? 2 : 1;
      const int n0 = ((int)blockIdx.x / cfg.n_mtiles) * BN;
      const int total_k0 = p.ntaps * p.kblocks;
      early_b = min(STAGES, total_k0);
      for (int it = 0; it < early_b; ++it) {
        const int t = it / p.kblocks, kb = it - t * p.kblocks;
        const int kcol = (p.wt[t] * p.kblocks + kb) * 64;
        mbar_expect_tx(full_bar(it), cfg.stage_bytes);
        const uint32_t sa = smem_base + it * cfg.stage_bytes;
        tma_load_2d(&maps.b, full_bar(it), sa + MT0 * kABytes, kcol, n0);
        if (cfg.wide) tma_load_2d(&maps.b, full_bar(it), sa + MT0 * kABytes + cfg.b_box_rows * 128, kcol, n0 + cfg.b_box_rows);
      }
    }
  }
  pdl_sync();     // everything above overlapped the previous kernel's tail; activations and outputs are touched only below
  if (threadIdx.x == 0) TC_TRACE(1);

  if (warp == 0) {
    if (lane == 0) {
      if (cfg.b_resident && !early_b) {
        mbar_expect_tx(bres_bar, cfg.bres_bytes);
        const int nblk = cfg.bres_bytes / b_block_bytes;
        for (int j = 0; j < nblk; ++j) tma_load_2d(&maps.b, bres_bar, bres_base + j * b_block_bytes, j * 64, 0);
      }
      if (cfg.b_resident) early_b = 0;                 // (the pipeline stages carry no weights in this mode)
      if (kHalo) {
        // one un-swizzled box per tile: the tile's whole input region, halo included; out-of-range pixels are zero-filled (= padding)
        int i = 0;
        for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x, ++i) {
          int mt = tile;
          const int tile_x = mt % p.tiles_x; mt /= p.tiles_x;
          const int tile_y = mt % p.tiles_y;
          const int img = mt / p.tiles_y;
          const int hb = i % cfg.halo_bufs;
          const uint32_t par = (uint32_t)(i / cfg.halo_bufs) & 1u;
          mbar_wait(hempty_bar(hb), par ^ 1u);
          if (cfg.dbg & 8) { mbar_arrive(hfull_bar(hb)); continue; }      // experiment: no halo load at all
          mbar_expect_tx(hfull_bar(hb), cfg.halo_bytes);
          const int fx = (tile_x * p.TW * cfg.halo_s + cfg.halo_x0) * cfg.halo_ci;       // first flattened (w, c) element
          const int hy = tile_y * p.TH * cfg.halo_s + cfg.halo_y0;
          HALO_TRACE(i, 0);
          for (int b = 0; b < cfg.halo_nbox; ++b)
            tma_load_3d(&maps.a[0], hfull_bar(hb), smem_base + halo_off + (uint32_t)(hb * cfg.halo_stride + b * cfg.halo_h * 512), fx + b * 256,
                        hy, img);
        }
      } else {
      int s = 0; uint32_t ph = 0;
      int issued = 0;                                  // k steps issued so far by this CTA (first tile first)
      const int MT = cfg.bm2 ? 2 : 1;
      for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x) {
        const int nt = tile / cfg.n_mtiles;
        const int mts = tile - nt * cfg.n_mtiles;
        int x0[2], y0[2], img[2];
        for (int sub = 0; sub < MT; ++sub) {
          int mt = mts * MT + sub;
          const bool real = mt < cfg.real_mtiles;
          const int tile_x = mt % p.tiles_x; mt /= p.tiles_x;
          const int tile_y = mt % p.tiles_y;
          img[sub] = real ? mt / p.tiles_y : p.out.n;          // past the last image: TMA zero-fills the box
          x0[sub] = tile_x * p.TW; y0[sub] = tile_y * p.TH;
        }
        const int n0 = nt * BN;
        for (int t = 0; t < p.ntaps; ++t) {
          const CUtensorMap* am = &maps.a[p.map_id[t]];
          const int kb0 = p.wt[t] * p.kblocks;
          for (int kb = 0; kb < p.kblocks; ++kb, ++issued) {
            const bool b_done = issued < early_b;      // this stage's weights (and its expect_tx) were issued before the wait
            if (!b_done) {
              mbar_wait(empty_bar(s), ph ^ 1u);
              mbar_expect_tx(full_bar(s), cfg.stage_bytes);
            }
            const uint32_t sa = smem_base + s * cfg.stage_bytes;
            tma_load_4d(am, full_bar(s), sa, kb * 64, x0[0] + p.qw[t], y0[0] + p.qh[t], img[0]);
            if (cfg.bm2) tma_load_4d(am, full_bar(s), sa + kABytes, kb * 64, x0[1] + p.qw[t], y0[1] + p.qh[t], img[1]);
            if (!cfg.b_resident && !b_done) {
              tma_load_2d(&maps.b, full_bar(s), sa + MT * kABytes, (kb0 + kb) * 64, n0);
              if (cfg.wide)
                tma_load_2d(&maps.b, full_bar(s), sa + MT * kABytes + cfg.b_box_rows * 128, (kb0 + kb) * 64, n0 + cfg.b_box_rows);
            }
            if (++s == STAGES) { s = 0; ph ^= 1u; }
          }
        }
      }
      }   // !kHalo
    }
  } else if (kHalo && warp >= 10) {
    // ---- halo mode: im2col producers.  A stage = 128 rows (tile pixels, row-major TH x TW) x 64 K elements, K-major, 128B
    //      swizzle; K element k = slice * Ci + ci, chunk j of k block kb covers k0 = kb*64 + j*8 .. +7 (Ci % 8 == 0, so a chunk
    //      never straddles two taps).  lane = (rsub, j): a warp instruction moves the 8 chunks (128 contiguous bytes of the
    //      destination row) of 4 rows; the warp owns rows 32*pw .. +31. ----
    const int pw = warp - 10, j = lane & 7, rsub = lane >> 3;
    const int Ci = cfg.halo_ci, hs = cfg.halo_s;
    // Producer warp pw builds the k blocks kb = pw, pw + 4, ... of every tile ALONE (all 128 rows, 32 warp instructions of
    // 4 rows x 8 chunks), so the four warps work on four different pipeline stages at once and every stage costs ONE
    // generic->async proxy fence and ONE barrier arrival (with all four warps on the same k block the per-block handshake
    // - 4 arrivals, fence, MMA commit, empty wait - serialised: 0.68 us per k block, measured round 2).
    int i = 0;
    for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x, ++i) {
      const int hb = i % cfg.halo_bufs;
      mbar_wait(hfull_bar(hb), (uint32_t)(i / cfg.halo_bufs) & 1u);
      if (pw == 0 && lane == 0) HALO_TRACE(i, 1);
      const uint8_t* hptr = smem_gen + halo_off + (uint32_t)(hb * cfg.halo_stride);
      // k block q (running number over this CTA's tiles) belongs to warp q % 4: five k blocks per tile then alternate between
      // the warps instead of always giving warp 0 two of them
      const int q0 = i * cfg.nkb;
      for (int kb = ((pw - q0) % 4 + 4) % 4; kb < cfg.nkb; kb += 4) {
        const int q = q0 + kb;                          // stage and phase follow from the running k-block number
        const int s = q % STAGES;
        const uint32_t ph = (uint32_t)(q / STAGES) & 1u;
        const int k0 = kb * 64 + j * 8;
        int t = -1;
        int tqh = 0, tf0 = 0;
        if (k0 < cfg.ktot) {
          const int slice = k0 / Ci, ci0 = k0 - slice * Ci;
          for (int qq = 0; qq < p.ntaps; ++qq) if (p.wt[qq] == slice) t = qq;      // tap that owns this weight slice (none: zeros)
          if (t >= 0) { tqh = p.qh[t]; tf0 = p.qw[t] * Ci + ci0; }
        }
        HALO_TRACE2(i, 0);
        mbar_wait(empty_bar(s), ph ^ 1u);
        HALO_TRACE2(i, 1);
        uint8_t* sptr = smem_gen + s * cfg.stage_bytes;
        const int boxpitch = cfg.halo_h * 256;            // elements per staged box: [halo_h rows][256 flattened (w, c) elements]
        const int nit = (cfg.dbg & 1) ? 0 : 4;
        // 4 batches of 8 warp instructions: eight independent 16-byte loads in flight, then their eight stores (a load -> store
        // chain per row ran at ~100 clk per row, 1.6 us per k block: tools/halo_trace.py, round 2).
        // row r = it*4 + rsub of the 8 x 16 tile: ty = it >> 2, tx = (it & 3)*4 + rsub; source pixel (ty*hs + qh, tx*hs + qw),
        // channel ci0 -> flattened column f -> box f / 256, element f % 256 of row hy
        for (int b8 = 0; b8 < nit; ++b8) {
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int it = b8 * 8 + u;
            const int ty = it >> 2, tx = (it & 3) * 4 + rsub;
            const int f = tx * hs * Ci + tf0, hyy = ty * hs + tqh;
            const int soff = ((f >> 8) * boxpitch + hyy * 256 + (f & 255)) * 2;
            v[u] = (t >= 0) ? *reinterpret_cast<const uint4*>(hptr + soff) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = (b8 * 8 + u) * 4 + rsub;
            *reinterpret_cast<uint4*>(sptr + r * 128 + ((j ^ (r & 7)) << 4)) = v[u];
          }
        }
        HALO_TRACE2(i, 2);
        fence_proxy_async();                           // generic-proxy writes -> visible to the tensor core's async proxy
        HALO_TRACE2(i, 3);
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(s));
        HALO_TRACE2(i, 4);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(hempty_bar(hb));      // this warp has read the halo buffer for the last time
      if (pw == 0 && lane == 0) HALO_TRACE(i, 2);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, cfg.wide ? cfg.bn_sub0 : BN, 0, 0);
      const uint32_t idesc1 = make_idesc(128, cfg.wide ? BN - cfg.bn_sub0 : 16, 0, 0);
      const bool single_acc = cfg.bm2 || cfg.wide;
      if (cfg.b_resident) { mbar_wait(bres_bar, 0); tc_fence_after(); }
      int s = 0; uint32_t ph = 0;
      int i = 0;
      const int MT = cfg.bm2 ? 2 : 1;
      for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x, ++i) {
        // bm2: one accumulator SET (two accumulators side by side); otherwise two alternating accumulators
        const int buf = single_acc ? 0 : (i & 1);
        const uint32_t par = single_acc ? ((uint32_t)i & 1u) : (((uint32_t)i >> 1) & 1u);
        mbar_wait(tempty_bar(buf), par ^ 1u);                            // epilogue has drained this accumulator
        tc_fence_after();
        if (kHalo) HALO_TRACE(i, 3);
        const uint32_t acc = tmem_base + (uint32_t)(buf * cfg.acc_stride);
        int it = 0;
        const int nt_loop = kHalo ? 1 : p.ntaps, nk_loop = kHalo ? cfg.nkb : p.kblocks;     // halo: K is dense, nkb blocks in order
        for (int t = 0; t < nt_loop; ++t) {
          const int kb0 = kHalo ? 0 : p.wt[t] * p.kblocks;
          for (int kb = 0; kb < nk_loop; ++kb, ++it) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            if (i == 0 && it == 0) TC_TRACE(2);
            const uint32_t sa = smem_base + s * cfg.stage_bytes;
            const uint32_t sb = cfg.b_resident ? (bres_base + (uint32_t)(kb0 + kb) * b_block_bytes) : (sa + MT * kABytes);
            const uint64_t da = make_smem_desc(sa, 16, 1024);
            const uint64_t db = make_smem_desc(sb, 16, 1024);
            // halo: K is dense, so the last k block may hold fewer than four 16-element steps (conv1: K = 144 = 64 + 64 + 16,
            // last_deconv's dgrad: 72 = 64 + 8) - the steps beyond ktot multiply zero columns by zero weight rows and are skipped
            const int kmax = kHalo ? min(4, (cfg.ktot - kb * 64 + 15) >> 4) : 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < kmax && !(kHalo && (cfg.dbg & 2))) umma_bf16(acc, da + 2u * k, db + 2u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
            if (cfg.wide) {
              // second N sub-tile: B rows bn_sub0.. of the same stage (row offset = bn_sub0 * 128 B, a multiple of 1024),
              // accumulator columns bn_sub0..BN-1
              const uint64_t db1 = db + (uint64_t)((cfg.bn_sub0 * 128) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(acc + (uint32_t)cfg.bn_sub0, da + 2u * k, db1 + 2u * k, idesc1, (it > 0 || k > 0) ? 1u : 0u);
            }
            if (cfg.bm2) {
              const uint64_t da1 = make_smem_desc(sa + kABytes, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(acc + (uint32_t)cfg.acc_stride, da1 + 2u * k, db + 2u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(s));
            if (++s == STAGES) { s = 0; ph ^= 1u; }
          }
        }
        umma_commit(tfull_bar(buf));
        TC_TRACE(3);
        if (kHalo) HALO_TRACE(i, 4);
      }
    }
  } else if (warp >= 2) {
    // ---- epilogue: warps 2..9 (two groups); warp w may only touch TMEM lanes 32*(w%4) .. +31 ----
    const int grp = (warp - 2) >> 2;
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int ty = row / p.TW, tx = row - ty * p.TW;
    const bool stage_f32 = (p.accumulate != 0) || (p.out.dtype == DC_F32);
    const int es = stage_f32 ? 4 : 2;
    const int pitch = 64 * es + 16;
    uint8_t* stg = smem_gen + stg_off + (size_t)lg * 32 * kV2StagePitchF32;
    const int V = stage_f32 ? 4 : 8;
    const bool raw_copy = !stage_f32 && p.out_vec_ok;
    int i = 0;
    const int MT = cfg.bm2 ? 2 : 1;
    // BatchNorm statistics carried across this CTA's tiles (see the statistics block below)
    double stat_acc0 = 0.0, stat_acc1 = 0.0, stat_acc2 = 0.0, stat_acc3 = 0.0;
    int stat_nt = -1;
    auto stats_flush = [&]() {
      if (p.stats == nullptr || stat_nt < 0) return;
      const int t = (int)threadIdx.x - 64 - 128 * grp;
      const int fn0 = stat_nt * BN, fncols = min(BN, p.out.c - fn0);
      double* dst = p.stats + (size_t)(t >> 6) * p.stats_C + fn0;
      const int cbase = grp * 64 + (t & 63);
      if (cbase < fncols && stat_acc0 != 0.0) atomicAdd(dst + cbase, stat_acc0);
      if (cbase + 128 < fncols && stat_acc1 != 0.0) atomicAdd(dst + cbase + 128, stat_acc1);
      if (cbase + 256 < fncols && stat_acc2 != 0.0) atomicAdd(dst + cbase + 256, stat_acc2);
      if (cbase + 384 < fncols && stat_acc3 != 0.0) atomicAdd(dst + cbase + 384, stat_acc3);
      stat_acc0 = stat_acc1 = stat_acc2 = stat_acc3 = 0.0;
    };
    for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x, ++i) {
     for (int sub = 0; sub < MT; ++sub) {
      const int nt = tile / cfg.n_mtiles;
      int mt = (tile - nt * cfg.n_mtiles) * MT + sub;
      const bool real = mt < cfg.real_mtiles;
      const int tile_x = mt % p.tiles_x; mt /= p.tiles_x;
      const int tile_y = mt % p.tiles_y;
      const int img = real ? mt / p.tiles_y : p.out.n;           // past the last image: the TMA store is clipped away
      const int oy = tile_y * p.TH + ty, ox = tile_x * p.TW + tx;
      const int n0 = nt * BN;
      const bool pix_ok = real && (oy < p.out.h) && (ox < p.out.w);
      const long long base = (long long)img * p.out.sn + (long long)oy * p.out.sh + (long long)ox * p.out.sw;
      const unsigned okmask = __ballot_sync(0xffffffffu, pix_ok);
      const int ncols = min(BN, p.out.c - n0);
      const bool single_acc = cfg.bm2 || cfg.wide;
      const int buf = single_acc ? 0 : (i & 1);
      if (sub == 0) {
        mbar_wait(tfull_bar(buf), single_acc ? ((uint32_t)i & 1u) : (((uint32_t)i >> 1) & 1u));
        tc_fence_after();
        if (threadIdx.x == 64) TC_TRACE(4);
        if (threadIdx.x == 192) TC_TRACE(12);
        if (kHalo && threadIdx.x == 64) HALO_TRACE(i, 5);
      }
      const uint32_t acc = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((cfg.bm2 ? sub : buf) * cfg.acc_stride);
      if (cfg.tma_store) {
        // ---- fast path: 128 rows x 64 bf16 columns per chunk, 128B-swizzled staging (two buffers), one TMA store
        //      (or TMA reduce-add for gradient accumulation) per chunk issued by one elected thread ----
        // group g drains chunks g, g+2, ... through its own 16 KB staging buffer, elected thread and named barrier
        const bool elected = (threadIdx.x == 64 + 128 * grp);
        const uint32_t stg_s = smem_base + stg_off + (uint32_t)grp * 16384u;
        if (grp * 64 >= ncols) {            // no chunk for this group in this N tile: nothing to drain
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
#pragma unroll 1
        for (int c0 = grp * 64; c0 < ncols; c0 += 128) {
          uint32_t v[2][32];
          const bool two = (c0 + 32 < ncols);
          tmem_ld32(acc + (uint32_t)c0, v[0]);
          if (two) tmem_ld32(acc + (uint32_t)(c0 + 32), v[1]);
          if (elected) tma_store_wait_read<0>();        // this group's previous store has read the staging buffer
          tmem_ld_wait();
          if (c0 + 128 >= ncols) {                      // this group's last chunk of the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
            if (kHalo && threadIdx.x == 64) HALO_TRACE(i, 6);
          }
          // eval-mode BatchNorm folded into the epilogue: the group's first 64 threads derive scale / shift of the chunk's 64
          // columns from the running statistics (same arithmetic as the eval branch of bn_apply) into the group's 2 KB scratch
          float* scoef = reinterpret_cast<float*>(smem_gen + stg_off + 32768 + grp * 2048);      // [scale 64 | shift 64]
          if (p.aff_gamma != nullptr) {
            const int t = (int)threadIdx.x - 64 - 128 * grp;
            if (t < 64) {
              const int c = n0 + c0 + t;
              float sc = 0.f, sh = 0.f;
              if (c0 + t < ncols) {
                const float inv = inv_sqrt_f32(p.aff_var[c] + p.aff_eps);
                sc = p.aff_gamma[c] * inv;
                sh = p.aff_beta[c] - p.aff_mean[c] * sc;
                if (p.bias) sh = fmaf(p.bias[c], sc, sh);
              }
              scoef[t] = sc;
              scoef[64 + t] = sh;
            }
          }
          epi_group_sync(grp);
          const uint32_t sbuf = stg_s + (uint32_t)row * 128u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
              if (h == 0 || two) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  f[j] = __uint_as_float(v[h][8 * q + j]);
                  if (p.aff_gamma != nullptr) {
                    f[j] = fmaf(f[j], scoef[32 * h + 8 * q + j], scoef[64 + 32 * h + 8 * q + j]);
                    if (p.aff_relu) f[j] = fmaxf(f[j], 0.f);
                  } else if (p.bias && c0 + 32 * h + 8 * q + j < ncols) {
                    f[j] += p.bias[n0 + c0 + 32 * h + 8 * q + j];
                  }
                }
                __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]), b3 = __floats2bfloat162_rn(f[6], f[7]);
                w0 = *reinterpret_cast<uint32_t*>(&b0); w1 = *reinterpret_cast<uint32_t*>(&b1);
                w2 = *reinterpret_cast<uint32_t*>(&b2); w3 = *reinterpret_cast<uint32_t*>(&b3);
              }
              if (p.stats != nullptr && !pix_ok) { w0 = 0; w1 = 0; w2 = 0; w3 = 0; }   // clipped by the store; must not count
              const uint32_t j16 = (uint32_t)(4 * h + q);
              const uint32_t dst = sbuf + ((j16 ^ ((uint32_t)row & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
            }
          }
          fence_proxy_async();
          epi_group_sync(grp);
          if (elected && c0 == 0) TC_TRACE(8);
          if (elected && !(kHalo && (cfg.dbg & 4))) {
            const uint32_t src = stg_s;
            if (p.accumulate) tma_reduce_add_4d(&maps.c, src, n0 + c0, tile_x * p.TW, tile_y * p.TH, img);
            else tma_store_4d(&maps.c, src, n0 + c0, tile_x * p.TW, tile_y * p.TH, img);
            tma_store_commit();
            if (c0 == 0) TC_TRACE(9);
          }
          if (p.stats != nullptr) {
            // ---- BatchNorm statistics of this 128 x 64 chunk, from the staged (bf16-rounded) values: warp = 32 rows, lane =
            //      one column pair (conflict-free 4-byte reads of the swizzled rows); the four row groups are combined through
            //      2 KB of shared memory and every thread issues ONE fp64 atomic (64 columns x {sum, sum of squares}) ----
            const uint32_t sb = stg_s;
            const uint32_t cj16 = (uint32_t)lane >> 2, coff = ((uint32_t)lane & 3u) * 4u;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const uint32_t rr = (uint32_t)(lg * 32 + r);
              uint32_t wv;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(sb + rr * 128u + ((cj16 ^ (rr & 7u)) << 4) + coff));
              const float a = __uint_as_float(wv << 16), b = __uint_as_float(wv & 0xffff0000u);
              s0 += a; s1 += b;
              q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
            }
            float* sred = reinterpret_cast<float*>(smem_gen + stg_off + 32768 + grp * 2048);      // [4 row groups][sum 64 | sumsq 64]
            sred[lg * 128 + 2 * lane] = s0; sred[lg * 128 + 2 * lane + 1] = s1;
            sred[lg * 128 + 64 + 2 * lane] = q0; sred[lg * 128 + 64 + 2 * lane + 1] = q1;
            epi_group_sync(grp);
            const int t = (int)threadIdx.x - 64 - 128 * grp;     // 0..127: statistic (t >> 6), column (t & 63)
            const float tot = sred[t] + sred[128 + t] + sred[256 + t] + sred[384 + t];
            // The chunk totals of ALL tiles this persistent CTA processes in one N tile are summed in registers (fp64) and
            // flushed with one atomic per (statistic, column) when the N tile changes / at the end: the large-M layers run
            // ~23 tiles per CTA, and 3456 tiles x 128 fp64 atomics on the same 64-128 addresses serialised in the L2
            // (conv1 / conv2: 62 / 93 us of which the tiles' own work is ~25 us).
            if (nt != stat_nt) { stats_flush(); stat_nt = nt; }
            const int slot = (c0 - grp * 64) >> 7;
            if (c0 + (t & 63) < ncols) {
              if (slot == 0) stat_acc0 += (double)tot;
              else if (slot == 1) stat_acc1 += (double)tot;
              else if (slot == 2) stat_acc2 += (double)tot;
              else stat_acc3 += (double)tot;                     // wide tiles of 512 columns: four chunks per group
            }
            if (elected && c0 == 0) TC_TRACE(10);
          }
        }
        if (kHalo && threadIdx.x == 64) HALO_TRACE(i, 7);
        continue;
      }
      if (grp == 1) {                       // generic (strided / fp32 / accumulating) epilogue: group A alone, per-warp staging
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(buf));
        continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < ncols; c0 += 64) {
        uint32_t v[2][32];
        const bool two = (c0 + 32 < ncols);
        tmem_ld32(acc + (uint32_t)c0, v[0]);
        if (two) tmem_ld32(acc + (uint32_t)(c0 + 32), v[1]);
        tmem_ld_wait();
        if (c0 + 64 >= ncols) {            // last chunk of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !two) break;
          const int cb = c0 + 32 * h;
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cb + j < ncols) v[h][j] = __float_as_uint(__uint_as_float(v[h][j]) + p.bias[n0 + cb + j]);
          }
          uint8_t* rp = stg + (size_t)lane * pitch + (size_t)(32 * h) * es;
          if (stage_f32) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(rp + q * 16) = make_uint4(v[h][4 * q], v[h][4 * q + 1], v[h][4 * q + 2], v[h][4 * q + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 r;
              __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(v[h][8 * q + 0]), __uint_as_float(v[h][8 * q + 1]));
              __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(v[h][8 * q + 2]), __uint_as_float(v[h][8 * q + 3]));
              __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[h][8 * q + 4]), __uint_as_float(v[h][8 * q + 5]));
              __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(v[h][8 * q + 6]), __uint_as_float(v[h][8 * q + 7]));
              r.x = *reinterpret_cast<uint32_t*>(&b0); r.y = *reinterpret_cast<uint32_t*>(&b1);
              r.z = *reinterpret_cast<uint32_t*>(&b2); r.w = *reinterpret_cast<uint32_t*>(&b3);
              *reinterpret_cast<uint4*>(rp + q * 16) = r;
            }
          }
        }
        __syncwarp();
        const int ccols = min(64, ncols - c0);
        int lpr = 1;
        while (lpr < 32 && lpr * V < ccols) lpr <<= 1;
        const int rpi = 32 / lpr;
        const int my_sub = lane / lpr, my_l = lane - my_sub * lpr;
#pragma unroll 2
        for (int r0 = 0; r0 < 32; r0 += rpi) {
          const int r = r0 + my_sub;
          const long long rbase = __shfl_sync(0xffffffffu, base, r);
          const bool ok = (okmask >> r) & 1u;
          const uint8_t* rp = stg + (size_t)r * pitch;
          for (int col = my_l * V; col < ccols; col += lpr * V) {
            if (!ok) continue;
            const int co = n0 + c0 + col;
            if (raw_copy && col + 8 <= ccols) {
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out.p) + rbase + co) =
                  *reinterpret_cast<const uint4*>(rp + (size_t)col * 2);
              continue;
            }
            float f[8];
            if (stage_f32) {
              const float4 t4 = *reinterpret_cast<const float4*>(rp + (size_t)col * 4);
              f[0] = t4.x; f[1] = t4.y; f[2] = t4.z; f[3] = t4.w;
            } else {
              const uint4 t4 = *reinterpret_cast<const uint4*>(rp + (size_t)col * 2);
              const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w4[j] << 16); f[2 * j + 1] = __uint_as_float(w4[j] & 0xffff0000u); }
            }
            if (stage_f32 && p.out_vec_ok && col + 4 <= ccols) {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out.p) + rbase + co;
              if (p.accumulate) {
                const uint2 old = *reinterpret_cast<const uint2*>(op);
                f[0] += __uint_as_float(old.x << 16); f[1] += __uint_as_float(old.x & 0xffff0000u);
                f[2] += __uint_as_float(old.y << 16); f[3] += __uint_as_float(old.y & 0xffff0000u);
              }
              __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
              uint2 o;
              o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
              *reinterpret_cast<uint2*>(op) = o;
            } else if (p.out.dtype == DC_F32 && p.out.sc == 1 && col + 4 <= ccols && ((rbase + out_col_off(p.out, co)) & 3) == 0 &&
                       ((reinterpret_cast<uintptr_t>(p.out.p) & 15) == 0)) {
              float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out.p) + rbase + out_col_off(p.out, co));
              float4 o = make_float4(f[0], f[1], f[2], f[3]);
              if (p.accumulate) { const float4 old = *q; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
              *q = o;
            } else {
              for (int j = 0; j < V && col + j < ccols; ++j) {
                const long long off = rbase + out_col_off(p.out, co + j);
                float val = f[j];
                if (p.out.dtype == DC_F32) {
                  float* q = reinterpret_cast<float*>(p.out.p) + off;
                  if (p.accumulate) val += *q;
                  *q = val;
                } else {
                  __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(p.out.p) + off;
                  if (p.accumulate) val += __bfloat162float(*q);
                  *q = __float2bfloat16_rn(val);
                }
              }
            }
          }
        }
        __syncwarp();       // staging rows are reused by the next chunk
      }
     }
    }
    stats_flush();
  }
  if (threadIdx.x == 64) TC_TRACE(5);
  if (threadIdx.x == 192) TC_TRACE(11);
  if (cfg.tma_store && (threadIdx.x == 64 || threadIdx.x == 192)) tma_store_wait_all();     // staging must stay valid until the last store has read it
  __syncwarp();        // warps 0/1: the single-lane role (producer / MMA issuer) rejoins its warp, so every warp arrives at the
                       // teardown barrier exactly once and the TMEM release below is ordered after the epilogue's last read
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_TRACE(6);
  if (warp == 1) tmem_dealloc(tmem_base, cfg.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// wgrad: D[co (128)][ci (BNW)] += sum over pixel tiles of dY^T X_t ; MN-major operands.
// ------------------------------------------------------------------------------------------------
struct TcWgradParams {
  int ntaps;
  int map_id[DC_MAX_TAPS];
  int qh[DC_MAX_TAPS], qw[DC_MAX_TAPS];
  int wt[DC_MAX_TAPS];
  int TH, TW, tiles_x, tiles_y, n_img;
  int mtiles_total, mtiles_per_split;
  int n_ci_tiles;
  int Co, Ci;
  int vec_red;          // G rows are 16-byte aligned and Ci % 4 == 0: use red.global.add.v4.f32
  int tma_red;          // ... and the fp32 gradient has a tensor map (maps.c): the epilogue stages 128 x 32 fp32 chunks in shared
                        // memory and hands them to TMA reduce-add (bulk L2 reductions instead of 8192 red instructions per CTA)
  int det_wtaps;        // > 0: deterministic two-stage reduction (dc_conv_wgrad_tc_det).  G is a workspace [splits][wtaps][Co][Ci] and
                        // pixel split z STORES its partial tile into slice z (one writer per element, no reduction in this kernel)
  float* G;
};

template <int BNW> struct WgradCfg {
  static constexpr int kABytes = 2 * 16384;            // 128 co = 2 boxes of [128 px][64 ch]
  static constexpr int kBBytes = (BNW / 64) * 16384;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kStageBytes > 65536) ? 2 : 3;
  static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
};

template <int BNW>
__global__ void __launch_bounds__(kTcThreads) conv_wgrad_tc_kernel(const __grid_constant__ TcMaps maps, const TcWgradParams p) {
  using Cfg = WgradCfg<BNW>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int co0 = (blockIdx.x / p.n_ci_tiles) * 128;
  const int ci0 = (blockIdx.x % p.n_ci_tiles) * BNW;
  const int t = blockIdx.y;
  const int mt_begin = blockIdx.z * p.mtiles_per_split;
  const int mt_end = min(p.mtiles_total, mt_begin + p.mtiles_per_split);
  const int n_iter = mt_end - mt_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BNW);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.a[p.map_id[t]]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_sync();     // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (n_iter > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const CUtensorMap* am = &maps.a[p.map_id[t]];
        int s = 0; uint32_t ph = 0;
        for (int mt = mt_begin; mt < mt_end; ++mt) {
          int r = mt;
          const int tile_x = r % p.tiles_x; r /= p.tiles_x;
          const int tile_y = r % p.tiles_y;
          const int img = r / p.tiles_y;
          const int x0 = tile_x * p.TW, y0 = tile_y * p.TH;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          // A = dY tile: two 64-channel boxes (co0, co0+64)
          tma_load_4d(&maps.b, full_bar(s), sa, co0, x0, y0, img);
          tma_load_4d(&maps.b, full_bar(s), sa + 16384, co0 + 64, x0, y0, img);
          // B = gathered X tile: BNW/64 boxes
#pragma unroll
          for (int j = 0; j < BNW / 64; ++j)
            tma_load_4d(am, full_bar(s), sa + Cfg::kABytes + j * 16384, ci0 + j * 64, x0 + p.qw[t], y0 + p.qh[t], img);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(128, BNW, 1, 1);
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < n_iter; ++it) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          // MN-major SW128: LBO = stride between 64-element MN blocks (one box = 16 KB), SBO = 8 k-rows = 1024 B
          const uint64_t da = make_smem_desc(sa, 16384, 1024);
          const uint64_t db = make_smem_desc(sa + Cfg::kABytes, 16384, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // advance 16 pixels (K) = 16 rows of 128 B = 2048 B -> +128 in the (>>4) address field
            umma_bf16(tmem_base, da + 128u * k, db + 128u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit(tmem_full_bar);
      }
    } else {
      const int lg = warp & 3;
      const int co = co0 + lg * 32 + lane;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const int wslice = p.det_wtaps > 0 ? (int)blockIdx.z * p.det_wtaps + p.wt[t] : p.wt[t];
      float* Grow = p.G + ((size_t)wslice * p.Co + (size_t)(co < p.Co ? co : 0)) * p.Ci;
      if (p.tma_red) {
        // all MMAs have completed (tmem_full), so every pipeline stage has been consumed: stage memory becomes two 16 KB
        // staging buffers of 128 rows x 128 B (32 fp32), 128B-swizzled like the tensor map of the gradient
        const bool elected = (threadIdx.x == 64);
        const int row = lg * 32 + lane;
        int chunk = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BNW; c0 += 32, ++chunk) {
          if (ci0 + c0 >= p.Ci) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
          if (elected) tma_store_wait_read<1>();       // the reduction that used this buffer two chunks ago has read it
          tmem_ld_wait();
          epi_bar_sync();
          const uint32_t sbuf = smem_base + (uint32_t)(chunk & 1) * 16384u + (uint32_t)row * 128u;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t dst = sbuf + (((uint32_t)q ^ ((uint32_t)row & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[4 * q]), "r"(v[4 * q + 1]), "r"(v[4 * q + 2]),
                         "r"(v[4 * q + 3]) : "memory");
          }
          fence_proxy_async();
          epi_bar_sync();
          if (elected) {
            if (p.det_wtaps > 0) tma_store_3d(&maps.c, smem_base + (uint32_t)(chunk & 1) * 16384u, ci0 + c0, co0, wslice);
            else tma_reduce_add_3d(&maps.c, smem_base + (uint32_t)(chunk & 1) * 16384u, ci0 + c0, co0, wslice);
            tma_store_commit();
          }
        }
        if (elected) tma_store_wait_all();             // shared memory must stay valid until the last reduction has read it
      } else
#pragma unroll 1
      for (int c0 = 0; c0 < BNW; c0 += 32) {
        if (ci0 + c0 >= p.Ci) break;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (co < p.Co && p.det_wtaps > 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int ci = ci0 + c0 + j;
            if (ci < p.Ci) Grow[ci] = __uint_as_float(v[j]);
          }
        } else if (co < p.Co) {
          if (p.vec_red) {
            // 16-byte reductions: 4 consecutive ci of this lane's co row per instruction (Ci % 4 == 0, G 16-byte aligned)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int ci = ci0 + c0 + 4 * q;
              if (ci < p.Ci) red_add_v4(Grow + ci, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int ci = ci0 + c0 + j;
              if (ci < p.Ci) atomicAdd(Grow + ci, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  __syncwarp();        // warps 0/1: the single-lane role (producer / MMA issuer) rejoins its warp, so every warp arrives at the
                       // teardown barrier exactly once and the TMEM release below is ordered after the epilogue's last read
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BNW);
}

// ------------------------------------------------------------------------------------------------
// wgrad, HALO mode: multi-tap weight gradients whose gathered operand has few channels (conv1 16 -> 32 k3 s2, conv2 32 -> 64 k3,
// last_deconv's 256 -> 3(8) k3 s2: DX:144,148,374).  conv_wgrad_tc_kernel runs ONE tap per CTA with a 64-channel TMA box per
// operand: 9 CTAs re-read the same dY tile, three quarters of every gathered box are zero fill, and each 128x64x16 MMA costs
// ~165 clk of issue for 16 useful columns (145 / 92 / 260 us for 13 / 13 / 39 us of HBM time, profiles/r02_roofline_table.md).
// Here a CTA owns ALL taps of a tap group: per 128-pixel tile it loads the dY tile once (A, MN-major, 128B swizzle, as before) and
// the tile's input REGION once (un-swizzled boxes, halo included, zero fill = padding); four builder warps copy the region into
// ONE MN-major B operand whose N index is (tap, channel) - 16-byte chunks, 128B-swizzled rows of 64 N elements - so that a tile
// takes 8 MMAs of N = taps x Ci (<= 192) instead of 8 per tap.  D[co][tap * Ci + ci] accumulates in TMEM over the CTA's pixel
// split and is added to G[wt[tap]][co][ci] at the end (red.global.add.v4.f32, or plain stores into the split's workspace slice
// in deterministic mode).
// ------------------------------------------------------------------------------------------------
struct TcWgradHaloParams {
  int tiles_x, tiles_y, n_img;
  int mtiles_total, mtiles_per_split;
  int Co, Ci;                       // Ci = gathered channels (multiple of 8, <= 64)
  int s;                            // gather stride 1 | 2
  int halo_w, halo_h, halo_nbox, halo_bytes, halo_stride;
  int halo_x0, halo_y0;             // smallest tap offset: input coordinate of the region's first pixel relative to stride * tile origin
  int a_boxes, a_bytes;             // 64-channel dY boxes per tile (1 when Co <= 64), bytes per A stage
  int sa, sr;                       // A stages, region stages
  int nblk;                         // 64-column blocks of the B operand (ceil(npad / 64))
  int ngroups;
  int g_first[3], g_count[3], g_npad[3];     // taps of tap group g: [g_first, g_first + g_count), N padded to a multiple of 16
  int qh[DC_MAX_TAPS], qw[DC_MAX_TAPS], wt[DC_MAX_TAPS];
  int det_wtaps;
  int tmem_cols;
  float* G;
};
constexpr int kWgradHaloThreads = 448;    // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue, warps 6..13 B builders

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(kWgradHaloThreads, 1) conv_wgrad_tc_halo_kernel(const __grid_constant__ TcMaps maps, const TcWgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [A stages][region stages][2 B buffers][barriers]
  const uint32_t a_off = 0u, r_off = (uint32_t)(p.sa * p.a_bytes), b_off = r_off + (uint32_t)(p.sr * p.halo_stride);
  const uint32_t b_bytes = (uint32_t)p.nblk * 16384u;
  const uint32_t bar_base = smem_base + b_off + 2u * b_bytes;
  auto afull = [&](int i) { return bar_base + 8u * i; };
  auto aempty = [&](int i) { return bar_base + 8u * (4 + i); };
  auto rfull = [&](int i) { return bar_base + 8u * (8 + i); };
  auto rempty = [&](int i) { return bar_base + 8u * (12 + i); };
  auto bfull = [&](int i) { return bar_base + 8u * (16 + i); };
  auto bempty = [&](int i) { return bar_base + 8u * (18 + i); };
  const uint32_t tmem_full_bar = bar_base + 8u * 20;
  const uint32_t tmem_slot = bar_base + 8u * 21;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * 128;
  const int g = blockIdx.y;
  const int t0 = p.g_first[g], nt = p.g_count[g], npad = p.g_npad[g];
  const int mt_begin = blockIdx.z * p.mtiles_per_split;
  const int mt_end = min(p.mtiles_total, mt_begin + p.mtiles_per_split);
  const int n_iter = mt_end - mt_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(afull(i), 1); mbar_init(aempty(i), 1); mbar_init(rfull(i), 1); mbar_init(rempty(i), 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(bfull(i), 8); mbar_init(bempty(i), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.a[0]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_sync();

  if (n_iter > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < n_iter; ++it) {
          int r = mt_begin + it;
          const int tile_x = r % p.tiles_x; r /= p.tiles_x;
          const int tile_y = r % p.tiles_y;
          const int img = r / p.tiles_y;
          const int sa = it % p.sa, sr = it % p.sr;
          // A = dY tile: one or two 64-channel boxes of 128 pixels (8 rows x 16 columns)
          mbar_wait(aempty(sa), (((uint32_t)(it / p.sa)) & 1u) ^ 1u);
          mbar_expect_tx(afull(sa), (uint32_t)p.a_bytes);
          for (int b = 0; b < p.a_boxes; ++b)
            tma_load_4d(&maps.b, afull(sa), smem_base + a_off + (uint32_t)(sa * p.a_bytes + b * 16384), co0 + b * 64, tile_x * kHaloTW,
                        tile_y * kHaloTH, img);
          // the tile's input region: halo_nbox boxes of {256 flattened (w, c) elements, halo_h rows}
          mbar_wait(rempty(sr), (((uint32_t)(it / p.sr)) & 1u) ^ 1u);
          mbar_expect_tx(rfull(sr), (uint32_t)p.halo_bytes);
          const int fx = (tile_x * kHaloTW * p.s + p.halo_x0) * p.Ci;
          const int hy = tile_y * kHaloTH * p.s + p.halo_y0;
          for (int b = 0; b < p.halo_nbox; ++b)
            tma_load_3d(&maps.a[0], rfull(sr), smem_base + r_off + (uint32_t)(sr * p.halo_stride + b * p.halo_h * 512), fx + b * 256, hy, img);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc(128, npad, 1, 1);
        for (int it = 0; it < n_iter; ++it) {
          const int sa = it % p.sa, bb = it & 1;
          mbar_wait(afull(sa), ((uint32_t)(it / p.sa)) & 1u);
          mbar_wait(bfull(bb), ((uint32_t)(it >> 1)) & 1u);
          tc_fence_after();
          // MN-major SW128 operands: LBO = stride between 64-element MN blocks (16 KB), SBO = 8 k-rows = 1024 B
          const uint64_t da = make_smem_desc(smem_base + a_off + (uint32_t)(sa * p.a_bytes), 16384, 1024);
          const uint64_t db = make_smem_desc(smem_base + b_off + (uint32_t)bb * b_bytes, 16384, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)      // 16 pixels (K) per instruction = 16 rows of 128 B = +128 in the (>>4) address field
            umma_bf16(tmem_base, da + 128u * k, db + 128u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(aempty(sa));
          umma_commit(bempty(bb));
        }
        umma_commit(tmem_full_bar);
      }
    } else if (warp >= 6) {
      // ---- B builders: row = tile pixel (ty, tx) = K index, N index = tap * Ci + ci.  Chunk q of a row (q = tap * Ci / 8 + c) is
      //      the 16 bytes of channels 8c .. 8c+7 of the pixel the tap gathers: region -> block (q / 8), logical chunk q % 8 of the
      //      row, physical chunk = logical ^ (row & 7) (the 128B swizzle the MN-major descriptor expects).
      //      lane = (rsub, j): a warp instruction moves chunks 8*qb + j (j = 0..7) of 4 rows, so a quarter-warp reads one row's
      //      consecutive chunks - the taps (kh, 0..2) of a pixel are adjacent pixels, i.e. one contiguous run of the region -
      //      and writes 8 distinct 16-byte columns of one 128-byte line: no bank conflicts on either side (one thread per row
      //      with a tap-major walk cost 4-way conflicts on every load: 1600 of the ~3000 clk per tile, round 2).  Warp bw owns rows
      //      16*bw .. +15 (4 passes of 4 rows); all source / destination offsets are tile-independent and live in registers. ----
      const int bw = warp - 6;
      const int rsub = lane >> 3, j = lane & 7;
      const int c16 = p.Ci >> 3;
      const int boxpitch = p.halo_h * 256;
      const int nchunks = nt * c16;
      const int nqb = ((npad >> 3) + 7) >> 3;                          // batches of 8 chunks, the zero padding chunk included
      int soff[3][4];                                                   // < 0: nothing to move; -2: write zeros (padding chunk)
      uint32_t doff[3][4];
#pragma unroll
      for (int qb = 0; qb < 3; ++qb)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = qb * 8 + j;
          const int row = bw * 16 + u * 4 + rsub;
          const int ty = row >> 4, tx = row & 15;
          soff[qb][u] = -1;
          doff[qb][u] = (uint32_t)row * 128u + (uint32_t)(q >> 3) * 16384u + (uint32_t)(((q & 7) ^ (row & 7)) << 4);
          if (q < nchunks) {
            const int t = q / c16, c = q - t * c16;
            const int hyy = ty * p.s + p.qh[t0 + t];
            const int f = (tx * p.s + p.qw[t0 + t]) * p.Ci + c * 8;
            soff[qb][u] = ((f >> 8) * boxpitch + hyy * 256 + (f & 255)) * 2;
          } else if (q == nchunks && q * 8 < npad) {
            soff[qb][u] = -2;
          }
        }
      for (int it = 0; it < n_iter; ++it) {
        const int sr = it % p.sr, bb = it & 1;
        mbar_wait(rfull(sr), ((uint32_t)(it / p.sr)) & 1u);
        mbar_wait(bempty(bb), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
        const uint8_t* hptr = smem_gen + r_off + (uint32_t)(sr * p.halo_stride);
        uint8_t* bptr = smem_gen + b_off + (uint32_t)bb * b_bytes;
#pragma unroll
        for (int qb = 0; qb < 3; ++qb) {
          if (qb < nqb) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              v[u] = soff[qb][u] >= 0 ? *reinterpret_cast<const uint4*>(hptr + soff[qb][u]) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (soff[qb][u] != -1) *reinterpret_cast<uint4*>(bptr + doff[qb][u]) = v[u];
          }
        }
        fence_proxy_async();                          // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) { mbar_arrive(bfull(bb)); mbar_arrive(rempty(sr)); }
      }
    } else {
      // ---- epilogue: lane = co row; tap by tap, 8 columns at a time ----
      const int lg = warp & 3;
      const int co = co0 + lg * 32 + lane;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      for (int t = 0; t < nt; ++t) {
        const int wt = p.wt[t0 + t];
        const int wslice = p.det_wtaps > 0 ? (int)blockIdx.z * p.det_wtaps + wt : wt;
        float* Grow = p.G + ((size_t)wslice * p.Co + (size_t)(co < p.Co ? co : 0)) * p.Ci;
        for (int c0 = 0; c0 < p.Ci; c0 += 8) {
          uint32_t v[8];
          tmem_ld8(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(t * p.Ci + c0), v);
          tmem_ld_wait();
          if (co < p.Co) {
            if (p.det_wtaps > 0) {
              *reinterpret_cast<float4*>(Grow + c0) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
              *reinterpret_cast<float4*>(Grow + c0 + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
            } else {
              red_add_v4(Grow + c0, v[0], v[1], v[2], v[3]);
              red_add_v4(Grow + c0 + 4, v[4], v[5], v[6], v[7]);
            }
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

// 4-D bf16 NHWC map over a (sub-)grid: dims {C, W, H, N}, box {64, TW, TH, 1}, 128B swizzle, zero OOB fill.
static int encode_act_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, long long sw, long long sh, long long sn,
                          int TW, int TH, const char* what) {
  PFN_encodeTiled enc = get_encode();
  DC_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled unavailable", what);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DC_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(act) failed with %d (C=%d W=%d H=%d N=%d sw=%lld sh=%lld sn=%lld)", what,
             (int)r, C, W, H, N, sw, sh, sn);
  return 0;
}

static int encode_weight_map(CUtensorMap* m, const void* ptr, long long Ktot, int Co, int BN, const char* what) {
  PFN_encodeTiled enc = get_encode();
  DC_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled unavailable", what);
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Co};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BN};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DC_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(weights) failed with %d (Ktot=%lld Co=%d)", what, (int)r, Ktot, Co);
  return 0;
}

static void pick_tile(int H, int W, int& TH, int& TW) {
  long long best = -1;
  for (int tw = 128; tw >= 1; tw >>= 1) {
    int th = 128 / tw;
    long long tiles = (long long)ceil_div(W, tw) * ceil_div(H, th);
    if (best < 0 || tiles < best) { best = tiles; TW = tw; TH = th; }
  }
}

static inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

static bool tc_view_ok(const dc_view& v) {
  return view_ok(v) && v.dtype == DC_BF16 && v.sc == 1 && (v.c % 8 == 0) && (v.sw % 8 == 0) && (v.sh % 8 == 0) && (v.sn % 8 == 0) &&
         ((reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0);
}

// Build the parity maps of the gathered operand and the per-tap (map, offset) table.
template <typename P>
static int build_gather(const char* what, const dc_conv_desc* d, const dc_view& in, int TH, int TW, TcMaps& maps, P& p) {
  const int s_h = d->stride_h, s_w = d->stride_w;
  DC_REQUIRE(s_h <= 2 && s_w <= 2, "%s: stride > 2 unsupported", what);
  bool used[4] = {false, false, false, false};
  for (int t = 0; t < d->ntaps; ++t) {
    int qh = floor_div(d->dh[t], s_h), rh = d->dh[t] - qh * s_h;
    int qw = floor_div(d->dw[t], s_w), rw = d->dw[t] - qw * s_w;
    int id = rh * 2 + rw;
    p.map_id[t] = id; p.qh[t] = qh; p.qw[t] = qw; p.wt[t] = d->wt[t];
    used[id] = true;
  }
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(in.ptr);
  int first = -1;
  for (int id = 0; id < 4; ++id) {
    if (!used[id]) continue;
    int rh = id >> 1, rw = id & 1;
    int Hs = (in.h - rh + s_h - 1) / s_h, Ws = (in.w - rw + s_w - 1) / s_w;
    DC_REQUIRE(Hs > 0 && Ws > 0, "%s: empty parity sub-grid", what);
    if (int r = encode_act_map(&maps.a[id], base + rh * in.sh + rw * in.sw, in.c, Ws, Hs, in.n, in.sw * s_w, in.sh * s_h, in.sn, TW, TH, what))
      return r;
    if (first < 0) first = id;
  }
  for (int id = 0; id < 4; ++id)
    if (!used[id]) maps.a[id] = maps.a[first];
  return 0;
}

static inline int round_up_i(int a, int b) { return (a + b - 1) / b * b; }

// Tile configuration: the widest N tile that still gives every SM a CTA (the kernel is L2-bandwidth bound, so
// wide tiles minimise operand re-reads), then as many pipeline stages as fit (two CTAs per SM when possible so
// that one CTA's epilogue overlaps the other's main loop).
static TcFpropCfg pick_fprop_cfg(int mtiles, int Co) {
  TcFpropCfg c;
  int BN = std::min(256, round_up_i(Co, 16));
  int ntiles = ceil_div(Co, BN);
  while ((long long)mtiles * ntiles < kNumSMs && BN > 64) {
    ++ntiles;
    int nb = round_up_i(ceil_div(Co, ntiles), 16);
    if (nb >= BN) nb = BN - 16;
    BN = std::max(nb, 64);
    ntiles = ceil_div(Co, BN);
  }
  c.BN = BN;
  c.stage_bytes = kABytes + BN * 128;
  const int two_cta_budget = 110 * 1024, one_cta_budget = 200 * 1024;
  if (3 * c.stage_bytes <= two_cta_budget) c.stages = std::min(6, two_cta_budget / c.stage_bytes);
  else c.stages = std::max(2, std::min(6, one_cta_budget / c.stage_bytes));
  const int staging = 4 * 32 * (round_up_i(BN, 32) * 4 + 16);
  while (c.stages * c.stage_bytes < staging) ++c.stages;
  c.tmem_cols = 32;
  while (c.tmem_cols < BN) c.tmem_cols <<= 1;
  c.smem_bytes = c.stages * c.stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
  return c;
}

static int launch_fprop(const TcMaps& maps, const TcFpropParams& p, const TcFpropCfg& cfg, int mtiles, int ntiles, cudaStream_t st) {
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail((int)e, "dc_conv_gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  dim3 grid(mtiles, ntiles, 1);
  launch_k(conv_gemm_tc_kernel, grid, dim3(kTcThreads), (size_t)cfg.smem_bytes, st, maps, p, cfg);
  return launch_status("dc_conv_gemm_tc");
}

// Tile configuration of the persistent kernel.  Cost model per CTA: rounds * (bytes pulled from L2 per k-block), since
// these GEMMs are bound by the L2->SM path (~42 B/clk/SM with all SMs pulling), not by the tensor pipe.
static TcV2Cfg pick_v2_cfg(int mtiles, int Co, int kblocks, int wtaps, bool tma_store) {
  TcV2Cfg c = TcV2Cfg();
  const int co16 = round_up_i(Co, 16);
  int best_bn = std::min(256, co16), best_bm2 = 0;
  double best_cost = 1e30;
  const double kiters = (double)kblocks * wtaps;
  // Cost model (cycles): these GEMMs are bound by the L2->SM path, ~6300 B/clk for the whole chip and ~64 B/clk for
  // one SM; add a fixed per-tile epilogue/drain cost.
  static int bm2_enabled = -1;
  if (bm2_enabled < 0) { const char* e = getenv("DEEPCAM_B200_TC_BM2"); bm2_enabled = (e && e[0] == '1') ? 1 : 0; }   // opt-in: measured slower in the full step (no epilogue overlap)
  static int wide_enabled = -1;
  if (wide_enabled < 0) { const char* e = getenv("DEEPCAM_B200_TC_WIDE"); wide_enabled = (e && e[0] == '0') ? 0 : 1; }
  for (int bm2 = 0; bm2 <= bm2_enabled; ++bm2) {
    if (bm2 && !tma_store) break;
    for (int nt = 1; nt <= 16; ++nt) {
      int bn = round_up_i(ceil_div(co16, nt), 16);
      // the TMA-store epilogue writes 64-column chunks: an N tile that is not the last one must be a multiple of 64
      if (tma_store && nt > 1) bn = round_up_i(bn, 64);
      // N tiles of 257..512 columns ("wide": two MMAs per k step, one accumulator): only as a one-round configuration with
      // the TMA-store epilogue, two B boxes per stage and at least two 2-box stages in shared memory
      const bool wide = bn > 256;
      if (wide && (!wide_enabled || bm2 || !tma_store || bn > 512 || bn % 64 != 0 ||
                   (long long)mtiles * ceil_div(Co, bn) > kNumSMs || 2 * (kABytes + bn * 128) > 227 * 1024 - 1280))
        continue;
      if (bn < 64 && co16 >= 64) break;
      const int ntiles = ceil_div(Co, bn);
      const int msup = bm2 ? ceil_div(mtiles, 2) : mtiles;
      const long long tiles = (long long)msup * ntiles;
      if (bm2 && tiles > kNumSMs) continue;              // 256-row tiles only when one round covers the problem
      const long long rounds = ceil_div64(tiles, kNumSMs);
      const double tile_bytes = kiters * ((bm2 ? 2 : 1) * 16384.0 + 128.0 * bn);
      const double chip = (double)tiles * tile_bytes / 6300.0;
      const double sm = (double)rounds * tile_bytes / 64.0;
      const double cost = std::max(chip, sm) + (double)rounds * 1500.0 * (bm2 ? 2 : 1) * (bn / 128.0);
      if (cost < best_cost - 1e-9) { best_cost = cost; best_bn = bn; best_bm2 = bm2; }
    }
  }
  {
    // experiment knob (tools/kbench.py sweeps): DEEPCAM_B200_TC_FORCE_BN=<n> overrides the N tile chosen by the cost model
    static int force_bn = -1;
    if (force_bn < 0) { const char* e = getenv("DEEPCAM_B200_TC_FORCE_BN"); force_bn = e ? atoi(e) : 0; }
    if (force_bn >= 16 && force_bn <= 256 && force_bn % 16 == 0 && (!tma_store || force_bn % 64 == 0 || force_bn >= co16)) {
      best_bn = std::min(force_bn, co16);
      best_bm2 = 0;
    }
  }
  c.BN = best_bn;
  c.bm2 = best_bm2;
  c.real_mtiles = mtiles;
  c.n_mtiles = c.bm2 ? ceil_div(mtiles, 2) : mtiles;
  c.n_ntiles = ceil_div(Co, c.BN);
  c.total_tiles = c.n_mtiles * c.n_ntiles;
  c.tma_store = tma_store ? 1 : 0;
  c.wide = c.BN > 256 ? 1 : 0;
  c.bn_sub0 = c.wide ? (c.BN / 2 + 15) / 16 * 16 : c.BN;       // 384 -> 192 + 192, 320 -> 160 + 160, 448 -> 224 + 224
  c.b_box_rows = c.wide ? c.BN / 2 : c.BN;                      // BN % 64 == 0 -> a multiple of 8 rows: 1024-byte aligned boxes
  c.acc_stride = 32;
  while (c.acc_stride < c.BN) c.acc_stride <<= 1;
  c.tmem_cols = c.wide ? c.acc_stride : 2 * c.acc_stride;
  const int budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - (c.wide ? 0 : kV2StagingBytes);
  c.bres_bytes = wtaps * kblocks * c.BN * 128;
  c.b_resident = (!c.bm2 && !c.wide && c.n_ntiles == 1 && c.bres_bytes <= 128 * 1024 && c.bres_bytes + 3 * kABytes <= budget) ? 1 : 0;
  if (!c.b_resident) c.bres_bytes = 0;
  c.stage_bytes = (c.bm2 ? 2 : 1) * kABytes + (c.b_resident ? 0 : c.BN * 128);
  c.stages = std::max(2, std::min(8, (budget - c.bres_bytes) / c.stage_bytes));
  c.smem_bytes = c.stages * c.stage_bytes + c.bres_bytes + (c.wide ? 0 : kV2StagingBytes) + 1024 + 256;
  return c;
}

static int launch_fprop_v2(const TcMaps& maps, const TcFpropParams& p, const TcV2Cfg& cfg, cudaStream_t st) {
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail((int)e, "dc_conv_gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int grid = std::min(cfg.total_tiles, kNumSMs);
  if (cfg.halo) launch_k(conv_gemm_tc2_kernel<true>, grid, dim3(kTc2HaloThreads), (size_t)cfg.smem_bytes, st, maps, p, cfg);
  else launch_k(conv_gemm_tc2_kernel<false>, grid, dim3(kTc2Threads), (size_t)cfg.smem_bytes, st, maps, p, cfg);
  return launch_status("dc_conv_gemm_tc");
}

static bool use_v1_kernel() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DEEPCAM_B200_TC_V1"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// ---- halo mode planning ------------------------------------------------------------------------------------------------
static bool halo_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DEEPCAM_B200_TC_HALO"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// Fills cfg (halo fields + the common ones) and returns true when the layer should and can run in halo mode: a multi-tap gather
// with uniform stride 1 | 2 whose narrow side (gathered channels or output channels) makes the per-tap re-reads dominate,
// one N tile, resident weights + at least one halo buffer within shared memory.
static bool plan_halo(const dc_conv_desc* d, const dc_view& in, const dc_view& out, bool tma_store, TcV2Cfg& c) {
  if (!halo_enabled() || use_v1_kernel()) return false;
  if (d->ntaps < 2 || d->stride_h != d->stride_w || (d->stride_h != 1 && d->stride_h != 2)) return false;
  // (gathered operands wider than 64 channels would need the halo staged in channel chunks to leave room for double buffering:
  //  the fused last_deconv forward, 256 -> 4 x 8 with a 78 KB region, stays on the per-tap boxes for now)
  if (in.c > 64 || in.c % 8 || out.c > 256) return false;
  int dh0 = d->dh[0], dh1 = d->dh[0], dw0 = d->dw[0], dw1 = d->dw[0];
  for (int t = 1; t < d->ntaps; ++t) {
    dh0 = std::min(dh0, d->dh[t]); dh1 = std::max(dh1, d->dh[t]);
    dw0 = std::min(dw0, d->dw[t]); dw1 = std::max(dw1, d->dw[t]);
  }
  if (dh1 - dh0 > 4 || dw1 - dw0 > 4) return false;
  const int s = d->stride_h;
  c = TcV2Cfg();
  c.halo = 1;
  c.halo_s = s; c.halo_x0 = dw0; c.halo_y0 = dh0; c.halo_ci = in.c;
  c.halo_h = (kHaloTH - 1) * s + (dh1 - dh0) + 1;
  c.halo_w = (kHaloTW - 1) * s + (dw1 - dw0) + 1;
  if (c.halo_h > 256 || c.halo_w > 256) return false;
  if (in.sw != in.c || in.sc != 1) return false;        // W and C must be one contiguous dimension (dense NHWC rows)
  c.halo_nbox = ceil_div(c.halo_w * in.c, 256);
  c.halo_bytes = c.halo_nbox * c.halo_h * 512;          // every box is charged in full (out-of-range parts are zero-filled)
  c.halo_stride = round_up_i(c.halo_bytes, 1024);
  c.ktot = d->wtaps * in.c;
  c.nkb = ceil_div(c.ktot, 64);
  c.BN = round_up_i(out.c, 16);
  c.bm2 = 0; c.wide = 0; c.bn_sub0 = c.BN; c.b_box_rows = c.BN;
  c.b_resident = 1;
  c.bres_bytes = c.nkb * c.BN * 128;
  c.stage_bytes = kABytes;
  const int fixed = 1024 + 256 + kV2StagingBytes + c.bres_bytes;
  if (227 * 1024 - fixed - 3 * c.stage_bytes < c.halo_stride) return false;
  // halo buffers = tiles of prefetch distance: the load for tile i + bufs is issued when the producers leave tile i, and a
  // load takes 0.8-1.2 us (tools/halo_trace.py), about as long as the producers spend on a tile: up to 4, at least 1
  c.halo_bufs = std::max(1, std::min(4, (227 * 1024 - fixed - 4 * c.stage_bytes) / c.halo_stride));
  {
    static int force_bufs = -1;
    if (force_bufs < 0) { const char* e = getenv("DEEPCAM_B200_TC_HALO_BUFS"); force_bufs = e ? atoi(e) : 0; }
    if (force_bufs >= 1 && force_bufs <= c.halo_bufs) c.halo_bufs = force_bufs;
  }
  // the four producer warps fill four stages at once: as many stages as fit (4..8), never fewer than 3
  c.stages = std::max(3, std::min(8, (227 * 1024 - fixed - c.halo_bufs * c.halo_stride) / c.stage_bytes));
  c.acc_stride = 32;
  while (c.acc_stride < c.BN) c.acc_stride <<= 1;
  c.tmem_cols = 2 * c.acc_stride;
  c.tma_store = tma_store ? 1 : 0;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("DEEPCAM_B200_TC_HALO_DBG"); dbg = e ? atoi(e) : 0; }
    c.dbg = dbg;
    static int force_stages = -1;
    if (force_stages < 0) { const char* e = getenv("DEEPCAM_B200_TC_HALO_STAGES"); force_stages = e ? atoi(e) : 0; }
    if (force_stages >= 2 && force_stages <= c.stages) c.stages = force_stages;
  }
  c.smem_bytes = c.stages * c.stage_bytes + c.bres_bytes + c.halo_bufs * c.halo_stride + kV2StagingBytes + 1024 + 256;
  return true;
}

// un-swizzled 4-D bf16 box {Ci, halo_w, halo_h, 1} over the NHWC input (zero fill outside = padding)
static int encode_halo_map(CUtensorMap* m, const dc_view& in, int halo_w, int halo_h, const char* what) {
  PFN_encodeTiled enc = get_encode();
  DC_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled unavailable", what);
  (void)halo_w;
  // {W * C, H, N}: a dense NHWC row is one contiguous run of W * C elements; boxes of 256 elements x halo_h rows
  cuuint64_t dims[3] = {(cuuint64_t)in.w * (cuuint64_t)in.c, (cuuint64_t)in.h, (cuuint64_t)in.n};
  cuuint64_t strides[2] = {(cuuint64_t)in.sh * 2, (cuuint64_t)in.sn * 2};
  cuuint32_t box[3] = {256, (cuuint32_t)halo_h, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(in.ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DC_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(halo) failed with %d (C=%d box %dx%d)", what, (int)r, in.c, halo_w, halo_h);
  return 0;
}


template <int BNW>
static int launch_wgrad(const TcMaps& maps, const TcWgradParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = WgradCfg<BNW>;
  cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<BNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
  if (e != cudaSuccess) return fail((int)e, "dc_conv_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  launch_k(conv_wgrad_tc_kernel<BNW>, grid, dim3(kTcThreads), (size_t)Cfg::kSmem, st, maps, p);
  return launch_status("dc_conv_wgrad_tc");
}

}  // namespace dc

using namespace dc;

extern "C" {

// stats != null: *stats_done tells the caller whether the epilogue accumulated the BatchNorm sums (TMA-store path)
struct TcAffine {           // eval-mode BatchNorm folded into the epilogue (null gamma = none)
  const float* gamma; const float* beta; const float* mean; const float* var; float eps; int relu;
};

static int conv_gemm_tc_impl(const dc_conv_desc* d, dc_view in, const void* w, const float* bias, dc_view out, double* stats,
                             bool* stats_done, void* stream, const TcAffine* aff = nullptr) {
  DC_REQUIRE(d != nullptr && d->ntaps >= 1 && d->ntaps <= DC_MAX_TAPS, "dc_conv_gemm_tc: bad descriptor");
  DC_REQUIRE(tc_view_ok(in), "dc_conv_gemm_tc: input must be bf16, channel-contiguous, C %% 8 == 0, 16-byte aligned strides");
  DC_REQUIRE(view_ok(out) && out.n == in.n, "dc_conv_gemm_tc: bad output view");
  DC_REQUIRE(w != nullptr && (reinterpret_cast<uintptr_t>(w) % 16) == 0, "dc_conv_gemm_tc: weights must be 16-byte aligned");
  DC_REQUIRE(d->wtaps >= 1 && d->wtaps <= DC_MAX_TAPS, "dc_conv_gemm_tc: bad wtaps");
  TcMaps maps;
  TcFpropParams p;
  p.ntaps = d->ntaps;
  pick_tile(out.h, out.w, p.TH, p.TW);
  p.tiles_x = ceil_div(out.w, p.TW);
  p.tiles_y = ceil_div(out.h, p.TH);
  p.kblocks = ceil_div(in.c, 64);
  p.accumulate = d->accumulate;
  p.bias = bias;
  p.out.p = out.ptr; p.out.n = out.n; p.out.h = out.h; p.out.w = out.w; p.out.c = out.c;
  p.out.sn = out.sn; p.out.sh = out.sh; p.out.sw = out.sw; p.out.sc = out.sc; p.out.dtype = out.dtype;
  DC_REQUIRE(d->out_csplit >= 0 && d->out_csplit % 4 == 0 && d->out_csplit < std::max(out.c, 1),
             "dc_conv_gemm_tc: out_csplit must be a multiple of 4 below the output channel count");
  p.out.csplit = d->out_csplit; p.out.split_off = d->out_split_off;
  p.out_vec_ok = (d->out_csplit == 0 && out.dtype == DC_BF16 && out.sc == 1 && out.sw % 8 == 0 && out.sh % 8 == 0 && out.sn % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(out.ptr) % 16) == 0) ? 1 : 0;
  p.stats = nullptr;
  p.stats_C = out.c;
  p.b_early = (d->flags & DC_CONV_WEIGHTS_STABLE) ? 1 : 0;
  p.aff_gamma = p.aff_beta = p.aff_mean = p.aff_var = nullptr;
  p.aff_eps = 0.f; p.aff_relu = 0;
  if (aff != nullptr) {
    // the folded epilogue exists on the persistent kernel's TMA-store path only
    if (use_v1_kernel() || !p.out_vec_ok || d->accumulate) return fail(-2, "dc_conv_gemm_tc_bn_eval: output layout needs the generic epilogue");
    p.aff_gamma = aff->gamma; p.aff_beta = aff->beta; p.aff_mean = aff->mean; p.aff_var = aff->var;
    p.aff_eps = aff->eps; p.aff_relu = aff->relu;
  }
  if (stats_done) *stats_done = false;
  {
    // HALO mode: the caller packed the weights K-dense ([Co][slice][Ci], DC_CONV_HALO_PACK) - or Ci is a multiple of 64, where
    // the dense and the 64-padded layouts coincide
    TcV2Cfg hc;
    const bool dense_pack = (d->flags & DC_CONV_HALO_PACK) || (in.c % 64 == 0);
    const bool halo = dense_pack && plan_halo(d, in, out, p.out_vec_ok != 0, hc);
    DC_REQUIRE(halo || !(d->flags & DC_CONV_HALO_PACK) || in.c % 64 == 0,
               "dc_conv_gemm_tc: DC_CONV_HALO_PACK weights but the layer cannot run in halo mode (query dc_conv_gemm_tc_halo_ok first)");
    if (halo) {
      p.TH = kHaloTH; p.TW = kHaloTW;
      p.tiles_x = ceil_div(out.w, p.TW);
      p.tiles_y = ceil_div(out.h, p.TH);
      p.kblocks = hc.nkb;
      for (int t = 0; t < d->ntaps; ++t) {
        p.map_id[t] = 0; p.wt[t] = d->wt[t];
        p.qh[t] = d->dh[t] - hc.halo_y0; p.qw[t] = d->dw[t] - hc.halo_x0;
      }
      hc.real_mtiles = p.tiles_x * p.tiles_y * out.n;
      hc.n_mtiles = hc.real_mtiles; hc.n_ntiles = 1; hc.total_tiles = hc.real_mtiles;
      if (int r = encode_halo_map(&maps.a[0], in, hc.halo_w, hc.halo_h, "dc_conv_gemm_tc")) return r;
      for (int id = 1; id < 4; ++id) maps.a[id] = maps.a[0];
      if (int r = encode_weight_map(&maps.b, w, hc.ktot, out.c, hc.BN, "dc_conv_gemm_tc")) return r;
      if (hc.tma_store) {
        if (int r = encode_act_map(&maps.c, out.ptr, out.c, out.w, out.h, out.n, out.sw, out.sh, out.sn, p.TW, p.TH, "dc_conv_gemm_tc(out)"))
          return r;
        if (stats != nullptr && !d->accumulate) { p.stats = stats; *stats_done = true; }
      }
      return launch_fprop_v2(maps, p, hc, as_stream(stream));
    }
  }
  if (int r = build_gather("dc_conv_gemm_tc", d, in, p.TH, p.TW, maps, p)) return r;
  const long long Ktot = (long long)d->wtaps * p.kblocks * 64;
  const int mtiles = p.tiles_x * p.tiles_y * out.n;
  if (!use_v1_kernel()) {
    const bool tma_store = p.out_vec_ok != 0;
    const TcV2Cfg c2 = pick_v2_cfg(mtiles, out.c, p.kblocks, d->wtaps, tma_store);
    if (int r = encode_weight_map(&maps.b, w, Ktot, out.c, c2.b_box_rows, "dc_conv_gemm_tc")) return r;
    if (tma_store) {
      if (int r = encode_act_map(&maps.c, out.ptr, out.c, out.w, out.h, out.n, out.sw, out.sh, out.sn, p.TW, p.TH, "dc_conv_gemm_tc(out)"))
        return r;
      if (stats != nullptr && !d->accumulate) {
        p.stats = stats;
        *stats_done = true;
      }
    }
    return launch_fprop_v2(maps, p, c2, as_stream(stream));
  }
  const TcFpropCfg cfg = pick_fprop_cfg(mtiles, out.c);
  if (int r = encode_weight_map(&maps.b, w, Ktot, out.c, cfg.BN, "dc_conv_gemm_tc")) return r;
  return launch_fprop(maps, p, cfg, mtiles, ceil_div(out.c, cfg.BN), as_stream(stream));
}

#ifdef DC_TC_TRACE
// trace build only (tools/tc_trace.py): copies the 148 x 16 timestamp table to the host
int dc_halo_trace_read(unsigned long long* host) {
  cudaError_t e = cudaMemcpyFromSymbol(host, g_halo_trace, sizeof(unsigned long long) * (32 * 8 + 8));
  return e == cudaSuccess ? 0 : (int)e;
}
int dc_tc_trace_read(unsigned long long* host) {
  cudaError_t e = cudaMemcpyFromSymbol(host, g_tc_trace, sizeof(unsigned long long) * 148 * 16);
  return e == cudaSuccess ? 0 : (int)e;
}
#endif

int dc_conv_gemm_tc(const dc_conv_desc* d, dc_view in, const void* w, const float* bias, dc_view out, void* stream) {
  return conv_gemm_tc_impl(d, in, w, bias, out, nullptr, nullptr, stream);
}

int dc_conv_gemm_tc_bnstats(const dc_conv_desc* d, dc_view in, const void* w, const float* bias, dc_view out, double* sums,
                            void* stream) {
  DC_REQUIRE(sums != nullptr, "dc_conv_gemm_tc_bnstats: null statistics workspace");
  DC_REQUIRE(d != nullptr && !d->accumulate, "dc_conv_gemm_tc_bnstats: statistics of an accumulating contraction are undefined");
  DC_REQUIRE(d->out_csplit == 0, "dc_conv_gemm_tc_bnstats: two-segment outputs are not supported");
  bool done = false;
  if (int r = conv_gemm_tc_impl(d, in, w, bias, out, sums, &done, stream)) return r;
  if (done) return 0;
  return dc::bn_accumulate_sums(out, sums, as_stream(stream));     // output layout without the TMA-store epilogue: one extra pass
}

int dc_conv_gemm_tc_halo_ok(const dc_conv_desc* d, dc_view in, dc_view out) {
  if (d == nullptr || d->ntaps < 1 || d->ntaps > DC_MAX_TAPS || !tc_view_ok(in) || !view_ok(out)) return 0;
  TcV2Cfg hc;
  return plan_halo(d, in, out, true, hc) ? 1 : 0;
}

int dc_conv_gemm_tc_bn_eval(const dc_conv_desc* d, dc_view in, const void* w, const float* bias, dc_view out, const float* gamma,
                            const float* beta, const float* running_mean, const float* running_var, float eps, int relu, void* stream) {
  DC_REQUIRE(gamma != nullptr && beta != nullptr && running_mean != nullptr && running_var != nullptr,
             "dc_conv_gemm_tc_bn_eval: null BatchNorm argument");
  DC_REQUIRE(d != nullptr && d->out_csplit == 0, "dc_conv_gemm_tc_bn_eval: two-segment outputs are not supported");
  const TcAffine aff = {gamma, beta, running_mean, running_var, eps, relu ? 1 : 0};
  return conv_gemm_tc_impl(d, in, w, bias, out, nullptr, nullptr, stream, &aff);
}

// ci tile and pixel-split count of the weight-gradient kernel (shared by the launch and by dc_conv_wgrad_tc_ws_elems)
static void wgrad_split_plan(const dc_conv_desc* d, const dc_view& in, const dc_view& dout, int& BNW, int& mtiles_total, int& per_split, int& splits) {
  int TH, TW;
  pick_tile(dout.h, dout.w, TH, TW);
  mtiles_total = ceil_div(dout.w, TW) * ceil_div(dout.h, TH) * dout.n;
  // ci tile: 256 when it does not waste columns (wider tiles halve the dY re-reads from L2), else 128 / 64
  BNW = in.c <= 64 ? 64 : ((in.c > 128 && (in.c % 256 == 0 || in.c % 256 > 128)) ? 256 : 128);
  const int tiles = ceil_div(dout.c, 128) * ceil_div(in.c, BNW) * d->ntaps;
  // one CTA per SM (the stages fill the shared memory): split the pixel reduction so that the grid is one wave
  static int split_div = -1;      // tuning knob (DEEPCAM_B200_WGRAD_SPLIT_DIV): divide the one-wave split count
  if (split_div < 0) { const char* e = getenv("DEEPCAM_B200_WGRAD_SPLIT_DIV"); split_div = e ? std::max(1, atoi(e)) : 1; }
  static int split_mul = -1;      // and DEEPCAM_B200_WGRAD_SPLIT_MUL: multiply it (more than one wave of CTAs)
  if (split_mul < 0) { const char* e = getenv("DEEPCAM_B200_WGRAD_SPLIT_MUL"); split_mul = e ? std::max(1, atoi(e)) : 1; }
  splits = std::max(1, std::min(mtiles_total, kNumSMs * split_mul / std::max(1, tiles) / split_div));
  per_split = ceil_div(mtiles_total, splits);
  splits = ceil_div(mtiles_total, per_split);
}

// Halo mode of the weight gradient (conv_wgrad_tc_halo_kernel): multi-tap, uniform stride 1 | 2, gathered operand with <= 64
// channels in dense NHWC rows.  Fills p (everything but G / det_wtaps) and the split count; false = use the per-tap kernel.
static bool plan_wgrad_halo(const dc_conv_desc* d, const dc_view& in, const dc_view& dout, TcWgradHaloParams& p, int& splits, int& smem_bytes) {
  static int enabled = -1;        // DEEPCAM_B200_WGRAD_HALO=0: per-tap kernel (A/B measurements)
  if (enabled < 0) { const char* e = getenv("DEEPCAM_B200_WGRAD_HALO"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return false;
  if (d->ntaps < 2 || d->stride_h != d->stride_w || (d->stride_h != 1 && d->stride_h != 2)) return false;
  if (in.c > 64 || in.c % 8 || in.sw != in.c || in.sc != 1) return false;
  int dh0 = d->dh[0], dh1 = d->dh[0], dw0 = d->dw[0], dw1 = d->dw[0];
  for (int t = 1; t < d->ntaps; ++t) {
    dh0 = std::min(dh0, d->dh[t]); dh1 = std::max(dh1, d->dh[t]);
    dw0 = std::min(dw0, d->dw[t]); dw1 = std::max(dw1, d->dw[t]);
  }
  if (dh1 - dh0 > 4 || dw1 - dw0 > 4) return false;
  for (int a = 0; a < d->ntaps; ++a)
    for (int b = a + 1; b < d->ntaps; ++b)
      if (d->wt[a] == d->wt[b]) return false;
  p = TcWgradHaloParams();
  p.s = d->stride_h;
  p.Co = dout.c; p.Ci = in.c;
  p.tiles_x = ceil_div(dout.w, kHaloTW);
  p.tiles_y = ceil_div(dout.h, kHaloTH);
  p.n_img = dout.n;
  p.mtiles_total = p.tiles_x * p.tiles_y * dout.n;
  p.halo_x0 = dw0; p.halo_y0 = dh0;
  p.halo_h = (kHaloTH - 1) * p.s + (dh1 - dh0) + 1;
  p.halo_w = (kHaloTW - 1) * p.s + (dw1 - dw0) + 1;
  p.halo_nbox = ceil_div(p.halo_w * in.c, 256);
  p.halo_bytes = p.halo_nbox * p.halo_h * 512;
  p.halo_stride = round_up_i(p.halo_bytes, 1024);
  // tap groups: N = taps x Ci <= 192 per CTA (three 64-column blocks of the B operand, double buffered)
  const int per_group = std::max(1, 192 / in.c);
  p.ngroups = ceil_div(d->ntaps, per_group);
  if (p.ngroups > 3) return false;
  const int even = ceil_div(d->ntaps, p.ngroups);
  int first = 0, max_npad = 0;
  for (int g = 0; g < p.ngroups; ++g) {
    p.g_first[g] = first;
    p.g_count[g] = std::min(even, d->ntaps - first);
    p.g_npad[g] = round_up_i(p.g_count[g] * in.c, 16);
    max_npad = std::max(max_npad, p.g_npad[g]);
    first += p.g_count[g];
  }
  for (int t = 0; t < d->ntaps; ++t) { p.qh[t] = d->dh[t] - dh0; p.qw[t] = d->dw[t] - dw0; p.wt[t] = d->wt[t]; }
  p.nblk = ceil_div(max_npad, 64);
  p.tmem_cols = 32;
  while (p.tmem_cols < max_npad) p.tmem_cols <<= 1;
  p.a_boxes = dout.c > 64 ? 2 : 1;
  p.a_bytes = p.a_boxes * 16384;
  const int fixed = 2 * p.nblk * 16384 + 1024 + 256;
  int st = 4;
  while (st >= 2 && fixed + st * (p.a_bytes + p.halo_stride) > 227 * 1024) --st;
  if (st < 2) return false;
  p.sa = p.sr = st;
  smem_bytes = fixed + st * (p.a_bytes + p.halo_stride);
  const int n_co_tiles = ceil_div(dout.c, 128);
  splits = std::max(1, std::min(p.mtiles_total, kNumSMs / (n_co_tiles * p.ngroups)));
  p.mtiles_per_split = ceil_div(p.mtiles_total, splits);
  splits = ceil_div(p.mtiles_total, p.mtiles_per_split);
  return true;
}

/* ws != null: deterministic two-stage form - the pixel splits store their partial tiles into ws[split][wtaps][Co][Ci] and a second
   launch adds the slices to G in split order */
static int conv_wgrad_tc_impl(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, float* ws, long long ws_elems, void* stream) {
  DC_REQUIRE(d != nullptr && d->ntaps >= 1 && d->ntaps <= DC_MAX_TAPS, "dc_conv_wgrad_tc: bad descriptor");
  DC_REQUIRE(tc_view_ok(in) && tc_view_ok(dout), "dc_conv_wgrad_tc: views must be bf16, channel-contiguous, C %% 8 == 0, aligned");
  DC_REQUIRE(in.n == dout.n && G != nullptr, "dc_conv_wgrad_tc: bad arguments");
  TcMaps maps;
  {
    TcWgradHaloParams hp;
    int hsplits = 1, hsmem = 0;
    if ((reinterpret_cast<uintptr_t>(G) % 16) == 0 && plan_wgrad_halo(d, in, dout, hp, hsplits, hsmem)) {
      const long long per_tap = (long long)dout.c * in.c, slice = per_tap * d->wtaps;
      const bool det = ws != nullptr && hsplits > 1;
      if (det) {
        DC_REQUIRE(ws_elems >= slice * hsplits, "dc_conv_wgrad_tc_det: workspace of %lld floats required, %lld given", slice * hsplits, ws_elems);
        DC_REQUIRE((reinterpret_cast<uintptr_t>(ws) % 16) == 0, "dc_conv_wgrad_tc_det: workspace must be 16-byte aligned");
      }
      hp.G = det ? ws : G;
      hp.det_wtaps = det ? d->wtaps : 0;
      if (int r = encode_act_map(&maps.b, dout.ptr, dout.c, dout.w, dout.h, dout.n, dout.sw, dout.sh, dout.sn, kHaloTW, kHaloTH, "dc_conv_wgrad_tc"))
        return r;
      if (int r = encode_halo_map(&maps.a[0], in, hp.halo_w, hp.halo_h, "dc_conv_wgrad_tc")) return r;
      for (int id = 1; id < 4; ++id) maps.a[id] = maps.a[0];
      static bool attr_set = false;
      if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail((int)e, "dc_conv_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
      }
      dim3 grid((unsigned)ceil_div(dout.c, 128), (unsigned)hp.ngroups, (unsigned)hsplits);
      cudaStream_t st = as_stream(stream);
      launch_k(conv_wgrad_tc_halo_kernel, grid, dim3(kWgradHaloThreads), (size_t)hsmem, st, maps, hp);
      if (int r = launch_status("dc_conv_wgrad_tc")) return r;
      if (!det) return 0;
      SplitReduceTaps taps;
      taps.n = d->ntaps;
      for (int t = 0; t < d->ntaps; ++t) taps.wt[t] = d->wt[t];
      return launch_split_reduce(ws, hsplits, slice, taps, per_tap, G, st);
    }
  }
  TcWgradParams p;
  p.ntaps = d->ntaps;
  pick_tile(dout.h, dout.w, p.TH, p.TW);
  p.tiles_x = ceil_div(dout.w, p.TW);
  p.tiles_y = ceil_div(dout.h, p.TH);
  p.n_img = dout.n;
  p.Co = dout.c; p.Ci = in.c; p.G = G;
  p.det_wtaps = 0;
  if (int r = build_gather("dc_conv_wgrad_tc", d, in, p.TH, p.TW, maps, p)) return r;
  if (int r = encode_act_map(&maps.b, dout.ptr, dout.c, dout.w, dout.h, dout.n, dout.sw, dout.sh, dout.sn, p.TW, p.TH, "dc_conv_wgrad_tc"))
    return r;
  int BNW, splits;
  wgrad_split_plan(d, in, dout, BNW, p.mtiles_total, p.mtiles_per_split, splits);
  p.n_ci_tiles = ceil_div(in.c, BNW);
  const int n_co_tiles = ceil_div(dout.c, 128);
  const long long per_tap = (long long)dout.c * in.c, slice = per_tap * d->wtaps;
  const bool det = ws != nullptr && splits > 1;     // one split: every element has one contribution, nothing to order
  if (det) {
    DC_REQUIRE(ws_elems >= slice * splits, "dc_conv_wgrad_tc_det: workspace of %lld floats required, %lld given", slice * splits, ws_elems);
    DC_REQUIRE((reinterpret_cast<uintptr_t>(ws) % 16) == 0, "dc_conv_wgrad_tc_det: workspace must be 16-byte aligned");
    for (int a = 0; a < d->ntaps; ++a)
      for (int b = a + 1; b < d->ntaps; ++b)
        DC_REQUIRE(d->wt[a] != d->wt[b], "dc_conv_wgrad_tc_det: two taps of one launch share weight tap %d", d->wt[a]);
    p.G = ws;
    p.det_wtaps = d->wtaps;
  }
  float* target = p.G;
  p.vec_red = (in.c % 4 == 0 && (reinterpret_cast<uintptr_t>(target) % 16) == 0) ? 1 : 0;
  static int tma_red_enabled = -1;
  if (tma_red_enabled < 0) { const char* e = getenv("DEEPCAM_B200_WGRAD_TMA_RED"); tma_red_enabled = (e && e[0] == '0') ? 0 : 1; }
  p.tma_red = 0;
  if (p.vec_red && tma_red_enabled) {
    // fp32 gradient G[wtaps][Co][Ci] (deterministic form: [splits * wtaps][Co][Ci]) as a 3-D tensor: boxes of 32 ci x 128 co of one
    // tap, 128B swizzle; rows >= Co and columns >= Ci of edge tiles are clipped by the TMA unit
    PFN_encodeTiled enc = get_encode();
    DC_REQUIRE(enc != nullptr, "dc_conv_wgrad_tc: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[3] = {(cuuint64_t)in.c, (cuuint64_t)dout.c, (cuuint64_t)d->wtaps * (cuuint64_t)(det ? splits : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)in.c * 4, (cuuint64_t)in.c * 4 * (cuuint64_t)dout.c};
    cuuint32_t box[3] = {32, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&maps.c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, target, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) p.tma_red = 1;
  }
  dim3 grid(n_co_tiles * p.n_ci_tiles, d->ntaps, splits);
  cudaStream_t st = as_stream(stream);
  int rc = BNW == 64 ? launch_wgrad<64>(maps, p, grid, st) : BNW == 256 ? launch_wgrad<256>(maps, p, grid, st) : launch_wgrad<128>(maps, p, grid, st);
  if (rc != 0 || !det) return rc;
  SplitReduceTaps taps;
  taps.n = d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) taps.wt[t] = d->wt[t];
  return launch_split_reduce(ws, splits, slice, taps, per_tap, G, st);
}

int dc_conv_wgrad_tc(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, void* stream) {
  return conv_wgrad_tc_impl(d, in, dout, G, nullptr, 0, stream);
}

long long dc_conv_wgrad_tc_ws_elems(const dc_conv_desc* d, dc_view in, dc_view dout) {
  if (d == nullptr || d->ntaps < 1 || d->ntaps > DC_MAX_TAPS || !tc_view_ok(in) || !tc_view_ok(dout)) return -1;
  int BNW, total, per, splits;
  TcWgradHaloParams hp;
  int hsmem = 0;
  if (!plan_wgrad_halo(d, in, dout, hp, splits, hsmem)) wgrad_split_plan(d, in, dout, BNW, total, per, splits);
  return splits > 1 ? (long long)splits * d->wtaps * dout.c * in.c : 0;
}

int dc_conv_wgrad_tc_det(const dc_conv_desc* d, dc_view in, dc_view dout, float* G, float* ws, long long ws_elems, void* stream) {
  DC_REQUIRE(ws != nullptr || dc_conv_wgrad_tc_ws_elems(d, in, dout) == 0, "dc_conv_wgrad_tc_det: workspace required (dc_conv_wgrad_tc_ws_elems)");
  return conv_wgrad_tc_impl(d, in, dout, G, ws, ws_elems, stream);
}

}  // extern "C"
