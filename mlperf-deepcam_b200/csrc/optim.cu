// Multi-tensor Adam / AdamW step: every parameter of the network in ONE launch (the reference builds
// optim.Adam / optim.AdamW over net.parameters() at TR:213-220 and calls optimizer.step() at TR:364; torch runs that as
// ~10 foreach kernels per step).  HBM-bound: reads p, g, m, v and writes p, m, v once = 28 bytes per parameter.
// Same update rule and state (step, exp_avg, exp_avg_sq) as torch.optim.Adam(W) without amsgrad / maximize.
#include "common.cuh"
#include <math.h>

namespace dc {

struct AdamScalars {       // derived on the host in double precision, exactly as torch.optim derives them in Python
  float lr_wd;             // lr * weight_decay (AdamW)
  float weight_decay;
  float one_minus_b1, beta2, one_minus_b2, eps;
  float step_size;         // lr / (1 - beta1^t)
  float bias_c2_sqrt;      // sqrt(1 - beta2^t)
  int adamw;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& a) {
  if (a.adamw) p *= 1.f - a.lr_wd;                         // decoupled weight decay (AdamW)
  else if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
  m = m + (g - m) * a.one_minus_b1;                        // exp_avg.lerp_(grad, 1 - beta1)
  v = a.beta2 * v + a.one_minus_b2 * g * g;                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / a.bias_c2_sqrt + a.eps;
  p = p - a.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const dc_adam_job* __restrict__ jobs, int njobs, AdamScalars a) {
  pdl_sync();
  int lo = 0, hi = njobs - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block_start <= b) lo = mid; else hi = mid - 1;
  }
  const dc_adam_job j = jobs[lo];
  const int lb = b - j.block_start;
  const long long n = j.numel;
  const bool vec = ((reinterpret_cast<uintptr_t>(j.p) | reinterpret_cast<uintptr_t>(j.g) | reinterpret_cast<uintptr_t>(j.m) |
                     reinterpret_cast<uintptr_t>(j.v)) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  float4* p4 = reinterpret_cast<float4*>(j.p);
  const float4* g4 = reinterpret_cast<const float4*>(j.g);
  float4* m4 = reinterpret_cast<float4*>(j.m);
  float4* v4 = reinterpret_cast<float4*>(j.v);
  const long long stride = (long long)j.n_blocks * 256;
  for (long long i = (long long)lb * 256 + threadIdx.x; i < n4; i += stride) {
    float4 p = p4[i], m = m4[i], v = v4[i];
    const float4 g = g4[i];
    adam_one(p.x, g.x, m.x, v.x, a);
    adam_one(p.y, g.y, m.y, v.y, a);
    adam_one(p.z, g.z, m.z, v.z, a);
    adam_one(p.w, g.w, m.w, v.w, a);
    p4[i] = p; m4[i] = m; v4[i] = v;
  }
  for (long long i = n4 * 4 + (long long)lb * 256 + threadIdx.x; i < n; i += stride) {
    float p = j.p[i], m = j.m[i], v = j.v[i];
    adam_one(p, j.g[i], m, v, a);
    j.p[i] = p; j.m[i] = m; j.v[i] = v;
  }
}

// ---- LAMB (the optimizer of the reference's DGX recipe: apex.optimizers.FusedLAMB at TR:217-218) ----------------------
// apex is not vendored in the reference tree; this follows its published algorithm (apex/optimizers/fused_lamb.py,
// csrc/multi_tensor_lamb.cu): global gradient-norm clipping, Adam moments on the clipped gradient, per-tensor trust ratio
// ||p|| / ||update||.  Three launches for all parameters: gradient norm; moments + update (written over the gradient, as
// apex does) + per-tensor norms; parameter update.  norms = double[1 + 2 * njobs]: [0] sum g^2, [1 + 2i] sum p_i^2,
// [2 + 2i] sum update_i^2 (zeroed by the caller).
__device__ __forceinline__ int find_job(const dc_adam_job* __restrict__ jobs, int njobs, int b) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block_start <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ void block_add(double v, double* dst) {
  __shared__ double s_part[8];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(dst, t);
  }
}

__global__ void __launch_bounds__(256) lamb_grad_norm_kernel(const dc_adam_job* __restrict__ jobs, int njobs, double* norms) {
  pdl_sync();
  const int ji = find_job(jobs, njobs, blockIdx.x);
  const dc_adam_job j = jobs[ji];
  const long long stride = (long long)j.n_blocks * 256;
  double acc = 0.0;
  for (long long i = (long long)(blockIdx.x - j.block_start) * 256 + threadIdx.x; i < j.numel; i += stride) {
    const float g = j.g[i];
    acc += (double)g * (double)g;
  }
  block_add(acc, norms);
}

struct LambScalars {
  float beta1, beta2, beta3, inv_bc1, inv_bc2, eps, decay, max_grad_norm;
  int adam_w_mode;
};
__global__ void __launch_bounds__(256) lamb_stage1_kernel(const dc_adam_job* __restrict__ jobs, int njobs, LambScalars a, double* norms) {
  pdl_sync();
  const int ji = find_job(jobs, njobs, blockIdx.x);
  const dc_adam_job j = jobs[ji];
  const float gnorm = (float)sqrt(norms[0]);
  const float clip = (a.max_grad_norm > 0.f && gnorm > a.max_grad_norm) ? gnorm / a.max_grad_norm : 1.f;
  const long long stride = (long long)j.n_blocks * 256;
  double pn = 0.0, un = 0.0;
  for (long long i = (long long)(blockIdx.x - j.block_start) * 256 + threadIdx.x; i < j.numel; i += stride) {
    const float p = j.p[i];
    float g = j.g[i] / clip;
    if (!a.adam_w_mode) g = g + a.decay * p;                       // L2 mode: decay enters the moments
    const float m = j.m[i] * a.beta1 + a.beta3 * g;
    const float v = j.v[i] * a.beta2 + (1.f - a.beta2) * g * g;
    j.m[i] = m;
    j.v[i] = v;
    const float denom = sqrtf(v * a.inv_bc2) + a.eps;
    float u = (m * a.inv_bc1) / denom;
    if (a.adam_w_mode) u = u + a.decay * p;
    const_cast<float*>(j.g)[i] = u;                                  // the update replaces the gradient (stage 2 reads it)
    pn += (double)p * (double)p;
    un += (double)u * (double)u;
  }
  block_add(pn, norms + 1 + 2 * ji);
  block_add(un, norms + 2 + 2 * ji);
}
__global__ void __launch_bounds__(256) lamb_stage2_kernel(const dc_adam_job* __restrict__ jobs, int njobs, float lr, int use_ratio,
                                                          const double* __restrict__ norms) {
  pdl_sync();
  const int ji = find_job(jobs, njobs, blockIdx.x);
  const dc_adam_job j = jobs[ji];
  float ratio = lr;
  if (use_ratio) {
    const float pn = (float)sqrt(norms[1 + 2 * ji]), un = (float)sqrt(norms[2 + 2 * ji]);
    if (pn != 0.f && un != 0.f) ratio = lr * (pn / un);
  }
  const long long stride = (long long)j.n_blocks * 256;
  for (long long i = (long long)(blockIdx.x - j.block_start) * 256 + threadIdx.x; i < j.numel; i += stride)
    j.p[i] = j.p[i] - ratio * j.g[i];
}

// ---- LARS (BASELINE.json configs[4] asks for a LARS sweep; the reference has no LARS, so this follows You et al. 2017 as
// commonly implemented: per-tensor trust ratio on SGD with momentum) -------------------------------------------------------
//   local_lr_i = trust * ||w_i|| / (||g_i|| + wd * ||w_i|| + eps)   (1 when either norm is 0)
//   v = momentum * v + lr * local_lr_i * (g + wd * w);   w -= v
// norms = double[2 * njobs]: [2i] sum w_i^2, [2i + 1] sum g_i^2 (zeroed by the entry point).
__global__ void __launch_bounds__(256) lars_norms_kernel(const dc_adam_job* __restrict__ jobs, int njobs, double* norms) {
  pdl_sync();
  const int ji = find_job(jobs, njobs, blockIdx.x);
  const dc_adam_job j = jobs[ji];
  const long long stride = (long long)j.n_blocks * 256;
  double wn = 0.0, gn = 0.0;
  for (long long i = (long long)(blockIdx.x - j.block_start) * 256 + threadIdx.x; i < j.numel; i += stride) {
    const float w = j.p[i], g = j.g[i];
    wn += (double)w * (double)w;
    gn += (double)g * (double)g;
  }
  block_add(wn, norms + 2 * ji);
  block_add(gn, norms + 2 * ji + 1);
}
__global__ void __launch_bounds__(256) lars_update_kernel(const dc_adam_job* __restrict__ jobs, int njobs, float lr, float momentum,
                                                          float wd, float trust, float eps, const double* __restrict__ norms) {
  pdl_sync();
  const int ji = find_job(jobs, njobs, blockIdx.x);
  const dc_adam_job j = jobs[ji];
  const float wn = (float)sqrt(norms[2 * ji]), gn = (float)sqrt(norms[2 * ji + 1]);
  float local_lr = 1.f;
  if (wn > 0.f && gn > 0.f) local_lr = trust * wn / (gn + wd * wn + eps);
  const float step = lr * local_lr;
  const long long stride = (long long)j.n_blocks * 256;
  for (long long i = (long long)(blockIdx.x - j.block_start) * 256 + threadIdx.x; i < j.numel; i += stride) {
    const float w = j.p[i];
    const float v = momentum * j.m[i] + step * (j.g[i] + wd * w);      // j.m = momentum buffer
    j.m[i] = v;
    j.p[i] = w - v;
  }
}

}  // namespace dc

using namespace dc;

extern "C" int dc_lars_step_multi(const dc_adam_job* jobs_dev, int njobs, int total_blocks, double lr, double momentum,
                                  double weight_decay, double trust_coefficient, double eps, double* norms, void* stream) {
  DC_REQUIRE(jobs_dev != nullptr && njobs > 0 && total_blocks > 0 && norms != nullptr, "dc_lars_step_multi: bad arguments");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(norms, 0, sizeof(double) * (size_t)(2 * njobs), st);
  if (e != cudaSuccess) return dc::fail((int)e, "dc_lars_step_multi: %s", cudaGetErrorString(e));
  launch_k(lars_norms_kernel, dim3(total_blocks), dim3(256), (size_t)0, st, jobs_dev, njobs, norms);
  launch_k(lars_update_kernel, dim3(total_blocks), dim3(256), (size_t)0, st, jobs_dev, njobs, (float)lr, (float)momentum,
           (float)weight_decay, (float)trust_coefficient, (float)eps, (const double*)norms);
  return launch_status("dc_lars_step_multi");
}

extern "C" int dc_lamb_step_multi(const dc_adam_job* jobs_dev, int njobs, int total_blocks, double lr, double beta1, double beta2,
                                  double eps, double weight_decay, double bias_c1, double bias_c2, int adam_w_mode,
                                  int grad_averaging, double max_grad_norm, int use_nvlamb, double* norms, void* stream) {
  DC_REQUIRE(jobs_dev != nullptr && njobs > 0 && total_blocks > 0 && norms != nullptr, "dc_lamb_step_multi: bad arguments");
  DC_REQUIRE(bias_c1 > 0.0 && bias_c2 > 0.0, "dc_lamb_step_multi: bias corrections must be positive");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(norms, 0, sizeof(double) * (size_t)(1 + 2 * njobs), st);
  if (e != cudaSuccess) return dc::fail((int)e, "dc_lamb_step_multi: %s", cudaGetErrorString(e));
  LambScalars a;
  a.beta1 = (float)beta1; a.beta2 = (float)beta2;
  a.beta3 = grad_averaging ? (float)(1.0 - beta1) : 1.f;
  a.inv_bc1 = (float)(1.0 / bias_c1); a.inv_bc2 = (float)(1.0 / bias_c2);
  a.eps = (float)eps; a.decay = (float)weight_decay; a.max_grad_norm = (float)max_grad_norm;
  a.adam_w_mode = adam_w_mode;
  launch_k(lamb_grad_norm_kernel, dim3(total_blocks), dim3(256), (size_t)0, st, jobs_dev, njobs, norms);
  launch_k(lamb_stage1_kernel, dim3(total_blocks), dim3(256), (size_t)0, st, jobs_dev, njobs, a, norms);
  launch_k(lamb_stage2_kernel, dim3(total_blocks), dim3(256), (size_t)0, st, jobs_dev, njobs, (float)lr,
           (use_nvlamb || weight_decay != 0.0) ? 1 : 0, (const double*)norms);
  return launch_status("dc_lamb_step_multi");
}

extern "C" int dc_adam_step_multi(const dc_adam_job* jobs_dev, int njobs, int total_blocks, double lr, double beta1, double beta2,
                                  double eps, double weight_decay, double bias_c1, double bias_c2, int adamw, void* stream) {
  DC_REQUIRE(jobs_dev != nullptr && njobs > 0 && total_blocks > 0, "dc_adam_step_multi: bad arguments");
  DC_REQUIRE(bias_c1 > 0.0 && bias_c2 > 0.0, "dc_adam_step_multi: bias corrections must be positive (step >= 1)");
  AdamScalars a;
  a.lr_wd = (float)(lr * weight_decay);
  a.weight_decay = (float)weight_decay;
  a.one_minus_b1 = (float)(1.0 - beta1);
  a.beta2 = (float)beta2;
  a.one_minus_b2 = (float)(1.0 - beta2);
  a.eps = (float)eps;
  a.step_size = (float)(lr / bias_c1);
  a.bias_c2_sqrt = (float)sqrt(bias_c2);
  a.adamw = adamw;
  launch_k(adam_multi_kernel, dim3(total_blocks), dim3(256), (size_t)0, as_stream(stream), jobs_dev, njobs, a);
  return launch_status("dc_adam_step_multi");
}
