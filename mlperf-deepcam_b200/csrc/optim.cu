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

}  // namespace dc

using namespace dc;

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
