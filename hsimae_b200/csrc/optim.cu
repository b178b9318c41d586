// Fused multi-tensor AdamW.
//
// Reference: the optimiser step of the pretraining / fine-tuning loops, /root/reference/Model_Pretraining.py:80-86,100-102
// (torch.optim.AdamW over two parameter groups: weight decay for everything whose name contains neither 'bias' nor
// 'norm', betas (0.9, 0.95)).  torch runs it as ~9 foreach kernels per group over 535 tensors (~0.55 ms per step at
// HSIMAE-Large, 8 % of it launch gaps); here ONE launch walks a device-resident job table (one entry per parameter,
// 4096-element tiles) and applies, element by element and in torch's operation order,
//     p *= 1 - lr*wd;  m += (1-b1)(g-m);  v = v*b2 + (1-b2) g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// (the scalars arrive as floats of the host's double-precision expressions, exactly what torch hands its kernels)
// reading p, g, m, v once and writing p, m, v once (28 B per element: HBM-bound).
#include "kernels.cuh"
#include "../../include/hsimae_b200.h"

namespace hsimae {

namespace {

constexpr int kOptTile = 4096;

struct OptJob {          // mirrored by hsimae_b200/optim.py (48 bytes)
  float* p;
  const float* g;
  float* m;
  float* v;
  int32_t n;
  float reserved;        // (was the baked-in decay: the table now depends on pointers only, the decay is a launch scalar)
  int32_t tile0;
  int32_t pad;
};

__device__ __forceinline__ void adamw1(float& p, float g, float& m, float& v, float decay, float w1, float b2, float w2, float eps,
                                       float inv_bc2s, float neg_step) {
  p = p * decay;                              // _foreach_mul_(params, 1 - lr * wd)
  m = fmaf(w1, g - m, m);                     // _foreach_lerp_(exp_avgs, grads, 1 - beta1)
  v = fmaf(w2 * g, g, v * b2);                // _foreach_mul_(v, beta2); _foreach_addcmul_(v, g, g, 1 - beta2)
  const float denom = sqrtf(v) / inv_bc2s + eps;   // sqrt(v) / sqrt(bias_correction2) + eps   (inv_bc2s holds sqrt(bc2))
  p = fmaf(neg_step, m / denom, p);           // _foreach_addcdiv_(params, m, denom, -lr / bias_correction1)
}

__global__ void __launch_bounds__(256)
adamw_kernel(const OptJob* __restrict__ jobs, const int* __restrict__ tile_job, float decay, float w1, float b2, float w2, float eps,
             float bc2_sqrt, float neg_step) {
  const OptJob j = jobs[tile_job[blockIdx.x]];
  const int e0 = (blockIdx.x - j.tile0) * kOptTile;
  const int e1 = e0 + kOptTile < j.n ? e0 + kOptTile : j.n;
  const bool vec = ((reinterpret_cast<uintptr_t>(j.p) | reinterpret_cast<uintptr_t>(j.g) | reinterpret_cast<uintptr_t>(j.m) |
                     reinterpret_cast<uintptr_t>(j.v)) & 15) == 0;
  if (vec) {
    const int q1 = e0 + ((e1 - e0) & ~3);
    for (int i = e0 + threadIdx.x * 4; i < q1; i += blockDim.x * 4) {
      float4 p = *reinterpret_cast<float4*>(j.p + i), m = *reinterpret_cast<float4*>(j.m + i), v = *reinterpret_cast<float4*>(j.v + i);
      const float4 g = *reinterpret_cast<const float4*>(j.g + i);
      adamw1(p.x, g.x, m.x, v.x, decay, w1, b2, w2, eps, bc2_sqrt, neg_step);
      adamw1(p.y, g.y, m.y, v.y, decay, w1, b2, w2, eps, bc2_sqrt, neg_step);
      adamw1(p.z, g.z, m.z, v.z, decay, w1, b2, w2, eps, bc2_sqrt, neg_step);
      adamw1(p.w, g.w, m.w, v.w, decay, w1, b2, w2, eps, bc2_sqrt, neg_step);
      *reinterpret_cast<float4*>(j.p + i) = p; *reinterpret_cast<float4*>(j.m + i) = m; *reinterpret_cast<float4*>(j.v + i) = v;
    }
    for (int i = q1 + threadIdx.x; i < e1; i += blockDim.x) {
      float p = j.p[i], m = j.m[i], v = j.v[i];
      adamw1(p, j.g[i], m, v, decay, w1, b2, w2, eps, bc2_sqrt, neg_step);
      j.p[i] = p; j.m[i] = m; j.v[i] = v;
    }
  } else {
    for (int i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
      float p = j.p[i], m = j.m[i], v = j.v[i];
      adamw1(p, j.g[i], m, v, decay, w1, b2, w2, eps, bc2_sqrt, neg_step);
      j.p[i] = p; j.m[i] = m; j.v[i] = v;
    }
  }
}

}  // namespace

}  // namespace hsimae

extern "C" int32_t hsimae_adamw_tile_elems(void) { return hsimae::kOptTile; }

extern "C" int hsimae_adamw_step(const void* jobs, int32_t njobs, int32_t ntiles, float decay, float one_minus_beta1, float beta2,
                                 float one_minus_beta2, float eps, float bias_correction2_sqrt, float neg_step_size, void* stream) {
  using namespace hsimae;
  if (njobs == 0 || ntiles == 0) return kOk;
  HS_REQUIRE(jobs != nullptr && njobs > 0 && ntiles > 0, "adamw: bad job table");
  HS_REQUIRE(bias_correction2_sqrt > 0.f, "adamw: bias correction must be positive (step >= 1)");
  const OptJob* j = static_cast<const OptJob*>(jobs);
  adamw_kernel<<<ntiles, 256, 0, (cudaStream_t)stream>>>(j, reinterpret_cast<const int*>(j + njobs), decay, one_minus_beta1, beta2,
                                                         one_minus_beta2, eps, bias_correction2_sqrt, neg_step_size);
  HS_CHECK_LAUNCH("adamw_kernel");
  return kOk;
}
