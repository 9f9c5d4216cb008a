// Fused gradient-clip + Adam step on one flat fp32 buffer.
//
// Reference: torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0) at
// src/core/training.py:492-498 followed by optim.Adam(weight_decay=...).step() at
// src/core/training.py:501-502 (optimizer built at src/train.py:49-55).  The multi-GPU step
// runs this directly on the buffer the NCCL allreduce summed (grad_scale = 1/world_size).
#include <cmath>

#include "common.cuh"

namespace eve {
namespace {

constexpr int kNormBlocks = 148 * 4;

__global__ void __launch_bounds__(256)
sqnorm_partial_kernel(const float* __restrict__ g, long long n, float scale,
                      float* __restrict__ part) {
  __shared__ float sm[8];
  float a = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(g4 + i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float v = g[i] * scale;
      a += v * v;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += sm[i];
    part[blockIdx.x] = s;
  }
}

// one block: norm = sqrt(sum partials) (deterministic order); coef = min(1, max_norm/(norm+1e-6))
__global__ void norm_final_kernel(const float* __restrict__ part, int nparts, float max_norm,
                                  float* __restrict__ stat, float* __restrict__ norm_out,
                                  int* __restrict__ step_dev, int step_host, float beta1,
                                  float beta2, const float* __restrict__ lr_dev, float lr_host) {
  __shared__ double sm[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) a += (double)part[i];
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float norm = (float)sqrt(sm[0]);
    float coef = 1.f;
    if (max_norm > 0.f) {
      coef = max_norm / (norm + 1e-6f);
      if (coef > 1.f) coef = 1.f;
    }
    stat[0] = norm;
    stat[1] = coef;
    if (norm_out) norm_out[0] = norm;
    // bias corrections for this step (device-side counter when the step runs from a graph)
    int step = step_host;
    if (step_dev) {
      step = *step_dev + 1;
      *step_dev = step;
    }
    stat[2] = (float)(1.0 - pow((double)beta1, (double)step));
    stat[3] = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    // learning rate of this step: a device scalar when a scheduler drives a captured graph
    stat[4] = lr_dev ? *lr_dev : lr_host;
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, const float* __restrict__ stat, float scale,
            float b1, float b2, float eps, float wd) {
  const float coef = stat[1] * scale;
  const float bc1 = stat[2], bc2_sqrt = stat[3], lr = stat[4];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float pv = p[i];
    float gv = fmaf(wd, pv, g[i] * coef);
    float mv = fmaf(b1, m[i], (1.f - b1) * gv);
    float vv = fmaf(b2, v[i], (1.f - b2) * gv * gv);
    m[i] = mv;
    v[i] = vv;
    float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] = pv - (lr / bc1) * (mv / denom);
  }
}

}  // namespace
}  // namespace eve

using namespace eve;

extern "C" size_t eve_adam_clip_workspace_bytes(const eve_adam_params* p) {
  (void)p;
  return (kNormBlocks + 64) * sizeof(float);
}

extern "C" int eve_adam_clip_step(const eve_adam_params* p, float* params, const float* grads,
                                  float* exp_avg, float* exp_avg_sq, float* norm_out,
                                  void* workspace, size_t workspace_bytes, eve_stream_t stream) {
  EVE_REQUIRE(p, EVE_ERR_NULL, "adam_clip_step: params is NULL");
  EVE_REQUIRE(p->count >= 0 && (p->step >= 1 || p->step_dev), EVE_ERR_SHAPE,
              "adam_clip_step: bad count/step");
  if (p->count == 0) return EVE_OK;
  EVE_REQUIRE(params && grads && exp_avg && exp_avg_sq && workspace, EVE_ERR_NULL,
              "adam_clip_step: NULL pointer");
  EVE_REQUIRE(workspace_bytes >= eve_adam_clip_workspace_bytes(p), EVE_ERR_WORKSPACE,
              "adam_clip_step: workspace too small");
  cudaStream_t s = as_stream(stream);
  float* part = (float*)workspace;
  float* stat = part + kNormBlocks;
  sqnorm_partial_kernel<<<kNormBlocks, 256, 0, s>>>(grads, p->count, p->grad_scale, part);
  EVE_LAUNCH_CHECK();
  norm_final_kernel<<<1, 256, 0, s>>>(part, kNormBlocks, p->max_norm, stat, norm_out, p->step_dev,
                                      p->step, p->beta1, p->beta2, p->lr_dev, p->lr);
  EVE_LAUNCH_CHECK();
  adam_kernel<<<kNormBlocks * 2, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, p->count, stat,
                                              p->grad_scale, p->beta1, p->beta2, p->eps,
                                              p->weight_decay);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
