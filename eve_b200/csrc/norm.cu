// Instance normalisation (per sample, per channel, over H x W), NHWC fp32.
//
// Reference: nn.InstanceNorm2d(affine=False) inside torchvision's ResNet (eye_net.py:50)
// and nn.InstanceNorm2d(affine=True) in RefineNet's pre-activation blocks
// (refine_net.py:46,50,59,215); biased variance, eps = 1e-5, never running statistics.
// HBM-bound: every kernel reads NHWC rows with 128-byte coalesced warps
// (a warp = 32 consecutive channels of one pixel) and reduces with shuffles / smem.
#include "common.cuh"

namespace eve {

namespace {

constexpr float kEps = 1e-5f;

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  return v;
}
// derivative from the saved forward output (sign(y) == sign(pre-activation))
__device__ __forceinline__ float act_grad(float y, int act) {
  if (act == ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == ACT_LEAKY) return y > 0.f ? 1.f : 0.01f;
  return 1.f;
}

// grid (N, C/CB); block = CB channels x L pixel lanes (CB*L = 256)
// Shifted sums (shift = first pixel) keep the variance accurate when |mean| >> std.
__global__ void __launch_bounds__(256) in_stats_kernel(const float* __restrict__ x, int HW, int C,
                                                       int CB, float* __restrict__ mean,
                                                       float* __restrict__ rstd) {
  __shared__ float s1[256], s2[256];
  const int n = blockIdx.x;
  const int cl = threadIdx.x % CB;
  const int c = blockIdx.y * CB + cl;
  const int lane = threadIdx.x / CB;
  const int L = 256 / CB;
  const float* xp = x + (size_t)n * HW * C;
  float a = 0.f, b = 0.f, shift = 0.f;
  if (c < C) {
    shift = __ldg(xp + c);
    // 4 independent accumulators for ILP
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    int p = lane;
    for (; p + L < HW; p += 2 * L) {
      float v0 = __ldg(xp + (size_t)p * C + c) - shift;
      float v1 = __ldg(xp + (size_t)(p + L) * C + c) - shift;
      a0 += v0; b0 = fmaf(v0, v0, b0);
      a1 += v1; b1 = fmaf(v1, v1, b1);
    }
    if (p < HW) {
      float v0 = __ldg(xp + (size_t)p * C + c) - shift;
      a0 += v0; b0 = fmaf(v0, v0, b0);
    }
    a = a0 + a1;
    b = b0 + b1;
  }
  s1[threadIdx.x] = a;
  s2[threadIdx.x] = b;
  __syncthreads();
  if (lane == 0 && c < C) {
    for (int l = 1; l < L; ++l) {
      a += s1[l * CB + cl];
      b += s2[l * CB + cl];
    }
    float inv = 1.f / (float)HW;
    float m = a * inv;
    float var = fmaxf(b * inv - m * m, 0.f);
    mean[(size_t)n * C + c] = m + shift;
    rstd[(size_t)n * C + c] = rsqrtf(var + kEps);
  }
}

// One thread per 4 channels of one pixel.
__global__ void __launch_bounds__(256)
in_apply_kernel(const float* __restrict__ x, long long total4, int HW, int C,
                const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ res, const float* __restrict__ res_mean,
                const float* __restrict__ res_rstd, int act, float* __restrict__ y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long pix = i / C4;
  int n = (int)(pix / HW);
  float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  float4 m = *reinterpret_cast<const float4*>(mean + (size_t)n * C + c);
  float4 r = *reinterpret_cast<const float4*>(rstd + (size_t)n * C + c);
  float o[4] = {(v.x - m.x) * r.x, (v.y - m.y) * r.y, (v.z - m.z) * r.z, (v.w - m.w) * r.w};
  if (gamma) {
    float4 g = *reinterpret_cast<const float4*>(gamma + c);
    float4 b = *reinterpret_cast<const float4*>(beta + c);
    o[0] = fmaf(o[0], g.x, b.x); o[1] = fmaf(o[1], g.y, b.y);
    o[2] = fmaf(o[2], g.z, b.z); o[3] = fmaf(o[3], g.w, b.w);
  }
  if (res) {
    float4 q = __ldg(reinterpret_cast<const float4*>(res) + i);
    if (res_mean) {
      float4 qm = *reinterpret_cast<const float4*>(res_mean + (size_t)n * C + c);
      float4 qr = *reinterpret_cast<const float4*>(res_rstd + (size_t)n * C + c);
      q.x = (q.x - qm.x) * qr.x; q.y = (q.y - qm.y) * qr.y;
      q.z = (q.z - qm.z) * qr.z; q.w = (q.w - qm.w) * qr.w;
    }
    o[0] += q.x; o[1] += q.y; o[2] += q.z; o[3] += q.w;
  }
  float4 out = make_float4(act_fwd(o[0], act), act_fwd(o[1], act), act_fwd(o[2], act),
                           act_fwd(o[3], act));
  reinterpret_cast<float4*>(y)[i] = out;
}

// Backward reductions per (n,c): sum g and sum g*xhat with g = dy*act'(y).
// grid (N, C/CB); same thread layout as in_stats_kernel.
__global__ void __launch_bounds__(256)
in_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ ymask,
                     const float* __restrict__ x, int HW, int C, int CB,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                     float* __restrict__ sum_g, float* __restrict__ sum_gx) {
  // fp64 accumulators: sum(g*xhat) cancels heavily (it is a covariance), and the affine
  // gradients are sums of these over the batch; B200 has the fp64 rate to hide this behind
  // the HBM reads (2 DFMA per 8-12 bytes loaded).
  __shared__ double s1[256], s2[256];
  const int n = blockIdx.x;
  const int cl = threadIdx.x % CB;
  const int c = blockIdx.y * CB + cl;
  const int lane = threadIdx.x / CB;
  const int L = 256 / CB;
  const size_t base = (size_t)n * HW * C;
  double a = 0.0, b = 0.0;
  if (c < C) {
    const float m = mean[(size_t)n * C + c], r = rstd[(size_t)n * C + c];
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    for (int p = lane; p < HW; p += L) {
      size_t o = base + (size_t)p * C + c;
      float g = __ldg(dy + o);
      float xh = (__ldg(x + o) - m) * r;
      if (act != ACT_NONE) g *= act_grad(ymask ? __ldg(ymask + o) : fmaf(xh, ga, be), act);
      a += (double)g;
      b += (double)g * (double)xh;
    }
  }
  s1[threadIdx.x] = a;
  s2[threadIdx.x] = b;
  __syncthreads();
  if (lane == 0 && c < C) {
    for (int l = 1; l < L; ++l) {
      a += s1[l * CB + cl];
      b += s2[l * CB + cl];
    }
    sum_g[(size_t)n * C + c] = (float)a;
    sum_gx[(size_t)n * C + c] = (float)b;
  }
}

__global__ void __launch_bounds__(256)
in_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ ymask,
                    const float* __restrict__ x, long long total4, int HW, int C,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ sum_g, const float* __restrict__ sum_gx, int act,
                    const float* __restrict__ addend, float* __restrict__ dx,
                    float* __restrict__ g_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long pix = i / C4;
  int n = (int)(pix / HW);
  float4 d = __ldg(reinterpret_cast<const float4*>(dy) + i);
  float g[4] = {d.x, d.y, d.z, d.w};
  float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
  float xs[4] = {xv.x, xv.y, xv.z, xv.w};
  float ym[4] = {0.f, 0.f, 0.f, 0.f};
  if (act != ACT_NONE && ymask) {
    float4 t = __ldg(reinterpret_cast<const float4*>(ymask) + i);
    ym[0] = t.x; ym[1] = t.y; ym[2] = t.z; ym[3] = t.w;
  }
  const float inv = 1.f / (float)HW;
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    size_t sc = (size_t)n * C + c + j;
    float r = rstd[sc];
    float xh = (xs[j] - mean[sc]) * r;
    float ga = gamma ? gamma[c + j] : 1.f;
    if (act != ACT_NONE)
      g[j] *= act_grad(ymask ? ym[j] : fmaf(xh, ga, beta ? beta[c + j] : 0.f), act);
    // sums were taken over g (without gamma); gamma scales all three terms equally
    o[j] = r * ga * (g[j] - sum_g[sc] * inv - xh * sum_gx[sc] * inv);
  }
  if (g_out) reinterpret_cast<float4*>(g_out)[i] = make_float4(g[0], g[1], g[2], g[3]);
  if (addend) {
    float4 a = __ldg(reinterpret_cast<const float4*>(addend) + i);
    o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
  }
  reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
}

// dgamma[c] (+)= sum_n sum_gx[n,c]; dbeta[c] (+)= sum_n sum_g[n,c]
__global__ void in_affine_grad_kernel(const float* __restrict__ sum_g,
                                      const float* __restrict__ sum_gx, int N, int C,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                      int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int n = 0; n < N; ++n) {
    a += (double)sum_gx[(size_t)n * C + c];
    b += (double)sum_g[(size_t)n * C + c];
  }
  dgamma[c] = accumulate ? dgamma[c] + (float)a : (float)a;
  dbeta[c] = accumulate ? dbeta[c] + (float)b : (float)b;
}

inline int chan_block(int C) { return C >= 32 ? 32 : C; }

}  // namespace

int in_stats(const float* x, int N, int HW, int C, float* mean, float* rstd, cudaStream_t s) {
  int CB = chan_block(C);
  EVE_REQUIRE(256 % CB == 0, EVE_ERR_SHAPE, "in_stats: C=%d unsupported", C);
  dim3 grid(N, cdiv(C, CB));
  in_stats_kernel<<<grid, 256, 0, s>>>(x, HW, C, CB, mean, rstd);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int in_apply(const float* x, int N, int HW, int C, const float* mean, const float* rstd,
             const float* gamma, const float* beta, const float* res, const float* res_mean,
             const float* res_rstd, int act, float* y, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "in_apply: C=%d must be a multiple of 4", C);
  long long total4 = (long long)N * HW * C / 4;
  in_apply_kernel<<<cdiv(total4, 256), 256, 0, s>>>(x, total4, HW, C, mean, rstd, gamma, beta, res,
                                                    res_mean, res_rstd, act, y);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

size_t in_backward_scratch_floats(int N, int C) { return (size_t)2 * N * C; }

int in_backward(const float* dy, const float* y_for_mask, const float* x, int N, int HW, int C,
                const float* mean, const float* rstd, const float* gamma, const float* beta,
                int act, const float* addend, float* dx, float* g_out, float* dgamma,
                float* dbeta, float* scratch, bool accumulate_affine, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "in_backward: C=%d must be a multiple of 4", C);
  int CB = chan_block(C);
  float* sum_g = scratch;
  float* sum_gx = scratch + (size_t)N * C;
  dim3 grid(N, cdiv(C, CB));
  in_bwd_reduce_kernel<<<grid, 256, 0, s>>>(dy, y_for_mask, x, HW, C, CB, mean, rstd, gamma, beta,
                                            act, sum_g, sum_gx);
  EVE_LAUNCH_CHECK();
  long long total4 = (long long)N * HW * C / 4;
  in_bwd_apply_kernel<<<cdiv(total4, 256), 256, 0, s>>>(dy, y_for_mask, x, total4, HW, C, mean,
                                                        rstd, gamma, beta, sum_g, sum_gx, act,
                                                        addend, dx, g_out);
  EVE_LAUNCH_CHECK();
  if (gamma && dgamma) {
    in_affine_grad_kernel<<<cdiv(C, 128), 128, 0, s>>>(sum_g, sum_gx, N, C, dgamma, dbeta,
                                                       accumulate_affine ? 1 : 0);
    EVE_LAUNCH_CHECK();
  }
  return EVE_OK;
}

}  // namespace eve
