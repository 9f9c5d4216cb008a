// Instance normalisation (per sample, per channel, over H x W), NHWC fp32.
//
// Reference: nn.InstanceNorm2d(affine=False) inside torchvision's ResNet (eye_net.py:50)
// and nn.InstanceNorm2d(affine=True) in RefineNet's pre-activation blocks
// (refine_net.py:46,50,59,215); biased variance, eps = 1e-5, never running statistics.
// HBM-bound: every kernel reads NHWC rows with 128-byte coalesced warps
// (a warp = 32 consecutive channels of one pixel) and reduces with shuffles / smem.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

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

// grid (N, ceil(C4/CQ)); block 256 = CQ channel-quads x L pixel lanes.  Every thread owns 4
// consecutive channels (one 16-byte load per pixel) and walks pixels lane, lane+L, ... with two
// loads in flight.  Shifted sums (shift = first pixel) keep the variance accurate when
// |mean| >> std.
__global__ void __launch_bounds__(256) in_stats_kernel(const float* __restrict__ x, int HW, int C,
                                                       int CQ, float* __restrict__ mean,
                                                       float* __restrict__ rstd) {
  __shared__ float4 s1[256], s2[256];
  const int n = blockIdx.x;
  const int ql = threadIdx.x % CQ;
  const int q = blockIdx.y * CQ + ql;          // channel quad
  const int lane = threadIdx.x / CQ;
  const int L = 256 / CQ;
  const int C4 = C >> 2;
  const float4* xp = reinterpret_cast<const float4*>(x + (size_t)n * HW * C);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, sh = a;
  if (q < C4) {
    sh = __ldg(xp + q);
    // four independent loads and partial sums per thread (a wide channel group leaves few blocks
    // per SM: the bytes in flight come from the unroll)
    float4 pa[4], pb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) pa[u] = pb[u] = a;
    auto acc = [&](float4 v, float4& sa, float4& sb) {
      v.x -= sh.x; v.y -= sh.y; v.z -= sh.z; v.w -= sh.w;
      sa.x += v.x; sa.y += v.y; sa.z += v.z; sa.w += v.w;
      sb.x = fmaf(v.x, v.x, sb.x); sb.y = fmaf(v.y, v.y, sb.y);
      sb.z = fmaf(v.z, v.z, sb.z); sb.w = fmaf(v.w, v.w, sb.w);
    };
    int p = lane;
    for (; p + 3 * L < HW; p += 4 * L) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(xp + (size_t)(p + u * L) * C4 + q);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc(v[u], pa[u], pb[u]);
    }
#pragma unroll
    for (int u = 0; u < 3; ++u)            // at most three pixels are left
      if (p + u * L < HW) acc(__ldg(xp + (size_t)(p + u * L) * C4 + q), pa[u], pb[u]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      pa[u].x += pa[u + 2].x; pa[u].y += pa[u + 2].y; pa[u].z += pa[u + 2].z; pa[u].w += pa[u + 2].w;
      pb[u].x += pb[u + 2].x; pb[u].y += pb[u + 2].y; pb[u].z += pb[u + 2].z; pb[u].w += pb[u + 2].w;
    }
    a.x = pa[0].x + pa[1].x; a.y = pa[0].y + pa[1].y; a.z = pa[0].z + pa[1].z; a.w = pa[0].w + pa[1].w;
    b.x = pb[0].x + pb[1].x; b.y = pb[0].y + pb[1].y; b.z = pb[0].z + pb[1].z; b.w = pb[0].w + pb[1].w;
  }
  s1[threadIdx.x] = a;
  s2[threadIdx.x] = b;
  __syncthreads();
  // pairwise tree over the L pixel lanes (L is a power of two): log2(L) steps, fixed order
  for (int half = L >> 1; half >= 1; half >>= 1) {
    if (lane < half) {
      const float4 t = s1[(lane + half) * CQ + ql], u = s2[(lane + half) * CQ + ql];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      b.x += u.x; b.y += u.y; b.z += u.z; b.w += u.w;
      s1[threadIdx.x] = a;
      s2[threadIdx.x] = b;
    }
    __syncthreads();
  }
  if (lane == 0 && q < C4) {
    const float inv = 1.f / (float)HW;
    float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    float sv[4] = {sh.x, sh.y, sh.z, sh.w};
    float mo[4], ro[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float m = av[j] * inv;
      float var = fmaxf(bv[j] * inv - m * m, 0.f);
      mo[j] = m + sv[j];
      ro[j] = rsqrtf(var + kEps);
    }
    reinterpret_cast<float4*>(mean + (size_t)n * C)[q] = make_float4(mo[0], mo[1], mo[2], mo[3]);
    reinterpret_cast<float4*>(rstd + (size_t)n * C)[q] = make_float4(ro[0], ro[1], ro[2], ro[3]);
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

// in_apply_kernel without a residual, emitting the result as the 16-bit hi/lo operand planes of
// the convolution that consumes it (x = hi + lo; fp16 planes forward, bf16 planes backward) and,
// optionally, as fp32.  Saves the separate split pass (one read + one launch per convolution) and
// lets the forward pass skip the fp32 activation altogether.
template <int FMT>
__global__ void __launch_bounds__(256)
in_apply_planes_kernel(const float* __restrict__ x, long long total4, int HW, int C,
                       const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                       float* __restrict__ y, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
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
  uint16_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[j] = act_fwd(o[j], act);
    if (FMT == TC_BF16) {
      __nv_bfloat16 hb = __float2bfloat16_rn(o[j]);
      h[j] = __bfloat16_as_ushort(hb);
      l[j] = __bfloat16_as_ushort(__float2bfloat16_rn(o[j] - __bfloat162float(hb)));
    } else {
      __half hh = __float2half_rn(o[j]);
      h[j] = __half_as_ushort(hh);
      l[j] = __half_as_ushort(__float2half_rn(o[j] - __half2float(hh)));
    }
  }
  if (y) reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
  reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
  reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
}

// Two affine sets from one read of x (planes only).
template <int FMT>
__global__ void __launch_bounds__(256)
in_apply_planes2_kernel(const float* __restrict__ x, long long total4, int HW, int C,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ gammaB, const float* __restrict__ betaB, int act,
                        uint16_t* __restrict__ hiA, uint16_t* __restrict__ loA,
                        uint16_t* __restrict__ hiB, uint16_t* __restrict__ loB) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long pix = i / C4;
  int n = (int)(pix / HW);
  float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  float4 m = *reinterpret_cast<const float4*>(mean + (size_t)n * C + c);
  float4 r = *reinterpret_cast<const float4*>(rstd + (size_t)n * C + c);
  const float xh[4] = {(v.x - m.x) * r.x, (v.y - m.y) * r.y, (v.z - m.z) * r.z, (v.w - m.w) * r.w};
#pragma unroll
  for (int set = 0; set < 2; ++set) {
    const float* gp = set == 0 ? gamma : gammaB;
    const float* bp = set == 0 ? beta : betaB;
    uint16_t* hi = set == 0 ? hiA : hiB;
    uint16_t* lo = set == 0 ? loA : loB;
    if (!hi) continue;
    float o[4] = {xh[0], xh[1], xh[2], xh[3]};
    if (gp) {
      float4 g = *reinterpret_cast<const float4*>(gp + c);
      float4 b = *reinterpret_cast<const float4*>(bp + c);
      o[0] = fmaf(o[0], g.x, b.x); o[1] = fmaf(o[1], g.y, b.y);
      o[2] = fmaf(o[2], g.z, b.z); o[3] = fmaf(o[3], g.w, b.w);
    }
    uint16_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = act_fwd(o[j], act);
      if (FMT == TC_BF16) {
        __nv_bfloat16 hb = __float2bfloat16_rn(o[j]);
        h[j] = __bfloat16_as_ushort(hb);
        l[j] = __bfloat16_as_ushort(__float2bfloat16_rn(o[j] - __bfloat162float(hb)));
      } else {
        __half hh = __float2half_rn(o[j]);
        h[j] = __half_as_ushort(hh);
        l[j] = __half_as_ushort(__float2half_rn(o[j] - __half2float(hh)));
      }
    }
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
    reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
  }
}

// Backward reductions per (n,c): sum g and sum g*xhat with g = dy*act'(y).
// grid (N, C/CB); same thread layout as in_stats_kernel.
__global__ void __launch_bounds__(256)
in_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ ymask,
                     const float* __restrict__ x, int HW, int C, int CQ,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                     float* __restrict__ sum_g, float* __restrict__ sum_gx) {
  // Same thread layout as in_stats_kernel (4 channels per thread, 16-byte loads).  fp64
  // accumulators: sum(g*xhat) cancels heavily (it is a covariance), and the affine
  // gradients are sums of these over the batch; B200 has the fp64 rate to hide this behind
  // the HBM reads (8 DFMA per 32-48 bytes loaded).
  __shared__ double s1[256 * 4];
  __shared__ double s2[256 * 4];
  const int n = blockIdx.x;
  const int ql = threadIdx.x % CQ;
  const int q = blockIdx.y * CQ + ql;
  const int lane = threadIdx.x / CQ;
  const int L = 256 / CQ;
  const int C4 = C >> 2;
  const size_t base4 = (size_t)n * HW * C4;
  const float4* dp = reinterpret_cast<const float4*>(dy) + base4;
  const float4* xp = reinterpret_cast<const float4*>(x) + base4;
  const float4* yp = ymask ? reinterpret_cast<const float4*>(ymask) + base4 : nullptr;
  double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
  if (q < C4) {
    const float4 m4 = reinterpret_cast<const float4*>(mean + (size_t)n * C)[q];
    const float4 r4 = reinterpret_cast<const float4*>(rstd + (size_t)n * C)[q];
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
    float ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
    if (gamma) {
      float4 g4 = reinterpret_cast<const float4*>(gamma)[q];
      ga[0] = g4.x; ga[1] = g4.y; ga[2] = g4.z; ga[3] = g4.w;
    }
    if (beta) {
      float4 b4 = reinterpret_cast<const float4*>(beta)[q];
      be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
    }
    auto body = [&](const float4& d4, const float4& x4, const float4& y4) {
      const float d[4] = {d4.x, d4.y, d4.z, d4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
      const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float xh = (xv[j] - m[j]) * r[j];
        float g = d[j];
        if (act != ACT_NONE) g *= act_grad(yp ? yv[j] : fmaf(xh, ga[j], be[j]), act);
        a[j] += (double)g;
        b[j] += (double)g * (double)xh;
      }
    };
    int p = lane;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (; p + L < HW; p += 2 * L) {
      const size_t o0 = (size_t)p * C4 + q, o1 = (size_t)(p + L) * C4 + q;
      float4 d0 = __ldg(dp + o0), d1 = __ldg(dp + o1);
      float4 x0 = __ldg(xp + o0), x1 = __ldg(xp + o1);
      float4 y0 = yp ? __ldg(yp + o0) : z4, y1 = yp ? __ldg(yp + o1) : z4;
      body(d0, x0, y0);
      body(d1, x1, y1);
    }
    if (p < HW) {
      const size_t o0 = (size_t)p * C4 + q;
      body(__ldg(dp + o0), __ldg(xp + o0), yp ? __ldg(yp + o0) : z4);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s1[threadIdx.x * 4 + j] = a[j];
    s2[threadIdx.x * 4 + j] = b[j];
  }
  __syncthreads();
  for (int half = L >> 1; half >= 1; half >>= 1) {
    if (lane < half) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] += s1[((lane + half) * CQ + ql) * 4 + j];
        b[j] += s2[((lane + half) * CQ + ql) * 4 + j];
        s1[threadIdx.x * 4 + j] = a[j];
        s2[threadIdx.x * 4 + j] = b[j];
      }
    }
    __syncthreads();
  }
  if (lane == 0 && q < C4) {
    reinterpret_cast<float4*>(sum_g + (size_t)n * C)[q] =
        make_float4((float)a[0], (float)a[1], (float)a[2], (float)a[3]);
    reinterpret_cast<float4*>(sum_gx + (size_t)n * C)[q] =
        make_float4((float)b[0], (float)b[1], (float)b[2], (float)b[3]);
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
// one warp per channel, lanes stride over n (fixed order: deterministic), fp64 partials
__global__ void __launch_bounds__(256)
in_affine_grad_kernel(const float* __restrict__ sum_g, const float* __restrict__ sum_gx, int N,
                      int C, float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int n = lane; n < N; n += 32) {
    a += (double)sum_gx[(size_t)n * C + c];
    b += (double)sum_g[(size_t)n * C + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    dgamma[c] = accumulate ? dgamma[c] + (float)a : (float)a;
    dbeta[c] = accumulate ? dbeta[c] + (float)b : (float)b;
  }
}

// ---- backward of  p = maxpool3x3s2p1(relu(IN(x)))  without the dense un-pooled gradient map.
// Only the argmax pixel of a pool window carries gradient, ReLU passes it iff the pooled value is
// positive, and the normalised value at that pixel IS the pooled value (no affine terms), so
//     sum_pixels g = sum_windows dp [p > 0]           sum_pixels g xhat = sum_windows dp p
// : the reductions read the two pooled tensors (a quarter of the map each) instead of the map
// and its gradient.  Same thread layout and fp64 accumulation as in_bwd_reduce_kernel.
__global__ void __launch_bounds__(256)
stem_pool_sums_kernel(const float* __restrict__ dp, const float* __restrict__ pooled, int OHW, int C,
                      int CQ, float* __restrict__ sum_g, float* __restrict__ sum_gx) {
  __shared__ double s1[256 * 4];
  __shared__ double s2[256 * 4];
  const int n = blockIdx.x;
  const int ql = threadIdx.x % CQ;
  const int q = blockIdx.y * CQ + ql;
  const int lane = threadIdx.x / CQ;
  const int L = 256 / CQ;
  const int C4 = C >> 2;
  const size_t base4 = (size_t)n * OHW * C4;
  const float4* dq = reinterpret_cast<const float4*>(dp) + base4;
  const float4* pq = reinterpret_cast<const float4*>(pooled) + base4;
  double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
  if (q < C4) {
    auto body = [&](const float4& d4, const float4& p4) {
      const float d[4] = {d4.x, d4.y, d4.z, d4.w}, pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g = pv[j] > 0.f ? d[j] : 0.f;
        a[j] += (double)g;
        b[j] += (double)g * (double)pv[j];
      }
    };
    int w = lane;
    for (; w + 3 * L < OHW; w += 4 * L) {
      const size_t o = (size_t)w * C4 + q, st = (size_t)L * C4;
      float4 d0 = __ldg(dq + o), d1 = __ldg(dq + o + st), d2 = __ldg(dq + o + 2 * st),
             d3 = __ldg(dq + o + 3 * st);
      float4 p0 = __ldg(pq + o), p1 = __ldg(pq + o + st), p2 = __ldg(pq + o + 2 * st),
             p3 = __ldg(pq + o + 3 * st);
      body(d0, p0);
      body(d1, p1);
      body(d2, p2);
      body(d3, p3);
    }
    for (; w < OHW; w += L) {
      const size_t o = (size_t)w * C4 + q;
      body(__ldg(dq + o), __ldg(pq + o));
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s1[threadIdx.x * 4 + j] = a[j];
    s2[threadIdx.x * 4 + j] = b[j];
  }
  __syncthreads();
  for (int half = L >> 1; half >= 1; half >>= 1) {
    if (lane < half) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] += s1[((lane + half) * CQ + ql) * 4 + j];
        b[j] += s2[((lane + half) * CQ + ql) * 4 + j];
        s1[threadIdx.x * 4 + j] = a[j];
        s2[threadIdx.x * 4 + j] = b[j];
      }
    }
    __syncthreads();
  }
  if (lane == 0 && q < C4) {
    reinterpret_cast<float4*>(sum_g + (size_t)n * C)[q] =
        make_float4((float)a[0], (float)a[1], (float)a[2], (float)a[3]);
    reinterpret_cast<float4*>(sum_gx + (size_t)n * C)[q] =
        make_float4((float)b[0], (float)b[1], (float)b[2], (float)b[3]);
  }
}

// One thread = a 2x2 pixel block x 4 channels.  The block's pixels are touched by the four
// windows (a, b), (a, b+1), (a+1, b), (a+1, b+1) only (pixel (2a+i, 2b+j) by rows a..a+i, columns
// b..b+j), so each window's (index, gradient) pair is loaded once per block instead of once per
// pixel; per pixel the matching windows are summed in the order of maxpool_bwd_kernel.
__global__ void __launch_bounds__(256)
stem_pool_in_bwd_apply_kernel(const float* __restrict__ dp, const int32_t* __restrict__ idx,
                              const float* __restrict__ x, long long total, int H, int W, int OH,
                              int OW, int C, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ sum_g,
                              const float* __restrict__ sum_gx, float* __restrict__ dx,
                              __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int C4 = C >> 2;
  const int cq = (int)(i % C4);
  long long t = i / C4;
  const int W2 = W >> 1, H2 = H >> 1;
  const int b = (int)(t % W2);
  t /= W2;
  const int a = (int)(t % H2);
  const int n = (int)(t / H2);
  float wd[2][2][4];
  int wi[2][2][4];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int oy = a + u, ox = b + v;
      if (oy < OH && ox < OW) {
        const size_t o = ((((size_t)n * OH + oy) * OW + ox) * C4 + cq);
        const int4 id = __ldg(reinterpret_cast<const int4*>(idx) + o);
        const float4 d = __ldg(reinterpret_cast<const float4*>(dp) + o);
        wi[u][v][0] = id.x; wi[u][v][1] = id.y; wi[u][v][2] = id.z; wi[u][v][3] = id.w;
        wd[u][v][0] = d.x; wd[u][v][1] = d.y; wd[u][v][2] = d.z; wd[u][v][3] = d.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          wi[u][v][j] = -1;
          wd[u][v][j] = 0.f;
        }
      }
    }
  const size_t sc4 = ((size_t)n * C4 + cq);
  const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean) + sc4);
  const float4 r4 = __ldg(reinterpret_cast<const float4*>(rstd) + sc4);
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(sum_g) + sc4);
  const float4 x4 = __ldg(reinterpret_cast<const float4*>(sum_gx) + sc4);
  const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
  const float sg[4] = {g4.x, g4.y, g4.z, g4.w}, sx[4] = {x4.x, x4.y, x4.z, x4.w};
  const float inv = 1.f / (float)(H * W);
  float4 xin[2][2];
#pragma unroll
  for (int ph = 0; ph < 2; ++ph)
#pragma unroll
    for (int pw = 0; pw < 2; ++pw) {
      const size_t o = (((size_t)n * H + 2 * a + ph) * W + 2 * b + pw) * C4 + cq;
      xin[ph][pw] = __ldg(reinterpret_cast<const float4*>(x) + o);
    }
#pragma unroll
  for (int ph = 0; ph < 2; ++ph)
#pragma unroll
    for (int pw = 0; pw < 2; ++pw) {
      const int me = (2 * a + ph) * W + 2 * b + pw;
      const size_t o = (((size_t)n * H + 2 * a + ph) * W + 2 * b + pw) * C4 + cq;
      const float xs[4] = {xin[ph][pw].x, xin[ph][pw].y, xin[ph][pw].z, xin[ph][pw].w};
      float out[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float g = 0.f;
#pragma unroll
        for (int u = 0; u <= ph; ++u)
#pragma unroll
          for (int v = 0; v <= pw; ++v)
            if (wi[u][v][j] == me) g += wd[u][v][j];
        const float xh = (xs[j] - m[j]) * r[j];
        g *= act_grad(xh, ACT_RELU);
        out[j] = r[j] * (g - sg[j] * inv - xh * sx[j] * inv);
      }
      if (dx) reinterpret_cast<float4*>(dx)[o] = make_float4(out[0], out[1], out[2], out[3]);
      if (d_hi) {
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h[j] = __float2bfloat16_rn(out[j]);
          l[j] = __float2bfloat16_rn(out[j] - __bfloat162float(h[j]));
        }
        reinterpret_cast<uint2*>(d_hi)[o] = *reinterpret_cast<uint2*>(h);
        if (d_lo) reinterpret_cast<uint2*>(d_lo)[o] = *reinterpret_cast<uint2*>(l);
      }
    }
}

// channel quads per block: up to 16 (64 channels), never more than the tensor has.  Wide maps get
// narrow channel groups (down to 2 quads = one 32-byte sector per pixel) so that the grid
// (one block per image and channel group) covers the SMs several times over and each of the
// 256 / CQ pixel lanes still walks >= 32 pixels (shorter walks
// are dominated by the block prologue and the cross-lane tree).  The choice depends on (C, HW) only -- never on
// N -- so a frame's statistics are bit-identical whatever batch it is processed in.
// in_bwd_reduce (three streams, fp64 accumulators) measured fastest with the widest groups: up to
// 16 quads, i.e. whole 256-byte pixel rows per warp
inline int quad_block_wide(int C) {
  const int q = C >> 2;
  return q >= 16 ? 16 : (q >= 8 ? 8 : (q >= 4 ? 4 : (q >= 2 ? 2 : 1)));
}
inline int quad_block(int C, int HW) {
  const int q = C >> 2;
  int cq = q >= 2 ? 2 : 1;
  while (cq < 16 && cq < q && 8192 / cq > HW) cq <<= 1;    // every lane walks >= 32 pixels
  return cq;
}

}  // namespace

int in_stats(const float* x, int N, int HW, int C, float* mean, float* rstd, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "in_stats: C=%d must be a multiple of 4", C);
  // maps of 4096 pixels and more (the EyeNet stem: 64 x 64 x 64 channels) read whole 256-byte pixel
  // rows per warp: with narrow groups eight blocks pulled 32-byte pieces out of every DRAM line
  // (2.6 TB/s measured); a function of (C, HW) only, as quad_block() is
  int CQ = HW >= 4096 ? quad_block_wide(C) : quad_block(C, HW);
  dim3 grid(N, cdiv(C >> 2, CQ));
  in_stats_kernel<<<grid, 256, 0, s>>>(x, HW, C, CQ, mean, rstd);
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

int in_apply_planes(const float* x, int N, int HW, int C, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, int act, int fmt, float* y, void* hi,
                    void* lo, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "in_apply_planes: C=%d must be a multiple of 4", C);
  EVE_REQUIRE(hi && lo, EVE_ERR_NULL, "in_apply_planes: plane pointers are NULL");
  long long total4 = (long long)N * HW * C / 4;
  if (total4 == 0) return EVE_OK;
  if (fmt == TC_BF16)
    in_apply_planes_kernel<TC_BF16><<<cdiv(total4, 256), 256, 0, s>>>(
        x, total4, HW, C, mean, rstd, gamma, beta, act, y, (uint16_t*)hi, (uint16_t*)lo);
  else
    in_apply_planes_kernel<TC_F16><<<cdiv(total4, 256), 256, 0, s>>>(
        x, total4, HW, C, mean, rstd, gamma, beta, act, y, (uint16_t*)hi, (uint16_t*)lo);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int in_apply_planes2(const float* x, int N, int HW, int C, const float* mean, const float* rstd,
                     const float* gamma, const float* beta, const float* gammaB,
                     const float* betaB, int act, int fmt, void* hiA, void* loA, void* hiB,
                     void* loB, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "in_apply_planes2: C=%d must be a multiple of 4", C);
  EVE_REQUIRE(hiA && loA && (!hiB || loB), EVE_ERR_NULL, "in_apply_planes2: plane pointers");
  long long total4 = (long long)N * HW * C / 4;
  if (total4 == 0) return EVE_OK;
  if (fmt == TC_BF16)
    in_apply_planes2_kernel<TC_BF16><<<cdiv(total4, 256), 256, 0, s>>>(
        x, total4, HW, C, mean, rstd, gamma, beta, gammaB, betaB, act, (uint16_t*)hiA,
        (uint16_t*)loA, (uint16_t*)hiB, (uint16_t*)loB);
  else
    in_apply_planes2_kernel<TC_F16><<<cdiv(total4, 256), 256, 0, s>>>(
        x, total4, HW, C, mean, rstd, gamma, beta, gammaB, betaB, act, (uint16_t*)hiA,
        (uint16_t*)loA, (uint16_t*)hiB, (uint16_t*)loB);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

size_t in_backward_scratch_floats(int N, int C) { return (size_t)2 * N * C; }

int in_backward(const float* dy, const float* y_for_mask, const float* x, int N, int HW, int C,
                const float* mean, const float* rstd, const float* gamma, const float* beta,
                int act, const float* addend, float* dx, float* g_out, float* dgamma,
                float* dbeta, float* scratch, bool accumulate_affine, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "in_backward: C=%d must be a multiple of 4", C);
  int CQ = quad_block_wide(C);
  float* sum_g = scratch;
  float* sum_gx = scratch + (size_t)N * C;
  dim3 grid(N, cdiv(C >> 2, CQ));
  in_bwd_reduce_kernel<<<grid, 256, 0, s>>>(dy, y_for_mask, x, HW, C, CQ, mean, rstd, gamma, beta,
                                            act, sum_g, sum_gx);
  EVE_LAUNCH_CHECK();
  long long total4 = (long long)N * HW * C / 4;
  in_bwd_apply_kernel<<<cdiv(total4, 256), 256, 0, s>>>(dy, y_for_mask, x, total4, HW, C, mean,
                                                        rstd, gamma, beta, sum_g, sum_gx, act,
                                                        addend, dx, g_out);
  EVE_LAUNCH_CHECK();
  if (gamma && dgamma) {
    in_affine_grad_kernel<<<cdiv(C, 8), 256, 0, s>>>(sum_g, sum_gx, N, C, dgamma, dbeta,
                                                       accumulate_affine ? 1 : 0);
    EVE_LAUNCH_CHECK();
  }
  return EVE_OK;
}

int stem_pool_in_backward(const float* dpool, const float* pooled, const int32_t* idx,
                          const float* x, int N, int H, int W, int C, const float* mean,
                          const float* rstd, float* dx, uint16_t* d_hi, uint16_t* d_lo,
                          float* scratch, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, EVE_ERR_SHAPE,
              "stem_pool_in_backward: C=%d must be a multiple of 4, H=%d and W=%d even", C, H, W);
  EVE_REQUIRE(dx || d_hi, EVE_ERR_NULL, "stem_pool_in_backward: no output");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const int CQ = quad_block_wide(C);
  float* sum_g = scratch;
  float* sum_gx = scratch + (size_t)N * C;
  dim3 grid(N, cdiv(C >> 2, CQ));
  stem_pool_sums_kernel<<<grid, 256, 0, s>>>(dpool, pooled, OH * OW, C, CQ, sum_g, sum_gx);
  EVE_LAUNCH_CHECK();
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
  stem_pool_in_bwd_apply_kernel<<<cdiv(total, 256), 256, 0, s>>>(
      dpool, idx, x, total, H, W, OH, OW, C, mean, rstd, sum_g, sum_gx, dx,
      reinterpret_cast<__nv_bfloat16*>(d_hi), reinterpret_cast<__nv_bfloat16*>(d_lo));
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

}  // namespace eve
