// Gaze <-> screen kernels: Gaussian heatmap raster, soft-argmax, pitch/yaw -> PoG projection.
//
// Reference: src/models/common.py:226-243 (make_heatmap / batch_make_heatmaps),
// :294-323 (soft_argmax), :149-179 (to_screen_coordinates) with :32-40 (pitchyaw_to_vector),
// :89-102 (apply_transformation / apply_rotation) and :109-126 (get_intersect_with_zero).
// One CTA per heatmap (9216 pixels, coalesced rows, warp-shuffle + smem reductions);
// one thread per sample for the 3x3 / 4x4 geometry.
#include "common.cuh"

namespace eve {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reductions for 256 threads; every thread receives the result
__device__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sm[i];
  return r;
}
__device__ float block_max(float v, float* sm) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sm[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, sm[i]);
  return r;
}

__global__ void __launch_bounds__(256)
heatmap_fwd_kernel(const float* __restrict__ centres, int W, int H, float kx, float ky,
                   float alpha, float* __restrict__ out) {
  const int n = blockIdx.x;
  const float cx = kx * centres[2 * n], cy = ky * centres[2 * n + 1];
  float* o = out + (size_t)n * W * H;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
    int y = i / W, x = i - y * W;
    float dx = (float)x - cx, dy = (float)y - cy;
    o[i] = 1e-8f + expf(alpha * (dx * dx + dy * dy));
  }
}

__global__ void __launch_bounds__(256)
heatmap_bwd_kernel(const float* __restrict__ centres, const float* __restrict__ dout, int W, int H,
                   float kx, float ky, float alpha, float* __restrict__ dcentres) {
  __shared__ float sm[8];
  const int n = blockIdx.x;
  const float cx = kx * centres[2 * n], cy = ky * centres[2 * n + 1];
  const float* d = dout + (size_t)n * W * H;
  float ax = 0.f, ay = 0.f;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
    int y = i / W, x = i - y * W;
    float dx = (float)x - cx, dy = (float)y - cy;
    float e = expf(alpha * (dx * dx + dy * dy)) * d[i];
    ax = fmaf(e, dx, ax);
    ay = fmaf(e, dy, ay);
  }
  ax = block_sum(ax, sm);
  ay = block_sum(ay, sm);
  if (threadIdx.x == 0) {
    dcentres[2 * n] = -2.f * alpha * kx * ax;
    dcentres[2 * n + 1] = -2.f * alpha * ky * ay;
  }
}

__device__ __forceinline__ float lin01(int i, int n) {
  return n > 1 ? (float)((double)i / (double)(n - 1)) : 0.f;
}

// softmax(100 h) statistics of one heatmap: max, sum, E[x], E[y]
__device__ void softargmax_stats(const float* __restrict__ h, int W, int H, float* sm, float& mx,
                                 float& sum, float& lx, float& ly) {
  float m = -INFINITY;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) m = fmaxf(m, 100.f * h[i]);
  mx = block_max(m, sm);
  float s = 0.f, sx = 0.f, sy = 0.f;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
    int y = i / W, x = i - y * W;
    float e = expf(100.f * h[i] - mx);
    s += e;
    sx = fmaf(e, lin01(x, W), sx);
    sy = fmaf(e, lin01(y, H), sy);
  }
  sum = block_sum(s, sm);
  lx = block_sum(sx, sm) / sum;
  ly = block_sum(sy, sm) / sum;
}

__global__ void __launch_bounds__(256)
soft_argmax_fwd_kernel(const float* __restrict__ hm, int W, int H, float sw, float sh,
                       float* __restrict__ pog) {
  __shared__ float sm[8];
  const int n = blockIdx.x;
  float mx, sum, lx, ly;
  softargmax_stats(hm + (size_t)n * W * H, W, H, sm, mx, sum, lx, ly);
  if (threadIdx.x == 0) {
    pog[2 * n] = fminf(fmaxf(sw * lx, 0.f), sw);
    pog[2 * n + 1] = fminf(fmaxf(sh * ly, 0.f), sh);
  }
}

__global__ void __launch_bounds__(256)
soft_argmax_bwd_kernel(const float* __restrict__ hm, const float* __restrict__ dpog, int W, int H,
                       float sw, float sh, float* __restrict__ dhm) {
  __shared__ float sm[8];
  const int n = blockIdx.x;
  const float* h = hm + (size_t)n * W * H;
  float mx, sum, lx, ly;
  softargmax_stats(h, W, H, sm, mx, sum, lx, ly);
  float px = sw * lx, py = sh * ly;
  float gx = (px >= 0.f && px <= sw) ? dpog[2 * n] * sw : 0.f;
  float gy = (py >= 0.f && py <= sh) ? dpog[2 * n + 1] * sh : 0.f;
  const float inv = 100.f / sum;
  float* d = dhm + (size_t)n * W * H;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
    int y = i / W, x = i - y * W;
    float p = expf(100.f * h[i] - mx) * inv;
    d[i] = p * (gx * (lin01(x, W) - lx) + gy * (lin01(y, H) - ly));
  }
}

struct PogMid {
  float sp, cp, sy, cy;
  float d2[3], o2[3], t, den;
};

__device__ __forceinline__ void pog_core(const float* o, const float* g, const float* R,
                                         const float* M, PogMid& m) {
  sincosf(g[0], &m.sp, &m.cp);
  sincosf(g[1], &m.sy, &m.cy);
  float d0[3] = {-(m.cp * m.sy), -m.sp, -(m.cp * m.cy)};
  float d1[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) d1[i] = R[0 * 3 + i] * d0[0] + R[1 * 3 + i] * d0[1] + R[2 * 3 + i] * d0[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    m.d2[i] = M[i * 4 + 0] * d1[0] + M[i * 4 + 1] * d1[1] + M[i * 4 + 2] * d1[2];
    m.o2[i] = M[i * 4 + 0] * o[0] + M[i * 4 + 1] * o[1] + M[i * 4 + 2] * o[2] + M[i * 4 + 3];
  }
  m.den = m.d2[2] + 1e-7f;
  m.t = (0.f - m.o2[2]) / m.den;
}

__global__ void pog_fwd_kernel(int n, const float* __restrict__ origin, const float* __restrict__ g,
                               const float* __restrict__ rot, const float* __restrict__ inv_cam,
                               const float* __restrict__ ppm, float sw, float sh,
                               float* __restrict__ mm, float* __restrict__ px) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PogMid m;
  pog_core(origin + 3 * i, g + 2 * i, rot + 9 * i, inv_cam + 16 * i, m);
  float x = m.o2[0] + m.t * m.d2[0], y = m.o2[1] + m.t * m.d2[1];
  mm[2 * i] = x;
  mm[2 * i + 1] = y;
  px[2 * i] = fminf(fmaxf(x * ppm[2 * i], 0.f), sw);
  px[2 * i + 1] = fminf(fmaxf(y * ppm[2 * i + 1], 0.f), sh);
}

__global__ void pog_bwd_kernel(int n, const float* __restrict__ origin, const float* __restrict__ g,
                               const float* __restrict__ rot, const float* __restrict__ inv_cam,
                               const float* __restrict__ ppm, float sw, float sh,
                               const float* __restrict__ dmm, const float* __restrict__ dpx,
                               float* __restrict__ dg) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* R = rot + 9 * i;
  const float* M = inv_cam + 16 * i;
  PogMid m;
  pog_core(origin + 3 * i, g + 2 * i, R, M, m);
  float x = m.o2[0] + m.t * m.d2[0], y = m.o2[1] + m.t * m.d2[1];
  float gx = dmm ? dmm[2 * i] : 0.f, gy = dmm ? dmm[2 * i + 1] : 0.f;
  if (dpx) {
    float pxx = x * ppm[2 * i], pxy = y * ppm[2 * i + 1];
    if (pxx >= 0.f && pxx <= sw) gx = fmaf(dpx[2 * i], ppm[2 * i], gx);
    if (pxy >= 0.f && pxy <= sh) gy = fmaf(dpx[2 * i + 1], ppm[2 * i + 1], gy);
  }
  float gd2[3] = {gx * m.t, gy * m.t, (gx * m.d2[0] + gy * m.d2[1]) * (-m.t / m.den)};
  float gd1[3], gd0[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) gd1[k] = M[0 * 4 + k] * gd2[0] + M[1 * 4 + k] * gd2[1] + M[2 * 4 + k] * gd2[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) gd0[k] = R[k * 3 + 0] * gd1[0] + R[k * 3 + 1] * gd1[1] + R[k * 3 + 2] * gd1[2];
  float gv[3] = {-gd0[0], -gd0[1], -gd0[2]};
  dg[2 * i] = gv[0] * (-m.sp * m.sy) + gv[1] * m.cp + gv[2] * (-m.sp * m.cy);
  dg[2 * i + 1] = gv[0] * (m.cp * m.cy) + gv[2] * (-m.cp * m.sy);
}

int check_hm(const eve_heatmap_params* p) {
  EVE_REQUIRE(p, EVE_ERR_NULL, "heatmap: params is NULL");
  EVE_REQUIRE(p->n >= 0 && p->hm_w > 0 && p->hm_h > 0 && p->screen_w > 0.f && p->screen_h > 0.f,
              EVE_ERR_SHAPE, "heatmap: bad shape n=%d w=%d h=%d", p->n, p->hm_w, p->hm_h);
  return EVE_OK;
}

}  // namespace
}  // namespace eve

using namespace eve;

extern "C" int eve_heatmap_fwd(const eve_heatmap_params* p, const float* centres_px, float* out,
                               eve_stream_t stream) {
  EVE_TRY(check_hm(p));
  EVE_REQUIRE(p->sigma > 0.f, EVE_ERR_SHAPE, "heatmap: sigma must be positive");
  if (p->n == 0) return EVE_OK;
  EVE_REQUIRE(centres_px && out, EVE_ERR_NULL, "heatmap_fwd: NULL pointer");
  heatmap_fwd_kernel<<<p->n, 256, 0, as_stream(stream)>>>(
      centres_px, p->hm_w, p->hm_h, (float)p->hm_w / p->screen_w, (float)p->hm_h / p->screen_h,
      -0.5f / (p->sigma * p->sigma), out);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_heatmap_bwd(const eve_heatmap_params* p, const float* centres_px,
                               const float* dout, float* dcentres, eve_stream_t stream) {
  EVE_TRY(check_hm(p));
  EVE_REQUIRE(p->sigma > 0.f, EVE_ERR_SHAPE, "heatmap: sigma must be positive");
  if (p->n == 0) return EVE_OK;
  EVE_REQUIRE(centres_px && dout && dcentres, EVE_ERR_NULL, "heatmap_bwd: NULL pointer");
  heatmap_bwd_kernel<<<p->n, 256, 0, as_stream(stream)>>>(
      centres_px, dout, p->hm_w, p->hm_h, (float)p->hm_w / p->screen_w,
      (float)p->hm_h / p->screen_h, -0.5f / (p->sigma * p->sigma), dcentres);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_soft_argmax_fwd(const eve_heatmap_params* p, const float* heatmaps,
                                   float* pog_px, eve_stream_t stream) {
  EVE_TRY(check_hm(p));
  if (p->n == 0) return EVE_OK;
  EVE_REQUIRE(heatmaps && pog_px, EVE_ERR_NULL, "soft_argmax_fwd: NULL pointer");
  soft_argmax_fwd_kernel<<<p->n, 256, 0, as_stream(stream)>>>(heatmaps, p->hm_w, p->hm_h,
                                                             p->screen_w, p->screen_h, pog_px);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_soft_argmax_bwd(const eve_heatmap_params* p, const float* heatmaps,
                                   const float* dpog, float* dheatmaps, eve_stream_t stream) {
  EVE_TRY(check_hm(p));
  if (p->n == 0) return EVE_OK;
  EVE_REQUIRE(heatmaps && dpog && dheatmaps, EVE_ERR_NULL, "soft_argmax_bwd: NULL pointer");
  soft_argmax_bwd_kernel<<<p->n, 256, 0, as_stream(stream)>>>(
      heatmaps, dpog, p->hm_w, p->hm_h, p->screen_w, p->screen_h, dheatmaps);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_pog_fwd(int n, const float* origin, const float* g, const float* rot,
                           const float* inv_cam, const float* ppm, float screen_w, float screen_h,
                           float* pog_mm, float* pog_px, eve_stream_t stream) {
  EVE_REQUIRE(n >= 0, EVE_ERR_SHAPE, "pog_fwd: n < 0");
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(origin && g && rot && inv_cam && ppm && pog_mm && pog_px, EVE_ERR_NULL,
              "pog_fwd: NULL pointer");
  pog_fwd_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(n, origin, g, rot, inv_cam, ppm,
                                                             screen_w, screen_h, pog_mm, pog_px);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_pog_bwd(int n, const float* origin, const float* g, const float* rot,
                           const float* inv_cam, const float* ppm, float screen_w, float screen_h,
                           const float* dpog_mm, const float* dpog_px, float* dg,
                           eve_stream_t stream) {
  EVE_REQUIRE(n >= 0, EVE_ERR_SHAPE, "pog_bwd: n < 0");
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(origin && g && rot && inv_cam && ppm && dg, EVE_ERR_NULL, "pog_bwd: NULL pointer");
  pog_bwd_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(
      n, origin, g, rot, inv_cam, ppm, screen_w, screen_h, dpog_mm, dpog_px, dg);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
