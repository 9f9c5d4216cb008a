// Per-frame gaze geometry kernels of EVE.forward / calculate_additional_labels:
//   calculate_combined_gaze_direction (common.py:129-146), apply_offset_augmentation (:182-218),
//   the label block of eve.py:441-543 (PoG in cm, averaged origin / PoG, validity AND, the
//   ground-truth combined gaze and the three validity-scaled label heatmaps) and the gaze-history
//   maps of common.py:249-287 as an O(T) recurrence.
// One thread per frame for the 3x3 algebra (math + hand-derived VJPs: gaze_math.cuh, checked on the
// CPU against torch autograd by tests/test_host_math.py); one CTA per heatmap for the rasters.
#include "common.cuh"
#include "gaze_math.cuh"

namespace eve {
namespace {

__global__ void combined_gaze_fwd_kernel(int n, const float* __restrict__ origin,
                                         const float* __restrict__ pog, const float* __restrict__ R,
                                         const float* __restrict__ cam, float* __restrict__ g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float out[2];
  gm::combined_gaze(origin + 3 * i, pog + 2 * i, R + 9 * i, cam + 16 * i, out);
  g[2 * i] = out[0];
  g[2 * i + 1] = out[1];
}

__global__ void combined_gaze_bwd_kernel(int n, const float* __restrict__ origin,
                                         const float* __restrict__ pog, const float* __restrict__ R,
                                         const float* __restrict__ cam, const float* __restrict__ dg,
                                         float* __restrict__ dpog) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d[2] = {0.f, 0.f};
  gm::combined_gaze_vjp(origin + 3 * i, pog + 2 * i, R + 9 * i, cam + 16 * i, dg + 2 * i, d);
  dpog[2 * i] = d[0];
  dpog[2 * i + 1] = d[1];
}

__global__ void offset_aug_fwd_kernel(int n, int frames_per_kappa, const float* __restrict__ g,
                                      const float* __restrict__ R, const float* __restrict__ kappa,
                                      int inverse, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  gm::OffsetAugMid<float> m;
  float o[2];
  gm::offset_augmentation(g + 2 * i, R + 9 * i, kappa + 2 * (i / frames_per_kappa), inverse != 0, o, m);
  out[2 * i] = o[0];
  out[2 * i + 1] = o[1];
}

__global__ void offset_aug_bwd_kernel(int n, int frames_per_kappa, const float* __restrict__ g,
                                      const float* __restrict__ R, const float* __restrict__ kappa,
                                      int inverse, const float* __restrict__ dout,
                                      float* __restrict__ dg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d[2] = {0.f, 0.f};
  gm::offset_augmentation_vjp(g + 2 * i, R + 9 * i, kappa + 2 * (i / frames_per_kappa), inverse != 0,
                              dout + 2 * i, d);
  dg[2 * i] = d[0];
  dg[2 * i + 1] = d[1];
}

// eve.py:449-456, 498-543: everything the label block derives per frame from the Tobii PoG
__global__ void labels_kernel(const eve_label_args a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const float kx = 0.1f * a.mm_per_px[2 * i], ky = 0.1f * a.mm_per_px[2 * i + 1];
  const float lx = a.left_pog_px[2 * i], ly = a.left_pog_px[2 * i + 1];
  const float rx = a.right_pog_px[2 * i], ry = a.right_pog_px[2 * i + 1];
  const float lcx = lx * kx, lcy = ly * ky, rcx = rx * kx, rcy = ry * ky;
  a.left_pog_cm[2 * i] = lcx;
  a.left_pog_cm[2 * i + 1] = lcy;
  a.right_pog_cm[2 * i] = rcx;
  a.right_pog_cm[2 * i + 1] = rcy;
  // torch.stack([l, r], -1).mean(-1) = (l + r) / 2
  float o[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[k] = (a.left_o[3 * i + k] + a.right_o[3 * i + k]) / 2.f;
    a.o[3 * i + k] = o[k];
  }
  a.pog_px[2 * i] = (lx + rx) / 2.f;
  a.pog_px[2 * i + 1] = (ly + ry) / 2.f;
  const float cm[2] = {(lcx + rcx) / 2.f, (lcy + rcy) / 2.f};
  a.pog_cm[2 * i] = cm[0];
  a.pog_cm[2 * i + 1] = cm[1];
  a.valid[i] = (a.left_valid[i] != 0 && a.right_valid[i] != 0) ? 1 : 0;
  const float mm[2] = {10.0f * cm[0], 10.0f * cm[1]};
  float g[2];
  gm::combined_gaze(o, mm, a.left_R + 9 * i, a.cam + 16 * i, g);
  a.g[2 * i] = g[0];
  a.g[2 * i + 1] = g[1];
}

// label heatmaps (eve.py:519-531): batch_make_heatmaps(PoG_px_tobii, sigma) * validity for up to
// three sigmas from one launch; grid (n, nsig)
__global__ void __launch_bounds__(256)
heatmap_labels_kernel(const float* __restrict__ centres, const unsigned char* __restrict__ valid,
                      int W, int H, float kx, float ky, float a0, float a1, float a2,
                      float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2) {
  const int n = blockIdx.x, s = blockIdx.y;
  const float alpha = s == 0 ? a0 : (s == 1 ? a1 : a2);
  float* o = (s == 0 ? o0 : (s == 1 ? o1 : o2)) + (size_t)n * W * H;
  const float cx = kx * centres[2 * n], cy = ky * centres[2 * n + 1];
  const float v = valid[n] ? 1.f : 0.f;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float dx = (float)x - cx, dy = (float)y - cy;
    o[i] = (1e-8f + expf(alpha * (dx * dx + dy * dy))) * v;
  }
}

// Gaze-history maps for EVERY prefix (common.py:249-287, eve.py:596-601) as a recurrence over t:
//   H_t = decay^((L_t - L_{t-1}) * 1e-6) * H_{t-1} + [ts_t != 0] * valid_t * heatmap_t
// with L_t the last non-zero timestamp up to t (padded frames carry ts == 0 and leave the map
// untouched).  The reference recomputes the weighted sum for each t: O(T^2).  One thread per
// (clip, pixel); `fac` holds the T per-step decay factors of the clip.
__global__ void __launch_bounds__(256)
history_fwd_kernel(int T, int HW, const float* __restrict__ fac, const float* __restrict__ add,
                   const float* __restrict__ hm, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float h = 0.f;
  for (int t = 0; t < T; ++t) {
    const size_t at = ((size_t)b * T + t) * HW + p;
    h = fmaf(fac[b * T + t], h, add[b * T + t] * __ldg(hm + at));
    out[at] = h;
  }
}
// d heatmap_t = add_t * G_t,  G_t = dOut_t + fac_{t+1} * G_{t+1}
__global__ void __launch_bounds__(256)
history_bwd_kernel(int T, int HW, const float* __restrict__ fac, const float* __restrict__ add,
                   const float* __restrict__ dout, float* __restrict__ dhm) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float g = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    const size_t at = ((size_t)b * T + t) * HW + p;
    g = __ldg(dout + at) + (t + 1 < T ? fac[b * T + t + 1] * g : 0.f);
    dhm[at] = add[b * T + t] * g;
  }
}
// per clip: the decay factor and the additive weight of every step (one thread per clip)
__global__ void history_factors_kernel(int B, int T, const long long* __restrict__ ts,
                                       const unsigned char* __restrict__ valid, float decay,
                                       float* __restrict__ fac, float* __restrict__ add) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long last = 0;
  bool have = false;
  for (int t = 0; t < T; ++t) {
    const long long cur = ts[b * T + t];
    const bool nz = cur != 0;
    float f = 1.f;
    if (nz) {
      // the reference weighs entry t' by decay ** ((ts_last - ts_t') * 1e-6), fp32 arithmetic
      if (have) f = powf(decay, (float)(cur - last) * 1e-6f);
      last = cur;
      have = true;
    }
    fac[b * T + t] = f;
    add[b * T + t] = (nz && valid[b * T + t]) ? 1.f : 0.f;
  }
}

}  // namespace
}  // namespace eve

using namespace eve;

extern "C" int eve_combined_gaze_fwd(int n, const float* origin, const float* pog_mm,
                                     const float* head_rot, const float* cam, float* g,
                                     eve_stream_t stream) {
  EVE_REQUIRE(n >= 0, EVE_ERR_SHAPE, "combined_gaze_fwd: n < 0");
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(origin && pog_mm && head_rot && cam && g, EVE_ERR_NULL, "combined_gaze_fwd: NULL pointer");
  combined_gaze_fwd_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(n, origin, pog_mm, head_rot, cam, g);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_combined_gaze_bwd(int n, const float* origin, const float* pog_mm,
                                     const float* head_rot, const float* cam, const float* dg,
                                     float* dpog_mm, eve_stream_t stream) {
  EVE_REQUIRE(n >= 0, EVE_ERR_SHAPE, "combined_gaze_bwd: n < 0");
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(origin && pog_mm && head_rot && cam && dg && dpog_mm, EVE_ERR_NULL,
              "combined_gaze_bwd: NULL pointer");
  combined_gaze_bwd_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(n, origin, pog_mm, head_rot, cam,
                                                                        dg, dpog_mm);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_offset_augmentation_fwd(int n, int frames_per_kappa, const float* g,
                                           const float* head_rot, const float* kappa, int inverse,
                                           float* out, eve_stream_t stream) {
  EVE_REQUIRE(n >= 0 && frames_per_kappa >= 1, EVE_ERR_SHAPE, "offset_augmentation_fwd: bad sizes");
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(g && head_rot && kappa && out, EVE_ERR_NULL, "offset_augmentation_fwd: NULL pointer");
  offset_aug_fwd_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(n, frames_per_kappa, g, head_rot,
                                                                     kappa, inverse, out);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_offset_augmentation_bwd(int n, int frames_per_kappa, const float* g,
                                           const float* head_rot, const float* kappa, int inverse,
                                           const float* dout, float* dg, eve_stream_t stream) {
  EVE_REQUIRE(n >= 0 && frames_per_kappa >= 1, EVE_ERR_SHAPE, "offset_augmentation_bwd: bad sizes");
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(g && head_rot && kappa && dout && dg, EVE_ERR_NULL, "offset_augmentation_bwd: NULL pointer");
  offset_aug_bwd_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(n, frames_per_kappa, g, head_rot,
                                                                     kappa, inverse, dout, dg);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_labels_fwd(const eve_label_args* a, eve_stream_t stream) {
  EVE_REQUIRE(a, EVE_ERR_NULL, "labels_fwd: args is NULL");
  EVE_REQUIRE(a->n >= 0, EVE_ERR_SHAPE, "labels_fwd: n < 0");
  if (a->n == 0) return EVE_OK;
  EVE_REQUIRE(a->left_pog_px && a->right_pog_px && a->left_valid && a->right_valid && a->mm_per_px &&
                  a->left_o && a->right_o && a->left_R && a->cam && a->left_pog_cm && a->right_pog_cm &&
                  a->o && a->pog_px && a->pog_cm && a->valid && a->g,
              EVE_ERR_NULL, "labels_fwd: NULL pointer");
  labels_kernel<<<cdiv(a->n, 128), 128, 0, as_stream(stream)>>>(*a);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_heatmap_labels_fwd(const eve_heatmap_params* p, const float* centres_px,
                                      const unsigned char* valid, int nsig, const float* sigmas,
                                      float* const* outs, eve_stream_t stream) {
  EVE_REQUIRE(p && sigmas && outs, EVE_ERR_NULL, "heatmap_labels: NULL pointer");
  EVE_REQUIRE(p->n >= 0 && p->hm_w > 0 && p->hm_h > 0 && p->screen_w > 0.f && p->screen_h > 0.f &&
                  nsig >= 1 && nsig <= 3,
              EVE_ERR_SHAPE, "heatmap_labels: bad shape n=%d nsig=%d", p->n, nsig);
  if (p->n == 0) return EVE_OK;
  EVE_REQUIRE(centres_px && valid, EVE_ERR_NULL, "heatmap_labels: NULL pointer");
  float al[3] = {0.f, 0.f, 0.f};
  float* o[3] = {nullptr, nullptr, nullptr};
  for (int s = 0; s < nsig; ++s) {
    EVE_REQUIRE(sigmas[s] > 0.f && outs[s], EVE_ERR_SHAPE, "heatmap_labels: bad sigma / NULL output %d", s);
    al[s] = -0.5f / (sigmas[s] * sigmas[s]);
    o[s] = outs[s];
  }
  heatmap_labels_kernel<<<dim3(p->n, nsig), 256, 0, as_stream(stream)>>>(
      centres_px, valid, p->hm_w, p->hm_h, (float)p->hm_w / p->screen_w, (float)p->hm_h / p->screen_h,
      al[0], al[1], al[2], o[0], o[1], o[2]);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" size_t eve_gaze_history_scratch_bytes(int batch, int steps) {
  return batch > 0 && steps > 0 ? (size_t)2 * batch * steps * sizeof(float) : 0;
}

extern "C" int eve_gaze_history_fwd(int batch, int steps, int hw, const long long* timestamps,
                                    const unsigned char* valid, float decay, const float* heatmaps,
                                    float* out, void* scratch, size_t scratch_bytes,
                                    eve_stream_t stream) {
  EVE_REQUIRE(batch >= 0 && steps >= 0 && hw > 0, EVE_ERR_SHAPE, "gaze_history_fwd: bad sizes");
  if (batch == 0 || steps == 0) return EVE_OK;
  EVE_REQUIRE(timestamps && valid && heatmaps && out && scratch, EVE_ERR_NULL, "gaze_history_fwd: NULL pointer");
  EVE_REQUIRE(scratch_bytes >= eve_gaze_history_scratch_bytes(batch, steps), EVE_ERR_WORKSPACE,
              "gaze_history_fwd: scratch too small");
  float* fac = (float*)scratch;
  float* add = fac + (size_t)batch * steps;
  history_factors_kernel<<<cdiv(batch, 64), 64, 0, as_stream(stream)>>>(batch, steps, timestamps, valid,
                                                                        decay, fac, add);
  EVE_LAUNCH_CHECK();
  history_fwd_kernel<<<dim3(cdiv(hw, 256), batch), 256, 0, as_stream(stream)>>>(steps, hw, fac, add,
                                                                                heatmaps, out);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_gaze_history_bwd(int batch, int steps, int hw, const long long* timestamps,
                                    const unsigned char* valid, float decay, const float* dout,
                                    float* dheatmaps, void* scratch, size_t scratch_bytes,
                                    eve_stream_t stream) {
  EVE_REQUIRE(batch >= 0 && steps >= 0 && hw > 0, EVE_ERR_SHAPE, "gaze_history_bwd: bad sizes");
  if (batch == 0 || steps == 0) return EVE_OK;
  EVE_REQUIRE(timestamps && valid && dout && dheatmaps && scratch, EVE_ERR_NULL, "gaze_history_bwd: NULL pointer");
  EVE_REQUIRE(scratch_bytes >= eve_gaze_history_scratch_bytes(batch, steps), EVE_ERR_WORKSPACE,
              "gaze_history_bwd: scratch too small");
  float* fac = (float*)scratch;
  float* add = fac + (size_t)batch * steps;
  history_factors_kernel<<<cdiv(batch, 64), 64, 0, as_stream(stream)>>>(batch, steps, timestamps, valid,
                                                                        decay, fac, add);
  EVE_LAUNCH_CHECK();
  history_bwd_kernel<<<dim3(cdiv(hw, 256), batch), 256, 0, as_stream(stream)>>>(steps, hw, fac, add, dout,
                                                                                dheatmaps);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
