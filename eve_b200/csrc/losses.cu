// Validity-masked sequence losses and metrics of EVE.calculate_losses_and_metrics
// (eve.py:286-439) with the loss objects of src/losses/*.py:
//   BaseLossWithValidity (base_loss_with_validity.py:32-73): per clip  sum_t(v * l) / n_valid
//   (divided only when n_valid > 1), then the mean over clips;  l = angular error in degrees
//   (angular.py), MSE / L1 (mean over the frame's features), Euclidean distance, or the per-frame
//   binary cross-entropy / MSE of 72x128 heatmaps (cross_entropy.py:29-35).
// The reference evaluates ~30 such terms with Python loops over the batch (and over time for the
// BCE); here ONE launch evaluates a whole table of terms (one CTA per term, one warp per clip,
// fixed-order reductions: deterministic) and one launch produces every gradient.
#include "common.cuh"
#include "gaze_math.cuh"

namespace eve {
namespace {

struct TermTable {
  int n;
  eve_loss_term t[EVE_LOSS_MAX_TERMS];
};

__device__ __forceinline__ float frame_loss(const eve_loss_term& tm, size_t f) {
  const float* a = tm.pred + f * tm.dim;
  if (tm.op == EVE_LOSS_IDENTITY) return a[0];
  const float* b = tm.gt + f * tm.dim;
  if (tm.op == EVE_LOSS_ANGULAR) return gm::angular_error_deg(a, b, -1.0f + 1e-8f, 1.0f - 1e-8f);
  float s = 0.f;
  for (int k = 0; k < tm.dim; ++k) {
    const float d = a[k] - b[k];
    s += tm.op == EVE_LOSS_L1 ? fabsf(d) : d * d;
  }
  if (tm.op == EVE_LOSS_EUCLIDEAN) return sqrtf(s);
  return s / (float)tm.dim;
}

__device__ __forceinline__ bool frame_valid(const eve_loss_term& tm, size_t f) {
  return tm.valid[f] != 0 && (tm.valid2 == nullptr || tm.valid2[f] != 0);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid = nterms, block = 256 (8 warps); warp w handles clips w, w + 8, ...
__global__ void __launch_bounds__(256)
masked_losses_fwd_kernel(const TermTable tab, int B, int T, float* __restrict__ out) {
  __shared__ float clip[8];
  const eve_loss_term& tm = tab.t[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;                       // this warp's sum over its clips (lane 0 holds it)
  for (int b = warp; b < B; b += 8) {
    float s = 0.f, n = 0.f;
    for (int t = lane; t < T; t += 32) {
      const size_t f = (size_t)b * T + t;
      if (frame_valid(tm, f)) {
        s += frame_loss(tm, f);
        n += 1.f;
      }
    }
    s = warp_sum_f(s);
    n = warp_sum_f(n);
    acc += n > 1.f ? s / n : s;
  }
  if (lane == 0) clip[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += clip[w];
    out[blockIdx.x] = tot / (float)B;
  }
}

// grid = ceil(B*T / 128): thread = frame; walks the table in order, so terms that share a
// prediction tensor accumulate into its gradient without races and in a fixed order
__global__ void __launch_bounds__(128)
masked_losses_bwd_kernel(const TermTable tab, int B, int T, const float* __restrict__ dout) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B * T) return;
  const int b = f / T;
  for (int i = 0; i < tab.n; ++i) {
    const eve_loss_term& tm = tab.t[i];
    if (tm.dpred == nullptr) continue;
    const float go = dout[i];
    if (go == 0.f || !frame_valid(tm, f)) continue;
    int n = 0;
    for (int t = 0; t < T; ++t) n += frame_valid(tm, (size_t)b * T + t) ? 1 : 0;
    const float w = go / (float)B / (n > 1 ? (float)n : 1.f);
    const float* a = tm.pred + (size_t)f * tm.dim;
    float* d = tm.dpred + (size_t)f * tm.dim;
    if (tm.op == EVE_LOSS_IDENTITY) {
      d[0] += w;
      continue;
    }
    const float* g = tm.gt + (size_t)f * tm.dim;
    if (tm.op == EVE_LOSS_ANGULAR) {
      float da[2] = {0.f, 0.f};
      gm::angular_error_deg_vjp(a, g, -1.0f + 1e-8f, 1.0f - 1e-8f, w, da);
      d[0] += da[0];
      d[1] += da[1];
      continue;
    }
    float s = 0.f;
    if (tm.op == EVE_LOSS_EUCLIDEAN) {
      for (int k = 0; k < tm.dim; ++k) s += (a[k] - g[k]) * (a[k] - g[k]);
      s = sqrtf(s);
    }
    for (int k = 0; k < tm.dim; ++k) {
      const float e = a[k] - g[k];
      float v;
      if (tm.op == EVE_LOSS_MSE) v = 2.f * e / (float)tm.dim;
      else if (tm.op == EVE_LOSS_L1) v = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) / (float)tm.dim;
      else v = e / s;
      d[k] += w * v;
    }
  }
}

__device__ float block_sum256(float v, float* sm) {
  v = warp_sum_f(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < 8; ++i) r += sm[i];
  return r;
}

// per-frame mean binary cross-entropy (torch semantics: logs clamped at -100) and mean squared
// error of one heatmap pair; one CTA per frame, float4 loads
__global__ void __launch_bounds__(256)
heatmap_frame_losses_fwd_kernel(int HW, const float* __restrict__ pred, const float* __restrict__ gt,
                                float* __restrict__ bce, float* __restrict__ mse) {
  __shared__ float sm[8];
  const size_t base = (size_t)blockIdx.x * HW;
  float sb = 0.f, sq = 0.f;
  for (int i = threadIdx.x * 4; i < HW; i += blockDim.x * 4) {
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(pred + base + i));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(gt + base + i));
    const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float la = fmaxf(logf(a[k]), -100.f), l1a = fmaxf(logf(1.f - a[k]), -100.f);
      sb -= b[k] * la + (1.f - b[k]) * l1a;
      const float e = a[k] - b[k];
      sq = fmaf(e, e, sq);
    }
  }
  sb = block_sum256(sb, sm);
  sq = block_sum256(sq, sm);
  if (threadIdx.x == 0) {
    if (bce) bce[blockIdx.x] = sb / (float)HW;
    if (mse) mse[blockIdx.x] = sq / (float)HW;
  }
}

// dpred = dbce[f]/HW * (a - b) / max((1 - a) a, 1e-12)  +  dmse[f]/HW * 2 (a - b)
__global__ void __launch_bounds__(256)
heatmap_frame_losses_bwd_kernel(int HW, const float* __restrict__ pred, const float* __restrict__ gt,
                                const float* __restrict__ dbce, const float* __restrict__ dmse,
                                float* __restrict__ dpred) {
  const size_t base = (size_t)blockIdx.x * HW;
  const float wb = dbce ? dbce[blockIdx.x] / (float)HW : 0.f;
  const float wm = dmse ? 2.f * dmse[blockIdx.x] / (float)HW : 0.f;
  for (int i = threadIdx.x * 4; i < HW; i += blockDim.x * 4) {
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(pred + base + i));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(gt + base + i));
    const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float e = a[k] - b[k];
      o[k] = wm * e;
      if (wb != 0.f) o[k] += wb * e / fmaxf((1.f - a[k]) * a[k], 1e-12f);
    }
    *reinterpret_cast<float4*>(dpred + base + i) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

int check_terms(int nterms, const eve_loss_term* terms, int batch, int steps, TermTable& tab) {
  EVE_REQUIRE(terms || nterms == 0, EVE_ERR_NULL, "masked_losses: terms is NULL");
  EVE_REQUIRE(nterms >= 0 && nterms <= EVE_LOSS_MAX_TERMS, EVE_ERR_SHAPE,
              "masked_losses: %d terms (at most %d per call)", nterms, EVE_LOSS_MAX_TERMS);
  EVE_REQUIRE(batch > 0 && steps > 0, EVE_ERR_SHAPE, "masked_losses: batch=%d steps=%d", batch, steps);
  tab.n = nterms;
  for (int i = 0; i < nterms; ++i) {
    const eve_loss_term& t = terms[i];
    EVE_REQUIRE(t.op >= EVE_LOSS_ANGULAR && t.op <= EVE_LOSS_IDENTITY, EVE_ERR_CONFIG,
                "masked_losses: term %d has unknown op %d", i, t.op);
    EVE_REQUIRE(t.dim >= 1 && t.dim <= 16 && (t.op != EVE_LOSS_ANGULAR || t.dim == 2) &&
                    (t.op != EVE_LOSS_IDENTITY || t.dim == 1),
                EVE_ERR_SHAPE, "masked_losses: term %d: op %d with %d features", i, t.op, t.dim);
    EVE_REQUIRE(t.pred && t.valid && (t.gt || t.op == EVE_LOSS_IDENTITY), EVE_ERR_NULL,
                "masked_losses: term %d has a NULL pointer", i);
    tab.t[i] = t;
  }
  return EVE_OK;
}

}  // namespace
}  // namespace eve

using namespace eve;

extern "C" int eve_masked_losses_fwd(int nterms, const eve_loss_term* terms, int batch, int steps,
                                     float* out, eve_stream_t stream) {
  TermTable tab;
  EVE_TRY(check_terms(nterms, terms, batch, steps, tab));
  if (nterms == 0) return EVE_OK;
  EVE_REQUIRE(out, EVE_ERR_NULL, "masked_losses_fwd: out is NULL");
  masked_losses_fwd_kernel<<<nterms, 256, 0, as_stream(stream)>>>(tab, batch, steps, out);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_masked_losses_bwd(int nterms, const eve_loss_term* terms, int batch, int steps,
                                     const float* dout, eve_stream_t stream) {
  TermTable tab;
  EVE_TRY(check_terms(nterms, terms, batch, steps, tab));
  if (nterms == 0) return EVE_OK;
  EVE_REQUIRE(dout, EVE_ERR_NULL, "masked_losses_bwd: dout is NULL");
  masked_losses_bwd_kernel<<<cdiv((long long)batch * steps, 128), 128, 0, as_stream(stream)>>>(
      tab, batch, steps, dout);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_heatmap_frame_losses_fwd(int n, int hw, const float* pred, const float* gt,
                                            float* bce, float* mse, eve_stream_t stream) {
  EVE_REQUIRE(n >= 0 && hw > 0 && hw % 4 == 0, EVE_ERR_SHAPE,
              "heatmap_frame_losses: n=%d hw=%d (hw must be a multiple of 4)", n, hw);
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(pred && gt && (bce || mse), EVE_ERR_NULL, "heatmap_frame_losses_fwd: NULL pointer");
  heatmap_frame_losses_fwd_kernel<<<n, 256, 0, as_stream(stream)>>>(hw, pred, gt, bce, mse);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

extern "C" int eve_heatmap_frame_losses_bwd(int n, int hw, const float* pred, const float* gt,
                                            const float* dbce, const float* dmse, float* dpred,
                                            eve_stream_t stream) {
  EVE_REQUIRE(n >= 0 && hw > 0 && hw % 4 == 0, EVE_ERR_SHAPE,
              "heatmap_frame_losses: n=%d hw=%d (hw must be a multiple of 4)", n, hw);
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(pred && gt && dpred, EVE_ERR_NULL, "heatmap_frame_losses_bwd: NULL pointer");
  heatmap_frame_losses_bwd_kernel<<<n, 256, 0, as_stream(stream)>>>(hw, pred, gt, dbce, dmse, dpred);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
