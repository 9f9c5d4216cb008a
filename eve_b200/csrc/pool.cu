// Layout changes, pooling and resampling -- HBM-bound NHWC fp32 kernels.
//
// Reference ops replaced: torchvision ResNet.maxpool / avgpool (via eye_net.py:106),
// nn.AdaptiveMaxPool2d (refine_net.py:93,121), nn.Upsample(bilinear, align_corners=False)
// + torch.cat (refine_net.py:101,124-126).  A warp always walks 32 consecutive channels of
// one pixel (128-byte coalesced rows); the arg-max index convention is torch's: first
// maximum in row-major window order.
#include "common.cuh"

namespace eve {

namespace {

// [N][R][Cc] -> [N][Cc][R] generic per-image transpose with a 32x32 smem tile.
// in : rows = R, cols = Cc (cols contiguous)   out: rows = Cc, cols = R
__global__ void transpose_kernel(const float* __restrict__ x, int R, int Cc,
                                 float* __restrict__ y) {
  __shared__ float tile[32][33];
  const size_t img = (size_t)blockIdx.z * R * Cc;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[j][threadIdx.x] = x[img + (size_t)r * Cc + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Cc) y[img + (size_t)c * R + r] = tile[threadIdx.x][j];
  }
}

// [N][3][HW] -> [N][HW][3] (image inputs): one thread = 4 pixels, three coalesced 16-byte loads (one
// per colour plane) and three consecutive 16-byte stores; the 32x32-tile kernel above leaves 29 of
// its 32 tile rows empty at C = 3 (0.26 ms for the 480 eye patches of a step instead of 0.04 ms)
__global__ void __launch_bounds__(256)
nchw3_to_nhwc_kernel(const float* __restrict__ x, long long quads, int HW4, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= quads) return;
  const long long n = i / HW4;
  const int q = (int)(i - n * HW4);
  const float4* src = reinterpret_cast<const float4*>(x) + n * 3 * HW4 + q;
  const float4 a = __ldg(src), b = __ldg(src + HW4), c = __ldg(src + 2 * HW4);
  float4* dst = reinterpret_cast<float4*>(y) + (n * HW4 + q) * 3;
  dst[0] = make_float4(a.x, b.x, c.x, a.y);
  dst[1] = make_float4(b.y, c.y, a.z, b.z);
  dst[2] = make_float4(c.z, a.w, b.w, c.w);
}

// ---- every kernel below: one thread = 4 consecutive channels (16-byte accesses) ----
struct F4 {
  float v[4];
  __device__ __forceinline__ F4() {}
  __device__ __forceinline__ F4(float4 f) { v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
  __device__ __forceinline__ float4 f4() const { return make_float4(v[0], v[1], v[2], v[3]); }
};
__device__ __forceinline__ F4 ld4(const float* p) { return F4(__ldg(reinterpret_cast<const float4*>(p))); }
__device__ __forceinline__ void st4(float* p, const F4& a) { *reinterpret_cast<float4*>(p) = a.f4(); }

__global__ void __launch_bounds__(256)
in_relu_maxpool_kernel(const float* __restrict__ x, long long total4, int H, int W, int C, int OH,
                       int OW, const float* __restrict__ mean, const float* __restrict__ rstd,
                       float* __restrict__ y, int32_t* __restrict__ idx) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  int ox = (int)(t % OW);
  t /= OW;
  int oy = (int)(t % OH);
  int n = (int)(t / OH);
  const F4 m = ld4(mean + (size_t)n * C + c), r = ld4(rstd + (size_t)n * C + c);
  const float* xp = x + (size_t)n * H * W * C + c;
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int bi[4] = {-1, -1, -1, -1};
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int h = oy * 2 - 1 + dy;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      int w = ox * 2 - 1 + dx;
      if (w < 0 || w >= W) continue;
      F4 xv = ld4(xp + (size_t)(h * W + w) * C);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = (xv.v[j] - m.v[j]) * r.v[j];
        v = v > 0.f ? v : 0.f;
        if (v > best[j] || bi[j] < 0) {
          best[j] = v;
          bi[j] = h * W + w;
        }
      }
    }
  }
  *reinterpret_cast<float4*>(y + i * 4) = make_float4(best[0], best[1], best[2], best[3]);
  *reinterpret_cast<int4*>(idx + i * 4) = make_int4(bi[0], bi[1], bi[2], bi[3]);
}

// gather form of the max-pool backward (3x3, stride 2, pad 1): deterministic, no atomics
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx, long long total4,
                   int H, int W, int OH, int OW, int C, float* __restrict__ g) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  int w = (int)(t % W);
  t /= W;
  int h = (int)(t % H);
  int n = (int)(t / H);
  const int me = h * W + w;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  // windows oy with oy*2-1 <= h <= oy*2+1
  int oy0 = h / 2, oy1 = (h + 1) / 2;
  int ox0 = w / 2, ox1 = (w + 1) / 2;
  for (int oy = oy0; oy <= oy1; ++oy) {
    if (oy >= OH) continue;
    for (int ox = ox0; ox <= ox1; ++ox) {
      if (ox >= OW) continue;
      size_t o = (((size_t)n * OH + oy) * OW + ox) * C + c;
      int4 id = __ldg(reinterpret_cast<const int4*>(idx + o));
      F4 d = ld4(dy + o);
      if (id.x == me) s[0] += d.v[0];
      if (id.y == me) s[1] += d.v[1];
      if (id.z == me) s[2] += d.v[2];
      if (id.w == me) s[3] += d.v[3];
    }
  }
  *reinterpret_cast<float4*>(g + i * 4) = make_float4(s[0], s[1], s[2], s[3]);
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ x, int HW, int C,
                                   float* __restrict__ y) {
  int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += x[((size_t)n * HW + p) * C + c];
    y[(size_t)n * C + c] = s / (float)HW;
  }
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, long long total, int HW, int C,
                                   float* __restrict__ dx) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  int n = (int)(i / ((long long)HW * C));
  dx[i] = dy[(size_t)n * C + c] / (float)HW;
}

__device__ __forceinline__ int win_start(int i, int L, int O) { return (i * L) / O; }
__device__ __forceinline__ int win_end(int i, int L, int O) { return ((i + 1) * L + O - 1) / O; }

__global__ void __launch_bounds__(256)
adaptive_maxpool_fwd_kernel(const float* __restrict__ x, long long total4, int H, int W, int C,
                            int OH, int OW, float* __restrict__ y, int32_t* __restrict__ idx,
                            float* __restrict__ copy, int cld) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  int ox = (int)(t % OW);
  t /= OW;
  int oy = (int)(t % OH);
  int n = (int)(t / OH);
  const float* xp = x + (size_t)n * H * W * C + c;
  int h0 = win_start(oy, H, OH), h1 = win_end(oy, H, OH);
  int w0 = win_start(ox, W, OW), w1 = win_end(ox, W, OW);
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int bi[4];
  bi[0] = bi[1] = bi[2] = bi[3] = h0 * W + w0;
  for (int h = h0; h < h1; ++h)
    for (int w = w0; w < w1; ++w) {
      F4 xv = ld4(xp + (size_t)(h * W + w) * C);
      // the input on its way into the concat buffer of the decoder (rows of cld floats); overlapping
      // windows write the same value twice
      if (copy) st4(copy + (((size_t)n * H + h) * W + w) * cld + c, xv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = xv.v[j];
        if (v > best[j] || v != v) {  // torch: (val > max) || isnan(val)
          best[j] = v;
          bi[j] = h * W + w;
        }
      }
    }
  *reinterpret_cast<float4*>(y + i * 4) = make_float4(best[0], best[1], best[2], best[3]);
  *reinterpret_cast<int4*>(idx + i * 4) = make_int4(bi[0], bi[1], bi[2], bi[3]);
}

// Windows that tile the input exactly (H = kh * OH, W = kw * OW: every level of GazeRefineNet but the
// 9 -> 5 rows of the last one): one thread per OUTPUT element quad reads (dy, idx) once and writes
// its kh x kw window -- no window search, no re-reads.  `addend` (optional, dx's layout) is the
// gradient arriving through the skip connection (refine_net.py:123-126), fused instead of a
// separate read-modify-write pass over dx.
template <int KH, int KW>
__global__ void __launch_bounds__(256)
adaptive_maxpool_bwd_exact_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx,
                                  long long total4, int W, int C, int OH, int OW,
                                  const float* __restrict__ addend, int ald,
                                  float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  const int c = (int)(i % C4) * 4;
  long long t = i / C4;
  const int ox = (int)(t % OW);
  t /= OW;
  const int oy = (int)(t % OH);
  const long long n = t / OH;
  const int4 id = __ldg(reinterpret_cast<const int4*>(idx) + i);
  const F4 d = ld4(dy + i * 4);
  const int H = OH * KH;
#pragma unroll
  for (int r = 0; r < KH; ++r) {
#pragma unroll
    for (int q = 0; q < KW; ++q) {
      const int h = oy * KH + r, w = ox * KW + q;
      const int me = h * W + w;
      const size_t o = (((size_t)n * H + h) * W + w) * C + c;
      F4 v;
      v.v[0] = id.x == me ? d.v[0] : 0.f;
      v.v[1] = id.y == me ? d.v[1] : 0.f;
      v.v[2] = id.z == me ? d.v[2] : 0.f;
      v.v[3] = id.w == me ? d.v[3] : 0.f;
      if (addend) {                  // pixel rows of `ald` floats (a channel slice of a wider tensor)
        const F4 a = ld4(addend + (((size_t)n * H + h) * W + w) * ald + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) v.v[j] += a.v[j];
      }
      st4(dx + o, v);
    }
  }
}

__global__ void __launch_bounds__(256)
adaptive_maxpool_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx,
                            long long total4, int H, int W, int C, int OH, int OW,
                            const float* __restrict__ addend, int ald, float* __restrict__ dx) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  const long long pix = t;
  int w = (int)(t % W);
  t /= W;
  int h = (int)(t % H);
  int n = (int)(t / H);
  const int me = h * W + w;
  int oy_lo = max(0, (h * OH) / H - 1), oy_hi = min(OH - 1, ((h + 1) * OH + H - 1) / H);
  int ox_lo = max(0, (w * OW) / W - 1), ox_hi = min(OW - 1, ((w + 1) * OW + W - 1) / W);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    if (h < win_start(oy, H, OH) || h >= win_end(oy, H, OH)) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      if (w < win_start(ox, W, OW) || w >= win_end(ox, W, OW)) continue;
      size_t o = (((size_t)n * OH + oy) * OW + ox) * C + c;
      int4 id = __ldg(reinterpret_cast<const int4*>(idx + o));
      F4 d = ld4(dy + o);
      if (id.x == me) s[0] += d.v[0];
      if (id.y == me) s[1] += d.v[1];
      if (id.z == me) s[2] += d.v[2];
      if (id.w == me) s[3] += d.v[3];
    }
  }
  if (addend) {
    const F4 a = ld4(addend + (size_t)pix * ald + c);
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] += a.v[j];
  }
  *reinterpret_cast<float4*>(dx + i * 4) = make_float4(s[0], s[1], s[2], s[3]);
}

// torch upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0
__device__ __forceinline__ void bilinear_src(int d, float scale, int in, int& i0, int& i1,
                                             float& l0, float& l1) {
  float s = scale * ((float)d + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const float* __restrict__ x, long long total4, int H, int W, int C, int OH,
                    int OW, float sh, float sw, float* __restrict__ y, int ldy, int coff) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  int ox = (int)(t % OW);
  t /= OW;
  int oy = (int)(t % OH);
  int n = (int)(t / OH);
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bilinear_src(oy, sh, H, y0, y1, ly0, ly1);
  bilinear_src(ox, sw, W, x0, x1, lx0, lx1);
  const float* xp = x + (size_t)n * H * W * C + c;
  F4 v00 = ld4(xp + (size_t)(y0 * W + x0) * C), v01 = ld4(xp + (size_t)(y0 * W + x1) * C);
  F4 v10 = ld4(xp + (size_t)(y1 * W + x0) * C), v11 = ld4(xp + (size_t)(y1 * W + x1) * C);
  F4 o;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o.v[j] = ly0 * (lx0 * v00.v[j] + lx1 * v01.v[j]) + ly1 * (lx0 * v10.v[j] + lx1 * v11.v[j]);
  st4(y + (((size_t)n * OH + oy) * OW + ox) * ldy + coff + c, o);
}

// gather form of the bilinear backward: each input pixel sums over the (few) output pixels
// whose 2x2 footprint contains it.  Output range bounded from the scale.
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const float* __restrict__ dy, int lddy, int coff, long long total4, int H,
                    int W, int C, int OH, int OW, float sh, float sw, float* __restrict__ dx) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  int w = (int)(t % W);
  t /= W;
  int h = (int)(t % H);
  int n = (int)(t / H);
  // candidate outputs: src in (h-1, h+1)  =>  dst in ((h-0.5)/s - 0.5 - 1, (h+1.5)/s - 0.5 + 1)
  int oy_lo = max(0, (int)floorf(((float)h - 0.5f) / sh - 1.5f));
  int oy_hi = min(OH - 1, (int)ceilf(((float)h + 1.5f) / sh + 0.5f));
  int ox_lo = max(0, (int)floorf(((float)w - 0.5f) / sw - 1.5f));
  int ox_hi = min(OW - 1, (int)ceilf(((float)w + 1.5f) / sw + 0.5f));
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  // horizontal weights of the candidate columns once per thread (not once per candidate pair);
  // in chunks of kMaxCand columns (one chunk for every upscale factor below 5)
  constexpr int kMaxCand = 12;
  for (int xb = ox_lo; xb <= ox_hi; xb += kMaxCand) {
    float wxs[kMaxCand];
    const int nx = min(ox_hi - xb + 1, kMaxCand);
#pragma unroll
    for (int k = 0; k < kMaxCand; ++k) {
      wxs[k] = 0.f;
      if (k < nx) {
        int x0, x1;
        float lx0, lx1;
        bilinear_src(xb + k, sw, W, x0, x1, lx0, lx1);
        wxs[k] = (x0 == w ? lx0 : 0.f) + (x1 == w ? lx1 : 0.f);
      }
    }
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float ly0, ly1;
      bilinear_src(oy, sh, H, y0, y1, ly0, ly1);
      float wy = (y0 == h ? ly0 : 0.f) + (y1 == h ? ly1 : 0.f);
      if (wy == 0.f) continue;
      const float* row = dy + (((size_t)n * OH + oy) * OW + xb) * lddy + coff + c;
#pragma unroll
      for (int k = 0; k < kMaxCand; ++k) {
        if (k < nx && wxs[k] != 0.f) {
          F4 d = ld4(row + (size_t)k * lddy);
          const float wt = wy * wxs[k];
#pragma unroll
          for (int j = 0; j < 4; ++j) s[j] = fmaf(wt, d.v[j], s[j]);
        }
      }
    }
  }
  *reinterpret_cast<float4*>(dx + i * 4) = make_float4(s[0], s[1], s[2], s[3]);
}

// ---- exact 2x upscale (RefineNet decoder levels 0-2: 36x64 -> 72x128, ...).  The general kernels
// above spend their time on index arithmetic (one thread per output quad: three runtime divisions
// and two source computations for one 16-byte store; the backward walks a computed candidate
// range).  Here a thread owns one INPUT pixel quad: the forward writes the 2x2 outputs it is the
// nearest source of from its 3x3 neighbourhood, the backward gathers the 4x4 outputs that read it.
// Weights come from the same bilinear_src() and every output / sum is formed by the same
// expression in the same order, so results are bit-identical to the general kernels.
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const float* __restrict__ x, int total4, int H, int W, int C4,
                      float* __restrict__ y, int ldy, int coff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int cq = i % C4;
  int t = i / C4;
  const int w = t % W;
  t /= W;
  const int h = t % H;
  const int n = t / H;
  const int OW = 2 * W;
  const float4* xp = reinterpret_cast<const float4*>(x) + (size_t)n * H * W * C4 + cq;
  const int rows[3] = {max(h - 1, 0), h, min(h + 1, H - 1)};
  const int cols[3] = {max(w - 1, 0), w, min(w + 1, W - 1)};
  F4 v[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) v[a][b] = F4(__ldg(xp + (size_t)(rows[a] * W + cols[b]) * C4));
  // bilinear_src(2h, 0.5): sources (h-1, h), weights (0.25, 0.75); at h = 0 sources (0, 1), weights
  // (1, 0) -- taken here from (rows[0], rows[1]) = (0, 0), the zero weight meets a finite value
  // either way.  bilinear_src(2h+1, 0.5): sources (h, min(h+1, H-1)), weights (0.75, 0.25).
  const float ly[2][2] = {{h == 0 ? 1.f : 0.25f, h == 0 ? 0.f : 0.75f}, {0.75f, 0.25f}};
  const float lx[2][2] = {{w == 0 ? 1.f : 0.25f, w == 0 ? 0.f : 0.75f}, {0.75f, 0.25f}};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      F4 o;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o.v[j] = ly[dy][0] * (lx[dx][0] * v[dy][dx].v[j] + lx[dx][1] * v[dy][dx + 1].v[j]) +
                 ly[dy][1] * (lx[dx][0] * v[dy + 1][dx].v[j] + lx[dx][1] * v[dy + 1][dx + 1].v[j]);
      st4(y + (((size_t)n * 2 * H + 2 * h + dy) * OW + 2 * w + dx) * ldy + coff + cq * 4, o);
    }
}

__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const float* __restrict__ dy, int lddy, int coff, int total4, int H, int W,
                      int C4, float* __restrict__ dx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int cq = i % C4;
  int t = i / C4;
  const int w = t % W;
  t /= W;
  const int h = t % H;
  const int n = t / H;
  const int OH = 2 * H, OW = 2 * W;
  // the outputs that read input h are 2h-1 .. 2h+2, with the weights bilinear_src() gives them:
  // 0.25 (if h > 0), 0.75 (1 at h = 0: clamped source), 0.75 (1 at h = H-1: both sources), 0.25
  // (if h < H-1)
  const float wy[4] = {h > 0 ? 0.25f : 0.f, h == 0 ? 1.f : 0.75f, h == H - 1 ? 1.f : 0.75f,
                       h < H - 1 ? 0.25f : 0.f};
  const float wx[4] = {w > 0 ? 0.25f : 0.f, w == 0 ? 1.f : 0.75f, w == W - 1 ? 1.f : 0.75f,
                       w < W - 1 ? 0.25f : 0.f};
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  const float* base = dy + coff + cq * 4;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (wy[a] == 0.f) continue;
    const int oy = 2 * h - 1 + a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (wx[b] == 0.f) continue;
      const int ox = 2 * w - 1 + b;
      const F4 d = ld4(base + (((size_t)n * OH + oy) * OW + ox) * lddy);
      const float wt = wy[a] * wx[b];
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] = fmaf(wt, d.v[j], s[j]);
    }
  }
  *reinterpret_cast<float4*>(dx + (size_t)i * 4) = make_float4(s[0], s[1], s[2], s[3]);
}

__global__ void __launch_bounds__(256)
copy_channels_kernel(const float* __restrict__ x, long long total, int C, int ldx, int xoff,
                     float* __restrict__ y, int ldy, int coff, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long r = i / C;
  float v = __ldg(x + r * ldx + xoff + c);
  float* o = y + r * ldy + coff + c;
  *o = accumulate ? *o + v : v;
}

__global__ void __launch_bounds__(256)
copy_channels4_kernel(const float* __restrict__ x, long long total4, int C, int ldx, int xoff,
                      float* __restrict__ y, int ldy, int coff, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  int c = (int)(i % C4) * 4;
  long long r = i / C4;
  F4 v = ld4(x + r * ldx + xoff + c);
  float* o = y + r * ldy + coff + c;
  if (accumulate) {
    F4 a(*reinterpret_cast<float4*>(o));
#pragma unroll
    for (int j = 0; j < 4; ++j) v.v[j] += a.v[j];
  }
  st4(o, v);
}

}  // namespace

static int transpose_images(const float* x, int N, int R, int Cc, float* y, cudaStream_t s) {
  if (N == 0) return EVE_OK;
  dim3 block(32, 8);
  for (int n0 = 0; n0 < N; n0 += 65535) {
    int nb = N - n0 < 65535 ? N - n0 : 65535;
    dim3 grid(cdiv(Cc, 32), cdiv(R, 32), nb);
    transpose_kernel<<<grid, block, 0, s>>>(x + (size_t)n0 * R * Cc, R, Cc,
                                            y + (size_t)n0 * R * Cc);
    EVE_LAUNCH_CHECK();
  }
  return EVE_OK;
}

int nchw_to_nhwc(const float* x, int N, int C, int H, int W, float* y, cudaStream_t s) {
  if (C == 3 && (H * W) % 4 == 0 && N > 0) {
    const long long quads = (long long)N * (H * W / 4);
    nchw3_to_nhwc_kernel<<<cdiv(quads, 256), 256, 0, s>>>(x, quads, H * W / 4, y);
    EVE_LAUNCH_CHECK();
    return EVE_OK;
  }
  return transpose_images(x, N, C, H * W, y, s);
}
int nhwc_to_nchw(const float* x, int N, int C, int H, int W, float* y, cudaStream_t s) {
  return transpose_images(x, N, H * W, C, y, s);
}

int in_relu_maxpool(const float* x, int N, int H, int W, int C, const float* mean,
                    const float* rstd, float* y, int32_t* idx, cudaStream_t s) {
  int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "in_relu_maxpool: C must be a multiple of 4");
  long long total = (long long)N * OH * OW * (C / 4);
  in_relu_maxpool_kernel<<<cdiv(total, 256), 256, 0, s>>>(x, total, H, W, C, OH, OW, mean, rstd, y,
                                                          idx);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int maxpool_bwd_scatter(const float* dy, const int32_t* idx, int N, int H, int W, int OH, int OW,
                        int C, float* g, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "maxpool_bwd: C must be a multiple of 4");
  long long total = (long long)N * H * W * (C / 4);
  maxpool_bwd_kernel<<<cdiv(total, 256), 256, 0, s>>>(dy, idx, total, H, W, OH, OW, C, g);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int avgpool_fwd(const float* x, int N, int HW, int C, float* y, cudaStream_t s) {
  avgpool_fwd_kernel<<<N, 256, 0, s>>>(x, HW, C, y);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int avgpool_bwd(const float* dy, int N, int HW, int C, float* dx, cudaStream_t s) {
  long long total = (long long)N * HW * C;
  avgpool_bwd_kernel<<<cdiv(total, 256), 256, 0, s>>>(dy, total, HW, C, dx);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int adaptive_maxpool_fwd(const float* x, int N, int H, int W, int C, int OH, int OW, float* y,
                         int32_t* idx, cudaStream_t s, float* copy, int copy_ld) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "adaptive_maxpool: C must be a multiple of 4");
  long long total = (long long)N * OH * OW * (C / 4);
  EVE_REQUIRE(!copy || (copy_ld % 4 == 0 && ((uintptr_t)copy & 15) == 0), EVE_ERR_SHAPE,
              "adaptive_maxpool_fwd: the copy's row pitch and start must be multiples of 4 floats");
  adaptive_maxpool_fwd_kernel<<<cdiv(total, 256), 256, 0, s>>>(x, total, H, W, C, OH, OW, y, idx, copy,
                                                               copy_ld);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int adaptive_maxpool_bwd(const float* dy, const int32_t* idx, int N, int H, int W, int C, int OH,
                         int OW, float* dx, cudaStream_t s, const float* addend, int addend_ld) {
  EVE_REQUIRE(C % 4 == 0, EVE_ERR_SHAPE, "adaptive_maxpool: C must be a multiple of 4");
  const int ald = addend_ld ? addend_ld : C;
  EVE_REQUIRE(!addend || (ald % 4 == 0 && ((uintptr_t)addend & 15) == 0), EVE_ERR_SHAPE,
              "adaptive_maxpool_bwd: the addend's row pitch and start must be multiples of 4 floats");
  if (H == 2 * OH && W == 2 * OW) {
    long long total = (long long)N * OH * OW * (C / 4);
    adaptive_maxpool_bwd_exact_kernel<2, 2><<<cdiv(total, 256), 256, 0, s>>>(dy, idx, total, W, C, OH, OW,
                                                                             addend, ald, dx);
    EVE_LAUNCH_CHECK();
    return EVE_OK;
  }
  long long total = (long long)N * H * W * (C / 4);
  adaptive_maxpool_bwd_kernel<<<cdiv(total, 256), 256, 0, s>>>(dy, idx, total, H, W, C, OH, OW,
                                                               addend, ald, dx);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int upsample_bilinear_fwd(const float* x, int N, int H, int W, int C, int OH, int OW, float* y,
                          int ldy, int coff, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0 && ldy % 4 == 0 && coff % 4 == 0, EVE_ERR_SHAPE,
              "upsample_bilinear: channel counts must be multiples of 4");
  if (OH == 2 * H && OW == 2 * W && (long long)N * H * W * (C / 4) < (1ll << 31) && N > 0) {
    const int total4 = N * H * W * (C / 4);
    upsample2x_fwd_kernel<<<cdiv(total4, 256), 256, 0, s>>>(x, total4, H, W, C / 4, y, ldy, coff);
    EVE_LAUNCH_CHECK();
    return EVE_OK;
  }
  long long total = (long long)N * OH * OW * (C / 4);
  float sh = (float)H / (float)OH, sw = (float)W / (float)OW;
  upsample_fwd_kernel<<<cdiv(total, 256), 256, 0, s>>>(x, total, H, W, C, OH, OW, sh, sw, y, ldy,
                                                       coff);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int upsample_bilinear_bwd(const float* dy, int lddy, int coff, int N, int H, int W, int C, int OH,
                          int OW, float* dx, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0 && lddy % 4 == 0 && coff % 4 == 0, EVE_ERR_SHAPE,
              "upsample_bilinear: channel counts must be multiples of 4");
  if (OH == 2 * H && OW == 2 * W && (long long)N * H * W * (C / 4) < (1ll << 31) && N > 0) {
    const int total4 = N * H * W * (C / 4);
    upsample2x_bwd_kernel<<<cdiv(total4, 256), 256, 0, s>>>(dy, lddy, coff, total4, H, W, C / 4, dx);
    EVE_LAUNCH_CHECK();
    return EVE_OK;
  }
  long long total = (long long)N * H * W * (C / 4);
  float sh = (float)H / (float)OH, sw = (float)W / (float)OW;
  upsample_bwd_kernel<<<cdiv(total, 256), 256, 0, s>>>(dy, lddy, coff, total, H, W, C, OH, OW, sh,
                                                       sw, dx);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int copy_channels(const float* x, long long rows, int C, int ldx, int xoff, float* y, int ldy,
                  int coff, bool accumulate, cudaStream_t s) {
  if (C % 4 == 0 && ldx % 4 == 0 && xoff % 4 == 0 && ldy % 4 == 0 && coff % 4 == 0) {
    long long total4 = rows * (C / 4);
    copy_channels4_kernel<<<cdiv(total4, 256), 256, 0, s>>>(x, total4, C, ldx, xoff, y, ldy, coff,
                                                            accumulate ? 1 : 0);
    EVE_LAUNCH_CHECK();
    return EVE_OK;
  }
  long long total = rows * C;
  copy_channels_kernel<<<cdiv(total, 256), 256, 0, s>>>(x, total, C, ldx, xoff, y, ldy, coff,
                                                        accumulate ? 1 : 0);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

}  // namespace eve
