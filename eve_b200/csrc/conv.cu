// Convolution dispatch: picks the tcgen05 tensor-core kernel (conv_tc.cu) when the geometry
// is dense enough for it and the fp32 CUDA-core implicit GEMM (conv_simt.cu) otherwise, and
// derives the operand layouts each of them needs inside the caller's scratch slice.
#include <algorithm>
#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace eve {

namespace {
int g_mode = -1;

// ---------------------------------------------------------------- stem (7x7, Cin = 3) --
// torchvision ResNet.conv1 (eye_net.py:48-50): K = 7*7*3 = 147 is too thin for a TMA box per
// tap, so the patch matrix is materialised once ([pixels][192] 16-bit hi/lo planes, K padded
// with zeros) and the convolution / its weight gradient run as 1x1 tensor-core GEMMs.
constexpr int kStemK = 192;

__device__ __forceinline__ void split16(float v, int fmt, uint16_t& h, uint16_t& l) {
  if (fmt == TC_BF16) {
    __nv_bfloat16 hb = __float2bfloat16_rn(v);
    h = __bfloat16_as_ushort(hb);
    l = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(hb)));
  } else {
    __half hh = __float2half_rn(v);
    h = __half_as_ushort(hh);
    l = __half_as_ushort(__float2half_rn(v - __half2float(hh)));
  }
}

// one thread = one output pixel x 8 consecutive k (one 16-byte store per plane)
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ x, long long total, int H, int W, int OH, int OW,
                   int fmt, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kg = (int)(i % (kStemK / 8));
  long long pix = i / (kStemK / 8);
  const int ox = (int)(pix % OW);
  long long t = pix / OW;
  const int oy = (int)(t % OH);
  const int n = (int)(t / OH);
  const float* xp = x + (size_t)n * H * W * 3;
  uint16_t h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = kg * 8 + j;
    float v = 0.f;
    if (k < 147) {
      const int tap = k / 3, c = k - tap * 3;
      const int r = tap / 7, q = tap - r * 7;
      const int iy = oy * 2 + r - 3, ix = ox * 2 + q - 3;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xp + ((size_t)iy * W + ix) * 3 + c);
    }
    split16(v, fmt, h[j], l[j]);
  }
  reinterpret_cast<uint4*>(hi)[i] = *reinterpret_cast<uint4*>(h);
  if (lo) reinterpret_cast<uint4*>(lo)[i] = *reinterpret_cast<uint4*>(l);
}

// OIHW [64][3][7][7] -> [64][192] (k = (r*7+q)*3 + c, zero padded)
__global__ void stem_prep_weights_kernel(const float* __restrict__ w, int Cout, int fmt,
                                         float scale, uint16_t* __restrict__ hi,
                                         uint16_t* __restrict__ lo) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * kStemK) return;
  const int k = i % kStemK, co = i / kStemK;
  float v = 0.f;
  if (k < 147) {
    const int tap = k / 3, c = k - tap * 3;
    v = w[((size_t)co * 3 + c) * 49 + tap] * scale;
  }
  uint16_t h, l;
  split16(v, fmt, h, l);
  hi[i] = h;
  if (lo) lo[i] = l;
}

// dw_oihw[co][c][r][q] (+)= sum_z part[z][co][(r*7+q)*3 + c]
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ part, int S, int Cout,
                                         float* __restrict__ dw, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 147) return;
  const int tap = i % 49;
  const int c = (i / 49) % 3;
  const int co = i / 147;
  float a = 0.f;
  for (int z = 0; z < S; ++z) a += part[((size_t)z * Cout + co) * kStemK + tap * 3 + c];
  dw[i] = accumulate ? dw[i] + a : a;
}

inline bool is_stem(const ConvGeom& g) {
  return g.KH == 7 && g.KW == 7 && g.stride == 2 && g.pad == 3 && g.Cin == 3 && g.Cout == 64 &&
         g.OW <= 64 && g.OW >= 1;
}
inline ConvGeom stem_gemm(const ConvGeom& g) {
  return make_conv(g.N, g.OH, g.OW, kStemK, g.Cout, 1, 1, 0);
}

struct Carve {
  char* p;
  char* end;
  template <typename T>
  T* get(size_t n) {
    size_t bytes = align_up(n * sizeof(T), 1024);
    char* r = p;
    p += bytes;
    return p <= end ? (T*)r : nullptr;
  }
};

// dgrad of a stride-1 "same" convolution is itself a stride-1 "same" convolution of dy with
// the flipped, channel-transposed filter.
ConvGeom dgrad_as_fwd(const ConvGeom& g) {
  return make_conv(g.N, g.OH, g.OW, g.Cout, g.Cin, g.KH, 1, g.KH - 1 - g.pad);
}
}  // namespace

// debugging aid: EVE_B200_TC_MASK selects which passes may use the tensor-core kernels
// (bit 0 forward, bit 1 data gradient, bit 2 weight gradient; default all)
static int tc_mask() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("EVE_B200_TC_MASK");
    m = e ? atoi(e) : 7;
  }
  return m;
}

int conv_mode() {
  if (g_mode < 0) {
    const char* e = getenv("EVE_B200_CONV_MODE");
    g_mode = e ? atoi(e) : 1;
    if (g_mode < 0 || g_mode > 2) g_mode = 1;
  }
  return g_mode;
}
void set_conv_mode(int mode) { g_mode = (mode < 0 || mode > 2) ? 1 : mode; }

size_t conv_scratch_bytes(size_t max_in, size_t max_out, size_t max_w, size_t wgrad_floats) {
  size_t m = max_in > max_out ? max_in : max_out;
  size_t weights = 2 * align_up(max_w * sizeof(float), 1024);          // fp32 layout or hi+lo
  size_t planes = 2 * align_up(m * 2, 1024);                           // hi + lo of one operand
  size_t partial = align_up(wgrad_floats * sizeof(float), 1024);
  return weights + 2 * planes + partial + 4096;
}

size_t conv_partial_floats(const ConvGeom& g) {
  size_t a = conv_wgrad_scratch_floats(g);
  size_t b = colsum_scratch_floats((long long)g.N * g.OH * g.OW, g.Cout);
  size_t c = conv_tc_wgrad_partial_floats(g);
  if (is_stem(g)) c = std::max(c, conv_tc_wgrad_partial_floats(stem_gemm(g)));
  return std::max(a, std::max(b, c));
}

size_t conv_operand_elems(const ConvGeom& g) {
  size_t m = std::max((size_t)g.in_elems(), (size_t)g.out_elems());
  if (is_stem(g)) m = std::max(m, (size_t)stem_gemm(g).in_elems());
  return m;
}

int conv_fwd(const ConvGeom& g, const float* x, const float* w, const float* bias,
             const float* addend, float* y, const ConvScratch& sc, cudaStream_t s) {
  Carve c{sc.base, sc.base + sc.bytes};
  const size_t wel = (size_t)g.Cout * g.K();
  const int mode = conv_mode();
  if (mode != 0 && (tc_mask() & 1) && conv_tc_supported(g)) {
    const int npass = mode == 1 ? 3 : 1;
    uint16_t* w_hi = c.get<uint16_t>(wel);
    uint16_t* w_lo = c.get<uint16_t>(wel);
    uint16_t* x_hi = c.get<uint16_t>((size_t)g.in_elems());
    uint16_t* x_lo = c.get<uint16_t>((size_t)g.in_elems());
    EVE_REQUIRE(x_lo, EVE_ERR_WORKSPACE, "conv_fwd: scratch too small");
    // forward, split mode: fp16 planes (22 mantissa bits for hi + lo); weights pre-scaled by
    // 2^6 so that their lo parts stay normal, undone exactly in the epilogue.  The single-pass
    // mode keeps bf16.
    const int fmt = npass == 3 ? TC_F16 : TC_BF16;
    const float wscale = npass == 3 ? 64.f : 1.f;
    EVE_TRY(conv_tc_prep_weights(g, w, false, w_hi, npass == 3 ? w_lo : nullptr, fmt, wscale, s));
    EVE_TRY(split_planes(x, g.in_elems(), x_hi, npass == 3 ? x_lo : nullptr, fmt, s));
    ProfScope prof(PROF_CONV_FWD, 2.0 * g.out_elems() * (double)g.K(),
                   4.0 * (g.in_elems() + g.out_elems() + (double)wel), s);
    return conv_tc_run(g, x_hi, x_lo, w_hi, w_lo, bias, addend, y, npass, fmt, 1.f / wscale, s);
  }
  if (mode != 0 && (tc_mask() & 1) && is_stem(g)) {
    const int npass = mode == 1 ? 3 : 1;
    const int fmt = npass == 3 ? TC_F16 : TC_BF16;
    const float wscale = npass == 3 ? 64.f : 1.f;
    const ConvGeom gg = stem_gemm(g);
    uint16_t* w_hi = c.get<uint16_t>((size_t)g.Cout * kStemK);
    uint16_t* w_lo = c.get<uint16_t>((size_t)g.Cout * kStemK);
    uint16_t* x_hi = c.get<uint16_t>((size_t)gg.in_elems());
    uint16_t* x_lo = c.get<uint16_t>((size_t)gg.in_elems());
    EVE_REQUIRE(x_lo, EVE_ERR_WORKSPACE, "conv_fwd(stem): scratch too small");
    stem_prep_weights_kernel<<<cdiv(g.Cout * kStemK, 256), 256, 0, s>>>(
        w, g.Cout, fmt, wscale, w_hi, npass == 3 ? w_lo : nullptr);
    EVE_LAUNCH_CHECK();
    const long long total = (long long)g.N * g.OH * g.OW * (kStemK / 8);
    stem_im2col_kernel<<<cdiv(total, 256), 256, 0, s>>>(x, total, g.H, g.W, g.OH, g.OW, fmt, x_hi,
                                                       npass == 3 ? x_lo : nullptr);
    EVE_LAUNCH_CHECK();
    ProfScope prof(PROF_CONV_FWD, 2.0 * g.out_elems() * (double)g.K(),
                   4.0 * (g.in_elems() + g.out_elems() + (double)wel), s);
    return conv_tc_run(gg, x_hi, x_lo, w_hi, w_lo, bias, addend, y, npass, fmt, 1.f / wscale, s);
  }
  float* wf = c.get<float>(wel);
  EVE_REQUIRE(wf, EVE_ERR_WORKSPACE, "conv_fwd: scratch too small");
  EVE_TRY(conv_prep_weights(g, w, wf, nullptr, s));
  return conv_fwd_simt(g, x, wf, bias, addend, y, g.Cout, s);
}

int conv_dgrad(const ConvGeom& g, const float* dy, const float* w, const float* addend, float* dx,
               const ConvScratch& sc, cudaStream_t s) {
  Carve c{sc.base, sc.base + sc.bytes};
  const size_t wel = (size_t)g.Cout * g.K();
  const int mode = conv_mode();
  if (mode != 0 && (tc_mask() & 2) && g.stride == 1) {
    ConvGeom f = dgrad_as_fwd(g);
    if (conv_tc_supported(f) && f.OH == g.H && f.OW == g.W) {
      const int npass = mode == 1 ? 3 : 1;
      uint16_t* w_hi = c.get<uint16_t>(wel);
      uint16_t* w_lo = c.get<uint16_t>(wel);
      uint16_t* d_hi = c.get<uint16_t>((size_t)g.out_elems());
      uint16_t* d_lo = c.get<uint16_t>((size_t)g.out_elems());
      EVE_REQUIRE(d_lo, EVE_ERR_WORKSPACE, "conv_dgrad: scratch too small");
      EVE_TRY(conv_tc_prep_weights(g, w, true, w_hi, npass == 3 ? w_lo : nullptr, TC_BF16, 1.f, s));
      EVE_TRY(split_planes(dy, g.out_elems(), d_hi, npass == 3 ? d_lo : nullptr, TC_BF16, s));
      ProfScope prof(PROF_CONV_DGRAD, 2.0 * g.out_elems() * (double)g.K(),
                     4.0 * (g.in_elems() + g.out_elems() + (double)wel), s);
      return conv_tc_run(f, d_hi, d_lo, w_hi, w_lo, nullptr, addend, dx, npass, TC_BF16, 1.f, s);
    }
  }
  if (mode != 0 && (tc_mask() & 2) && conv_tc_dgrad_s2_supported(g)) {
    const int npass = mode == 1 ? 3 : 1;
    uint16_t* w_hi = c.get<uint16_t>(wel);
    uint16_t* w_lo = c.get<uint16_t>(wel);
    uint16_t* d_hi = c.get<uint16_t>((size_t)g.out_elems());
    uint16_t* d_lo = c.get<uint16_t>((size_t)g.out_elems());
    EVE_REQUIRE(d_lo, EVE_ERR_WORKSPACE, "conv_dgrad: scratch too small");
    EVE_TRY(conv_tc_prep_weights(g, w, true, w_hi, npass == 3 ? w_lo : nullptr, TC_BF16, 1.f, s));
    EVE_TRY(split_planes(dy, g.out_elems(), d_hi, npass == 3 ? d_lo : nullptr, TC_BF16, s));
    if (g.KH == 1) {
      // only even/even pixels receive data: the rest is the addend (or zero)
      if (addend) {
        if (addend != dx)
          EVE_CUDA(cudaMemcpyAsync(dx, addend, (size_t)g.in_elems() * sizeof(float),
                                   cudaMemcpyDeviceToDevice, s));
      } else {
        EVE_TRY(fill_zero(dx, g.in_elems(), s));
      }
    }
    ProfScope prof(PROF_CONV_DGRAD, 2.0 * g.out_elems() * (double)g.K(),
                   4.0 * (g.in_elems() + g.out_elems() + (double)wel), s);
    return conv_tc_dgrad_s2_run(g, d_hi, d_lo, w_hi, w_lo, addend, dx, npass, s);
  }
  float* wd = c.get<float>(wel);
  EVE_REQUIRE(wd, EVE_ERR_WORKSPACE, "conv_dgrad: scratch too small");
  EVE_TRY(conv_prep_weights(g, w, nullptr, wd, s));
  return conv_dgrad_simt(g, dy, g.Cout, wd, addend, dx, s);
}

int conv_wgrad(const ConvGeom& g, const float* x, const float* dy, float* dw, float* dbias,
               bool accumulate, const ConvScratch& sc, cudaStream_t s) {
  Carve c{sc.base, sc.base + sc.bytes};
  const int mode = conv_mode();
  if (dw && mode != 0 && (tc_mask() & 4) && conv_tc_wgrad_supported(g)) {
    const int npass = mode == 1 ? 3 : 1;
    size_t pf = conv_tc_wgrad_partial_floats(g);
    size_t cs = colsum_scratch_floats((long long)g.N * g.OH * g.OW, g.Cout);
    float* part = c.get<float>(pf > cs ? pf : cs);
    uint16_t* d_hi = c.get<uint16_t>((size_t)g.out_elems());
    uint16_t* d_lo = c.get<uint16_t>((size_t)g.out_elems());
    uint16_t* x_hi = c.get<uint16_t>((size_t)g.in_elems());
    uint16_t* x_lo = c.get<uint16_t>((size_t)g.in_elems());
    EVE_REQUIRE(x_lo, EVE_ERR_WORKSPACE, "conv_wgrad: scratch too small");
    // tcgen05 kind::f16 needs both operands in ONE 16-bit format (an fp16 x bf16 descriptor is an
    // illegal instruction on sm_100a), so x is split again as bf16 next to the bf16 dy planes
    const int xfmt = TC_BF16;
    EVE_TRY(split_planes(dy, g.out_elems(), d_hi, npass == 3 ? d_lo : nullptr, TC_BF16, s));
    EVE_TRY(split_planes(x, g.in_elems(), x_hi, npass == 3 ? x_lo : nullptr, xfmt, s));
    int splits = 0;
    {
      ProfScope prof(PROF_CONV_WGRAD, 2.0 * g.out_elems() * (double)g.K(),
                     4.0 * (g.in_elems() + g.out_elems() + (double)g.Cout * g.K()), s);
      EVE_TRY(conv_tc_wgrad_run(g, d_hi, d_lo, x_hi, x_lo, part, npass, &splits, s, xfmt));
      EVE_TRY(wgrad_reduce(part, splits, g, dw, accumulate, s));
    }
    if (dbias)
      EVE_TRY(colsum(dy, (long long)g.N * g.OH * g.OW, g.Cout, g.Cout, dbias, part, accumulate, s));
    return EVE_OK;
  }
  if (dw && mode != 0 && (tc_mask() & 4) && is_stem(g) && conv_tc_wgrad_supported(stem_gemm(g))) {
    const int npass = mode == 1 ? 3 : 1;
    const ConvGeom gg = stem_gemm(g);
    size_t pf = conv_tc_wgrad_partial_floats(gg);
    size_t cs = colsum_scratch_floats((long long)g.N * g.OH * g.OW, g.Cout);
    float* part = c.get<float>(pf > cs ? pf : cs);
    uint16_t* d_hi = c.get<uint16_t>((size_t)g.out_elems());
    uint16_t* d_lo = c.get<uint16_t>((size_t)g.out_elems());
    uint16_t* x_hi = c.get<uint16_t>((size_t)gg.in_elems());
    uint16_t* x_lo = c.get<uint16_t>((size_t)gg.in_elems());
    EVE_REQUIRE(x_lo, EVE_ERR_WORKSPACE, "conv_wgrad(stem): scratch too small");
    EVE_TRY(split_planes(dy, g.out_elems(), d_hi, npass == 3 ? d_lo : nullptr, TC_BF16, s));
    const long long total = (long long)g.N * g.OH * g.OW * (kStemK / 8);
    stem_im2col_kernel<<<cdiv(total, 256), 256, 0, s>>>(x, total, g.H, g.W, g.OH, g.OW, TC_BF16,
                                                       x_hi, npass == 3 ? x_lo : nullptr);
    EVE_LAUNCH_CHECK();
    int splits = 0;
    {
      ProfScope prof(PROF_CONV_WGRAD, 2.0 * g.out_elems() * (double)g.K(),
                     4.0 * (g.in_elems() + g.out_elems() + (double)g.Cout * g.K()), s);
      EVE_TRY(conv_tc_wgrad_run(gg, d_hi, d_lo, x_hi, x_lo, part, npass, &splits, s));
      stem_wgrad_reduce_kernel<<<cdiv(g.Cout * 147, 256), 256, 0, s>>>(part, splits, g.Cout, dw,
                                                                      accumulate ? 1 : 0);
      EVE_LAUNCH_CHECK();
    }
    if (dbias)
      EVE_TRY(colsum(dy, (long long)g.N * g.OH * g.OW, g.Cout, g.Cout, dbias, part, accumulate, s));
    return EVE_OK;
  }
  size_t need = conv_wgrad_scratch_floats(g);
  size_t cs = colsum_scratch_floats((long long)g.N * g.OH * g.OW, g.Cout);
  float* part = c.get<float>(need > cs ? need : cs);
  EVE_REQUIRE(part, EVE_ERR_WORKSPACE, "conv_wgrad: scratch too small");
  if (dw) EVE_TRY(conv_wgrad_simt(g, x, dy, g.Cout, dw, part, accumulate, s));
  if (dbias)
    EVE_TRY(colsum(dy, (long long)g.N * g.OH * g.OW, g.Cout, g.Cout, dbias, part, accumulate, s));
  return EVE_OK;
}

// Weight gradient and data gradient of one convolution sharing ONE split of dy (both tensor-core
// passes read the same bf16 hi/lo planes).  Falls back to the two separate entry points when
// either pass is not taken by the tensor-core kernels.
int conv_bwd(const ConvGeom& g, const float* x, const float* dy, const float* w, float* dw,
             float* dbias, bool accumulate, const float* addend, float* dx, const ConvScratch& sc,
             cudaStream_t s) {
  const int mode = conv_mode();
  const bool s1 = g.stride == 1 && conv_tc_supported(dgrad_as_fwd(g));
  const bool s2 = conv_tc_dgrad_s2_supported(g);
  const bool fused = dw && dx && mode != 0 && (tc_mask() & 6) == 6 && conv_tc_wgrad_supported(g) &&
                     (s1 || s2);
  if (!fused) {
    if (dw || dbias) EVE_TRY(conv_wgrad(g, x, dy, dw, dbias, accumulate, sc, s));
    if (dx) EVE_TRY(conv_dgrad(g, dy, w, addend, dx, sc, s));
    return EVE_OK;
  }
  Carve c{sc.base, sc.base + sc.bytes};
  const int npass = mode == 1 ? 3 : 1;
  const size_t wel = (size_t)g.Cout * g.K();
  size_t pf = conv_tc_wgrad_partial_floats(g);
  size_t cs = colsum_scratch_floats((long long)g.N * g.OH * g.OW, g.Cout);
  float* part = c.get<float>(pf > cs ? pf : cs);
  uint16_t* w_hi = c.get<uint16_t>(wel);
  uint16_t* w_lo = c.get<uint16_t>(wel);
  uint16_t* d_hi = c.get<uint16_t>((size_t)g.out_elems());
  uint16_t* d_lo = c.get<uint16_t>((size_t)g.out_elems());
  uint16_t* x_hi = c.get<uint16_t>((size_t)g.in_elems());
  uint16_t* x_lo = c.get<uint16_t>((size_t)g.in_elems());
  EVE_REQUIRE(x_lo, EVE_ERR_WORKSPACE, "conv_bwd: scratch too small");
  const int xfmt = TC_BF16;     // same format as dy: mixed fp16 x bf16 MMAs are illegal
  EVE_TRY(split_planes(dy, g.out_elems(), d_hi, npass == 3 ? d_lo : nullptr, TC_BF16, s));
  EVE_TRY(split_planes(x, g.in_elems(), x_hi, npass == 3 ? x_lo : nullptr, xfmt, s));
  EVE_TRY(conv_tc_prep_weights(g, w, true, w_hi, npass == 3 ? w_lo : nullptr, TC_BF16, 1.f, s));
  const double flops = 2.0 * g.out_elems() * (double)g.K();
  const double bytes = 4.0 * (g.in_elems() + g.out_elems() + (double)wel);
  {
    int splits = 0;
    ProfScope prof(PROF_CONV_WGRAD, flops, bytes, s);
    EVE_TRY(conv_tc_wgrad_run(g, d_hi, d_lo, x_hi, x_lo, part, npass, &splits, s, xfmt));
    EVE_TRY(wgrad_reduce(part, splits, g, dw, accumulate, s));
  }
  if (dbias)
    EVE_TRY(colsum(dy, (long long)g.N * g.OH * g.OW, g.Cout, g.Cout, dbias, part, accumulate, s));
  if (s2 && g.KH == 1) {
    if (addend) {
      if (addend != dx)
        EVE_CUDA(cudaMemcpyAsync(dx, addend, (size_t)g.in_elems() * sizeof(float),
                                 cudaMemcpyDeviceToDevice, s));
    } else {
      EVE_TRY(fill_zero(dx, g.in_elems(), s));
    }
  }
  ProfScope prof(PROF_CONV_DGRAD, flops, bytes, s);
  if (s1)
    return conv_tc_run(dgrad_as_fwd(g), d_hi, d_lo, w_hi, w_lo, nullptr, addend, dx, npass, TC_BF16,
                       1.f, s);
  return conv_tc_dgrad_s2_run(g, d_hi, d_lo, w_hi, w_lo, addend, dx, npass, s);
}

}  // namespace eve
