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
constexpr int kStemK = 160;        // 147 rounded up to five 32-channel K chunks

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

// One block = one output row of one image.  The seven input rows it touches are staged in shared
// memory with coalesced loads (the gather itself then never leaves the SM); every work item is
// one output pixel x 8 consecutive k, i.e. one 16-byte store per plane, consecutive items writing
// consecutive addresses.
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ x, int H, int W, int OH, int OW, int fmt,
                   uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  extern __shared__ float srow[];                 // [7][W * 3]
  // source of patch element k = (r * 7 + q) * 3 + c of output column ox inside the staged rows:
  // srow[r * rowlen + (2 ox + q - 3) * 3 + c] = srow[tab_off[k] + 6 ox]; decoded once per block
  // instead of per element (the kernel was bound by this index arithmetic: 404 us for 1.35 GB)
  __shared__ int tab[kStemK];                     // (offset << 8) | q, q = 127: zero padding of K
  const int oy = blockIdx.x % OH;
  const int n = blockIdx.x / OH;
  const int rowlen = W * 3;
  if (threadIdx.x < kStemK) {
    const int k = threadIdx.x;
    int e = 127;
    if (k < 147) {
      const int tap = k / 3, c = k - tap * 3;
      const int r = tap / 7, q = tap - r * 7;
      e = ((r * rowlen + 3 * q - 9 + c) << 8) | q;
    }
    tab[k] = e;
  }
  for (int r = 0; r < 7; ++r) {
    const int iy = oy * 2 + r - 3;
    float* dst = srow + r * rowlen;
    if (iy >= 0 && iy < H) {
      const float* src = x + ((size_t)n * H + iy) * rowlen;
      for (int i = threadIdx.x; i < rowlen; i += blockDim.x) dst[i] = __ldg(src + i);
    } else {
      for (int i = threadIdx.x; i < rowlen; i += blockDim.x) dst[i] = 0.f;
    }
  }
  __syncthreads();
  constexpr int kGroups = kStemK / 8;
  const size_t out0 = ((size_t)n * OH + oy) * OW * kGroups;
  for (int item = threadIdx.x; item < OW * kGroups; item += blockDim.x) {
    const int ox = item / kGroups;
    const int kg = item - ox * kGroups;
    uint16_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = tab[kg * 8 + j];
      const int q = e & 255;
      const int ix = ox * 2 + q - 3;
      float v = 0.f;
      if (q != 127 && ix >= 0 && ix < W) v = srow[(e >> 8) + 6 * ox];
      split16(v, fmt, h[j], l[j]);
    }
    reinterpret_cast<uint4*>(hi)[out0 + item] = *reinterpret_cast<uint4*>(h);
    if (lo) reinterpret_cast<uint4*>(lo)[out0 + item] = *reinterpret_cast<uint4*>(l);
  }
}

// Stem windows (conv_tc_stem_run): 16-bit hi/lo planes of the zero-padded input, either
//   mode 1: [N][H + 6][OW][32]: the window of output column ox (input pixels 2 ox - 3 .. 2 ox + 4,
//           4 channels each, channel 3 and everything outside the image zero), or
//   mode 2: [N][H + 6][W + 8][4]: the padded image itself (3 zero pixels left, 5 right); windows are
//           read overlapping, 2 pixels apart.
// One thread writes 8 values (two pixels) = one 16-byte store per plane.
__global__ void __launch_bounds__(256)
stem_windows_kernel(const float* __restrict__ x, int N, int H, int W, int mode, int fmt,
                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const int Hp = H + 6;
  const int groups = mode == 1 ? (W / 2) * 4 : (W + 8) / 2;   // 16-byte groups per padded row
  const long long total = (long long)N * Hp * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gidx = (int)(i % groups);
    const long long row = i / groups;
    const int ihp = (int)(row % Hp);
    const int n = (int)(row / Hp);
    const int ih = ihp - 3;
    int ix0;                                       // first of the two input pixels of this group
    if (mode == 1) {
      const int ox = gidx >> 2, pg = gidx & 3;
      ix0 = 2 * ox - 3 + 2 * pg;
    } else {
      ix0 = 2 * gidx - 3;
    }
    uint16_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ix = ix0 + (j >> 2), c = j & 3;
      float v = 0.f;
      // mode 1: the 8th pixel of a window (pg == 3, second pixel) is outside the 7-tap filter; its
      // weights are zero, the value is irrelevant but kept finite
      if (c < 3 && ih >= 0 && ih < H && ix >= 0 && ix < W)
        v = __ldg(x + (((size_t)n * H + ih) * W + ix) * 3 + c);
      split16(v, fmt, h[j], l[j]);
    }
    reinterpret_cast<uint4*>(hi)[i] = *reinterpret_cast<uint4*>(h);
    if (lo) reinterpret_cast<uint4*>(lo)[i] = *reinterpret_cast<uint4*>(l);
  }
}

// OIHW [64][3][7][7] -> [64][7][32] (k = r*32 + q*4 + c; q == 7 and c == 3 zero)
__global__ void stem_prep_weights_win_kernel(const float* __restrict__ w, int Cout, int fmt,
                                             float scale, uint16_t* __restrict__ hi,
                                             uint16_t* __restrict__ lo) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 224) return;
  const int k = i % 224, co = i / 224;
  const int r = k >> 5, q = (k >> 2) & 7, c = k & 3;
  float v = 0.f;
  if (q < 7 && c < 3) v = w[((size_t)co * 3 + c) * 49 + r * 7 + q] * scale;
  uint16_t h, l;
  split16(v, fmt, h, l);
  hi[i] = h;
  if (lo) lo[i] = l;
}

// OIHW [64][3][7][7] -> [64][kStemK] (k = (r*7+q)*3 + c, zero padded)
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
         g.OW <= 64 && g.OW >= 1 && g.W <= 512;
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

// ---- prepared weights: a caller that applies the same convolution many times (the ConvRNN cells
// step through time) prepares the tensor-core weight layout once, at the TOP of the conv scratch,
// and conv_fwd / conv_dgrad find it here by weight pointer instead of re-deriving it per call.
struct PreparedW {
  const float* w;
  int dgrad, npass;
  const uint16_t *hi, *lo;
};
constexpr int kMaxPrepared = 192;
static thread_local PreparedW g_prepared[kMaxPrepared];
static thread_local int g_nprepared = 0;

static const PreparedW* find_prepared(const float* w, bool dgrad, int npass) {
  for (int i = 0; i < g_nprepared; ++i)
    if (g_prepared[i].w == w && g_prepared[i].dgrad == (dgrad ? 1 : 0) && g_prepared[i].npass == npass)
      return &g_prepared[i];
  return nullptr;
}

// scratch layout of the split tensor-core paths (shared by the kernels' callers and by the
// producers that write operand planes in place, see conv_x_planes_*)
struct FwdCarve {
  uint16_t *w_hi, *w_lo, *x_hi, *x_lo;
};
static bool carve_fwd(const ConvGeom& g, const ConvScratch& sc, FwdCarve& f) {
  Carve c{sc.base, sc.base + sc.bytes};
  const size_t wel = (size_t)g.Cout * g.K();
  f.w_hi = c.get<uint16_t>(wel);
  f.w_lo = c.get<uint16_t>(wel);
  f.x_hi = c.get<uint16_t>((size_t)g.in_elems());
  f.x_lo = c.get<uint16_t>((size_t)g.in_elems());
  return f.x_lo != nullptr;
}
struct BwdCarve {
  float* part;
  uint16_t *w_hi, *w_lo, *d_hi, *d_lo, *x_hi, *x_lo;
};
static bool carve_bwd(const ConvGeom& g, const ConvScratch& sc, BwdCarve& b) {
  Carve c{sc.base, sc.base + sc.bytes};
  const size_t wel = (size_t)g.Cout * g.K();
  size_t pf = conv_tc_wgrad_partial_floats(g);
  size_t cs = colsum_scratch_floats((long long)g.N * g.OH * g.OW, g.Cout);
  b.part = c.get<float>(pf > cs ? pf : cs);
  b.w_hi = c.get<uint16_t>(wel);
  b.w_lo = c.get<uint16_t>(wel);
  b.d_hi = c.get<uint16_t>((size_t)g.out_elems());
  b.d_lo = c.get<uint16_t>((size_t)g.out_elems());
  b.x_hi = c.get<uint16_t>((size_t)g.in_elems());
  b.x_lo = c.get<uint16_t>((size_t)g.in_elems());
  return b.x_lo != nullptr;
}

static bool bwd_dgrad_s1(const ConvGeom& g) {
  return g.stride == 1 && conv_tc_supported(dgrad_as_fwd(g));
}

bool conv_x_fusable(const ConvGeom& g) {
  if (!get_option(OPT_FUSED_PLANES) || conv_mode() != 1 || tc_mask() != 7) return false;
  if (!conv_tc_supported(g) || !conv_tc_wgrad_supported(g)) return false;
  return bwd_dgrad_s1(g) || conv_tc_dgrad_s2_supported(g);
}
void conv_x_planes_fwd(const ConvGeom& g, const ConvScratch& sc, void** hi, void** lo) {
  FwdCarve f;
  carve_fwd(g, sc, f);
  *hi = f.x_hi;
  *lo = f.x_lo;
}
void conv_x_planes_bwd(const ConvGeom& g, const ConvScratch& sc, void** hi, void** lo) {
  BwdCarve b;
  carve_bwd(g, sc, b);
  *hi = b.x_hi;
  *lo = b.x_lo;
}

int conv_fwd(const ConvGeom& g, const float* x, const float* w, const float* bias,
             const float* addend, float* y, const ConvScratch& sc, cudaStream_t s) {
  Carve c{sc.base, sc.base + sc.bytes};
  const size_t wel = (size_t)g.Cout * g.K();
  const int mode = conv_mode();
  EVE_REQUIRE(x || conv_x_fusable(g), EVE_ERR_NULL,
              "conv_fwd: x is NULL but the convolution does not take pre-split operand planes");
  if (mode != 0 && (tc_mask() & 1) && conv_tc_supported(g)) {
    const int npass = mode == 1 ? 3 : 1;
    FwdCarve f;
    EVE_REQUIRE(carve_fwd(g, sc, f), EVE_ERR_WORKSPACE, "conv_fwd: scratch too small");
    // forward, split mode: fp16 planes (22 mantissa bits for hi + lo); weights pre-scaled by
    // 2^6 so that their lo parts stay normal, undone exactly in the epilogue.  The single-pass
    // mode keeps bf16.
    const int fmt = npass == 3 ? TC_F16 : TC_BF16;
    const float wscale = npass == 3 ? 64.f : 1.f;
    if (const PreparedW* pw = find_prepared(w, false, npass)) {
      f.w_hi = const_cast<uint16_t*>(pw->hi);
      f.w_lo = const_cast<uint16_t*>(pw->lo);
    } else {
      EVE_TRY(conv_tc_prep_weights(g, w, false, f.w_hi, npass == 3 ? f.w_lo : nullptr, fmt, wscale, s));
    }
    if (x) EVE_TRY(split_planes(x, g.in_elems(), f.x_hi, npass == 3 ? f.x_lo : nullptr, fmt, s));
    ProfScope prof(PROF_CONV_FWD, 2.0 * g.out_elems() * (double)g.K(),
                   4.0 * (g.in_elems() + g.out_elems() + (double)wel), s, &g);
    return conv_tc_run(g, f.x_hi, f.x_lo, f.w_hi, f.w_lo, bias, addend, y, npass, fmt,
                       1.f / wscale, s);
  }
  if (mode != 0 && (tc_mask() & 1) && is_stem(g)) {
    const int npass = mode == 1 ? 3 : 1;
    const int fmt = npass == 3 ? TC_F16 : TC_BF16;
    const float wscale = npass == 3 ? 64.f : 1.f;
    const ConvGeom gg = stem_gemm(g);
    const int win = get_option(OPT_STEM_WINDOWS);
    if (win != 0 && g.H % 2 == 0 && g.W % 2 == 0 && g.W % 8 == 0 && !addend) {
      // no im2col matrix: 32-value windows of the zero-padded planes, seven taps (conv_tc_stem_run)
      const int Hp = g.H + 6;
      const size_t row_el = win == 1 ? (size_t)(g.W / 2) * 32 : (size_t)(g.W + 8) * 4;
      const size_t pel = (size_t)g.N * Hp * row_el;      // <= the im2col matrix the scratch is sized for
      uint16_t* w_hi = c.get<uint16_t>((size_t)g.Cout * 224);
      uint16_t* w_lo = c.get<uint16_t>((size_t)g.Cout * 224);
      uint16_t* x_hi = c.get<uint16_t>(pel);
      uint16_t* x_lo = c.get<uint16_t>(pel);
      EVE_REQUIRE(x_lo, EVE_ERR_WORKSPACE, "conv_fwd(stem windows): scratch too small");
      stem_prep_weights_win_kernel<<<cdiv(g.Cout * 224, 256), 256, 0, s>>>(
          w, g.Cout, fmt, wscale, w_hi, npass == 3 ? w_lo : nullptr);
      EVE_LAUNCH_CHECK();
      stem_windows_kernel<<<kNumSMs * 8, 256, 0, s>>>(x, g.N, g.H, g.W, win, fmt, x_hi,
                                                     npass == 3 ? x_lo : nullptr);
      EVE_LAUNCH_CHECK();
      ProfScope prof(PROF_CONV_FWD, 2.0 * g.out_elems() * (double)g.K(),
                     4.0 * (g.in_elems() + g.out_elems() + (double)wel), s, &g);
      return conv_tc_stem_run(g.N, g.H, g.W, x_hi, x_lo, win == 1 ? 64 : 16, (int)(row_el * 2), w_hi,
                              w_lo, bias, y, npass, fmt, 1.f / wscale, s);
    }
    uint16_t* w_hi = c.get<uint16_t>((size_t)g.Cout * kStemK);
    uint16_t* w_lo = c.get<uint16_t>((size_t)g.Cout * kStemK);
    uint16_t* x_hi = c.get<uint16_t>((size_t)gg.in_elems());
    uint16_t* x_lo = c.get<uint16_t>((size_t)gg.in_elems());
    EVE_REQUIRE(x_lo, EVE_ERR_WORKSPACE, "conv_fwd(stem): scratch too small");
    stem_prep_weights_kernel<<<cdiv(g.Cout * kStemK, 256), 256, 0, s>>>(
        w, g.Cout, fmt, wscale, w_hi, npass == 3 ? w_lo : nullptr);
    EVE_LAUNCH_CHECK();
    stem_im2col_kernel<<<g.N * g.OH, 256, 7 * g.W * 3 * sizeof(float), s>>>(
        x, g.H, g.W, g.OH, g.OW, fmt, x_hi, npass == 3 ? x_lo : nullptr);
    EVE_LAUNCH_CHECK();
    ProfScope prof(PROF_CONV_FWD, 2.0 * g.out_elems() * (double)g.K(),
                   4.0 * (g.in_elems() + g.out_elems() + (double)wel), s, &g);
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
      if (const PreparedW* pw = find_prepared(w, true, npass)) {
        w_hi = const_cast<uint16_t*>(pw->hi);
        w_lo = const_cast<uint16_t*>(pw->lo);
      } else {
        EVE_TRY(conv_tc_prep_weights(g, w, true, w_hi, npass == 3 ? w_lo : nullptr, TC_BF16, 1.f, s));
      }
      EVE_TRY(split_planes(dy, g.out_elems(), d_hi, npass == 3 ? d_lo : nullptr, TC_BF16, s));
      ProfScope prof(PROF_CONV_DGRAD, 2.0 * g.out_elems() * (double)g.K(),
                     4.0 * (g.in_elems() + g.out_elems() + (double)wel), s, &g);
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
    if (const PreparedW* pw = find_prepared(w, true, npass)) {
      w_hi = const_cast<uint16_t*>(pw->hi);
      w_lo = const_cast<uint16_t*>(pw->lo);
    } else {
      EVE_TRY(conv_tc_prep_weights(g, w, true, w_hi, npass == 3 ? w_lo : nullptr, TC_BF16, 1.f, s));
    }
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
                   4.0 * (g.in_elems() + g.out_elems() + (double)wel), s, &g);
    return conv_tc_dgrad_s2_run(g, d_hi, d_lo, w_hi, w_lo, addend, dx, npass, s);
  }
  float* wd = c.get<float>(wel);
  EVE_REQUIRE(wd, EVE_ERR_WORKSPACE, "conv_dgrad: scratch too small");
  EVE_TRY(conv_prep_weights(g, w, nullptr, wd, s));
  return conv_dgrad_simt(g, dy, g.Cout, wd, addend, dx, s);
}

// ---- the stem's weight gradient on the tensor cores: bf16 hi/lo planes of dy (split here, or
// written directly by the producer of dy: stem_pool_in_backward) x the bf16 patch matrix of x
struct StemWgradCarve {
  float* part;
  uint16_t *d_hi, *d_lo, *x_hi, *x_lo;
};

bool conv_wgrad_stem_takes_planes(const ConvGeom& g) {
  return conv_mode() != 0 && (tc_mask() & 4) && !conv_tc_wgrad_supported(g) && is_stem(g) &&
         conv_tc_wgrad_supported(stem_gemm(g));
}

static bool carve_stem_wgrad(const ConvGeom& g, const ConvScratch& sc, StemWgradCarve& k) {
  Carve c{sc.base, sc.base + sc.bytes};
  const int npass = conv_mode() == 1 ? 3 : 1;
  const ConvGeom gg = stem_gemm(g);
  size_t pf = conv_tc_wgrad_partial_floats(gg);
  size_t cs = colsum_scratch_floats((long long)g.N * g.OH * g.OW, g.Cout);
  k.part = c.get<float>(pf > cs ? pf : cs);
  k.d_hi = c.get<uint16_t>((size_t)g.out_elems());
  k.d_lo = c.get<uint16_t>((size_t)g.out_elems());
  k.x_hi = c.get<uint16_t>((size_t)gg.in_elems());
  k.x_lo = c.get<uint16_t>((size_t)gg.in_elems());
  if (!k.x_lo) return false;
  if (npass != 3) k.d_lo = k.x_lo = nullptr;
  return true;
}

int conv_wgrad_stem_planes(const ConvGeom& g, const ConvScratch& sc, uint16_t** d_hi,
                           uint16_t** d_lo) {
  StemWgradCarve k;
  EVE_REQUIRE(conv_wgrad_stem_takes_planes(g) && carve_stem_wgrad(g, sc, k), EVE_ERR_WORKSPACE,
              "conv_wgrad_stem_planes: not the tensor-core stem path, or scratch too small");
  *d_hi = k.d_hi;
  *d_lo = k.d_lo;
  return EVE_OK;
}

int conv_wgrad_stem_run(const ConvGeom& g, const float* x, float* dw, bool accumulate,
                        const ConvScratch& sc, cudaStream_t s) {
  StemWgradCarve k;
  EVE_REQUIRE(conv_wgrad_stem_takes_planes(g) && carve_stem_wgrad(g, sc, k), EVE_ERR_WORKSPACE,
              "conv_wgrad(stem): scratch too small");
  const int npass = conv_mode() == 1 ? 3 : 1;
  const ConvGeom gg = stem_gemm(g);
  stem_im2col_kernel<<<g.N * g.OH, 256, 7 * g.W * 3 * sizeof(float), s>>>(
      x, g.H, g.W, g.OH, g.OW, TC_BF16, k.x_hi, k.x_lo);
  EVE_LAUNCH_CHECK();
  int splits = 0;
  ProfScope prof(PROF_CONV_WGRAD, 2.0 * g.out_elems() * (double)g.K(),
                 4.0 * (g.in_elems() + g.out_elems() + (double)g.Cout * g.K()), s, &g);
  EVE_TRY(conv_tc_wgrad_run(gg, k.d_hi, k.d_lo, k.x_hi, k.x_lo, k.part, npass, &splits, s));
  stem_wgrad_reduce_kernel<<<cdiv(g.Cout * 147, 256), 256, 0, s>>>(k.part, splits, g.Cout, dw,
                                                                  accumulate ? 1 : 0);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
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
                     4.0 * (g.in_elems() + g.out_elems() + (double)g.Cout * g.K()), s, &g);
      EVE_TRY(conv_tc_wgrad_run(g, d_hi, d_lo, x_hi, x_lo, part, npass, &splits, s, xfmt));
      EVE_TRY(wgrad_reduce(part, splits, g, dw, accumulate, s));
    }
    if (dbias)
      EVE_TRY(colsum(dy, (long long)g.N * g.OH * g.OW, g.Cout, g.Cout, dbias, part, accumulate, s));
    return EVE_OK;
  }
  if (dw && conv_wgrad_stem_takes_planes(g)) {
    StemWgradCarve k;
    EVE_REQUIRE(carve_stem_wgrad(g, sc, k), EVE_ERR_WORKSPACE, "conv_wgrad(stem): scratch too small");
    EVE_TRY(split_planes(dy, g.out_elems(), k.d_hi, k.d_lo, TC_BF16, s));
    EVE_TRY(conv_wgrad_stem_run(g, x, dw, accumulate, sc, s));
    if (dbias)
      EVE_TRY(colsum(dy, (long long)g.N * g.OH * g.OW, g.Cout, g.Cout, dbias, k.part, accumulate, s));
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
  const bool s1 = bwd_dgrad_s1(g);
  const bool s2 = conv_tc_dgrad_s2_supported(g);
  const bool fused = dw && mode != 0 && (tc_mask() & 6) == 6 && conv_tc_wgrad_supported(g) &&
                     (s1 || s2);
  if (!x && !dw) {   // only the data gradient is wanted: x is not needed at all
    if (dbias)
      EVE_TRY(conv_wgrad(g, nullptr, dy, nullptr, dbias, accumulate, sc, s));
    if (dx) EVE_TRY(conv_dgrad(g, dy, w, addend, dx, sc, s));
    return EVE_OK;
  }
  EVE_REQUIRE(x || (fused && conv_x_fusable(g)), EVE_ERR_NULL,
              "conv_bwd: x is NULL but the convolution does not take pre-split operand planes");
  if (!fused) {
    if (dw || dbias) EVE_TRY(conv_wgrad(g, x, dy, dw, dbias, accumulate, sc, s));
    if (dx) EVE_TRY(conv_dgrad(g, dy, w, addend, dx, sc, s));
    return EVE_OK;
  }
  const int npass = mode == 1 ? 3 : 1;
  const size_t wel = (size_t)g.Cout * g.K();
  BwdCarve b;
  EVE_REQUIRE(carve_bwd(g, sc, b), EVE_ERR_WORKSPACE, "conv_bwd: scratch too small");
  float* part = b.part;
  // tcgen05 kind::f16 needs both operands in ONE 16-bit format (an fp16 x bf16 instruction
  // descriptor is an illegal instruction on sm_100a), so x is split as bf16 next to dy
  EVE_TRY(split_planes(dy, g.out_elems(), b.d_hi, npass == 3 ? b.d_lo : nullptr, TC_BF16, s));
  if (x) EVE_TRY(split_planes(x, g.in_elems(), b.x_hi, npass == 3 ? b.x_lo : nullptr, TC_BF16, s));
  if (dx) {
    if (const PreparedW* pw = find_prepared(w, true, npass)) {
      b.w_hi = const_cast<uint16_t*>(pw->hi);
      b.w_lo = const_cast<uint16_t*>(pw->lo);
    } else {
      EVE_TRY(conv_tc_prep_weights(g, w, true, b.w_hi, npass == 3 ? b.w_lo : nullptr, TC_BF16, 1.f, s));
    }
  }
  const double flops = 2.0 * g.out_elems() * (double)g.K();
  const double bytes = 4.0 * (g.in_elems() + g.out_elems() + (double)wel);
  {
    int splits = 0;
    ProfScope prof(PROF_CONV_WGRAD, flops, bytes, s, &g);
    EVE_TRY(conv_tc_wgrad_run(g, b.d_hi, b.d_lo, b.x_hi, b.x_lo, part, npass, &splits, s));
    EVE_TRY(wgrad_reduce(part, splits, g, dw, accumulate, s));
  }
  if (dbias)
    EVE_TRY(colsum(dy, (long long)g.N * g.OH * g.OW, g.Cout, g.Cout, dbias, part, accumulate, s));
  if (!dx) return EVE_OK;
  if (!s1 && g.KH == 1) {
    if (addend) {
      if (addend != dx)
        EVE_CUDA(cudaMemcpyAsync(dx, addend, (size_t)g.in_elems() * sizeof(float),
                                 cudaMemcpyDeviceToDevice, s));
    } else {
      EVE_TRY(fill_zero(dx, g.in_elems(), s));
    }
  }
  ProfScope prof(PROF_CONV_DGRAD, flops, bytes, s, &g);
  if (s1)
    return conv_tc_run(dgrad_as_fwd(g), b.d_hi, b.d_lo, b.w_hi, b.w_lo, nullptr, addend, dx, npass,
                       TC_BF16, 1.f, s);
  return conv_tc_dgrad_s2_run(g, b.d_hi, b.d_lo, b.w_hi, b.w_lo, addend, dx, npass, s);
}

// ---- plane-to-plane entry points: only the weight planes (and the split-K partials) come out of
// the scratch slice; the activations' planes belong to the caller
int conv_fwd_planes(const ConvGeom& g, const void* x_hi, const void* x_lo, const float* w,
                    const float* bias, const float* addend, float* y, const ConvScratch& sc,
                    cudaStream_t s) {
  EVE_REQUIRE(conv_x_fusable(g), EVE_ERR_CONFIG,
              "conv_fwd_planes: the convolution does not take the split tensor-core path");
  EVE_REQUIRE(x_hi && x_lo && w && y, EVE_ERR_NULL, "conv_fwd_planes: NULL pointer");
  Carve c{sc.base, sc.base + sc.bytes};
  const size_t wel = (size_t)g.Cout * g.K();
  uint16_t* w_hi = c.get<uint16_t>(wel);
  uint16_t* w_lo = c.get<uint16_t>(wel);
  EVE_REQUIRE(w_lo, EVE_ERR_WORKSPACE, "conv_fwd_planes: scratch too small");
  const float wscale = 64.f;
  if (const PreparedW* pw = find_prepared(w, false, 3)) {
    w_hi = const_cast<uint16_t*>(pw->hi);
    w_lo = const_cast<uint16_t*>(pw->lo);
  } else {
    EVE_TRY(conv_tc_prep_weights(g, w, false, w_hi, w_lo, TC_F16, wscale, s));
  }
  ProfScope prof(PROF_CONV_FWD, 2.0 * g.out_elems() * (double)g.K(),
                 4.0 * (g.in_elems() + g.out_elems() + (double)wel), s, &g);
  return conv_tc_run(g, x_hi, x_lo, w_hi, w_lo, bias, addend, y, 3, TC_F16, 1.f / wscale, s);
}

int conv_bwd_planes(const ConvGeom& g, const void* x_hi, const void* x_lo, const void* d_hi,
                    const void* d_lo, const float* w, float* dw, bool accumulate,
                    const float* addend, float* dx, const ConvScratch& sc, cudaStream_t s) {
  EVE_REQUIRE(conv_x_fusable(g), EVE_ERR_CONFIG,
              "conv_bwd_planes: the convolution does not take the split tensor-core path");
  EVE_REQUIRE(d_hi && d_lo && (!dw || (x_hi && x_lo)) && (!dx || w), EVE_ERR_NULL,
              "conv_bwd_planes: NULL pointer");
  Carve c{sc.base, sc.base + sc.bytes};
  const size_t wel = (size_t)g.Cout * g.K();
  float* part = c.get<float>(conv_tc_wgrad_partial_floats(g));
  uint16_t* w_hi = c.get<uint16_t>(wel);
  uint16_t* w_lo = c.get<uint16_t>(wel);
  EVE_REQUIRE(w_lo, EVE_ERR_WORKSPACE, "conv_bwd_planes: scratch too small");
  const double flops = 2.0 * g.out_elems() * (double)g.K();
  const double bytes = 4.0 * (g.in_elems() + g.out_elems() + (double)wel);
  if (dw) {
    int splits = 0;
    ProfScope prof(PROF_CONV_WGRAD, flops, bytes, s, &g);
    EVE_TRY(conv_tc_wgrad_run(g, d_hi, d_lo, x_hi, x_lo, part, 3, &splits, s));
    EVE_TRY(wgrad_reduce(part, splits, g, dw, accumulate, s));
  }
  if (!dx) return EVE_OK;
  const bool s1 = bwd_dgrad_s1(g);
  if (const PreparedW* pw = find_prepared(w, true, 3)) {      // same layout for stride 1 and 2
    w_hi = const_cast<uint16_t*>(pw->hi);
    w_lo = const_cast<uint16_t*>(pw->lo);
  } else {
    EVE_TRY(conv_tc_prep_weights(g, w, true, w_hi, w_lo, TC_BF16, 1.f, s));
  }
  if (!s1 && g.KH == 1) {
    // stride-2 1x1: only even/even pixels receive data, the rest is the addend (or zero)
    if (addend) {
      if (addend != dx)
        EVE_CUDA(cudaMemcpyAsync(dx, addend, (size_t)g.in_elems() * sizeof(float),
                                 cudaMemcpyDeviceToDevice, s));
    } else {
      EVE_TRY(fill_zero(dx, g.in_elems(), s));
    }
  }
  ProfScope prof(PROF_CONV_DGRAD, flops, bytes, s, &g);
  if (s1)
    return conv_tc_run(dgrad_as_fwd(g), d_hi, d_lo, w_hi, w_lo, nullptr, addend, dx, 3, TC_BF16, 1.f,
                       s);
  return conv_tc_dgrad_s2_run(g, d_hi, d_lo, w_hi, w_lo, addend, dx, 3, s);
}

void conv_prepared_clear() { g_nprepared = 0; }
int conv_prepared_mark() { return g_nprepared; }
void conv_prepared_truncate(int mark) {
  if (mark >= 0 && mark < g_nprepared) g_nprepared = mark;
}

// Prepare the weights of `g` for repeated conv_fwd (dgrad = false) or stride-1 conv_dgrad
// (dgrad = true) calls.  Storage is taken from the top of the scratch slice, below `*top_used`
// bytes already handed out there; nothing is registered (and the per-call preparation stays in
// place) when the convolution would not take the split tensor-core path or when the region could
// collide with the operand planes conv_fwd / conv_dgrad carve from the bottom.
int conv_prepare_weights(const ConvGeom& g, const float* w, bool dgrad, const ConvScratch& sc,
                         size_t* top_used, cudaStream_t s) {
  const int mode = conv_mode();
  if (mode == 0 || g_nprepared >= kMaxPrepared) return EVE_OK;
  const int npass = mode == 1 ? 3 : 1;
  if (!dgrad && !((tc_mask() & 1) && conv_tc_supported(g))) return EVE_OK;
  if (dgrad && !((tc_mask() & 2) && g.stride == 1 && conv_tc_supported(dgrad_as_fwd(g)))) return EVE_OK;
  const size_t wel = (size_t)g.Cout * g.K();
  const size_t plane = align_up(wel * sizeof(uint16_t), 1024);
  // bottom-up carve of the consumer: two weight planes + two operand planes
  const size_t opnd = (size_t)(dgrad ? g.out_elems() : g.in_elems());
  const size_t bottom = 2 * plane + 2 * align_up(opnd * sizeof(uint16_t), 1024);
  if (bottom + *top_used + 2 * plane + 1024 > sc.bytes) return EVE_OK;
  char* top = sc.base + ((sc.bytes - *top_used - 2 * plane) & ~(size_t)1023);
  *top_used = (size_t)(sc.base + sc.bytes - top);
  uint16_t* hi = (uint16_t*)top;
  uint16_t* lo = (uint16_t*)(top + plane);
  const int fmt = (!dgrad && npass == 3) ? TC_F16 : TC_BF16;
  const float wscale = (!dgrad && npass == 3) ? 64.f : 1.f;
  EVE_TRY(conv_tc_prep_weights(g, w, dgrad, hi, npass == 3 ? lo : nullptr, fmt, wscale, s));
  g_prepared[g_nprepared++] = PreparedW{w, dgrad ? 1 : 0, npass, hi, lo};
  return EVE_OK;
}

// ---- batched weight preparation: the tensor-core layouts of ALL convolutions of a network pass
// from ONE launch (they only change at the optimizer step; preparing them per call cost ~125 tiny
// launches per training step).  Entries are registered like conv_prepare_weights() does.
namespace {
struct PrepEntry {
  const float* w;
  uint16_t *hi, *lo;
  int Cout, Cin, KH, KW;
  int flip, fmt;
  float scale;
  long long start;             // first block of this entry in the launch
};
constexpr int kPrepBatch = 56;
struct PrepTable {
  int n;
  PrepEntry e[kPrepBatch];
};

// One block = one tile of one entry, staged through shared memory so that both the OIHW reads and
// the 16-bit writes are contiguous runs (one thread per output element read the weights with a stride
// of `taps` floats -- or of a whole filter for the flipped layout -- and spent ~100 instructions on
// 64-bit index arithmetic: 345 us per training step for 200 MB of traffic):
//   forward  [co][tap][ci]        : one output channel x up to 512 input channels (its taps * Cin
//                                   floats are one contiguous run of the source)
//   flipped  [ci][flipped tap][co]: 64 output x 8 input channels (8 * taps contiguous floats per
//                                   output channel in, 64 contiguous values per (ci, tap) out)
constexpr int kPrepCi = 512, kPrepFlipCo = 64, kPrepFlipCi = 8;
constexpr int kPrepPitch = kPrepFlipCi * 9 + 1;        // odd pitch: conflict-free transposed reads

__host__ __device__ inline int prep_entry_blocks(int Cout, int Cin, int flip) {
  return flip ? ((Cout + kPrepFlipCo - 1) / kPrepFlipCo) * ((Cin + kPrepFlipCi - 1) / kPrepFlipCi)
              : Cout * ((Cin + kPrepCi - 1) / kPrepCi);
}

__global__ void __launch_bounds__(256)
prep_weights_batch_kernel(const PrepTable tab) {
  __shared__ float tile[kPrepFlipCo * kPrepPitch > kPrepCi * 9 ? kPrepFlipCo * kPrepPitch : kPrepCi * 9];
  int k = 0;                               // block-uniform search
  while (k + 1 < tab.n && tab.e[k + 1].start <= (long long)blockIdx.x) ++k;
  const PrepEntry& e = tab.e[k];
  const int b = (int)((long long)blockIdx.x - e.start);
  const int taps = e.KH * e.KW;
  const float scale = e.scale;
  if (!e.flip) {
    const int chunks = (e.Cin + kPrepCi - 1) / kPrepCi;
    const int co = b / chunks, ci0 = (b - co * chunks) * kPrepCi;
    const int nci = min(kPrepCi, e.Cin - ci0);
    const int n = nci * taps;
    const float* src = e.w + ((size_t)co * e.Cin + ci0) * taps;
    for (int i = threadIdx.x; i < n; i += 256) tile[i] = __ldg(src + i) * scale;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) {
      const int tap = i / nci, cl = i - tap * nci;
      uint16_t h, l;
      split16(tile[cl * taps + tap], e.fmt, h, l);
      const size_t o = ((size_t)co * taps + tap) * e.Cin + ci0 + cl;
      e.hi[o] = h;
      if (e.lo) e.lo[o] = l;
    }
  } else {
    const int cib = (e.Cin + kPrepFlipCi - 1) / kPrepFlipCi;
    const int cot = b / cib, ci0 = (b - cot * cib) * kPrepFlipCi, co0 = cot * kPrepFlipCo;
    const int nco = min(kPrepFlipCo, e.Cout - co0), nci = min(kPrepFlipCi, e.Cin - ci0);
    const int run = nci * taps;            // contiguous source floats per output channel
    for (int i = threadIdx.x; i < nco * run; i += 256) {
      const int cl = i / run, j = i - cl * run;
      tile[cl * kPrepPitch + j] = __ldg(e.w + ((size_t)(co0 + cl) * e.Cin + ci0) * taps + j) * scale;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nco * run; i += 256) {
      const int j = i / nco, cl = i - j * nco;          // j = (local ci, flipped tap)
      const int cil = j / taps, ftap = j - cil * taps;
      uint16_t h, l;
      split16(tile[cl * kPrepPitch + cil * taps + (taps - 1 - ftap)], e.fmt, h, l);
      const size_t o = ((size_t)(ci0 + cil) * taps + ftap) * e.Cout + co0 + cl;
      e.hi[o] = h;
      if (e.lo) e.lo[o] = l;
    }
  }
}
}  // namespace

static bool prep_wanted(const ConvGeom& g, bool dgrad) {
  const int mode = conv_mode();
  if (mode == 0) return false;
  if (!dgrad) return (tc_mask() & 1) && conv_tc_supported(g);
  if (!(tc_mask() & 2)) return false;
  return bwd_dgrad_s1(g) || conv_tc_dgrad_s2_supported(g);
}

size_t conv_prepare_batch_bytes(const ConvPrepReq* reqs, int n) {
  size_t bytes = 0;
  for (int i = 0; i < n; ++i)
    bytes += 2 * align_up((size_t)reqs[i].g.Cout * reqs[i].g.K() * sizeof(uint16_t), 1024);
  return bytes + 1024;
}

int conv_prepare_batch(const ConvPrepReq* reqs, int n, bool dgrad, void* region, size_t bytes,
                       cudaStream_t s) {
  const int mode = conv_mode();
  if (mode == 0 || !region) return EVE_OK;
  const int npass = mode == 1 ? 3 : 1;
  char* p = (char*)align_up((size_t)region, 1024);
  char* end = (char*)region + bytes;
  PrepTable tab;
  tab.n = 0;
  long long total = 0;
  auto flush = [&]() -> int {
    if (tab.n == 0) return EVE_OK;
    prep_weights_batch_kernel<<<(unsigned)total, 256, 0, s>>>(tab);
    EVE_LAUNCH_CHECK();
    tab.n = 0;
    total = 0;
    return EVE_OK;
  };
  for (int i = 0; i < n; ++i) {
    const ConvGeom& g = reqs[i].g;
    if (!reqs[i].w || !prep_wanted(g, dgrad) || g_nprepared >= kMaxPrepared) continue;
    if (g.KH * g.KW > 9) continue;                                   // the kernel's tiles hold up to nine taps
    if (find_prepared(reqs[i].w, dgrad, npass)) continue;          // shared weights
    const size_t wel = (size_t)g.Cout * g.K();
    const size_t plane = align_up(wel * sizeof(uint16_t), 1024);
    if (p + 2 * plane > end) break;                                  // the rest is prepared per call
    PrepEntry e;
    e.w = reqs[i].w;
    e.hi = (uint16_t*)p;
    e.lo = npass == 3 ? (uint16_t*)(p + plane) : nullptr;
    p += 2 * plane;
    e.Cout = g.Cout; e.Cin = g.Cin; e.KH = g.KH; e.KW = g.KW;
    e.flip = dgrad ? 1 : 0;
    e.fmt = (!dgrad && npass == 3) ? TC_F16 : TC_BF16;
    e.scale = (!dgrad && npass == 3) ? 64.f : 1.f;
    e.start = total;
    total += prep_entry_blocks(g.Cout, g.Cin, e.flip);
    tab.e[tab.n++] = e;
    g_prepared[g_nprepared++] = PreparedW{e.w, dgrad ? 1 : 0, npass, e.hi, e.lo ? e.lo : e.hi};
    if (tab.n == kPrepBatch) EVE_TRY(flush());
  }
  return flush();
}

// act(IN(x)) feeding convolution `g`: either as fp32 `y` (then conv_* split it themselves) or,
// when the convolution takes pre-split operands, straight into its 16-bit operand planes inside
// the conv scratch (returns with *fused = true; the caller then passes x = nullptr to conv_*).
int norm_act_into_conv(const ConvGeom& g, bool backward, const float* x, int N, int HW, int C,
                              const float* mean, const float* rstd, const float* gamma,
                              const float* beta, int act, float* y, const ConvScratch& cs,
                              bool* fused, cudaStream_t s) {
  *fused = conv_x_fusable(g);
  if (*fused) {
    void *hi = nullptr, *lo = nullptr;
    if (backward) conv_x_planes_bwd(g, cs, &hi, &lo);
    else conv_x_planes_fwd(g, cs, &hi, &lo);
    EVE_REQUIRE(lo, EVE_ERR_WORKSPACE, "norm_act_into_conv: conv scratch too small");
    return in_apply_planes(x, N, HW, C, mean, rstd, gamma, beta, act, backward ? TC_BF16 : TC_F16,
                           nullptr, hi, lo, s);
  }
  if (backward) return EVE_OK;       // the forward pass kept the fp32 activation in the tape
  return in_apply(x, N, HW, C, mean, rstd, gamma, beta, nullptr, nullptr, nullptr, act, y, s);
}

}  // namespace eve
