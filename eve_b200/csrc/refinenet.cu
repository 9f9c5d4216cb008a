// GazeRefineNet: pre-activation encoder / ConvRNN bottleneck / decoder, forward + backward.
//
// Reference: src/models/refine_net.py:35-255 (BasicBlock :35-67, WrapEncoderDecoder :70-129,
// Bottleneck :132-176, RefineNet :179-255) and the cells in src/models/common.py:331-415.
// Encoder and decoder run for all batch*steps frames at once (per-sample norms, SURVEY.md
// 3.3); only the 5x8 bottleneck cells walk over time, on time-major copies of the features.
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace eve {
namespace {

constexpr int kLevels = 5;
constexpr int kLevelH[kLevels] = {72, 36, 18, 9, 5};
constexpr int kLevelW[kLevels] = {128, 64, 32, 16, 8};
constexpr int kLevelC[kLevels] = {16, 32, 64, 128, 256};  // channels entering each level
constexpr int kMaxRCells = 4;
constexpr int kInitC = 16;   // initial.0 input channels after zero padding

struct RBlock {
  int ic, oc, H, W, act, slot;
  bool skipconv;
  ConvGeom g1, g2, gs;
  const float* x;
  float *m0, *r0, *y1, *s, *c1, *m1, *r1, *y2, *out;
};

struct CellTape {
  // time-major [T][B][P][.]
  float *xh, *cat2, *r, *z, *n, *h, *h0;
};

struct RNet {
  eve_refinenet_params p;
  int N, B, T, nf;
  std::vector<std::string> names;
  int slot_initial, slot_final, slot_rnn;
  ConvGeom gi0, gi3, gf0, gf2;
  float *x0, *i0, *im, *ir, *i1, *i2;
  std::vector<RBlock> enc[kLevels];
  RBlock dec[kLevels];
  float* pooled[kLevels];
  int32_t* pidx[kLevels];
  float* cat[kLevels];    // decoder inputs
  float* inner[kLevels];  // output of the module wrapped by level l (at level l+1 resolution)
  float *bx, *by;         // bottleneck in/out, time-major [T][B][P][nf]
  CellTape cell[kMaxRCells];
  float *f0, *f1, *f2, *sig;
  size_t max_act;
};

void add_block_names(std::vector<std::string>& v, const std::string& p, bool skip) {
  const char* base[] = {"layers.0.weight", "layers.0.bias", "layers.2.weight", "layers.2.bias",
                        "layers.3.weight", "layers.3.bias", "layers.5.weight", "layers.5.bias"};
  for (const char* b : base) v.push_back(p + b);
  if (skip) {
    const char* sk[] = {"skip_layer.0.weight", "skip_layer.0.bias", "skip_layer.2.weight",
                        "skip_layer.2.bias"};
    for (const char* b : sk) v.push_back(p + b);
  }
}

RBlock make_block(int N, int ic, int oc, int H, int W, int act, int slot) {
  RBlock k;
  k.ic = ic; k.oc = oc; k.H = H; k.W = W; k.act = act; k.slot = slot;
  k.skipconv = ic != oc;
  k.g1 = make_conv(N, H, W, ic, oc, 3, 1, 1);
  k.g2 = make_conv(N, H, W, oc, oc, 3, 1, 1);
  k.gs = make_conv(N, H, W, ic, oc, 1, 1, 0);
  k.x = nullptr;
  return k;
}

void alloc_block(RBlock& k, int N, Arena& sv, size_t& max_act) {
  size_t pin = (size_t)N * k.H * k.W * k.ic, pout = (size_t)N * k.H * k.W * k.oc;
  k.m0 = sv.get<float>((size_t)N * k.ic);
  k.r0 = sv.get<float>((size_t)N * k.ic);
  k.y1 = sv.get<float>(pin);
  k.s = k.skipconv ? sv.get<float>(pin) : nullptr;
  k.c1 = sv.get<float>(pout);
  k.m1 = sv.get<float>((size_t)N * k.oc);
  k.r1 = sv.get<float>((size_t)N * k.oc);
  k.y2 = sv.get<float>(pout);
  k.out = sv.get<float>(pout);
  if (pin > max_act) max_act = pin;
  if (pout > max_act) max_act = pout;
}

int check_refine(const eve_refinenet_params* p) {
  EVE_REQUIRE(p, EVE_ERR_NULL, "refinenet: params is NULL");
  EVE_REQUIRE(p->batch >= 0 && p->steps >= 0, EVE_ERR_SHAPE, "refinenet: bad batch/steps");
  EVE_REQUIRE(p->in_channels == 4 || p->in_channels == 1, EVE_ERR_CONFIG,
              "refinenet: in_channels must be 4 (screen + heatmap) or 1, got %d", p->in_channels);
  EVE_REQUIRE(p->nf > 0 && p->nf % 4 == 0 && p->nf <= 512, EVE_ERR_CONFIG,
              "refinenet: refine_net_num_features=%d unsupported", p->nf);
  EVE_REQUIRE(p->rnn_type >= EVE_CRNN_NONE && p->rnn_type <= EVE_CRNN_CGRU, EVE_ERR_CONFIG,
              "Unknown RNN type for RefineNet: %d", p->rnn_type);
  EVE_REQUIRE(p->rnn_type == EVE_CRNN_NONE || (p->rnn_cells >= 1 && p->rnn_cells <= kMaxRCells),
              EVE_ERR_CONFIG, "refinenet: rnn_cells=%d unsupported (1..%d)", p->rnn_cells,
              kMaxRCells);
  return EVE_OK;
}

// Builds names, geometry and (when `sv` is real or dry) the saved-activation layout.
bool build_rnet(const eve_refinenet_params& p, Arena& sv, RNet& n) {
  n.p = p;
  n.B = p.batch; n.T = p.steps; n.N = p.batch * p.steps; n.nf = p.nf;
  const int N = n.N;
  n.max_act = 0;
  auto& nm = n.names;
  nm.clear();
  n.slot_initial = 0;
  for (const char* s : {"initial.0.weight", "initial.0.bias", "initial.1.weight", "initial.1.bias",
                        "initial.3.weight", "initial.3.bias"})
    nm.push_back(s);
  // ---- geometry + names
  std::string prefix = "network.";
  const int nenc[kLevels] = {1, 2, 2, 2, 2};
  for (int l = 0; l < kLevels; ++l) {
    const int c = kLevelC[l];
    const int bic = l + 1 < kLevels ? kLevelC[l + 1] : p.nf;
    const int H = kLevelH[l], W = kLevelW[l];
    n.enc[l].clear();
    for (int j = 0; j < nenc[l]; ++j) {
      int ic = j == 0 ? c : bic;
      n.enc[l].push_back(make_block(N, ic, bic, H, W, ACT_RELU, (int)nm.size()));
      add_block_names(nm, prefix + "encoder_blocks." + std::to_string(j) + ".", ic != bic);
    }
    const int dic = bic + (p.use_skip ? bic : 0);
    n.dec[l] = make_block(N, dic, c, H, W, ACT_LEAKY, (int)nm.size());
    add_block_names(nm, prefix + "decoder_blocks.0.", dic != c);
    prefix += "between_module.";
  }
  n.slot_rnn = (int)nm.size();
  if (p.rnn_type != EVE_CRNN_NONE) {
    for (int i = 0; i < p.rnn_cells; ++i) {
      std::string q = prefix + "rnn_cells." + std::to_string(i) + ".";
      if (p.rnn_type == EVE_CRNN_CRNN) {
        nm.push_back(q + "cell.weight"); nm.push_back(q + "cell.bias");
      } else if (p.rnn_type == EVE_CRNN_CLSTM) {
        nm.push_back(q + "gates.weight"); nm.push_back(q + "gates.bias");
      } else {
        nm.push_back(q + "gates_1.weight"); nm.push_back(q + "gates_1.bias");
        nm.push_back(q + "gate_2.weight"); nm.push_back(q + "gate_2.bias");
      }
    }
  }
  n.slot_final = (int)nm.size();
  for (const char* s : {"final.0.weight", "final.0.bias", "final.2.weight", "final.2.bias"})
    nm.push_back(s);

  // ---- saved activations
  const int H0 = kLevelH[0], W0 = kLevelW[0];
  const size_t P0 = (size_t)N * H0 * W0;
  n.gi0 = make_conv(N, H0, W0, kInitC, 16, 3, 1, 1);   // input zero-padded to 16 channels
  n.gi3 = make_conv(N, H0, W0, 16, 16, 3, 1, 1);
  n.gf0 = make_conv(N, H0, W0, 16, 16, 3, 1, 1);
  n.gf2 = make_conv(N, H0, W0, 16, 1, 1, 1, 0);
  n.x0 = sv.get<float>(P0 * kInitC);
  n.i0 = sv.get<float>(P0 * 16);
  n.im = sv.get<float>((size_t)N * 16);
  n.ir = sv.get<float>((size_t)N * 16);
  n.i1 = sv.get<float>(P0 * 16);
  n.i2 = sv.get<float>(P0 * 16);
  n.max_act = P0 * 16;
  const float* cur = n.i2;
  for (int l = 0; l < kLevels; ++l) {
    for (auto& k : n.enc[l]) {
      k.x = cur;
      alloc_block(k, N, sv, n.max_act);
      cur = k.out;
    }
    const int bic = n.enc[l].back().oc;
    if (l + 1 < kLevels) {
      size_t pe = (size_t)N * kLevelH[l + 1] * kLevelW[l + 1] * bic;
      n.pooled[l] = sv.get<float>(pe);
      n.pidx[l] = sv.get<int32_t>(pe);
      cur = n.pooled[l];
    } else {
      n.pooled[l] = nullptr;
      n.pidx[l] = nullptr;
    }
  }
  // bottleneck (time-major)
  const int P = kLevelH[4] * kLevelW[4];
  const size_t be = (size_t)N * P * p.nf;
  n.bx = sv.get<float>(be);
  n.by = sv.get<float>(be);
  for (int i = 0; i < kMaxRCells; ++i) n.cell[i] = CellTape{};
  if (p.rnn_type == EVE_CRNN_CGRU || p.rnn_type == EVE_CRNN_CRNN) {
    for (int i = 0; i < p.rnn_cells; ++i) {
      CellTape& c = n.cell[i];
      c.xh = sv.get<float>(be * 2);
      c.h = sv.get<float>(be);
      c.h0 = sv.get<float>((size_t)n.B * P * p.nf);
      if (p.rnn_type == EVE_CRNN_CGRU) {
        c.cat2 = sv.get<float>(be * 2);
        c.r = sv.get<float>(be);
        c.z = sv.get<float>(be);
        c.n = sv.get<float>(be);
      }
    }
  }
  // decoder, innermost first
  for (int l = kLevels - 1; l >= 0; --l) {
    RBlock& k = n.dec[l];
    n.cat[l] = sv.get<float>((size_t)N * k.H * k.W * k.ic);
    k.x = n.cat[l];
    alloc_block(k, N, sv, n.max_act);
  }
  n.f0 = sv.get<float>(P0 * 16);
  n.f1 = nullptr;
  n.sig = sv.get<float>(P0);
  return sv.ok();
}

// ------------------------------------------------------------------ small kernels --
// x0[n,h,w,16] = (screen[n,0..2,h,w], heatmap[n,0,h,w], 0 ...)  or  (heatmap, 0 ...)
__global__ void __launch_bounds__(256)
pack_input_kernel(const float* __restrict__ screen, const float* __restrict__ hm, long long total,
                  int HW, float* __restrict__ x0) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float4 v;
  if (screen) {
    long long n = i / HW;
    int p = (int)(i % HW);
    const float* sp = screen + n * 3 * HW + p;
    v = make_float4(sp[0], sp[HW], sp[2 * HW], hm[i]);
  } else {
    v = make_float4(hm[i], 0.f, 0.f, 0.f);
  }
  float4* o = reinterpret_cast<float4*>(x0 + i * kInitC);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  o[0] = v; o[1] = z; o[2] = z; o[3] = z;
}

// Fused head (refine_net.py:220-224 after final.0): LeakyReLU -> conv1x1 (16 -> 1) -> sigmoid.
// One thread per pixel reads its 16 channels (64 bytes) once.
constexpr int kHeadC = 16;
__global__ void __launch_bounds__(256)
final_head_fwd_kernel(const float* __restrict__ f0, const float* __restrict__ w,
                      const float* __restrict__ b, long long rows, float* __restrict__ out,
                      float* __restrict__ sig) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float4* fp = reinterpret_cast<const float4*>(f0 + i * kHeadC);
  float acc = __ldg(b);
#pragma unroll
  for (int q = 0; q < kHeadC / 4; ++q) {
    float4 v = __ldg(fp + q);
    float4 ww = __ldg(reinterpret_cast<const float4*>(w) + q);
    acc = fmaf(v.x > 0.f ? v.x : 0.01f * v.x, ww.x, acc);
    acc = fmaf(v.y > 0.f ? v.y : 0.01f * v.y, ww.y, acc);
    acc = fmaf(v.z > 0.f ? v.z : 0.01f * v.z, ww.z, acc);
    acc = fmaf(v.w > 0.f ? v.w : 0.01f * v.w, ww.w, acc);
  }
  float sg = 1.f / (1.f + expf(-acc));
  out[i] = sg;
  sig[i] = sg;
}

// backward of the fused head: df0 = dpre * w * leaky'(f0), per-block partial sums of
// dW[c] = sum dpre * leaky(f0)[c] and db = sum dpre  (part[block][17])
__global__ void __launch_bounds__(256)
final_head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ sig,
                      const float* __restrict__ f0, const float* __restrict__ w, long long rows,
                      float* __restrict__ df0, float* __restrict__ part) {
  __shared__ float sm[8][kHeadC + 1];
  float acc[kHeadC + 1];
#pragma unroll
  for (int c = 0; c <= kHeadC; ++c) acc[c] = 0.f;
  float wv[kHeadC];
#pragma unroll
  for (int c = 0; c < kHeadC; ++c) wv[c] = __ldg(w + c);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows;
       i += (long long)gridDim.x * blockDim.x) {
    const float sg = __ldg(sig + i);
    const float dpre = __ldg(dout + i) * sg * (1.f - sg);
    const float4* fp = reinterpret_cast<const float4*>(f0 + i * kHeadC);
    float4* dp = reinterpret_cast<float4*>(df0 + i * kHeadC);
#pragma unroll
    for (int q = 0; q < kHeadC / 4; ++q) {
      float4 v = __ldg(fp + q);
      float x[4] = {v.x, v.y, v.z, v.w}, d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool pos = x[j] > 0.f;
        acc[q * 4 + j] = fmaf(dpre, pos ? x[j] : 0.01f * x[j], acc[q * 4 + j]);
        d[j] = dpre * wv[q * 4 + j] * (pos ? 1.f : 0.01f);
      }
      dp[q] = make_float4(d[0], d[1], d[2], d[3]);
    }
    acc[kHeadC] += dpre;
  }
#pragma unroll
  for (int c = 0; c <= kHeadC; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][c] = v;
  }
  __syncthreads();
  if (threadIdx.x <= kHeadC) {
    float v = 0.f;
    for (int wp = 0; wp < 8; ++wp) v += sm[wp][threadIdx.x];
    part[(size_t)blockIdx.x * (kHeadC + 1) + threadIdx.x] = v;
  }
}

__global__ void final_head_reduce_kernel(const float* __restrict__ part, int nblocks,
                                         float* __restrict__ dw, float* __restrict__ db,
                                         int accumulate) {
  int c = threadIdx.x;
  if (c > kHeadC) return;
  double a = 0.0;
  for (int b = 0; b < nblocks; ++b) a += (double)part[(size_t)b * (kHeadC + 1) + c];
  float* dst = c < kHeadC ? (dw ? dw + c : nullptr) : db;
  if (dst) *dst = accumulate ? *dst + (float)a : (float)a;
}
constexpr int kHeadBlocks = 148 * 4;

// initial.0 runs on the tensor cores with its 4 (or 1) input channels zero-padded to 16:
// w[16][in_c][3][3] <-> wp[16][16][3][3]
__global__ void pad_cin_kernel(const float* __restrict__ w, int cout, int cin, int cpad,
                               float* __restrict__ wp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cpad * 9) return;
  int t = i % 9;
  int ci = (i / 9) % cpad;
  int co = i / (9 * cpad);
  wp[i] = ci < cin ? w[((size_t)co * cin + ci) * 9 + t] : 0.f;
}
__global__ void unpad_cin_kernel(const float* __restrict__ dwp, int cout, int cin, int cpad,
                                 float* __restrict__ dw, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin * 9) return;
  int t = i % 9;
  int ci = (i / 9) % cin;
  int co = i / (9 * cin);
  float v = dwp[((size_t)co * cpad + ci) * 9 + t];
  dw[i] = accumulate ? dw[i] + v : v;
}
// x[A][Bd][P][ldx] (first C channels) -> y[Bd][A][P][ldy] (first C channels)
__global__ void __launch_bounds__(256)
swap_bt_kernel(const float* __restrict__ x, long long total, int A, int Bd, int P, int C, int ldx,
               float* __restrict__ y, int ldy) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long t = i / C;
  int p = (int)(t % P);
  t /= P;
  int b = (int)(t % Bd);
  int a = (int)(t / Bd);
  y[(((size_t)b * A + a) * P + p) * ldy + c] = x[(((size_t)a * Bd + b) * P + p) * ldx + c];
}

// OIHW slice along the input channels: out[co][ci][3][3] = w[co][c0 + ci][3][3]
__global__ void __launch_bounds__(256)
slice_cin_kernel(const float* __restrict__ w, int cout, int cin_total, int c0, int cin,
                 float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin * 9) return;
  int t = i % 9;
  int ci = (i / 9) % cin;
  int co = i / (9 * cin);
  out[i] = w[((size_t)co * cin_total + c0 + ci) * 9 + t];
}

// xh[row, 0:nf] = x[row], xh[row, nf:2nf] = h[row]
__global__ void __launch_bounds__(256)
cat2_kernel(const float* __restrict__ a, const float* __restrict__ b, long long rows, int nf,
            float* __restrict__ y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 2 * nf) return;
  int c = (int)(i % (2 * nf));
  long long r = i / (2 * nf);
  y[i] = c < nf ? a[r * nf + c] : b[r * nf + c - nf];
}

// CGRU middle: g1[row, 2nf] (pre-sigmoid) -> r, z; cat2 = [r*h, x]
__global__ void __launch_bounds__(256)
cgru_mid_kernel(const float* __restrict__ g1, const float* __restrict__ x,
                const float* __restrict__ h, long long rows, int nf, float* __restrict__ r,
                float* __restrict__ z, float* __restrict__ cat2) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * nf) return;
  int c = (int)(i % nf);
  long long row = i / nf;
  float rv = 1.f / (1.f + expf(-g1[row * 2 * nf + c]));
  float zv = 1.f / (1.f + expf(-g1[row * 2 * nf + nf + c]));
  r[i] = rv;
  z[i] = zv;
  cat2[row * 2 * nf + c] = rv * h[i];
  cat2[row * 2 * nf + nf + c] = x[i];
}

// CGRU out: n = tanh(g2); h' = (1-z) n + z h
__global__ void __launch_bounds__(256)
cgru_out_kernel(const float* __restrict__ g2, const float* __restrict__ z,
                const float* __restrict__ h, long long total, float* __restrict__ n,
                float* __restrict__ hn) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float nv = tanhf(g2[i]);
  float zv = z[i];
  n[i] = nv;
  hn[i] = (1.f - zv) * nv + zv * h[i];
}

// CGRU backward, first half: dh' -> dg2pre, dz (kept), direct dh
__global__ void __launch_bounds__(256)
cgru_bwd1_kernel(const float* __restrict__ dhn, const float* __restrict__ dcarry,
                 const float* __restrict__ z, const float* __restrict__ n,
                 const float* __restrict__ h, long long total, float* __restrict__ dg2,
                 float* __restrict__ dzbuf, float* __restrict__ dhdir) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float d = dhn[i] + (dcarry ? dcarry[i] : 0.f);
  float zv = z[i], nv = n[i];
  dg2[i] = d * (1.f - zv) * (1.f - nv * nv);
  dzbuf[i] = d * (h[i] - nv) * zv * (1.f - zv);  // already through the sigmoid
  dhdir[i] = d * zv;
}

// CGRU backward, second half: dcat2 = [d(rh), dx2] -> dg1pre = [dr*r(1-r), dz']; dh += d(rh)*r
__global__ void __launch_bounds__(256)
cgru_bwd2_kernel(const float* __restrict__ dcat2, const float* __restrict__ r,
                 const float* __restrict__ h, const float* __restrict__ dzbuf, long long rows,
                 int nf, float* __restrict__ dg1, float* __restrict__ dhdir) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * nf) return;
  int c = (int)(i % nf);
  long long row = i / nf;
  float drh = dcat2[row * 2 * nf + c];
  float rv = r[i];
  dg1[row * 2 * nf + c] = drh * h[i] * rv * (1.f - rv);
  dg1[row * 2 * nf + nf + c] = dzbuf[i];
  dhdir[i] += drh * rv;
}

// dx = dcat2[:, nf:] + dxh[:, :nf] ; dh_carry = dhdir + dxh[:, nf:]
__global__ void __launch_bounds__(256)
cgru_bwd3_kernel(const float* __restrict__ dcat2, const float* __restrict__ dxh,
                 const float* __restrict__ dhdir, long long rows, int nf, float* __restrict__ dx,
                 float* __restrict__ dcarry) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * nf) return;
  int c = (int)(i % nf);
  long long row = i / nf;
  dx[i] = (dcat2 ? dcat2[row * 2 * nf + nf + c] : 0.f) + dxh[row * 2 * nf + c];
  dcarry[i] = (dhdir ? dhdir[i] : 0.f) + dxh[row * 2 * nf + nf + c];
}

// CRNN backward: dpre = (dh' + carry) * (1 - h'^2)
__global__ void __launch_bounds__(256)
crnn_bwd_kernel(const float* __restrict__ dhn, const float* __restrict__ dcarry,
                const float* __restrict__ hn, long long total, float* __restrict__ dpre) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float d = dhn[i] + (dcarry ? dcarry[i] : 0.f);
  float v = hn[i];
  dpre[i] = d * (1.f - v * v);
}

// CLSTM pointwise: gates[row,4nf] = (in, forget, out, cell) pre-activation (common.py:376)
__global__ void __launch_bounds__(256)
clstm_out_kernel(const float* __restrict__ gates, const float* __restrict__ c, long long rows,
                 int nf, float* __restrict__ hn, float* __restrict__ cn) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * nf) return;
  int ch = (int)(i % nf);
  long long row = i / nf;
  const float* g = gates + row * 4 * nf + ch;
  float ig = 1.f / (1.f + expf(-g[0]));
  float fg = 1.f / (1.f + expf(-g[nf]));
  float og = 1.f / (1.f + expf(-g[2 * nf]));
  float cg = tanhf(g[3 * nf]);
  float cv = fg * c[i] + ig * cg;
  cn[i] = cv;
  hn[i] = og * tanhf(cv);
}

#define LAUNCH1D(kernel, total, ...)                                      \
  do {                                                                    \
    long long tot__ = (total);                                            \
    if (tot__ > 0) {                                                      \
      kernel<<<cdiv(tot__, 256), 256, 0, s>>>(__VA_ARGS__);               \
      EVE_LAUNCH_CHECK();                                                 \
    }                                                                     \
  } while (0)

// ------------------------------------------------------------------ block fwd/bwd --
int block_fwd(const RBlock& k, int N, const float* const* w, const ConvScratch& cs,
              cudaStream_t s) {
  const float* const* bw = w + k.slot;
  const int HW = k.H * k.W;
  bool fused = false;
  EVE_TRY(in_stats(k.x, N, HW, k.ic, k.m0, k.r0, s));
  EVE_TRY(norm_act_into_conv(k.g1, false, k.x, N, HW, k.ic, k.m0, k.r0, bw[0], bw[1], k.act, k.y1,
                             cs, &fused, s));
  EVE_TRY(conv_fwd(k.g1, fused ? nullptr : k.y1, bw[2], bw[3], nullptr, k.c1, cs, s));
  EVE_TRY(in_stats(k.c1, N, HW, k.oc, k.m1, k.r1, s));
  const float* addend = k.x;
  if (k.skipconv) {
    EVE_TRY(norm_act_into_conv(k.gs, false, k.x, N, HW, k.ic, k.m0, k.r0, bw[8], bw[9], k.act, k.s,
                               cs, &fused, s));
    EVE_TRY(conv_fwd(k.gs, fused ? nullptr : k.s, bw[10], bw[11], nullptr, k.out, cs, s));
    addend = k.out;
  }
  EVE_TRY(norm_act_into_conv(k.g2, false, k.c1, N, HW, k.oc, k.m1, k.r1, bw[4], bw[5], k.act, k.y2,
                             cs, &fused, s));
  EVE_TRY(conv_fwd(k.g2, fused ? nullptr : k.y2, bw[6], bw[7], addend, k.out, cs, s));
  return EVE_OK;
}

struct BwdScratch {
  float *inb, *t0, *t1, *t2, *ga, *gb;
  ConvScratch cs;
};

int conv_param_grads(const ConvGeom& g, const float* x, const float* dy, float* dw, float* db,
                     const ConvScratch& cs, bool acc, cudaStream_t s) {
  return conv_wgrad(g, x, dy, dw, db, acc, cs, s);
}

// dout -> dx (grad w.r.t. k.x).  dx must not alias t0..t2.
int block_bwd(const RBlock& k, int N, const float* const* w, float* const* gr, bool acc,
              const float* dout, float* dx, const BwdScratch& sc, cudaStream_t s) {
  const float* const* bw = w + k.slot;
  float* const* bg = gr + k.slot;
  const int HW = k.H * k.W;
  bool fused = false;
  // the convolutions' x operands (y2, y1, s) are re-derived from the saved pre-norm tensors as bf16
  // planes when the forward pass did not keep them (conv_x_fusable)
  EVE_TRY(norm_act_into_conv(k.g2, true, k.c1, N, HW, k.oc, k.m1, k.r1, bw[4], bw[5], k.act, nullptr,
                             sc.cs, &fused, s));
  EVE_TRY(conv_bwd(k.g2, fused ? nullptr : k.y2, dout, bw[6], bg[6], bg[7], acc, nullptr, sc.t0,
                   sc.cs, s));
  // activation masks are recomputed from (x, mean, rstd, gamma, beta) -- the same arithmetic as
  // in_apply -- instead of being read back from the saved activations (4 bytes/element less in
  // both backward passes of every normalisation)
  EVE_TRY(in_backward(sc.t0, nullptr, k.c1, N, HW, k.oc, k.m1, k.r1, bw[4], bw[5], k.act, nullptr,
                      sc.t1, nullptr, bg[4], bg[5], sc.inb, acc, s));
  EVE_TRY(norm_act_into_conv(k.g1, true, k.x, N, HW, k.ic, k.m0, k.r0, bw[0], bw[1], k.act, nullptr,
                             sc.cs, &fused, s));
  EVE_TRY(conv_bwd(k.g1, fused ? nullptr : k.y1, sc.t1, bw[2], bg[2], bg[3], acc, nullptr, sc.t0,
                   sc.cs, s));
  const float* addend = dout;
  if (k.skipconv) {
    EVE_TRY(norm_act_into_conv(k.gs, true, k.x, N, HW, k.ic, k.m0, k.r0, bw[8], bw[9], k.act,
                               nullptr, sc.cs, &fused, s));
    EVE_TRY(conv_bwd(k.gs, fused ? nullptr : k.s, dout, bw[10], bg[10], bg[11], acc, nullptr, sc.t2,
                     sc.cs, s));
    EVE_TRY(in_backward(sc.t2, nullptr, k.x, N, HW, k.ic, k.m0, k.r0, bw[8], bw[9], k.act, nullptr,
                        sc.t1, nullptr, bg[8], bg[9], sc.inb, acc, s));
    addend = sc.t1;
  }
  EVE_TRY(in_backward(sc.t0, nullptr, k.x, N, HW, k.ic, k.m0, k.r0, bw[0], bw[1], k.act, addend, dx,
                      nullptr, bg[0], bg[1], sc.inb, acc, s));
  return EVE_OK;
}

// ------------------------------------------------- plane-to-plane block pipeline (fused_norm) --
// Every normalisation is ONE pass (in_fused.cu) that writes the consuming convolution's operand
// planes; backward normalisations emit the bf16 dy planes and the bias-gradient column sums of the
// convolution in front of them, so no fp32 activation, split or colsum pass remains inside a block.
struct Planes {
  uint16_t *hi, *lo;
};

bool block_fusable(const RBlock& k) {
  if (!conv_x_fusable(k.g1) || !conv_x_fusable(k.g2)) return false;
  if (k.skipconv && !conv_x_fusable(k.gs)) return false;
  const int HW = k.H * k.W;
  return in_fused_supported(HW, k.ic, 2) && in_fused_supported(HW, k.oc, 2);
}

bool rnet_fused(const RNet& n) {
  if (!get_option(OPT_FUSED_NORM)) return false;
  if (!conv_x_fusable(n.gi3) || !in_fused_supported(kLevelH[0] * kLevelW[0], 16, 2)) return false;
  for (int l = 0; l < kLevels; ++l) {
    for (auto& k : n.enc[l])
      if (!block_fusable(k)) return false;
    if (!block_fusable(n.dec[l])) return false;
  }
  return true;
}

int block_fwd_fused(const RBlock& k, int N, const float* const* w, const ConvScratch& cs,
                    const Planes& PA, const Planes& PB, cudaStream_t s) {
  const float* const* bw = w + k.slot;
  const int HW = k.H * k.W;
  const bool sk = k.skipconv;
  // one read of x: statistics, the main path's planes and (skip convolution) the skip path's
  EVE_TRY(in_fwd_fused(k.x, N, HW, k.ic, nullptr, 0, bw[0], bw[1], sk ? bw[8] : nullptr,
                       sk ? bw[9] : nullptr, k.act, TC_F16, k.m0, k.r0, nullptr, nullptr, nullptr,
                       PA.hi, PA.lo, sk ? PB.hi : nullptr, sk ? PB.lo : nullptr, s));
  EVE_TRY(conv_fwd_planes(k.g1, PA.hi, PA.lo, bw[2], bw[3], nullptr, k.c1, cs, s));
  const float* addend = k.x;
  if (sk) {
    EVE_TRY(conv_fwd_planes(k.gs, PB.hi, PB.lo, bw[10], bw[11], nullptr, k.out, cs, s));
    addend = k.out;
  }
  EVE_TRY(in_fwd_fused(k.c1, N, HW, k.oc, nullptr, 0, bw[4], bw[5], nullptr, nullptr, k.act, TC_F16,
                       k.m1, k.r1, nullptr, nullptr, nullptr, PA.hi, PA.lo, nullptr, nullptr, s));
  return conv_fwd_planes(k.g2, PA.hi, PA.lo, bw[6], bw[7], addend, k.out, cs, s);
}

struct FusedBwd {
  Planes DA, DB, D2, XP, XP2;   // dout planes (ping-pong), dy planes of conv1, x planes
  float* col;                   // reduction scratch of in_bwd_fused / split_colsum
};

size_t fused_col_floats(int N) {
  return (size_t)12 * N * 512 + (size_t)2049 * 512;
}

// dout -> dx.  `din`: the planes of dout (already written, together with the bias gradients of
// conv2 / the skip convolution, by whoever produced dout) or null (derived here).  `dnext`: where
// to write the planes of dx (null: fp32 only); nb_a / nb_b: bias gradients fed by dx's column sums.
int block_bwd_fused(const RBlock& k, int N, const float* const* w, float* const* gr, bool acc,
                    const float* dout, const Planes* din, float* dx, const Planes* dnext,
                    float* nb_a, float* nb_b, const BwdScratch& sc, const FusedBwd& f,
                    cudaStream_t s) {
  const float* const* bw = w + k.slot;
  float* const* bg = gr + k.slot;
  const int HW = k.H * k.W;
  const bool sk = k.skipconv;
  Planes D = f.DA;
  if (din) {
    D = *din;
  } else {
    EVE_TRY(split_colsum(dout, (long long)N * HW, k.oc, D.hi, D.lo, bg[7], sk ? bg[11] : nullptr,
                         acc, f.col, s));
  }
  // Data gradients first, weight gradients after the norm's backward kernel: that kernel computes
  // xhat of the saved pre-norm tensor anyway and writes act(xhat * gamma + beta) -- the x operand
  // of the weight gradient -- as bf16 planes next to dx (no separate re-derivation pass).
  EVE_TRY(conv_bwd_planes(k.g2, nullptr, nullptr, D.hi, D.lo, bw[6], nullptr, acc, nullptr, sc.t0,
                          sc.cs, s));
  // norm 1: dy planes of conv1, its bias gradient, the affine gradients, conv2's x planes
  EVE_TRY(in_bwd_fused(sc.t0, nullptr, nullptr, k.c1, N, HW, k.oc, k.m1, k.r1, bw[4], bw[5], nullptr,
                       nullptr, k.act, nullptr, nullptr, f.D2.hi, f.D2.lo, nullptr, bg[4], bg[5],
                       nullptr, nullptr, bg[3], nullptr, acc, f.col, s, f.XP.hi, f.XP.lo));
  EVE_TRY(conv_bwd_planes(k.g2, f.XP.hi, f.XP.lo, D.hi, D.lo, bw[6], bg[6], acc, nullptr, nullptr,
                          sc.cs, s));
  EVE_TRY(conv_bwd_planes(k.g1, nullptr, nullptr, f.D2.hi, f.D2.lo, bw[2], nullptr, acc, nullptr,
                          sc.t0, sc.cs, s));
  if (sk)
    EVE_TRY(conv_bwd_planes(k.gs, nullptr, nullptr, D.hi, D.lo, bw[10], nullptr, acc, nullptr,
                            sc.t2, sc.cs, s));
  // norm 0 (both affine sets of the same statistics when there is a skip convolution); the
  // identity residual passes dout straight through when there is none
  EVE_TRY(in_bwd_fused(sc.t0, sk ? sc.t2 : nullptr, nullptr, k.x, N, HW, k.ic, k.m0, k.r0, bw[0],
                       bw[1], sk ? bw[8] : nullptr, sk ? bw[9] : nullptr, k.act,
                       sk ? nullptr : dout, dx,
                       dnext ? dnext->hi : nullptr, dnext ? dnext->lo : nullptr, nullptr, bg[0],
                       bg[1], sk ? bg[8] : nullptr, sk ? bg[9] : nullptr, nb_a, nb_b, acc, f.col, s,
                       f.XP.hi, f.XP.lo, sk ? f.XP2.hi : nullptr, sk ? f.XP2.lo : nullptr));
  EVE_TRY(conv_bwd_planes(k.g1, f.XP.hi, f.XP.lo, f.D2.hi, f.D2.lo, bw[2], bg[2], acc, nullptr,
                          nullptr, sc.cs, s));
  if (sk)
    EVE_TRY(conv_bwd_planes(k.gs, f.XP2.hi, f.XP2.lo, D.hi, D.lo, bw[10], bg[10], acc, nullptr,
                            nullptr, sc.cs, s));
  return EVE_OK;
}

size_t rnet_conv_scratch_bytes(const RNet& n) {
  size_t mi = 0, mo = 0, mw = 0, mp = 0;
  auto upd = [&](const ConvGeom& g) {
    mi = std::max(mi, conv_operand_elems(g));
    mo = std::max(mo, (size_t)g.out_elems());
    mw = std::max(mw, (size_t)g.Cout * g.K());
    mp = std::max(mp, conv_partial_floats(g));
  };
  upd(n.gi0); upd(n.gi3); upd(n.gf0); upd(n.gf2);
  for (int l = 0; l < kLevels; ++l) {
    for (auto& k : n.enc[l]) { upd(k.g1); upd(k.g2); upd(k.gs); }
    upd(n.dec[l].g1); upd(n.dec[l].g2); upd(n.dec[l].gs);
  }
  upd(make_conv(n.N, kLevelH[4], kLevelW[4], 2 * n.nf, 4 * n.nf, 3, 1, 1));
  return conv_scratch_bytes(mi, mo, mw, mp);
}

// every convolution whose tensor-core weight layout can be prepared ahead of the pass (w == null:
// geometry only, for sizing); initial.0 runs on a zero-padded scratch copy and is prepared per call
std::vector<ConvPrepReq> rnet_prep_list(const RNet& n, const float* const* w) {
  std::vector<ConvPrepReq> v;
  auto add = [&](const ConvGeom& g, int slot) { v.push_back(ConvPrepReq{g, w ? w[slot] : nullptr}); };
  add(n.gi3, 4);
  auto block = [&](const RBlock& k) {
    add(k.g1, k.slot + 2);
    add(k.g2, k.slot + 6);
    if (k.skipconv) add(k.gs, k.slot + 10);
  };
  for (int l = 0; l < kLevels; ++l) {
    for (auto& k : n.enc[l]) block(k);
    block(n.dec[l]);
  }
  add(n.gf0, n.slot_final);
  return v;
}
size_t rnet_prep_bytes(const RNet& n) {
  const std::vector<ConvPrepReq> v = rnet_prep_list(n, nullptr);
  return conv_prepare_batch_bytes(v.data(), (int)v.size());
}

bool build_bwd_scratch(const RNet& n, Arena& ws, BwdScratch& sc) {
  sc.cs.bytes = rnet_conv_scratch_bytes(n);
  sc.cs.base = ws.get<char>(sc.cs.bytes);
  sc.inb = ws.get<float>(in_backward_scratch_floats(n.N, 512));
  sc.t0 = ws.get<float>(n.max_act);
  sc.t1 = ws.get<float>(n.max_act);
  sc.t2 = ws.get<float>(n.max_act);
  sc.ga = ws.get<float>(n.max_act);
  sc.gb = ws.get<float>(n.max_act);
  return ws.ok();
}

Planes get_planes(Arena& ws, size_t elems) {
  Planes p;
  p.hi = ws.get<uint16_t>(elems);
  p.lo = ws.get<uint16_t>(elems);
  return p;
}

bool build_fused_bwd(const RNet& n, Arena& ws, FusedBwd& f) {
  f.DA = get_planes(ws, n.max_act);
  f.DB = get_planes(ws, n.max_act);
  f.D2 = get_planes(ws, n.max_act);
  f.XP = get_planes(ws, n.max_act);
  f.XP2 = get_planes(ws, n.max_act);
  f.col = ws.get<float>(fused_col_floats(n.N));
  return ws.ok();
}

// Backward-only scratch beyond BwdScratch.
struct BwdExtra {
  float* dskip[kLevels];      // grad w.r.t. the encoder output of level l coming from the concat:
  int dskip_ld[kLevels];      // levels 0-3 a channel slice of dcat[l] (read in place), level 4 its own copy
  float* dcat[kLevels];       // levels 0-3: the gradient of the decoder block's concatenated input
  float* cb[6];               // bottleneck temporaries, each B*P*4nf
  float* dcarry[kMaxRCells];  // B*P*nf
  float *dby, *dbx;           // time-major [T][B][P][nf]
  float* dg1all[kMaxRCells];  // [T][B][P][2nf]
  float* dg2all[kMaxRCells];  // [T][B][P][nf]
};

bool build_bwd_extra(const RNet& n, Arena& ws, BwdExtra& e) {
  for (int l = 0; l < kLevels; ++l) {
    const RBlock& k = n.enc[l].back();
    e.dcat[l] = nullptr;
    e.dskip_ld[l] = k.oc;
    if (!n.p.use_skip) {
      e.dskip[l] = nullptr;
    } else if (l < kLevels - 1 && n.dec[l + 1].oc % 4 == 0 && n.dec[l].ic % 4 == 0) {
      // the decoder block writes its input gradient here and it stays until the encoder's backward
      // has added the skip half (channels [inner.oc, ic)) -- no copy out of a recycled buffer
      const RBlock& d = n.dec[l];
      e.dcat[l] = ws.get<float>((size_t)n.N * d.H * d.W * d.ic);
      e.dskip[l] = e.dcat[l] ? e.dcat[l] + n.dec[l + 1].oc : nullptr;
      e.dskip_ld[l] = d.ic;
    } else {
      e.dskip[l] = ws.get<float>((size_t)n.N * k.H * k.W * k.oc);
    }
  }
  const int P = kLevelH[4] * kLevelW[4];
  for (int i = 0; i < 6; ++i) e.cb[i] = ws.get<float>((size_t)n.B * P * 4 * n.nf);
  size_t be = (size_t)n.N * P * n.nf;
  e.dby = ws.get<float>(be);
  e.dbx = ws.get<float>(be);
  const bool rec = n.p.rnn_type == EVE_CRNN_CGRU || n.p.rnn_type == EVE_CRNN_CRNN;
  for (int i = 0; i < kMaxRCells; ++i) {
    bool on = rec && i < n.p.rnn_cells;
    e.dcarry[i] = on ? ws.get<float>((size_t)n.B * P * n.nf) : nullptr;
    e.dg1all[i] = on ? ws.get<float>(be * 2) : nullptr;
    e.dg2all[i] = on && n.p.rnn_type == EVE_CRNN_CGRU ? ws.get<float>(be) : nullptr;
  }
  return ws.ok();
}

size_t rnet_fwd_scratch_bytes(const RNet& n) {
  const int P = kLevelH[4] * kLevelW[4];
  return align_up(rnet_conv_scratch_bytes(n), 256) +
         4 * align_up(n.max_act * sizeof(uint16_t), 256) +      // operand planes A, B (hi + lo)
         4 * align_up((size_t)n.B * P * 4 * n.nf * sizeof(float), 256) +
         2 * align_up((size_t)kMaxRCells * n.B * P * n.nf * sizeof(float), 256) +
         align_up(rnet_prep_bytes(n), 256) + 4096 + 16384;
}

}  // namespace
}  // namespace eve

using namespace eve;

extern "C" int eve_refinenet_num_weights(const eve_refinenet_params* p) {
  if (check_refine(p) != EVE_OK) return -1;
  Arena sv(nullptr, 0);
  RNet n;
  build_rnet(*p, sv, n);
  return (int)n.names.size();
}

extern "C" const char* eve_refinenet_weight_name(const eve_refinenet_params* p, int i) {
  static thread_local std::string name;
  if (check_refine(p) != EVE_OK) return nullptr;
  Arena sv(nullptr, 0);
  RNet n;
  build_rnet(*p, sv, n);
  if (i < 0 || i >= (int)n.names.size()) return nullptr;
  name = n.names[i];
  return name.c_str();
}

extern "C" size_t eve_refinenet_saved_bytes(const eve_refinenet_params* p) {
  if (check_refine(p) != EVE_OK) return 0;
  Arena sv(nullptr, 0);
  RNet n;
  build_rnet(*p, sv, n);
  return sv.off + 256;
}

extern "C" size_t eve_refinenet_workspace_bytes(const eve_refinenet_params* p) {
  if (check_refine(p) != EVE_OK) return 0;
  Arena sv(nullptr, 0);
  RNet n;
  build_rnet(*p, sv, n);
  Arena ws(nullptr, 0);
  BwdScratch sc;
  BwdExtra ex;
  FusedBwd fb;
  build_bwd_scratch(n, ws, sc);
  build_bwd_extra(n, ws, ex);
  build_fused_bwd(n, ws, fb);
  ws.get<char>(rnet_prep_bytes(n));
  size_t f = rnet_fwd_scratch_bytes(n);
  return (ws.off > f ? ws.off : f) + 256;
}

extern "C" int eve_refinenet_fwd(const eve_refinenet_params* p, const float* screen,
                                 const float* heatmap, const float* h0, const float* c0,
                                 const float* const* w, float* out, float* hT, float* cT,
                                 void* saved, size_t saved_bytes, void* workspace,
                                 size_t workspace_bytes, eve_stream_t stream) {
  EVE_TRY(check_refine(p));
  if (p->batch == 0 || p->steps == 0) return EVE_OK;
  EVE_REQUIRE(heatmap && w && out && saved && workspace, EVE_ERR_NULL,
              "refinenet_fwd: NULL pointer");
  EVE_REQUIRE(p->in_channels == 1 || screen, EVE_ERR_NULL, "refinenet_fwd: screen is NULL");
  conv_prepared_clear();     // an earlier call that failed half-way must not leave entries behind
  cudaStream_t s = as_stream(stream);
  Arena sv(saved, saved_bytes);
  RNet n;
  EVE_REQUIRE(build_rnet(*p, sv, n), EVE_ERR_WORKSPACE,
              "refinenet_fwd: saved buffer too small (%zu < %zu)", saved_bytes, sv.off);
  EVE_REQUIRE(workspace_bytes >= rnet_fwd_scratch_bytes(n), EVE_ERR_WORKSPACE,
              "refinenet_fwd: workspace too small");
  Arena ws(workspace, workspace_bytes);
  const int N = n.N, B = n.B, T = n.T, nf = n.nf;
  const int P = kLevelH[4] * kLevelW[4];
  ConvScratch cs;
  cs.bytes = rnet_conv_scratch_bytes(n);
  cs.base = ws.get<char>(cs.bytes);
  float* cb[4];
  for (int i = 0; i < 4; ++i) cb[i] = ws.get<float>((size_t)B * P * 4 * nf);
  float* h0n = ws.get<float>((size_t)kMaxRCells * B * P * nf);
  float* c0n = ws.get<float>((size_t)kMaxRCells * B * P * nf);
  const Planes PA = get_planes(ws, n.max_act), PB = get_planes(ws, n.max_act);
  const bool fusedn = rnet_fused(n);
  const int HW0 = kLevelH[0] * kLevelW[0];

  // ---- input + initial
  LAUNCH1D(pack_input_kernel, (long long)N * HW0, p->in_channels == 4 ? screen : nullptr, heatmap,
           (long long)N * HW0, HW0, n.x0);
  float* w0p = ws.get<float>((size_t)16 * kInitC * 9);
  EVE_REQUIRE(w0p, EVE_ERR_WORKSPACE, "refinenet_fwd: workspace too small");
  {
    // forward tensor-core layouts of all convolution weights: one launch
    const size_t pb = rnet_prep_bytes(n);
    char* region = ws.get<char>(pb);
    EVE_REQUIRE(region, EVE_ERR_WORKSPACE, "refinenet_fwd: workspace too small");
    const std::vector<ConvPrepReq> reqs = rnet_prep_list(n, w);
    EVE_TRY(conv_prepare_batch(reqs.data(), (int)reqs.size(), false, region, pb, s));
  }
  LAUNCH1D(pad_cin_kernel, 16 * kInitC * 9, w[0], 16, p->in_channels, kInitC, w0p);
  EVE_TRY(conv_fwd(n.gi0, n.x0, w0p, w[1], nullptr, n.i0, cs, s));
  if (fusedn) {
    EVE_TRY(in_fwd_fused(n.i0, N, HW0, 16, nullptr, 0, w[2], w[3], nullptr, nullptr, ACT_RELU, TC_F16,
                         n.im, n.ir, nullptr, nullptr, nullptr, PA.hi, PA.lo, nullptr, nullptr, s));
    EVE_TRY(conv_fwd_planes(n.gi3, PA.hi, PA.lo, w[4], w[5], nullptr, n.i2, cs, s));
  } else {
    EVE_TRY(in_stats(n.i0, N, HW0, 16, n.im, n.ir, s));
    bool i1_fused = false;
    EVE_TRY(norm_act_into_conv(n.gi3, false, n.i0, N, HW0, 16, n.im, n.ir, w[2], w[3], ACT_RELU,
                               n.i1, cs, &i1_fused, s));
    EVE_TRY(conv_fwd(n.gi3, i1_fused ? nullptr : n.i1, w[4], w[5], nullptr, n.i2, cs, s));
  }
  auto run_block = [&](const RBlock& k) -> int {
    return fusedn ? block_fwd_fused(k, N, w, cs, PA, PB, s) : block_fwd(k, N, w, cs, s);
  };
  // ---- encoder
  bool skip_in_cat[kLevels] = {false, false, false, false, false};
  for (int l = 0; l < kLevels; ++l) {
    for (auto& k : n.enc[l]) EVE_TRY(run_block(k));
    if (l + 1 < kLevels) {
      const RBlock& k = n.enc[l].back();
      // the skip connection's half of cat[l] (channels [inner.oc, ic)) is written by the same pass
      const bool cp = p->use_skip && n.dec[l + 1].oc % 4 == 0 && n.dec[l].ic % 4 == 0;
      skip_in_cat[l] = cp;
      EVE_TRY(adaptive_maxpool_fwd(k.out, N, k.H, k.W, k.oc, kLevelH[l + 1], kLevelW[l + 1],
                                   n.pooled[l], n.pidx[l], s, cp ? n.cat[l] + n.dec[l + 1].oc : nullptr,
                                   n.dec[l].ic));
    }
  }
  // ---- bottleneck over time (time-major)
  const float* benc = n.enc[4].back().out;  // [B][T][P][nf]
  const long long E = (long long)P * nf;
  const long long rows = (long long)B * P;
  LAUNCH1D(swap_bt_kernel, (long long)N * E, benc, (long long)N * E, B, T, P, nf, nf, n.bx, nf);
  const float* bout = n.bx;  // what the decoder sees
  if (p->rnn_type != EVE_CRNN_NONE) {
    const int nc = p->rnn_cells;
    if (h0) EVE_TRY(nchw_to_nhwc(h0, nc * B, nf, kLevelH[4], kLevelW[4], h0n, s));
    else EVE_TRY(fill_zero(h0n, (long long)nc * B * E, s));
    if (p->rnn_type == EVE_CRNN_CLSTM) {
      if (c0) EVE_TRY(nchw_to_nhwc(c0, nc * B, nf, kLevelH[4], kLevelW[4], c0n, s));
      else EVE_TRY(fill_zero(c0n, (long long)nc * B * E, s));
    }
    if (p->rnn_type == EVE_CRNN_CLSTM) {
      // refine_net.py:168-174: with tuple states the features handed on are NOT updated, so
      // every cell sees the encoder features and the decoder input is unchanged.
      ConvGeom gg = make_conv(B, kLevelH[4], kLevelW[4], 2 * nf, 4 * nf, 3, 1, 1);
      for (int i = 0; i < nc; ++i) {
        const float* const* cw = w + n.slot_rnn + 2 * i;
        float* hcur = h0n + (size_t)i * B * E;
        float* ccur = c0n + (size_t)i * B * E;
        for (int t = 0; t < T; ++t) {
          const float* xt = n.bx + (size_t)t * B * E;
          LAUNCH1D(cat2_kernel, rows * 2 * nf, xt, hcur, rows, nf, cb[0]);
          EVE_TRY(conv_fwd(gg, cb[0], cw[0], cw[1], nullptr, cb[1], cs, s));
          LAUNCH1D(clstm_out_kernel, rows * nf, cb[1], ccur, rows, nf, hcur, ccur);
        }
      }
    } else {
      const bool gru = p->rnn_type == EVE_CRNN_CGRU;
      ConvGeom g1 = make_conv(B, kLevelH[4], kLevelW[4], 2 * nf, gru ? 2 * nf : nf, 3, 1, 1);
      ConvGeom g2 = make_conv(B, kLevelH[4], kLevelW[4], 2 * nf, nf, 3, 1, 1);
      const int wpc = gru ? 4 : 2;
      for (int i = 0; i < nc; ++i)
        EVE_CUDA(cudaMemcpyAsync(n.cell[i].h0, h0n + (size_t)i * B * E, (size_t)B * E * sizeof(float),
                                 cudaMemcpyDeviceToDevice, s));
      const bool persistent = gru && cgru_seq_supported(nf, kLevelH[4], kLevelW[4]) &&
                              (size_t)N * P * 2 * nf * sizeof(float) <= n.max_act * sizeof(uint16_t);
      if (persistent) {
        // ---- persistent ConvGRU (conv_tc.cu: cgru_seq_fwd_kernel), cell by cell: the x halves of
        // both gate convolutions for all T steps at once, then ONE kernel walks the recurrence.
        // Scratch: the block pipelines' operand-plane buffers are idle during the bottleneck.
        float* gx1 = reinterpret_cast<float*>(PA.hi);           // [T*B][P][2nf]
        float* gx2 = reinterpret_cast<float*>(PA.lo);           // [T*B][P][nf]
        const size_t w1s = (size_t)2 * nf * nf * 9, w2s = (size_t)nf * nf * 9;
        float* W1x = reinterpret_cast<float*>(PB.hi);
        float* W1h = W1x + w1s;
        float* W2h = W1h + w1s;
        float* W2x = W2h + w2s;
        uint16_t* p1h = PB.lo;                                   // K-major fp16 planes of W1h, W2h
        uint16_t* p1l = p1h + w1s;
        uint16_t* p2h = p1l + w1s;
        uint16_t* p2l = p2h + w2s;
        EVE_REQUIRE((2 * w1s + 2 * w2s) * sizeof(float) <= n.max_act * sizeof(uint16_t), EVE_ERR_WORKSPACE,
                    "refinenet_fwd: operand-plane scratch too small for the ConvGRU weights");
        const ConvGeom gx1g = make_conv(N, kLevelH[4], kLevelW[4], nf, 2 * nf, 3, 1, 1);
        const ConvGeom gx2g = make_conv(N, kLevelH[4], kLevelW[4], nf, nf, 3, 1, 1);
        for (int i = 0; i < nc; ++i) {
          CellTape& c = n.cell[i];
          const float* const* cw = w + n.slot_rnn + wpc * i;
          const float* xseq = i == 0 ? n.bx : n.cell[i - 1].h;   // [T][B][P][nf], time-major
          LAUNCH1D(slice_cin_kernel, (long long)w1s, cw[0], 2 * nf, 2 * nf, 0, nf, W1x);
          LAUNCH1D(slice_cin_kernel, (long long)w1s, cw[0], 2 * nf, 2 * nf, nf, nf, W1h);
          LAUNCH1D(slice_cin_kernel, (long long)w2s, cw[2], nf, 2 * nf, 0, nf, W2h);
          LAUNCH1D(slice_cin_kernel, (long long)w2s, cw[2], nf, 2 * nf, nf, nf, W2x);
          EVE_TRY(conv_fwd(gx1g, xseq, W1x, cw[1], nullptr, gx1, cs, s));
          EVE_TRY(conv_fwd(gx2g, xseq, W2x, cw[3], nullptr, gx2, cs, s));
          // x halves of the two concatenated inputs the weight gradients read (xh = [x, h], cat2 = [r*h, x])
          EVE_TRY(copy_channels(xseq, (long long)N * P, nf, nf, 0, c.xh, 2 * nf, 0, false, s));
          EVE_TRY(copy_channels(xseq, (long long)N * P, nf, nf, 0, c.cat2, 2 * nf, nf, false, s));
          EVE_TRY(conv_tc_prep_weights(gx1g, W1h, false, p1h, p1l, TC_F16, 64.f, s));
          EVE_TRY(conv_tc_prep_weights(gx2g, W2h, false, p2h, p2l, TC_F16, 64.f, s));
          EVE_TRY(cgru_seq_fwd(B, T, p1h, p1l, p2h, p2l, gx1, gx2, c.h0, c.r, c.z, c.n, c.h, c.xh,
                               c.cat2, 1.f / 64.f, s));
        }
        EVE_CUDA(cudaMemcpyAsync(n.by, n.cell[nc - 1].h, (size_t)N * E * sizeof(float),
                                 cudaMemcpyDeviceToDevice, s));
      }
      // the gate weights do not change over the T steps: lay them out for the tensor cores once
      const int prep_mark = conv_prepared_mark();
      size_t top_used = 0;
      for (int i = 0; i < nc && !persistent; ++i) {
        const float* const* cw = w + n.slot_rnn + wpc * i;
        EVE_TRY(conv_prepare_weights(g1, cw[0], false, cs, &top_used, s));
        if (gru) EVE_TRY(conv_prepare_weights(g2, cw[2], false, cs, &top_used, s));
      }
      for (int t = 0; t < T && !persistent; ++t) {
        const float* xt = n.bx + (size_t)t * B * E;
        for (int i = 0; i < nc; ++i) {
          CellTape& c = n.cell[i];
          const float* const* cw = w + n.slot_rnn + wpc * i;
          const float* hprev = t == 0 ? c.h0 : c.h + (size_t)(t - 1) * B * E;
          float* xh = c.xh + (size_t)t * B * E * 2;
          float* hn = c.h + (size_t)t * B * E;
          LAUNCH1D(cat2_kernel, rows * 2 * nf, xt, hprev, rows, nf, xh);
          if (gru) {
            float* rr = c.r + (size_t)t * B * E;
            float* zz = c.z + (size_t)t * B * E;
            float* nn = c.n + (size_t)t * B * E;
            float* cat2 = c.cat2 + (size_t)t * B * E * 2;
            EVE_TRY(conv_fwd(g1, xh, cw[0], cw[1], nullptr, cb[0], cs, s));
            LAUNCH1D(cgru_mid_kernel, rows * nf, cb[0], xt, hprev, rows, nf, rr, zz, cat2);
            EVE_TRY(conv_fwd(g2, cat2, cw[2], cw[3], nullptr, cb[1], cs, s));
            LAUNCH1D(cgru_out_kernel, rows * nf, cb[1], zz, hprev, rows * nf, nn, hn);
          } else {
            EVE_TRY(conv_fwd(g1, xh, cw[0], cw[1], nullptr, cb[0], cs, s));
            EVE_TRY(ew_fwd(EW_TANH, cb[0], rows * nf, hn, s));
          }
          xt = hn;
        }
        EVE_CUDA(cudaMemcpyAsync(n.by + (size_t)t * B * E, xt, (size_t)B * E * sizeof(float),
                                 cudaMemcpyDeviceToDevice, s));
      }
      conv_prepared_truncate(prep_mark);
      bout = n.by;
      for (int i = 0; i < nc; ++i)
        EVE_CUDA(cudaMemcpyAsync(h0n + (size_t)i * B * E, n.cell[i].h + (size_t)(T - 1) * B * E,
                                 (size_t)B * E * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    if (hT) EVE_TRY(nhwc_to_nchw(h0n, nc * B, nf, kLevelH[4], kLevelW[4], hT, s));
    if (cT && p->rnn_type == EVE_CRNN_CLSTM)
      EVE_TRY(nhwc_to_nchw(c0n, nc * B, nf, kLevelH[4], kLevelW[4], cT, s));
  }
  // ---- decoder, innermost first.  cat[l] = [module output (upsampled), encoder skip]
  {
    const RBlock& k = n.dec[4];
    // time-major bottleneck output back to batch-major, straight into the concat buffer
    LAUNCH1D(swap_bt_kernel, (long long)N * E, bout, (long long)N * E, T, B, P, nf, nf, n.cat[4],
             k.ic);
    if (p->use_skip)
      EVE_TRY(copy_channels(n.enc[4].back().out, (long long)N * P, nf, nf, 0, n.cat[4], k.ic, nf,
                            false, s));
    EVE_TRY(run_block(k));
  }
  for (int l = 3; l >= 0; --l) {
    const RBlock& k = n.dec[l];
    const RBlock& inner = n.dec[l + 1];
    const RBlock& e = n.enc[l].back();
    EVE_TRY(upsample_bilinear_fwd(inner.out, N, inner.H, inner.W, inner.oc, k.H, k.W, n.cat[l],
                                  k.ic, 0, s));
    if (p->use_skip && !skip_in_cat[l])
      EVE_TRY(copy_channels(e.out, (long long)N * k.H * k.W, e.oc, e.oc, 0, n.cat[l], k.ic,
                            inner.oc, false, s));
    EVE_TRY(run_block(k));
  }
  // ---- final: conv3x3 -> LeakyReLU -> conv1x1 -> sigmoid
  const float* const* fw = w + n.slot_final;
  EVE_TRY(conv_fwd(n.gf0, n.dec[0].out, fw[0], fw[1], nullptr, n.f0, cs, s));
  LAUNCH1D(final_head_fwd_kernel, (long long)N * HW0, n.f0, fw[2], fw[3], (long long)N * HW0, out,
           n.sig);
  conv_prepared_clear();
  return EVE_OK;
}


extern "C" int eve_refinenet_bwd(const eve_refinenet_params* p, const float* dout,
                                 const float* dhT, const float* dcT, const float* const* w,
                                 float* dheatmap, float* dh0, float* dc0, float* const* gr,
                                 int accumulate, const void* saved, size_t saved_bytes,
                                 void* workspace, size_t workspace_bytes, eve_stream_t stream) {
  EVE_TRY(check_refine(p));
  if (p->batch == 0 || p->steps == 0) return EVE_OK;
  EVE_REQUIRE(dout && w && gr && saved && workspace, EVE_ERR_NULL, "refinenet_bwd: NULL pointer");
  EVE_REQUIRE(!(p->rnn_type == EVE_CRNN_CLSTM && (dhT || dcT)), EVE_ERR_CONFIG,
              "refinenet_bwd: gradients into CLSTM states are not supported (the CLSTM state "
              "never reaches the heatmap, refine_net.py:168-174)");
  conv_prepared_clear();
  cudaStream_t s = as_stream(stream);
  Arena sv(const_cast<void*>(saved), saved_bytes);
  RNet n;
  EVE_REQUIRE(build_rnet(*p, sv, n), EVE_ERR_WORKSPACE, "refinenet_bwd: saved buffer too small");
  Arena ws(workspace, workspace_bytes);
  BwdScratch sc;
  BwdExtra ex;
  FusedBwd fb;
  bool ok = build_bwd_scratch(n, ws, sc);
  ok = build_bwd_extra(n, ws, ex) && ok;
  ok = build_fused_bwd(n, ws, fb) && ok;
  const size_t prep_bytes = rnet_prep_bytes(n);
  char* prep_region = ws.get<char>(prep_bytes);
  ok = ok && prep_region != nullptr;
  const bool fusedn = rnet_fused(n);
  EVE_REQUIRE(ok, EVE_ERR_WORKSPACE, "refinenet_bwd: workspace too small (%zu < %zu)",
              workspace_bytes, ws.off);
  {
    // flipped data-gradient layouts of all convolution weights: one launch
    const std::vector<ConvPrepReq> reqs = rnet_prep_list(n, w);
    EVE_TRY(conv_prepare_batch(reqs.data(), (int)reqs.size(), true, prep_region, prep_bytes, s));
  }
  const int N = n.N, B = n.B, T = n.T, nf = n.nf;
  const int P = kLevelH[4] * kLevelW[4];
  const int HW0 = kLevelH[0] * kLevelW[0];
  const long long E = (long long)P * nf;
  const long long rows = (long long)B * P;
  const bool acc = accumulate != 0;

  // ---- final
  const float* const* fw = w + n.slot_final;
  float* const* fg = gr + n.slot_final;
  {
    float* hpart = sc.t0;   // [kHeadBlocks][17]
    final_head_bwd_kernel<<<kHeadBlocks, 256, 0, s>>>(dout, n.sig, n.f0, fw[2], (long long)N * HW0,
                                                      sc.t2, hpart);
    EVE_LAUNCH_CHECK();
    final_head_reduce_kernel<<<1, 32, 0, s>>>(hpart, kHeadBlocks, fg[2], fg[3], acc ? 1 : 0);
    EVE_LAUNCH_CHECK();
  }
  float* cur = sc.ga;
  float* other = sc.gb;
  EVE_TRY(conv_bwd(n.gf0, n.dec[0].out, sc.t2, fw[0], fg[0], fg[1], acc, nullptr, cur, sc.cs, s));

  // ---- decoder, outermost first
  for (int l = 0; l < 4; ++l) {
    const RBlock& k = n.dec[l];
    const RBlock& inner = n.dec[l + 1];
    const RBlock& e = n.enc[l].back();
    float* din = ex.dcat[l] ? ex.dcat[l] : other;   // gradient of cat[upsampled inner, skip]
    if (fusedn)
      EVE_TRY(block_bwd_fused(k, N, w, gr, acc, cur, nullptr, din, nullptr, nullptr, nullptr, sc,
                              fb, s));
    else
      EVE_TRY(block_bwd(k, N, w, gr, acc, cur, din, sc, s));
    if (p->use_skip && !ex.dcat[l])
      EVE_TRY(copy_channels(din, (long long)N * k.H * k.W, e.oc, k.ic, inner.oc, ex.dskip[l],
                            e.oc, 0, false, s));
    EVE_TRY(upsample_bilinear_bwd(din, k.ic, 0, N, inner.H, inner.W, inner.oc, k.H, k.W, cur, s));
  }
  {
    const RBlock& k = n.dec[4];
    if (fusedn)
      EVE_TRY(block_bwd_fused(k, N, w, gr, acc, cur, nullptr, other, nullptr, nullptr, nullptr, sc,
                              fb, s));
    else
      EVE_TRY(block_bwd(k, N, w, gr, acc, cur, other, sc, s));
    if (p->use_skip)
      EVE_TRY(copy_channels(other, (long long)N * P, nf, k.ic, nf, ex.dskip[4], nf, 0, false, s));
    // batch-major [B][T][P][ic] -> time-major [T][B][P][nf]
    LAUNCH1D(swap_bt_kernel, (long long)N * E, other, (long long)N * E, B, T, P, nf, k.ic, ex.dby,
             nf);
  }

  // ---- bottleneck BPTT
  const float* dbx = ex.dby;  // pass-through unless a recurrent cell transforms the features
  if (p->rnn_type == EVE_CRNN_CLSTM) {
    if (!acc)
      for (int i = 0; i < p->rnn_cells; ++i) {
        float* const* cg = gr + n.slot_rnn + 2 * i;
        if (cg[0]) EVE_TRY(fill_zero(cg[0], (long long)4 * nf * 2 * nf * 9, s));
        if (cg[1]) EVE_TRY(fill_zero(cg[1], 4 * nf, s));
      }
    if (dh0) EVE_TRY(fill_zero(dh0, (long long)p->rnn_cells * B * E, s));
    if (dc0) EVE_TRY(fill_zero(dc0, (long long)p->rnn_cells * B * E, s));
  } else if (p->rnn_type != EVE_CRNN_NONE) {
    const bool gru = p->rnn_type == EVE_CRNN_CGRU;
    const int nc = p->rnn_cells;
    const int wpc = gru ? 4 : 2;
    ConvGeom g1 = make_conv(B, kLevelH[4], kLevelW[4], 2 * nf, gru ? 2 * nf : nf, 3, 1, 1);
    ConvGeom g2 = make_conv(B, kLevelH[4], kLevelW[4], 2 * nf, nf, 3, 1, 1);
    // initial carries = gradient into the final states
    for (int i = 0; i < nc; ++i) {
      if (dhT)
        EVE_TRY(nchw_to_nhwc(dhT + (size_t)i * B * E, B, nf, kLevelH[4], kLevelW[4], ex.dcarry[i], s));
      else
        EVE_TRY(fill_zero(ex.dcarry[i], (long long)B * E, s));
    }
    float* dcat2 = ex.cb[0];
    float* dzbuf = ex.cb[1];
    float* dhdir = ex.cb[2];
    float* dxh = ex.cb[3];
    float* dxa = ex.cb[4];
    float* dxb = ex.cb[5];
    const bool persistent = gru && cgru_seq_supported(nf, kLevelH[4], kLevelW[4]) &&
                            (size_t)N * E * sizeof(float) <= n.max_act * sizeof(uint16_t);
    if (persistent) {
      // ---- persistent BPTT (conv_tc.cu: cgru_seq_bwd_kernel), top cell first: ONE kernel walks the
      // recurrence and stores the gate-gradient sequences; the x halves of the data gradients (what
      // the cell below / the encoder receives) follow as two convolutions batched over all T steps.
      // Scratch: the fused block pipeline's plane buffers are idle between decoder and encoder.
      const size_t w1s = (size_t)2 * nf * nf * 9, w2s = (size_t)nf * nf * 9;
      EVE_REQUIRE((2 * w1s + 2 * w2s) * sizeof(float) <= n.max_act * sizeof(uint16_t), EVE_ERR_WORKSPACE,
                  "refinenet_bwd: operand-plane scratch too small for the ConvGRU weights");
      float* W1x = reinterpret_cast<float*>(fb.XP.hi);
      float* W1h = W1x + w1s;
      float* W2h = W1h + w1s;
      float* W2x = W2h + w2s;
      uint16_t* d1h = fb.XP.lo;                // data-gradient planes of W1h [nf][9 * 2nf], W2h [nf][9 * nf]
      uint16_t* d1l = d1h + w1s;
      uint16_t* d2h = d1l + w1s;
      uint16_t* d2l = d2h + w2s;
      float* dxa = reinterpret_cast<float*>(fb.XP2.hi);      // [T][B][P][nf] sequences
      float* dxb = reinterpret_cast<float*>(fb.XP2.lo);
      const ConvGeom gx1g = make_conv(N, kLevelH[4], kLevelW[4], nf, 2 * nf, 3, 1, 1);
      const ConvGeom gx2g = make_conv(N, kLevelH[4], kLevelW[4], nf, nf, 3, 1, 1);
      const float* dcur = ex.dby;
      for (int i = nc - 1; i >= 0; --i) {
        const CellTape& c = n.cell[i];
        const float* const* cw = w + n.slot_rnn + wpc * i;
        LAUNCH1D(slice_cin_kernel, (long long)w1s, cw[0], 2 * nf, 2 * nf, 0, nf, W1x);
        LAUNCH1D(slice_cin_kernel, (long long)w1s, cw[0], 2 * nf, 2 * nf, nf, nf, W1h);
        LAUNCH1D(slice_cin_kernel, (long long)w2s, cw[2], nf, 2 * nf, 0, nf, W2h);
        LAUNCH1D(slice_cin_kernel, (long long)w2s, cw[2], nf, 2 * nf, nf, nf, W2x);
        EVE_TRY(conv_tc_prep_weights(gx1g, W1h, true, d1h, d1l, TC_BF16, 1.f, s));
        EVE_TRY(conv_tc_prep_weights(gx2g, W2h, true, d2h, d2l, TC_BF16, 1.f, s));
        EVE_TRY(cgru_seq_bwd(B, T, d2h, d2l, d1h, d1l, dcur, ex.dcarry[i], c.r, c.z, c.n, c.h, c.h0,
                             ex.dg1all[i], ex.dg2all[i], s));
        float* dxo = (i == 0) ? ex.dbx : (dcur == dxa ? dxb : dxa);
        float* tmp = (dxo == dxa || dcur == dxa) ? dxb : dxa;
        if (tmp == dxo || tmp == dcur) tmp = ex.dbx;          // (i > 0 with both sequences busy)
        EVE_TRY(conv_dgrad(gx1g, ex.dg1all[i], W1x, nullptr, tmp, sc.cs, s));
        EVE_TRY(conv_dgrad(gx2g, ex.dg2all[i], W2x, tmp, dxo, sc.cs, s));
        dcur = dxo;
      }
    }
    const int prep_mark = conv_prepared_mark();
    size_t top_used = 0;
    for (int i = 0; i < nc && !persistent; ++i) {
      const float* const* cw = w + n.slot_rnn + wpc * i;
      EVE_TRY(conv_prepare_weights(g1, cw[0], true, sc.cs, &top_used, s));
      if (gru) EVE_TRY(conv_prepare_weights(g2, cw[2], true, sc.cs, &top_used, s));
    }
    for (int t = T - 1; t >= 0 && !persistent; --t) {
      const float* dcur = ex.dby + (size_t)t * B * E;
      for (int i = nc - 1; i >= 0; --i) {
        const CellTape& c = n.cell[i];
        const float* const* cw = w + n.slot_rnn + wpc * i;
        const float* hprev = t == 0 ? c.h0 : c.h + (size_t)(t - 1) * B * E;
        float* dg1 = ex.dg1all[i] + (size_t)t * B * E * (gru ? 2 : 1);
        float* dxo = (i == 0) ? ex.dbx + (size_t)t * B * E : (dcur == dxa ? dxb : dxa);
        if (gru) {
          float* dg2 = ex.dg2all[i] + (size_t)t * B * E;
          LAUNCH1D(cgru_bwd1_kernel, rows * nf, dcur, ex.dcarry[i], c.z + (size_t)t * B * E,
                   c.n + (size_t)t * B * E, hprev, rows * nf, dg2, dzbuf, dhdir);
          EVE_TRY(conv_dgrad(g2, dg2, cw[2], nullptr, dcat2, sc.cs, s));
          LAUNCH1D(cgru_bwd2_kernel, rows * nf, dcat2, c.r + (size_t)t * B * E, hprev, dzbuf, rows,
                   nf, dg1, dhdir);
          EVE_TRY(conv_dgrad(g1, dg1, cw[0], nullptr, dxh, sc.cs, s));
          LAUNCH1D(cgru_bwd3_kernel, rows * nf, dcat2, dxh, dhdir, rows, nf, dxo, ex.dcarry[i]);
        } else {
          LAUNCH1D(crnn_bwd_kernel, rows * nf, dcur, ex.dcarry[i], c.h + (size_t)t * B * E,
                   rows * nf, dg1);
          EVE_TRY(conv_dgrad(g1, dg1, cw[0], nullptr, dxh, sc.cs, s));
          LAUNCH1D(cgru_bwd3_kernel, rows * nf, (const float*)nullptr, dxh, (const float*)nullptr,
                   rows, nf, dxo, ex.dcarry[i]);
        }
        dcur = dxo;
      }
    }
    conv_prepared_truncate(prep_mark);
    dbx = ex.dbx;
    // weight gradients: one batched wgrad over all T*B bottleneck images per conv
    ConvGeom G1 = make_conv(N, kLevelH[4], kLevelW[4], 2 * nf, g1.Cout, 3, 1, 1);
    ConvGeom G2 = make_conv(N, kLevelH[4], kLevelW[4], 2 * nf, nf, 3, 1, 1);
    for (int i = 0; i < nc; ++i) {
      float* const* cg = gr + n.slot_rnn + wpc * i;
      EVE_TRY(conv_param_grads(G1, n.cell[i].xh, ex.dg1all[i], cg[0], cg[1], sc.cs, acc, s));
      if (gru)
        EVE_TRY(conv_param_grads(G2, n.cell[i].cat2, ex.dg2all[i], cg[2], cg[3], sc.cs, acc, s));
      if (dh0)
        EVE_TRY(nhwc_to_nchw(ex.dcarry[i], B, nf, kLevelH[4], kLevelW[4], dh0 + (size_t)i * B * E, s));
    }
  }

  // ---- encoder, innermost first.  `cur` <- grad w.r.t. the encoder output of level 4
  LAUNCH1D(swap_bt_kernel, (long long)N * E, dbx, (long long)N * E, T, B, P, nf, nf, cur, nf);
  if (p->use_skip) EVE_TRY(ew_add(cur, ex.dskip[4], (long long)N * E, cur, s));
  // In the fused pipeline a block hands the bf16 planes of its dx (and the bias-gradient column
  // sums) straight to the block in front of it; `have` tells whether `cur` comes with planes.
  Planes pcur = fb.DA, pnext = fb.DB;
  bool have = false;
  for (int l = kLevels - 1; l >= 0; --l) {
    for (int j = (int)n.enc[l].size() - 1; j >= 0; --j) {
      const RBlock& k = n.enc[l][j];
      if (fusedn) {
        // consumer of this block's dx: the previous block of the level, or initial.3 at the very end
        const bool to_block = j > 0;
        const bool to_initial = l == 0 && j == 0;
        float* const* ng = to_block ? gr + n.enc[l][j - 1].slot : nullptr;
        float* nb_a = to_block ? ng[7] : (to_initial ? gr[5] : nullptr);
        float* nb_b = to_block && n.enc[l][j - 1].skipconv ? ng[11] : nullptr;
        const bool emit = to_block || to_initial;
        FusedBwd fl = fb;
        fl.DA = pcur;          // where split_colsum puts the planes when `cur` came without them
        EVE_TRY(block_bwd_fused(k, N, w, gr, acc, cur, have ? &pcur : nullptr, other,
                                emit ? &pnext : nullptr, nb_a, nb_b, sc, fl, s));
        have = emit;
        Planes tp = pcur; pcur = pnext; pnext = tp;
      } else {
        EVE_TRY(block_bwd(k, N, w, gr, acc, cur, other, sc, s));
      }
      float* tmp = cur; cur = other; other = tmp;
    }
    if (l > 0) {
      const RBlock& e = n.enc[l - 1].back();
      // + the gradient that reached this encoder output through the skip connection
      EVE_TRY(adaptive_maxpool_bwd(cur, n.pidx[l - 1], N, e.H, e.W, e.oc, kLevelH[l], kLevelW[l],
                                   other, s, p->use_skip ? ex.dskip[l - 1] : nullptr,
                                   ex.dskip_ld[l - 1]));
      float* tmp = cur; cur = other; other = tmp;
      have = false;
    }
  }
  // ---- initial
  if (fusedn) {
    // `cur` arrives with its bf16 planes in pcur and initial.3's bias gradient already reduced
    EVE_TRY(in_apply_planes2(n.i0, N, HW0, 16, n.im, n.ir, w[2], w[3], nullptr, nullptr, ACT_RELU,
                             TC_BF16, fb.XP.hi, fb.XP.lo, nullptr, nullptr, s));
    EVE_TRY(conv_bwd_planes(n.gi3, fb.XP.hi, fb.XP.lo, pcur.hi, pcur.lo, w[4], gr[4], acc, nullptr,
                            sc.t0, sc.cs, s));
    EVE_TRY(in_bwd_fused(sc.t0, nullptr, nullptr, n.i0, N, HW0, 16, n.im, n.ir, w[2], w[3], nullptr,
                         nullptr, ACT_RELU, nullptr, sc.t1, nullptr, nullptr, nullptr, gr[2], gr[3],
                         nullptr, nullptr, gr[1], nullptr, acc, fb.col, s));
  } else {
    bool i1_fused = false;
    EVE_TRY(norm_act_into_conv(n.gi3, true, n.i0, N, HW0, 16, n.im, n.ir, w[2], w[3], ACT_RELU,
                               nullptr, sc.cs, &i1_fused, s));
    EVE_TRY(conv_bwd(n.gi3, i1_fused ? nullptr : n.i1, cur, w[4], gr[4], gr[5], acc, nullptr, sc.t0,
                     sc.cs, s));
    EVE_TRY(in_backward(sc.t0, nullptr, n.i0, N, HW0, 16, n.im, n.ir, w[2], w[3], ACT_RELU, nullptr,
                        sc.t1, nullptr, gr[2], gr[3], sc.inb, acc, s));
  }
  {
    // zero-padded weights (and their gradient) live at the head of t2, which is free here
    float* w0p = sc.t2;
    float* dw0p = sc.t2 + 16 * kInitC * 9;
    LAUNCH1D(pad_cin_kernel, 16 * kInitC * 9, w[0], 16, p->in_channels, kInitC, w0p);
    EVE_TRY(conv_bwd(n.gi0, n.x0, sc.t1, w0p, gr[0] ? dw0p : nullptr, nullptr, false, nullptr,
                     dheatmap ? sc.t0 : nullptr, sc.cs, s));
    if (gr[1] && !fusedn)
      EVE_TRY(colsum(sc.t1, (long long)N * HW0, 16, 16, gr[1], sc.t2 + 2 * 16 * kInitC * 9, acc, s));
    if (gr[0])
      LAUNCH1D(unpad_cin_kernel, 16 * p->in_channels * 9, dw0p, 16, p->in_channels, kInitC, gr[0],
               acc ? 1 : 0);
    if (dheatmap)
      EVE_TRY(copy_channels(sc.t0, (long long)N * HW0, 1, kInitC, p->in_channels - 1, dheatmap, 1, 0,
                            false, s));
  }
  conv_prepared_clear();
  return EVE_OK;
}
