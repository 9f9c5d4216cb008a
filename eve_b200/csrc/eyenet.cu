// EyeNet: ResNet-18 / InstanceNorm encoder (forward + backward) and the recurrent tail.
//
// Reference: src/models/eye_net.py:37-150 and torchvision.models.resnet (BasicBlock, ResNet)
// as instantiated at eye_net.py:48-50.  The CNN runs for all patches of a step at once
// (norms are per sample); only the RNN cell walks over time.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace eve {

namespace {

struct BlockTape {
  ConvGeom g1, g2, gd;
  bool down;
  const float* in;  // block input (previous output)
  float *a, *am, *ar, *y, *b, *bm, *br, *d, *dm, *dr, *out;
};

struct CnnTape {
  int N, nf;
  ConvGeom stem;
  float *x, *c1, *c1m, *c1r, *p;
  int32_t* pidx;
  BlockTape blk[8];
  float* pooled;
  size_t max_act;  // largest activation (floats) among block tensors
};

// Lays out every saved activation in the caller's `saved` buffer.  Called identically by
// forward, backward and the size query (dry arena).
bool build_cnn_tape(const eve_eyenet_cnn_params& p, Arena& sv, CnnTape& t) {
  const int N = p.n;
  t.N = N;
  t.nf = p.nf;
  t.stem = make_conv(N, p.h, p.w, 3, 64, 7, 2, 3);
  t.x = sv.get<float>(t.stem.in_elems());
  t.c1 = sv.get<float>(t.stem.out_elems());
  t.c1m = sv.get<float>((size_t)N * 64);
  t.c1r = sv.get<float>((size_t)N * 64);
  int H = (t.stem.OH + 2 - 3) / 2 + 1, W = (t.stem.OW + 2 - 3) / 2 + 1;
  t.p = sv.get<float>((size_t)N * H * W * 64);
  t.pidx = sv.get<int32_t>((size_t)N * H * W * 64);
  t.max_act = (size_t)N * H * W * 64;
  const float* in = t.p;
  int cin = 64;
  const int couts[4] = {64, 128, 256, 512};
  for (int l = 0; l < 4; ++l) {
    for (int b = 0; b < 2; ++b) {
      BlockTape& k = t.blk[l * 2 + b];
      int cout = couts[l];
      int stride = (l > 0 && b == 0) ? 2 : 1;
      k.down = (l > 0 && b == 0);
      k.in = in;
      k.g1 = make_conv(N, H, W, cin, cout, 3, stride, 1);
      k.g2 = make_conv(N, k.g1.OH, k.g1.OW, cout, cout, 3, 1, 1);
      k.gd = make_conv(N, H, W, cin, cout, 1, stride, 0);
      size_t oe = (size_t)k.g1.out_elems();
      k.a = sv.get<float>(oe);
      k.am = sv.get<float>((size_t)N * cout);
      k.ar = sv.get<float>((size_t)N * cout);
      k.y = sv.get<float>(oe);
      k.b = sv.get<float>(oe);
      k.bm = sv.get<float>((size_t)N * cout);
      k.br = sv.get<float>((size_t)N * cout);
      if (k.down) {
        k.d = sv.get<float>(oe);
        k.dm = sv.get<float>((size_t)N * cout);
        k.dr = sv.get<float>((size_t)N * cout);
      } else {
        k.d = k.dm = k.dr = nullptr;
      }
      k.out = sv.get<float>(oe);
      if (oe > t.max_act) t.max_act = oe;
      in = k.out;
      cin = cout;
      H = k.g1.OH;
      W = k.g1.OW;
    }
  }
  t.pooled = sv.get<float>((size_t)N * 512);
  return sv.ok();
}

inline size_t conv_w_floats(const ConvGeom& g) { return (size_t)g.Cout * g.K(); }

// weight table slot of block (l,b): returns slot of conv1; conv2 = +1; downsample = +2
int block_slot(int blk) {
  // slots: 0 stem; layer1: 1,2 | 3,4 ; layer2: 5,6,7 | 8,9 ; layer3: 10,11,12 | 13,14 ;
  // layer4: 15,16,17 | 18,19 ; fc.w 20 ; fc.b 21
  static const int s[8] = {1, 3, 5, 8, 10, 13, 15, 18};
  return s[blk];
}

int check_cnn(const eve_eyenet_cnn_params* p) {
  EVE_REQUIRE(p, EVE_ERR_NULL, "eyenet_cnn: params is NULL");
  EVE_REQUIRE(p->n >= 0 && p->nf > 0 && p->h >= 32 && p->w >= 32 && p->h % 32 == 0 &&
                  p->w % 32 == 0,
              EVE_ERR_SHAPE, "eyenet_cnn: unsupported shape n=%d nf=%d h=%d w=%d", p->n, p->nf,
              p->h, p->w);
  return EVE_OK;
}

// conv scratch (operand planes, weight layouts, split-K partials) big enough for every
// convolution of the CNN
size_t cnn_conv_scratch_bytes(const CnnTape& t) {
  size_t mi = 0, mo = 0, mw = 0, mp = 0;
  auto upd = [&](const ConvGeom& g) {
    mi = std::max(mi, conv_operand_elems(g));
    mo = std::max(mo, (size_t)g.out_elems());
    mw = std::max(mw, (size_t)g.Cout * g.K());
    mp = std::max(mp, conv_partial_floats(g));
  };
  upd(t.stem);
  for (int i = 0; i < 8; ++i) {
    upd(t.blk[i].g1);
    upd(t.blk[i].g2);
    if (t.blk[i].down) upd(t.blk[i].gd);
  }
  return conv_scratch_bytes(mi, mo, mw, mp);
}

// the sixteen 3x3 + three 1x1 block convolutions (the stem has its own patch-matrix layout)
std::vector<ConvPrepReq> cnn_prep_list(const CnnTape& t, const float* const* w) {
  std::vector<ConvPrepReq> v;
  for (int i = 0; i < 8; ++i) {
    const BlockTape& k = t.blk[i];
    const int slot = block_slot(i);
    v.push_back(ConvPrepReq{k.g1, w ? w[slot] : nullptr});
    v.push_back(ConvPrepReq{k.g2, w ? w[slot + 1] : nullptr});
    if (k.down) v.push_back(ConvPrepReq{k.gd, w ? w[slot + 2] : nullptr});
  }
  return v;
}
size_t cnn_prep_bytes(const CnnTape& t) {
  const std::vector<ConvPrepReq> v = cnn_prep_list(t, nullptr);
  return conv_prepare_batch_bytes(v.data(), (int)v.size());
}

size_t cnn_fwd_scratch(const CnnTape& t) {
  return align_up(cnn_conv_scratch_bytes(t), 256) +
         align_up((size_t)512 * t.nf * sizeof(float), 256) +
         4 * align_up(t.max_act * sizeof(uint16_t), 256) +         // operand planes IN, Y
         align_up(cnn_prep_bytes(t), 256) + 1024;
}

struct Planes {
  uint16_t *hi, *lo;
};
Planes get_planes(Arena& ws, size_t elems) {
  Planes p;
  p.hi = ws.get<uint16_t>(elems);
  p.lo = ws.get<uint16_t>(elems);
  return p;
}

// plane-to-plane pipeline (one-pass normalisations, in_fused.cu): every 3x3 / 1x1 convolution of
// the eight blocks takes the split tensor-core path and every map fits the cluster kernels
bool cnn_fused(const CnnTape& t) {
  if (!get_option(OPT_FUSED_NORM)) return false;
  for (int i = 0; i < 8; ++i) {
    const BlockTape& k = t.blk[i];
    if (!conv_x_fusable(k.g1) || !conv_x_fusable(k.g2)) return false;
    if (k.down && !conv_x_fusable(k.gd)) return false;
    if (!in_fused_supported(k.g1.OH * k.g1.OW, k.g1.Cout, 2)) return false;
  }
  return true;
}

}  // namespace

}  // namespace eve

using namespace eve;

extern "C" size_t eve_eyenet_cnn_saved_bytes(const eve_eyenet_cnn_params* p) {
  if (check_cnn(p) != EVE_OK) return 0;
  Arena sv(nullptr, 0);
  CnnTape t;
  build_cnn_tape(*p, sv, t);
  return sv.off;
}

namespace eve {
namespace {

// Scratch plan of the backward pass (dry-run capable).
struct CnnBwdScratch {
  float *wg, *inb, *ga, *gb, *t0, *t1, *t2, *t3, *stem_g, *stem_d, *dpooled;
  ConvScratch cs;
  Planes XP, XI, DB, DA;   // fused pipeline: x planes (y / block input), dy planes
  float* col;
  char* prep;              // batched weight layouts (conv_prepare_batch)
  size_t prep_bytes;
};

bool build_cnn_bwd_scratch(const CnnTape& t, Arena& ws, CnnBwdScratch& s) {
  s.cs.bytes = cnn_conv_scratch_bytes(t);
  s.cs.base = ws.get<char>(s.cs.bytes);
  s.wg = ws.get<float>(linear_wgrad_scratch_floats(t.N, 512, t.nf));
  s.inb = ws.get<float>(in_backward_scratch_floats(t.N, 512));
  s.ga = ws.get<float>(t.max_act);
  s.gb = ws.get<float>(t.max_act);
  s.t0 = ws.get<float>(t.max_act);
  s.t1 = ws.get<float>(t.max_act);
  s.t2 = ws.get<float>(t.max_act);
  s.t3 = ws.get<float>(t.max_act);
  s.stem_g = ws.get<float>((size_t)t.stem.out_elems());
  s.stem_d = ws.get<float>((size_t)t.stem.out_elems());
  s.dpooled = ws.get<float>((size_t)t.N * 512);
  s.XP = get_planes(ws, t.max_act);
  s.XI = get_planes(ws, t.max_act);
  s.DB = get_planes(ws, t.max_act);
  s.DA = get_planes(ws, t.max_act);
  s.col = ws.get<float>((size_t)12 * t.N * 512);
  s.prep_bytes = cnn_prep_bytes(t);
  s.prep = ws.get<char>(s.prep_bytes);
  return ws.ok();
}

}  // namespace
}  // namespace eve

extern "C" size_t eve_eyenet_cnn_workspace_bytes(const eve_eyenet_cnn_params* p) {
  if (check_cnn(p) != EVE_OK) return 0;
  Arena sv(nullptr, 0);
  CnnTape t;
  build_cnn_tape(*p, sv, t);
  Arena ws(nullptr, 0);
  CnnBwdScratch s;
  build_cnn_bwd_scratch(t, ws, s);
  size_t f = cnn_fwd_scratch(t);
  return ws.off > f ? ws.off : f;
}

extern "C" int eve_eyenet_cnn_fwd(const eve_eyenet_cnn_params* p, const float* x,
                                  const float* const* w, float* feat, void* saved,
                                  size_t saved_bytes, void* workspace, size_t workspace_bytes,
                                  eve_stream_t stream) {
  EVE_TRY(check_cnn(p));
  if (p->n == 0) return EVE_OK;
  EVE_REQUIRE(x && w && feat && saved && workspace, EVE_ERR_NULL, "eyenet_cnn_fwd: NULL pointer");
  cudaStream_t s = as_stream(stream);
  Arena sv(saved, saved_bytes);
  CnnTape t;
  EVE_REQUIRE(build_cnn_tape(*p, sv, t), EVE_ERR_WORKSPACE,
              "eyenet_cnn_fwd: saved buffer too small (%zu < %zu)", saved_bytes, sv.off);
  EVE_REQUIRE(workspace_bytes >= cnn_fwd_scratch(t), EVE_ERR_WORKSPACE,
              "eyenet_cnn_fwd: workspace too small");
  Arena ws(workspace, workspace_bytes);
  ConvScratch cs;
  cs.bytes = cnn_conv_scratch_bytes(t);
  cs.base = ws.get<char>(cs.bytes);
  float* wfc = ws.get<float>((size_t)512 * t.nf);
  const int N = t.N;
  conv_prepared_clear();
  {
    // forward tensor-core layouts of the 19 block convolutions: one launch
    const size_t pb = cnn_prep_bytes(t);
    char* region = ws.get<char>(pb);
    EVE_REQUIRE(region, EVE_ERR_WORKSPACE, "eyenet_cnn_fwd: workspace too small");
    const std::vector<ConvPrepReq> reqs = cnn_prep_list(t, w);
    EVE_TRY(conv_prepare_batch(reqs.data(), (int)reqs.size(), false, region, pb, s));
  }

  EVE_TRY(nchw_to_nhwc(x, N, 3, p->h, p->w, t.x, s));
  EVE_TRY(conv_fwd(t.stem, t.x, w[0], nullptr, nullptr, t.c1, cs, s));
  EVE_TRY(in_stats(t.c1, N, t.stem.OH * t.stem.OW, 64, t.c1m, t.c1r, s));
  EVE_TRY(in_relu_maxpool(t.c1, N, t.stem.OH, t.stem.OW, 64, t.c1m, t.c1r, t.p, t.pidx, s));
  if (cnn_fused(t)) {
    const Planes PI = get_planes(ws, t.max_act), PY = get_planes(ws, t.max_act);
    EVE_REQUIRE(ws.ok(), EVE_ERR_WORKSPACE, "eyenet_cnn_fwd: workspace too small");
    // PI always holds the fp16 planes of the current block input: written here for the first block,
    // afterwards by the kernel that finishes the previous block
    EVE_TRY(split_planes(t.p, (long long)t.blk[0].g1.in_elems(), PI.hi, PI.lo, TC_F16, s));
    for (int i = 0; i < 8; ++i) {
      BlockTape& k = t.blk[i];
      const int slot = block_slot(i);
      const int C = k.g1.Cout, HW = k.g1.OH * k.g1.OW;
      EVE_TRY(conv_fwd_planes(k.g1, PI.hi, PI.lo, w[slot], nullptr, nullptr, k.a, cs, s));
      EVE_TRY(in_fwd_fused(k.a, N, HW, C, nullptr, 0, nullptr, nullptr, nullptr, nullptr, ACT_RELU,
                           TC_F16, k.am, k.ar, nullptr, nullptr, nullptr, PY.hi, PY.lo, nullptr,
                           nullptr, s));
      EVE_TRY(conv_fwd_planes(k.g2, PY.hi, PY.lo, w[slot + 1], nullptr, nullptr, k.b, cs, s));
      if (k.down)
        EVE_TRY(conv_fwd_planes(k.gd, PI.hi, PI.lo, w[slot + 2], nullptr, nullptr, k.d, cs, s));
      // out = relu(IN(b) + (x | IN(d))): statistics of b (and d), the fp32 output (residual of the
      // next block, mask of the backward pass) and the next block's operand planes in one pass
      const bool last = i == 7;
      EVE_TRY(in_fwd_fused(k.b, N, HW, C, k.down ? k.d : k.in, k.down ? 2 : 1, nullptr, nullptr,
                           nullptr, nullptr, ACT_RELU, TC_F16, k.bm, k.br, k.dm, k.dr, k.out,
                           last ? nullptr : PI.hi, last ? nullptr : PI.lo, nullptr, nullptr, s));
    }
  } else
  for (int i = 0; i < 8; ++i) {
    BlockTape& k = t.blk[i];
    const int slot = block_slot(i);
    const int C = k.g1.Cout, HW = k.g1.OH * k.g1.OW;
    EVE_TRY(conv_fwd(k.g1, k.in, w[slot], nullptr, nullptr, k.a, cs, s));
    EVE_TRY(in_stats(k.a, N, HW, C, k.am, k.ar, s));
    bool fused = false;
    EVE_TRY(norm_act_into_conv(k.g2, false, k.a, N, HW, C, k.am, k.ar, nullptr, nullptr, ACT_RELU,
                               k.y, cs, &fused, s));
    EVE_TRY(conv_fwd(k.g2, fused ? nullptr : k.y, w[slot + 1], nullptr, nullptr, k.b, cs, s));
    EVE_TRY(in_stats(k.b, N, HW, C, k.bm, k.br, s));
    if (k.down) {
      EVE_TRY(conv_fwd(k.gd, k.in, w[slot + 2], nullptr, nullptr, k.d, cs, s));
      EVE_TRY(in_stats(k.d, N, HW, C, k.dm, k.dr, s));
      EVE_TRY(in_apply(k.b, N, HW, C, k.bm, k.br, nullptr, nullptr, k.d, k.dm, k.dr, ACT_RELU,
                       k.out, s));
    } else {
      EVE_TRY(in_apply(k.b, N, HW, C, k.bm, k.br, nullptr, nullptr, k.in, nullptr, nullptr,
                       ACT_RELU, k.out, s));
    }
  }
  const BlockTape& last = t.blk[7];
  EVE_TRY(avgpool_fwd(last.out, N, last.g1.OH * last.g1.OW, 512, t.pooled, s));
  EVE_TRY(linear_fwd(t.pooled, N, 512, w[20], w[21], t.nf, feat, wfc, s));
  conv_prepared_clear();
  return EVE_OK;
}

extern "C" int eve_eyenet_cnn_bwd(const eve_eyenet_cnn_params* p, const float* dfeat,
                                  const float* const* w, float* const* gr, int accumulate,
                                  const void* saved, size_t saved_bytes, void* workspace,
                                  size_t workspace_bytes, eve_stream_t stream) {
  EVE_TRY(check_cnn(p));
  if (p->n == 0) return EVE_OK;
  EVE_REQUIRE(dfeat && w && gr && saved && workspace, EVE_ERR_NULL,
              "eyenet_cnn_bwd: NULL pointer");
  cudaStream_t s = as_stream(stream);
  Arena sv(const_cast<void*>(saved), saved_bytes);
  CnnTape t;
  EVE_REQUIRE(build_cnn_tape(*p, sv, t), EVE_ERR_WORKSPACE,
              "eyenet_cnn_bwd: saved buffer too small");
  Arena ws(workspace, workspace_bytes);
  CnnBwdScratch sc;
  EVE_REQUIRE(build_cnn_bwd_scratch(t, ws, sc), EVE_ERR_WORKSPACE,
              "eyenet_cnn_bwd: workspace too small (%zu < %zu)", workspace_bytes, ws.off);
  const int N = t.N;
  const bool acc = accumulate != 0;
  conv_prepared_clear();
  {
    // flipped data-gradient layouts of the block convolutions: one launch
    const std::vector<ConvPrepReq> reqs = cnn_prep_list(t, w);
    EVE_TRY(conv_prepare_batch(reqs.data(), (int)reqs.size(), true, sc.prep, sc.prep_bytes, s));
  }

  // fc
  EVE_TRY(linear_wgrad(t.pooled, dfeat, N, 512, t.nf, gr[20], gr[21], sc.wg, acc, s));
  EVE_TRY(linear_dgrad(dfeat, N, t.nf, w[20], 512, nullptr, sc.dpooled, s));
  const BlockTape& last = t.blk[7];
  float* dout = sc.ga;
  float* dnext = sc.gb;
  EVE_TRY(avgpool_bwd(sc.dpooled, N, last.g1.OH * last.g1.OW, 512, dout, s));

  if (cnn_fused(t)) {
    for (int i = 7; i >= 0; --i) {
      const BlockTape& k = t.blk[i];
      const int slot = block_slot(i);
      const int C = k.g1.Cout, HW = k.g1.OH * k.g1.OW;
      float* gskip = sc.t1;
      // out = relu(IN(b) + skip): db as the bf16 dy planes of conv2, gskip = dout * relu'(out)
      EVE_TRY(in_bwd_fused(dout, nullptr, k.out, k.b, N, HW, C, k.bm, k.br, nullptr, nullptr, nullptr,
                           nullptr, ACT_RELU, nullptr, nullptr, sc.DB.hi, sc.DB.lo, gskip, nullptr,
                           nullptr, nullptr, nullptr, nullptr, nullptr, false, sc.col, s));
      // conv2: data gradient first; the weight gradient's x operand relu(IN(a)) comes out of the
      // norm's backward kernel (it computes xhat anyway), not out of a separate pass over a
      EVE_TRY(conv_bwd_planes(k.g2, nullptr, nullptr, sc.DB.hi, sc.DB.lo, w[slot + 1], nullptr, acc,
                              nullptr, sc.t2, sc.cs, s));
      // y = relu(IN(a)): da as the dy planes of conv1, y as the x planes of conv2's weight gradient
      EVE_TRY(in_bwd_fused(sc.t2, nullptr, nullptr, k.a, N, HW, C, k.am, k.ar, nullptr, nullptr,
                           nullptr, nullptr, ACT_RELU, nullptr, nullptr, sc.DA.hi, sc.DA.lo, nullptr,
                           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, false, sc.col, s,
                           sc.XP.hi, sc.XP.lo));
      EVE_TRY(conv_bwd_planes(k.g2, sc.XP.hi, sc.XP.lo, sc.DB.hi, sc.DB.lo, w[slot + 1], gr[slot + 1],
                              acc, nullptr, nullptr, sc.cs, s));
      // the block input feeds conv1 and the downsample convolution: one bf16 split for both
      EVE_TRY(split_planes(k.in, (long long)k.g1.in_elems(), sc.XI.hi, sc.XI.lo, TC_BF16, s));
      const float* addend = gskip;
      if (k.down) {
        EVE_TRY(in_bwd_fused(gskip, nullptr, nullptr, k.d, N, HW, C, k.dm, k.dr, nullptr, nullptr,
                             nullptr, nullptr, ACT_NONE, nullptr, nullptr, sc.DB.hi, sc.DB.lo, nullptr,
                             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, false, sc.col, s));
        EVE_TRY(conv_bwd_planes(k.gd, sc.XI.hi, sc.XI.lo, sc.DB.hi, sc.DB.lo, w[slot + 2],
                                gr[slot + 2], acc, nullptr, sc.t3, sc.cs, s));
        addend = sc.t3;
      }
      EVE_TRY(conv_bwd_planes(k.g1, sc.XI.hi, sc.XI.lo, sc.DA.hi, sc.DA.lo, w[slot], gr[slot], acc,
                              addend, dnext, sc.cs, s));
      float* tmp = dout;
      dout = dnext;
      dnext = tmp;
    }
  } else
  for (int i = 7; i >= 0; --i) {
    const BlockTape& k = t.blk[i];
    const int slot = block_slot(i);
    const int C = k.g1.Cout, HW = k.g1.OH * k.g1.OW;
    float* db = sc.t0;     // grad wrt conv2 output
    float* gskip = sc.t1;  // grad wrt the residual branch
    // out = relu(IN(b) + skip)
    EVE_TRY(in_backward(dout, k.out, k.b, N, HW, C, k.bm, k.br, nullptr, nullptr, ACT_RELU, nullptr, db,
                        gskip, nullptr, nullptr, sc.inb, false, s));
    // conv2: weight gradient + data gradient from one split of db
    float* dy = sc.t2;
    bool fused = false;
    EVE_TRY(norm_act_into_conv(k.g2, true, k.a, N, HW, C, k.am, k.ar, nullptr, nullptr, ACT_RELU,
                               nullptr, sc.cs, &fused, s));
    EVE_TRY(conv_bwd(k.g2, fused ? nullptr : k.y, db, w[slot + 1], gr[slot + 1], nullptr, acc,
                     nullptr, dy, sc.cs, s));
    // y = relu(IN(a))
    float* da = sc.t0;
    EVE_TRY(in_backward(dy, nullptr, k.a, N, HW, C, k.am, k.ar, nullptr, nullptr, ACT_RELU, nullptr, da,
                        nullptr, nullptr, nullptr, sc.inb, false, s));
    const float* addend = gskip;
    if (k.down) {
      float* dd = sc.t2;
      EVE_TRY(in_backward(gskip, nullptr, k.d, N, HW, C, k.dm, k.dr, nullptr, nullptr, ACT_NONE,
                          nullptr, dd, nullptr, nullptr, nullptr, sc.inb, false, s));
      EVE_TRY(conv_bwd(k.gd, k.in, dd, w[slot + 2], gr[slot + 2], nullptr, acc, nullptr, sc.t3, sc.cs,
                       s));
      addend = sc.t3;
    }
    EVE_TRY(conv_bwd(k.g1, k.in, da, w[slot], gr[slot], nullptr, acc, addend, dnext, sc.cs, s));
    float* tmp = dout;
    dout = dnext;
    dnext = tmp;
  }
  // stem: p = maxpool(relu(IN(c1)))
  if (gr[0]) {
    const int OH = t.stem.OH, OW = t.stem.OW;
    const int PH = (OH + 2 - 3) / 2 + 1, PW = (OW + 2 - 3) / 2 + 1;
    if (get_option(OPT_STEM_FUSED_BWD) && OH % 2 == 0 && OW % 2 == 0) {
      // no dense un-pooled gradient: window sums + one gather pass (norm.cu)
      if (conv_wgrad_stem_takes_planes(t.stem)) {
        uint16_t *d_hi, *d_lo;
        EVE_TRY(conv_wgrad_stem_planes(t.stem, sc.cs, &d_hi, &d_lo));
        EVE_TRY(stem_pool_in_backward(dout, t.p, t.pidx, t.c1, N, OH, OW, 64, t.c1m, t.c1r, nullptr,
                                      d_hi, d_lo, sc.inb, s));
        EVE_TRY(conv_wgrad_stem_run(t.stem, t.x, gr[0], acc, sc.cs, s));
      } else {
        EVE_TRY(stem_pool_in_backward(dout, t.p, t.pidx, t.c1, N, OH, OW, 64, t.c1m, t.c1r,
                                      sc.stem_d, nullptr, nullptr, sc.inb, s));
        EVE_TRY(conv_wgrad(t.stem, t.x, sc.stem_d, gr[0], nullptr, acc, sc.cs, s));
      }
    } else {
      EVE_TRY(maxpool_bwd_scatter(dout, t.pidx, N, OH, OW, PH, PW, 64, sc.stem_g, s));
      EVE_TRY(in_backward(sc.stem_g, nullptr, t.c1, N, OH * OW, 64, t.c1m, t.c1r, nullptr, nullptr,
                          ACT_RELU, nullptr, sc.stem_d, nullptr, nullptr, nullptr, sc.inb, false, s));
      EVE_TRY(conv_wgrad(t.stem, t.x, sc.stem_d, gr[0], nullptr, acc, sc.cs, s));
    }
  }
  conv_prepared_clear();
  return EVE_OK;
}

// =========================================================================== tail ====
// eye_net.py:109-140 for whole sequences.  Everything that is not recurrent is a batched
// GEMM over R = batch*steps rows; the cell recurrences run in one persistent CTA per
// sequence with warp-coalesced weight reads (the weights stay L1/L2 resident).
namespace eve {
namespace {

constexpr int kMaxCells = 8;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

inline int rnn_gates(int type) { return type == EVE_RNN_GRU ? 3 : type == EVE_RNN_LSTM ? 4 : 1; }
inline int rnn_saved_per_unit(int type) {
  return type == EVE_RNN_GRU ? 4 : type == EVE_RNN_LSTM ? 5 : 0;
}

// grid = batch, block = nf.  gi[B,T,G*nf]; whT[nf][G*nf] (transposed weight_hh); bhh[G*nf].
template <int TYPE>
__global__ void rnn_seq_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ whT,
                                   const float* __restrict__ bhh, const float* __restrict__ h0,
                                   const float* __restrict__ c0, int T, int nf,
                                   float* __restrict__ H, float* __restrict__ Hprev,
                                   float* __restrict__ gates, float* __restrict__ hT,
                                   float* __restrict__ cT) {
  extern __shared__ float hs[];
  constexpr int G = TYPE == EVE_RNN_GRU ? 3 : TYPE == EVE_RNN_LSTM ? 4 : 1;
  constexpr int S = TYPE == EVE_RNN_GRU ? 4 : TYPE == EVE_RNN_LSTM ? 5 : 0;
  const int b = blockIdx.x, j = threadIdx.x;
  float h = h0 ? h0[(size_t)b * nf + j] : 0.f;
  float c = (TYPE == EVE_RNN_LSTM && c0) ? c0[(size_t)b * nf + j] : 0.f;
  const int ld = G * nf;
  for (int t = 0; t < T; ++t) {
    const size_t row = (size_t)b * T + t;
    hs[j] = h;
    Hprev[row * nf + j] = h;
    __syncthreads();
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = bhh[g * nf + j];
    for (int k = 0; k < nf; ++k) {
      const float hk = hs[k];
      const float* wr = whT + (size_t)k * ld + j;
#pragma unroll
      for (int g = 0; g < G; ++g) acc[g] = fmaf(__ldg(wr + g * nf), hk, acc[g]);
    }
    const float* gir = gi + row * ld + j;
    float* sv = gates + row * (size_t)(S * nf) + j;
    if (TYPE == EVE_RNN_GRU) {
      float r = sigmoidf_(gir[0] + acc[0]);
      float z = sigmoidf_(gir[nf] + acc[1]);
      float n = tanhf(gir[2 * nf] + r * acc[2]);
      sv[0] = r; sv[nf] = z; sv[2 * nf] = n; sv[3 * nf] = acc[2];
      h = (1.f - z) * n + z * h;
    } else if (TYPE == EVE_RNN_LSTM) {
      float ig = sigmoidf_(gir[0] + acc[0]);
      float fg = sigmoidf_(gir[nf] + acc[1 % G]);
      float gg = tanhf(gir[2 * nf] + acc[2 % G]);
      float og = sigmoidf_(gir[3 * nf] + acc[3 % G]);
      c = fg * c + ig * gg;
      h = og * tanhf(c);
      sv[0] = ig; sv[nf] = fg; sv[2 * nf] = gg; sv[3 * nf] = og; sv[4 * nf] = c;
    } else {
      h = tanhf(gir[0] + acc[0]);
    }
    H[row * nf + j] = h;
    __syncthreads();
  }
  if (hT) hT[(size_t)b * nf + j] = h;
  if (TYPE == EVE_RNN_LSTM && cT) cT[(size_t)b * nf + j] = c;
}

// grid = batch, block = nf.  whh[G*nf][nf] (torch layout).  dgh aliases dgi unless GRU.
template <int TYPE>
__global__ void rnn_seq_bwd_kernel(const float* __restrict__ dHext, const float* __restrict__ dhT,
                                   const float* __restrict__ dcT, const float* __restrict__ whh,
                                   const float* __restrict__ H, const float* __restrict__ Hprev,
                                   const float* __restrict__ gates, const float* __restrict__ c0,
                                   int T, int nf, float* __restrict__ dgi,
                                   float* __restrict__ dgh, float* __restrict__ dh0,
                                   float* __restrict__ dc0) {
  extern __shared__ float ds[];  // G*nf
  constexpr int G = TYPE == EVE_RNN_GRU ? 3 : TYPE == EVE_RNN_LSTM ? 4 : 1;
  constexpr int S = TYPE == EVE_RNN_GRU ? 4 : TYPE == EVE_RNN_LSTM ? 5 : 0;
  const int b = blockIdx.x, k = threadIdx.x;
  const int ld = G * nf;
  float dh_carry = dhT ? dhT[(size_t)b * nf + k] : 0.f;
  float dc_carry = (TYPE == EVE_RNN_LSTM && dcT) ? dcT[(size_t)b * nf + k] : 0.f;
  for (int t = T - 1; t >= 0; --t) {
    const size_t row = (size_t)b * T + t;
    const float dh = dHext[row * nf + k] + dh_carry;
    const float* sv = gates + row * (size_t)(S * nf) + k;
    float direct = 0.f;
    float* gi_o = dgi + row * ld + k;
    float* gh_o = dgh + row * ld + k;
    if (TYPE == EVE_RNN_GRU) {
      float r = sv[0], z = sv[nf], n = sv[2 * nf], hn = sv[3 * nf];
      float hp = Hprev[row * nf + k];
      float dn = dh * (1.f - z);
      float dz = dh * (hp - n);
      direct = dh * z;
      float dpn = dn * (1.f - n * n);
      float dr = dpn * hn;
      float dpr = dr * r * (1.f - r);
      float dpz = dz * z * (1.f - z);
      gi_o[0] = dpr; gi_o[nf] = dpz; gi_o[2 * nf] = dpn;
      float dghn = dpn * r;
      gh_o[0] = dpr; gh_o[nf] = dpz; gh_o[2 * nf] = dghn;
      ds[k] = dpr; ds[nf + k] = dpz; ds[2 * nf + k] = dghn;
    } else if (TYPE == EVE_RNN_LSTM) {
      float ig = sv[0], fg = sv[nf], gg = sv[2 * nf], og = sv[3 * nf], c = sv[4 * nf];
      float cp = t > 0 ? gates[(row - 1) * (size_t)(S * nf) + 4 * nf + k]
                       : (c0 ? c0[(size_t)b * nf + k] : 0.f);
      float tc = tanhf(c);
      float d_o = dh * tc;
      float dc = dh * og * (1.f - tc * tc) + dc_carry;
      float di = dc * gg, dg = dc * ig, df = dc * cp;
      dc_carry = dc * fg;
      float pi = di * ig * (1.f - ig), pf = df * fg * (1.f - fg);
      float pg = dg * (1.f - gg * gg), po = d_o * og * (1.f - og);
      gi_o[0] = pi; gi_o[nf] = pf; gi_o[2 * nf] = pg; gi_o[3 * nf] = po;
      ds[k] = pi; ds[nf + k] = pf; ds[2 * nf + k] = pg; ds[3 * nf + k] = po;
    } else {
      float h = H[row * nf + k];
      float dp = dh * (1.f - h * h);
      gi_o[0] = dp;
      ds[k] = dp;
    }
    __syncthreads();
    float acc = direct;
    for (int j = 0; j < ld; ++j) acc = fmaf(__ldg(whh + (size_t)j * nf + k), ds[j], acc);
    dh_carry = acc;
    __syncthreads();
  }
  if (dh0) dh0[(size_t)b * nf + k] = dh_carry;
  if (TYPE == EVE_RNN_LSTM && dc0) dc0[(size_t)b * nf + k] = dc_carry;
}

struct TailTape {
  int R, nf, in0, G, S;
  float *xin, *u1, *f1, *f2;
  float *H[kMaxCells], *Hprev[kMaxCells], *gates[kMaxCells], *c0[kMaxCells];
  float *u3, *f3;
  const float* feat_final;
  float *ug, *sg, *gout, *up, *sp, *vp;
};

bool build_tail_tape(const eve_eyenet_tail_params& p, Arena& sv, TailTape& t) {
  const size_t R = (size_t)p.batch * p.steps;
  t.R = (int)R;
  t.nf = p.nf;
  t.in0 = p.nf + (p.use_head_pose ? 2 : 0);
  t.G = rnn_gates(p.rnn_type);
  t.S = rnn_saved_per_unit(p.rnn_type);
  t.xin = sv.get<float>(R * t.in0);
  t.u1 = sv.get<float>(R * p.nf);
  t.f1 = sv.get<float>(R * p.nf);
  t.f2 = sv.get<float>(R * p.nf);
  t.feat_final = t.f2;
  t.u3 = t.f3 = nullptr;
  if (p.rnn_type != EVE_RNN_NONE) {
    for (int i = 0; i < p.rnn_cells; ++i) {
      t.H[i] = sv.get<float>(R * p.nf);
      t.Hprev[i] = sv.get<float>(R * p.nf);
      t.gates[i] = sv.get<float>(R * p.nf * (t.S > 0 ? t.S : 1));
      t.c0[i] = p.rnn_type == EVE_RNN_LSTM ? sv.get<float>((size_t)p.batch * p.nf) : nullptr;
      t.feat_final = t.H[i];
    }
  } else {
    t.u3 = sv.get<float>(R * p.nf);
    t.f3 = sv.get<float>(R * p.nf);
    t.feat_final = t.f3;
  }
  t.ug = sv.get<float>(R * p.nf);
  t.sg = sv.get<float>(R * p.nf);
  t.gout = sv.get<float>(R * 2);
  t.up = sv.get<float>(R * p.nf);
  t.sp = sv.get<float>(R * p.nf);
  t.vp = sv.get<float>(R);
  return sv.ok();
}

int check_tail(const eve_eyenet_tail_params* p) {
  EVE_REQUIRE(p, EVE_ERR_NULL, "eyenet_tail: params is NULL");
  EVE_REQUIRE(p->batch >= 0 && p->steps >= 0 && p->nf > 0 && p->nf <= 1024 && p->nf % 4 == 0,
              EVE_ERR_SHAPE, "eyenet_tail: unsupported shape batch=%d steps=%d nf=%d", p->batch,
              p->steps, p->nf);
  EVE_REQUIRE(p->rnn_type >= EVE_RNN_NONE && p->rnn_type <= EVE_RNN_GRU, EVE_ERR_CONFIG,
              "Unknown RNN type for EyeNet: %d", p->rnn_type);
  EVE_REQUIRE(p->rnn_type == EVE_RNN_NONE || (p->rnn_cells >= 1 && p->rnn_cells <= kMaxCells),
              EVE_ERR_CONFIG, "eyenet_tail: rnn_cells=%d unsupported (1..%d)", p->rnn_cells,
              kMaxCells);
  return EVE_OK;
}

inline int tail_head_base(const eve_eyenet_tail_params& p) {
  return 4 + (p.rnn_type == EVE_RNN_NONE ? 2 : 4 * p.rnn_cells);
}

size_t tail_ws_bytes(const eve_eyenet_tail_params& p) {
  const size_t R = (size_t)p.batch * p.steps;
  const size_t wide = (size_t)5 * p.nf + 8;
  size_t floats = 16 * R * wide                               // row buffers
                  + 4 * wide * (size_t)(p.nf + 8)              // transposed weights
                  + ((R + 255) / 256 + 1) * wide * (size_t)(p.nf + 8)   // split-K partials
                  + 4096 * wide;
  return floats * sizeof(float) + 64 * 256;
}

template <int TYPE>
int launch_rnn_fwd(const float* gi, const float* whT, const float* bhh, const float* h0,
                   const float* c0, int B, int T, int nf, float* H, float* Hprev, float* gates,
                   float* hT, float* cT, cudaStream_t s) {
  rnn_seq_fwd_kernel<TYPE><<<B, nf, nf * sizeof(float), s>>>(gi, whT, bhh, h0, c0, T, nf, H, Hprev,
                                                            gates, hT, cT);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
template <int TYPE>
int launch_rnn_bwd(const float* dHext, const float* dhT, const float* dcT, const float* whh,
                   const float* H, const float* Hprev, const float* gates, const float* c0, int B,
                   int T, int nf, int G, float* dgi, float* dgh, float* dh0, float* dc0,
                   cudaStream_t s) {
  rnn_seq_bwd_kernel<TYPE><<<B, nf, G * nf * sizeof(float), s>>>(
      dHext, dhT, dcT, whh, H, Hprev, gates, c0, T, nf, dgi, dgh, dh0, dc0);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

}  // namespace
}  // namespace eve

extern "C" int eve_eyenet_tail_num_weights(const eve_eyenet_tail_params* p) {
  if (check_tail(p) != EVE_OK) return -1;
  return tail_head_base(*p) + 7;
}

extern "C" size_t eve_eyenet_tail_saved_bytes(const eve_eyenet_tail_params* p) {
  if (check_tail(p) != EVE_OK) return 0;
  Arena sv(nullptr, 0);
  TailTape t;
  build_tail_tape(*p, sv, t);
  return sv.off + 256;
}

extern "C" size_t eve_eyenet_tail_workspace_bytes(const eve_eyenet_tail_params* p) {
  if (check_tail(p) != EVE_OK) return 0;
  return tail_ws_bytes(*p);
}

extern "C" int eve_eyenet_tail_fwd(const eve_eyenet_tail_params* p, const float* feat,
                                   const float* head_pose, const float* h0, const float* c0,
                                   const float* const* w, float* g, float* pupil, float* hT,
                                   float* cT, void* saved, size_t saved_bytes, void* workspace,
                                   size_t workspace_bytes, eve_stream_t stream) {
  EVE_TRY(check_tail(p));
  if (p->batch == 0 || p->steps == 0) return EVE_OK;
  EVE_REQUIRE(feat && w && g && pupil && saved && workspace, EVE_ERR_NULL,
              "eyenet_tail_fwd: NULL pointer");
  EVE_REQUIRE(!p->use_head_pose || head_pose, EVE_ERR_NULL, "eyenet_tail_fwd: head_pose is NULL");
  cudaStream_t s = as_stream(stream);
  Arena sv(saved, saved_bytes);
  TailTape t;
  EVE_REQUIRE(build_tail_tape(*p, sv, t), EVE_ERR_WORKSPACE,
              "eyenet_tail_fwd: saved buffer too small");
  EVE_REQUIRE(workspace_bytes >= tail_ws_bytes(*p), EVE_ERR_WORKSPACE,
              "eyenet_tail_fwd: workspace too small");
  Arena ws(workspace, workspace_bytes);
  const int R = t.R, nf = t.nf, B = p->batch, T = p->steps;
  float* wt = ws.get<float>((size_t)(5 * nf) * (nf + 8));
  float* gi = ws.get<float>((size_t)R * 5 * nf);

  EVE_TRY(copy_channels(feat, R, nf, nf, 0, t.xin, t.in0, 0, false, s));
  if (p->use_head_pose) EVE_TRY(copy_channels(head_pose, R, 2, 2, 0, t.xin, t.in0, nf, false, s));
  EVE_TRY(linear_fwd(t.xin, R, t.in0, w[0], w[1], nf, t.u1, wt, s));
  EVE_TRY(ew_fwd(EW_SELU, t.u1, (long long)R * nf, t.f1, s));
  EVE_TRY(linear_fwd(t.f1, R, nf, w[2], w[3], nf, t.f2, wt, s));
  if (p->rnn_type != EVE_RNN_NONE) {
    const float* xin = t.f2;
    const int G = t.G;
    for (int i = 0; i < p->rnn_cells; ++i) {
      const float* const* cw = w + 4 + 4 * i;
      EVE_TRY(linear_fwd(xin, R, nf, cw[0], cw[2], G * nf, gi, wt, s));
      // transposed weight_hh: [nf][G*nf]
      ConvGeom tg = make_conv(1, 1, 1, nf, G * nf, 1, 1, 0);
      EVE_TRY(conv_prep_weights(tg, cw[1], wt, nullptr, s));
      const float* h0i = h0 ? h0 + (size_t)i * B * nf : nullptr;
      const float* c0i = c0 ? c0 + (size_t)i * B * nf : nullptr;
      float* hTi = hT ? hT + (size_t)i * B * nf : nullptr;
      float* cTi = cT ? cT + (size_t)i * B * nf : nullptr;
      if (t.c0[i]) {  // LSTM backward needs c_{-1}
        if (c0i)
          EVE_CUDA(cudaMemcpyAsync(t.c0[i], c0i, (size_t)B * nf * sizeof(float),
                                   cudaMemcpyDeviceToDevice, s));
        else
          EVE_TRY(fill_zero(t.c0[i], (long long)B * nf, s));
      }
      if (p->rnn_type == EVE_RNN_GRU)
        EVE_TRY(launch_rnn_fwd<EVE_RNN_GRU>(gi, wt, cw[3], h0i, c0i, B, T, nf, t.H[i], t.Hprev[i],
                                            t.gates[i], hTi, cTi, s));
      else if (p->rnn_type == EVE_RNN_LSTM)
        EVE_TRY(launch_rnn_fwd<EVE_RNN_LSTM>(gi, wt, cw[3], h0i, c0i, B, T, nf, t.H[i], t.Hprev[i],
                                             t.gates[i], hTi, cTi, s));
      else
        EVE_TRY(launch_rnn_fwd<EVE_RNN_RNN>(gi, wt, cw[3], h0i, c0i, B, T, nf, t.H[i], t.Hprev[i],
                                            t.gates[i], hTi, cTi, s));
      xin = t.H[i];
    }
  } else {
    EVE_TRY(linear_fwd(t.f2, R, nf, w[4], w[5], nf, t.u3, wt, s));
    EVE_TRY(ew_fwd(EW_SELU, t.u3, (long long)R * nf, t.f3, s));
  }
  const float* const* hw = w + tail_head_base(*p);
  const float* f = t.feat_final;
  EVE_TRY(linear_fwd(f, R, nf, hw[0], hw[1], nf, t.ug, wt, s));
  EVE_TRY(ew_fwd(EW_SELU, t.ug, (long long)R * nf, t.sg, s));
  float* vg = gi;  // reuse
  EVE_TRY(linear_fwd(t.sg, R, nf, hw[2], nullptr, 2, vg, wt, s));
  EVE_TRY(ew_fwd(EW_TANH_HALFPI, vg, (long long)R * 2, t.gout, s));
  EVE_CUDA(cudaMemcpyAsync(g, t.gout, (size_t)R * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  EVE_TRY(linear_fwd(f, R, nf, hw[3], hw[4], nf, t.up, wt, s));
  EVE_TRY(ew_fwd(EW_SELU, t.up, (long long)R * nf, t.sp, s));
  EVE_TRY(linear_fwd(t.sp, R, nf, hw[5], hw[6], 1, t.vp, wt, s));
  EVE_TRY(ew_fwd(EW_RELU, t.vp, R, pupil, s));
  return EVE_OK;
}

extern "C" int eve_eyenet_tail_bwd(const eve_eyenet_tail_params* p, const float* dg,
                                   const float* dpupil, const float* dhT, const float* dcT,
                                   const float* const* w, float* dfeat, float* dh0, float* dc0,
                                   float* const* gr, int accumulate, const void* saved,
                                   size_t saved_bytes, void* workspace, size_t workspace_bytes,
                                   eve_stream_t stream) {
  EVE_TRY(check_tail(p));
  if (p->batch == 0 || p->steps == 0) return EVE_OK;
  EVE_REQUIRE(dg && dpupil && w && gr && dfeat && saved && workspace, EVE_ERR_NULL,
              "eyenet_tail_bwd: NULL pointer");
  cudaStream_t s = as_stream(stream);
  Arena sv(const_cast<void*>(saved), saved_bytes);
  TailTape t;
  EVE_REQUIRE(build_tail_tape(*p, sv, t), EVE_ERR_WORKSPACE,
              "eyenet_tail_bwd: saved buffer too small");
  EVE_REQUIRE(workspace_bytes >= tail_ws_bytes(*p), EVE_ERR_WORKSPACE,
              "eyenet_tail_bwd: workspace too small");
  Arena ws(workspace, workspace_bytes);
  const int R = t.R, nf = t.nf, B = p->batch, T = p->steps, G = t.G;
  const bool acc = accumulate != 0;
  const size_t wide = (size_t)5 * nf + 8;
  float* sc = ws.get<float>(((size_t)(R + 255) / 256 + 1) * wide * (nf + 8) + 4096 * wide);
  float* b0 = ws.get<float>((size_t)R * wide);
  float* b1 = ws.get<float>((size_t)R * wide);
  float* b2 = ws.get<float>((size_t)R * wide);
  float* b3 = ws.get<float>((size_t)R * wide);
  float* df = ws.get<float>((size_t)R * nf);
  EVE_REQUIRE(ws.ok(), EVE_ERR_WORKSPACE, "eyenet_tail_bwd: workspace too small");
  const int hb = tail_head_base(*p);
  const float* const* hw = w + hb;
  float* const* hg = gr + hb;
  const float* f = t.feat_final;

  // gaze head
  float* dvg = b0;
  EVE_TRY(ew_bwd(EW_TANH_HALFPI, dg, t.gout, (long long)R * 2, dvg, s));
  EVE_TRY(linear_wgrad(t.sg, dvg, R, nf, 2, hg[2], nullptr, sc, acc, s));
  float* dsg = b1;
  EVE_TRY(linear_dgrad(dvg, R, 2, hw[2], nf, nullptr, dsg, s));
  float* dug = b2;
  EVE_TRY(ew_bwd(EW_SELU, dsg, t.ug, (long long)R * nf, dug, s));
  EVE_TRY(linear_wgrad(f, dug, R, nf, nf, hg[0], hg[1], sc, acc, s));
  EVE_TRY(linear_dgrad(dug, R, nf, hw[0], nf, nullptr, df, s));
  // pupil head
  float* dvp = b0;
  EVE_TRY(ew_bwd(EW_RELU, dpupil, t.vp, R, dvp, s));
  EVE_TRY(linear_wgrad(t.sp, dvp, R, nf, 1, hg[5], hg[6], sc, acc, s));
  float* dsp = b1;
  EVE_TRY(linear_dgrad(dvp, R, 1, hw[5], nf, nullptr, dsp, s));
  float* dup = b2;
  EVE_TRY(ew_bwd(EW_SELU, dsp, t.up, (long long)R * nf, dup, s));
  EVE_TRY(linear_wgrad(f, dup, R, nf, nf, hg[3], hg[4], sc, acc, s));
  EVE_TRY(linear_dgrad(dup, R, nf, hw[3], nf, df, df, s));

  float* df2 = b3;  // gradient w.r.t. f2
  if (p->rnn_type != EVE_RNN_NONE) {
    float* dHext = df;
    for (int i = p->rnn_cells - 1; i >= 0; --i) {
      const float* const* cw = w + 4 + 4 * i;
      float* const* cg = gr + 4 + 4 * i;
      const float* xin = i == 0 ? t.f2 : t.H[i - 1];
      float* dgi = b0;
      float* dgh = p->rnn_type == EVE_RNN_GRU ? b1 : b0;
      const float* dhTi = dhT ? dhT + (size_t)i * B * nf : nullptr;
      const float* dcTi = dcT ? dcT + (size_t)i * B * nf : nullptr;
      float* dh0i = dh0 ? dh0 + (size_t)i * B * nf : nullptr;
      float* dc0i = dc0 ? dc0 + (size_t)i * B * nf : nullptr;
      const float* c0i = t.c0[i];
      if (p->rnn_type == EVE_RNN_GRU)
        EVE_TRY(launch_rnn_bwd<EVE_RNN_GRU>(dHext, dhTi, dcTi, cw[1], t.H[i], t.Hprev[i],
                                            t.gates[i], c0i, B, T, nf, G, dgi, dgh, dh0i, dc0i, s));
      else if (p->rnn_type == EVE_RNN_LSTM)
        EVE_TRY(launch_rnn_bwd<EVE_RNN_LSTM>(dHext, dhTi, dcTi, cw[1], t.H[i], t.Hprev[i],
                                             t.gates[i], c0i, B, T, nf, G, dgi, dgh, dh0i, dc0i,
                                             s));
      else
        EVE_TRY(launch_rnn_bwd<EVE_RNN_RNN>(dHext, dhTi, dcTi, cw[1], t.H[i], t.Hprev[i],
                                            t.gates[i], c0i, B, T, nf, G, dgi, dgh, dh0i, dc0i, s));
      EVE_TRY(linear_wgrad(xin, dgi, R, nf, G * nf, cg[0], cg[2], sc, acc, s));
      EVE_TRY(linear_wgrad(t.Hprev[i], dgh, R, nf, G * nf, cg[1], cg[3], sc, acc, s));
      float* dx = (i == 0) ? df2 : (dHext == df ? b2 : df);
      EVE_TRY(linear_dgrad(dgi, R, G * nf, cw[0], nf, nullptr, dx, s));
      dHext = dx;
    }
  } else {
    float* du3 = b0;
    EVE_TRY(ew_bwd(EW_SELU, df, t.u3, (long long)R * nf, du3, s));
    EVE_TRY(linear_wgrad(t.f2, du3, R, nf, nf, gr[4], gr[5], sc, acc, s));
    EVE_TRY(linear_dgrad(du3, R, nf, w[4], nf, nullptr, df2, s));
  }
  // fc_common
  EVE_TRY(linear_wgrad(t.f1, df2, R, nf, nf, gr[2], gr[3], sc, acc, s));
  float* df1 = b0;
  EVE_TRY(linear_dgrad(df2, R, nf, w[2], nf, nullptr, df1, s));
  float* du1 = b1;
  EVE_TRY(ew_bwd(EW_SELU, df1, t.u1, (long long)R * nf, du1, s));
  EVE_TRY(linear_wgrad(t.xin, du1, R, t.in0, nf, gr[0], gr[1], sc, acc, s));
  float* dxin = b2;
  EVE_TRY(linear_dgrad(du1, R, nf, w[0], t.in0, nullptr, dxin, s));
  EVE_TRY(copy_channels(dxin, R, nf, t.in0, 0, dfeat, nf, 0, false, s));
  return EVE_OK;
}
