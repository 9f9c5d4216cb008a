// C ABI: error reporting and the single-op entry points (see include/eve_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace eve {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int ensure_dynamic_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  if (bytes <= 48 * 1024) return EVE_OK;
  int dev = 0;
  EVE_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = done[std::make_pair(kernel, dev)];
  if (bytes > cur) {
    EVE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
  }
  return EVE_OK;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- profiler state (host side only)
struct ProfRec {
  int kind;
  double flops, bytes;
  cudaEvent_t a, b;
  int geom[8];        // N, H, W, Cin, Cout, k, stride, tagged (0 when the caller gave no geometry)
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int kind, double flops, double bytes, cudaStream_t stream, const ConvGeom* g)
    : slot(-1), s(stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{kind, flops, bytes, prof_event(), prof_event(), {0, 0, 0, 0, 0, 0, 0, 0}};
  if (g) {
    const int t[8] = {g->N, g->H, g->W, g->Cin, g->Cout, g->KH, g->stride, 1};
    for (int i = 0; i < 8; ++i) r.geom[i] = t[i];
  }
  if (!r.a || !r.b) return;
  cudaEventRecord(r.a, s);
  g_prof_recs.push_back(r);
  slot = (int)g_prof_recs.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[slot].b, s);
}

static int check_conv(const eve_conv_params* p, ConvGeom& g) {
  EVE_REQUIRE(p, EVE_ERR_NULL, "conv2d: params is NULL");
  EVE_REQUIRE(p->n >= 0 && p->h > 0 && p->w > 0 && p->cin > 0 && p->cout > 0 && p->ksize > 0 &&
                  p->stride > 0 && p->pad >= 0 && p->h + 2 * p->pad >= p->ksize &&
                  p->w + 2 * p->pad >= p->ksize,
              EVE_ERR_SHAPE, "conv2d: bad geometry n=%d h=%d w=%d cin=%d cout=%d k=%d s=%d p=%d",
              p->n, p->h, p->w, p->cin, p->cout, p->ksize, p->stride, p->pad);
  g = make_conv(p->n, p->h, p->w, p->cin, p->cout, p->ksize, p->stride, p->pad);
  return EVE_OK;
}

static size_t conv_ws_bytes(const ConvGeom& g) {
  return conv_scratch_bytes(conv_operand_elems(g), (size_t)g.out_elems(), (size_t)g.Cout * g.K(),
                            conv_partial_floats(g));
}

}  // namespace eve

using namespace eve;

extern "C" int eve_version(void) { return 100; }
extern "C" long long eve_launch_count(void) { return g_launches.load(); }

extern "C" void eve_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
}

extern "C" void eve_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) {
    g_prof_pool.push_back(r.a);
    g_prof_pool.push_back(r.b);
  }
  g_prof_recs.clear();
}

extern "C" int eve_profile_read(int kind, double* ms, double* flops, double* bytes,
                                long long* launches) {
  EVE_REQUIRE(kind >= 0 && kind < PROF_KINDS, EVE_ERR_SHAPE, "profile_read: bad kind %d", kind);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double t = 0.0, f = 0.0, b = 0.0;
  long long n = 0;
  for (auto& r : g_prof_recs) {
    if (r.kind != kind) continue;
    EVE_CUDA(cudaEventSynchronize(r.b));
    float e = 0.f;
    EVE_CUDA(cudaEventElapsedTime(&e, r.a, r.b));
    t += e;
    f += r.flops;
    b += r.bytes;
    ++n;
  }
  if (ms) *ms = t;
  if (flops) *flops = f;
  if (bytes) *bytes = b;
  if (launches) *launches = n;
  return EVE_OK;
}
// One text line per recorded launch: "kind N H W Cin Cout k stride ms flops bytes"; returns the
// number of bytes the full listing needs (call with cap = 0 to size the buffer).
extern "C" long long eve_profile_dump(char* buf, long long cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  long long off = 0;
  for (auto& r : g_prof_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float e = 0.f;
    if (cudaEventElapsedTime(&e, r.a, r.b) != cudaSuccess) continue;
    char line[192];
    int n = snprintf(line, sizeof(line), "%d %d %d %d %d %d %d %d %.6f %.0f %.0f\n", r.kind,
                     r.geom[0], r.geom[1], r.geom[2], r.geom[3], r.geom[4], r.geom[5], r.geom[6],
                     (double)e, r.flops, r.bytes);
    if (buf && off + n < cap) memcpy(buf + off, line, (size_t)n);
    off += n;
  }
  if (buf && cap > 0) buf[off < cap ? off : cap - 1] = 0;
  return off + 1;
}
extern "C" const char* eve_last_error(void) { return g_error; }

extern "C" int eve_conv2d_describe(const eve_conv_params* p, char* buf, size_t cap) {
  ConvGeom g;
  EVE_TRY(check_conv(p, g));
  EVE_REQUIRE(buf && cap > 0, EVE_ERR_NULL, "conv2d_describe: NULL buffer");
  conv_tc_describe(g, buf, cap);
  return EVE_OK;
}

extern "C" size_t eve_conv2d_workspace_bytes(const eve_conv_params* p) {
  ConvGeom g;
  if (check_conv(p, g) != EVE_OK) return 0;
  return conv_ws_bytes(g);
}

// ---- tuning options (process-wide; see include/eve_b200.h)
namespace eve {
struct Opt {
  const char* name;
  int value, lo, hi;
  const char* env;
  bool env_read;
};
static Opt g_opts[OPT_COUNT] = {
    {"tc_stage_cap", 24, 2, 24, "EVE_B200_TC_STAGE_CAP", false},
    {"tc_row_kernel", 1, 0, 1, "EVE_B200_TC_ROW_KERNEL", false},
    {"tc_row_strips", 0, 0, 128, "EVE_B200_TC_ROW_STRIPS", false},
    {"tc_row_wgrad", 2, 0, 2, "EVE_B200_TC_ROW_WGRAD", false},
    {"tc_wgrad_waves", 3, 1, 8, "EVE_B200_TC_WGRAD_WAVES", false},
    {"fused_planes", 1, 0, 1, "EVE_B200_FUSED_PLANES", false},
    {"fused_norm", 1, 0, 1, "EVE_B200_FUSED_NORM", false},
    {"tc_strip", 1, 0, 2, "EVE_B200_TC_STRIP", false},
    {"tc_wgrad_strip", 1, 0, 1, "EVE_B200_TC_WGRAD_STRIP", false},
    {"cgru_persistent", 1, 0, 1, "EVE_B200_CGRU_PERSISTENT", false},
    {"tc_dual", 0, 0, 1, "EVE_B200_TC_DUAL", false},
    {"tc_pair", 0, 0, 2, "EVE_B200_TC_PAIR", false},
    {"stem_windows", 1, 0, 2, "EVE_B200_STEM_WINDOWS", false},
    {"in_stream", 2, 0, 2, "EVE_B200_IN_STREAM", false},
    {"stem_fused_bwd", 1, 0, 1, "EVE_B200_STEM_FUSED_BWD", false},
};
int get_option(int key) {
  if (key < 0 || key >= OPT_COUNT) return 0;
  Opt& o = g_opts[key];
  if (!o.env_read) {
    o.env_read = true;
    const char* e = getenv(o.env);
    if (e) {
      int v = atoi(e);
      if (v >= o.lo && v <= o.hi) o.value = v;
    }
  }
  return o.value;
}
}  // namespace eve

extern "C" int eve_set_option(const char* name, int value) {
  EVE_REQUIRE(name, EVE_ERR_NULL, "eve_set_option: name is NULL");
  for (int k = 0; k < OPT_COUNT; ++k)
    if (strcmp(g_opts[k].name, name) == 0) {
      EVE_REQUIRE(value >= g_opts[k].lo && value <= g_opts[k].hi, EVE_ERR_CONFIG,
                  "eve_set_option: %s=%d outside [%d, %d]", name, value, g_opts[k].lo, g_opts[k].hi);
      g_opts[k].env_read = true;
      g_opts[k].value = value;
      return EVE_OK;
    }
  EVE_REQUIRE(false, EVE_ERR_CONFIG, "eve_set_option: unknown option '%s'", name);
  return EVE_ERR_CONFIG;
}
extern "C" int eve_get_option(const char* name, int* value) {
  EVE_REQUIRE(name && value, EVE_ERR_NULL, "eve_get_option: NULL argument");
  for (int k = 0; k < OPT_COUNT; ++k)
    if (strcmp(g_opts[k].name, name) == 0) {
      *value = get_option(k);
      return EVE_OK;
    }
  EVE_REQUIRE(false, EVE_ERR_CONFIG, "eve_get_option: unknown option '%s'", name);
  return EVE_ERR_CONFIG;
}

extern "C" void eve_set_conv_mode(int mode) { set_conv_mode(mode); }
extern "C" int eve_get_conv_mode(void) { return conv_mode(); }

extern "C" int eve_conv2d_fwd(const eve_conv_params* p, const float* x, const float* w,
                              const float* bias, float* y, void* workspace,
                              size_t workspace_bytes, eve_stream_t stream) {
  ConvGeom g;
  EVE_TRY(check_conv(p, g));
  if (g.N == 0) return EVE_OK;
  EVE_REQUIRE(x && w && y && workspace, EVE_ERR_NULL, "conv2d_fwd: NULL pointer");
  EVE_REQUIRE(workspace_bytes >= conv_ws_bytes(g), EVE_ERR_WORKSPACE,
              "conv2d_fwd: workspace too small");
  ConvScratch sc{(char*)workspace, workspace_bytes};
  return conv_fwd(g, x, w, bias, nullptr, y, sc, as_stream(stream));
}

extern "C" int eve_conv2d_dgrad(const eve_conv_params* p, const float* dy, const float* w,
                                float* dx, void* workspace, size_t workspace_bytes,
                                eve_stream_t stream) {
  ConvGeom g;
  EVE_TRY(check_conv(p, g));
  if (g.N == 0) return EVE_OK;
  EVE_REQUIRE(dy && w && dx && workspace, EVE_ERR_NULL, "conv2d_dgrad: NULL pointer");
  EVE_REQUIRE(workspace_bytes >= conv_ws_bytes(g), EVE_ERR_WORKSPACE,
              "conv2d_dgrad: workspace too small");
  ConvScratch sc{(char*)workspace, workspace_bytes};
  return conv_dgrad(g, dy, w, nullptr, dx, sc, as_stream(stream));
}

extern "C" int eve_conv2d_wgrad(const eve_conv_params* p, const float* x, const float* dy,
                                float* dw, float* dbias, void* workspace, size_t workspace_bytes,
                                eve_stream_t stream) {
  ConvGeom g;
  EVE_TRY(check_conv(p, g));
  EVE_REQUIRE(dw && workspace && (g.N == 0 || (x && dy)), EVE_ERR_NULL,
              "conv2d_wgrad: NULL pointer");
  EVE_REQUIRE(workspace_bytes >= conv_ws_bytes(g), EVE_ERR_WORKSPACE,
              "conv2d_wgrad: workspace too small");
  cudaStream_t s = as_stream(stream);
  if (g.N == 0) {
    EVE_TRY(fill_zero(dw, (long long)g.Cout * g.K(), s));
    if (dbias) EVE_TRY(fill_zero(dbias, g.Cout, s));
    return EVE_OK;
  }
  ConvScratch sc{(char*)workspace, workspace_bytes};
  return conv_wgrad(g, x, dy, dw, dbias, false, sc, s);
}

extern "C" int eve_instnorm_act_fwd(const float* x, int n, int hw, int c, const float* gamma,
                                    const float* beta, int act, float* y, float* mean,
                                    float* rstd, eve_stream_t stream) {
  EVE_REQUIRE(x && y && mean && rstd, EVE_ERR_NULL, "instnorm_act_fwd: NULL pointer");
  EVE_REQUIRE(n >= 0 && hw > 0 && c > 0 && c % 4 == 0 && act >= 0 && act <= 2, EVE_ERR_SHAPE,
              "instnorm_act_fwd: bad shape n=%d hw=%d c=%d act=%d", n, hw, c, act);
  EVE_REQUIRE((gamma == nullptr) == (beta == nullptr), EVE_ERR_NULL,
              "instnorm_act_fwd: gamma and beta must both be given or both be NULL");
  if (n == 0) return EVE_OK;
  cudaStream_t s = as_stream(stream);
  EVE_TRY(in_stats(x, n, hw, c, mean, rstd, s));
  return in_apply(x, n, hw, c, mean, rstd, gamma, beta, nullptr, nullptr, nullptr, act, y, s);
}

extern "C" int eve_instnorm_act_bwd(const float* dy, const float* y, const float* x, int n, int hw,
                                    int c, const float* mean, const float* rstd,
                                    const float* gamma, int act, float* dx, float* dgamma,
                                    float* dbeta, void* workspace, size_t workspace_bytes,
                                    eve_stream_t stream) {
  EVE_REQUIRE(dy && x && mean && rstd && dx && workspace, EVE_ERR_NULL,
              "instnorm_act_bwd: NULL pointer");
  EVE_REQUIRE(act == ACT_NONE || y, EVE_ERR_NULL, "instnorm_act_bwd: y is NULL");
  EVE_REQUIRE(n >= 0 && hw > 0 && c > 0 && c % 4 == 0 && act >= 0 && act <= 2, EVE_ERR_SHAPE,
              "instnorm_act_bwd: bad shape n=%d hw=%d c=%d act=%d", n, hw, c, act);
  EVE_REQUIRE(workspace_bytes >= in_backward_scratch_floats(n, c) * sizeof(float),
              EVE_ERR_WORKSPACE, "instnorm_act_bwd: workspace too small");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    if (gamma && dgamma) {
      EVE_TRY(fill_zero(dgamma, c, s));
      EVE_TRY(fill_zero(dbeta, c, s));
    }
    return EVE_OK;
  }
  return in_backward(dy, y, x, n, hw, c, mean, rstd, gamma, nullptr, act, nullptr, dx, nullptr, dgamma,
                     dbeta, (float*)workspace, false, s);
}

extern "C" size_t eve_instnorm_fused_workspace_bytes(int n, int hw, int c) {
  return (in_bwd_fused_scratch_floats(n, hw, c) + 64) * sizeof(float);
}

extern "C" int eve_instnorm_fused_fwd(const float* x, const float* x2, int x2_mode, int n, int hw,
                                      int c, const float* gamma, const float* beta,
                                      const float* gamma_b, const float* beta_b, int act, int fmt,
                                      float* mean, float* rstd, float* mean2, float* rstd2,
                                      float* y, void* hi_a, void* lo_a, void* hi_b, void* lo_b,
                                      eve_stream_t stream) {
  EVE_REQUIRE(n >= 0 && hw > 0 && c > 0 && c % 4 == 0 && act >= 0 && act <= 2 && x2_mode >= 0 &&
                  x2_mode <= 2 && (fmt == TC_F16 || fmt == TC_BF16),
              EVE_ERR_SHAPE, "instnorm_fused_fwd: bad arguments n=%d hw=%d c=%d act=%d", n, hw, c, act);
  EVE_REQUIRE(in_fused_supported(hw, c, x2_mode == 2 ? 2 : 1), EVE_ERR_SHAPE,
              "instnorm_fused_fwd: hw=%d c=%d does not fit the cluster kernel", hw, c);
  return in_fwd_fused(x, n, hw, c, x2, x2_mode, gamma, beta, gamma_b, beta_b, act, fmt, mean, rstd,
                      mean2, rstd2, y, hi_a, lo_a, hi_b, lo_b, as_stream(stream));
}

extern "C" int eve_instnorm_fused_bwd(const float* dy, const float* dy2, const float* ymask,
                                      const float* x, int n, int hw, int c, const float* mean,
                                      const float* rstd, const float* gamma, const float* beta,
                                      const float* gamma2, const float* beta2, int act,
                                      const float* addend, float* dx, void* dx_hi, void* dx_lo,
                                      float* g_out, float* dgamma, float* dbeta, float* dgamma2,
                                      float* dbeta2, float* dbias, void* workspace,
                                      size_t workspace_bytes, eve_stream_t stream) {
  EVE_REQUIRE(n >= 0 && hw > 0 && c > 0 && c % 4 == 0 && act >= 0 && act <= 2, EVE_ERR_SHAPE,
              "instnorm_fused_bwd: bad arguments n=%d hw=%d c=%d act=%d", n, hw, c, act);
  EVE_REQUIRE(in_fused_supported(hw, c, 2), EVE_ERR_SHAPE,
              "instnorm_fused_bwd: hw=%d c=%d does not fit the cluster kernel", hw, c);
  EVE_REQUIRE(workspace && workspace_bytes >= eve_instnorm_fused_workspace_bytes(n, hw, c),
              EVE_ERR_WORKSPACE, "instnorm_fused_bwd: workspace too small");
  return in_bwd_fused(dy, dy2, ymask, x, n, hw, c, mean, rstd, gamma, beta, gamma2, beta2, act,
                      addend, dx, dx_hi, dx_lo, g_out, dgamma, dbeta, dgamma2, dbeta2, dbias, nullptr,
                      false, (float*)workspace, as_stream(stream));
}

extern "C" int eve_in_relu_maxpool_fwd(const float* x, int n, int h, int w, int c, float* mean,
                                       float* rstd, float* y, int32_t* idx, eve_stream_t stream) {
  EVE_REQUIRE(x && mean && rstd && y && idx, EVE_ERR_NULL, "in_relu_maxpool_fwd: NULL pointer");
  EVE_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, EVE_ERR_SHAPE,
              "in_relu_maxpool_fwd: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  if (n == 0) return EVE_OK;
  cudaStream_t s = as_stream(stream);
  EVE_TRY(in_stats(x, n, h * w, c, mean, rstd, s));
  return in_relu_maxpool(x, n, h, w, c, mean, rstd, y, idx, s);
}

extern "C" int eve_in_relu_maxpool_bwd(const float* dy, const float* y, const int32_t* idx,
                                       const float* x, int n, int h, int w, int c,
                                       const float* mean, const float* rstd, float* dx,
                                       uint16_t* dx_hi, uint16_t* dx_lo, float* scratch,
                                       eve_stream_t stream) {
  EVE_REQUIRE(dy && y && idx && x && mean && rstd && scratch && (dx || dx_hi), EVE_ERR_NULL,
              "in_relu_maxpool_bwd: NULL pointer");
  EVE_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0 && h % 2 == 0 && w % 2 == 0,
              EVE_ERR_SHAPE, "in_relu_maxpool_bwd: bad shape n=%d h=%d w=%d c=%d (h, w even)", n, h,
              w, c);
  if (n == 0) return EVE_OK;
  return stem_pool_in_backward(dy, y, idx, x, n, h, w, c, mean, rstd, dx, dx_hi, dx_lo, scratch,
                               as_stream(stream));
}

extern "C" int eve_adaptive_maxpool_fwd(const float* x, int n, int h, int w, int c, int oh, int ow,
                                        float* y, int32_t* idx, eve_stream_t stream) {
  EVE_REQUIRE(x && y && idx, EVE_ERR_NULL, "adaptive_maxpool_fwd: NULL pointer");
  EVE_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, EVE_ERR_SHAPE,
              "adaptive_maxpool_fwd: bad shape");
  if (n == 0) return EVE_OK;
  return adaptive_maxpool_fwd(x, n, h, w, c, oh, ow, y, idx, as_stream(stream));
}

extern "C" int eve_adaptive_maxpool_bwd(const float* dy, const int32_t* idx, int n, int h, int w,
                                        int c, int oh, int ow, float* dx, eve_stream_t stream) {
  EVE_REQUIRE(dy && idx && dx, EVE_ERR_NULL, "adaptive_maxpool_bwd: NULL pointer");
  EVE_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, EVE_ERR_SHAPE,
              "adaptive_maxpool_bwd: bad shape");
  if (n == 0) return EVE_OK;
  return adaptive_maxpool_bwd(dy, idx, n, h, w, c, oh, ow, dx, as_stream(stream));
}

extern "C" int eve_upsample_bilinear_fwd(const float* x, int n, int h, int w, int c, int oh,
                                         int ow, float* y, eve_stream_t stream) {
  EVE_REQUIRE(x && y, EVE_ERR_NULL, "upsample_bilinear_fwd: NULL pointer");
  EVE_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, EVE_ERR_SHAPE,
              "upsample_bilinear_fwd: bad shape");
  if (n == 0) return EVE_OK;
  return upsample_bilinear_fwd(x, n, h, w, c, oh, ow, y, c, 0, as_stream(stream));
}

extern "C" int eve_upsample_bilinear_bwd(const float* dy, int n, int h, int w, int c, int oh,
                                         int ow, float* dx, eve_stream_t stream) {
  EVE_REQUIRE(dy && dx, EVE_ERR_NULL, "upsample_bilinear_bwd: NULL pointer");
  EVE_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, EVE_ERR_SHAPE,
              "upsample_bilinear_bwd: bad shape");
  if (n == 0) return EVE_OK;
  return upsample_bilinear_bwd(dy, c, 0, n, h, w, c, oh, ow, dx, as_stream(stream));
}

extern "C" int eve_nchw_to_nhwc(const float* x, int n, int c, int h, int w, float* y,
                                eve_stream_t stream) {
  EVE_REQUIRE(x && y, EVE_ERR_NULL, "nchw_to_nhwc: NULL pointer");
  return nchw_to_nhwc(x, n, c, h, w, y, as_stream(stream));
}

extern "C" int eve_nhwc_to_nchw(const float* x, int n, int c, int h, int w, float* y,
                                eve_stream_t stream) {
  EVE_REQUIRE(x && y, EVE_ERR_NULL, "nhwc_to_nchw: NULL pointer");
  return nhwc_to_nchw(x, n, c, h, w, y, as_stream(stream));
}
