// Shared declarations for libeve_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "../../include/eve_b200.h"

namespace eve {

// Records a message retrievable through eve_last_error(); thread-local.
void set_error(const char* fmt, ...);

#define EVE_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::eve::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                    \
                       cudaGetErrorString(e__));                                        \
      return EVE_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// Every kernel launch goes through this check, which also feeds eve_launch_count().
#define EVE_LAUNCH_CHECK()             \
  do {                                 \
    ::eve::count_launch();             \
    EVE_CUDA(cudaGetLastError());      \
  } while (0)

#define EVE_TRY(expr)                                                                   \
  do {                                                                                  \
    int rc__ = (expr);                                                                  \
    if (rc__ != EVE_OK) return rc__;                                                    \
  } while (0)

#define EVE_REQUIRE(cond, code, ...)                                                    \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      ::eve::set_error(__VA_ARGS__);                                                    \
      return (code);                                                                    \
    }                                                                                   \
  } while (0)

void count_launch();
// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): sets it once per pair
// (and again if a larger value is asked for)
int ensure_dynamic_smem(const void* kernel, size_t bytes);

// Optional per-kernel-family device timing (eve_profile_*): CUDA events recorded on the
// launching stream around a launch, summed on read.  Costs nothing when disabled.
enum ProfKind { PROF_CONV_FWD = 0, PROF_CONV_DGRAD = 1, PROF_CONV_WGRAD = 2, PROF_KINDS = 3 };
struct ConvGeom;
struct ProfScope {
  int slot;
  cudaStream_t s;
  // `g` (optional) tags the record with the convolution's geometry for eve_profile_dump()
  ProfScope(int kind, double flops, double bytes, cudaStream_t stream, const ConvGeom* g = nullptr);
  ~ProfScope();
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kNumSMs = 148;

// Bump allocator over the caller-owned workspace (the library never allocates
// device memory itself; see include/eve_b200.h "Ownership").
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  bool dry;  // size-counting pass: hand out fake pointers
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), dry(p == nullptr) {}
  template <typename T>
  T* get(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    size_t at = off;
    off += bytes;
    if (dry) return (T*)(uintptr_t)(256 + at);
    if (off > cap) return nullptr;
    return (T*)(base + at);
  }
  bool ok() const { return dry || off <= cap; }
};

// --------------------------------------------------------------------- conv geometry --
struct ConvGeom {
  int N, H, W, Cin;     // input, NHWC
  int OH, OW, Cout;     // output, NHWC
  int KH, KW, stride, pad;
  long long in_elems() const { return (long long)N * H * W * Cin; }
  long long out_elems() const { return (long long)N * OH * OW * Cout; }
  int K() const { return KH * KW * Cin; }
};
inline ConvGeom make_conv(int N, int H, int W, int Cin, int Cout, int k, int stride, int pad) {
  ConvGeom g;
  g.N = N; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.KH = k; g.KW = k;
  g.stride = stride; g.pad = pad;
  g.OH = (H + 2 * pad - k) / stride + 1;
  g.OW = (W + 2 * pad - k) / stride + 1;
  return g;
}

// ------------------------------------------------------------------------ conv (SIMT) --
// Weight layouts derived from the OIHW fp32 master (caches in the workspace):
//   fwd  : wf[(r*KW+q)*Cin + ci][co]
//   dgrad: wd[(r*KW+q)*Cout + co][ci]
int conv_prep_weights(const ConvGeom& g, const float* w_oihw, float* wf, float* wd,
                      cudaStream_t s);
// y[N,OH,OW,ldo>=Cout] (+bias) (+addend, same layout as y)
int conv_fwd_simt(const ConvGeom& g, const float* x, const float* wf, const float* bias,
                  const float* addend, float* y, int ldo, cudaStream_t s);
// dx[N,H,W,Cin] = dgrad(dy) (+addend)
int conv_dgrad_simt(const ConvGeom& g, const float* dy, int lddy, const float* wd,
                    const float* addend, float* dx, cudaStream_t s);
// dw_oihw (+)= wgrad(x, dy); scratch must hold conv_wgrad_scratch_floats(g)
size_t conv_wgrad_scratch_floats(const ConvGeom& g);
int conv_wgrad_simt(const ConvGeom& g, const float* x, const float* dy, int lddy, float* dw_oihw,
                    float* scratch, bool accumulate, cudaStream_t s);
// db[c] (+)= sum over rows of dy[rows, ld] (first C columns)
int colsum(const float* dy, long long rows, int C, int ld, float* db, float* scratch,
           bool accumulate, cudaStream_t s);
size_t colsum_scratch_floats(long long rows, int C);

// --------------------------------------------------------------------- conv (tcgen05) --
bool conv_tc_supported(const ConvGeom& g);
// operand planes: fmt 1 = bf16 (fp32 exponent range: gradients), fmt 0 = fp16 (11-bit mantissa:
// forward activations, which the normalisations keep far below 65504)
enum TcFormat { TC_F16 = 0, TC_BF16 = 1 };
int split_planes(const float* x, long long n, void* hi, void* lo, int fmt, cudaStream_t s);
int conv_tc_prep_weights(const ConvGeom& g, const float* w_oihw, bool dgrad, void* hi, void* lo,
                         int fmt, float scale, cudaStream_t s);
int conv_tc_run(const ConvGeom& g, const void* x_hi, const void* x_lo, const void* w_hi,
                const void* w_lo, const float* bias, const float* addend, float* y, int npass,
                int fmt, float out_scale, cudaStream_t s);

int conv_tc_describe(const ConvGeom& g, char* buf, size_t cap);
// persistent ConvGRU forward (conv_tc.cu): one CTA per clip walks all T steps
bool cgru_seq_supported(int nf, int H, int W);
int cgru_seq_fwd(int B, int T, const void* w1h_hi, const void* w1h_lo, const void* w2h_hi,
                 const void* w2h_lo, const float* gx1, const float* gx2, const float* h0, float* r,
                 float* z, float* n, float* h, float* xh, float* cat2, float out_scale,
                 cudaStream_t s);
int cgru_seq_bwd(int B, int T, const void* wd2_hi, const void* wd2_lo, const void* wd1_hi,
                 const void* wd1_lo, const float* dout, float* dcarry, const float* r, const float* z,
                 const float* n, const float* h, const float* h0, float* dg1, float* dg2,
                 cudaStream_t s);
bool conv_tc_dgrad_s2_supported(const ConvGeom& g);
int conv_tc_dgrad_s2_run(const ConvGeom& g, const void* d_hi, const void* d_lo, const void* w_hi,
                         const void* w_lo, const float* addend, float* dx, int npass,
                         cudaStream_t s);
int conv_tc_stem_run(int N, int H, int W, const void* x_hi, const void* x_lo, int win_bytes,
                     int pitch_bytes, const void* w_hi, const void* w_lo, const float* bias, float* y,
                     int npass, int fmt, float out_scale, cudaStream_t s);
bool conv_tc_wgrad_supported(const ConvGeom& g);
size_t conv_tc_wgrad_partial_floats(const ConvGeom& g);
int conv_tc_wgrad_run(const ConvGeom& g, const void* d_hi, const void* d_lo, const void* x_hi,
                      const void* x_lo, float* part, int npass, int* splits_out, cudaStream_t s, int x_fmt = TC_BF16);
// dw_oihw (+)= sum over splits of part[z][co][(r*KW+q)*Cin+ci]
int wgrad_reduce(const float* part, int splits, const ConvGeom& g, float* dw_oihw, bool accumulate,
                 cudaStream_t s);

// ------------------------------------------------------------------- conv (dispatch) --
// The entry points the networks call.  They take the fp32 NHWC activations and the OIHW fp32
// master weights, derive whatever operand layouts the chosen kernel needs inside `scratch`
// (a slice of the caller's workspace) and launch either the tcgen05 kernel (conv_tc.cu) or
// the fp32 CUDA-core kernel (conv_simt.cu).
//   mode 0: fp32 CUDA cores everywhere (exact)      mode 1: tcgen05, split-bf16 x3 (default)
//   mode 2: tcgen05, single bf16 pass (fastest, ~1e-2 relative error)
int conv_mode();
void set_conv_mode(int mode);
// tuning options (eve_set_option / eve_get_option; every one has an EVE_B200_* env default)
enum OptKey {
  OPT_TC_STAGE_CAP = 0,      // max ring stages of the tcgen05 conv kernel
  OPT_TC_ROW_KERNEL,         // halo-row tcgen05 kernel for 3x3 stride-1 layers with W == 128
  OPT_TC_ROW_STRIPS,         // 0 = automatic; otherwise the number of row strips per image
  OPT_TC_ROW_WGRAD,          // halo-row weight-gradient kernel (W == 128)
  OPT_TC_WGRAD_WAVES,        // full waves of CTAs the split-K weight gradient is sized for
  OPT_FUSED_PLANES,          // InstanceNorm kernels emit the consuming convolution's operand planes
  OPT_FUSED_NORM,            // one-pass cluster InstanceNorm kernels + plane-to-plane block pipelines
  OPT_TC_STRIP,              // padded-strip tcgen05 kernel for 3x3 stride-1 layers: 0 off, 1 auto, 2 whenever it fits
  OPT_TC_WGRAD_STRIP,        // padded-strip weight-gradient kernel (3x3 stride 1, 64 output channels)
  OPT_CGRU_PERSISTENT,       // ConvGRU (64 features, 5x8 maps): whole sequence in one persistent kernel
  OPT_TC_DUAL,               // box kernel with two MMA-issuing warps (even / odd ring stages, two partial accumulators)
  OPT_TC_PAIR,               // box kernel as clusters of two CTAs that multicast the weight stages to each other
  OPT_STEM_WINDOWS,          // stem forward without an im2col matrix: 0 off, 1 one window per output column, 2 overlapping windows in the padded image
  OPT_IN_STREAM,             // InstanceNorm backward without shared-memory staging: 0 off, 1 maps that need one CTA per SM, 2 always (default)
  OPT_STEM_FUSED_BWD,        // EyeNet stem backward: max-pool gather inside the norm's backward, sums taken over pool windows, dy planes written directly
  OPT_COUNT
};
int get_option(int key);
inline int tc_stage_cap() { return get_option(OPT_TC_STAGE_CAP); }
struct ConvScratch {
  char* base;
  size_t bytes;
};
// bytes needed for any convolution with at most these many input / output / weight elements
size_t conv_scratch_bytes(size_t max_in_elems, size_t max_out_elems, size_t max_w_elems,
                          size_t wgrad_partial_floats);
// split-K partial / bias-reduction floats the weight gradient of `g` may need (any kernel)
size_t conv_partial_floats(const ConvGeom& g);
// largest operand (elements) any pass of `g` stages as 16-bit planes (covers the stem's
// materialised patch matrix)
size_t conv_operand_elems(const ConvGeom& g);
// x == nullptr: the caller has already written the input's operand planes where
// conv_x_planes_fwd() points (only when conv_x_fusable(g)).
int conv_fwd(const ConvGeom& g, const float* x, const float* w_oihw, const float* bias,
             const float* addend, float* y, const ConvScratch& sc, cudaStream_t s);
// Producer/consumer fusion of the operand split: true when conv_fwd AND conv_bwd of `g` both take
// the split tensor-core path, so that the producer of x may write its 16-bit planes straight into
// the scratch slice (fp16 planes for conv_fwd, bf16 planes for conv_bwd) and pass x = nullptr.
bool conv_x_fusable(const ConvGeom& g);
void conv_x_planes_fwd(const ConvGeom& g, const ConvScratch& sc, void** hi, void** lo);
void conv_x_planes_bwd(const ConvGeom& g, const ConvScratch& sc, void** hi, void** lo);
int conv_dgrad(const ConvGeom& g, const float* dy, const float* w_oihw, const float* addend,
               float* dx, const ConvScratch& sc, cudaStream_t s);
int conv_wgrad(const ConvGeom& g, const float* x, const float* dy, float* dw_oihw, float* dbias,
               bool accumulate, const ConvScratch& sc, cudaStream_t s);
// The stem's weight gradient from dy planes its producer wrote (stem_pool_in_backward):
// conv_wgrad_stem_planes() tells where they go inside `sc`, conv_wgrad_stem_run() consumes them.
bool conv_wgrad_stem_takes_planes(const ConvGeom& g);
int conv_wgrad_stem_planes(const ConvGeom& g, const ConvScratch& sc, uint16_t** d_hi,
                           uint16_t** d_lo);
int conv_wgrad_stem_run(const ConvGeom& g, const float* x, float* dw, bool accumulate,
                        const ConvScratch& sc, cudaStream_t s);
// Plane-to-plane entry points (split tensor-core path only; conv_x_fusable(g) must hold): the
// producers of x / dy have already written the 16-bit hi/lo NHWC planes (x: fp16 for the forward
// pass, bf16 for the backward pass; dy: bf16).  Bias gradients come from the producer of dy.
int conv_fwd_planes(const ConvGeom& g, const void* x_hi, const void* x_lo, const float* w_oihw,
                    const float* bias, const float* addend, float* y, const ConvScratch& sc,
                    cudaStream_t s);
int conv_bwd_planes(const ConvGeom& g, const void* x_hi, const void* x_lo, const void* d_hi,
                    const void* d_lo, const float* w_oihw, float* dw, bool accumulate,
                    const float* addend, float* dx, const ConvScratch& sc, cudaStream_t s);
// Weights of a convolution that is applied many times in a row (ConvRNN cells): prepared once into
// the top of the scratch slice; conv_fwd / stride-1 conv_dgrad calls with the same weight pointer
// then skip their own preparation until conv_prepared_clear().  *top_used accumulates the bytes
// taken from the top (start at 0).  A no-op when the tensor-core path would not be taken.
int conv_prepare_weights(const ConvGeom& g, const float* w_oihw, bool dgrad, const ConvScratch& sc,
                         size_t* top_used, cudaStream_t s);
void conv_prepared_clear();
int conv_prepared_mark();
void conv_prepared_truncate(int mark);
// All tensor-core weight layouts of a network pass from one launch (forward layouts, or the
// flipped data-gradient layouts); `region` must stay untouched until conv_prepared_clear().
struct ConvPrepReq {
  ConvGeom g;
  const float* w;
};
size_t conv_prepare_batch_bytes(const ConvPrepReq* reqs, int n);
int conv_prepare_batch(const ConvPrepReq* reqs, int n, bool dgrad, void* region, size_t bytes,
                       cudaStream_t s);
// both gradients of one convolution (dw/dbias and dx, each optional) from one split of dy
int conv_bwd(const ConvGeom& g, const float* x, const float* dy, const float* w_oihw, float* dw,
             float* dbias, bool accumulate, const float* addend, float* dx, const ConvScratch& sc,
             cudaStream_t s);

// ------------------------------------------------------------------------------ norms --
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

// Per-(n,c) mean / rstd over HW of x[N,HW,C] (biased variance, eps 1e-5).
int in_stats(const float* x, int N, int HW, int C, float* mean, float* rstd, cudaStream_t s);
// y = act( (x-mean)*rstd*gamma + beta  +  residual ), residual optional; if res_mean != null
// the residual is itself normalised on the fly with (res_mean, res_rstd).  gamma/beta may be
// null (non-affine).  y may be written with row stride ldy and channel offset (concat).
int in_apply(const float* x, int N, int HW, int C, const float* mean, const float* rstd,
             const float* gamma, const float* beta, const float* res, const float* res_mean,
             const float* res_rstd, int act, float* y, cudaStream_t s);
// y = act(IN(x)*gamma+beta) written as the 16-bit hi/lo operand planes (fmt: TcFormat) of the
// convolution that consumes it, and as fp32 when y != null
int in_apply_planes(const float* x, int N, int HW, int C, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, int act, int fmt, float* y, void* hi,
                    void* lo, cudaStream_t s);
// act(IN(x)) feeding convolution `g`: as fp32 `y` (forward; conv_* then split it themselves) or,
// when conv_x_fusable(g), straight into the convolution's 16-bit operand planes inside its scratch
// slice (*fused = true: pass x = nullptr to conv_fwd / conv_bwd).  backward = true re-derives the
// bf16 planes conv_bwd needs from the saved pre-norm tensor (nothing to do when not fusable: the
// forward pass kept y).
int norm_act_into_conv(const ConvGeom& g, bool backward, const float* x, int N, int HW, int C,
                       const float* mean, const float* rstd, const float* gamma, const float* beta,
                       int act, float* y, const ConvScratch& cs, bool* fused, cudaStream_t s);
// in_apply_planes with two affine sets from one read of x (RefineNet blocks with a skip
// convolution normalise the same tensor twice, refine_net.py:45-62); set B optional
int in_apply_planes2(const float* x, int N, int HW, int C, const float* mean, const float* rstd,
                     const float* gamma, const float* beta, const float* gammaB,
                     const float* betaB, int act, int fmt, void* hiA, void* loA, void* hiB,
                     void* loB, cudaStream_t s);
// ---- one-pass cluster kernels (in_fused.cu)
bool in_fused_supported(int HW, int C, int tensors);
// statistics + normalisation from ONE read of x: mean/rstd out, y = act(IN(x)*gamma+beta (+x2)) as
// fp32 (y, optional) and as operand planes A (optional); planes B = act(IN(x)*gammaB+betaB)
// (optional).  x2_mode: 0 none, 1 x2 is a residual added before the activation, 2 x2 is itself
// instance-normalised (non-affine; mean2/rstd2 out) and added.
int in_fwd_fused(const float* x, int N, int HW, int C, const float* x2, int x2_mode,
                 const float* gamma, const float* beta, const float* gammaB, const float* betaB,
                 int act, int fmt, float* mean, float* rstd, float* mean2, float* rstd2, float* y,
                 void* hiA, void* loA, void* hiB, void* loB, cudaStream_t s);
// reductions + dx from ONE read of (dy, x); see in_fused.cu.  Outputs (each optional): dx fp32,
// dx as bf16 planes, g_out = dy*act', dgamma/dbeta (and the second affine set's), dbias = column
// sums of dx (the bias gradient of the convolution that produced x).
size_t in_bwd_fused_scratch_floats(int N, int HW, int C);
int in_bwd_fused(const float* dy, const float* dy2, const float* ymask, const float* x, int N,
                 int HW, int C, const float* mean, const float* rstd, const float* gamma,
                 const float* beta, const float* gamma2, const float* beta2, int act,
                 const float* addend, float* dx, void* dx_hi, void* dx_lo, float* g_out,
                 float* dgamma, float* dbeta, float* dgamma2, float* dbeta2, float* dbias,
                 float* dbias2, bool accumulate, float* scratch, cudaStream_t s,
                 void* ya_hi = nullptr, void* ya_lo = nullptr, void* yb_hi = nullptr,
                 void* yb_lo = nullptr);
// dy -> bf16 hi/lo planes and (dbias != null) dbias (+)= column sums, one read of dy
size_t split_colsum_scratch_floats(long long rows, int C);
int split_colsum(const float* dy, long long rows, int C, void* hi, void* lo, float* dbias,
                 float* dbias2, bool accumulate, float* scratch, cudaStream_t s);
// Backward of y = act(IN(x)*gamma+beta + res):
//   g   = dy * act'(y)                          (written to g_out if non-null: grad wrt res)
//   dx  = rstd*gamma*( g - mean_hw(g) - xhat*mean_hw(g*xhat) )  (+ addend, optional)
//   dgamma (+)= sum g*xhat, dbeta (+)= sum g   (if gamma != null)
// y_for_mask: the saved forward output (sign decides act'); when null the pre-activation
// xhat*gamma+beta is recomputed instead (only valid when there was no residual).
int in_backward(const float* dy, const float* y_for_mask, const float* x, int N, int HW, int C,
                const float* mean, const float* rstd, const float* gamma, const float* beta,
                int act, const float* addend, float* dx, float* g_out, float* dgamma,
                float* dbeta, float* scratch, bool accumulate_affine, cudaStream_t s);
size_t in_backward_scratch_floats(int N, int C);
// Backward of  p = maxpool3x3s2p1(relu(IN(x)))  (non-affine norm; torchvision ResNet stem,
// eye_net.py:48-50) in two kernels: the two per-(n, c) sums of the norm's backward are taken over
// the POOL WINDOWS (only the argmax of a window carries gradient, and its normalised value is the
// pooled output p itself), then one pass gathers the pooled gradient per pixel and writes the
// norm's input gradient as fp32 (dx) and / or as the bf16 hi / lo planes the stem's weight
// gradient reads (d_lo may be NULL).  H and W even.  scratch: in_backward_scratch_floats(N, C).
int stem_pool_in_backward(const float* dpool, const float* pooled, const int32_t* idx,
                          const float* x, int N, int H, int W, int C, const float* mean,
                          const float* rstd, float* dx, uint16_t* d_hi, uint16_t* d_lo,
                          float* scratch, cudaStream_t s);

// ------------------------------------------------------------------------------ pools --
int nchw_to_nhwc(const float* x, int N, int C, int H, int W, float* y, cudaStream_t s);
int nhwc_to_nchw(const float* x, int N, int C, int H, int W, float* y, cudaStream_t s);
// Fused stem tail: y = maxpool3x3s2p1( relu( IN(x) ) ), idx = argmax position (first max,
// row-major window order, flat h*W+w in the input plane) as int32.
int in_relu_maxpool(const float* x, int N, int H, int W, int C, const float* mean,
                    const float* rstd, float* y, int32_t* idx, cudaStream_t s);
// dx (pre-norm grad input g wrt relu(IN(x)) output, scattered) : g[N,H,W,C] = scatter(dy)
int maxpool_bwd_scatter(const float* dy, const int32_t* idx, int N, int H, int W, int OH, int OW,
                        int C, float* g, cudaStream_t s);
int avgpool_fwd(const float* x, int N, int HW, int C, float* y, cudaStream_t s);
int avgpool_bwd(const float* dy, int N, int HW, int C, float* dx, cudaStream_t s);
// torch AdaptiveMaxPool2d semantics: window [floor(i*L/O), ceil((i+1)*L/O))
// copy (optional): x is also written, in its own pixel order, into rows of copy_ld floats (the skip
// half of the decoder's concat buffer: the kernel reads every input element anyway)
int adaptive_maxpool_fwd(const float* x, int N, int H, int W, int C, int OH, int OW, float* y,
                         int32_t* idx, cudaStream_t s, float* copy = nullptr, int copy_ld = 0);
// dx = scatter(dy) (+ addend: the skip-connection gradient, dx's pixel order with rows of
// addend_ld floats -- 0 = C -- so that it can be a channel slice of the concat's gradient)
int adaptive_maxpool_bwd(const float* dy, const int32_t* idx, int N, int H, int W, int C, int OH,
                         int OW, float* dx, cudaStream_t s, const float* addend = nullptr,
                         int addend_ld = 0);
// bilinear (align_corners=False) upsample of x[N,H,W,C] to [OH,OW], written into
// y[N,OH,OW,ldy] at channel offset coff (the concat of refine_net.py:123-126)
int upsample_bilinear_fwd(const float* x, int N, int H, int W, int C, int OH, int OW, float* y,
                          int ldy, int coff, cudaStream_t s);
int upsample_bilinear_bwd(const float* dy, int lddy, int coff, int N, int H, int W, int C, int OH,
                          int OW, float* dx, cudaStream_t s);
// y[rows, ldy] (cols coff..coff+C) = x[rows, ldx] (cols xoff..xoff+C)
int copy_channels(const float* x, long long rows, int C, int ldx, int xoff, float* y, int ldy,
                  int coff, bool accumulate, cudaStream_t s);

}  // namespace eve

namespace eve {
// ----------------------------------------------------------------------------- linear --
// nn.Linear as a 1x1 convolution on a 1x1 image.  W is torch's [N_out, K_in] row-major.
// wt_scratch: K*N floats (transposed copy used by the forward GEMM).
int linear_fwd(const float* x, int M, int K, const float* W, const float* b, int N, float* y,
               float* wt_scratch, cudaStream_t s);
// dx[M,K] = dy[M,N] . W  (+ addend)
int linear_dgrad(const float* dy, int M, int N, const float* W, int K, const float* addend,
                 float* dx, cudaStream_t s);
// dW[N,K] (+)= dy^T x ; db[N] (+)= colsum(dy) when db != null.
// scratch: linear_wgrad_scratch_floats(M, K, N)
size_t linear_wgrad_scratch_floats(int M, int K, int N);
int linear_wgrad(const float* x, const float* dy, int M, int K, int N, float* dW, float* db,
                 float* scratch, bool accumulate, cudaStream_t s);

// ------------------------------------------------------------------------ elementwise --
enum Ew { EW_SELU = 0, EW_RELU = 1, EW_TANH_HALFPI = 2, EW_SIGMOID = 3, EW_LEAKY = 4, EW_TANH = 5 };
// y = f(x)
int ew_fwd(int op, const float* x, long long n, float* y, cudaStream_t s);
// dx = dy * f'(.) ; `ref` is x for SELU/RELU/LEAKY and y for TANH_HALFPI/SIGMOID/TANH
int ew_bwd(int op, const float* dy, const float* ref, long long n, float* dx, cudaStream_t s);
// y = a + b
int ew_add(const float* a, const float* b, long long n, float* y, cudaStream_t s);
int fill_zero(float* p, long long n, cudaStream_t s);

inline cudaStream_t as_stream(eve_stream_t s) { return (cudaStream_t)s; }
}  // namespace eve
