/* libeve_b200.so -- C ABI of the B200-native EVE hot path (EyeNet + GazeRefineNet, fwd + bwd).
 *
 * The reference (swook/EVE) is pure Python/PyTorch and has no FFI of its own; this header
 * IS the boundary a maintainer binds with ctypes.  Each entry point
 * names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer (fp32 unless stated) owned by the caller.  The library
 *    never allocates or frees device memory; scratch comes from `workspace`, activations that
 *    backward needs are kept in the caller's `saved` buffer.  Sizes: eve_*_saved_bytes(),
 *    eve_*_workspace_bytes().
 *  - Tensors crossing the boundary use the reference's layouts (NCHW images, [rows, features]
 *    vectors, OIHW conv weights, [out, in] linear weights).  NHWC is internal.
 *  - All work is enqueued on `stream` (a cudaStream_t); no entry synchronises the device.
 *  - Return value: EVE_OK or an EVE_ERR_* code; eve_last_error() gives the message for the
 *    calling thread.  Nothing aborts and nothing falls back to another implementation.
 *  - `weights` / `grads` are tables of device pointers in the order documented per entry;
 *    a grads slot may be NULL (that gradient is skipped).  `accumulate` != 0 adds into grads.
 */
#ifndef EVE_B200_H
#define EVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVE_OK 0
#define EVE_ERR_SHAPE 1     /* bad dimension / unsupported size            */
#define EVE_ERR_CONFIG 2    /* unsupported knob combination                */
#define EVE_ERR_CUDA 3      /* a CUDA runtime call or launch failed        */
#define EVE_ERR_WORKSPACE 4 /* workspace / saved buffer too small          */
#define EVE_ERR_NULL 5      /* a required pointer is NULL                  */

typedef void* eve_stream_t; /* cudaStream_t */

int eve_version(void);
const char* eve_last_error(void);
/* number of kernels this library has launched since it was loaded (all threads) */
long long eve_launch_count(void);
/* Optional device-side timing of the convolution kernels, used by bench.py for the roofline
 * line: when enabled, every conv launch is bracketed by CUDA events on its own stream.
 * kind: 0 forward, 1 data-gradient, 2 weight-gradient.  eve_profile_read() synchronises the
 * recorded events and returns summed milliseconds, algorithmic FLOPs (2*M*N*K) and algorithmic
 * bytes (fp32 input + output + weights, once each) and the launch count of that kind. */
void eve_profile_enable(int on);
void eve_profile_reset(void);
int eve_profile_read(int kind, double* ms, double* flops, double* bytes, long long* launches);
/* Per-launch listing of the recorded convolution passes, one text line each:
 * "kind N H W Cin Cout ksize stride ms flops bytes".  Returns the bytes the listing needs
 * (including the terminating NUL); call with buf = NULL to size the buffer. */
long long eve_profile_dump(char* buf, long long cap);
/* Hardware probe (measurement aid): `grid` CTAs each issue reps x nmma
 * tcgen05.mma (M=128, N=n, K=16, both operands in shared memory, A start shifted by
 * a_shift_bytes) and report their clock64 span in cycles_out[grid] (device memory). */
int eve_probe_mma_rate(int n, int nmma, int reps, int a_shift_bytes, int distinct_a, int grid,
                       long long* cycles_out, eve_stream_t stream);
/* the same with row_bytes = 128 / 64 / 32: SWIZZLE_128B / 64B / 32B operand rows (64 / 32 / 16
 * channels per pixel row) */
int eve_probe_mma_rate_swizzle(int n, int nmma, int reps, int a_shift_bytes, int distinct_a,
                               int row_bytes, int grid, long long* cycles_out, eve_stream_t stream);
/* the same with 1..4 warps issuing independent chains concurrently (each into its own accumulator
 * columns); cycles_out is the span of warp 0, i.e. divide by nmma * reps for cycles per MMA of ONE
 * chain while `issuers` chains run */
int eve_probe_mma_rate_issuers(int n, int nmma, int reps, int a_shift_bytes, int distinct_a,
                               int row_bytes, int issuers, int grid, long long* cycles_out,
                               eve_stream_t stream);

/* the same for CTA PAIRS (tcgen05.mma.cta_group::2, M = 256 over two SMs, B split in halves):
 * `pairs` clusters of two CTAs; cycles_out[pairs] = span of each leader */
int eve_probe_mma_rate_pair(int n, int nmma, int reps, int pairs, long long* cycles_out,
                            eve_stream_t stream);

/* ------------------------------------------------------------------ building blocks --
 * Exposed so that each kernel family can be parity-tested on its own.  NHWC fp32. */

/* nn.Conv2d (refine_net.py:47,51,60,214,216,221,223; torchvision resnet conv3x3/conv1x1/conv1;
 * common.py:338,362,395-398).  x[N,H,W,Cin], w OIHW, bias[Cout] or NULL, y[N,OH,OW,Cout]. */
typedef struct {
  int n, h, w, cin, cout, ksize, stride, pad;
} eve_conv_params;
/* Kernel selection for every convolution in the library (process-wide):
 *   0 = fp32 CUDA-core implicit GEMM everywhere (exact fp32 products)
 *   1 = tcgen05 tensor cores with split-bf16 operands (hi*hi + hi*lo + lo*hi, fp32 accumulate in
 *       TMEM; products exact to ~2^-16) wherever the geometry allows -- the default
 *   2 = tcgen05 with a single bf16 pass (fastest; ~1e-2 relative error, outside the 1e-3 bar)
 * The default can also be set with the environment variable EVE_B200_CONV_MODE. */
void eve_set_conv_mode(int mode);
int eve_get_conv_mode(void);
/* Tuning switches of the tensor-core path (process-wide, default = fastest validated setting;
 * each also reads an EVE_B200_<NAME> environment default).  They select between kernels that
 * compute the same result, so that every variant can be parity-tested and A/B-timed:
 *   "tc_stage_cap"        2..24  ring depth limit of the implicit-GEMM kernel
 *   "tc_row_kernel"       0/1    halo-row kernel (input rows staged once, filter taps taken as
 *                                shifted shared-memory descriptors) for 3x3 stride-1, W == 128
 *   "tc_row_strips"       0..128 row strips per image (0 = automatic)
 *   "tc_row_wgrad"        0..2   halo-row weight-gradient kernel (128-pixel-wide maps): 0 off, 1 one
 *                                instruction group per filter row, 2 (default) the three filter rows
 *                                of a 3x3 filter stacked along N (one instruction per dy row, K step
 *                                and operand product)
 *   "tc_wgrad_waves"      1..8   full waves of CTAs the split-K weight gradient fills
 *   "fused_planes"        0/1    normalise+activate kernels write the consuming convolution's
 *                                16-bit operand planes directly (no fp32 activation, no split pass);
 *                                must not change between a forward and its backward
 *   "fused_norm"          0/1    one-pass cluster InstanceNorm kernels (statistics + normalisation
 *                                from a single read; backward emits dy planes and bias-gradient
 *                                sums) and plane-to-plane residual blocks; 0 = the separate
 *                                stats / apply / reduce / split passes of round 1
 *   "tc_strip"            0..2   padded-strip kernel for 3x3 stride-1 layers narrower than 128
 *                                pixels (input staged once per K chunk as a zero-padded strip, the
 *                                nine taps taken as descriptor shifts, weight tiles shared by all
 *                                tiles of the strip): 0 off, 1 where its plan wastes at most a
 *                                seventh more MMA rows than the per-tap box kernel, 2 wherever it fits
 *   "tc_wgrad_strip"      0/1    padded-strip weight-gradient kernel (3x3 stride 1, 64 output
 *                                channels, maps narrower than 128 pixels): x and dy strips staged
 *                                once per item, the nine taps as descriptor shifts
 *   "cgru_persistent"     0/1    ConvGRU bottleneck (64 features, 5x8 maps) as ONE persistent kernel
 *                                per sequence (x halves of the gate convolutions batched over
 *                                time, recurrence in shared memory / TMEM); 0 = one convolution
 *                                launch pair per time step
 *   "tc_dual"             0/1    per-tap box kernel (64 / 128 output channels per tile) with TWO
 *                                MMA-issuing warps: even and odd ring stages of a tile go to two
 *                                partial accumulators in TMEM that the epilogue adds.  Default 0:
 *                                measured 4-11 % slower (the gaps between MMAs are operand waits)
 *   "tc_pair"             0..2   per-tap box kernel with 64 / 128 output channels per tile launched as
 *                                clusters of two CTAs: each loads half of every weight stage and
 *                                TMA-multicasts it to both (the kernel is bound by its L2 -> shared
 *                                memory operand stream); bit-identical results.  0 = off (default:
 *                                measured 2-8 % SLOWER per layer -- every SM still ingests the whole
 *                                stage, so the multicast relieves L2 but not the SM's input port, and
 *                                the two rings now advance in lockstep), 1 = when every pair has
 *                                work, 2 = whenever a layer has two tiles
 *   "stem_windows"        0..2   EyeNet stem (7x7 stride 2, 3 channels) forward without an im2col
 *                                matrix: a filter row of an output pixel is one 32-value window
 *                                (8 pixels x 4 zero-padded channels) and the convolution a 7-tap
 *                                tensor-core pass over window rows.  1 = one window per output
 *                                column written once (default), 2 = windows overlapping inside the
 *                                zero-padded image (TMA strides smaller than the box), 0 = im2col
 *   "in_stream"           0..2   InstanceNorm backward without shared-memory staging (second read of
 *                                dy / x served by L2, two CTAs per SM): 0 = never, 1 = for maps whose
 *                                staged form needs one CTA per SM, 2 = always (default: measured
 *                                fastest inside the training step, where dy is usually still in L2)
 *   "stem_fused_bwd"      0/1    EyeNet stem backward (max-pool -> ReLU -> InstanceNorm) without the
 *                                dense un-pooled gradient map: the norm's two sums are taken over the
 *                                pool windows, one pass gathers the pooled gradient per pixel and
 *                                writes the weight gradient's bf16 dy planes.  0 = scatter, reduce,
 *                                apply and split as four passes
 * Unknown names / out-of-range values return EVE_ERR_CONFIG. */
int eve_set_option(const char* name, int value);
int eve_get_option(const char* name, int* value);
size_t eve_conv2d_workspace_bytes(const eve_conv_params* p);
/* Text description of the kernel (and tiling plan) eve_conv2d_fwd picks for this geometry in
 * conv mode 1 (diagnostics, docs). */
int eve_conv2d_describe(const eve_conv_params* p, char* buf, size_t cap);
int eve_conv2d_fwd(const eve_conv_params* p, const float* x, const float* w, const float* bias,
                   float* y, void* workspace, size_t workspace_bytes, eve_stream_t stream);
/* dx = d(loss)/dx given dy */
int eve_conv2d_dgrad(const eve_conv_params* p, const float* dy, const float* w, float* dx,
                     void* workspace, size_t workspace_bytes, eve_stream_t stream);
/* dw (OIHW) and dbias (may be NULL) */
int eve_conv2d_wgrad(const eve_conv_params* p, const float* x, const float* dy, float* dw,
                     float* dbias, void* workspace, size_t workspace_bytes, eve_stream_t stream);

/* nn.InstanceNorm2d(eps=1e-5, biased variance, no running stats) + activation
 * (eye_net.py:50; refine_net.py:46,50,59,215).  act: 0 none, 1 ReLU, 2 LeakyReLU(0.01).
 * gamma/beta NULL => non-affine.  mean/rstd [N,C] are outputs of fwd, inputs of bwd. */
int eve_instnorm_act_fwd(const float* x, int n, int hw, int c, const float* gamma,
                         const float* beta, int act, float* y, float* mean, float* rstd,
                         eve_stream_t stream);
int eve_instnorm_act_bwd(const float* dy, const float* y, const float* x, int n, int hw, int c,
                         const float* mean, const float* rstd, const float* gamma, int act,
                         float* dx, float* dgamma, float* dbeta, void* workspace,
                         size_t workspace_bytes, eve_stream_t stream);

/* The same normalisation as ONE pass over the tensor (csrc/in_fused.cu): a thread-block cluster
 * per (image, channel group) stages its pixels in shared memory, reduces the statistics through
 * distributed shared memory and normalises from there.  These are the kernels the networks use;
 * exposed so that they can be parity-tested alone.
 *  fwd: y = act(IN(x)*gamma+beta (+ x2)) with x2_mode 0 none / 1 residual / 2 x2 is itself
 *       instance-normalised (non-affine; mean2/rstd2 out).  Outputs, each optional: fp32 y, the
 *       16-bit hi/lo operand planes of y (fmt 0 = fp16, 1 = bf16; y ~ hi + lo) and planes B of
 *       act(IN(x)*gamma_b+beta_b) (a second affine set over the same statistics,
 *       refine_net.py:45-62 layers.0 / skip_layer.0).
 *  bwd: given dy (and dy2 for the second affine set) produces dx = d/dx (+ addend) as fp32 and/or
 *       bf16 hi/lo planes, g_out = dy*act'(.), the affine gradients and dbias[c] = sum of dx
 *       (the bias gradient of the convolution that produced x).  ymask: saved forward output
 *       when a residual was added (its sign gives act'), else NULL (recomputed). */
size_t eve_instnorm_fused_workspace_bytes(int n, int hw, int c);
int eve_instnorm_fused_fwd(const float* x, const float* x2, int x2_mode, int n, int hw, int c,
                           const float* gamma, const float* beta, const float* gamma_b,
                           const float* beta_b, int act, int fmt, float* mean, float* rstd,
                           float* mean2, float* rstd2, float* y, void* hi_a, void* lo_a,
                           void* hi_b, void* lo_b, eve_stream_t stream);
int eve_instnorm_fused_bwd(const float* dy, const float* dy2, const float* ymask, const float* x,
                           int n, int hw, int c, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, const float* gamma2,
                           const float* beta2, int act, const float* addend, float* dx,
                           void* dx_hi, void* dx_lo, float* g_out, float* dgamma, float* dbeta,
                           float* dgamma2, float* dbeta2, float* dbias, void* workspace,
                           size_t workspace_bytes, eve_stream_t stream);

/* The ResNet stem tail as one kernel: maxpool3x3 s2 p1 (relu (InstanceNorm(x))) (torchvision
 * ResNet.bn1 / relu / maxpool behind eye_net.py:48-50,106).  x[n,h,w,c] NHWC -> y[n,oh,ow,c] with
 * oh = (h+2-3)/2+1; idx = int32 flat ih*w+iw of the first maximum in row-major window order
 * (what ATen returns; padding never wins).  mean/rstd [n,c] are outputs. */
int eve_in_relu_maxpool_fwd(const float* x, int n, int h, int w, int c, float* mean, float* rstd,
                            float* y, int32_t* idx, eve_stream_t stream);
/* Its backward in two passes, without a dense un-pooled gradient map: dy[n,oh,ow,c] is the gradient
 * of the pooled output, y / idx / mean / rstd what the forward call wrote, h and w even.  The
 * gradient of x is written as fp32 (dx, may be NULL) and / or as bf16 hi + lo planes (dx_hi,
 * dx_lo: the operand form of the stem convolution's weight gradient; NULL to skip, dx_lo alone may
 * be NULL).  scratch: 2*n*c floats. */
int eve_in_relu_maxpool_bwd(const float* dy, const float* y, const int32_t* idx, const float* x,
                            int n, int h, int w, int c, const float* mean, const float* rstd,
                            float* dx, uint16_t* dx_hi, uint16_t* dx_lo, float* scratch,
                            eve_stream_t stream);

/* nn.AdaptiveMaxPool2d (refine_net.py:93,121): idx = int32 flat h*W+w of the first maximum. */
int eve_adaptive_maxpool_fwd(const float* x, int n, int h, int w, int c, int oh, int ow, float* y,
                             int32_t* idx, eve_stream_t stream);
int eve_adaptive_maxpool_bwd(const float* dy, const int32_t* idx, int n, int h, int w, int c,
                             int oh, int ow, float* dx, eve_stream_t stream);
/* nn.Upsample(bilinear, align_corners=False) (refine_net.py:101,124) */
int eve_upsample_bilinear_fwd(const float* x, int n, int h, int w, int c, int oh, int ow, float* y,
                              eve_stream_t stream);
int eve_upsample_bilinear_bwd(const float* dy, int n, int h, int w, int c, int oh, int ow,
                              float* dx, eve_stream_t stream);
/* layout helpers */
int eve_nchw_to_nhwc(const float* x, int n, int c, int h, int w, float* y, eve_stream_t stream);
int eve_nhwc_to_nchw(const float* x, int n, int c, int h, int w, float* y, eve_stream_t stream);

/* -------------------------------------------------------------------- EyeNet: CNN --
 * torchvision ResNet(BasicBlock,[2,2,2,2], num_classes=nf, norm_layer=InstanceNorm2d) as called
 * at eye_net.py:48-50,106: conv7x7s2 -> IN -> ReLU -> maxpool3x3s2 -> 8 BasicBlocks -> avgpool
 * -> fc.  x[n,3,128,128] NCHW -> feat[n,nf].
 * weights (22+1): conv1, then per layer l=1..4, block b=0..1: conv1, conv2, (downsample.0 when
 * l>1 && b==0), then fc.weight, fc.bias.  grads: same order. */
#define EVE_EYENET_CNN_NUM_WEIGHTS 22
typedef struct {
  int n;  /* number of eye patches (B*T*2 when time-batched) */
  int nf; /* fc output features (eye_net_rnn_num_features / eye_net_static_num_features) */
  int h, w; /* patch size, 128 x 128 */
} eve_eyenet_cnn_params;
size_t eve_eyenet_cnn_saved_bytes(const eve_eyenet_cnn_params* p);
size_t eve_eyenet_cnn_workspace_bytes(const eve_eyenet_cnn_params* p);
int eve_eyenet_cnn_fwd(const eve_eyenet_cnn_params* p, const float* x, const float* const* weights,
                       float* feat, void* saved, size_t saved_bytes, void* workspace,
                       size_t workspace_bytes, eve_stream_t stream);
int eve_eyenet_cnn_bwd(const eve_eyenet_cnn_params* p, const float* dfeat,
                       const float* const* weights, float* const* grads, int accumulate,
                       const void* saved, size_t saved_bytes, void* workspace,
                       size_t workspace_bytes, eve_stream_t stream);

/* ------------------------------------------------------------------- EyeNet: tail --
 * eye_net.py:109-140 over whole sequences: cat(head pose) -> fc_common -> RNN cells (or
 * static_fc) -> gaze head (tanh * pi/2) and pupil head (ReLU).
 * feat[batch,steps,nf], head_pose[batch,steps,2] (NULL iff !use_head_pose),
 * h0/c0[cells,batch,nf] initial states (NULL => zeros; c0 only for LSTM),
 * g[batch,steps,2], pupil[batch,steps], hT/cT[cells,batch,nf] final states (may be NULL).
 * weights: fc_common.0.{w,b}, fc_common.2.{w,b}, then rnn: per cell {w_ih,w_hh,b_ih,b_hh} or
 * static: static_fc.0.{w,b}; then fc_to_gaze.0.{w,b}, fc_to_gaze.2.w, fc_to_pupil.0.{w,b},
 * fc_to_pupil.2.{w,b}. */
#define EVE_RNN_NONE 0
#define EVE_RNN_RNN 1
#define EVE_RNN_LSTM 2
#define EVE_RNN_GRU 3
typedef struct {
  int batch, steps, nf;
  int use_head_pose;
  int rnn_type;  /* EVE_RNN_*; NONE = static_fc path (eye_net_use_rnn = False) */
  int rnn_cells;
} eve_eyenet_tail_params;
int eve_eyenet_tail_num_weights(const eve_eyenet_tail_params* p);
size_t eve_eyenet_tail_saved_bytes(const eve_eyenet_tail_params* p);
size_t eve_eyenet_tail_workspace_bytes(const eve_eyenet_tail_params* p);
int eve_eyenet_tail_fwd(const eve_eyenet_tail_params* p, const float* feat, const float* head_pose,
                        const float* h0, const float* c0, const float* const* weights, float* g,
                        float* pupil, float* hT, float* cT, void* saved, size_t saved_bytes,
                        void* workspace, size_t workspace_bytes, eve_stream_t stream);
/* dg, dpupil required; dhT/dcT may be NULL (no gradient into the final state);
 * dfeat required; dh0/dc0 may be NULL. */
int eve_eyenet_tail_bwd(const eve_eyenet_tail_params* p, const float* dg, const float* dpupil,
                        const float* dhT, const float* dcT, const float* const* weights,
                        float* dfeat, float* dh0, float* dc0, float* const* grads, int accumulate,
                        const void* saved, size_t saved_bytes, void* workspace,
                        size_t workspace_bytes, eve_stream_t stream);

/* ---------------------------------------------------------------------- RefineNet --
 * refine_net.py:237-255 over whole sequences: cat(screen, heatmap) -> initial -> 5-level
 * pre-activation encoder (AdaptiveMaxPool between levels) -> ConvRNN bottleneck stepped over
 * `steps` -> decoder (bilinear upsample + skip concat) -> final -> sigmoid.
 * screen[batch,steps,3,72,128] NCHW (NULL iff in_channels == 1), heatmap[batch,steps,1,72,128],
 * h0/c0[cells,batch,nf,5,8] NCHW (NULL => zeros), out[batch,steps,1,72,128],
 * hT/cT like h0 (may be NULL).  Weight order: eve_refinenet_weight_name(). */
#define EVE_CRNN_NONE 0
#define EVE_CRNN_CRNN 1
#define EVE_CRNN_CLSTM 2
#define EVE_CRNN_CGRU 3
typedef struct {
  int batch, steps;
  int in_channels; /* 4 with screen content, 1 without (refine_net.py:183) */
  int use_skip;    /* refine_net_use_skip_connections */
  int rnn_type;    /* EVE_CRNN_*; NONE = refine_net_use_rnn False */
  int rnn_cells;
  int nf;          /* refine_net_num_features (64) */
} eve_refinenet_params;
int eve_refinenet_num_weights(const eve_refinenet_params* p);
/* name of weight slot i relative to `refine_net.` (e.g. "network.encoder_blocks.0.layers.2.weight");
 * NULL when i is out of range.  The string lives in thread-local storage. */
const char* eve_refinenet_weight_name(const eve_refinenet_params* p, int i);
size_t eve_refinenet_saved_bytes(const eve_refinenet_params* p);
size_t eve_refinenet_workspace_bytes(const eve_refinenet_params* p);
int eve_refinenet_fwd(const eve_refinenet_params* p, const float* screen, const float* heatmap,
                      const float* h0, const float* c0, const float* const* weights, float* out,
                      float* hT, float* cT, void* saved, size_t saved_bytes, void* workspace,
                      size_t workspace_bytes, eve_stream_t stream);
int eve_refinenet_bwd(const eve_refinenet_params* p, const float* dout, const float* dhT,
                      const float* dcT, const float* const* weights, float* dheatmap, float* dh0,
                      float* dc0, float* const* grads, int accumulate, const void* saved,
                      size_t saved_bytes, void* workspace, size_t workspace_bytes,
                      eve_stream_t stream);

/* ------------------------------------------------------------- gaze <-> screen ops --
 * common.py:226-243 make_heatmap / batch_make_heatmaps: centres_px[n,2] (screen pixels) ->
 * out[n,1,hm_h,hm_w] = exp(-((x-cx)^2+(y-cy)^2)/(2 sigma^2)) + 1e-8. */
typedef struct {
  int n, hm_w, hm_h;
  float screen_w, screen_h; /* actual_screen_size (1920, 1080) */
  float sigma;
} eve_heatmap_params;
int eve_heatmap_fwd(const eve_heatmap_params* p, const float* centres_px, float* out,
                    eve_stream_t stream);
int eve_heatmap_bwd(const eve_heatmap_params* p, const float* centres_px, const float* dout,
                    float* dcentres, eve_stream_t stream);
/* common.py:294-323 soft_argmax: heatmaps[n,1,hm_h,hm_w] -> pog_px[n,2] (softmax(100 h),
 * expectation over linspace(0,1) grids, scaled to the screen and clamped). */
int eve_soft_argmax_fwd(const eve_heatmap_params* p, const float* heatmaps, float* pog_px,
                        eve_stream_t stream);
int eve_soft_argmax_bwd(const eve_heatmap_params* p, const float* heatmaps, const float* dpog,
                        float* dheatmaps, eve_stream_t stream);
/* common.py:149-179 to_screen_coordinates (+ :32-40, :89-126): per sample
 * origin[n,3], g[n,2] (pitch, yaw), rot[n,3,3], inv_cam[n,4,4], ppm[n,2] ->
 * pog_mm[n,2], pog_px[n,2] (clamped to the screen).  bwd: gradient w.r.t. g only. */
int eve_pog_fwd(int n, const float* origin, const float* g, const float* rot,
                const float* inv_cam, const float* ppm, float screen_w, float screen_h,
                float* pog_mm, float* pog_px, eve_stream_t stream);
int eve_pog_bwd(int n, const float* origin, const float* g, const float* rot,
                const float* inv_cam, const float* ppm, float screen_w, float screen_h,
                const float* dpog_mm, const float* dpog_px, float* dg, eve_stream_t stream);

/* common.py:129-146 calculate_combined_gaze_direction: per frame origin[n,3] (averaged eye origin),
 * pog_mm[n,2], head_rot[n,3,3], cam[n,4,4] (camera_transformation) -> g[n,2] (pitch, yaw).
 * bwd: gradient w.r.t. pog_mm only. */
int eve_combined_gaze_fwd(int n, const float* origin, const float* pog_mm, const float* head_rot,
                          const float* cam, float* g, eve_stream_t stream);
int eve_combined_gaze_bwd(int n, const float* origin, const float* pog_mm, const float* head_rot,
                          const float* cam, const float* dg, float* dpog_mm, eve_stream_t stream);
/* common.py:182-218 apply_offset_augmentation: g[n,2], head_rot[n,3,3], kappa[n/frames_per_kappa,2]
 * (eve.py:466-477 draws one kappa per clip and repeats it over time: frames_per_kappa = T; pass 1
 * for a per-frame kappa) -> out[n,2].  bwd: gradient w.r.t. g only (kappa is data). */
int eve_offset_augmentation_fwd(int n, int frames_per_kappa, const float* g, const float* head_rot,
                                const float* kappa, int inverse_kappa, float* out,
                                eve_stream_t stream);
int eve_offset_augmentation_bwd(int n, int frames_per_kappa, const float* g, const float* head_rot,
                                const float* kappa, int inverse_kappa, const float* dout, float* dg,
                                eve_stream_t stream);
/* eve.py:441-543 EVE.calculate_additional_labels, the per-frame block: PoG labels in cm
 * (:449-456), averaged origin / PoG and the left&right validity (:498-517), ground-truth combined
 * gaze g (:534-543).  All arrays [n, .]; validity arrays are bytes (torch.bool). */
typedef struct {
  int n;
  const float *left_pog_px, *right_pog_px;          /* [n,2] left/right_PoG_tobii */
  const unsigned char *left_valid, *right_valid;    /* [n] */
  const float* mm_per_px;                           /* [n,2] millimeters_per_pixel */
  const float *left_o, *right_o;                    /* [n,3] */
  const float* left_R;                              /* [n,3,3] */
  const float* cam;                                 /* [n,4,4] camera_transformation */
  float *left_pog_cm, *right_pog_cm;                /* out [n,2] */
  float* o;                                         /* out [n,3] */
  float *pog_px, *pog_cm;                           /* out [n,2] */
  unsigned char* valid;                             /* out [n] */
  float* g;                                         /* out [n,2] */
} eve_label_args;
int eve_labels_fwd(const eve_label_args* a, eve_stream_t stream);
/* eve.py:519-531: nsig (<= 3) label heatmaps per frame from one launch,
 * outs[s][n,1,hm_h,hm_w] = make_heatmap(centres_px[n], sigmas[s]) * valid[n]
 * (sigmas and outs are HOST arrays; p->sigma is ignored). */
int eve_heatmap_labels_fwd(const eve_heatmap_params* p, const float* centres_px,
                           const unsigned char* valid, int nsig, const float* sigmas,
                           float* const* outs, eve_stream_t stream);
/* common.py:249-287 gaze-history maps for every prefix of a clip as an O(T) recurrence (the
 * reference re-sums the history at every step, eve.py:596-601): timestamps[batch,steps] int64
 * (0 = padded frame), valid[batch,steps] bytes, heatmaps / out [batch,steps,hw].
 * bwd: gradient w.r.t. heatmaps.  scratch: eve_gaze_history_scratch_bytes(). */
size_t eve_gaze_history_scratch_bytes(int batch, int steps);
int eve_gaze_history_fwd(int batch, int steps, int hw, const long long* timestamps,
                         const unsigned char* valid, float decay, const float* heatmaps, float* out,
                         void* scratch, size_t scratch_bytes, eve_stream_t stream);
int eve_gaze_history_bwd(int batch, int steps, int hw, const long long* timestamps,
                         const unsigned char* valid, float decay, const float* dout,
                         float* dheatmaps, void* scratch, size_t scratch_bytes,
                         eve_stream_t stream);

/* ------------------------------------------------------------- losses and metrics --
 * eve.py:286-439 + losses/base_loss_with_validity.py:32-73: a table of validity-masked sequence
 * losses, each  mean_b( sum_t(valid * l) / n_valid_b  [divided only if n_valid_b > 1] ),
 * evaluated by ONE launch; pred/gt are [batch,steps,dim], valid [batch,steps] bytes. */
enum {
  EVE_LOSS_ANGULAR = 0,   /* losses/angular.py: angle between two pitch/yaw pairs, degrees (dim 2) */
  EVE_LOSS_MSE = 1,       /* mean over dim of (a-b)^2 */
  EVE_LOSS_L1 = 2,        /* mean over dim of |a-b| */
  EVE_LOSS_EUCLIDEAN = 3, /* sqrt(sum over dim of (a-b)^2) */
  EVE_LOSS_IDENTITY = 4   /* pred[batch,steps] already holds the per-frame loss (heatmap terms) */
};
#define EVE_LOSS_MAX_TERMS 40
typedef struct {
  int op, dim;
  const float* pred;
  const float* gt;              /* NULL for EVE_LOSS_IDENTITY */
  const unsigned char* valid;
  const unsigned char* valid2;  /* optional second mask, ANDed (lr-consistency terms) */
  float* dpred;                 /* bwd only: gradient buffer of pred, ACCUMULATED into (terms that
                                   share a prediction share it; zero it first); NULL = no gradient */
} eve_loss_term;
/* out[nterms] / dout[nterms] are DEVICE arrays; `terms` is a HOST array (copied into the launch) */
int eve_masked_losses_fwd(int nterms, const eve_loss_term* terms, int batch, int steps, float* out,
                          eve_stream_t stream);
int eve_masked_losses_bwd(int nterms, const eve_loss_term* terms, int batch, int steps,
                          const float* dout, eve_stream_t stream);
/* cross_entropy.py:29-35 / mse.py on heatmaps: per frame f, bce[f] = mean over hw of
 * F.binary_cross_entropy(pred, gt) (logs clamped at -100) and mse[f] = mean (pred-gt)^2; either
 * output may be NULL.  bwd: dpred = dbce[f] * dBCE + dmse[f] * dMSE (either may be NULL). */
int eve_heatmap_frame_losses_fwd(int n, int hw, const float* pred, const float* gt, float* bce,
                                 float* mse, eve_stream_t stream);
int eve_heatmap_frame_losses_bwd(int n, int hw, const float* pred, const float* gt,
                                 const float* dbce, const float* dmse, float* dpred,
                                 eve_stream_t stream);

/* ------------------------------------------------------------------ input pipeline --
 * datasources/eve_sequences.py:196-211 (preprocess_frames / preprocess_screen_frames) and the
 * eye-patch split of :283-285 on the device: frames[n,h,w_in,c] uint8 (decoder layout) ->
 * out[n,c,h,w_out] float32 = frames[:, :, x_offset : x_offset + w_out, :] * scale + bias, rounded
 * like numpy's two in-place float32 ops (bit-identical to the reference's arrays).
 * frames_per_clip (optional, DEVICE int[n / steps]): frames t >= frames_per_clip[clip] are written
 * as zeros (the reference zero-pads short clips after preprocessing, :287-299). */
int eve_preprocess_frames(const unsigned char* frames, int n, int h, int w_in, int c, int x_offset,
                          int w_out, float scale, float bias, const int* frames_per_clip, int steps,
                          float* out, eve_stream_t stream);

/* ------------------------------------------------------------------ optimiser step --
 * training.py:492-502 + train.py:49-55: clip_grad_norm_(max_norm) then Adam with L2
 * weight decay, on ONE flat fp32 buffer (the buffer the NCCL allreduce ran on).
 * grad_scale is applied to the gradients first (1/world_size after a sum-allreduce).
 * norm_out[1] receives the pre-clip global L2 norm (device pointer, may be NULL).
 * workspace: eve_adam_clip_workspace_bytes(). */
typedef struct {
  long long count;
  float lr, beta1, beta2, eps, weight_decay;
  float max_norm;   /* <= 0 disables clipping */
  float grad_scale;
  int step;         /* 1-based Adam step for bias correction (ignored when step_dev is set) */
  int* step_dev;    /* optional DEVICE counter: the kernel increments it and uses the new value as
                       the step -- lets one captured CUDA graph be replayed for every step */
  const float* lr_dev; /* optional DEVICE scalar overriding `lr`: a learning-rate schedule
                          (training.py:382-440,576) can update it between replays of one graph */
} eve_adam_params;
size_t eve_adam_clip_workspace_bytes(const eve_adam_params* p);
int eve_adam_clip_step(const eve_adam_params* p, float* params, const float* grads, float* exp_avg,
                       float* exp_avg_sq, float* norm_out, void* workspace, size_t workspace_bytes,
                       eve_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EVE_B200_H */
