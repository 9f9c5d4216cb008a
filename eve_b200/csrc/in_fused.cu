// One-pass InstanceNorm kernels: statistics AND normalisation from a single read of the tensor.
//
// Reference semantics: nn.InstanceNorm2d (eps 1e-5, biased variance, no running statistics) as
// used by torchvision's ResNet inside EyeNet (eye_net.py:48-50, non-affine) and by RefineNet's
// pre-activation blocks (refine_net.py:45-62,213-216, affine), followed by ReLU / LeakyReLU(0.01).
//
// Round 1 ran every normalisation as separate HBM passes (in_stats, in_apply_planes; backward:
// in_bwd_reduce, in_bwd_apply, split, colsum).  Here a thread-block CLUSTER owns one work item =
// one (image, channel group): its CTAs stage disjoint pixel ranges of the group in shared memory
// with cp.async, reduce the per-channel sums across the cluster through distributed shared memory
// in a fixed order (deterministic, independent of the batch size) and then normalise straight out
// of shared memory.  (A persistent, double-buffered variant was measured slower: the kernels are
// bound by instruction issue and occupancy, not by exposed load latency -- profiles/README.)
// Outputs
//   forward : the consuming convolution's 16-bit hi/lo operand planes (two affine sets from one
//             read when a block has a skip convolution), optionally the fp32 tensor;
//   backward: dx as fp32 and/or as the bf16 hi/lo planes the preceding convolution's gradients
//             read as dy, the per-(n,c) sums the affine gradients need, and per-CTA column sums of
//             dx (= that convolution's bias gradient).
// HBM traffic per element: forward 4 B read + 4 B planes (was 8 + 4), backward 8 B read + 4..8 B
// written (was 16 read + 4 written + 8 split + 4 colsum).
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace eve {
namespace {

constexpr float kEps = 1e-5f;
constexpr int kThr = 512;                    // 16 warps per CTA
constexpr int kWarps = kThr / 32;
constexpr int kMaxQ = 64;                    // channel quads per CTA (<= 256 channels)
constexpr size_t kSmemTwo = 100 * 1024;      // staging bytes per CTA that still allow two CTAs per SM
constexpr size_t kSmemOne = 200 * 1024;      // ... one CTA per SM

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// activations as one select: slope = 0 (ReLU), 0.01 (LeakyReLU), 1 (none)
__host__ __device__ __forceinline__ float act_slope(int act) {
  return act == ACT_RELU ? 0.f : (act == ACT_LEAKY ? 0.01f : 1.f);
}
__device__ __forceinline__ float act_apply(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float act_deriv(float y, float slope) { return y > 0.f ? 1.f : slope; }

// x ~ hi + lo as two 16-bit values each; four elements at a time with the packed converts
template <int FMT>
__device__ __forceinline__ void split4(const float* o, uint2& hi, uint2& lo) {
  if (FMT == TC_BF16) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(o[0], o[1]);
    const __nv_bfloat162 h23 = __floats2bfloat162_rn(o[2], o[3]);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(o[0] - f01.x, o[1] - f01.y);
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(o[2] - f23.x, o[3] - f23.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  } else {
    const __half2 h01 = __floats2half2_rn(o[0], o[1]);
    const __half2 h23 = __floats2half2_rn(o[2], o[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(o[0] - f01.x, o[1] - f01.y);
    const __half2 l23 = __floats2half2_rn(o[2] - f23.x, o[3] - f23.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  }
}

// ------------------------------------------------------------------------- decomposition --
// Depends on (C, HW, staged tensors) only -- never on N -- so that a frame's statistics are
// bit-identical whatever batch it is processed in.
struct FusedPlan {
  int CG, Q, CS, ppc;
  bool one_cta;        // the staged form takes more than half an SM's shared memory
};
int stream_cluster() {        // tuning knob: cluster size of the streaming backward (0 = the staged plan's)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVE_B200_IN_STREAM_CS");
    v = e ? atoi(e) : 0;
    if (v != 1 && v != 2 && v != 4 && v != 8) v = 0;
  }
  return v;
}
// tuning knob: the same for maps of 8192+ pixels only (0 = no exception).  Their second read misses L2
// with clusters of two (3.5 GB read for 1.7 GB of tensors on the 64-channel 72x128 launch); clusters of
// four / eight would keep the tensors in flight below the L2 size -- measured in the training step:
// 30.47 ms (2) vs 30.68 (4) vs 30.69 (8): the barrier and prologue cost outweighs the DRAM bytes saved
int stream_cluster_big() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVE_B200_IN_STREAM_CS_BIG");
    v = e ? atoi(e) : 0;
    if (v != 1 && v != 2 && v != 4 && v != 8) v = 0;
  }
  return v;
}
int max_cluster() {           // tuning knob (environment, read once)
  static int v = 0;
  if (!v) {
    const char* e = getenv("EVE_B200_IN_MAXCS");
    v = e ? atoi(e) : 8;
    if (v != 1 && v != 2 && v != 4 && v != 8 && v != 16) v = 8;
  }
  return v;
}
size_t smem_two() {
  static size_t v = 0;
  if (!v) {
    const char* e = getenv("EVE_B200_IN_SMEM");
    v = e ? (size_t)atol(e) : kSmemTwo;
    if (v < 8192 || v > kSmemOne) v = kSmemTwo;
  }
  return v;
}
// tensors: fp32 slices the kernel stages (forward 1 or 2, backward 2)
bool plan_fused(int C, int HW, int tensors, FusedPlan& p) {
  if (C % 4 != 0 || C < 4 || HW < 1) return false;
  auto bytes = [&](int cg_, int cs_) { return (size_t)cdiv(HW, cs_) * cg_ * 4 * tensors; };
  const int maxcs = max_cluster();
  const int top = C < 4 * kMaxQ ? C : 4 * kMaxQ;
  // 1) groups of >= 16 channels (64-byte pieces), two CTAs per SM; 2) the same with one CTA per
  // SM; 3) / 4) 8-channel groups (one 32-byte sector per pixel: measurably worse DRAM efficiency)
  static const int small_first = [] {
    const char* e = getenv("EVE_B200_IN_PLAN");
    return e ? atoi(e) : 0;     // measured: no gain for the staged kernels, -0.25 ms/step worse for the streaming one
  }();
  for (int stage = 0; stage < 4; ++stage) {
    const size_t budget = (stage & 1) ? kSmemOne : smem_two();
    const int min_cg = stage < 2 ? (C < 16 ? C : 16) : (C < 8 ? C : 8);
    auto take = [&](int CG, int CS) {
      p.CG = CG; p.Q = CG / 4; p.CS = CS; p.ppc = cdiv(HW, CS);
      p.one_cta = (stage & 1) != 0;
    };
    if (small_first) {
      // the smallest cluster first, then the widest channel group that fits it: the same bytes per
      // CTA and the same number of CTAs as a wide group over a large cluster, but fewer CTAs per
      // cluster-wide barrier (64 -> 16 channels still moves 64-byte pieces)
      for (int CS = 1; CS <= maxcs && CS <= HW; CS <<= 1)
        for (int CG = top; CG >= min_cg; CG -= 4) {
          if (C % CG != 0) continue;
          if (bytes(CG, CS) <= budget) {
            take(CG, CS);
            return true;
          }
        }
      continue;
    }
    for (int CG = top; CG >= min_cg; CG -= 4) {
      if (C % CG != 0) continue;
      for (int CS = 1; CS <= maxcs && CS <= HW; CS <<= 1)
        if (bytes(CG, CS) <= budget) {
          take(CG, CS);
          return true;
        }
    }
  }
  return false;
}

// Sum v[0..3] over all threads that own the same channel quad q (thread t: q = t % Q, pixel lane
// t / Q).  The totals land in out[q*4 + j] (t < 4Q), valid after the call (ends with
// __syncthreads()).  Fixed order -> deterministic.  wred: kThr*4 elements of scratch.
template <typename T>
__device__ __forceinline__ void quad_reduce(T* v, int Q, int L, T* wred, T* out) {
  const int t = threadIdx.x, warp = t >> 5, ln = t & 31;
  if ((Q & (Q - 1)) == 0 && Q <= 32) {
    // the lanes of a warp that share q are Q apart: xor tree, then one partial per warp
    for (int off = 16; off >= Q; off >>= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
    }
    if (ln < Q) {
#pragma unroll
      for (int j = 0; j < 4; ++j) wred[(warp * 32 + ln) * 4 + j] = v[j];
    }
    __syncthreads();
    if (t < Q * 4) {
      const int q = t >> 2, j = t & 3;
      T s = (T)0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += wred[(w * 32 + q) * 4 + j];
      out[t] = s;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) wred[t * 4 + j] = v[j];
    __syncthreads();
    if (t < Q * 4) {
      const int q = t >> 2, j = t & 3;
      T s = (T)0;
      for (int l = 0; l < L; ++l) s += wred[(l * Q + q) * 4 + j];
      out[t] = s;
    }
  }
  __syncthreads();
}

// ================================================================================ forward ==
struct InFwdArgs {
  const float* x;
  const float* x2;       // MODE 1: residual added after the affine; 2: second normalised input
  int HW, C, Q, CS, ppc;
  const float *gamma, *beta, *gammaB, *betaB;
  float slope;                          // activation
  float *mean, *rstd, *mean2, *rstd2;   // [N, C]
  float* y;                             // fp32 result (optional)
  uint16_t *hiA, *loA, *hiB, *loB;      // operand planes (A optional, B optional)
};

// grid (CS, C/CG, N), cluster (CS, 1, 1), block 512.  Thread t owns channel quad q = t % Q of the
// pixels lane, lane + L, ... (lane = t / Q, L = 512 / Q) of its CTA's pixel range in every phase.
template <int FMT, int MODE>
__global__ void __launch_bounds__(kThr)
in_fwd_fused_kernel(const InFwdArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = blockIdx.x, cgi = blockIdx.y, n = blockIdx.z;
  const int Q = a.Q, C4 = a.C >> 2;
  constexpr bool two = MODE == 2;
  const int p0 = rank * a.ppc;
  const int np = max(0, min(a.HW, p0 + a.ppc) - p0);
  const int slice = a.ppc * Q;
  float4* sx = reinterpret_cast<float4*>(smraw);
  float4* sx2 = sx + slice;
  float* wred = reinterpret_cast<float*>(sx2 + (two ? slice : 0));   // [kThr*4]
  float* cpart = wred + kThr * 4;        // [2 passes][2 tensors][4*kMaxQ]
  float* stat = cpart + 4 * 4 * kMaxQ;   // mean, rstd, mean2, rstd2: [4][4*kMaxQ] (index q*4+j)
  constexpr int kS = 4 * kMaxQ;

  const int L = kThr / Q;
  const int q = threadIdx.x % Q, lane = threadIdx.x / Q;
  const bool active = lane < L;
  const int step_s = L * Q;                       // smem stride (float4) between a thread's pixels
  const size_t step_g = (size_t)L * C4;           // global stride
  const size_t goff = ((size_t)n * a.HW + p0 + lane) * C4 + (size_t)cgi * Q + q;
  const int s0 = lane * Q + q;
  if (active) {
    const float4* gx = reinterpret_cast<const float4*>(a.x) + goff;
    float4* d = sx + s0;
    for (int p = lane; p < np; p += L, gx += step_g, d += step_s) cp_async16(d, gx);
    if (two) {
      const float4* gx2 = reinterpret_cast<const float4*>(a.x2) + goff;
      float4* d2 = sx2 + s0;
      for (int p = lane; p < np; p += L, gx2 += step_g, d2 += step_s) cp_async16(d2, gx2);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const float inv = 1.f / (float)a.HW;
  // ---- pass 0: mean; pass 1: variance about the mean (exact two-pass: the data sits in smem)
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
      const float4* d = sx + s0;
      const float4* d2 = sx2 + s0;
      if (pass == 0) {
        for (int p = lane; p < np; p += L, d += step_s, d2 += step_s) {
          const float4 v = *d;
          s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
          if (two) {
            const float4 u = *d2;
            s2[0] += u.x; s2[1] += u.y; s2[2] += u.z; s2[3] += u.w;
          }
        }
      } else {
        const float4 m = reinterpret_cast<const float4*>(stat)[q];
        const float4 m2 = reinterpret_cast<const float4*>(stat + 2 * kS)[q];
        for (int p = lane; p < np; p += L, d += step_s, d2 += step_s) {
          float4 v = *d;
          v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
          s[0] = fmaf(v.x, v.x, s[0]); s[1] = fmaf(v.y, v.y, s[1]);
          s[2] = fmaf(v.z, v.z, s[2]); s[3] = fmaf(v.w, v.w, s[3]);
          if (two) {
            float4 u = *d2;
            u.x -= m2.x; u.y -= m2.y; u.z -= m2.z; u.w -= m2.w;
            s2[0] = fmaf(u.x, u.x, s2[0]); s2[1] = fmaf(u.y, u.y, s2[1]);
            s2[2] = fmaf(u.z, u.z, s2[2]); s2[3] = fmaf(u.w, u.w, s2[3]);
          }
        }
      }
    }
    quad_reduce<float>(s, Q, L, wred, cpart + (pass * 2 + 0) * kS);
    if (two) quad_reduce<float>(s2, Q, L, wred, cpart + (pass * 2 + 1) * kS);
    cluster.sync();
    if (threadIdx.x < Q * 4) {      // fixed order over the cluster ranks: same result in every CTA
      float t = 0.f, t2 = 0.f;
      for (int r = 0; r < a.CS; ++r) {
        const float* rp = cluster.map_shared_rank(cpart, r);
        t += rp[(pass * 2 + 0) * kS + threadIdx.x];
        if (two) t2 += rp[(pass * 2 + 1) * kS + threadIdx.x];
      }
      if (pass == 0) {
        stat[threadIdx.x] = t * inv;
        stat[2 * kS + threadIdx.x] = t2 * inv;
      } else {
        stat[kS + threadIdx.x] = rsqrtf(t * inv + kEps);
        stat[3 * kS + threadIdx.x] = rsqrtf(t2 * inv + kEps);
      }
    }
    __syncthreads();
  }
  if (rank == 0 && threadIdx.x < Q * 4) {
    const size_t so = (size_t)n * a.C + (size_t)cgi * Q * 4 + threadIdx.x;
    a.mean[so] = stat[threadIdx.x];
    a.rstd[so] = stat[kS + threadIdx.x];
    if (two) {
      a.mean2[so] = stat[2 * kS + threadIdx.x];
      a.rstd2[so] = stat[3 * kS + threadIdx.x];
    }
  }

  // ---- normalise out of shared memory
  if (active) {
    const int c = (cgi * Q + q) * 4;
    const float4 m = reinterpret_cast<const float4*>(stat)[q];
    const float4 r = reinterpret_cast<const float4*>(stat + kS)[q];
    const float4 m2 = reinterpret_cast<const float4*>(stat + 2 * kS)[q];
    const float4 r2 = reinterpret_cast<const float4*>(stat + 3 * kS)[q];
    float4 gA = make_float4(1.f, 1.f, 1.f, 1.f), bA = make_float4(0.f, 0.f, 0.f, 0.f), gB = gA, bB = bA;
    if (a.gamma) {
      gA = *reinterpret_cast<const float4*>(a.gamma + c);
      bA = *reinterpret_cast<const float4*>(a.beta + c);
    }
    const bool hasB = a.hiB != nullptr;
    if (hasB) {
      gB = *reinterpret_cast<const float4*>(a.gammaB + c);
      bB = *reinterpret_cast<const float4*>(a.betaB + c);
    }
    const float slope = a.slope;
    const float4* d = sx + s0;
    const float4* d2 = sx2 + s0;
    size_t go = goff;
    for (int p = lane; p < np; p += L, d += step_s, d2 += step_s, go += step_g) {
      const float4 v = *d;
      const float xh[4] = {(v.x - m.x) * r.x, (v.y - m.y) * r.y, (v.z - m.z) * r.z, (v.w - m.w) * r.w};
      float o[4] = {fmaf(xh[0], gA.x, bA.x), fmaf(xh[1], gA.y, bA.y), fmaf(xh[2], gA.z, bA.z),
                    fmaf(xh[3], gA.w, bA.w)};
      if (MODE == 2) {
        const float4 u = *d2;
        o[0] += (u.x - m2.x) * r2.x; o[1] += (u.y - m2.y) * r2.y;
        o[2] += (u.z - m2.z) * r2.z; o[3] += (u.w - m2.w) * r2.w;
      } else if (MODE == 1) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(a.x2) + go);
        o[0] += u.x; o[1] += u.y; o[2] += u.z; o[3] += u.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = act_apply(o[j], slope);
      if (a.y) reinterpret_cast<float4*>(a.y)[go] = make_float4(o[0], o[1], o[2], o[3]);
      if (a.hiA) {
        uint2 h, l;
        split4<FMT>(o, h, l);
        reinterpret_cast<uint2*>(a.hiA)[go] = h;
        reinterpret_cast<uint2*>(a.loA)[go] = l;
      }
      if (hasB) {
        const float ob[4] = {act_apply(fmaf(xh[0], gB.x, bB.x), slope), act_apply(fmaf(xh[1], gB.y, bB.y), slope),
                             act_apply(fmaf(xh[2], gB.z, bB.z), slope), act_apply(fmaf(xh[3], gB.w, bB.w), slope)};
        uint2 h, l;
        split4<FMT>(ob, h, l);
        reinterpret_cast<uint2*>(a.hiB)[go] = h;
        reinterpret_cast<uint2*>(a.loB)[go] = l;
      }
    }
  }
  cluster.sync();     // nobody leaves while a peer may still read its partial sums
}

// =============================================================================== backward ==
// y = act(xhat*gamma + beta [+ res]),  xhat = (x - mean) * rstd.  Given dy:
//   g   = dy * act'(.)                                  (g_out: gradient w.r.t. res)
//   dx  = rstd * ( G - mean_hw(G) - xhat * mean_hw(G * xhat) ),  G = gamma * g
// With a second affine set (RefineNet blocks with a skip convolution normalise the same x twice,
// refine_net.py:45-62) G = gamma*g + gamma2*g2 and both pairs of sums are produced.
struct InBwdArgs {
  const float* dy;
  const float* dy2;
  const float* ymask;     // saved forward output (residual case); null: recompute xhat*gamma+beta
  const float* x;
  int HW, C, Q, CS, ppc;
  const float *mean, *rstd, *gamma, *beta, *gamma2, *beta2;
  float slope;
  const float* addend;
  float* dx;
  uint16_t *dx_hi, *dx_lo;
  float* g_out;
  float *sum_g, *sum_gx, *sum_g2, *sum_gx2;    // [N, C] (optional, for the affine gradients)
  float* colpart;                              // [N * CS][C] column sums of dx (optional)
  // optional: the forward activation act(xhat * gamma + beta) of this norm as bf16 hi/lo planes (and
  // of the second affine set) -- the x operand of the weight gradient of the convolution(s) the norm
  // feeds, re-derived here from the xhat this kernel computes anyway instead of by a separate pass
  uint16_t *ya_hi, *ya_lo, *yb_hi, *yb_lo;
};

template <bool DUAL>
__global__ void __launch_bounds__(kThr, 2)
in_bwd_fused_kernel(const InBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = blockIdx.x, cgi = blockIdx.y, n = blockIdx.z;
  const int Q = a.Q, C4 = a.C >> 2;
  const int p0 = rank * a.ppc;
  const int np = max(0, min(a.HW, p0 + a.ppc) - p0);
  const int slice = a.ppc * Q;
  float4* sg = reinterpret_cast<float4*>(smraw);                  // dy, then G
  float4* sx = sg + slice;                                        // x, then xhat
  double* wred = reinterpret_cast<double*>(sx + slice);           // [kThr*4]
  double* cpart = wred + kThr * 4;                                // [4 sums][4*kMaxQ] (index q*4+j)
  float* tot = reinterpret_cast<float*>(cpart + 4 * 4 * kMaxQ);   // A, B: [2][4*kMaxQ]
  constexpr int kS = 4 * kMaxQ;

  const int L = kThr / Q;
  const int q = threadIdx.x % Q, lane = threadIdx.x / Q;
  const bool active = lane < L;
  const int step_s = L * Q;
  const size_t step_g = (size_t)L * C4;
  const size_t goff = ((size_t)n * a.HW + p0 + lane) * C4 + (size_t)cgi * Q + q;
  const int s0 = lane * Q + q;
  if (active) {
    const float4* gd = reinterpret_cast<const float4*>(a.dy) + goff;
    const float4* gx = reinterpret_cast<const float4*>(a.x) + goff;
    float4 *d = sg + s0, *e = sx + s0;
    for (int p = lane; p < np; p += L, gd += step_g, gx += step_g, d += step_s, e += step_s) {
      cp_async16(d, gd);
      cp_async16(e, gx);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int c = (cgi * Q + q) * 4;
  const float slope = a.slope;
  float m[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0};
  float ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
  float ga2[4] = {0.f, 0.f, 0.f, 0.f}, be2[4] = {0.f, 0.f, 0.f, 0.f};
  // ---- phase 1: g, G and the sums.  Each thread sums its few pixels in fp32; everything across
  // threads, CTAs and (later) images is accumulated in fp64: sum(g*xhat) is a covariance and
  // cancels heavily.
  float f_g[4] = {0, 0, 0, 0}, f_gx[4] = {0, 0, 0, 0}, f_g2[4] = {0, 0, 0, 0}, f_gx2[4] = {0, 0, 0, 0};
  if (active) {
    const size_t so = (size_t)n * a.C + c;
    const float4 m4 = *reinterpret_cast<const float4*>(a.mean + so);
    const float4 r4 = *reinterpret_cast<const float4*>(a.rstd + so);
    m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
    r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w;
    if (a.gamma) {
      const float4 g4 = *reinterpret_cast<const float4*>(a.gamma + c);
      const float4 b4 = *reinterpret_cast<const float4*>(a.beta + c);
      ga[0] = g4.x; ga[1] = g4.y; ga[2] = g4.z; ga[3] = g4.w;
      be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
    }
    if (DUAL) {
      const float4 g4 = *reinterpret_cast<const float4*>(a.gamma2 + c);
      const float4 b4 = *reinterpret_cast<const float4*>(a.beta2 + c);
      ga2[0] = g4.x; ga2[1] = g4.y; ga2[2] = g4.z; ga2[3] = g4.w;
      be2[0] = b4.x; be2[1] = b4.y; be2[2] = b4.z; be2[3] = b4.w;
    }
    const bool has_mask = a.ymask != nullptr;
    const bool has_gout = a.g_out != nullptr;
    float4 *d = sg + s0, *e = sx + s0;
    size_t go = goff;
    for (int p = lane; p < np; p += L, d += step_s, e += step_s, go += step_g) {
      const float4 d4 = *d, x4 = *e;
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
      float pre[4], d2[4] = {0.f, 0.f, 0.f, 0.f};
      float xh[4], g[4], G[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xh[j] = (xv[j] - m[j]) * r[j];
        pre[j] = fmaf(xh[j], ga[j], be[j]);
      }
      if (has_mask) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.ymask) + go);
        pre[0] = t.x; pre[1] = t.y; pre[2] = t.z; pre[3] = t.w;
      }
      if (DUAL) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.dy2) + go);
        d2[0] = t.x; d2[1] = t.y; d2[2] = t.z; d2[3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        g[j] = dv[j] * act_deriv(pre[j], slope);
        f_g[j] += g[j];
        f_gx[j] = fmaf(g[j], xh[j], f_gx[j]);
        G[j] = ga[j] * g[j];
        if (DUAL) {
          const float g2 = d2[j] * act_deriv(fmaf(xh[j], ga2[j], be2[j]), slope);
          f_g2[j] += g2;
          f_gx2[j] = fmaf(g2, xh[j], f_gx2[j]);
          G[j] = fmaf(ga2[j], g2, G[j]);
        }
      }
      if (has_gout) reinterpret_cast<float4*>(a.g_out)[go] = make_float4(g[0], g[1], g[2], g[3]);
      *d = make_float4(G[0], G[1], G[2], G[3]);
      *e = make_float4(xh[0], xh[1], xh[2], xh[3]);
    }
  }
  {
    double t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = (double)f_g[j];
    quad_reduce<double>(t, Q, L, wred, cpart);
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = (double)f_gx[j];
    quad_reduce<double>(t, Q, L, wred, cpart + kS);
    if (DUAL) {
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = (double)f_g2[j];
      quad_reduce<double>(t, Q, L, wred, cpart + 2 * kS);
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = (double)f_gx2[j];
      quad_reduce<double>(t, Q, L, wred, cpart + 3 * kS);
    }
  }
  cluster.sync();
  if (threadIdx.x < Q * 4) {
    // thread t < 4Q totals channel cgi*CG + t
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    constexpr int nsum = DUAL ? 4 : 2;
    for (int rk = 0; rk < a.CS; ++rk) {
      const double* rp = cluster.map_shared_rank(cpart, rk);
#pragma unroll
      for (int k = 0; k < nsum; ++k) t[k] += rp[k * kS + threadIdx.x];
    }
    const int ch = cgi * Q * 4 + threadIdx.x;
    const double g1 = a.gamma ? (double)a.gamma[ch] : 1.0;
    const double g2 = DUAL ? (double)a.gamma2[ch] : 0.0;
    const double inv = 1.0 / (double)a.HW;
    tot[threadIdx.x] = (float)((g1 * t[0] + g2 * t[2]) * inv);
    tot[kS + threadIdx.x] = (float)((g1 * t[1] + g2 * t[3]) * inv);
    if (rank == 0 && a.sum_g) {
      const size_t o = (size_t)n * a.C + ch;
      a.sum_g[o] = (float)t[0];
      a.sum_gx[o] = (float)t[1];
      if (DUAL && a.sum_g2) {
        a.sum_g2[o] = (float)t[2];
        a.sum_gx2[o] = (float)t[3];
      }
    }
  }
  __syncthreads();

  // ---- phase 2: dx out of shared memory
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    const float4 A4 = reinterpret_cast<const float4*>(tot)[q];
    const float4 B4 = reinterpret_cast<const float4*>(tot + kS)[q];
    const float A[4] = {A4.x, A4.y, A4.z, A4.w}, B[4] = {B4.x, B4.y, B4.z, B4.w};
    const bool has_add = a.addend != nullptr, has_dx = a.dx != nullptr, has_pl = a.dx_hi != nullptr;
    const bool has_ya = a.ya_hi != nullptr;
    const float4 *d = sg + s0, *e = sx + s0;
    size_t go = goff;
    for (int p = lane; p < np; p += L, d += step_s, e += step_s, go += step_g) {
      const float4 G4 = *d, h4 = *e;
      const float G[4] = {G4.x, G4.y, G4.z, G4.w}, xh[4] = {h4.x, h4.y, h4.z, h4.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = r[j] * (G[j] - A[j] - xh[j] * B[j]);
      if (has_add) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.addend) + go);
        o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) cs[j] += o[j];
      if (has_dx) reinterpret_cast<float4*>(a.dx)[go] = make_float4(o[0], o[1], o[2], o[3]);
      if (has_pl) {
        uint2 h, l;
        split4<TC_BF16>(o, h, l);
        reinterpret_cast<uint2*>(a.dx_hi)[go] = h;
        reinterpret_cast<uint2*>(a.dx_lo)[go] = l;
      }
      if (has_ya) {
        float y[4];
        uint2 h, l;
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = act_apply(fmaf(xh[j], ga[j], be[j]), slope);
        split4<TC_BF16>(y, h, l);
        reinterpret_cast<uint2*>(a.ya_hi)[go] = h;
        reinterpret_cast<uint2*>(a.ya_lo)[go] = l;
        if (DUAL && a.yb_hi) {
#pragma unroll
          for (int j = 0; j < 4; ++j) y[j] = act_apply(fmaf(xh[j], ga2[j], be2[j]), slope);
          split4<TC_BF16>(y, h, l);
          reinterpret_cast<uint2*>(a.yb_hi)[go] = h;
          reinterpret_cast<uint2*>(a.yb_lo)[go] = l;
        }
      }
    }
  }
  if (a.colpart) {
    float* fred = reinterpret_cast<float*>(wred);
    float* fout = fred + kThr * 4;
    quad_reduce<float>(cs, Q, L, fred, fout);
    if (threadIdx.x < Q * 4)
      a.colpart[((size_t)n * a.CS + rank) * a.C + cgi * Q * 4 + threadIdx.x] = fout[threadIdx.x];
  }
  cluster.sync();
}

// ---- streaming variant ---------------------------------------------------------------------
// Large maps (RefineNet level 0 / 1: 9216 and 2304 pixels) do not fit a cluster's shared memory
// unless every CTA takes ~150 KB, i.e. ONE CTA per SM whose load, reduce, cluster-sync and store
// phases run back to back with nothing else to hide them (measured: 28 % of the DRAM peak).  This
// variant keeps the work split (cluster = one image x channel group, fixed-order DSMEM reduction,
// same arithmetic per element) but stages nothing: phase 1 streams dy and x from HBM through
// registers, phase 2 reads them AGAIN -- the cluster touched them microseconds ago and the ~20
// clusters in flight hold ~25 MB, so the second read is served by the 126 MB L2 -- and recomputes
// g / xhat instead of loading them from shared memory.  With ~27 KB of reduction scratch per CTA
// two CTAs share an SM and one cluster's stores overlap another's loads.
template <bool DUAL>
struct BwdElem {
  float xh[4], g[4], G[4];
};

template <bool DUAL>
__device__ __forceinline__ void bwd_elem(const float4 d4, const float4 x4, const float4 y4,
                                         const bool has_mask, const float4 e4, const float* m,
                                         const float* r, const float* ga, const float* be,
                                         const float* ga2, const float* be2, const float slope,
                                         BwdElem<DUAL>& o, float* g2) {
  const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
  const float yv[4] = {y4.x, y4.y, y4.z, y4.w}, ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o.xh[j] = (xv[j] - m[j]) * r[j];
    const float pre = has_mask ? yv[j] : fmaf(o.xh[j], ga[j], be[j]);
    o.g[j] = dv[j] * act_deriv(pre, slope);
    o.G[j] = ga[j] * o.g[j];
    if (DUAL) {
      g2[j] = ev[j] * act_deriv(fmaf(o.xh[j], ga2[j], be2[j]), slope);
      o.G[j] = fmaf(ga2[j], g2[j], o.G[j]);
    }
  }
}

template <int N>
__device__ __forceinline__ void cp_async_wait_pending() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// D = depth of the per-thread cp.async ring the element loops read through: every thread keeps D
// iterations (x up to five tensors) of ITS OWN 16-byte pieces in flight in shared memory -- loads
// held in registers (an unrolled __ldg loop under the 64-register cap of two CTAs per SM) kept only
// one or two iterations in flight and the kernel waited on the long scoreboard at 36 % of the DRAM
// peak.  No thread reads another thread's pieces, so cp.async.wait_group is the only synchronisation.
template <bool DUAL, int D>
__global__ void __launch_bounds__(kThr, 2)
in_bwd_stream_kernel(const InBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = blockIdx.x, cgi = blockIdx.y, n = blockIdx.z;
  const int Q = a.Q, C4 = a.C >> 2;
  const int p0 = rank * a.ppc;
  const int np = max(0, min(a.HW, p0 + a.ppc) - p0);
  double* wred = reinterpret_cast<double*>(smraw);                // [kThr*4]
  double* cpart = wred + kThr * 4;                                // [4 sums][4*kMaxQ]
  float* tot = reinterpret_cast<float*>(cpart + 4 * 4 * kMaxQ);   // A, B: [2][4*kMaxQ]
  float4* ring = reinterpret_cast<float4*>(tot + 2 * 4 * kMaxQ);  // [D][tensors][kThr]
  constexpr int kS = 4 * kMaxQ;

  const int L = kThr / Q;
  const int q = threadIdx.x % Q, lane = threadIdx.x / Q;
  const bool active = lane < L;
  const size_t step_g = (size_t)L * C4;
  const size_t goff = ((size_t)n * a.HW + p0 + lane) * C4 + (size_t)cgi * Q + q;
  const int c = (cgi * Q + q) * 4;
  const float slope = a.slope;
  const float4* gd = reinterpret_cast<const float4*>(a.dy);
  const float4* gx = reinterpret_cast<const float4*>(a.x);
  const float4* gy = reinterpret_cast<const float4*>(a.ymask);
  const float4* ge = reinterpret_cast<const float4*>(a.dy2);
  const float4* gt = reinterpret_cast<const float4*>(a.addend);
  const bool has_mask = a.ymask != nullptr;
  const bool has_add = a.addend != nullptr;
  // ring layout: tensor slots dy, x, [ymask], [dy2], [addend] (the last in phase 2 only)
  const int i_mask = 2, i_e = 2 + (has_mask ? 1 : 0), i_add = i_e + (DUAL ? 1 : 0);
  const int TT = i_add + (has_add ? 1 : 0);
  float4* mine = ring + threadIdx.x;
  auto issue = [&](int slot, size_t go, bool with_add) {
    float4* d = mine + (size_t)slot * TT * kThr;
    cp_async16(d, gd + go);
    cp_async16(d + kThr, gx + go);
    if (has_mask) cp_async16(d + i_mask * kThr, gy + go);
    if (DUAL) cp_async16(d + i_e * kThr, ge + go);
    if (with_add) cp_async16(d + i_add * kThr, gt + go);
  };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float m[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0};
  float ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
  float ga2[4] = {0.f, 0.f, 0.f, 0.f}, be2[4] = {0.f, 0.f, 0.f, 0.f};
  float f_g[4] = {0, 0, 0, 0}, f_gx[4] = {0, 0, 0, 0}, f_g2[4] = {0, 0, 0, 0}, f_gx2[4] = {0, 0, 0, 0};
  if (active) {
    // ---- phase 1: the sums (fp32 per thread over its few pixels, fp64 across threads and CTAs,
    // exactly as the staged kernel)
    int pi = lane;
    size_t gi = goff;
#pragma unroll
    for (int k = 0; k < D; ++k, pi += L, gi += step_g) {
      if (pi < np) issue(k, gi, false);
      cp_async_commit();
    }
    const size_t so = (size_t)n * a.C + c;
    const float4 m4 = *reinterpret_cast<const float4*>(a.mean + so);
    const float4 r4 = *reinterpret_cast<const float4*>(a.rstd + so);
    m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
    r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w;
    if (a.gamma) {
      const float4 g4 = *reinterpret_cast<const float4*>(a.gamma + c);
      const float4 b4 = *reinterpret_cast<const float4*>(a.beta + c);
      ga[0] = g4.x; ga[1] = g4.y; ga[2] = g4.z; ga[3] = g4.w;
      be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
    }
    if (DUAL) {
      const float4 g4 = *reinterpret_cast<const float4*>(a.gamma2 + c);
      const float4 b4 = *reinterpret_cast<const float4*>(a.beta2 + c);
      ga2[0] = g4.x; ga2[1] = g4.y; ga2[2] = g4.z; ga2[3] = g4.w;
      be2[0] = b4.x; be2[1] = b4.y; be2[2] = b4.z; be2[3] = b4.w;
    }
    const bool has_gout = a.g_out != nullptr;
    size_t go = goff;
    int slot = 0;
#pragma unroll 1
    for (int p = lane; p < np; p += L, go += step_g, pi += L, gi += step_g) {
      cp_async_wait_pending<D - 1>();
      const float4* d = mine + (size_t)slot * TT * kThr;
      const float4 d4 = d[0], x4 = d[kThr];
      const float4 y4 = has_mask ? d[i_mask * kThr] : z4;
      const float4 e4 = DUAL ? d[i_e * kThr] : z4;
      if (pi < np) issue(slot, gi, false);       // refill the slot just read (iteration + D)
      cp_async_commit();
      slot = slot + 1 == D ? 0 : slot + 1;
      BwdElem<DUAL> o;
      float g2[4] = {0.f, 0.f, 0.f, 0.f};
      bwd_elem<DUAL>(d4, x4, y4, has_mask, e4, m, r, ga, be, ga2, be2, slope, o, g2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f_g[j] += o.g[j];
        f_gx[j] = fmaf(o.g[j], o.xh[j], f_gx[j]);
        if (DUAL) {
          f_g2[j] += g2[j];
          f_gx2[j] = fmaf(g2[j], o.xh[j], f_gx2[j]);
        }
      }
      if (has_gout)
        reinterpret_cast<float4*>(a.g_out)[go] = make_float4(o.g[0], o.g[1], o.g[2], o.g[3]);
    }
    cp_async_wait_pending<0>();
    // phase 2's first D iterations: on their way while the sums cross the cluster
    pi = lane;
    gi = goff;
#pragma unroll
    for (int k = 0; k < D; ++k, pi += L, gi += step_g) {
      if (pi < np) issue(k, gi, has_add);
      cp_async_commit();
    }
  }
  {
    double t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = (double)f_g[j];
    quad_reduce<double>(t, Q, L, wred, cpart);
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = (double)f_gx[j];
    quad_reduce<double>(t, Q, L, wred, cpart + kS);
    if (DUAL) {
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = (double)f_g2[j];
      quad_reduce<double>(t, Q, L, wred, cpart + 2 * kS);
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = (double)f_gx2[j];
      quad_reduce<double>(t, Q, L, wred, cpart + 3 * kS);
    }
  }
  cluster.sync();
  if (threadIdx.x < Q * 4) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    constexpr int nsum = DUAL ? 4 : 2;
    for (int rk = 0; rk < a.CS; ++rk) {
      const double* rp = cluster.map_shared_rank(cpart, rk);
#pragma unroll
      for (int k = 0; k < nsum; ++k) t[k] += rp[k * kS + threadIdx.x];
    }
    const int ch = cgi * Q * 4 + threadIdx.x;
    const double g1 = a.gamma ? (double)a.gamma[ch] : 1.0;
    const double g2 = DUAL ? (double)a.gamma2[ch] : 0.0;
    const double inv = 1.0 / (double)a.HW;
    tot[threadIdx.x] = (float)((g1 * t[0] + g2 * t[2]) * inv);
    tot[kS + threadIdx.x] = (float)((g1 * t[1] + g2 * t[3]) * inv);
    if (rank == 0 && a.sum_g) {
      const size_t o = (size_t)n * a.C + ch;
      a.sum_g[o] = (float)t[0];
      a.sum_gx[o] = (float)t[1];
      if (DUAL && a.sum_g2) {
        a.sum_g2[o] = (float)t[2];
        a.sum_gx2[o] = (float)t[3];
      }
    }
  }
  __syncthreads();

  // ---- phase 2: dx from a second read of dy / x (L2)
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    const float4 A4 = reinterpret_cast<const float4*>(tot)[q];
    const float4 B4 = reinterpret_cast<const float4*>(tot + kS)[q];
    const float A[4] = {A4.x, A4.y, A4.z, A4.w}, B[4] = {B4.x, B4.y, B4.z, B4.w};
    const bool has_dx = a.dx != nullptr, has_pl = a.dx_hi != nullptr;
    const bool has_ya = a.ya_hi != nullptr;
    size_t go = goff;
    int pi = lane + D * L;
    size_t gi = goff + (size_t)D * step_g;
    int slot = 0;
#pragma unroll 1
    for (int p = lane; p < np; p += L, go += step_g, pi += L, gi += step_g) {
      cp_async_wait_pending<D - 1>();
      const float4* d = mine + (size_t)slot * TT * kThr;
      const float4 d4 = d[0], x4 = d[kThr];
      const float4 y4 = has_mask ? d[i_mask * kThr] : z4;
      const float4 e4 = DUAL ? d[i_e * kThr] : z4;
      const float4 t4 = has_add ? d[i_add * kThr] : z4;
      if (pi < np) issue(slot, gi, has_add);
      cp_async_commit();
      slot = slot + 1 == D ? 0 : slot + 1;
      BwdElem<DUAL> e;
      float g2[4];
      bwd_elem<DUAL>(d4, x4, y4, has_mask, e4, m, r, ga, be, ga2, be2, slope, e, g2);
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = r[j] * (e.G[j] - A[j] - e.xh[j] * B[j]);
      if (has_add) {
        o[0] += t4.x; o[1] += t4.y; o[2] += t4.z; o[3] += t4.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) cs[j] += o[j];
      if (has_dx) reinterpret_cast<float4*>(a.dx)[go] = make_float4(o[0], o[1], o[2], o[3]);
      if (has_pl) {
        uint2 h, l;
        split4<TC_BF16>(o, h, l);
        reinterpret_cast<uint2*>(a.dx_hi)[go] = h;
        reinterpret_cast<uint2*>(a.dx_lo)[go] = l;
      }
      if (has_ya) {
        float y[4];
        uint2 h, l;
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = act_apply(fmaf(e.xh[j], ga[j], be[j]), slope);
        split4<TC_BF16>(y, h, l);
        reinterpret_cast<uint2*>(a.ya_hi)[go] = h;
        reinterpret_cast<uint2*>(a.ya_lo)[go] = l;
        if (DUAL && a.yb_hi) {
#pragma unroll
          for (int j = 0; j < 4; ++j) y[j] = act_apply(fmaf(e.xh[j], ga2[j], be2[j]), slope);
          split4<TC_BF16>(y, h, l);
          reinterpret_cast<uint2*>(a.yb_hi)[go] = h;
          reinterpret_cast<uint2*>(a.yb_lo)[go] = l;
        }
      }
    }
    cp_async_wait_pending<0>();
  }
  if (a.colpart) {
    float* fred = reinterpret_cast<float*>(wred);
    float* fout = fred + kThr * 4;
    quad_reduce<float>(cs, Q, L, fred, fout);
    if (threadIdx.x < Q * 4)
      a.colpart[((size_t)n * a.CS + rank) * a.C + cgi * Q * 4 + threadIdx.x] = fout[threadIdx.x];
  }
  cluster.sync();
}

__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// dy (fp32) -> bf16 hi/lo planes + per-block column sums, one read of dy.
// grid: blocks of `rpb` rows; block 256 = Q channel quads x L row lanes (Q = C/4 <= 256).
__global__ void __launch_bounds__(kThr)
split_colsum_kernel(const float* __restrict__ dy, long long rows, int C, int rpb,
                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                    float* __restrict__ colpart) {
  __shared__ float4 red[kThr];
  const int Q = C >> 2;
  const int L = kThr / Q;
  const int q = threadIdx.x % Q, lane = threadIdx.x / Q;
  const long long r0 = (long long)blockIdx.x * rpb;
  const long long r1 = min(rows, r0 + rpb);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < L) {
    for (long long rr = r0 + lane; rr < r1; rr += L) {
      const size_t i = (size_t)rr * Q + q;
      const float4 v = __ldg(reinterpret_cast<const float4*>(dy) + i);
      const float o[4] = {v.x, v.y, v.z, v.w};
      uint2 h, l;
      split4<TC_BF16>(o, h, l);
      reinterpret_cast<uint2*>(hi)[i] = h;
      reinterpret_cast<uint2*>(lo)[i] = l;
      s = f4add(s, v);
    }
  }
  if (!colpart) return;
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < Q) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < L; ++l) t = f4add(t, red[l * Q + threadIdx.x]);
    *reinterpret_cast<float4*>(colpart + (size_t)blockIdx.x * C + 4 * threadIdx.x) = t;
  }
}

// out[c] (+)= sum over rows of part[rows][C]; one warp per channel, lanes stride over rows in a
// fixed order, fp64 partials
__global__ void __launch_bounds__(256)
rowsum_kernel(const float* __restrict__ part, int rows, int C, float* __restrict__ out,
              float* __restrict__ out2, int accumulate) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0;
  for (int rr = lane; rr < rows; rr += 32) s += (double)part[(size_t)rr * C + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (out) out[c] = accumulate ? out[c] + (float)s : (float)s;
    if (out2) out2[c] = accumulate ? out2[c] + (float)s : (float)s;
  }
}

// up to five such reductions (dgamma, dbeta, the second affine set's, the bias column sums) that
// follow one in_bwd_fused launch, as ONE launch: blockIdx.y = job
struct RowsumJobs {
  int n;
  const float* part[5];
  int rows[5];
  float* out[5];
  float* out2[5];
};
__global__ void __launch_bounds__(256)
rowsum_multi_kernel(const RowsumJobs jobs, int C, int accumulate) {
  const int j = blockIdx.y;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const float* part = jobs.part[j];
  const int rows = jobs.rows[j];
  double s = 0.0;
  for (int rr = lane; rr < rows; rr += 32) s += (double)part[(size_t)rr * C + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    float* out = jobs.out[j];
    float* out2 = jobs.out2[j];
    if (out) out[c] = accumulate ? out[c] + (float)s : (float)s;
    if (out2) out2[c] = accumulate ? out2[c] + (float)s : (float)s;
  }
}

size_t fwd_smem(const FusedPlan& p, int tensors) {
  return (size_t)p.ppc * p.Q * 16 * tensors +
         (size_t)(kThr * 4 + 4 * 4 * kMaxQ + 4 * 4 * kMaxQ) * sizeof(float);
}
size_t bwd_smem(const FusedPlan& p) {
  return (size_t)p.ppc * p.Q * 16 * 2 + (size_t)(kThr * 4 + 4 * 4 * kMaxQ) * sizeof(double) +
         2 * 4 * kMaxQ * sizeof(float);
}

template <typename K, typename A>
int launch_cluster(K kernel, dim3 grid, int cs, size_t smem, const A& args, cudaStream_t s) {
  EVE_TRY(ensure_dynamic_smem((const void*)kernel, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThr, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cs > 8) {
    static std::mutex mu;
    static std::map<const void*, bool> allowed;
    std::lock_guard<std::mutex> lk(mu);
    if (!allowed[(const void*)kernel]) {
      EVE_CUDA(cudaFuncSetAttribute((const void*)kernel,
                                    cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      allowed[(const void*)kernel] = true;
    }
  }
  EVE_CUDA(cudaLaunchKernelEx(&cfg, kernel, args));
  count_launch();
  return EVE_OK;
}

template <int FMT>
int launch_fwd(int mode, dim3 grid, int cs, size_t smem, const InFwdArgs& a, cudaStream_t s) {
  if (mode == 2) return launch_cluster(in_fwd_fused_kernel<FMT, 2>, grid, cs, smem, a, s);
  if (mode == 1) return launch_cluster(in_fwd_fused_kernel<FMT, 1>, grid, cs, smem, a, s);
  return launch_cluster(in_fwd_fused_kernel<FMT, 0>, grid, cs, smem, a, s);
}

}  // namespace

bool in_fused_supported(int HW, int C, int tensors) {
  FusedPlan p;
  return plan_fused(C, HW, tensors, p);
}

int in_fwd_fused(const float* x, int N, int HW, int C, const float* x2, int x2_mode,
                 const float* gamma, const float* beta, const float* gammaB, const float* betaB,
                 int act, int fmt, float* mean, float* rstd, float* mean2, float* rstd2, float* y,
                 void* hiA, void* loA, void* hiB, void* loB, cudaStream_t s) {
  if (N == 0) return EVE_OK;
  const int tensors = x2_mode == 2 ? 2 : 1;
  FusedPlan p;
  EVE_REQUIRE(plan_fused(C, HW, tensors, p), EVE_ERR_SHAPE,
              "in_fwd_fused: unsupported shape HW=%d C=%d", HW, C);
  EVE_REQUIRE(x && mean && rstd, EVE_ERR_NULL, "in_fwd_fused: NULL pointer");
  EVE_REQUIRE(x2_mode == 0 || x2, EVE_ERR_NULL, "in_fwd_fused: x2 is NULL");
  EVE_REQUIRE(x2_mode != 2 || (mean2 && rstd2), EVE_ERR_NULL, "in_fwd_fused: mean2/rstd2 are NULL");
  EVE_REQUIRE((gamma == nullptr) == (beta == nullptr), EVE_ERR_NULL, "in_fwd_fused: gamma/beta");
  EVE_REQUIRE(!hiB || (gammaB && betaB && loB), EVE_ERR_NULL, "in_fwd_fused: second affine set");
  InFwdArgs a;
  a.x = x; a.x2 = x2;
  a.HW = HW; a.C = C; a.Q = p.Q; a.CS = p.CS; a.ppc = p.ppc;
  a.gamma = gamma; a.beta = beta; a.gammaB = gammaB; a.betaB = betaB;
  a.slope = act_slope(act);
  a.mean = mean; a.rstd = rstd; a.mean2 = mean2; a.rstd2 = rstd2;
  a.y = y;
  a.hiA = (uint16_t*)hiA; a.loA = (uint16_t*)loA; a.hiB = (uint16_t*)hiB; a.loB = (uint16_t*)loB;
  const size_t smem = fwd_smem(p, tensors);
  dim3 grid(p.CS, C / p.CG, N);
  if (fmt == TC_BF16) return launch_fwd<TC_BF16>(x2_mode, grid, p.CS, smem, a, s);
  return launch_fwd<TC_F16>(x2_mode, grid, p.CS, smem, a, s);
}

size_t in_bwd_fused_scratch_floats(int N, int HW, int C) {
  FusedPlan p;
  if (!plan_fused(C, HW, 2, p)) return 0;
  return (size_t)4 * N * C + (size_t)N * p.CS * C;
}

int in_bwd_fused(const float* dy, const float* dy2, const float* ymask, const float* x, int N,
                 int HW, int C, const float* mean, const float* rstd, const float* gamma,
                 const float* beta, const float* gamma2, const float* beta2, int act,
                 const float* addend, float* dx, void* dx_hi, void* dx_lo, float* g_out,
                 float* dgamma, float* dbeta, float* dgamma2, float* dbeta2, float* dbias,
                 float* dbias2, bool accumulate, float* scratch, cudaStream_t s, void* ya_hi,
                 void* ya_lo, void* yb_hi, void* yb_lo) {
  if (N == 0) return EVE_OK;
  FusedPlan p;
  EVE_REQUIRE(plan_fused(C, HW, 2, p), EVE_ERR_SHAPE,
              "in_bwd_fused: unsupported shape HW=%d C=%d", HW, C);
  EVE_REQUIRE(!ya_hi || (ya_lo && !ymask), EVE_ERR_NULL,
              "in_bwd_fused: activation planes need ya_lo and the recomputed pre-activation");
  EVE_REQUIRE(!yb_hi || (yb_lo && ya_hi && dy2), EVE_ERR_NULL, "in_bwd_fused: second plane set");
  EVE_REQUIRE(dy && x && mean && rstd && scratch, EVE_ERR_NULL, "in_bwd_fused: NULL pointer");
  EVE_REQUIRE(!dy2 || (gamma2 && beta2 && gamma), EVE_ERR_NULL, "in_bwd_fused: second affine set");
  EVE_REQUIRE(!dx_hi || dx_lo, EVE_ERR_NULL, "in_bwd_fused: dx_lo is NULL");
  // the streaming form reads dy / x / ymask / dy2 twice: not when g_out overwrites one of them
  const bool inplace = g_out && (g_out == dy || g_out == dy2 || g_out == ymask || g_out == x);
  const int sopt = get_option(OPT_IN_STREAM);
  const bool stream = !inplace && (sopt == 2 || (sopt == 1 && p.one_cta));
  if (stream) {
    // Nothing is staged, so the cluster need not be as large as the shared-memory plan wants, and
    // smaller is faster (measured, tools/bench_in.py: 8 -> 2 CTAs +9-13 % at 9216 pixels, +30-45 % at
    // 1024-2304 pixels; in the training step 34.52 -> 33.85 ms with 2 everywhere, 34.10 with 4, 34.13
    // with 1): fewer cluster-wide barriers and per-CTA prologues, a longer steady state of the ring,
    // and clusters that pack the GPCs' CTA slots without remainder.  Never a function of N.
    int cs = stream_cluster();
    if (!cs) cs = 2;
    if (HW >= 8192 && stream_cluster_big()) cs = stream_cluster_big();
    if (cs < p.CS) {
      p.CS = cs;
      p.ppc = cdiv(HW, p.CS);
    }
  }
  InBwdArgs a;
  a.dy = dy; a.dy2 = dy2; a.ymask = ymask; a.x = x;
  a.HW = HW; a.C = C; a.Q = p.Q; a.CS = p.CS; a.ppc = p.ppc;
  a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta; a.gamma2 = gamma2; a.beta2 = beta2;
  a.slope = act_slope(act);
  a.addend = addend; a.dx = dx; a.dx_hi = (uint16_t*)dx_hi; a.dx_lo = (uint16_t*)dx_lo;
  a.g_out = g_out;
  a.ya_hi = (uint16_t*)ya_hi; a.ya_lo = (uint16_t*)ya_lo;
  a.yb_hi = (uint16_t*)yb_hi; a.yb_lo = (uint16_t*)yb_lo;
  const size_t nc = (size_t)N * C;
  const bool affine = gamma && dgamma;
  a.sum_g = affine ? scratch : nullptr;
  a.sum_gx = affine ? scratch + nc : nullptr;
  a.sum_g2 = affine && dy2 ? scratch + 2 * nc : nullptr;
  a.sum_gx2 = affine && dy2 ? scratch + 3 * nc : nullptr;
  a.colpart = (dbias || dbias2) ? scratch + 4 * nc : nullptr;
  dim3 grid(p.CS, C / p.CG, N);
  if (stream) {
    // ring depth: tensors x depth >= 8 sixteen-byte pieces per thread in flight, two CTAs per SM
    const int tensors = 2 + (ymask ? 1 : 0) + (dy2 ? 1 : 0) + (addend ? 1 : 0);
    const int depth = tensors <= 2 ? 4 : (tensors == 3 ? 3 : 2);
    const size_t smem = (size_t)(kThr * 4 + 4 * 4 * kMaxQ) * sizeof(double) + 2 * 4 * kMaxQ * sizeof(float) +
                        (size_t)depth * tensors * kThr * 16;
    if (dy2) {
      if (depth == 3) EVE_TRY(launch_cluster(in_bwd_stream_kernel<true, 3>, grid, p.CS, smem, a, s));
      else EVE_TRY(launch_cluster(in_bwd_stream_kernel<true, 2>, grid, p.CS, smem, a, s));
    } else if (depth == 4) {
      EVE_TRY(launch_cluster(in_bwd_stream_kernel<false, 4>, grid, p.CS, smem, a, s));
    } else if (depth == 3) {
      EVE_TRY(launch_cluster(in_bwd_stream_kernel<false, 3>, grid, p.CS, smem, a, s));
    } else {
      EVE_TRY(launch_cluster(in_bwd_stream_kernel<false, 2>, grid, p.CS, smem, a, s));
    }
  } else if (dy2) {
    EVE_TRY(launch_cluster(in_bwd_fused_kernel<true>, grid, p.CS, bwd_smem(p), a, s));
  } else {
    EVE_TRY(launch_cluster(in_bwd_fused_kernel<false>, grid, p.CS, bwd_smem(p), a, s));
  }
  const int acc = accumulate ? 1 : 0;
  RowsumJobs jobs;
  jobs.n = 0;
  auto add = [&](const float* part, int rows, float* out, float* out2) {
    jobs.part[jobs.n] = part; jobs.rows[jobs.n] = rows; jobs.out[jobs.n] = out; jobs.out2[jobs.n] = out2;
    ++jobs.n;
  };
  if (affine) {
    add(a.sum_gx, N, dgamma, nullptr);
    add(a.sum_g, N, dbeta, nullptr);
    if (dy2 && dgamma2) {
      add(a.sum_gx2, N, dgamma2, nullptr);
      add(a.sum_g2, N, dbeta2, nullptr);
    }
  }
  if (a.colpart) add(a.colpart, N * p.CS, dbias, dbias2);
  if (jobs.n > 0) {
    rowsum_multi_kernel<<<dim3(cdiv(C, 8), jobs.n), 256, 0, s>>>(jobs, C, acc);
    EVE_LAUNCH_CHECK();
  }
  return EVE_OK;
}

size_t split_colsum_scratch_floats(long long rows, int C) {
  return (size_t)(cdiv(rows, 256) < 2048 ? cdiv(rows, 256) : 2048) * C + C;
}

// dy -> bf16 planes; dbias (+)= column sums of dy when dbias != null
int split_colsum(const float* dy, long long rows, int C, void* hi, void* lo, float* dbias,
                 float* dbias2, bool accumulate, float* scratch, cudaStream_t s) {
  EVE_REQUIRE(C % 4 == 0 && C <= 1024, EVE_ERR_SHAPE, "split_colsum: C=%d", C);
  if (rows == 0) return EVE_OK;
  int nblk = cdiv(rows, 256);
  if (nblk > 2048) nblk = 2048;
  const int rpb = cdiv(rows, nblk);
  nblk = cdiv(rows, rpb);
  split_colsum_kernel<<<nblk, kThr, 0, s>>>(dy, rows, C, rpb, (uint16_t*)hi, (uint16_t*)lo,
                                            (dbias || dbias2) ? scratch : nullptr);
  EVE_LAUNCH_CHECK();
  if (dbias || dbias2) {
    rowsum_kernel<<<cdiv(C, 8), 256, 0, s>>>(scratch, nblk, C, dbias, dbias2, accumulate ? 1 : 0);
    EVE_LAUNCH_CHECK();
  }
  return EVE_OK;
}

}  // namespace eve
