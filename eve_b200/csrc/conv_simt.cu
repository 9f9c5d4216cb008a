// fp32 CUDA-core implicit-GEMM convolution: forward, data-gradient, weight-gradient.
//
// Used for (a) the layers whose arithmetic intensity is below the tensor-core ridge
// (stem 7x7 with Cin=3, RefineNet's 16/32-channel 72x128 and 36x64 levels, 1x1
// down-sample / skip convs), (b) every linear layer (a 1x1 conv on a 1x1 image) and
// (c) as the exact-fp32 path that the tcgen05 kernels in conv_tc.cu are checked against.
//
// Reference call sites replaced: every nn.Conv2d / nn.Linear on the hot path
// (torchvision ResNet via eye_net.py:48-50,106; refine_net.py:45-62,213-224;
// common.py:338,362,395-398) and their autograd backward.
//
// GEMM view, all operands NHWC fp32:
//   fwd  : C[m=(n,oy,ox)][co] = sum_{k=(r,q,ci)} X[n, oy*s+r-p, ox*s+q-p, ci] * Wf[k][co]
//   dgrad: C[m=(n,y,x)][ci]   = sum_{k=(r,q,co)} dY[n,(y+p-r)/s,(x+p-q)/s,co] * Wd[k][ci]
//   wgrad: C[co][(r,q,ci)]    = sum_{pix}        dY[pix][co] * X[gather(pix,r,q)][ci]
#include "common.cuh"

namespace eve {

namespace {

constexpr int BK = 16;
constexpr int NT = 256;

struct GatherArgs {
  const float* src;     // [N, Hs, Ws, lds] (first Cs channels used)
  const float* wmat;    // [K][Nn]
  const float* bias;    // [Nn] or null
  const float* addend;  // [M][ldo] or null
  float* out;           // [M][ldo]
  int Hs, Ws, Cs, lds;
  int Ho, Wo, Nn, KH, KW;
  int num, den, dr, base;  // ys*den = y*num + r*dr + base
  int M, K, ldo;
};

// BM = TM * (NT / (BN/TN)) rows per CTA.
template <int BN, int TM, int TN>
__global__ void __launch_bounds__(NT) igemm_gather_kernel(GatherArgs a) {
  constexpr int CT = BN / TN;        // threads along n
  constexpr int RT = NT / CT;        // threads along m
  constexpr int BM = RT * TM;
  constexpr int AL = BM / 16;        // A elements each thread loads per k-tile
  constexpr int BL = (BK * BN) / NT; // B elements each thread loads per k-tile
  static_assert(BL >= 1, "tile too small");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A loader state: this thread always loads column kk of rows (t/16 + 16*i)
  const int kk = t & 15;
  const int rbase = t >> 4;
  int pix0[AL];   // n*Hs*Ws
  int yx[AL];     // (y*num+base) in hi16 (biased), (x*num+base) in lo16 (biased)
#pragma unroll
  for (int i = 0; i < AL; ++i) {
    int m = m0 + rbase + 16 * i;
    if (m < a.M) {
      int x = m % a.Wo;
      int tmp = m / a.Wo;
      int y = tmp % a.Ho;
      int n = tmp / a.Ho;
      pix0[i] = n * a.Hs * a.Ws;
      yx[i] = ((y * a.num + a.base + 1024) << 16) | (x * a.num + a.base + 1024);
    } else {
      pix0[i] = -1;
      yx[i] = 0;
    }
  }
  // ---- B loader: column nn of rows (t/BN + (NT/BN)*j)
  const int nn = t % BN;
  const int kb0 = t / BN;
  constexpr int KSTEP = NT / BN;

  float ra[AL], rb[BL];
  auto load_tile = [&](int k0) {
    int k = k0 + kk;
    bool kvalid = k < a.K;
    int tap = kvalid ? k / a.Cs : 0;
    int ci = k - tap * a.Cs;
    int r = tap / a.KW;
    int q = tap - r * a.KW;
    int ro = r * a.dr, qo = q * a.dr;
#pragma unroll
    for (int i = 0; i < AL; ++i) {
      float v = 0.f;
      if (kvalid && pix0[i] >= 0) {
        int ys = (yx[i] >> 16) - 1024 + ro;
        int xs = (yx[i] & 0xffff) - 1024 + qo;
        bool ok = ys >= 0 && xs >= 0;
        if (a.den != 1) {
          ok = ok && (ys % a.den == 0) && (xs % a.den == 0);
          ys /= a.den;
          xs /= a.den;
        }
        if (ok && ys < a.Hs && xs < a.Ws)
          v = __ldg(a.src + (size_t)(pix0[i] + ys * a.Ws + xs) * a.lds + ci);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int j = 0; j < BL; ++j) {
      int kr = k0 + kb0 + KSTEP * j;
      int n = n0 + nn;
      rb[j] = (kr < a.K && n < a.Nn) ? __ldg(a.wmat + (size_t)kr * a.Nn + n) : 0.f;
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = t % CT;
  const int ty = t / CT;
  const int ntiles = (a.K + BK - 1) / BK;
  load_tile(0);
  for (int kt = 0; kt < ntiles; ++kt) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < AL; ++i) As[kk][rbase + 16 * i] = ra[i];
#pragma unroll
    for (int j = 0; j < BL; ++j) Bs[kb0 + KSTEP * j][nn] = rb[j];
    __syncthreads();
    if (kt + 1 < ntiles) load_tile((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
        av[i] = v.x; av[i + 1] = v.y; av[i + 2] = v.z; av[i + 3] = v.w;
      }
      if (TN == 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[k][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= a.Nn) continue;
      float v = acc[i][j];
      if (a.bias) v += __ldg(a.bias + n);
      size_t o = (size_t)m * a.ldo + n;
      if (a.addend) v += __ldg(a.addend + o);
      a.out[o] = v;
    }
  }
}

template <int BN, int TM, int TN>
int launch_gather(const GatherArgs& a, cudaStream_t s) {
  constexpr int BM = (NT / (BN / TN)) * TM;
  dim3 grid(cdiv(a.M, BM), cdiv(a.Nn, BN));
  igemm_gather_kernel<BN, TM, TN><<<grid, NT, 0, s>>>(a);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int dispatch_gather(const GatherArgs& a, cudaStream_t s) {
  EVE_REQUIRE(a.Hs < 30000 && a.Ws < 30000, EVE_ERR_SHAPE, "conv: spatial size too large");
  if (a.Nn > 32) return launch_gather<64, 8, 4>(a, s);
  if (a.Nn > 16) return launch_gather<32, 8, 4>(a, s);
  return launch_gather<16, 8, 2>(a, s);
}

// ------------------------------------------------------------------------------ wgrad --
struct WgradArgs {
  const float* x;    // [N,H,W,Cin]
  const float* dy;   // [N,OH,OW,lddy]
  float* part;       // [S][Cout][KK]  (KK = KH*KW*Cin)
  int H, W, Cin, OH, OW, Cout, lddy, KH, KW, stride, pad;
  int P;             // number of output pixels N*OH*OW
  int KK;
  int chunk;         // pixels per split
};

// 64 x 64 tile, 4x4 per thread.
__global__ void __launch_bounds__(NT) igemm_wgrad_kernel(WgradArgs a) {
  constexpr int BM = 64, BN = 64;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int p_begin = blockIdx.z * a.chunk;
  const int p_end = min(a.P, p_begin + a.chunk);

  const int lane64 = t & 63;
  const int kr0 = t >> 6;  // 0..3, rows kr0 + 4*j
  // B column owned by this thread: (tap, ci)
  const int nB = n0 + lane64;
  const bool nvalid = nB < a.KK;
  int tap = nvalid ? nB / a.Cin : 0;
  const int ci = nB - tap * a.Cin;
  const int r = tap / a.KW;
  const int q = tap - r * a.KW;
  const int mA = m0 + lane64;
  const bool mvalid = mA < a.Cout;

  float ra[4], rb[4];
  auto load_tile = [&](int p0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int p = p0 + kr0 + 4 * j;
      float va = 0.f, vb = 0.f;
      if (p < p_end) {
        if (mvalid) va = __ldg(a.dy + (size_t)p * a.lddy + mA);
        if (nvalid) {
          int ox = p % a.OW;
          int tmp = p / a.OW;
          int oy = tmp % a.OH;
          int n = tmp / a.OH;
          int ys = oy * a.stride + r - a.pad;
          int xs = ox * a.stride + q - a.pad;
          if (ys >= 0 && ys < a.H && xs >= 0 && xs < a.W)
            vb = __ldg(a.x + ((size_t)(n * a.H + ys) * a.W + xs) * a.Cin + ci);
        }
      }
      ra[j] = va;
      rb[j] = vb;
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tx = t & 15, ty = t >> 4;

  if (p_begin < p_end) load_tile(p_begin);
  for (int p0 = p_begin; p0 < p_end; p0 += BK) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[kr0 + 4 * j][lane64] = ra[j];
      Bs[kr0 + 4 * j][lane64] = rb[j];
    }
    __syncthreads();
    if (p0 + BK < p_end) load_tile(p0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float aa[4] = {av.x, av.y, av.z, av.w};
      float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  float* out = a.part + (size_t)blockIdx.z * a.Cout * a.KK;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= a.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < a.KK) out[(size_t)m * a.KK + n] = acc[i][j];
    }
  }
}

// dw_oihw[co][ci][r][q] (+)= sum_z part[z][co][(r*KW+q)*Cin+ci]   (deterministic order)
// block = 64 consecutive source elements (coalesced reads of every split) x 4 split lanes;
// lane l sums splits l, l+4, ... with four loads in flight, lanes are combined in fixed order.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, int S, int Cout, int Cin, int KHW,
                    float* __restrict__ dw, int accumulate) {
  __shared__ float sm[256];
  const size_t KK = (size_t)KHW * Cin;
  const size_t total = (size_t)Cout * KK;
  const int tx = threadIdx.x & 63, tz = threadIdx.x >> 6;
  const size_t src = (size_t)blockIdx.x * 64 + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (src < total) {
    const float* p = part + src;
    int z = tz;
    for (; z + 12 < S; z += 16) {
      a0 += p[(size_t)z * total];
      a1 += p[(size_t)(z + 4) * total];
      a2 += p[(size_t)(z + 8) * total];
      a3 += p[(size_t)(z + 12) * total];
    }
    for (; z < S; z += 4) a0 += p[(size_t)z * total];
  }
  sm[threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (tz == 0 && src < total) {
    const float s = (sm[tx] + sm[64 + tx]) + (sm[128 + tx] + sm[192 + tx]);
    const int co = (int)(src / KK);
    const int rem = (int)(src - (size_t)co * KK);
    const int tap = rem / Cin, ci = rem - tap * Cin;
    const size_t i = ((size_t)co * Cin + ci) * KHW + tap;
    dw[i] = accumulate ? dw[i] + s : s;
  }
}

// Derived weight layouts.
__global__ void prep_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int KHW,
                                    float* __restrict__ wf, float* __restrict__ wd) {
  size_t total = (size_t)Cout * Cin * KHW;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int tap = (int)(i % KHW);
  size_t tmp = i / KHW;
  int ci = (int)(tmp % Cin);
  int co = (int)(tmp / Cin);
  float v = w[i];
  if (wf) wf[((size_t)tap * Cin + ci) * Cout + co] = v;
  if (wd) wd[((size_t)tap * Cout + co) * Cin + ci] = v;
}

__global__ void colsum_partial_kernel(const float* __restrict__ dy, long long rows, int C, int ld,
                                      int rows_per_block, float* __restrict__ part) {
  // block: 256 threads = CB channels x (256/CB) row lanes
  extern __shared__ float sm[];
  int CB = min(C, 32);
  int lanes = blockDim.x / CB;
  int c = blockIdx.y * CB + threadIdx.x % CB;
  int lane = threadIdx.x / CB;
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < C)
    for (long long r = r0 + lane; r < r1; r += lanes) s += dy[r * ld + c];
  sm[threadIdx.x] = s;
  __syncthreads();
  if (lane == 0 && c < C) {
    for (int l = 1; l < lanes; ++l) s += sm[l * CB + threadIdx.x % CB];
    part[(size_t)blockIdx.x * C + c] = s;
  }
}

// one warp per channel: lanes stride over the partial blocks in a fixed order, then a shuffle
// tree (deterministic); the serial one-thread-per-channel loop cost 35-60 us per bias gradient
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ db,
                    int accumulate) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float s0 = 0.f, s1 = 0.f;
  int b = lane;
  for (; b + 32 < nblk; b += 64) {
    s0 += part[(size_t)b * C + c];
    s1 += part[(size_t)(b + 32) * C + c];
  }
  if (b < nblk) s0 += part[(size_t)b * C + c];
  float s = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) db[c] = accumulate ? db[c] + s : s;
}

int colsum_blocks(long long rows) {
  long long b = (rows + 255) / 256;
  if (b > 1024) b = 1024;
  if (b < 1) b = 1;
  return (int)b;
}

// algorithmic work of one convolution pass (fwd, dgrad and wgrad all do 2*M*N*K flops)
inline double conv_flops(const ConvGeom& g) {
  return 2.0 * (double)g.N * g.OH * g.OW * (double)g.Cout * (double)g.K();
}
// algorithmic fp32 bytes: input + output activations + weights, each touched once
inline double conv_bytes(const ConvGeom& g) {
  return 4.0 * ((double)g.in_elems() + (double)g.out_elems() + (double)g.Cout * g.K());
}

int wgrad_splits(const ConvGeom& g) {
  int tiles = cdiv(g.Cout, 64) * cdiv(g.K(), 64);
  long long P = (long long)g.N * g.OH * g.OW;
  int S = cdiv(4 * kNumSMs, tiles);
  long long maxS = (P + 255) / 256;
  if (S > maxS) S = (int)maxS;
  if (S < 1) S = 1;
  return S;
}

}  // namespace

int conv_prep_weights(const ConvGeom& g, const float* w, float* wf, float* wd, cudaStream_t s) {
  size_t total = (size_t)g.Cout * g.Cin * g.KH * g.KW;
  prep_weights_kernel<<<cdiv(total, 256), 256, 0, s>>>(w, g.Cout, g.Cin, g.KH * g.KW, wf, wd);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int conv_fwd_simt(const ConvGeom& g, const float* x, const float* wf, const float* bias,
                  const float* addend, float* y, int ldo, cudaStream_t s) {
  GatherArgs a;
  a.src = x; a.wmat = wf; a.bias = bias; a.addend = addend; a.out = y;
  a.Hs = g.H; a.Ws = g.W; a.Cs = g.Cin; a.lds = g.Cin;
  a.Ho = g.OH; a.Wo = g.OW; a.Nn = g.Cout; a.KH = g.KH; a.KW = g.KW;
  a.num = g.stride; a.den = 1; a.dr = 1; a.base = -g.pad;
  a.M = g.N * g.OH * g.OW; a.K = g.K(); a.ldo = ldo;
  ProfScope prof(PROF_CONV_FWD, conv_flops(g), conv_bytes(g), s, &g);
  return dispatch_gather(a, s);
}

int conv_dgrad_simt(const ConvGeom& g, const float* dy, int lddy, const float* wd,
                    const float* addend, float* dx, cudaStream_t s) {
  GatherArgs a;
  a.src = dy; a.wmat = wd; a.bias = nullptr; a.addend = addend; a.out = dx;
  a.Hs = g.OH; a.Ws = g.OW; a.Cs = g.Cout; a.lds = lddy;
  a.Ho = g.H; a.Wo = g.W; a.Nn = g.Cin; a.KH = g.KH; a.KW = g.KW;
  a.num = 1; a.den = g.stride; a.dr = -1; a.base = g.pad;
  a.M = g.N * g.H * g.W; a.K = g.KH * g.KW * g.Cout; a.ldo = g.Cin;
  ProfScope prof(PROF_CONV_DGRAD, conv_flops(g), conv_bytes(g), s, &g);
  return dispatch_gather(a, s);
}

size_t conv_wgrad_scratch_floats(const ConvGeom& g) {
  return (size_t)wgrad_splits(g) * g.Cout * g.K();
}

int conv_wgrad_simt(const ConvGeom& g, const float* x, const float* dy, int lddy, float* dw,
                    float* scratch, bool accumulate, cudaStream_t s) {
  WgradArgs a;
  a.x = x; a.dy = dy; a.part = scratch;
  a.H = g.H; a.W = g.W; a.Cin = g.Cin; a.OH = g.OH; a.OW = g.OW; a.Cout = g.Cout;
  a.lddy = lddy; a.KH = g.KH; a.KW = g.KW; a.stride = g.stride; a.pad = g.pad;
  a.P = g.N * g.OH * g.OW; a.KK = g.K();
  int S = wgrad_splits(g);
  a.chunk = cdiv(cdiv(a.P, S), BK) * BK;
  S = cdiv(a.P, a.chunk);
  dim3 grid(cdiv(g.Cout, 64), cdiv(a.KK, 64), S);
  ProfScope prof(PROF_CONV_WGRAD, conv_flops(g), conv_bytes(g), s, &g);
  igemm_wgrad_kernel<<<grid, NT, 0, s>>>(a);
  EVE_LAUNCH_CHECK();
  size_t total = (size_t)g.Cout * g.K();
  wgrad_reduce_kernel<<<cdiv(total, 64), 256, 0, s>>>(scratch, S, g.Cout, g.Cin, g.KH * g.KW, dw,
                                                      accumulate ? 1 : 0);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int wgrad_reduce(const float* part, int splits, const ConvGeom& g, float* dw, bool accumulate,
                 cudaStream_t s) {
  size_t total = (size_t)g.Cout * g.K();
  wgrad_reduce_kernel<<<cdiv(total, 64), 256, 0, s>>>(part, splits, g.Cout, g.Cin, g.KH * g.KW, dw,
                                                      accumulate ? 1 : 0);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

size_t colsum_scratch_floats(long long rows, int C) { return (size_t)colsum_blocks(rows) * C; }

int colsum(const float* dy, long long rows, int C, int ld, float* db, float* scratch,
           bool accumulate, cudaStream_t s) {
  int nblk = colsum_blocks(rows);
  int rpb = (int)((rows + nblk - 1) / nblk);
  nblk = (int)((rows + rpb - 1) / rpb);
  int CB = C < 32 ? C : 32;
  int threads = (256 / CB) * CB;
  dim3 grid(nblk, cdiv(C, CB));
  colsum_partial_kernel<<<grid, threads, threads * sizeof(float), s>>>(dy, rows, C, ld, rpb,
                                                                        scratch);
  EVE_LAUNCH_CHECK();
  colsum_final_kernel<<<cdiv(C, 8), 256, 0, s>>>(scratch, nblk, C, db, accumulate ? 1 : 0);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

}  // namespace eve
