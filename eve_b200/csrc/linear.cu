// nn.Linear (as 1x1 convolutions over 1x1 images) and small elementwise kernels.
//
// Reference call sites: torchvision ResNet.fc, eye_net.py:51-56 (fc_common), :74-78
// (static_fc), :81-92 (fc_to_gaze, fc_to_pupil; nn.SELU, nn.Tanh, nn.ReLU), :139 (pi/2 scale).
#include "common.cuh"

namespace eve {

namespace {
constexpr float kSeluAlpha = 1.6732632423543772f;
constexpr float kSeluScale = 1.0507009873554805f;
constexpr float kHalfPi = 1.5707963267948966f;

__global__ void __launch_bounds__(256)
ew_fwd_kernel(int op, const float* __restrict__ x, long long n, float* __restrict__ y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i], o;
  switch (op) {
    case EW_SELU: o = kSeluScale * (v > 0.f ? v : kSeluAlpha * expm1f(v)); break;
    case EW_RELU: o = v > 0.f ? v : 0.f; break;
    case EW_TANH_HALFPI: o = kHalfPi * tanhf(v); break;
    case EW_SIGMOID: o = 1.f / (1.f + expf(-v)); break;
    case EW_TANH: o = tanhf(v); break;
    default: o = v > 0.f ? v : 0.01f * v; break;
  }
  y[i] = o;
}

__global__ void __launch_bounds__(256)
ew_bwd_kernel(int op, const float* __restrict__ dy, const float* __restrict__ ref, long long n,
              float* __restrict__ dx) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d = dy[i], r = ref[i], o;
  switch (op) {
    case EW_SELU: o = d * kSeluScale * (r > 0.f ? 1.f : kSeluAlpha * expf(r)); break;
    case EW_RELU: o = r > 0.f ? d : 0.f; break;
    case EW_TANH_HALFPI: { float t = r * (1.f / kHalfPi); o = d * kHalfPi * (1.f - t * t); break; }
    case EW_SIGMOID: o = d * r * (1.f - r); break;
    case EW_TANH: o = d * (1.f - r * r); break;
    default: o = r > 0.f ? d : 0.01f * d; break;
  }
  dx[i] = o;
}

__global__ void __launch_bounds__(256)
ew_add_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
              float* __restrict__ y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}
}  // namespace

int ew_fwd(int op, const float* x, long long n, float* y, cudaStream_t s) {
  if (n == 0) return EVE_OK;
  ew_fwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(op, x, n, y);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
int ew_bwd(int op, const float* dy, const float* ref, long long n, float* dx, cudaStream_t s) {
  if (n == 0) return EVE_OK;
  ew_bwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(op, dy, ref, n, dx);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
int ew_add(const float* a, const float* b, long long n, float* y, cudaStream_t s) {
  if (n == 0) return EVE_OK;
  ew_add_kernel<<<cdiv(n, 256), 256, 0, s>>>(a, b, n, y);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
int fill_zero(float* p, long long n, cudaStream_t s) {
  if (n == 0) return EVE_OK;
  EVE_CUDA(cudaMemsetAsync(p, 0, (size_t)n * sizeof(float), s));
  return EVE_OK;
}

int linear_fwd(const float* x, int M, int K, const float* W, const float* b, int N, float* y,
               float* wt, cudaStream_t s) {
  ConvGeom g = make_conv(M, 1, 1, K, N, 1, 1, 0);
  EVE_TRY(conv_prep_weights(g, W, wt, nullptr, s));
  return conv_fwd_simt(g, x, wt, b, nullptr, y, N, s);
}

int linear_dgrad(const float* dy, int M, int N, const float* W, int K, const float* addend,
                 float* dx, cudaStream_t s) {
  // dgrad layout wd[(tap*Cout+co)][ci] of a 1x1 conv is W[N][K] itself
  ConvGeom g = make_conv(M, 1, 1, K, N, 1, 1, 0);
  return conv_dgrad_simt(g, dy, N, W, addend, dx, s);
}

size_t linear_wgrad_scratch_floats(int M, int K, int N) {
  ConvGeom g = make_conv(M, 1, 1, K, N, 1, 1, 0);
  size_t a = conv_wgrad_scratch_floats(g), b = colsum_scratch_floats(M, N);
  return a > b ? a : b;
}

int linear_wgrad(const float* x, const float* dy, int M, int K, int N, float* dW, float* db,
                 float* scratch, bool accumulate, cudaStream_t s) {
  ConvGeom g = make_conv(M, 1, 1, K, N, 1, 1, 0);
  if (dW) EVE_TRY(conv_wgrad_simt(g, x, dy, N, dW, scratch, accumulate, s));
  if (db) EVE_TRY(colsum(dy, M, N, N, db, scratch, accumulate, s));
  return EVE_OK;
}

}  // namespace eve
