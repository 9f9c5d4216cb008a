// Input pipeline, device side: the frame preprocessing of
// datasources/eve_sequences.py:196-211 on the GPU, so that decoded frames cross PCIe as uint8
// (4x fewer bytes than the float32 tensors the reference's DataLoader ships, training.py:256-263).
//   preprocess_frames        : N x H x W x C uint8 -> N x C x H x W float32,  x * (2/255) - 1
//   preprocess_screen_frames : same layout change,                            x * (1/255)
// and the eye-patch split of :283-285 (the "eyes" video holds both patches side by side: the left
// eye is columns [ew, 2 ew), the right eye [0, ew)) as a column window.  Frames past the end of a
// clip are zero (the reference pads AFTER preprocessing, :287-299).  The products are rounded
// exactly like numpy's in-place float32 ops (multiply, then subtract: no fused multiply-add), so
// the result is bit-identical to the reference's arrays.
#include "common.cuh"

namespace eve {
namespace {

// one thread per 4 output pixels of one (frame, channel, row); coalesced float4 stores, the
// 4 x C uint8 loads of a thread are contiguous
__global__ void __launch_bounds__(256)
preprocess_kernel(const unsigned char* __restrict__ in, int N, int H, int Win, int C, int x0,
                  int Wout, float scale, float bias, const int* __restrict__ frames_per_clip,
                  int steps, float* __restrict__ out) {
  const long long quads = (long long)N * C * H * (Wout / 4);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= quads) return;
  const int wq = (int)(i % (Wout / 4));
  long long t = i / (Wout / 4);
  const int h = (int)(t % H);
  t /= H;
  const int c = (int)(t % C);
  const int n = (int)(t / C);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  const bool live = frames_per_clip == nullptr || (n % steps) < frames_per_clip[n / steps];
  if (live) {
    const unsigned char* src = in + (((size_t)n * H + h) * Win + x0 + wq * 4) * C + c;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = (float)src[(size_t)j * C];
      v[j] = __fadd_rn(__fmul_rn(x, scale), bias);     // numpy: x *= scale; x -= 1 (two roundings)
    }
  }
  *reinterpret_cast<float4*>(out + (((size_t)n * C + c) * H + h) * Wout + wq * 4) =
      make_float4(v[0], v[1], v[2], v[3]);
}

}  // namespace
}  // namespace eve

using namespace eve;

extern "C" int eve_preprocess_frames(const unsigned char* frames, int n, int h, int w_in, int c,
                                     int x_offset, int w_out, float scale, float bias,
                                     const int* frames_per_clip, int steps, float* out,
                                     eve_stream_t stream) {
  EVE_REQUIRE(n >= 0 && h > 0 && w_in > 0 && c > 0 && w_out > 0 && x_offset >= 0 &&
                  x_offset + w_out <= w_in && w_out % 4 == 0,
              EVE_ERR_SHAPE, "preprocess_frames: n=%d h=%d w_in=%d c=%d window [%d, %d)", n, h, w_in, c,
              x_offset, x_offset + w_out);
  EVE_REQUIRE(frames_per_clip == nullptr || (steps > 0 && n % steps == 0), EVE_ERR_SHAPE,
              "preprocess_frames: %d frames are not whole clips of %d steps", n, steps);
  if (n == 0) return EVE_OK;
  EVE_REQUIRE(frames && out, EVE_ERR_NULL, "preprocess_frames: NULL pointer");
  const long long quads = (long long)n * c * h * (w_out / 4);
  preprocess_kernel<<<cdiv(quads, 256), 256, 0, as_stream(stream)>>>(
      frames, n, h, w_in, c, x_offset, w_out, scale, bias, frames_per_clip, steps > 0 ? steps : 1, out);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
