// upconv.cu — the auto-encoder's up-convolution decoder for sm_100a.
//
// Replaces UpConv.forward (models/layers.py:214-240: nn.Upsample(scale_factor=2) + Conv2d 3x3 pad 1
// + BatchNorm2d + ReLU) as used by DecoderConv (models/networks.py:394-431).
//
// Nearest x2 up-sampling followed by a 3x3 convolution never needs the up-sampled image: output
// pixel (2i+py, 2j+px) only sees the 2x2 low-resolution neighbourhood rows {i-1+py, i+py} x
// columns {j-1+px, j+px}, and the three kernel rows (columns) collapse onto those two — rows
// {0}|{1,2} for py = 0 and {0,1}|{2} for py = 1. The layer is therefore FOUR GEMMs (one per output
// parity) with K = 4*Cin instead of one with K = 9*Cin over 4x the pixels: 2.25x fewer flops, and
// the zero padding of the up-sampled image coincides with zero padding of the low-resolution one.
// The host folds BN and the tap sums into four [Cout, 4*Cin] matrices once per weight change
// (sonet_b200/layers.py); here:
//   upconv_im2col_kernel    gathers the 2x2 neighbourhoods of all four parities,
//                           xcol[g][b][t*Cin+ci][i*W+j], coalesced along j;
//   sonet_pointwise_tc_grouped_forward (csrc/pointwise_tc.cu) runs the four GEMMs as ONE grouped
//                           tcgen05 launch (fp16 hi/lo split, fp32 accumulate) whose epilogue —
//                           or, for the small maps that need a K split to fill the SMs, the
//                           split-K reduce kernel — adds the folded shift, applies ReLU and
//                           interleaves the parities into the [B, Cout, 2H, 2W] output.
#include <algorithm>

#include "common.cuh"

namespace sonet {

__global__ void __launch_bounds__(256)
    upconv_im2col_kernel(const float* __restrict__ in, int B, int Cin, int H, int W,
                         float* __restrict__ xcol) {
  const int HW = H * W;
  const long long per_g = static_cast<long long>(B) * 4 * Cin * HW;
  const long long total = per_g * 4;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(e / per_g);
    long long r = e - g * per_g;
    const int p = static_cast<int>(r % HW);
    r /= HW;
    const int k = static_cast<int>(r % (4 * Cin));
    const int b = static_cast<int>(r / (4 * Cin));
    const int t = k / Cin, ci = k - t * Cin;
    const int i = p / W, j = p - i * W;
    const int y = i + (t >> 1) - 1 + (g >> 1), x = j + (t & 1) - 1 + (g & 1);
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W)
      v = __ldg(in + (static_cast<size_t>(b) * Cin + ci) * HW + y * W + x);
    xcol[e] = v;
  }
}

}  // namespace sonet

extern "C" int sonet_upconv_im2col_f32(const float* in, int B, int Cin, int H, int W, float* xcol,
                                       sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(B >= 0 && Cin >= 1 && H >= 1 && W >= 1, "upconv_im2col: bad dimension");
  if (B == 0) return SONET_OK;
  SONET_REQUIRE(in && xcol, "upconv_im2col: null pointer");
  const long long total = 16LL * B * Cin * H * W;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 16LL * sm_count()));
  upconv_im2col_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, B, Cin, H, W, xcol);
  return check_launch("upconv_im2col");
}
