// Optional cell-softmax heatmap decode (north_star: "softmax/depth-to-space ... border mask").
//
// The reference CODE has no such mode — its detector head is a 1-channel sigmoid at patch resolution
// (models/keypoint_selector.py:61-62, "SIGMOID (not softmax!)"); the 65-channel softmax +
// depth-to-space decode exists only in the author's reading notes
// (papers/pdfs/SuperPoint_DeTone.md:49-58).  This kernel provides it as an OFF-BY-DEFAULT front stage
// of the decode: it turns cell logits into a pixel-resolution heatmap that sslam_decode_topk_f32
// then consumes (NMS, threshold, top-k).  It is checked against a plain PyTorch restatement written
// in the test (there is no reference function to compare with).
//
//   logits [B, cell*cell + 1, Hc, Wc]  ->  softmax over channels, drop the last ("dustbin") channel,
//   depth-to-space: heat[b, hc*cell + i, wc*cell + j] = p[b, i*cell + j, hc, wc],
//   border mask: pixels closer than `border` to an image edge are set to 0.
#include "common.cuh"

namespace sslam {
namespace {

template <int CELL>
__global__ void __launch_bounds__(128) cell_softmax_kernel(const float* __restrict__ logits, int B, int Hc, int Wc,
                                                          int border, float* __restrict__ heat) {
  constexpr int CH = CELL * CELL + 1;
  const long long cells = (long long)B * Hc * Wc;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const int wc = (int)(c % Wc), hc = (int)((c / Wc) % Hc), b = (int)(c / ((long long)Wc * Hc));
  const size_t plane = (size_t)Hc * Wc;
  const float* src = logits + (size_t)b * CH * plane + (size_t)hc * Wc + wc;   // channel stride = plane
  float v[CH];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < CH; ++k) { v[k] = __ldg(src + (size_t)k * plane); mx = fmaxf(mx, v[k]); }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < CH; ++k) { v[k] = expf(v[k] - mx); sum += v[k]; }
  const float inv = 1.0f / sum;
  const int H = Hc * CELL, W = Wc * CELL;
  float* dst = heat + (size_t)b * H * W + (size_t)(hc * CELL) * W + wc * CELL;
#pragma unroll
  for (int i = 0; i < CELL; ++i) {
    const int y = hc * CELL + i;
    const bool row_in = y >= border && y < H - border;
#pragma unroll
    for (int j = 0; j < CELL; ++j) {
      const int x = wc * CELL + j;
      const bool in = row_in && x >= border && x < W - border;
      dst[(size_t)i * W + j] = in ? v[i * CELL + j] * inv : 0.f;
    }
  }
}

}  // namespace
}  // namespace sslam

using namespace sslam;

extern "C" int sslam_heatmap_from_cells_f32(const float* logits, int B, int Hc, int Wc, int cell, int border,
                                            float* heat, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(B >= 0 && Hc > 0 && Wc > 0 && border >= 0, SSLAM_EINVAL, "heatmap: bad size");
  SSLAM_REQUIRE(cell == 8 || cell == 4 || cell == 2, SSLAM_EUNSUPPORTED, "heatmap: cell size %d (2, 4 or 8)", cell);
  if (B == 0) return SSLAM_OK;
  SSLAM_REQUIRE(logits && heat, SSLAM_EINVAL, "heatmap: null pointer");
  const long long cells = (long long)B * Hc * Wc;
  const unsigned grid = (unsigned)((cells + 127) / 128);
  if (cell == 8) SSLAM_LAUNCH(KK_HEATMAP, stream, cell_softmax_kernel<8><<<grid, 128, 0, stream>>>(logits, B, Hc, Wc, border, heat));
  else if (cell == 4) SSLAM_LAUNCH(KK_HEATMAP, stream, cell_softmax_kernel<4><<<grid, 128, 0, stream>>>(logits, B, Hc, Wc, border, heat));
  else SSLAM_LAUNCH(KK_HEATMAP, stream, cell_softmax_kernel<2><<<grid, 128, 0, stream>>>(logits, B, Hc, Wc, border, heat));
  return SSLAM_OK;
}
