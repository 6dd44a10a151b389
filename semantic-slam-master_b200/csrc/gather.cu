// Bilinear descriptor sampling (NHWC, one warp per keypoint) and row-wise L2 normalisation.
//
// Replaces DinoBackbone.extract_at_keypoints / pixel_to_patch (models/dino_backbone.py:114-152,
// 167-178) and F.normalize at the end of DescriptorRefiner.forward (descriptor_refiner.py:86).
//
// The coordinate arithmetic is the reference's, operation for operation, with explicit
// round-to-nearest intrinsics so that nvcc cannot contract it:
//   nx = 2x/(w-1) - 1                         (dino_backbone.py:135)
//   ix = ((nx+1)/2)*(w-1)                     (ATen GridSampler.h:27-31, align_corners=True)
//   taps nw, ne, sw, se with weights (x1-ix)(y1-iy) ... ; out-of-range taps read 0
//   out = fma(se, w_se, fma(sw, w_sw, fma(ne, w_ne, nw*w_nw)))   (torch 2.11 CPU kernel order)
// Each tap of a keypoint is one contiguous C*4-byte row of the NHWC map: a warp reads it with
// 128-bit loads (C % 4 == 0) and writes the (B,N,C) output the same way.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace sslam {
namespace {

constexpr int WARPS_PER_BLOCK = 8;

struct Taps {
  int off[4];      // element offset of the tap row inside the image, or -1 when outside
  float w[4];
};

__device__ __forceinline__ Taps make_taps(float x, float y, int H, int W, int C, int coords) {
  if (coords == 1) {                                    // pixel_to_patch: (p - 8) / 16  (:177)
    x = __fdiv_rn(__fsub_rn(x, 8.0f), 16.0f);
    y = __fdiv_rn(__fsub_rn(y, 8.0f), 16.0f);
  }
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float nx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, x), wm1), 1.0f);
  float ny = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, y), hm1), 1.0f);
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(nx, 1.0f), 2.0f), wm1);
  float iy = __fmul_rn(__fdiv_rn(__fadd_rn(ny, 1.0f), 2.0f), hm1);
  float x0 = floorf(ix), y0 = floorf(iy);
  float x1 = __fadd_rn(x0, 1.0f), y1 = __fadd_rn(y0, 1.0f);
  float wx0 = __fsub_rn(x1, ix), wx1 = __fsub_rn(ix, x0);
  float wy0 = __fsub_rn(y1, iy), wy1 = __fsub_rn(iy, y0);
  Taps t;
  t.w[0] = __fmul_rn(wx0, wy0); t.w[1] = __fmul_rn(wx1, wy0);
  t.w[2] = __fmul_rn(wx0, wy1); t.w[3] = __fmul_rn(wx1, wy1);
  // float -> int conversion saturates, so wild coordinates stay "outside"
  int xi0 = (int)x0, yi0 = (int)y0, xi1 = (int)x1, yi1 = (int)y1;
  bool vx0 = (x0 >= 0.f) && (xi0 < W), vx1 = (x1 >= 0.f) && (xi1 < W);
  bool vy0 = (y0 >= 0.f) && (yi0 < H), vy1 = (y1 >= 0.f) && (yi1 < H);
  t.off[0] = (vx0 && vy0) ? (yi0 * W + xi0) * C : -1;
  t.off[1] = (vx1 && vy0) ? (yi0 * W + xi1) * C : -1;
  t.off[2] = (vx0 && vy1) ? (yi1 * W + xi0) * C : -1;
  t.off[3] = (vx1 && vy1) ? (yi1 * W + xi1) * C : -1;
  return t;
}

__device__ __forceinline__ float blend(float a, float b, float c, float d, const Taps& t) {
  float acc = __fmul_rn(a, t.w[0]);
  acc = __fmaf_rn(b, t.w[1], acc);
  acc = __fmaf_rn(c, t.w[2], acc);
  acc = __fmaf_rn(d, t.w[3], acc);
  return acc;
}

template <bool VEC4>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
gather_kernel(const float* __restrict__ feat, const float* __restrict__ kpts, int B, int H, int W,
              int C, int N, int coords, float* __restrict__ out, __half* __restrict__ out_hi,
              __half* __restrict__ out_lo) {
  const int lane = threadIdx.x & 31;
  const long long kp = (long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (kp >= (long long)B * N) return;
  const int b = (int)(kp / N);
  const float x = __ldg(kpts + 2 * kp), y = __ldg(kpts + 2 * kp + 1);
  const Taps t = make_taps(x, y, H, W, C, coords);
  const float* img = feat + (size_t)b * H * W * C;
  float* o = out + (size_t)kp * C;
  if (VEC4) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = lane * 4; c < C; c += 128) {
      float4 v0 = t.off[0] >= 0 ? __ldg(reinterpret_cast<const float4*>(img + t.off[0] + c)) : zero;
      float4 v1 = t.off[1] >= 0 ? __ldg(reinterpret_cast<const float4*>(img + t.off[1] + c)) : zero;
      float4 v2 = t.off[2] >= 0 ? __ldg(reinterpret_cast<const float4*>(img + t.off[2] + c)) : zero;
      float4 v3 = t.off[3] >= 0 ? __ldg(reinterpret_cast<const float4*>(img + t.off[3] + c)) : zero;
      float4 r;
      r.x = blend(v0.x, v1.x, v2.x, v3.x, t);
      r.y = blend(v0.y, v1.y, v2.y, v3.y, t);
      r.z = blend(v0.z, v1.z, v2.z, v3.z, t);
      r.w = blend(v0.w, v1.w, v2.w, v3.w, t);
      if (out) *reinterpret_cast<float4*>(o + c) = r;
      if (out_hi) {                                       // fp16 pair for the tensor-core refiner
        __half h4[4], l4[4];
        const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h4[i] = __float2half_rn(rr[i]);
          l4[i] = __float2half_rn(__fmul_rn(__fsub_rn(rr[i], __half2float(h4[i])), 2048.0f));
        }
        *reinterpret_cast<uint2*>(out_hi + (size_t)kp * C + c) = *reinterpret_cast<uint2*>(h4);
        *reinterpret_cast<uint2*>(out_lo + (size_t)kp * C + c) = *reinterpret_cast<uint2*>(l4);
      }
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      float v0 = t.off[0] >= 0 ? __ldg(img + t.off[0] + c) : 0.f;
      float v1 = t.off[1] >= 0 ? __ldg(img + t.off[1] + c) : 0.f;
      float v2 = t.off[2] >= 0 ? __ldg(img + t.off[2] + c) : 0.f;
      float v3 = t.off[3] >= 0 ? __ldg(img + t.off[3] + c) : 0.f;
      const float r = blend(v0, v1, v2, v3, t);
      if (out) o[c] = r;
      if (out_hi) {
        const __half h = __float2half_rn(r);
        out_hi[(size_t)kp * C + c] = h;
        out_lo[(size_t)kp * C + c] = __float2half_rn(__fmul_rn(__fsub_rn(r, __half2float(h)), 2048.0f));
      }
    }
  }
}

// one warp per row; two passes over a row that stays in registers when D <= 1024
template <bool VEC4>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_kernel(const float* __restrict__ in, int rows, int D, float eps, float* __restrict__ out_f32,
              __nv_bfloat16* __restrict__ out_bf16) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* src = in + (size_t)row * D;
  float ss = 0.f;
  if (VEC4) {
    for (int c = lane * 4; c < D; c += 128) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
      ss = __fmaf_rn(v.x, v.x, ss); ss = __fmaf_rn(v.y, v.y, ss);
      ss = __fmaf_rn(v.z, v.z, ss); ss = __fmaf_rn(v.w, v.w, ss);
    }
  } else {
    for (int c = lane; c < D; c += 32) { float v = __ldg(src + c); ss = __fmaf_rn(v, v, ss); }
  }
  ss = warp_reduce_sum(ss);
  const float denom = fmaxf(sqrtf(ss), eps);              // clamp_min(eps) (descriptor_refiner.py:86)
  if (VEC4) {
    for (int c = lane * 4; c < D; c += 128) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
      float4 r = make_float4(__fdiv_rn(v.x, denom), __fdiv_rn(v.y, denom), __fdiv_rn(v.z, denom),
                             __fdiv_rn(v.w, denom));
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)row * D + c) = r;
      if (out_bf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
        uint2 pk = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
        *reinterpret_cast<uint2*>(out_bf16 + (size_t)row * D + c) = pk;
      }
    }
  } else {
    for (int c = lane; c < D; c += 32) {
      float r = __fdiv_rn(__ldg(src + c), denom);
      if (out_f32) out_f32[(size_t)row * D + c] = r;
      if (out_bf16) out_bf16[(size_t)row * D + c] = __float2bfloat16_rn(r);
    }
  }
}

}  // namespace
}  // namespace sslam

using namespace sslam;

extern "C" int sslam_gather_bilinear_f32(const float* feat, const float* kpts, int B, int h, int w,
                                         int C, int N, int coords, float* out, void* out_hi_, void* out_lo_,
                                         void* stream_) {
  __half* out_hi = static_cast<__half*>(out_hi_);
  __half* out_lo = static_cast<__half*>(out_lo_);
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(B >= 0 && h > 0 && w > 0 && C > 0 && N >= 0, SSLAM_EINVAL, "gather: bad size");
  SSLAM_REQUIRE(coords == 0 || coords == 1, SSLAM_EINVAL, "gather: coords must be 0 or 1");
  if (B == 0 || N == 0) return SSLAM_OK;
  SSLAM_REQUIRE(feat && kpts && (out || (out_hi && out_lo)), SSLAM_EINVAL, "gather: null pointer");
  SSLAM_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), SSLAM_EINVAL, "gather: pair outputs go together");
  SSLAM_REQUIRE((long long)h * w * C < (1ll << 31), SSLAM_EUNSUPPORTED, "gather: map too large");
  const long long total = (long long)B * N;
  const unsigned blocks = (unsigned)((total + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out_hi) & 7) == 0) && ((reinterpret_cast<uintptr_t>(out_lo) & 7) == 0);
  SSLAM_LAUNCH(KK_GATHER, stream,
               if (vec)
      gather_kernel<true><<<blocks, WARPS_PER_BLOCK * 32, 0, stream>>>(feat, kpts, B, h, w, C, N,
                                                                      coords, out, out_hi, out_lo);
    else
      gather_kernel<false><<<blocks, WARPS_PER_BLOCK * 32, 0, stream>>>(feat, kpts, B, h, w, C, N,
                                                                     coords, out, out_hi, out_lo));
  return SSLAM_OK;
}

extern "C" int sslam_l2norm_rows(const float* in, int rows, int D, float eps, float* out_f32,
                                 void* out_bf16, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(rows >= 0 && D > 0, SSLAM_EINVAL, "l2norm: bad size");
  if (rows == 0) return SSLAM_OK;
  SSLAM_REQUIRE(in && (out_f32 || out_bf16), SSLAM_EINVAL, "l2norm: null pointer");
  const unsigned blocks = (unsigned)((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                   (!out_f32 || (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0) &&
                   (!out_bf16 || (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0);
  SSLAM_LAUNCH(KK_L2NORM, stream,
               if (vec)
      l2norm_kernel<true><<<blocks, WARPS_PER_BLOCK * 32, 0, stream>>>(
          in, rows, D, eps, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16));
    else
      l2norm_kernel<false><<<blocks, WARPS_PER_BLOCK * 32, 0, stream>>>(
          in, rows, D, eps, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16)));
  return SSLAM_OK;
}
