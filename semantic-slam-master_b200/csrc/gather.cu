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

// tap geometry of one keypoint: integer corner, validity of its two columns / rows, the weights
struct TapGeom {
  int xi0, yi0;
  bool vx0, vx1, vy0, vy1;
  float w[4];      // nw, ne, sw, se
};

__device__ __forceinline__ TapGeom tap_geometry(float x, float y, int H, int W, int coords) {
  // divisions by 16 and by 2 are exact scalings: multiplying by 2^-4 / 2^-1 gives the same bits as
  // the reference's divisions (no underflow at these magnitudes) without the IEEE division sequence
  if (coords == 1) {                                    // pixel_to_patch: (p - 8) / 16  (:177)
    x = __fmul_rn(__fsub_rn(x, 8.0f), 0.0625f);
    y = __fmul_rn(__fsub_rn(y, 8.0f), 0.0625f);
  }
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float nx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, x), wm1), 1.0f);
  float ny = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, y), hm1), 1.0f);
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.0f), 0.5f), wm1);
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.0f), 0.5f), hm1);
  float x0 = floorf(ix), y0 = floorf(iy);
  float x1 = __fadd_rn(x0, 1.0f), y1 = __fadd_rn(y0, 1.0f);
  float wx0 = __fsub_rn(x1, ix), wx1 = __fsub_rn(ix, x0);
  float wy0 = __fsub_rn(y1, iy), wy1 = __fsub_rn(iy, y0);
  TapGeom g;
  g.w[0] = __fmul_rn(wx0, wy0); g.w[1] = __fmul_rn(wx1, wy0);
  g.w[2] = __fmul_rn(wx0, wy1); g.w[3] = __fmul_rn(wx1, wy1);
  // float -> int conversion saturates, so wild coordinates stay "outside"
  g.xi0 = (int)x0; g.yi0 = (int)y0;
  const int xi1 = (int)x1, yi1 = (int)y1;
  g.vx0 = (x0 >= 0.f) && (g.xi0 < W); g.vx1 = (x1 >= 0.f) && (xi1 < W);
  g.vy0 = (y0 >= 0.f) && (g.yi0 < H); g.vy1 = (y1 >= 0.f) && (yi1 < H);
  return g;
}

__device__ __forceinline__ Taps make_taps(float x, float y, int H, int W, int C, int coords) {
  const TapGeom g = tap_geometry(x, y, H, W, coords);
  Taps t;
#pragma unroll
  for (int i = 0; i < 4; ++i) t.w[i] = g.w[i];
  const int xi1 = g.xi0 + 1, yi1 = g.yi0 + 1;            // only used when the flags say they are in range
  t.off[0] = (g.vx0 && g.vy0) ? (g.yi0 * W + g.xi0) * C : -1;
  t.off[1] = (g.vx1 && g.vy0) ? (g.yi0 * W + xi1) * C : -1;
  t.off[2] = (g.vx0 && g.vy1) ? (yi1 * W + g.xi0) * C : -1;
  t.off[3] = (g.vx1 && g.vy1) ? (yi1 * W + xi1) * C : -1;
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

// Row-band variant.  The warp-per-keypoint kernel above fetches four C*4-byte tap rows per keypoint
// from L2 (6 KB for 1.5 KB of output): with keypoints in score order nothing is reused in L1 and the
// kernel saturates the L2 slices, not HBM.  Here a CTA owns a group of consecutive patch rows of one
// frame: it streams those rows of the NHWC map through a three-slot shared-memory ring with 1-D bulk
// async copies (a row is one contiguous w*C*4-byte piece; the next row is in flight while a band
// is processed), bins the frame's keypoints by the row of their upper taps, and each warp blends
// its keypoints from shared memory.  The map crosses L2 -> SM once (plus one shared row per group)
// instead of ~7 times; arithmetic and operation order are those of gather_kernel.
constexpr int BAND_THREADS = 512;

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void band_bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr_u32(bar);
  uint32_t ok = 0;
  while (!ok)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
}

// per keypoint, computed once per CTA: weights, left tap column (-2: no tap column inside) and band
struct __align__(4) BandGeo { float w[4]; short xi0; short band; };

template <bool PAIR>
__global__ void __launch_bounds__(BAND_THREADS, 1)
gather_band_kernel(const float* __restrict__ feat, const float* __restrict__ kpts, int B, int H, int W, int C,
                   int N, int coords, float* __restrict__ out, __half* __restrict__ out_hi,
                   __half* __restrict__ out_lo, int bands_per_group, int ngroups) {
  extern __shared__ __align__(128) unsigned char gsm[];
  const int row_bytes = W * C * 4, row_floats = W * C;
  const float* ring = reinterpret_cast<const float*>(gsm);                  // three rows of the map
  BandGeo* geo = reinterpret_cast<BandGeo*>(gsm + 3 * row_bytes);            // [N]
  unsigned short* list = reinterpret_cast<unsigned short*>(geo + N);        // [N] keypoints of the current band
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(list + N) + 7) & ~(uintptr_t)7);
  __shared__ int list_n[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / ngroups, g = blockIdx.x - b * ngroups;
  // bands are indexed by the row of the upper taps: -1 .. H-1 (row -1 and row H are all zeros);
  // keypoints with no tap inside the map at all go to the very first band
  const int band_lo = g * bands_per_group - 1;
  const int band_hi = min(band_lo + bands_per_group, H);                     // exclusive
  const float* img = feat + (size_t)b * H * W * C;
  const float* kp_b = kpts + (size_t)b * N * 2;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    list_n[0] = 0; list_n[1] = 0;
  }
  const int first_row = max(band_lo, 0);
  const int last_row = min(band_hi, H - 1);                // band y0 needs rows y0 and y0 + 1
  __syncthreads();
  auto issue_row = [&](int y) {                            // tid 0: row y -> slot y mod 3 (y in [0, H))
    const int slot = y % 3;
    const uint32_t bar = smem_addr_u32(&bars[slot]);
    const int use = (y - first_row) / 3;                   // rows are requested in increasing order
    if (use > 0) band_bar_wait(&bars[slot], (uint32_t)(use - 1) & 1u);       // previous fill has landed
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)row_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr_u32(gsm + (size_t)slot * row_bytes)),
                 "l"(img + (size_t)y * W * C), "r"((uint32_t)row_bytes), "r"(bar)
                 : "memory");
  };
  if (tid == 0) {                                          // the first two rows stream in during the pre-pass
    if (first_row <= last_row) issue_row(first_row);
    if (first_row + 1 <= last_row) issue_row(first_row + 1);
  }
  // pre-pass: tap geometry of every keypoint of the frame, once (the IEEE divisions of the
  // reference's coordinate arithmetic are the expensive part)
  for (int i = tid; i < N; i += BAND_THREADS) {
    const float2 xy = __ldg(reinterpret_cast<const float2*>(kp_b) + i);
    const TapGeom t = tap_geometry(xy.x, xy.y, H, W, coords);
    BandGeo gg;
#pragma unroll
    for (int k = 0; k < 4; ++k) gg.w[k] = t.w[k];
    const bool any_x = t.vx0 || t.vx1, any_y = t.vy0 || t.vy1;
    gg.xi0 = (short)((any_x && any_y) ? (t.vx0 ? t.xi0 : -1) : -2);
    gg.band = (short)((any_x && any_y && t.vy0) ? t.yi0 : -1);
    geo[i] = gg;
  }
  __syncthreads();

  int it = 0;
  for (int y0 = band_lo; y0 < band_hi; ++y0, ++it) {
    // the slot of row y0 + 2 held row y0 - 1, which the previous band (synchronised below) was the last to read
    if (tid == 0 && y0 + 2 <= last_row && y0 + 2 >= first_row + 2) issue_row(y0 + 2);
    const bool r0_ok = y0 >= 0 && y0 < H, r1_ok = y0 + 1 >= 0 && y0 + 1 < H;
    const int o0 = ((y0 + 3) % 3) * row_floats, o1 = ((y0 + 1) % 3) * row_floats;
    int* cnt = &list_n[it & 1];
    for (int i = tid; i < N; i += BAND_THREADS)
      if (geo[i].band == y0) list[atomicAdd(cnt, 1)] = (unsigned short)i;
    __syncthreads();
    const int n = *cnt;
    if (tid == 0) list_n[(it + 1) & 1] = 0;                 // next band's counter (not in use now)
    // every requested row is waited for (also by bands without keypoints: no copy may be in flight
    // when the CTA exits)
    if (r0_ok) band_bar_wait(&bars[y0 % 3], (uint32_t)((y0 - first_row) / 3) & 1u);
    if (r1_ok) band_bar_wait(&bars[(y0 + 1) % 3], (uint32_t)((y0 + 1 - first_row) / 3) & 1u);
    for (int e = warp; e < n; e += BAND_THREADS / 32) {
      const int i = list[e];
      const BandGeo gg = geo[i];
      const size_t kp = (size_t)b * N + i;
      Taps w4;
#pragma unroll
      for (int k = 0; k < 4; ++k) w4.w[k] = gg.w[k];
      const int xi0 = gg.xi0;
      const bool vx0 = xi0 >= 0 && xi0 < W, vx1 = xi0 >= -1 && xi0 + 1 < W;
      const int t0 = (vx0 && r0_ok) ? o0 + xi0 * C : -1, t1 = (vx1 && r0_ok) ? o0 + (xi0 + 1) * C : -1;
      const int t2 = (vx0 && r1_ok) ? o1 + xi0 * C : -1, t3 = (vx1 && r1_ok) ? o1 + (xi0 + 1) * C : -1;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = lane * 4; c < C; c += 128) {
        const float4 a0 = t0 >= 0 ? *reinterpret_cast<const float4*>(ring + t0 + c) : zero;
        const float4 a1 = t1 >= 0 ? *reinterpret_cast<const float4*>(ring + t1 + c) : zero;
        const float4 a2 = t2 >= 0 ? *reinterpret_cast<const float4*>(ring + t2 + c) : zero;
        const float4 a3 = t3 >= 0 ? *reinterpret_cast<const float4*>(ring + t3 + c) : zero;
        float4 r;
        r.x = blend(a0.x, a1.x, a2.x, a3.x, w4);
        r.y = blend(a0.y, a1.y, a2.y, a3.y, w4);
        r.z = blend(a0.z, a1.z, a2.z, a3.z, w4);
        r.w = blend(a0.w, a1.w, a2.w, a3.w, w4);
        if (!PAIR) {
          *reinterpret_cast<float4*>(out + kp * C + c) = r;
        } else {                                            // fp16 pair for the tensor-core refiner
          const __half2 h01 = __floats2half2_rn(r.x, r.y), h23 = __floats2half2_rn(r.z, r.w);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(__fmul_rn(__fsub_rn(r.x, f01.x), 2048.0f),
                                                __fmul_rn(__fsub_rn(r.y, f01.y), 2048.0f));
          const __half2 l23 = __floats2half2_rn(__fmul_rn(__fsub_rn(r.z, f23.x), 2048.0f),
                                                __fmul_rn(__fsub_rn(r.w, f23.y), 2048.0f));
          uint2 hv, lv;
          hv.x = *reinterpret_cast<const unsigned*>(&h01); hv.y = *reinterpret_cast<const unsigned*>(&h23);
          lv.x = *reinterpret_cast<const unsigned*>(&l01); lv.y = *reinterpret_cast<const unsigned*>(&l23);
          *reinterpret_cast<uint2*>(out_hi + kp * C + c) = hv;
          *reinterpret_cast<uint2*>(out_lo + kp * C + c) = lv;
        }
      }
    }
    __syncthreads();                                        // band done: its upper row's slot may be refilled
  }
}

// one warp per row; two passes over a row that stays in registers when D <= 1024
template <bool VEC4>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_kernel(const float* __restrict__ in, int rows, int D, float eps, float* __restrict__ out_f32,
              __nv_bfloat16* __restrict__ out_bf16, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
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
      if (out_hi) {                                         // fp16 pair for the f16x3 matcher
        __half h4[4], l4[4];
        const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h4[i] = __float2half_rn(rr[i]);
          l4[i] = __float2half_rn(__fmul_rn(__fsub_rn(rr[i], __half2float(h4[i])), 2048.0f));
        }
        *reinterpret_cast<uint2*>(out_hi + (size_t)row * D + c) = *reinterpret_cast<uint2*>(h4);
        *reinterpret_cast<uint2*>(out_lo + (size_t)row * D + c) = *reinterpret_cast<uint2*>(l4);
      }
    }
  } else {
    for (int c = lane; c < D; c += 32) {
      float r = __fdiv_rn(__ldg(src + c), denom);
      if (out_f32) out_f32[(size_t)row * D + c] = r;
      if (out_bf16) out_bf16[(size_t)row * D + c] = __float2bfloat16_rn(r);
      if (out_hi) {
        const __half h = __float2half_rn(r);
        out_hi[(size_t)row * D + c] = h;
        out_lo[(size_t)row * D + c] = __float2half_rn(__fmul_rn(__fsub_rn(r, __half2float(h)), 2048.0f));
      }
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
  // row-band kernel: three map rows, the band table and the band list must fit in shared memory
  const size_t band_smem = 3 * (size_t)w * C * 4 + (size_t)N * (sizeof(BandGeo) + 2) + 64;
  if (vec && ((size_t)w * C * 4) % 16 == 0 && N <= 65535 && band_smem <= 227 * 1024 - 256 &&
      (reinterpret_cast<uintptr_t>(kpts) & 7) == 0 && ((out != nullptr) != (out_hi != nullptr))) {
    static DeviceOnce once;
    if (once.first_use()) {
      SSLAM_CHECK_CUDA(cudaFuncSetAttribute(gather_band_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            227 * 1024 - 256));
      SSLAM_CHECK_CUDA(cudaFuncSetAttribute(gather_band_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            227 * 1024 - 256));
    }
    // bands -1 .. h-1 in groups of consecutive bands; about eight CTAs per SM over the launch so
    // that the tail is short, at least three bands per group so that the shared row stays cheap
    int ngroups = (8 * num_sms() + B - 1) / B;
    if (ngroups > (h + 1) / 3) ngroups = (h + 1) / 3;
    if (ngroups < 1) ngroups = 1;
    const int per = (h + 1 + ngroups - 1) / ngroups;
    ngroups = (h + 1 + per - 1) / per;
    SSLAM_LAUNCH(KK_GATHER, stream,
                 if (out_hi)
        gather_band_kernel<true><<<(unsigned)(B * ngroups), BAND_THREADS, band_smem, stream>>>(
            feat, kpts, B, h, w, C, N, coords, out, out_hi, out_lo, per, ngroups);
      else
        gather_band_kernel<false><<<(unsigned)(B * ngroups), BAND_THREADS, band_smem, stream>>>(
            feat, kpts, B, h, w, C, N, coords, out, out_hi, out_lo, per, ngroups));
    return SSLAM_OK;
  }
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
                                 void* out_bf16, void* out_hi_, void* out_lo_, void* stream_) {
  __half* out_hi = static_cast<__half*>(out_hi_);
  __half* out_lo = static_cast<__half*>(out_lo_);
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(rows >= 0 && D > 0, SSLAM_EINVAL, "l2norm: bad size");
  if (rows == 0) return SSLAM_OK;
  SSLAM_REQUIRE(in && (out_f32 || out_bf16 || out_hi), SSLAM_EINVAL, "l2norm: null pointer");
  SSLAM_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), SSLAM_EINVAL, "l2norm: pair outputs go together");
  const unsigned blocks = (unsigned)((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                   (!out_f32 || (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0) &&
                   (!out_bf16 || (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out_hi) & 7) == 0) && ((reinterpret_cast<uintptr_t>(out_lo) & 7) == 0);
  SSLAM_LAUNCH(KK_L2NORM, stream,
               if (vec)
      l2norm_kernel<true><<<blocks, WARPS_PER_BLOCK * 32, 0, stream>>>(
          in, rows, D, eps, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_hi, out_lo);
    else
      l2norm_kernel<false><<<blocks, WARPS_PER_BLOCK * 32, 0, stream>>>(
          in, rows, D, eps, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_hi, out_lo));
  return SSLAM_OK;
}
