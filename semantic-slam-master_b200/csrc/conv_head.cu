// KeypointSelector head on tensor cores.
//
// Replaces KeypointSelector.conv + sigmoid (models/keypoint_selector.py:30-34, 56-65):
//   Conv2d(C, hidden, 3, padding=1) -> ReLU -> Conv2d(hidden, 1, 1) -> sigmoid
// on the NHWC patch-feature map the backbone produces, as ONE kernel:
//   * the 3x3 convolution is an implicit GEMM, M = pixels, N = hidden, K = 9*C.  A 128-row A tile is a
//     (box_h x box_w) window of pixels; for tap (ky, kx) and channel block cb the producer issues ONE
//     4-D TMA load of the NHWC tensor at (cb*64, x0+kx-1, y0+ky-1, b): the box lands in shared memory
//     as 128 rows of 128 bytes — exactly the K-major SWIZZLE_128B operand tile — and the parts of the
//     window that fall outside the image are zero-filled by the TMA unit (= the convolution's zero
//     padding).  No im2col buffer exists anywhere.
//   * fp32-level accuracy at the 16-bit tensor rate: features and weights are fp16 (hi, lo) pairs,
//     x = hi + lo * 2^-11, three kind::f16 MMAs per product with the cross terms in their own TMEM
//     accumulator (same scheme as refiner_tc.cu); the features are split once per call, the weights
//     once per weight set (sslam_selector_pack_weights, which also reorders them to [hidden][tap][C]).
//   * the epilogue (thread = pixel, tcgen05.ld) adds the bias, applies ReLU and contracts the hidden
//     activations with the 1x1 convolution's weights on the fly: the (pixels x hidden) activation
//     never reaches HBM; the kernel writes one logit (or sigmoid) per pixel — the saliency map the
//     decode kernels consume.
// Persistent: one CTA per SM walks the pixel tiles; roles as in gemm_f16x3_kernel (TMA producer warp,
// MMA warp, eight epilogue warps, 3-stage 64 KB operand ring, double-buffered accumulators).
#include "tc_common.cuh"

namespace sslam {

using namespace tc;

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int BLOCK_BYTES = BM * 128;
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 4 * BLOCK_BYTES;             // A_hi, A_lo, B_hi, B_lo
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int UNIT = 16;
constexpr int MAX_HIDDEN = 512;
constexpr int SMEM_OPERANDS = STAGES * STAGE_BYTES;
constexpr int SMEM_VECS = 2 * MAX_HIDDEN * 4;            // bias of the 3x3 conv, weights of the 1x1 conv
constexpr int SMEM_PART = 2 * 2 * BM * 4;                // [parity][half][row] partial dot products
constexpr int SMEM_BARS = (2 * STAGES + 4) * 8 + 16;
constexpr int SMEM_TOTAL = SMEM_OPERANDS + SMEM_VECS + SMEM_PART + SMEM_BARS + 1024;

struct ConvParams {
  int B, H, W, C, hidden;
  int bw, bh, tiles_x, tiles_y;
  const float* b1;     // [hidden]
  const float* w2;     // [hidden]
  const float* b2;     // [1]
  int apply_sigmoid;
  float* out;          // [B,H,W]
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_head_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 ConvParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* operands = smem;
  float* sb1 = reinterpret_cast<float*>(smem + SMEM_OPERANDS);
  float* sw2 = sb1 + MAX_HIDDEN;
  float* spart = reinterpret_cast<float*>(smem + SMEM_OPERANDS + SMEM_VECS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_OPERANDS + SMEM_VECS + SMEM_PART);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_img = p.tiles_x * p.tiles_y;
  const int ntiles = p.B * tiles_img;
  const int ntile_n = (p.hidden + BN - 1) / BN;
  const int kpc = p.C / BK;                              // channel blocks per tap
  const int nkb = 9 * kpc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  for (int i = threadIdx.x; i < MAX_HIDDEN; i += NUM_THREADS) {
    sb1[i] = i < p.hidden ? __ldg(p.b1 + i) : 0.f;
    sw2[i] = i < p.hidden ? __ldg(p.w2 + i) : 0.f;       // zero weight: padded columns contribute nothing
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {                                            // ---- TMA producer
      prefetch_tensormap(&tmA_hi); prefetch_tensormap(&tmA_lo);
      prefetch_tensormap(&tmB_hi); prefetch_tensormap(&tmB_lo);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / tiles_img, t = tile - b * tiles_img;
        const int y0 = (t / p.tiles_x) * p.bh, x0 = (t % p.tiles_x) * p.bw;
        for (int ct = 0; ct < ntile_n; ++ct) {
          for (int kb = 0; kb < nkb; ++kb) {
            const int tap = kb / kpc, cb = kb - tap * kpc;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* st = operands + stage * STAGE_BYTES;
            mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
            tma_load_4d(st, &tmA_hi, &full[stage], cb * BK, x0 + dx, y0 + dy, b);
            tma_load_4d(st + BLOCK_BYTES, &tmA_lo, &full[stage], cb * BK, x0 + dx, y0 + dy, b);
            tma_load_2d(st + 2 * BLOCK_BYTES, &tmB_hi, &full[stage], kb * BK, ct * BN);
            tma_load_2d(st + 3 * BLOCK_BYTES, &tmB_lo, &full[stage], kb * BK, ct * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                            // ---- MMA issuer
      const uint32_t idesc = make_instr_desc(FMT_F16, BM, BN);
      const uint32_t idesc_cat = make_instr_desc(FMT_F16, BM, 2 * BN);
      int stage = 0; uint32_t phase = 0;
      const int my_tiles = ((ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * ntile_n;
      for (int tc = 0; tc < my_tiles; ++tc) {
        const int acc = tc & 1;
        mbar_wait(&tempty[acc], ((tc >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 2 * BN;
        const uint32_t tmem_s = tmem_d + BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(operands + stage * STAGE_BYTES);
          const uint64_t a_hi = make_smem_desc_sw128(sa);
          const uint64_t a_lo = make_smem_desc_sw128(sa + BLOCK_BYTES);
          const uint64_t b_hi = make_smem_desc_sw128(sa + 2 * BLOCK_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            const uint32_t first = (kb | k) ? 1u : 0u;
            // A_hi x [B_hi ; B_lo] as one N=256 instruction (hi.hi -> columns [0,128), hi.lo -> [128,256)),
            // then lo.hi accumulates into the latter
            umma_ss<false>(tmem_d, a_hi + adv, b_hi + adv, idesc_cat, first);
            umma_ss<false>(tmem_s, a_lo + adv, b_hi + adv, idesc, 1u);
          }
          tcgen05_commit(&empty[stage]);
          if (kb == nkb - 1) tcgen05_commit(&tfull[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ---- epilogue warps 2..9: thread = pixel row of the tile; warps w and w+4 share a TMEM lane
    // quarter and take 64 hidden columns each per column tile
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int yy = row / p.bw, xx = row - yy * p.bw;
    const float bias2 = __ldg(p.b2);
    int tc = 0, it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_img, t = tile - b * tiles_img;
      const int y = (t / p.tiles_x) * p.bh + yy, x = (t % p.tiles_x) * p.bw + xx;
      float part = 0.f;
      for (int ct = 0; ct < ntile_n; ++ct, ++tc) {
        const int acc = tc & 1;
        mbar_wait(&tfull[acc], (tc >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll 1
        for (int un = 0; un < 64 / UNIT; ++un) {
          const int col0 = half * 64 + un * UNIT;
          const int gc0 = ct * BN + col0;                 // hidden index of the unit's first column
          if (gc0 >= p.hidden) continue;                  // warp-uniform
          uint32_t r[UNIT], rs[UNIT];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN + col0;
          tmem_ld_32x16(taddr, r);
          tmem_ld_32x16(taddr + BN, rs);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < UNIT; ++j) {
            float v = __fmaf_rn(__uint_as_float(rs[j]), F16_LO_INV, __uint_as_float(r[j]));   // fold the cross terms
            v = fmaxf(__fadd_rn(v, sb1[gc0 + j]), 0.f);                                       // bias + ReLU
            part = __fmaf_rn(v, sw2[gc0 + j], part);                                          // 1x1 convolution
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
      // combine the two column halves of each pixel
      float* sp = spart + (it & 1) * (2 * BM);
      sp[half * BM + row] = part;
      named_bar_sync(1, 32 * EPI_WARPS);
      if (half == 0 && y < p.H && x < p.W) {
        float logit = __fadd_rn(__fadd_rn(sp[row], sp[BM + row]), bias2);
        if (p.apply_sigmoid) logit = sigmoid_f32(logit);
        p.out[((size_t)b * p.H + y) * p.W + x] = logit;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// fp32 [n] -> fp16 pair
__global__ void conv_split_kernel(const float4* __restrict__ src, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                  size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 x = __ldg(src + i);
    __half h[4], l[4];
    split_f16(x.x, h[0], l[0]); split_f16(x.y, h[1], l[1]);
    split_f16(x.z, h[2], l[2]); split_f16(x.w, h[3], l[3]);
    hi[i] = *reinterpret_cast<uint2*>(h);
    lo[i] = *reinterpret_cast<uint2*>(l);
  }
}

// torch Conv2d weight [hidden, C, 3, 3] -> fp16 pair [hidden][tap = ky*3+kx][C]
__global__ void conv_pack_kernel(const float* __restrict__ w1, int hidden, int C, __half* __restrict__ hi,
                                 __half* __restrict__ lo) {
  const size_t total = (size_t)hidden * 9 * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int tap = (int)((i / C) % 9);
    const int n = (int)(i / ((size_t)9 * C));
    const float w = w1[((size_t)n * C + c) * 9 + tap];
    __half h, l;
    split_f16(w, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

size_t half_bytes(size_t n) { return align_up(n * 2, 256); }

}  // namespace
}  // namespace sslam

using namespace sslam;

extern "C" size_t sslam_selector_packed_bytes(int C, int hidden) {
  if (C <= 0 || hidden <= 0) return 0;
  return 2 * half_bytes((size_t)hidden * 9 * C);
}

extern "C" int sslam_selector_pack_weights(const float* conv1_weight, int C, int hidden, void* packed,
                                           size_t packed_bytes, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(conv1_weight && packed, SSLAM_EINVAL, "selector_pack: null pointer");
  SSLAM_REQUIRE(C > 0 && C % BK == 0 && hidden > 0 && hidden % 8 == 0 && hidden <= MAX_HIDDEN, SSLAM_EUNSUPPORTED,
                "selector: C must be a multiple of 64, hidden a multiple of 8 and <= 512 (C=%d hidden=%d)", C, hidden);
  SSLAM_REQUIRE(packed_bytes >= sslam_selector_packed_bytes(C, hidden), SSLAM_EWORKSPACE,
                "selector_pack: packed buffer too small");
  const size_t n = (size_t)hidden * 9 * C;
  __half* hi = static_cast<__half*>(packed);
  __half* lo = reinterpret_cast<__half*>(static_cast<char*>(packed) + half_bytes(n));
  SSLAM_LAUNCH(KK_SPLIT, stream, conv_pack_kernel<<<num_sms() * 4, 256, 0, stream>>>(conv1_weight, hidden, C, hi, lo));
  return SSLAM_OK;
}

extern "C" size_t sslam_selector_workspace_bytes(int B, int H, int W, int C) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
  return 2 * half_bytes((size_t)B * H * W * C) + 256;
}

extern "C" int sslam_selector_head_f32(const float* feat, const void* packed, const float* conv1_bias,
                                       const float* conv2_weight, const float* conv2_bias, int B, int H, int W,
                                       int C, int hidden, int apply_sigmoid, float* out, void* ws,
                                       size_t ws_bytes, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(B >= 0 && H > 0 && W > 0, SSLAM_EINVAL, "selector: bad size");
  if (B == 0) return SSLAM_OK;
  SSLAM_REQUIRE(feat && packed && conv1_bias && conv2_weight && conv2_bias && out && ws, SSLAM_EINVAL,
                "selector: null pointer");
  SSLAM_REQUIRE(C > 0 && C % BK == 0 && hidden > 0 && hidden % 8 == 0 && hidden <= MAX_HIDDEN, SSLAM_EUNSUPPORTED,
                "selector: C must be a multiple of 64, hidden a multiple of 8 and <= 512 (C=%d hidden=%d)", C, hidden);
  SSLAM_REQUIRE((reinterpret_cast<uintptr_t>(feat) & 15) == 0, SSLAM_EINVAL, "selector: features must be 16-byte aligned");
  SSLAM_REQUIRE(ws_bytes >= sslam_selector_workspace_bytes(B, H, W, C), SSLAM_EWORKSPACE,
                "selector: workspace %zu < %zu", ws_bytes, sslam_selector_workspace_bytes(B, H, W, C));
  const size_t n = (size_t)B * H * W * C;
  char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  __half* f_hi = reinterpret_cast<__half*>(w);
  __half* f_lo = reinterpret_cast<__half*>(w + half_bytes(n));
  SSLAM_LAUNCH(KK_SPLIT, stream,
               conv_split_kernel<<<num_sms() * 8, 256, 0, stream>>>(reinterpret_cast<const float4*>(feat),
                                                                    reinterpret_cast<uint2*>(f_hi),
                                                                    reinterpret_cast<uint2*>(f_lo), n / 4));
  // pixel window of a 128-row tile: the (box_h x box_w) shape that covers the grid with the fewest tiles
  ConvParams p;
  p.B = B; p.H = H; p.W = W; p.C = C; p.hidden = hidden;
  long best = -1;
  for (int bw = 8; bw <= 128; bw <<= 1) {
    const int bh = BM / bw;
    if (bw > 8 && bw >= 2 * W) break;                      // wider than the grid (and than its next power of two)
    const long tiles = (long)((W + bw - 1) / bw) * ((H + bh - 1) / bh);
    if (best < 0 || tiles <= best) { best = tiles; p.bw = bw; p.bh = bh; }
  }
  p.tiles_x = (W + p.bw - 1) / p.bw;
  p.tiles_y = (H + p.bh - 1) / p.bh;
  p.b1 = conv1_bias; p.w2 = conv2_weight; p.b2 = conv2_bias; p.apply_sigmoid = apply_sigmoid; p.out = out;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  if ((rc = make_tensor_map_nhwc(&ta_hi, f_hi, B, H, W, C, p.bh, p.bw, BK))) return rc;
  if ((rc = make_tensor_map_nhwc(&ta_lo, f_lo, B, H, W, C, p.bh, p.bw, BK))) return rc;
  const size_t nw = (size_t)hidden * 9 * C;
  const char* pk = static_cast<const char*>(packed);
  if ((rc = make_tensor_map_2d(&tb_hi, pk, hidden, 9 * (uint64_t)C, BN, BK, 2))) return rc;
  if ((rc = make_tensor_map_2d(&tb_lo, pk + half_bytes(nw), hidden, 9 * (uint64_t)C, BN, BK, 2))) return rc;
  static DeviceOnce once;
  if (once.first_use())
    SSLAM_CHECK_CUDA(cudaFuncSetAttribute(conv_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  const int ntiles = B * p.tiles_x * p.tiles_y;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  SSLAM_LAUNCH(KK_CONV_HEAD, stream,
               conv_head_kernel<<<grid, NUM_THREADS, SMEM_TOTAL, stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, p));
  return SSLAM_OK;
}
