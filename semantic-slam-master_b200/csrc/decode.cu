// Heatmap decode for B maps at once, no host synchronisation.
//
// Replaces KeypointSelector.select_keypoints/_apply_nms (models/keypoint_selector.py:69-226).
//
// Data flow (per call, all images in every launch):
//   K1 scan     : ONE DRAM read of the map.  Row slabs stream through a shared-memory ring (bulk async
//                 copies, mbarrier pipeline: the bytes in flight per SM no longer depend on registers),
//                 (2r+1)^2 NMS from a rolling register window, local maxima appended to a per-image
//                 candidate list as 64-bit keys (score bits << 32 | ~linear index) and binned into a
//                 per-image 1024-bin score histogram; while the pixels are in registers the kernel also
//                 counts those below a speculative threshold t_b (see K3)
//   K2 topk     : per image, the histogram locates the bin of the K-th candidate, ONE pass over the
//                 candidate list keeps the keys at or above that bin (K + a few hundred) in shared
//                 memory, bitonic sort -> tentative main-branch output (score desc, index asc), the
//                 K-th key and the count of boundary ties.  (Plateau maps whose boundary bin overflows
//                 the buffer take an exact radix select over the whole list instead.)
//   K3 count    : how many pixels are < the K-th score.  The scan already counted the pixels below
//                 t_b; when t_b <= K-th score that count is a lower bound and usually proves the main
//                 branch, and this kernel exits without touching the map.  Only images whose
//                 speculation failed are read a second time.  t_b is a hint kept in the workspace
//                 (0.95 x the image slot's last K-th score): it never changes a result, only whether
//                 the second pass is needed.
//   K4 resolve  : the main branch (keypoint_selector.py:120-128) is taken iff at least K NMS
//                 survivors exceed thr = max(quantile(map, p), floor).  Because the quantile lies
//                 between order statistics v[lo] <= thr32 <= v[hi], "count(pixels < kth score) >
//                 hi" proves thr32 < kth score without computing the quantile.  When the proof
//                 fails (few maxima, tiny grids, constant maps) this kernel computes the exact
//                 quantile(s) and reproduces the reference's fallback branches (:130-184).
//
// The common case therefore never sorts or histograms the full map; the exact-quantile machinery
// only runs for images that really need the reference's fallback arithmetic.
#include "common.cuh"

namespace sslam {
bool g_decode_no_stream = false;         // tools / tests: force the register-prefetching scan kernels
int g_decode_ns = 0, g_decode_band = 0;  // tools: stages per CTA / rows per band of the streaming scan (0 = default)
namespace {

constexpr int TILE_W = 128;
constexpr int TILE_H = 32;
constexpr int MAX_R = 8;
constexpr int SCAN_THREADS = 256;
constexpr int SEL_THREADS = 1024;
constexpr int RES_THREADS = 512;       // resolve: 128 registers per thread available, no spills
constexpr float LOWER_FLOOR = 0.05f;     // keypoint_selector.py:141

struct __align__(16) ImgHeader {
  u32 cand_count;      // local maxima appended by K1
  u32 below_count;     // pixels strictly below the K-th candidate score (K3)
  u64 kth_key;         // K-th largest candidate key (K2); 0 when fewer than K candidates
  u32 have_tentative;  // K2 wrote a tentative main-branch result
  u32 ties;            // candidates equal to the K-th score left unselected
  u32 spec_below;      // pixels strictly below the speculative threshold hint[b] (K1)
  u32 pad;
};

constexpr int HIST_BINS = 1024;          // candidate scores in [2^-8, 1): 8 binades x 128 bins
__device__ __forceinline__ int score_bin(u32 score_bits) {           // monotone for positive scores
  const int b = (int)(score_bits >> 16) - 0x3b80;
  return b < 0 ? 0 : (b > HIST_BINS - 1 ? HIST_BINS - 1 : b);
}

struct DecodeParams {
  const float* sal;
  int from_logits;
  int B, H, W, K, r;
  float pct, floor;
  float* kpts;
  float* scores;
  int32_t* info;
  ImgHeader* hdr;
  float* hint;         // [B] speculative count thresholds; persists across calls in the workspace
  u32* chist;          // [B][HIST_BINS] candidate score histogram (stream scan only), or null
  u64* cand;           // [B][H*W]
};

__device__ __forceinline__ float load_px(const DecodeParams& p, size_t off) {
  float v = __ldg(p.sal + off);
  return p.from_logits ? sigmoid_f32(v) : v;
}

// ------------------------------------------------------------------------------------------ K1
__global__ void __launch_bounds__(SCAN_THREADS) decode_scan_kernel(DecodeParams p) {
  extern __shared__ float smem[];
  const int r = p.r, H = p.H, W = p.W;
  const int inW = TILE_W + 2 * r, inH = TILE_H + 2 * r;
  float* tin = smem;                 // [inH][inW]
  float* hmx = smem + inH * inW;     // [inH][TILE_W]
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
  const size_t img = (size_t)b * H * W;
  const float NEG_INF = __int_as_float(0xff800000);

  for (int i = threadIdx.x; i < inH * inW; i += SCAN_THREADS) {
    int ty = i / inW, tx = i - ty * inW;
    int y = y0 + ty - r, x = x0 + tx - r;
    float v = NEG_INF;                                    // max_pool2d pads with -inf (:215-220)
    if (y >= 0 && y < H && x >= 0 && x < W) v = load_px(p, img + (size_t)y * W + x);
    tin[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < inH * TILE_W; i += SCAN_THREADS) {
    int ty = i / TILE_W, tx = i - ty * TILE_W;
    const float* row = tin + ty * inW + tx;
    float m = row[0];
    for (int d = 1; d <= 2 * r; ++d) m = fmaxf(m, row[d]);
    hmx[i] = m;
  }
  __syncthreads();
  const float min_keep = fminf(p.floor, LOWER_FLOOR);     // nothing at or below can ever pass
  ImgHeader* hdr = p.hdr + b;
  u64* cand = p.cand + (size_t)b * H * W;
  for (int i = threadIdx.x; i < TILE_H * TILE_W; i += SCAN_THREADS) {
    int ty = i / TILE_W, tx = i - ty * TILE_W;
    int y = y0 + ty, x = x0 + tx;
    bool is_max = false;
    float c = 0.f;
    if (y < H && x < W) {
      const float* col = hmx + ty * TILE_W + tx;
      float m = col[0];
      for (int d = 1; d <= 2 * r; ++d) m = fmaxf(m, col[d * TILE_W]);
      c = tin[(ty + r) * inW + tx + r];
      is_max = (c == m) && (c > min_keep);                // plateaus survive (:223)
    }
    unsigned bal = __ballot_sync(0xffffffffu, is_max);
    if (bal) {
      int lane = threadIdx.x & 31;
      int leader = __ffs(bal) - 1;
      u32 base = 0;
      if (lane == leader) base = atomicAdd(&hdr->cand_count, (u32)__popc(bal));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (is_max) {
        u32 pos = base + __popc(bal & ((1u << lane) - 1));
        u32 lin = (u32)(y * W + x);
        cand[pos] = ((u64)__float_as_uint(c) << 32) | (u64)(0xffffffffu - lin);   // c > 0
      }
    }
  }
}

// K1, fast path for radius 1..3: no shared memory.  One warp owns a vertical strip of
// 32-2R output columns (its 32 lanes load 32 columns, R halo columns per side) and walks down a
// segment of rows: the horizontal (2R+1)-max comes from warp shuffles, the vertical one from a
// rolling register window, rows are prefetched PF deep.  Local maxima are staged in a per-CTA
// shared list and appended to the image's candidate list with ONE global atomic per CTA (one
// same-address L2 atomic costs ~18 ns on B200; per warp-row they dominated the kernel).  Reads each
// pixel once from DRAM (halo re-reads are L1/L2 hits).
template <int R>
__global__ void __launch_bounds__(256) decode_scan_strip_kernel(DecodeParams p, int nstrips, int nseg,
                                                                int seg_rows) {
  constexpr int WIN = 2 * R + 1, OUTW = 32 - 2 * R, PF = 4, CAP = 2048;
  __shared__ u64 sbuf[CAP];
  __shared__ u32 scount, scut, sbase;
  const int lane = threadIdx.x & 31;
  const int per_img = nstrips * nseg;
  const int ctas_per_img = (per_img + 7) / 8;                 // a CTA never straddles two images
  const int b = blockIdx.x / ctas_per_img;
  const int rem = (blockIdx.x - b * ctas_per_img) * 8 + (threadIdx.x >> 5);
  const bool active = rem < per_img;
  const int seg = rem / nstrips, strip = rem - seg * nstrips;
  if (threadIdx.x == 0) { scount = 0; scut = 0xffffffffu; }
  __syncthreads();
  const int H = p.H, W = p.W;
  const int x = strip * OUTW - R + lane;
  const bool col_ok = (x >= 0) && (x < W);
  const bool out_lane = (lane >= R) && (lane < 32 - R) && (x < W);
  const int y0 = seg * seg_rows, y1 = active ? min(H, y0 + seg_rows) : y0 - R;
  const int y_end = y1 + R;                                   // rows [y0-R, y_end) are loaded
  const size_t img = (size_t)b * H * W;
  const float NEG_INF = __int_as_float(0xff800000);
  const float min_keep = fminf(p.floor, LOWER_FLOOR);
  ImgHeader* hdr = p.hdr + b;
  u64* cand = p.cand + (size_t)b * H * W;

  auto ldrow = [&](int y) -> float {
    return (col_ok && y >= 0 && y < H && y < y_end) ? load_px(p, img + (size_t)y * W + x) : NEG_INF;
  };
  float pre[PF];
#pragma unroll
  for (int i = 0; i < PF; ++i) pre[i] = ldrow(y0 - R + i);
  float hwin[WIN], cwin[R + 1];                              // horizontal maxima / centre values
#pragma unroll
  for (int i = 0; i < WIN; ++i) hwin[i] = NEG_INF;
#pragma unroll
  for (int i = 0; i <= R; ++i) cwin[i] = NEG_INF;

  for (int yy = y0 - R; yy < y_end; yy += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int yc = yy + u;
      const float v = pre[u];
      pre[u] = ldrow(yc + PF);
      if (yc < y_end) {                                       // warp-uniform
        float hm = v;
#pragma unroll
        for (int d = 1; d <= R; ++d) {
          hm = fmaxf(hm, __shfl_up_sync(0xffffffffu, v, d));
          hm = fmaxf(hm, __shfl_down_sync(0xffffffffu, v, d));
        }
#pragma unroll
        for (int i = 0; i < WIN - 1; ++i) hwin[i] = hwin[i + 1];
        hwin[WIN - 1] = hm;
#pragma unroll
        for (int i = 0; i < R; ++i) cwin[i] = cwin[i + 1];
        cwin[R] = v;
        const int yo = yc - R;                                // output row whose window is complete
        if (yo >= y0) {                                       // warp-uniform
          float m = hwin[0];
#pragma unroll
          for (int i = 1; i < WIN; ++i) m = fmaxf(m, hwin[i]);
          const float c = cwin[0];
          const bool is_max = out_lane && (c == m) && (c > min_keep);
          const unsigned bal = __ballot_sync(0xffffffffu, is_max);
          if (bal) {
            const int leader = __ffs(bal) - 1;
            const u32 n = (u32)__popc(bal);
            u32 base = 0;
            bool in_smem = false;
            if (lane == leader) {
              base = atomicAdd(&scount, n);
              in_smem = (base + n <= (u32)CAP);
              if (!in_smem) {                                  // staging list full (plateau maps)
                atomicMin(&scut, base);
                base = atomicAdd(&hdr->cand_count, n);
              }
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            in_smem = __shfl_sync(0xffffffffu, (int)in_smem, leader) != 0;
            if (is_max) {
              const u32 pos = base + __popc(bal & ((1u << lane) - 1));
              const u32 lin = (u32)(yo * W + x);
              const u64 key = ((u64)__float_as_uint(c) << 32) | (u64)(0xffffffffu - lin);
              if (in_smem) sbuf[pos] = key; else cand[pos] = key;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const u32 nstaged = min(scount, scut);
  if (threadIdx.x == 0 && nstaged) sbase = atomicAdd(&hdr->cand_count, nstaged);
  __syncthreads();
  for (u32 i = threadIdx.x; i < nstaged; i += blockDim.x) cand[sbase + i] = sbuf[i];
}

// K1, vector path (radius 1..3, W % 4 == 0, 16-byte aligned maps): as above, but every lane owns FOUR
// consecutive columns loaded as one float4, so a warp row is one 512-byte request and the per-pixel
// instruction count drops ~3x (the scalar kernel is issue bound at 14 % of HBM bandwidth).  Lanes
// 0 and 31 are halo lanes (a whole float4 each side, which keeps the loads aligned); lanes 1..30
// produce 120 output columns.  The horizontal maxima use the neighbours' edge values (2R
// shuffles), the vertical ones a ring of the last 2R+1 row maxima with compile-time slots.
template <int R>
__global__ void __launch_bounds__(256) decode_scan_vec_kernel(DecodeParams p, int nstrips, int nseg,
                                                              int seg_rows) {
  constexpr int WIN = 2 * R + 1, OUTW = 120, CAP = 2048;
  __shared__ u64 sbuf[CAP];
  __shared__ u32 scount, sbase;
  const int lane = threadIdx.x & 31;
  const int per_img = nstrips * nseg;
  const int ctas_per_img = (per_img + 7) / 8;                 // a CTA never straddles two images
  const int b = blockIdx.x / ctas_per_img;
  const int rem = (blockIdx.x - b * ctas_per_img) * 8 + (threadIdx.x >> 5);
  const bool active = rem < per_img;
  const int seg = rem / nstrips, strip = rem - seg * nstrips;
  if (threadIdx.x == 0) scount = 0;
  __syncthreads();
  const int H = p.H, W = p.W;
  const int x = strip * OUTW - 4 + 4 * lane;                  // first of this lane's four columns
  const bool col_ok = (x >= 0) && (x < W);                    // W % 4 == 0: all four or none
  const bool out_lane = (lane >= 1) && (lane <= 30) && col_ok;
  const int y0 = seg * seg_rows, y1 = active ? min(H, y0 + seg_rows) : y0 - R;
  const int y_end = y1 + R;                                   // rows [y0-R, y_end) are loaded
  const size_t img = (size_t)b * H * W;
  const float NEG_INF = __int_as_float(0xff800000);
  const float min_keep = fminf(p.floor, LOWER_FLOOR);
  ImgHeader* hdr = p.hdr + b;
  u64* cand = p.cand + (size_t)b * H * W;

  auto ldrow = [&](int y) -> float4 {
    float4 v = make_float4(NEG_INF, NEG_INF, NEG_INF, NEG_INF);
    if (col_ok && y >= 0 && y < H && y < y_end) {
      v = __ldg(reinterpret_cast<const float4*>(p.sal + img + (size_t)y * W + x));
      if (p.from_logits) { v.x = sigmoid_f32(v.x); v.y = sigmoid_f32(v.y); v.z = sigmoid_f32(v.z); v.w = sigmoid_f32(v.w); }
    }
    return v;
  };
  float4 pre[WIN];
#pragma unroll
  for (int i = 0; i < WIN; ++i) pre[i] = ldrow(y0 - R + i);
  float hwin[WIN][4], cwin[WIN][4];                          // rings: row maxima / centre values
#pragma unroll
  for (int i = 0; i < WIN; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) { hwin[i][c] = NEG_INF; cwin[i][c] = NEG_INF; }

  for (int yy = y0 - R; yy < y_end; yy += WIN) {
#pragma unroll
    for (int u = 0; u < WIN; ++u) {
      const int yc = yy + u;
      const float4 v4 = pre[u];
      pre[u] = ldrow(yc + WIN);
      if (yc < y_end) {                                       // warp-uniform
        // e[] = R values of the left neighbour, own four, R values of the right neighbour
        float e[4 + 2 * R];
        const float own[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) e[R + c] = own[c];
#pragma unroll
        for (int d = 0; d < R; ++d) {
          e[d] = __shfl_up_sync(0xffffffffu, own[4 - R + d], 1);
          e[R + 4 + d] = __shfl_down_sync(0xffffffffu, own[d], 1);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float hm = e[c];
#pragma unroll
          for (int d = 1; d < WIN; ++d) hm = fmaxf(hm, e[c + d]);
          hwin[u][c] = hm;
          cwin[u][c] = own[c];
        }
        const int yo = yc - R;                                // output row whose window is complete
        if (yo >= y0) {                                       // warp-uniform
          unsigned mask = 0;
          float cv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float m = hwin[0][c];
#pragma unroll
            for (int i = 1; i < WIN; ++i) m = fmaxf(m, hwin[i][c]);
            cv[c] = cwin[(u + WIN - R) % WIN][c];
            if (out_lane && (cv[c] == m) && (cv[c] > min_keep)) mask |= 1u << c;
          }
          // local maxima are sparse (~4 % of the pixels) and a lane rarely holds more than one: the
          // lanes that found any append them themselves, one loop trip per maximum (order in the
          // list is irrelevant, K2 sorts keys)
          while (mask) {
            const int c = __ffs((int)mask) - 1;
            mask &= mask - 1;
            const float val = c == 0 ? cv[0] : c == 1 ? cv[1] : c == 2 ? cv[2] : cv[3];
            const u32 lin = (u32)(yo * W + x + c);
            const u64 key = ((u64)__float_as_uint(val) << 32) | (u64)(0xffffffffu - lin);
            const u32 pos = atomicAdd(&scount, 1u);
            if (pos < (u32)CAP) sbuf[pos] = key;
            else cand[atomicAdd(&hdr->cand_count, 1u)] = key;   // staging list full (plateau maps)
          }
        }
      }
    }
  }
  __syncthreads();
  const u32 nstaged = min(scount, (u32)CAP);
  if (threadIdx.x == 0 && nstaged) sbase = atomicAdd(&hdr->cand_count, nstaged);
  __syncthreads();
  for (u32 i = threadIdx.x; i < nstaged; i += blockDim.x) cand[sbase + i] = sbuf[i];
}

// K1, streaming path (radius 1..3, W % 4 == 0, 16-byte aligned maps, W <= 1800): the arithmetic of the
// vector kernel above, fed from shared memory.  A CTA owns a band of output rows of one image over
// the full width; a producer warp streams the band (plus R halo rows each side) through a ring of
// NS stages of WIN = 2R+1 rows with 1-D bulk async copies (rows of a band are contiguous), each
// completing on the stage's `full` mbarrier; consumer warp w owns the 120-column strip w, reads its
// lane's float4 of every row from the stage, and arrives on the stage's `empty` mbarrier when done
// with it.  Bytes in flight per SM = stages x CTAs, independent of register count (the register-
// prefetching kernel had 16 warps x 5 rows = 40 KB per SM in flight and reached 0.22 of HBM peak).
// Also fills the candidate score histogram and the speculative below-threshold count.
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_wait(u32 bar, u32 parity) {
  u32 ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  }
}

constexpr int STREAM_MAX_WARPS = 15;     // consumer warps = strips of 120 columns

template <int R, bool LOGITS>
__global__ void __launch_bounds__(32 * (STREAM_MAX_WARPS + 1), R <= 2 ? 2 : 1)   // R <= 2: <= 64 registers, four 224-thread CTAs per SM at W = 640
decode_scan_stream_kernel(DecodeParams p, int nstrips, int nbands, int band_rows, int NS) {
  constexpr int WIN = 2 * R + 1, OUTW = 120, CAP = 2048;
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ u64 sbuf[CAP];
  __shared__ u32 shist[HIST_BINS];
  __shared__ __align__(8) unsigned long long bars[2 * 16];     // full[NS], empty[NS]  (NS <= 16)
  __shared__ u32 scount, sbase;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = p.H, W = p.W;
  const int b = blockIdx.x / nbands, band = blockIdx.x - b * nbands;
  const int y0 = band * band_rows, y1 = min(H, y0 + band_rows);
  const int first = y0 - R, y_end = y1 + R;                   // rows [first, y_end) enter the window
  const int nstage = (y_end - first + WIN - 1) / WIN;
  const size_t img = (size_t)b * H * W;
  const size_t stage_bytes = (size_t)WIN * W * 4;
  const u32 full0 = smem_addr(&bars[0]), empty0 = smem_addr(&bars[16]);

  for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) shist[i] = 0;
  if (threadIdx.x == 0) {
    scount = 0;
    for (int s = 0; s < NS; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0 + 8 * s) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * s), "r"(nstrips) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == nstrips) {
    // ---- producer: stage s holds rows first + s*WIN .. +WIN-1; rows outside the image are not
    // loaded (the consumers mask them by index)
    if (lane == 0) {
      int slot = 0; u32 lap = 0;                              // s = lap * NS + slot
      for (int s = 0; s < nstage; ++s) {
        if (lap) bar_wait(empty0 + 8 * slot, (lap - 1) & 1);
        const int r0 = first + s * WIN;
        const int a = max(r0, 0), e = min(min(r0 + WIN, H), y_end);
        const u32 bytes = e > a ? (u32)(e - a) * (u32)W * 4u : 0u;
        const u32 bar = full0 + 8 * slot;
        if (bytes) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          const u32 dst = smem_addr(dsm + slot * stage_bytes + (size_t)(a - r0) * W * 4);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(p.sal + img + (size_t)a * W), "r"(bytes), "r"(bar) : "memory");
        } else {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
        if (++slot == NS) { slot = 0; ++lap; }
      }
    }
  } else if (warp < nstrips) {
    // ---- consumer warp = one 120-column strip
    const int x = warp * OUTW - 4 + 4 * lane;                 // first of this lane's four columns
    const bool col_ok = (x >= 0) && (x < W);                  // W % 4 == 0: all four or none
    const bool out_lane = (lane >= 1) && (lane <= 30) && col_ok;
    const float NEG_INF = __int_as_float(0xff800000);
    const float min_keep = fminf(p.floor, LOWER_FLOOR);
    // speculative count: pixels of output rows / output lanes below hint[b] (counted exactly once)
    const float tspec = out_lane ? p.hint[b] : NEG_INF;
    u32 below = 0;
    ImgHeader* hdr = p.hdr + b;
    u64* cand = p.cand + (size_t)b * H * W;
    float hwin[WIN][4], cwin[WIN][4];                          // rings: row maxima / centre values
#pragma unroll
    for (int i = 0; i < WIN; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) { hwin[i][c] = NEG_INF; cwin[i][c] = NEG_INF; }

    int slot = 0; u32 lap = 0;
    for (int s = 0; s < nstage; ++s) {
      bar_wait(full0 + 8 * slot, lap & 1);
      const unsigned char* st = dsm + slot * stage_bytes;
      const int yy = first + s * WIN;
#pragma unroll
      for (int u = 0; u < WIN; ++u) {
        const int yc = yy + u;
        if (yc < y_end) {                                      // warp-uniform
          float4 v4 = make_float4(NEG_INF, NEG_INF, NEG_INF, NEG_INF);
          if (col_ok && yc >= 0 && yc < H) {
            v4 = *reinterpret_cast<const float4*>(st + ((size_t)u * W + x) * 4);
            if (LOGITS) { v4.x = sigmoid_f32(v4.x); v4.y = sigmoid_f32(v4.y); v4.z = sigmoid_f32(v4.z); v4.w = sigmoid_f32(v4.w); }
          }
          float e[4 + 2 * R];
          const float own[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) e[R + c] = own[c];
#pragma unroll
          for (int d = 0; d < R; ++d) {
            e[d] = __shfl_up_sync(0xffffffffu, own[4 - R + d], 1);
            e[R + 4 + d] = __shfl_down_sync(0xffffffffu, own[d], 1);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float hm = e[c];
#pragma unroll
            for (int d = 1; d < WIN; ++d) hm = fmaxf(hm, e[c + d]);
            hwin[u][c] = hm;
            cwin[u][c] = own[c];
          }
          const int yo = yc - R;                               // output row whose window is complete
          if (yo >= y0) {                                      // warp-uniform
            unsigned mask = 0;
            float cv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float m = hwin[0][c];
#pragma unroll
              for (int i = 1; i < WIN; ++i) m = fmaxf(m, hwin[i][c]);
              cv[c] = cwin[(u + WIN - R) % WIN][c];
              below += (cv[c] < tspec) ? 1u : 0u;
              if (out_lane && (cv[c] == m) && (cv[c] > min_keep)) mask |= 1u << c;
            }
            while (mask) {                                     // sparse: ~4 % of the pixels
              const int c = __ffs((int)mask) - 1;
              mask &= mask - 1;
              const float val = c == 0 ? cv[0] : c == 1 ? cv[1] : c == 2 ? cv[2] : cv[3];
              const u32 lin = (u32)(yo * W + x + c);
              const u32 bits = __float_as_uint(val);
              const u64 key = ((u64)bits << 32) | (u64)(0xffffffffu - lin);
              atomicAdd(&shist[score_bin(bits)], 1u);
              const u32 pos = atomicAdd(&scount, 1u);
              if (pos < (u32)CAP) sbuf[pos] = key;
              else cand[atomicAdd(&hdr->cand_count, 1u)] = key;   // staging list full (plateau maps)
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * slot) : "memory");
      if (++slot == NS) { slot = 0; ++lap; }
    }
    below = warp_reduce_sum(below);
    if (lane == 0 && below) atomicAdd(&hdr->spec_below, below);
  }
  __syncthreads();
  const u32 nstaged = min(scount, (u32)CAP);
  if (threadIdx.x == 0 && nstaged) sbase = atomicAdd(&p.hdr[b].cand_count, nstaged);
  __syncthreads();
  u64* cand = p.cand + (size_t)b * H * W;
  for (u32 i = threadIdx.x; i < nstaged; i += blockDim.x) cand[sbase + i] = sbuf[i];
  u32* gh = p.chist + (size_t)b * HIST_BINS;
  for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) {
    const u32 h = shist[i];
    if (h) atomicAdd(gh + i, h);
  }
}

// Plain NMS output (drop-in for _apply_nms)
__global__ void __launch_bounds__(SCAN_THREADS) nms_kernel(const float* sal, int H, int W, int r,
                                                           float* out) {
  extern __shared__ float smem[];
  const int inW = TILE_W + 2 * r, inH = TILE_H + 2 * r;
  float* tin = smem;
  float* hmx = smem + inH * inW;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
  const size_t img = (size_t)b * H * W;
  const float NEG_INF = __int_as_float(0xff800000);
  for (int i = threadIdx.x; i < inH * inW; i += SCAN_THREADS) {
    int ty = i / inW, tx = i - ty * inW;
    int y = y0 + ty - r, x = x0 + tx - r;
    float v = NEG_INF;
    if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(sal + img + (size_t)y * W + x);
    tin[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < inH * TILE_W; i += SCAN_THREADS) {
    int ty = i / TILE_W, tx = i - ty * TILE_W;
    const float* row = tin + ty * inW + tx;
    float m = row[0];
    for (int d = 1; d <= 2 * r; ++d) m = fmaxf(m, row[d]);
    hmx[i] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TILE_H * TILE_W; i += SCAN_THREADS) {
    int ty = i / TILE_W, tx = i - ty * TILE_W;
    int y = y0 + ty, x = x0 + tx;
    if (y < H && x < W) {
      const float* col = hmx + ty * TILE_W + tx;
      float m = col[0];
      for (int d = 1; d <= 2 * r; ++d) m = fmaxf(m, col[d * TILE_W]);
      float c = tin[(ty + r) * inW + tx + r];
      out[img + (size_t)y * W + x] = __fmul_rn(c, (c == m) ? 1.0f : 0.0f);          // :223-224
    }
  }
}

// ------------------------------------------------------------------------------------------ K2
struct CandSrc {
  const u64* keys;
  __device__ __forceinline__ bool operator()(int i, u64& k) const { k = keys[i]; return true; }
};

// Select the `count` largest keys of a filtered stream, sort them descending in `buf`
// (capacity pow2 >= count) and return through `buf`.  Returns the count-th largest key.
template <typename Src>
__device__ u64 block_topk_sorted(Src src, int n, int count, u64* buf, int cap_pow2,
                                 SelectScratch* ss, u32* counter) {
  u64 kth = block_select_kth_largest(src, n, (u32)count, ss);
  if (threadIdx.x == 0) *counter = 0;
  for (int i = threadIdx.x; i < cap_pow2; i += blockDim.x) buf[i] = 0;
  __syncthreads();
  // keys strictly above kth first, then as many == kth as needed (keys may repeat in raw mode)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    u64 k;
    if (src(i, k) && k > kth) buf[atomicAdd(counter, 1u)] = k;
  }
  __syncthreads();
  u32 above = *counter;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    u64 k;
    if (src(i, k) && k == kth) {
      u32 pos = atomicAdd(counter, 1u);
      if (pos < (u32)count) buf[pos] = k;
    }
  }
  (void)above;
  __syncthreads();
  block_bitonic_sort_desc(buf, cap_pow2);
  return kth;
}

__device__ __forceinline__ void write_row(const DecodeParams& p, int b, int row, u32 lin, float sc) {
  size_t o = (size_t)b * p.K + row;
  p.kpts[2 * o] = (float)(lin % (u32)p.W);       // x = column  (:127)
  p.kpts[2 * o + 1] = (float)(lin / (u32)p.W);   // y = row
  p.scores[o] = sc;
}

__global__ void __launch_bounds__(SEL_THREADS) decode_topk_kernel(DecodeParams p, int kpad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* buf = reinterpret_cast<u64*>(smem_raw);                        // [kpad]
  SelectScratch* ss = reinterpret_cast<SelectScratch*>(buf + kpad);
  __shared__ u32 counter;
  __shared__ u32 eq_total;
  const int b = blockIdx.x;
  ImgHeader* hdr = p.hdr + b;
  const int n = (int)hdr->cand_count;
  if (n < p.K) {                       // cannot be the main branch; K4 takes the exact path
    if (threadIdx.x == 0) { hdr->kth_key = 0; hdr->have_tentative = 0; hdr->ties = 0; }
    return;
  }
  CandSrc src{p.cand + (size_t)b * p.H * p.W};
  if (threadIdx.x == 0) eq_total = 0;
  u64 kth = block_topk_sorted(src, n, p.K, buf, kpad, ss, &counter);
  for (int i = threadIdx.x; i < p.K; i += blockDim.x) {
    u64 k = buf[i];
    write_row(p, b, i, key_index(k), __uint_as_float((u32)(k >> 32)));
  }
  // ties at the k-th boundary: same score, not selected
  u32 ksc = (u32)(kth >> 32);
  u32 local = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    u64 k = src.keys[i];
    if ((u32)(k >> 32) == ksc && k < kth) ++local;
  }
  local = warp_reduce_sum(local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&eq_total, local);
  __syncthreads();
  if (threadIdx.x == 0) { hdr->kth_key = kth; hdr->have_tentative = 1; hdr->ties = eq_total; }
}

// K2, histogram-guided: the stream scan binned every candidate's score into chist[b][1024].  A suffix
// sum from the top bin finds the bin of the K-th candidate; one pass over the candidate list keeps
// the keys at or above that bin in shared memory (K plus the population of the boundary bin) and a
// bitonic sort orders them.  256 threads and ~K*16 bytes of shared memory per CTA: several images
// per SM are in flight (the per-image work is a chain of block-wide barriers, so residency, not
// issue slots, sets its throughput; the 1024-thread kernel below ran one image per SM).
// Falls back to the exact radix select over the whole list when the boundary bin overflows `cap`.
constexpr int TOPK2_THREADS = 256;

__global__ void __launch_bounds__(TOPK2_THREADS) decode_topk_hist_kernel(DecodeParams p, int cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* buf = reinterpret_cast<u64*>(smem_raw);                        // [cap], cap = pow2 >= 2K
  __shared__ u32 wsum[TOPK2_THREADS / 32];
  __shared__ u32 counter, cut_bin, list_len, eq_total;
  __shared__ SelectScratch ss_fallback;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ImgHeader* hdr = p.hdr + b;
  const int n = (int)hdr->cand_count;
  if (n < p.K) {                       // cannot be the main branch; K4 takes the exact path
    if (tid == 0) { hdr->kth_key = 0; hdr->have_tentative = 0; hdr->ties = 0; }
    return;
  }
  const u64* cand = p.cand + (size_t)b * p.H * p.W;
  const u32* hist = p.chist + (size_t)b * HIST_BINS;
  // thread t owns bins [4*(255-t) .. +3] so that thread order = descending score
  const int top = HIST_BINS - 1 - 4 * tid;
  u32 h[4], mine = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { h[i] = hist[top - i]; mine += h[i]; }
  u32 incl = mine;
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) {
    const u32 t = __shfl_up_sync(0xffffffffu, incl, sft);
    if (lane >= sft) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  if (tid == 0) { counter = 0; eq_total = 0; }
  __syncthreads();
  u32 before = 0;
  for (int w = 0; w < warp; ++w) before += wsum[w];
  const u32 excl = before + incl - mine;                    // candidates in bins above this thread's
  if ((u32)p.K > excl && (u32)p.K <= excl + mine) {         // exactly one thread
    u32 c = excl;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (c + h[i] >= (u32)p.K) { cut_bin = (u32)(top - i); list_len = c + h[i]; break; }
      c += h[i];
    }
  }
  __syncthreads();
  const int cut = (int)cut_bin;
  const int L = (int)list_len;                               // K <= L
  u64 kth;
  if (L <= cap) {
    int cap2 = 1;
    while (cap2 < L) cap2 <<= 1;
    for (int i = tid; i < cap2; i += TOPK2_THREADS) buf[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += TOPK2_THREADS) {
      const u64 k = cand[i];
      if (score_bin((u32)(k >> 32)) >= cut) buf[atomicAdd(&counter, 1u)] = k;
    }
    __syncthreads();
    block_bitonic_sort_desc(buf, cap2);
    kth = buf[p.K - 1];
    const u32 ksc = (u32)(kth >> 32);
    u32 local = 0;
    for (int i = p.K + tid; i < L; i += TOPK2_THREADS) local += ((u32)(buf[i] >> 32) == ksc) ? 1u : 0u;
    local = warp_reduce_sum(local);
    if (lane == 0 && local) atomicAdd(&eq_total, local);
  } else {
    // boundary bin holds more keys than the buffer (plateaus / quantised maps): exact select
    CandSrc src{cand};
    kth = block_topk_sorted(src, n, p.K, buf, cap >= 2 * p.K ? (cap >> 1) : cap, &ss_fallback, &counter);
    const u32 ksc = (u32)(kth >> 32);
    u32 local = 0;
    for (int i = tid; i < n; i += TOPK2_THREADS) {
      const u64 k = cand[i];
      if ((u32)(k >> 32) == ksc && k < kth) ++local;
    }
    local = warp_reduce_sum(local);
    if (lane == 0 && local) atomicAdd(&eq_total, local);
  }
  __syncthreads();
  for (int i = tid; i < p.K; i += TOPK2_THREADS) {
    const u64 k = buf[i];
    write_row(p, b, i, key_index(k), __uint_as_float((u32)(k >> 32)));
  }
  if (tid == 0) { hdr->kth_key = kth; hdr->have_tentative = 1; hdr->ties = eq_total; }
}

// ------------------------------------------------------------------------------------------ K3
__global__ void __launch_bounds__(256) decode_count_kernel(DecodeParams p, int chunks_per_img) {
  __shared__ u32 red[33];
  const int b = blockIdx.x / chunks_per_img;
  const int chunk = blockIdx.x - b * chunks_per_img;
  ImgHeader* hdr = p.hdr + b;
  if (!hdr->have_tentative) return;
  const float kth = __uint_as_float((u32)(hdr->kth_key >> 32));
  const int n = p.H * p.W;
  {
    // the scan counted the pixels below hint[b]; when hint[b] <= kth that is a lower bound of what this
    // kernel would count, and if it already exceeds the quantile's upper rank the map is not read again
    const int hi = (int)ceilf(__fmul_rn(p.pct, (float)(n - 1)));
    if (p.hint[b] <= kth && hdr->spec_below > (u32)hi) return;
  }
  const size_t img = (size_t)b * n;
  int per = (n + chunks_per_img - 1) / chunks_per_img;
  per = (per + 3) & ~3;                                       // chunk boundaries stay 16-byte aligned
  const int beg = chunk * per, end = min(n, beg + per);
  u32 c = 0;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(p.sal) & 15) == 0) {
    for (int i = beg + 4 * (int)threadIdx.x; i < end; i += 4 * 256) {        // n % 4 == 0: whole float4s
      float4 v = __ldg(reinterpret_cast<const float4*>(p.sal + img + i));
      if (p.from_logits) { v.x = sigmoid_f32(v.x); v.y = sigmoid_f32(v.y); v.z = sigmoid_f32(v.z); v.w = sigmoid_f32(v.w); }
      c += (v.x < kth ? 1u : 0u) + (v.y < kth ? 1u : 0u) + (v.z < kth ? 1u : 0u) + (v.w < kth ? 1u : 0u);
    }
  } else {
    for (int i = beg + threadIdx.x; i < end; i += 256) c += (load_px(p, img + i) < kth) ? 1u : 0u;
  }
  c = block_reduce_sum(c, red);
  if (threadIdx.x == 0 && c) atomicAdd(&hdr->below_count, c);
}

// ------------------------------------------------------------------------------------------ K4
struct PixelValSrc {                    // all pixels, key = ordered value (for quantiles)
  DecodeParams p; size_t img;
  __device__ __forceinline__ bool operator()(int i, u64& k) const {
    k = (u64)ordered_from_float(load_px(p, img + i));
    return true;
  }
};
struct PixelKeySrc {                    // all pixels, key = (value, ~index)  (raw topk, :166,178)
  DecodeParams p; size_t img;
  __device__ __forceinline__ bool operator()(int i, u64& k) const {
    k = pack_key(load_px(p, img + i), (u32)i);
    return true;
  }
};
struct CandBandSrc {                    // candidates with lo < score <= hi  (hi = +inf: score > lo)
  const u64* keys; float lo, hi;
  __device__ __forceinline__ bool operator()(int i, u64& k) const {
    k = keys[i];
    float s = __uint_as_float((u32)(k >> 32));
    return (s > lo) && !(s > hi);
  }
};

// torch.quantile(flat, q) (keypoint_selector.py:106,140), then max(., floor) as fp32
__device__ float block_exact_threshold(const DecodeParams& p, size_t img, float q, float floor,
                                       SelectScratch* ss) {
  const int n = p.H * p.W;
  float rank = __fmul_rn(q, (float)(n - 1));
  int lo = (int)floorf(rank), hi = (int)ceilf(rank);
  float w = __fsub_rn(rank, (float)lo);
  PixelValSrc src{p, img};
  // ascending index lo == (n - lo)-th largest
  float vlo = float_from_ordered((u32)block_select_kth_largest(src, n, (u32)(n - lo), ss));
  float vhi = vlo;
  if (hi != lo) vhi = float_from_ordered((u32)block_select_kth_largest(src, n, (u32)(n - hi), ss));
  return fmaxf(lerp_aten(vlo, vhi, w), floor);
}

__device__ u32 block_count_band(const u64* keys, int n, float lo, float hi, u32* red) {
  CandBandSrc s{keys, lo, hi};
  u32 c = 0;
  u64 k;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += s(i, k) ? 1u : 0u;
  return block_reduce_sum(c, red);
}

__global__ void __launch_bounds__(RES_THREADS) decode_resolve_kernel(DecodeParams p, int kpad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* buf = reinterpret_cast<u64*>(smem_raw);
  SelectScratch* ss = reinterpret_cast<SelectScratch*>(buf + kpad);
  __shared__ u32 counter;
  __shared__ u32 red[33];
  const int b = blockIdx.x;
  ImgHeader* hdr = p.hdr + b;
  const int npx = p.H * p.W;
  const size_t img = (size_t)b * npx;
  const int ncand = (int)hdr->cand_count;
  const u64* cand = p.cand + (size_t)b * npx;
  const float INF = __int_as_float(0x7f800000);
  int32_t* info = p.info ? p.info + 4 * b : nullptr;

  // ---- fast proof of the main branch
  {
    float rank = __fmul_rn(p.pct, (float)(npx - 1));
    int hi = (int)ceilf(rank);
    float kth = __uint_as_float((u32)(hdr->kth_key >> 32));
    const float hint = p.hint[b];
    const bool have = hdr->have_tentative != 0;
    const bool proven = hdr->below_count > (u32)hi || (hint <= kth && hdr->spec_below > (u32)hi);
    __syncthreads();                                           // every thread has read the old hint
    if (have && threadIdx.x == 0) p.hint[b] = 0.95f * kth;     // next call's speculative threshold
    if (have && proven && kth > p.floor) {
      if (info && threadIdx.x == 0) {
        info[0] = 0; info[1] = -1; info[2] = (int32_t)hdr->ties; info[3] = ncand;
      }
      return;
    }
  }
  // ---- exact path (keypoint_selector.py:105-117)
  const float thr = block_exact_threshold(p, img, p.pct, p.floor, ss);
  const int n = (int)block_count_band(cand, ncand, thr, INF, red);
  if (n >= p.K) {                                            // main branch after all (:120-128)
    if (info && threadIdx.x == 0) {
      info[0] = 0; info[1] = n; info[2] = (int32_t)hdr->ties; info[3] = ncand;
    }
    return;                                                  // K2's result stands
  }
  int branch, ties = 0, row0 = 0, remaining = p.K;
  if (n > 0) {                                               // :130-137
    // first group: the n candidates in row-major order -> sort by ascending linear index
    if (threadIdx.x == 0) counter = 0;
    for (int i = threadIdx.x; i < kpad; i += blockDim.x) buf[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
      u64 k = cand[i];
      if (__uint_as_float((u32)(k >> 32)) > thr)
        buf[atomicAdd(&counter, 1u)] = ((k & 0xffffffffull) << 32) | (k >> 32);   // (~lin, score)
    }
    __syncthreads();
    block_bitonic_sort_desc(buf, kpad);                      // ~lin descending == lin ascending
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      u64 k = buf[i];
      write_row(p, b, i, 0xffffffffu - (u32)(k >> 32), __uint_as_float((u32)k));
    }
    __syncthreads();
    row0 = n;
    remaining = p.K - n;                                     // :136
    branch = -2;
    const float pcts[4] = {0.40f, 0.30f, 0.20f, 0.10f};      // :139
    for (int t = 0; t < 4; ++t) {
      float lthr = block_exact_threshold(p, img, pcts[t], LOWER_FLOOR, ss);   // :140-141
      // additional = nms > lthr and not valid (:143)
      int m = (int)block_count_band(cand, ncand, lthr, thr, red);
      if (m >= remaining) {                                  // :147-156
        CandBandSrc src{cand, lthr, thr};
        u64 kth = block_topk_sorted(src, ncand, remaining, buf, kpad, ss, &counter);
        for (int i = threadIdx.x; i < remaining; i += blockDim.x) {
          u64 k = buf[i];
          write_row(p, b, row0 + i, key_index(k), __uint_as_float((u32)(k >> 32)));
        }
        u32 ksc = (u32)(kth >> 32), local = 0;
        u64 k;
        for (int i = threadIdx.x; i < ncand; i += blockDim.x)
          if (src(i, k) && (u32)(k >> 32) == ksc && k < kth) ++local;
        ties = (int)block_reduce_sum(local, red);
        branch = 1;
        break;
      }
    }
    if (branch == -2) branch = 2;                            // for/else (:157-173)
  } else {
    branch = 3;                                              // :174-184
  }
  if (branch == 2 || branch == 3) {
    if (remaining > npx) {                                   // topk raises (:166,178)
      if (info && threadIdx.x == 0) { info[0] = -1; info[1] = n; info[2] = 0; info[3] = ncand; }
      return;
    }
    PixelKeySrc src{p, img};
    u64 kth = block_topk_sorted(src, npx, remaining, buf, kpad, ss, &counter);
    for (int i = threadIdx.x; i < remaining; i += blockDim.x) {
      u64 k = buf[i];
      write_row(p, b, row0 + i, key_index(k), key_value(k));
    }
    u32 ksc = (u32)(kth >> 32), local = 0;
    u64 k;
    for (int i = threadIdx.x; i < npx; i += blockDim.x)
      if (src(i, k) && (u32)(k >> 32) == ksc && k < kth) ++local;
    ties = (int)block_reduce_sum(local, red);
  }
  // keypoint_selector.py:186-199 (trim / duplicate padding) is unreachable: every branch above
  // produced exactly K rows.
  if (info && threadIdx.x == 0) { info[0] = branch; info[1] = n; info[2] = ties; info[3] = ncand; }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

size_t hdr_bytes(int B) { return align_up((size_t)B * sizeof(ImgHeader), 256); }
size_t hint_bytes(int B) { return align_up((size_t)B * sizeof(float), 256); }
size_t hist_bytes(int B) { return align_up((size_t)B * HIST_BINS * sizeof(u32), 256); }

}  // namespace
}  // namespace sslam

using namespace sslam;

extern "C" size_t sslam_decode_workspace_bytes(int B, int H, int W, int K) {
  (void)K;
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return hdr_bytes(B) + hint_bytes(B) + hist_bytes(B) + (size_t)B * H * W * sizeof(u64);
}

extern "C" int sslam_decode_topk_f32(const float* sal, int from_logits, int B, int H, int W, int K,
                                     int nms_radius, float pct, float floor, float* kpts_xy,
                                     float* scores, int32_t* info, void* ws, size_t ws_bytes,
                                     void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(B >= 0 && H > 0 && W > 0 && K >= 0, SSLAM_EINVAL, "decode: negative size");
  if (B == 0 || K == 0) return SSLAM_OK;
  SSLAM_REQUIRE(sal && kpts_xy && scores && ws, SSLAM_EINVAL, "decode: null pointer");
  SSLAM_REQUIRE(floor >= 0.f, SSLAM_EINVAL, "decode: floor must be >= 0");
  SSLAM_REQUIRE(pct >= 0.f && pct <= 1.f, SSLAM_EINVAL,
                "decode: quantile() q values must be in the range [0, 1]");
  SSLAM_REQUIRE(nms_radius >= 0 && nms_radius <= MAX_R, SSLAM_EUNSUPPORTED,
                "decode: nms_radius %d outside 0..%d", nms_radius, MAX_R);
  SSLAM_REQUIRE(K <= 16384, SSLAM_EUNSUPPORTED, "decode: K=%d > 16384", K);
  SSLAM_REQUIRE((long long)H * W <= (1ll << 24), SSLAM_EUNSUPPORTED,
                "decode: quantile() input tensor is too large (H*W > 2^24)");
  SSLAM_REQUIRE(ws_bytes >= sslam_decode_workspace_bytes(B, H, W, K), SSLAM_EWORKSPACE,
                "decode: workspace %zu < %zu", ws_bytes, sslam_decode_workspace_bytes(B, H, W, K));

  DecodeParams p;
  p.sal = sal; p.from_logits = from_logits; p.B = B; p.H = H; p.W = W; p.K = K; p.r = nms_radius;
  p.pct = pct; p.floor = floor; p.kpts = kpts_xy; p.scores = scores; p.info = info;
  // workspace: [headers | hints (NOT cleared: they persist from call to call) | candidate histograms | keys]
  p.hdr = reinterpret_cast<ImgHeader*>(ws);
  p.hint = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + hdr_bytes(B));
  p.chist = reinterpret_cast<u32*>(reinterpret_cast<char*>(ws) + hdr_bytes(B) + hint_bytes(B));
  p.cand = reinterpret_cast<u64*>(reinterpret_cast<char*>(ws) + hdr_bytes(B) + hint_bytes(B) + hist_bytes(B));
  SSLAM_CHECK_CUDA(cudaMemsetAsync(p.hdr, 0, (size_t)B * sizeof(ImgHeader), stream));

  const int r = nms_radius;
  const bool vec_ok = (W % 4 == 0) && (W >= 64) && ((reinterpret_cast<uintptr_t>(sal) & 15) == 0);
  const int stream_strips = (W + 119) / 120;
  const size_t stream_stage = (size_t)(2 * r + 1) * W * 4;
  const bool stream_ok = r >= 1 && r <= 3 && vec_ok && stream_strips <= STREAM_MAX_WARPS && H >= 8 &&
                         2 * stream_stage <= 88 * 1024 && !g_decode_no_stream;
  bool hist_topk = false;
  if (stream_ok) {
    SSLAM_CHECK_CUDA(cudaMemsetAsync(p.chist, 0, (size_t)B * HIST_BINS * sizeof(u32), stream));
    // Residency beats ring depth (c2, 300 maps per launch): 7 stages x 2 CTAs per SM 0.26 ms, 4 x 3 CTAs 0.19,
    // 2 x 4 CTAs 0.177 — the scan is bound by instruction issue and latency, not by bytes in flight, so
    // the ring is kept at about 26 KB (two stages at W = 640) and the bands at 64 rows (a band's ~1500
    // local maxima then fit the 2048-key staging list: no overflow appends to the global list)
    int NS = (int)((26 * 1024) / stream_stage);
    if (NS < 2) NS = 2;
    if (NS > 16) NS = 16;
    if (g_decode_ns > 0) NS = g_decode_ns < 2 ? 2 : (g_decode_ns > 16 ? 16 : g_decode_ns);
    if ((size_t)NS * stream_stage > 88 * 1024) NS = (int)((88 * 1024) / stream_stage);
    const int band_target = g_decode_band > 0 ? g_decode_band : 64;
    const int nbands = (H + band_target - 1) / band_target;
    const int band_rows = (H + nbands - 1) / nbands;
    const size_t dyn = (size_t)NS * stream_stage;
    static DeviceOnce once_stream;
    if (once_stream.first_use()) {
#define SSLAM_SCAN_ATTR(R_, L_) SSLAM_CHECK_CUDA(cudaFuncSetAttribute(decode_scan_stream_kernel<R_, L_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 88 * 1024))
      SSLAM_SCAN_ATTR(1, false); SSLAM_SCAN_ATTR(2, false); SSLAM_SCAN_ATTR(3, false);
      SSLAM_SCAN_ATTR(1, true); SSLAM_SCAN_ATTR(2, true); SSLAM_SCAN_ATTR(3, true);
#undef SSLAM_SCAN_ATTR
    }
    const unsigned blocks = (unsigned)(B * nbands);
    const unsigned threads = 32u * (unsigned)(stream_strips + 1);
#define SSLAM_SCAN_GO(R_, L_) decode_scan_stream_kernel<R_, L_><<<blocks, threads, dyn, stream>>>(p, stream_strips, nbands, band_rows, NS)
    SSLAM_LAUNCH(KK_DECODE_SCAN, stream,
                 if (from_logits) { if (r == 1) SSLAM_SCAN_GO(1, true); else if (r == 2) SSLAM_SCAN_GO(2, true); else SSLAM_SCAN_GO(3, true); }
                 else { if (r == 1) SSLAM_SCAN_GO(1, false); else if (r == 2) SSLAM_SCAN_GO(2, false); else SSLAM_SCAN_GO(3, false); });
#undef SSLAM_SCAN_GO
    hist_topk = true;
  } else if (r >= 1 && r <= 3 && vec_ok) {
    const int nstrips = (W + 119) / 120;                       // 120 output columns per warp
    int nseg = (H + 63) / 64;                                  // ~64 rows per warp
    const int seg_rows = (H + nseg - 1) / nseg;
    nseg = (H + seg_rows - 1) / seg_rows;
    const unsigned blocks = (unsigned)(((nstrips * nseg + 7) / 8) * B);
    SSLAM_LAUNCH(KK_DECODE_SCAN, stream,
                 if (r == 1) decode_scan_vec_kernel<1><<<blocks, 256, 0, stream>>>(p, nstrips, nseg, seg_rows);
                 else if (r == 2) decode_scan_vec_kernel<2><<<blocks, 256, 0, stream>>>(p, nstrips, nseg, seg_rows);
                 else decode_scan_vec_kernel<3><<<blocks, 256, 0, stream>>>(p, nstrips, nseg, seg_rows));
  } else if (r >= 1 && r <= 3) {
    const int outw = 32 - 2 * r;
    const int nstrips = (W + outw - 1) / outw;
    int nseg = (H + 63) / 64;                                  // ~64 rows per warp
    const int seg_rows = (H + nseg - 1) / nseg;
    nseg = (H + seg_rows - 1) / seg_rows;
    const unsigned blocks = (unsigned)(((nstrips * nseg + 7) / 8) * B);
    SSLAM_LAUNCH(KK_DECODE_SCAN, stream,
                 if (r == 1) decode_scan_strip_kernel<1><<<blocks, 256, 0, stream>>>(p, nstrips, nseg, seg_rows);
                 else if (r == 2) decode_scan_strip_kernel<2><<<blocks, 256, 0, stream>>>(p, nstrips, nseg, seg_rows);
                 else decode_scan_strip_kernel<3><<<blocks, 256, 0, stream>>>(p, nstrips, nseg, seg_rows));
  } else {
    size_t scan_smem = (size_t)((TILE_H + 2 * r) * (TILE_W + 2 * r) + (TILE_H + 2 * r) * TILE_W) * 4;
    dim3 g1((W + TILE_W - 1) / TILE_W, (H + TILE_H - 1) / TILE_H, B);
    SSLAM_LAUNCH(KK_DECODE_SCAN, stream, decode_scan_kernel<<<g1, SCAN_THREADS, scan_smem, stream>>>(p));
  }

  const int kpad = next_pow2(K);
  size_t sel_smem = (size_t)kpad * 8 + sizeof(SelectScratch);
  static DeviceOnce once;
  if (once.first_use()) {
    SSLAM_CHECK_CUDA(cudaFuncSetAttribute(decode_topk_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SSLAM_CHECK_CUDA(cudaFuncSetAttribute(decode_resolve_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  if (hist_topk) {
    const int cap = 2 * kpad;                                  // K + boundary bin; pow2
    const size_t smem2 = (size_t)cap * 8;
    static DeviceOnce once2;
    if (once2.first_use())
      SSLAM_CHECK_CUDA(cudaFuncSetAttribute(decode_topk_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            200 * 1024));
    SSLAM_LAUNCH(KK_DECODE_TOPK, stream,
                 decode_topk_hist_kernel<<<B, TOPK2_THREADS, smem2, stream>>>(p, cap));
  } else {
    SSLAM_LAUNCH(KK_DECODE_TOPK, stream,
                 decode_topk_kernel<<<B, SEL_THREADS, sel_smem, stream>>>(p, kpad));
  }

  int chunks = (H * W + 16383) / 16384;
  if (chunks < 1) chunks = 1;
  SSLAM_LAUNCH(KK_DECODE_COUNT, stream,
               decode_count_kernel<<<B * chunks, 256, 0, stream>>>(p, chunks));
  SSLAM_LAUNCH(KK_DECODE_RESOLVE, stream,
               decode_resolve_kernel<<<B, RES_THREADS, sel_smem, stream>>>(p, kpad));
  return SSLAM_OK;
}

extern "C" int sslam_nms_f32(const float* sal, int B, int H, int W, int nms_radius, float* out,
                             void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(B >= 0 && H > 0 && W > 0, SSLAM_EINVAL, "nms: negative size");
  if (B == 0) return SSLAM_OK;
  SSLAM_REQUIRE(sal && out, SSLAM_EINVAL, "nms: null pointer");
  SSLAM_REQUIRE(nms_radius >= 0 && nms_radius <= MAX_R, SSLAM_EUNSUPPORTED,
                "nms: nms_radius %d outside 0..%d", nms_radius, MAX_R);
  if (nms_radius == 0) {                                     // identity (:211-212)
    SSLAM_CHECK_CUDA(cudaMemcpyAsync(out, sal, (size_t)B * H * W * 4, cudaMemcpyDeviceToDevice,
                                     stream));
    return SSLAM_OK;
  }
  const int r = nms_radius;
  size_t smem = (size_t)((TILE_H + 2 * r) * (TILE_W + 2 * r) + (TILE_H + 2 * r) * TILE_W) * 4;
  dim3 g((W + TILE_W - 1) / TILE_W, (H + TILE_H - 1) / TILE_H, B);
  SSLAM_LAUNCH(KK_NMS, stream,
               nms_kernel<<<g, SCAN_THREADS, smem, stream>>>(sal, H, W, r, out));
  return SSLAM_OK;
}

// Debug aid for tests / tools (include/sslam_b200_debug.h): 0 forces the register-prefetching scan
// kernels + radix-select top-k instead of the shared-memory streaming scan + histogram top-k.
extern "C" void sslam_debug_decode_stream(int on) { sslam::g_decode_no_stream = (on == 0); }
extern "C" void sslam_debug_decode_tune(int stages, int band_rows) { sslam::g_decode_ns = stages; sslam::g_decode_band = band_rows; }
