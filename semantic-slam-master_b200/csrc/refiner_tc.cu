// DescriptorRefiner forward on tensor cores.
//
// Replaces the body of DescriptorRefiner.forward / ResidualBlock.forward
// (models/descriptor_refiner.py:73-86, 108-126): Linear+ReLU, [LN, Linear+ReLU, LN, Linear,
// +identity, ReLU] x blocks, Linear, L2 normalise — with fp32-level accuracy at the full 16-bit
// tensor-core rate:
//
//   * numbers are carried as fp16 PAIRS, x ~= hi + lo * 2^-11 (hi = fp16(x), lo = fp16((x-hi) * 2^11):
//     22 significant bits), and every Linear is three tcgen05 kind::f16 MMAs
//         x.w ~= hi.hi'  +  2^-11 * (hi.lo' + lo.hi')
//     with fp32 accumulation in TMEM; the cross terms live in their own accumulator and are folded
//     in by the epilogue with one FMA (the tensor core truncates when it adds into an accumulator,
//     so small terms must not be added to a large one inside it).  The dropped lo.lo' term is 2^-22.
//   * activations travel between layers as such pairs (two fp16 arrays = 4 bytes per element, the
//     same HBM traffic as fp32) and are loaded by TMA straight into 128B-swizzled operand tiles —
//     no conversion stage, no fp32 activations in HBM; weights are split once by
//     sslam_refiner_pack_weights.  Requires |activation| < 65504.
//   * the GEMM epilogue (thread = row, tcgen05.ld, 16 columns at a time) fuses bias / folded
//     LayerNorm, residual add, ReLU and the row statistics, packs the next layer's pair (or fp32 for
//     the last layer) into a dense shared-memory box and hands it to a TMA store, which also clips
//     ragged rows / columns; no per-lane global stores.
//   * LayerNorm never runs as a kernel.  For y = LN(h).W^T + c with LN(h) = (h - mu) * rho * g + b,
//         y[r,n] = rho_r * ( (h.W'^T)[r,n] - mu_r * s1[n] ) + c0[n],
//     W' = W * g (column scale), s1[n] = sum_k W'[n,k], c0[n] = sum_k b[k] W[n,k] + c[n]  (all folded
//     once by sslam_refiner_pack_weights).  So the GEMM multiplies the *un-normalised* activations
//     and its epilogue applies the two per-row scalars; those (mu_r, rho_r) are produced for free by
//     the epilogue of the GEMM that wrote h (row sums of v and v^2 while storing).
//
// Two kernels.  gemm_pair_kernel (further down) is the production path for K <= 384, N <= 512:
// weight-stationary CTA pairs (cta_group::2, M = 256), the pairs of the column tiles of one strip
// set forming a cluster that multicasts the activation tiles.  gemm_f16x3_kernel (directly below)
// is the general single-CTA fallback: persistent, one CTA per SM, each CTA walks 128-row strips
// (blockIdx, +gridDim, ...) and, inside a strip, the N/128 column tiles of the layer, streaming
// both operands.
#include "tc_common.cuh"

namespace sslam {

using namespace tc;

long long* g_gemm_dbg = nullptr;    // set by sslam_debug_gemm_stalls (tools only, not part of the ABI)

namespace {

constexpr int BM = 128, BN = 128, BK = 64;               // BK fp16 = 128 bytes of K
constexpr int BLOCK_BYTES = BM * 128;
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 4 * BLOCK_BYTES;             // A_hi, A_lo, B_hi, B_lo
constexpr int EPI_WARPS = 8;                             // two warps per TMEM lane quarter
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;                           // 2 x (128 main + 128 cross)
constexpr int UNIT = 16;                                 // columns per epilogue step
constexpr int STG_BYTES = 32 * UNIT * 4;                 // per warp: [32][16] fp16 hi | lo, or [32][16] fp32
constexpr int SMEM_OPERANDS = STAGES * STAGE_BYTES;
constexpr int SMEM_TRANSP = EPI_WARPS * STG_BYTES;
constexpr int SMEM_STATS = 2 * 4 * 32 * 2 * 2 * 4;         // [parity][quarter][row][half][sum, sumsq]
constexpr int MAX_N = 1024;
constexpr int SMEM_VECS = 2 * MAX_N * 4;                 // bias / c0 and s1 of the layer, staged once per CTA
constexpr int SMEM_BARS = (2 * STAGES + 4) * 8 + 16;
constexpr int SMEM_TOTAL = SMEM_OPERANDS + SMEM_TRANSP + SMEM_STATS + SMEM_VECS + SMEM_BARS + 1024;

struct GemmParams {
  int rows, N, K;
  const float* bias;        // [N]
  const __half* res_hi;     // residual pair [rows, N] or null
  const __half* res_lo;
  int relu;
  float* out_f32;           // [rows, N] or null
  __half* out_hi;           // pair [rows, N] or null
  __half* out_lo;
  // folded LayerNorm on the A operand (null = plain bias): partial row sums / sums of squares of A
  // ([a_parts][stat_stride], written by the GEMM that produced A), per-column s1; `bias` then holds c0
  const float* a_sum;
  const float* a_sq;
  int a_parts;
  const float* s1;
  // partial row sums of the OUTPUT (after residual + ReLU) for the next folded LayerNorm, or null:
  // part = column tile of the writer (pair kernel) or 0 (single-CTA kernel)
  float* out_sum;
  float* out_sq;
  size_t stat_stride;
  long long* dbg;           // optional per-CTA stall counters (debug aid), or null
};

// 32 contiguous, 32-byte aligned bytes through the read-only path as one 256-bit load
__device__ __forceinline__ void ldg_256(const void* ptr, uint4 (&d)[2]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(d[0].x), "=r"(d[0].y), "=r"(d[0].z), "=r"(d[0].w), "=r"(d[1].x), "=r"(d[1].y), "=r"(d[1].z),
                 "=r"(d[1].w)
               : "l"(ptr));
}

// mean / rstd of row `grow` of the A operand from the producer's partial sums (fixed summation order)
__device__ __forceinline__ void row_layernorm_scalars(const GemmParams& p, int grow, float& mean, float& rstd) {
  float sm = 0.f, sq = 0.f;
  for (int i = 0; i < p.a_parts; ++i) {
    sm += __ldg(p.a_sum + (size_t)i * p.stat_stride + grow);
    sq += __ldg(p.a_sq + (size_t)i * p.stat_stride + grow);
  }
  mean = sm / (float)p.K;
  const float var = fmaxf(sq / (float)p.K - mean * mean, 0.f);
  rstd = 1.0f / sqrtf(var + 1e-5f);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
                  const __grid_constant__ CUtensorMap tmO_f32, GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* operands = smem;
  unsigned char* staging = smem + SMEM_OPERANDS;
  float* stats = reinterpret_cast<float*>(smem + SMEM_OPERANDS + SMEM_TRANSP);
  float* sbias = reinterpret_cast<float*>(smem + SMEM_OPERANDS + SMEM_TRANSP + SMEM_STATS);
  float* ss1 = sbias + MAX_N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_OPERANDS + SMEM_TRANSP + SMEM_STATS + SMEM_VECS);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nstrips = (p.rows + BM - 1) / BM;       // persistent: strips blockIdx.x, +gridDim.x, ...
  const int ntile = (p.N + BN - 1) / BN;
  const int nkb = (p.K + BK - 1) / BK;
  // Every CTA multiplies by the same weight matrix; walking its tiles in lock-step would make all
  // 148 SMs hit the same few L2 lines at once.  Each CTA therefore starts at its own column tile
  // and its own k-block (the k order only changes the fp32 summation order).
  const int ct_rot = blockIdx.x % ntile;
  const int kb_rot = (blockIdx.x / ntile) % nkb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  for (int i = threadIdx.x; i < p.N; i += NUM_THREADS) {        // per-column vectors -> smem
    sbias[i] = __ldg(p.bias + i);
    ss1[i] = p.a_sum ? __ldg(p.s1 + i) : 0.f;
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
      for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
        const int row0 = strip * BM;
        for (int ct = 0; ct < ntile; ++ct) {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* st = operands + stage * STAGE_BYTES;
            mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
            int kr = kb + kb_rot; if (kr >= nkb) kr -= nkb;
            int cr = ct + ct_rot; if (cr >= ntile) cr -= ntile;
            const int kc = kr * BK;
            tma_load_2d(st, &tmA_hi, &full[stage], kc, row0);
            tma_load_2d(st + BLOCK_BYTES, &tmA_lo, &full[stage], kc, row0);
            tma_load_2d(st + 2 * BLOCK_BYTES, &tmB_hi, &full[stage], kc, cr * BN);
            tma_load_2d(st + 3 * BLOCK_BYTES, &tmB_lo, &full[stage], kc, cr * BN);
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
      const int my_tiles = ((nstrips - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * ntile;
      for (int tc = 0; tc < my_tiles; ++tc) {           // tc: running tile count of this CTA
        const int acc = tc & 1;
        mbar_wait(&tempty[acc], ((tc >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 2 * BN;
        const uint32_t tmem_s = tmem_d + BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);               // TMA bytes have landed
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(operands + stage * STAGE_BYTES);
          const uint64_t a_hi = make_smem_desc_sw128(sa);
          const uint64_t a_lo = make_smem_desc_sw128(sa + BLOCK_BYTES);
          const uint64_t b_hi = make_smem_desc_sw128(sa + 2 * BLOCK_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            const uint32_t first = (kb | k) ? 1u : 0u;
            // B_hi and B_lo tiles are adjacent: A_hi x [B_hi ; B_lo] is one N=256 instruction
            // (hi.hi -> columns [0,128), hi.lo -> [128,256)); lo.hi then accumulates into the latter
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
    // ---- epilogue warps 2..9.  Warps w and w+4 share a TMEM lane quarter and each take half of
    // the tile's 128 columns, 16 at a time: thread = row, all arithmetic in registers, result rows
    // packed into a dense [32][16] box in shared memory and written by TMA.
    const int q = warp & 3;
    const int ew = warp - 2;
    const int half = ew >> 2;                           // which 64 columns of the tile
    unsigned char* stg = staging + ew * STG_BYTES;
    if (lane == 0) { prefetch_tensormap(&tmO_hi); prefetch_tensormap(&tmO_lo); prefetch_tensormap(&tmO_f32); }
    int tc = 0, sidx = 0;
    for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x, ++sidx) {
      const int wrow0 = strip * BM + q * 32;            // first global row of this warp
      const int grow = wrow0 + lane;                    // this thread's row
      const bool row_ok = grow < p.rows;
      float am = 0.f, ar = 0.f;                         // folded-LN scalars of this row
      if (p.a_sum && row_ok) row_layernorm_scalars(p, grow, am, ar);
      float rsum = 0.f, rsq = 0.f;
      // residual of one unit = this row's 16 values of each half of the pair; requested one unit
      // ahead so that its HBM latency overlaps the arithmetic of the current unit
      auto unit_col = [&](int ct_, int un_) {
        int c = ct_ + ct_rot; if (c >= ntile) c -= ntile;
        return c * BN + half * 64 + un_ * UNIT;
      };
      auto load_res = [&](int gc, uint4 (&h)[2], uint4 (&l)[2]) {
        h[0] = h[1] = l[0] = l[1] = make_uint4(0u, 0u, 0u, 0u);
        if (p.res_hi && row_ok && gc < p.N) {
          const size_t o = (size_t)grow * p.N + gc;
          h[0] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o));
          l[0] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + o));
          if (gc + 8 < p.N) {
            h[1] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o + 8));
            l[1] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + o + 8));
          }
        }
      };
      uint4 nh[2], nl[2];
      load_res(unit_col(0, 0), nh, nl);
      for (int ct = 0; ct < ntile; ++ct, ++tc) {
        const int acc = tc & 1;
        mbar_wait(&tfull[acc], (tc >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll 1
        for (int un = 0; un < 64 / UNIT; ++un) {
          const int col0 = half * 64 + un * UNIT;       // first column of this unit inside the tile
          const int gc0 = unit_col(ct, un);             // ... and in the output (warp-uniform)
          uint4 rh[2] = {nh[0], nh[1]}, rl[2] = {nl[0], nl[1]};
          {
            int nct = ct, nun = un + 1;
            if (nun == 64 / UNIT) { nun = 0; ++nct; }
            if (nct < ntile) load_res(unit_col(nct, nun), nh, nl);
          }
          if (gc0 >= p.N) continue;
          uint32_t r[UNIT], rs[UNIT];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN + col0;
          tmem_ld_32x16(taddr, r);
          tmem_ld_32x16(taddr + BN, rs);
          tmem_ld_wait();
          float v[UNIT];
#pragma unroll
          for (int j = 0; j < UNIT; j += 4) {
            const bool cok = gc0 + j < p.N;             // N % 4 == 0
            const float4 b4 = *reinterpret_cast<const float4*>(sbias + gc0 + j);
            const float4 s4 = *reinterpret_cast<const float4*>(ss1 + gc0 + j);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
            const float ss[4] = {s4.x, s4.y, s4.z, s4.w};
            const __half* h8 = reinterpret_cast<const __half*>(&rh[j >> 3]) + (j & 7);
            const __half* l8 = reinterpret_cast<const __half*>(&rl[j >> 3]) + (j & 7);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = __fmaf_rn(__uint_as_float(rs[j + e]), F16_LO_INV, __uint_as_float(r[j + e]));
              if (p.a_sum) x = __fmaf_rn(ar, __fmaf_rn(-am, ss[e], x), bb[e]);    // rho*(acc - mu*s1) + c0
              else x = __fadd_rn(x, bb[e]);
              if (p.res_hi) x = __fadd_rn(x, join_f16(h8[e], l8[e]));
              if (p.relu) x = fmaxf(x, 0.f);
              if (!cok) x = 0.f;
              v[j + e] = x;
              rsum += x;
              rsq = __fmaf_rn(x, x, rsq);
            }
          }
          // the previous TMA store of this warp must have finished reading the staging box
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          if (p.out_f32) {
            float4* dst = reinterpret_cast<float4*>(stg + lane * (UNIT * 4));
#pragma unroll
            for (int j = 0; j < UNIT; j += 4) dst[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            __half hh[UNIT], ll[UNIT];
#pragma unroll
            for (int j = 0; j < UNIT; ++j) split_f16(v[j], hh[j], ll[j]);
            uint4* dh = reinterpret_cast<uint4*>(stg + lane * (UNIT * 2));
            uint4* dl = reinterpret_cast<uint4*>(stg + 32 * UNIT * 2 + lane * (UNIT * 2));
            dh[0] = *reinterpret_cast<uint4*>(&hh[0]); dh[1] = *reinterpret_cast<uint4*>(&hh[8]);
            dl[0] = *reinterpret_cast<uint4*>(&ll[0]); dl[1] = *reinterpret_cast<uint4*>(&ll[8]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (p.out_f32) {
              tma_store_2d(&tmO_f32, stg, gc0, wrow0);
            } else {
              tma_store_2d(&tmO_hi, stg, gc0, wrow0);
              tma_store_2d(&tmO_lo, stg + 32 * UNIT * 2, gc0, wrow0);
            }
            tma_store_commit();
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
      if (p.out_sum) {
        // row sums of what this strip stored: combine the two column halves of each row
        float* sc = stats + (sidx & 1) * (4 * 32 * 2 * 2);
        float* e = sc + ((q * 32 + lane) * 2 + half) * 2;
        e[0] = rsum; e[1] = rsq;
        named_bar_sync(1, 32 * EPI_WARPS);
        if (half == 0 && row_ok) {
          const float* f = sc + (q * 32 + lane) * 4;
          p.out_sum[grow] = f[0] + f[2];
          p.out_sq[grow] = f[1] + f[3];
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();             // all output boxes are in global memory
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2), weight-stationary.  Used whenever K <= 384.
//
// The single-CTA kernel above streams both operands of every 128x128 tile through L2 and shared
// memory (2 x 196 KB per tile at K = 384): at one CTA per SM that saturates the L2 slices and the
// 128 B/cycle shared-memory port long before the tensor pipe.  Here two CTAs on the SMs of one TPC
// share each MMA (M = 256: each CTA owns a 128-row strip and its accumulators; each holds HALF of the
// weight tile), and a pair keeps its 128 weight columns resident for the whole launch:
//   * grid = groups x ceil(N/128) pairs; pair (g, ct) keeps W[ct*128 .. +128, :] (64 rows per CTA,
//     hi+lo, K/64 x 16 KB <= 96 KB) in shared memory and walks the 256-row strip pairs g, g+groups, ...
//   * only activations stream: 32 KB per k-block and CTA, 3 stages; per tile a CTA pulls 196 KB
//     instead of 393 KB through L2 and its MMAs read 6 KB instead of 8 KB of operands each;
//   * the three pairs that own the column tiles of the same strips run in step, so a strip is
//     fetched from DRAM once and served from L2 to the other two;
//   * row statistics for the next folded LayerNorm are written as one partial (sum, sum of squares)
//     per column tile and added up in a fixed order by the consumer (no atomics: results do not
//     depend on scheduling).
// Barriers: TMA of both CTAs signals the LEADER's `full` barriers (rank 0 issues every MMA);
// tcgen05.commit multicasts `empty` / `tfull` to both CTAs; both epilogues arrive on the leader's
// `tempty`.
constexpr int P_STAGES = 3;
constexpr int P_MAX_KB = 6;                               // K <= 384
constexpr int B_HALF = 64 * 128;                          // 64 weight rows x 128 bytes of K
constexpr int P_STAGE_BYTES = 2 * BLOCK_BYTES;            // A_hi, A_lo
constexpr int P_SMEM_B = P_MAX_KB * 2 * B_HALF;
constexpr int P_SMEM_A = P_STAGES * P_STAGE_BYTES;
constexpr int P_SMEM_VECS = 2 * BN * 4;
constexpr int P_SMEM_BARS = (2 * P_STAGES + 5) * 8 + 16;
constexpr int P_SMEM_TOTAL = P_SMEM_B + P_SMEM_A + SMEM_TRANSP + SMEM_STATS + P_SMEM_VECS + P_SMEM_BARS + 1024;

template <bool LN, bool RES, bool RELU, bool OUTF32, bool STATS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
                 const __grid_constant__ CUtensorMap tmO_f32, GemmParams p, int groups) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* bres = smem;
  unsigned char* a_stages = smem + P_SMEM_B;
  unsigned char* staging = a_stages + P_SMEM_A;
  float* stats = reinterpret_cast<float*>(staging + SMEM_TRANSP);
  float* sbias = reinterpret_cast<float*>(staging + SMEM_TRANSP + SMEM_STATS);
  float* ss1 = sbias + BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + SMEM_TRANSP + SMEM_STATS + P_SMEM_VECS);
  uint64_t* full = bars;
  uint64_t* empty = bars + P_STAGES;
  uint64_t* bfull = bars + 2 * P_STAGES;
  uint64_t* tfull = bfull + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // cluster = the ntile CTA pairs that own the column tiles of the same strips; cluster rank = 2*ct + rank
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;                       // rank inside the CTA pair; 0 = MMA leader
  const uint32_t leader = crank & ~1u;                    // cluster rank of this pair's leader
  const int ntile = (p.N + BN - 1) / BN;                  // = multicast width
  const int ct = (int)(crank >> 1), g = blockIdx.x / (2 * ntile);
  const int col_base = ct * BN;
  const int nsp = (p.rows + 2 * BM - 1) / (2 * BM);      // 256-row strip pairs
  const int nkb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    // a stage is free again when the MMAs of ALL pairs of the cluster have read it (it is refilled
    // by a multicast that writes every CTA of the same rank)
    // `full` (used in the pair leader): the leader's expect_tx arrival plus one arrival of the peer's
    // producer.  The peer's arrival carries no bytes; it is there so that EVERY producer of the
    // cluster takes part in every use of every stage.  A producer that only waits on `empty` when it
    // is not its turn to fetch could otherwise be overtaken by two phases of that barrier, see the
    // old phase as incomplete again (parity aliasing) and never issue its own stage's loads:
    // a rare deadlock (about one launch in 300 when the epilogue is the bottleneck).
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], (uint32_t)ntile); }
    mbar_init(bfull, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
  for (int i = threadIdx.x; i < BN; i += NUM_THREADS) {          // this pair's 128 columns of the per-column vectors
    const int c = col_base + i;
    sbias[i] = c < p.N ? __ldg(p.bias + c) : 0.f;
    ss1[i] = (LN && c < p.N) ? __ldg(p.s1 + c) : 0.f;
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {                                            // ---- TMA producer (both CTAs)
      prefetch_tensormap(&tmA_hi); prefetch_tensormap(&tmA_lo);
      prefetch_tensormap(&tmB_hi); prefetch_tensormap(&tmB_lo);
      // resident weights: this CTA's 64 of the pair's 128 rows
      const uint32_t bfull_leader = mapa_u32(smem_u32(bfull), leader);
      if (rank == 0) mbar_arrive_expect_tx(bfull, 2u * (uint32_t)nkb * 2u * B_HALF);
      for (int kb = 0; kb < nkb; ++kb) {
        tma_load_2d_pair(bres + kb * 2 * B_HALF, &tmB_hi, bfull_leader, kb * BK, col_base + (int)rank * 64);
        tma_load_2d_pair(bres + kb * 2 * B_HALF + B_HALF, &tmB_lo, bfull_leader, kb * BK, col_base + (int)rank * 64);
      }
      // activations: the ntile CTAs of the same rank need the same 128-row tile of every k-block;
      // they take turns fetching it and multicast it to all of them (L2 -> SM traffic / ntile)
      uint16_t mc_mask = 0;
      for (int j = 0; j < ntile; ++j) mc_mask |= (uint16_t)(1u << (2 * j + (int)rank));
      int stage = 0; uint32_t phase = 0;
      int turn = 0;                                               // running k-block count modulo ntile
      for (int sp = g; sp < nsp; sp += groups) {
        const int row0 = sp * 2 * BM + (int)rank * BM;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = a_stages + stage * P_STAGE_BYTES;
          const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), leader);
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2u * P_STAGE_BYTES);
          else mbar_arrive_cluster(full_leader);
          if (turn == ct) {
            if (ntile == 1) {
              tma_load_2d_pair(st, &tmA_hi, full_leader, kb * BK, row0);
              tma_load_2d_pair(st + BLOCK_BYTES, &tmA_lo, full_leader, kb * BK, row0);
            } else {
              tma_load_2d_pair_mc(st, &tmA_hi, full_leader, kb * BK, row0, mc_mask);
              tma_load_2d_pair_mc(st + BLOCK_BYTES, &tmA_lo, full_leader, kb * BK, row0, mc_mask);
            }
            // the activations come from DRAM (a layer's input is far larger than L2): start the
            // fetch of the same k-block of the NEXT strip pair now, one whole tile ahead
            if (sp + groups < nsp) {
              tma_prefetch_l2_2d(&tmA_hi, kb * BK, row0 + groups * 2 * BM);
              tma_prefetch_l2_2d(&tmA_lo, kb * BK, row0 + groups * 2 * BM);
            }
          }
          if (++turn == ntile) turn = 0;
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      // Producer tail: the last `empty` arrivals of every stage are tcgen05.commit multicasts from
      // the MMA threads of ALL pairs of the cluster, delivered asynchronously into this CTA's shared
      // memory.  Nobody would otherwise wait for them, and the cluster barrier at the end does not
      // order them: if this CTA exited first they would land in the shared memory of whatever CTA
      // runs here next (observed as sporadic launch failures of back-to-back GEMMs).  Drain them.
      for (int i = 0; i < P_STAGES; ++i) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {                               // ---- MMA issuer (leader CTA only)
      const uint32_t idesc = make_instr_desc(FMT_F16, 2 * BM, BN);
      const uint16_t pair_mask = (uint16_t)(3u << leader);        // both CTAs of this pair
      const uint16_t all_mask = (uint16_t)((1u << (2 * ntile)) - 1u);
      int stage = 0; uint32_t phase = 0;
      const bool dbg_on = p.dbg != nullptr;
      long long w_full = 0, w_tempty = 0, tq = 0, t_begin = clock64();
      mbar_wait(bfull, 0);
      tcgen05_fence_after();
      const long long w_b = clock64() - t_begin;
      int tc = 0;
      for (int sp = g; sp < nsp; sp += groups, ++tc) {
        const int acc = tc & 1;
        if (dbg_on) tq = clock64();
        mbar_wait(&tempty[acc], (((uint32_t)tc >> 1) & 1u) ^ 1u);
        if (dbg_on) w_tempty += clock64() - tq;
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 2 * BN;
        const uint32_t tmem_s = tmem_d + BN;
        for (int kb = 0; kb < nkb; ++kb) {
          if (dbg_on) tq = clock64();
          mbar_wait(&full[stage], phase);
          if (dbg_on) w_full += clock64() - tq;
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(a_stages + stage * P_STAGE_BYTES);
          const uint32_t sb = smem_u32(bres + kb * 2 * B_HALF);
          const uint64_t a_hi = make_smem_desc_sw128(sa);
          const uint64_t a_lo = make_smem_desc_sw128(sa + BLOCK_BYTES);
          const uint64_t b_hi = make_smem_desc_sw128(sb);
          const uint64_t b_lo = make_smem_desc_sw128(sb + B_HALF);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            const uint32_t first = (kb | k) ? 1u : 0u;
            umma_ss_pair(tmem_s, a_lo + adv, b_hi + adv, idesc, first);
            umma_ss_pair(tmem_s, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_ss_pair(tmem_d, a_hi + adv, b_hi + adv, idesc, first);
          }
          tcgen05_commit_pair(&empty[stage], all_mask);            // every CTA of the cluster: this pair is done with the stage
          if (kb == nkb - 1) tcgen05_commit_pair(&tfull[acc], pair_mask);
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (p.dbg) {
        long long* d = p.dbg + 8 * blockIdx.x;
        d[0] = clock64() - t_begin; d[1] = w_full; d[2] = w_tempty; d[3] = w_b;
      }
    }
  } else {
    // ---- epilogue warps 2..9 of both CTAs, each CTA on its own 128 rows.  The layer's options are
    // template parameters: the per-element code has no branches on kernel arguments (the epilogue
    // is instruction-issue bound — 8 warps share 4 schedulers with the two producer warps).
    const int q = warp & 3;
    const int ew = warp - 2;
    const int half = ew >> 2;
    unsigned char* stg = staging + ew * STG_BYTES;
    if (lane == 0) { prefetch_tensormap(&tmO_hi); prefetch_tensormap(&tmO_lo); prefetch_tensormap(&tmO_f32); }
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty[0]), leader);
    const uint32_t tempty_leader1 = mapa_u32(smem_u32(&tempty[1]), leader);
    const bool full_cols = col_base + BN <= p.N;          // no ragged columns in this pair's tile
    int tc = 0;
    for (int sp = g; sp < nsp; sp += groups, ++tc) {
      const int wrow0 = sp * 2 * BM + (int)rank * BM + q * 32;
      const int grow = wrow0 + lane;
      const bool row_ok = grow < p.rows;
      float am = 0.f, ar = 0.f, nam = 0.f;
      if (LN && row_ok) { row_layernorm_scalars(p, grow, am, ar); nam = -am; }
      float rsum = 0.f, rsq = 0.f;
      // residual of one unit = this row's 16 values of each half of the pair, requested one unit ahead
      const __half* res_h = RES ? p.res_hi + (size_t)(row_ok ? grow : 0) * p.N : nullptr;
      const __half* res_l = RES ? p.res_lo + (size_t)(row_ok ? grow : 0) * p.N : nullptr;
      auto load_res = [&](int gc, uint4 (&h)[2], uint4 (&l)[2]) {
        if (gc + UNIT <= p.N) {                         // one 32-byte sector per array: 256-bit loads
          ldg_256(res_h + gc, h);
          ldg_256(res_l + gc, l);
        } else {
          h[0] = h[1] = l[0] = l[1] = make_uint4(0u, 0u, 0u, 0u);
          if (gc < p.N) {
            h[0] = __ldg(reinterpret_cast<const uint4*>(res_h + gc));
            l[0] = __ldg(reinterpret_cast<const uint4*>(res_l + gc));
          }
        }
      };
      uint4 nh[2], nl[2];
      if (RES) load_res(col_base + half * 64, nh, nl); // in flight while the accumulator completes
      {
        // pull what the NEXT strip pair of this warp will read row-wise (residual pair, LayerNorm
        // partial sums) into L2 now, so that those loads are L2 hits when their turn comes
        const int nrow = grow + groups * 2 * BM;
        if (nrow < p.rows) {
          if (RES) {
            const size_t o = (size_t)nrow * p.N + col_base + half * 64;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res_hi + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res_lo + o));
          }
          if (LN && half == 0 && (lane & 7) == 0) {
            for (int i = 0; i < p.a_parts; ++i) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.a_sum + (size_t)i * p.stat_stride + nrow));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.a_sq + (size_t)i * p.stat_stride + nrow));
            }
          }
        }
      }
      const int acc = tc & 1;
      mbar_wait(&tfull[acc], ((uint32_t)tc >> 1) & 1u);
      tcgen05_fence_after();
      uint32_t r[UNIT], rs[UNIT];
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN + half * 64;
        tmem_ld_32x16(taddr, r);
        tmem_ld_32x16(taddr + BN, rs);
      }
#pragma unroll
      for (int un = 0; un < 64 / UNIT; ++un) {
        const int col0 = half * 64 + un * UNIT;         // first column of this unit inside the tile
        const int gc0 = col_base + col0;                // ... and in the output (warp-uniform)
        // (one unit ahead only: a deeper prefetch makes ptxas spill, and registers that are the
        // target of an in-flight tcgen05.ld should not be spilled before tcgen05.wait::ld; the
        // Makefile builds with --warn-on-spills.  Requesting the whole slab up front and giving up
        // the accumulator pipelining instead measured 5 % slower.)
        uint4 rh[2], rl[2];
        if (RES) {
          rh[0] = nh[0]; rh[1] = nh[1]; rl[0] = nl[0]; rl[1] = nl[1];
          if (un + 1 < 64 / UNIT) load_res(gc0 + UNIT, nh, nl);
        }
        float v[UNIT];
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < UNIT; ++j) {
          v[j] = __fmaf_rn(__uint_as_float(rs[j]), F16_LO_INV, __uint_as_float(r[j]));   // fold the cross terms
        }
        if (un + 1 < 64 / UNIT) {                       // next unit's accumulators: in flight during the math
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN + col0 + UNIT;
          tmem_ld_32x16(taddr, r);
          tmem_ld_32x16(taddr + BN, rs);
        }
        if (gc0 >= p.N) continue;
#pragma unroll
        for (int j = 0; j < UNIT; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sbias + col0 + j);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
          float ss[4] = {0.f, 0.f, 0.f, 0.f};
          if (LN) {
            const float4 s4 = *reinterpret_cast<const float4*>(ss1 + col0 + j);
            ss[0] = s4.x; ss[1] = s4.y; ss[2] = s4.z; ss[3] = s4.w;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = v[j + e];
            if (LN) x = __fmaf_rn(ar, __fmaf_rn(nam, ss[e], x), bb[e]);         // rho*(acc - mu*s1) + c0
            else x = __fadd_rn(x, bb[e]);
            if (RES)
              x = __fadd_rn(__fmaf_rn(__half2float(reinterpret_cast<const __half*>(rl)[j + e]), F16_LO_INV, x),
                            __half2float(reinterpret_cast<const __half*>(rh)[j + e]));
            if (RELU) x = fmaxf(x, 0.f);
            v[j + e] = x;
          }
        }
        if (!full_cols) {                               // ragged last column tile (warp-uniform)
#pragma unroll
          for (int j = 0; j < UNIT; ++j)
            if (gc0 + j >= p.N) v[j] = 0.f;
        }
        if (STATS) {
#pragma unroll
          for (int j = 0; j < UNIT; ++j) { rsum += v[j]; rsq = __fmaf_rn(v[j], v[j], rsq); }
        }
        if (lane == 0) tma_store_wait_read<0>();        // previous box of this warp has left smem
        __syncwarp();
        if (OUTF32) {
          float4* dst = reinterpret_cast<float4*>(stg + lane * (UNIT * 4));
#pragma unroll
          for (int j = 0; j < UNIT; j += 4) dst[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          __half2 hh[UNIT / 2], ll[UNIT / 2];
#pragma unroll
          for (int j = 0; j < UNIT; j += 2) {
            const __half2 h2 = __floats2half2_rn(v[j], v[j + 1]);
            const float2 hf = __half22float2(h2);
            hh[j >> 1] = h2;
            ll[j >> 1] = __floats2half2_rn(__fmul_rn(__fsub_rn(v[j], hf.x), F16_LO_SCALE),
                                           __fmul_rn(__fsub_rn(v[j + 1], hf.y), F16_LO_SCALE));
          }
          uint4* dh = reinterpret_cast<uint4*>(stg + lane * (UNIT * 2));
          uint4* dl = reinterpret_cast<uint4*>(stg + 32 * UNIT * 2 + lane * (UNIT * 2));
          dh[0] = *reinterpret_cast<uint4*>(&hh[0]); dh[1] = *reinterpret_cast<uint4*>(&hh[4]);
          dl[0] = *reinterpret_cast<uint4*>(&ll[0]); dl[1] = *reinterpret_cast<uint4*>(&ll[4]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (OUTF32) {
            tma_store_2d(&tmO_f32, stg, gc0, wrow0);
          } else {
            tma_store_2d(&tmO_hi, stg, gc0, wrow0);
            tma_store_2d(&tmO_lo, stg + 32 * UNIT * 2, gc0, wrow0);
          }
          tma_store_commit();
        }
      }
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
      if (STATS) {
        // partial row sums of this column tile: combine the two 64-column halves of each row
        float* sc = stats + (tc & 1) * (4 * 32 * 2 * 2);
        float* e = sc + ((q * 32 + lane) * 2 + half) * 2;
        e[0] = rsum; e[1] = rsq;
        named_bar_sync(1, 32 * EPI_WARPS);
        if (half == 0 && row_ok) {
          const float* f = sc + (q * 32 + lane) * 4;
          p.out_sum[(size_t)ct * p.stat_stride + grow] = f[0] + f[2];
          p.out_sq[(size_t)ct * p.stat_stride + grow] = f[1] + f[3];
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

// fp32 [n] -> fp16 pair
__global__ void split_kernel(const float4* __restrict__ src, uint2* __restrict__ hi, uint2* __restrict__ lo,
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

// Weight folding for a Linear fed by LayerNorm(gamma, beta): one warp per output row n.
//   W'[n,k] = W[n,k] * gamma[k]  -> fp16 pair;  s1[n] = sum_k W'[n,k] (of the pair actually used);
//   c0[n] = sum_k beta[k] * W[n,k] + bias[n]
__global__ void fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ bias, int N, int K,
                               __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ s1,
                               float* __restrict__ c0) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float a = 0.f, c = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[(size_t)n * K + k];
    const float wp = __fmul_rn(w, gamma[k]);
    __half h, l;
    split_f16(wp, h, l);
    hi[(size_t)n * K + k] = h;
    lo[(size_t)n * K + k] = l;
    a += join_f16(h, l);
    c = __fmaf_rn(beta[k], w, c);
  }
  a = warp_reduce_sum(a);
  c = warp_reduce_sum(c);
  if (lane == 0) { s1[n] = a; c0[n] = c + bias[n]; }
}

struct Pair { __half* hi; __half* lo; };

int launch_split(const float* src, Pair dst, size_t n, cudaStream_t stream) {
  SSLAM_LAUNCH(KK_SPLIT, stream,
               split_kernel<<<num_sms() * 8, 256, 0, stream>>>(reinterpret_cast<const float4*>(src),
                                                               reinterpret_cast<uint2*>(dst.hi),
                                                               reinterpret_cast<uint2*>(dst.lo), n / 4));
  return SSLAM_OK;
}

struct RowStats { float* sum; float* sq; int parts; };   // [parts][stat_stride] partial row sums

size_t stat_stride_of(int rows) { return align_up((size_t)rows, 64); }
int stat_parts_of(int N) { return (N + BN - 1) / BN; }

int launch_gemm(Pair a, Pair w, int rows, int N, int K, const float* bias, RowStats a_ln, const float* s1,
                Pair residual, int relu, float* out_f32, Pair out, RowStats* out_stats, cudaStream_t stream) {
  // weight-stationary CTA-pair kernel: weights must fit (K <= 384) and the ntile pairs form one cluster
  const bool pair = (K + BK - 1) / BK <= P_MAX_KB && (N + BN - 1) / BN <= 4;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if ((rc = make_tensor_map_2d(&ta_hi, a.hi, rows, K, BM, BK, 2))) return rc;
  if ((rc = make_tensor_map_2d(&ta_lo, a.lo, rows, K, BM, BK, 2))) return rc;
  if ((rc = make_tensor_map_2d(&tb_hi, w.hi, N, K, pair ? 64 : BN, BK, 2))) return rc;
  if ((rc = make_tensor_map_2d(&tb_lo, w.lo, N, K, pair ? 64 : BN, BK, 2))) return rc;
  // outputs: dense [32 rows][16 cols] boxes, no swizzle; unused maps alias a valid one
  CUtensorMap to_hi, to_lo, to_f32;
  if (out.hi) {
    if ((rc = make_tensor_map_2d(&to_hi, out.hi, rows, N, 32, UNIT, 2, 0))) return rc;
    if ((rc = make_tensor_map_2d(&to_lo, out.lo, rows, N, 32, UNIT, 2, 0))) return rc;
    to_f32 = to_hi;
  } else {
    if ((rc = make_tensor_map_2d(&to_f32, out_f32, rows, N, 32, UNIT, 4, 0))) return rc;
    to_hi = to_f32; to_lo = to_f32;
  }
  static DeviceOnce once;
  if (once.first_use()) {
    SSLAM_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          SMEM_TOTAL));
  }
  GemmParams gp;
  gp.rows = rows; gp.N = N; gp.K = K; gp.bias = bias; gp.res_hi = residual.hi; gp.res_lo = residual.lo;
  gp.relu = relu; gp.out_f32 = out_f32; gp.out_hi = out.hi; gp.out_lo = out.lo;
  gp.a_sum = a_ln.sum; gp.a_sq = a_ln.sq; gp.a_parts = a_ln.parts; gp.s1 = s1;
  gp.out_sum = out_stats ? out_stats->sum : nullptr;
  gp.out_sq = out_stats ? out_stats->sq : nullptr;
  gp.stat_stride = stat_stride_of(rows);
  gp.dbg = g_gemm_dbg;
  if (pair) {
    const int ntile = (N + BN - 1) / BN;
    const int nsp = (rows + 2 * BM - 1) / (2 * BM);
    int groups = 0;
    if (out_stats) out_stats->parts = ntile;
    const bool ln = a_ln.sum != nullptr, res = residual.hi != nullptr, st = out_stats != nullptr;
#define SSLAM_PAIR_LAUNCH(LN_, RES_, RELU_, F32_, ST_)                                                     \
  do {                                                                                                     \
    auto kfn = gemm_pair_kernel<LN_, RES_, RELU_, F32_, ST_>;                                              \
    static std::atomic<int> max_clusters[64][5] = {};             /* per device and cluster shape, 0 = not queried */ \
    int dev_ = 0;                                                                                          \
    SSLAM_CHECK_CUDA(cudaGetDevice(&dev_));                                                                \
    dev_ &= 63;                                                                                            \
    cudaLaunchConfig_t cfg = {};                                                                           \
    cudaLaunchAttribute attr[1];                                                                           \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                      \
    attr[0].val.clusterDim.x = 2 * ntile; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;      \
    cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = P_SMEM_TOTAL; cfg.stream = stream;            \
    cfg.attrs = attr; cfg.numAttrs = 1;                                                                    \
    int mcl = max_clusters[dev_][ntile].load();                                                               \
    if (mcl == 0) {                                                                                        \
      SSLAM_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_TOTAL)); \
      cfg.gridDim = dim3(2 * ntile);                                                                       \
      SSLAM_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&mcl, kfn, &cfg));                                   \
      SSLAM_REQUIRE(mcl >= 1, SSLAM_EUNSUPPORTED, "refiner: no cluster of %d CTAs fits on this device", 2 * ntile); \
      max_clusters[dev_][ntile].store(mcl);                                                                \
    }                                                                                                      \
    groups = mcl < nsp ? mcl : nsp;                               /* all clusters co-resident */            \
    cfg.gridDim = dim3(2u * ntile * groups);                                                               \
    SSLAM_LAUNCH(KK_GEMM, stream,                                                                          \
                 cudaLaunchKernelEx(&cfg, kfn, ta_hi, ta_lo, tb_hi, tb_lo, to_hi, to_lo, to_f32, gp, groups)); \
  } while (0)
    // the layer shapes of DescriptorRefiner: input projection, fc1, fc2 (+identity), output projection
    if (!ln && !res && relu && !out_f32 && st) SSLAM_PAIR_LAUNCH(false, false, true, false, true);
    else if (!ln && !res && relu && !out_f32 && !st) SSLAM_PAIR_LAUNCH(false, false, true, false, false);
    else if (ln && !res && relu && !out_f32 && st) SSLAM_PAIR_LAUNCH(true, false, true, false, true);
    else if (ln && res && relu && !out_f32 && st) SSLAM_PAIR_LAUNCH(true, true, true, false, true);
    else if (ln && res && relu && !out_f32 && !st) SSLAM_PAIR_LAUNCH(true, true, true, false, false);
    else if (!ln && !res && !relu && out_f32 && !st) SSLAM_PAIR_LAUNCH(false, false, false, true, false);
    else SSLAM_REQUIRE(false, SSLAM_EUNSUPPORTED, "refiner: layer option combination not instantiated");
#undef SSLAM_PAIR_LAUNCH
    return SSLAM_OK;
  }
  if (out_stats) out_stats->parts = 1;
  const int strips = (rows + BM - 1) / BM;
  const int grid = strips < num_sms() ? strips : num_sms();        // persistent: one CTA per SM
  SSLAM_LAUNCH(KK_GEMM, stream,
               gemm_f16x3_kernel<<<grid, NUM_THREADS, SMEM_TOTAL, stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, to_hi, to_lo,
                                                                            to_f32, gp));
  return SSLAM_OK;
}

// packed weights: per Linear the fp16 hi then lo copies of the [out, in] matrix; the LayerNorm-fed
// ones (fc1, fc2 of every block) are stored folded and followed by s1[out], c0[out] (fp32)
size_t pair_bytes(size_t n) { return 2 * align_up(n * 2, 256); }
size_t vec_bytes(size_t n) { return 2 * align_up(n * 4, 256); }
size_t packed_total(int C, int Hd, int D, int blocks) {
  return pair_bytes((size_t)Hd * C) + (size_t)blocks * 2 * (pair_bytes((size_t)Hd * Hd) + vec_bytes(Hd)) +
         pair_bytes((size_t)D * Hd);
}

}  // namespace
}  // namespace sslam

using namespace sslam;

// params order (device pointers, fp32):
//   [0] input_proj.weight [Hd,C]   [1] input_proj.bias [Hd]
//   per block b (8 entries from 2 + 8b): norm1.weight, norm1.bias, fc1.weight [Hd,Hd], fc1.bias,
//                                        norm2.weight, norm2.bias, fc2.weight [Hd,Hd], fc2.bias
//   [2+8*blocks] output_proj.weight [D,Hd]   [3+8*blocks] output_proj.bias [D]
extern "C" size_t sslam_refiner_packed_bytes(int C, int Hd, int D, int blocks) {
  if (C <= 0 || Hd <= 0 || D <= 0 || blocks < 0) return 0;
  return packed_total(C, Hd, D, blocks) + 256;
}

extern "C" int sslam_refiner_pack_weights(const float* const* params, int C, int Hd, int D, int blocks,
                                          void* packed, size_t packed_bytes, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(params && packed, SSLAM_EINVAL, "refiner_pack: null pointer");
  SSLAM_REQUIRE(C % 8 == 0 && Hd % 8 == 0 && D % 4 == 0, SSLAM_EUNSUPPORTED,
                "refiner: C and hidden must be multiples of 8, D of 4 (C=%d Hd=%d D=%d)", C, Hd, D);
  SSLAM_REQUIRE(packed_bytes >= sslam_refiner_packed_bytes(C, Hd, D, blocks), SSLAM_EWORKSPACE,
                "refiner_pack: packed buffer too small");
  char* w = static_cast<char*>(packed);
  auto take_pair = [&](size_t n) {
    Pair d{reinterpret_cast<__half*>(w), reinterpret_cast<__half*>(w + align_up(n * 2, 256))};
    w += pair_bytes(n);
    return d;
  };
  if ((rc = launch_split(params[0], take_pair((size_t)Hd * C), (size_t)Hd * C, stream))) return rc;
  for (int b = 0; b < blocks; ++b) {
    const float* const* bp = params + 2 + 8 * b;
    for (int half = 0; half < 2; ++half) {                 // fc1 folded with norm1, fc2 with norm2
      Pair d = take_pair((size_t)Hd * Hd);
      float* s1 = reinterpret_cast<float*>(w);
      float* c0 = reinterpret_cast<float*>(w + align_up((size_t)Hd * 4, 256));
      w += vec_bytes(Hd);
      SSLAM_LAUNCH(KK_SPLIT, stream,
                   fold_ln_kernel<<<(Hd + 7) / 8, 256, 0, stream>>>(bp[4 * half + 2], bp[4 * half], bp[4 * half + 1],
                                                                    bp[4 * half + 3], Hd, Hd, d.hi, d.lo, s1, c0));
    }
  }
  return launch_split(params[2 + 8 * blocks], take_pair((size_t)D * Hd), (size_t)D * Hd, stream);
}

extern "C" size_t sslam_refiner_workspace_bytes(int rows, int C, int Hd, int D, int blocks) {
  (void)blocks;
  if (rows <= 0) return 0;
  const size_t r = (size_t)rows;
  // pairs (4 B/element): x [r,C]; h_a, h_b, u [r,Hd];  fp32 raw [r,D];  3 sets of partial row sums
  return (r * C + 3 * r * Hd + r * D) * 4 + 6 * (size_t)stat_parts_of(Hd) * stat_stride_of(rows) * 4 + 16 * 256;
}

extern "C" int sslam_refiner_forward_f32(const float* const* params, const void* packed, const float* x,
                                         const void* x_hi, const void* x_lo,
                                         int rows, int C, int Hd, int D, int blocks, float eps_norm,
                                         float* out_f32, void* out_bf16, void* out_hi, void* out_lo,
                                         void* ws, size_t ws_bytes, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(rows >= 0, SSLAM_EINVAL, "refiner: negative rows");
  if (rows == 0) return SSLAM_OK;
  SSLAM_REQUIRE(params && packed && (x || (x_hi && x_lo)) && ws && (out_f32 || out_bf16 || out_hi), SSLAM_EINVAL,
                "refiner: null pointer");
  SSLAM_REQUIRE(C % 8 == 0 && Hd % 8 == 0 && D % 4 == 0 && Hd <= MAX_N && D <= MAX_N, SSLAM_EUNSUPPORTED,
                "refiner: C and hidden must be multiples of 8, D of 4, hidden and D <= 1024 (C=%d Hd=%d D=%d)", C, Hd, D);
  SSLAM_REQUIRE(!x || (reinterpret_cast<uintptr_t>(x) & 15) == 0, SSLAM_EINVAL, "refiner: x must be 16-byte aligned");
  SSLAM_REQUIRE(ws_bytes >= sslam_refiner_workspace_bytes(rows, C, Hd, D, blocks), SSLAM_EWORKSPACE,
                "refiner: workspace %zu < %zu", ws_bytes, sslam_refiner_workspace_bytes(rows, C, Hd, D, blocks));
  const size_t r = (size_t)rows;
  char* wp = static_cast<char*>(ws);
  auto take_pair = [&](size_t n) {
    Pair q{reinterpret_cast<__half*>(wp), reinterpret_cast<__half*>(wp + align_up(n * 2, 256))};
    wp += pair_bytes(n);
    return q;
  };
  auto take_stats = [&]() {
    const size_t n = (size_t)stat_parts_of(Hd) * stat_stride_of(rows);
    RowStats st{reinterpret_cast<float*>(wp), reinterpret_cast<float*>(wp) + n, 0};
    wp += 2 * n * 4;
    return st;
  };
  Pair xs = take_pair(r * C), h_a = take_pair(r * Hd), h_b = take_pair(r * Hd), u = take_pair(r * Hd);
  RowStats st_a = take_stats(), st_b = take_stats(), st_u = take_stats();
  float* raw = reinterpret_cast<float*>(wp);
  const char* pk = static_cast<const char*>(packed);
  auto next_w = [&](size_t n) {
    Pair q{reinterpret_cast<__half*>(const_cast<char*>(pk)),
           reinterpret_cast<__half*>(const_cast<char*>(pk) + align_up(n * 2, 256))};
    pk += pair_bytes(n);
    return q;
  };
  const Pair none{nullptr, nullptr};
  const RowStats no_stats{nullptr, nullptr, 0};

  if (x) {
    if ((rc = launch_split(x, xs, r * C, stream))) return rc;
  } else {                                                                // pair written by the gather kernel
    xs.hi = static_cast<__half*>(const_cast<void*>(x_hi));
    xs.lo = static_cast<__half*>(const_cast<void*>(x_lo));
  }
  Pair w = next_w((size_t)Hd * C);                                        // descriptor_refiner.py:76
  if ((rc = launch_gemm(xs, w, rows, Hd, C, params[1], no_stats, nullptr, none, 1, nullptr, h_a,
                        blocks ? &st_a : nullptr, stream)))
    return rc;
  Pair h_cur = h_a, h_nxt = h_b;
  RowStats st_cur = st_a, st_nxt = st_b;
  for (int b = 0; b < blocks; ++b) {                                      // :79-80, :108-126
    // fc1( LN1(h) ) + ReLU, LayerNorm folded into the weights and the epilogue
    w = next_w((size_t)Hd * Hd);
    const float* s1 = reinterpret_cast<const float*>(pk);
    const float* c0 = reinterpret_cast<const float*>(pk + align_up((size_t)Hd * 4, 256));
    pk += vec_bytes(Hd);
    if ((rc = launch_gemm(h_cur, w, rows, Hd, Hd, c0, st_cur, s1, none, 1, nullptr, u, &st_u, stream))) return rc;
    // fc2( LN2(u) ) + identity, ReLU
    w = next_w((size_t)Hd * Hd);
    s1 = reinterpret_cast<const float*>(pk);
    c0 = reinterpret_cast<const float*>(pk + align_up((size_t)Hd * 4, 256));
    pk += vec_bytes(Hd);
    const bool last = (b == blocks - 1);
    if ((rc = launch_gemm(u, w, rows, Hd, Hd, c0, st_u, s1, h_cur, 1, nullptr, h_nxt, last ? nullptr : &st_nxt,
                          stream)))
      return rc;
    Pair tp = h_cur; h_cur = h_nxt; h_nxt = tp;
    RowStats ts = st_cur; st_cur = st_nxt; st_nxt = ts;
  }
  w = next_w((size_t)D * Hd);                                             // :83
  if ((rc = launch_gemm(h_cur, w, rows, D, Hd, params[3 + 8 * blocks], no_stats, nullptr, none, 0, raw, none,
                        nullptr, stream)))
    return rc;
  return sslam_l2norm_rows(raw, rows, D, eps_norm, out_f32, out_bf16, out_hi, out_lo, stream_);   // :86
}

// Debug aid for tools/: per-CTA cycle counters of gemm_pair_kernel ({MMA thread: total, wait_full,
// wait_tempty, wait_weights; epilogue warp 2: total, wait_tfull, wait_store, tiles}, 8 int64 per CTA)
// are written to buf (device) while buf != NULL.
extern "C" void sslam_debug_gemm_stalls(long long* buf) { sslam::g_gemm_dbg = buf; }

// Debug aid for tools/: host-mapped buffer (device pointer) that receives watchdog records of this
// translation unit's kernels; see tc_common.cuh.
extern "C" int sslam_debug_watchdog_gemm(unsigned long long* buf) {
  return (int)cudaMemcpyToSymbol(sslam::tc::g_watchdog_buf, &buf, sizeof(buf));
}
