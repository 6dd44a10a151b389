// Exact-mode (fp32 FMA) similarity with fused row top-2 / column argmax, plus the column-key
// unpack and the per-variant finalise kernels.  The similarity matrix is never stored.
//
// Replaces, for every matcher variant:
//   S = D1.D2^T, argmax over rows and columns, second-best per row
//   (visualize_matches.py:105-109,117-119; visualize_matches_sequence.py:144-146;
//    test/test_descriptor_quality.py:116-130; train.py:422-424; test/test_tracking.py:159-160)
// and the acceptance rules + ascending-i compaction
//   (visualize_matches.py:112-122; visualize_matches_sequence.py:147-192;
//    test/test_descriptor_quality.py:126-140; train.py:425-433; test/test_tracking.py:161).
//
// Layout: one CTA owns a 128-row strip of one pair, keeps the strip of D1 resident in shared
// memory (k-major) and streams 128-column tiles of D2 through a double-buffered 16-deep k chunk.
// Row top-2 lives in registers across the column loop; column argmax is reduced per tile and
// merged across strips with a 64-bit atomicMax on (ordered value << 32 | ~row), which yields the
// largest value and, among equals, the lowest row.
#include "common.cuh"

namespace sslam {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int LDS_ = BM + 4;          // padded leading dimension (floats), keeps 16B alignment
constexpr int THREADS = 256;
constexpr int MAX_D = 256;

struct Top2 {
  float best, second;
  int idx;
  __device__ __forceinline__ void init() {
    best = __int_as_float(0xff800000); second = best; idx = 0x7fffffff;
  }
  // candidates arrive in ascending index order -> strict '>' keeps the lowest index
  __device__ __forceinline__ void push(float v, int i) {
    if (v > best) { second = best; best = v; idx = i; }
    else second = fmaxf(second, v);
  }
  __device__ __forceinline__ void merge(float b2, float s2, int i2) {
    if (b2 > best || (b2 == best && i2 < idx)) { second = fmaxf(best, s2); best = b2; idx = i2; }
    else second = fmaxf(second, b2);
  }
};

struct MatchParams {
  const float* bank1;
  const float* bank2;
  const int32_t* pair_index;
  int P, N, M, D;
  int32_t* nn12;
  float* best12;
  float* second12;
  u64* colkeys;       // [P][M]
};

__device__ __forceinline__ int row_of(int t, int i) { return (i < 4) ? t * 4 + i : 64 + t * 4 + (i - 4); }

__global__ void __launch_bounds__(THREADS, 1) match_f32_kernel(MatchParams p) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, N = p.N, M = p.M;
  const int dpad = (D + BK - 1) / BK * BK;
  float* As = smem;                                   // [dpad][LDS_]
  float* Bs = As + (size_t)dpad * LDS_;               // [2][BK][LDS_]
  u64* colpart = reinterpret_cast<u64*>(Bs + 2 * BK * LDS_);   // [8][BN]

  const int pair = blockIdx.y;
  int ia = pair, ib = pair;
  if (p.pair_index) { ia = p.pair_index[2 * pair]; ib = p.pair_index[2 * pair + 1]; }
  const float* A = p.bank1 + (size_t)ia * N * D;
  const float* Bm = p.bank2 + (size_t)ib * M * D;
  const int row0 = blockIdx.x * BM;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;

  // ---- resident A strip, k-major
  const int dq = dpad / 4;
  for (int e = tid; e < BM * dq; e += THREADS) {
    int r = e / dq, kq = e - r * dq;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N && kq * 4 < D) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(row0 + r) * D + kq * 4));
    As[(kq * 4 + 0) * LDS_ + r] = v.x; As[(kq * 4 + 1) * LDS_ + r] = v.y;
    As[(kq * 4 + 2) * LDS_ + r] = v.z; As[(kq * 4 + 3) * LDS_ + r] = v.w;
  }

  const int nkb = dpad / BK;
  const int ntile = (M + BN - 1) / BN;
  const int total = nkb * ntile;

  // each thread stages two float4 of the B chunk: e = tid + 256*j -> col = e/4, kq = e%4
  float4 stage[2];
  auto fetch = [&](int it) {
    int ct = it / nkb, kb = it - ct * nkb;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int e = tid + THREADS * j, col = e >> 2, kq = e & 3;
      int gc = ct * BN + col, gk = kb * BK + kq * 4;
      stage[j] = (gc < M && gk < D) ? __ldg(reinterpret_cast<const float4*>(Bm + (size_t)gc * D + gk))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto commit = [&](int buf) {
    float* dst = Bs + buf * BK * LDS_;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int e = tid + THREADS * j, col = e >> 2, kq = e & 3;
      dst[(kq * 4 + 0) * LDS_ + col] = stage[j].x; dst[(kq * 4 + 1) * LDS_ + col] = stage[j].y;
      dst[(kq * 4 + 2) * LDS_ + col] = stage[j].z; dst[(kq * 4 + 3) * LDS_ + col] = stage[j].w;
    }
  };

  Top2 rowtop[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) rowtop[i].init();
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  fetch(0);
  commit(0);
  __syncthreads();

  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    const int ct = it / nkb, kb = it - ct * nkb;
    if (it + 1 < total) fetch(it + 1);
    const float* a_base = As + (size_t)(kb * BK) * LDS_;
    const float* b_base = Bs + buf * BK * LDS_;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(a_base + kk * LDS_ + ty * 4);
      float4 a1 = *reinterpret_cast<const float4*>(a_base + kk * LDS_ + 64 + ty * 4);
      float4 b0 = *reinterpret_cast<const float4*>(b_base + kk * LDS_ + tx * 4);
      float4 b1 = *reinterpret_cast<const float4*>(b_base + kk * LDS_ + 64 + tx * 4);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    if (kb == nkb - 1) {
      // ---------------- fused epilogue for column tile ct
      const int c0 = ct * BN;
      // rows: local top-2 over this thread's 8 columns (ascending), then across the 16 tx lanes
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        Top2 t; t.init();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int gc = c0 + row_of(tx, j);
          if (gc < M) t.push(acc[i][j], gc);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          float b2 = __shfl_xor_sync(0xffffffffu, t.best, o);
          float s2 = __shfl_xor_sync(0xffffffffu, t.second, o);
          int i2 = __shfl_xor_sync(0xffffffffu, t.idx, o);
          t.merge(b2, s2, i2);
        }
        rowtop[i].merge(t.best, t.second, t.idx);
      }
      // columns: max over this thread's 8 rows (ascending), lanes tx / tx+16, then the 8 warps
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        u64 key = 0;
        float bv = 0.f; int bi = -1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int gr = row0 + row_of(ty, i);
          if (gr < N && (bi < 0 || acc[i][j] > bv)) { bv = acc[i][j]; bi = gr; }
        }
        if (bi >= 0) key = pack_key(bv, (u32)bi);
        u64 other = __shfl_xor_sync(0xffffffffu, key, 16);
        key = other > key ? other : key;
        if (lane < 16) colpart[warp * BN + row_of(tx, j)] = key;
      }
      __syncthreads();
      if (tid < BN) {
        u64 k = colpart[tid];
#pragma unroll
        for (int w = 1; w < 8; ++w) { u64 o = colpart[w * BN + tid]; k = o > k ? o : k; }
        if (c0 + tid < M && k) atomicMax(p.colkeys + (size_t)pair * M + c0 + tid, k);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    if (it + 1 < total) commit(buf ^ 1);
    __syncthreads();
  }

  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int gr = row0 + row_of(ty, i);
      if (gr < N) {
        size_t o = (size_t)pair * N + gr;
        p.nn12[o] = rowtop[i].idx;
        p.best12[o] = rowtop[i].best;
        p.second12[o] = rowtop[i].second;
      }
    }
  }
}

__global__ void unpack_cols_kernel(const u64* __restrict__ keys, long long total, int32_t* nn21,
                                   float* best21) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  u64 k = keys[i];
  nn21[i] = (int32_t)key_index(k);
  best21[i] = key_value(k);
}

// ------------------------------------------------------------------------------ finalise
struct FinalizeParams {
  int variant;
  float prm[8];
  double prm_d[2];                                          // M1 with legacy promotion: ratio as a double
  const int32_t* pair_index;
  int P, N, M;
  const int32_t* nn12; const float* best12; const float* second12;
  const int32_t* nn21; const float* best21;
  const float* scores1; const float* scores2; const float* inten1; const float* inten2;
  int32_t* pairs; float* pair_scores; int32_t* counts;
};

__device__ __forceinline__ bool accept_row(const FinalizeParams& f, int pair, int ia, int ib, int i,
                                           int& j, float& score) {
  const size_t ro = (size_t)pair * f.N + i;
  j = f.nn12[ro];
  const float best = f.best12[ro];
  const bool mutual = (f.nn21[(size_t)pair * f.M + j] == i);
  switch (f.variant) {
    case SSLAM_MATCH_M1: {                                 // visualize_matches.py:114-122
      float second = fmaxf(f.second12[ro], -1.0f);         // best column overwritten with -1 (:118)
      score = best;
      // `sim > second_best * ratio_thresh` on NumPy scalars: fp32 under NumPy 2 (NEP 50, the Python float
      // is cast to fp32), double under the NumPy < 2 the reference pins (requirements.txt:1)
      if (f.prm[1] != 0.f) return mutual && ((double)best > __dmul_rn((double)second, f.prm_d[0]));
      return mutual && (best > __fmul_rn(second, f.prm[0]));
    }
    case SSLAM_MATCH_M2: {                                 // visualize_matches_sequence.py:158-192
      float s1 = f.scores1[(size_t)ia * f.N + i], s2 = f.scores2[(size_t)ib * f.M + j];
      float avg = __fdiv_rn(__fadd_rn(s1, s2), 2.0f);
      bool ok = mutual && (avg >= f.prm[1]) && (best >= f.prm[2]);
      if (f.inten1 && f.inten2) {
        float a = f.inten1[(size_t)ia * f.N + i], b = f.inten2[(size_t)ib * f.M + j];
        ok = ok && (__fdiv_rn(__fadd_rn(a, b), 2.0f) >= f.prm[3]);
      }
      score = __fadd_rn(__fmul_rn(f.prm[4], best), __fmul_rn(f.prm[0], avg));
      return ok;
    }
    case SSLAM_MATCH_M3: {                                 // test_descriptor_quality.py:126-140
      float ratio = __fdiv_rn(f.second12[ro], __fadd_rn(best, 1e-8f));
      score = __fsub_rn(1.0f, best);
      return mutual && (ratio < f.prm[0]);
    }
    case SSLAM_MATCH_M4:                                   // train.py:425-428
      score = best;
      return mutual;
    default:                                               // M5, test_tracking.py:160-161
      score = best;
      return best > f.prm[0];
  }
}

__global__ void __launch_bounds__(256) finalize_kernel(FinalizeParams f) {
  __shared__ int scan[256];
  __shared__ int total_s;
  const int pair = blockIdx.x, tid = threadIdx.x;
  int ia = pair, ib = pair;
  if (f.pair_index) { ia = f.pair_index[2 * pair]; ib = f.pair_index[2 * pair + 1]; }
  const int seg = (f.N + 255) / 256;
  const int beg = min(f.N, tid * seg), end = min(f.N, beg + seg);
  int cnt = 0;
  for (int i = beg; i < end; ++i) { int j; float s; cnt += accept_row(f, pair, ia, ib, i, j, s) ? 1 : 0; }
  scan[tid] = cnt;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {                       // Hillis-Steele inclusive scan
    int v = (tid >= o) ? scan[tid - o] : 0;
    __syncthreads();
    scan[tid] += v;
    __syncthreads();
  }
  int pos = scan[tid] - cnt;
  if (tid == 255) total_s = scan[255];
  __syncthreads();
  const int total = total_s;
  int32_t* pr = f.pairs + (size_t)pair * f.N * 2;
  float* ps = f.pair_scores + (size_t)pair * f.N;
  for (int i = beg; i < end; ++i) {
    int j; float s;
    if (accept_row(f, pair, ia, ib, i, j, s)) { pr[2 * pos] = i; pr[2 * pos + 1] = j; ps[pos] = s; ++pos; }
  }
  for (int i = total + tid; i < f.N; i += 256) { pr[2 * i] = -1; pr[2 * i + 1] = -1; ps[i] = 0.f; }
  if (tid == 0) f.counts[pair] = total;
}

}  // namespace

// implemented in match_tc.cu
int match_top2_tc(const void* bank1, const void* bank1_lo, int F1, const void* bank2, const void* bank2_lo,
                  int F2, const int32_t* pair_index,
                  int dtype, int P, int N, int M, int D, int32_t* nn12, float* best12, float* second12,
                  u64* colkeys, void* ws_extra, size_t ws_extra_bytes, cudaStream_t stream);
size_t match_tc_extra_workspace(int P, int N, int M, int D, int dtype, int F1, int F2);

}  // namespace sslam

using namespace sslam;

extern "C" size_t sslam_match_workspace_bytes(int F1, int F2, int P, int N, int M, int D, int dtype) {
  if (P <= 0 || N <= 0 || M <= 0) return 0;
  size_t base = align_up((size_t)P * M * sizeof(u64), 256);
  if (dtype != SSLAM_SIM_F32) base += match_tc_extra_workspace(P, N, M, D, dtype, F1, F2);
  return base;
}

extern "C" int sslam_match_top2(const void* bank1, const void* bank1_lo, int F1, const void* bank2,
                                const void* bank2_lo, int F2, const int32_t* pair_index,
                                int dtype, int P, int N, int M, int D, int32_t* nn12, float* best12,
                                float* second12, int32_t* nn21, float* best21, void* ws,
                                size_t ws_bytes, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(P >= 0 && N >= 0 && M >= 0 && D > 0, SSLAM_EINVAL, "match: bad size");
  if (P == 0 || N == 0) return SSLAM_OK;
  SSLAM_REQUIRE(M > 0, SSLAM_EINVAL, "match: M must be > 0 (argmax of an empty row)");
  SSLAM_REQUIRE(bank1 && bank2 && nn12 && best12 && second12 && nn21 && best21 && ws, SSLAM_EINVAL,
                "match: null pointer");
  SSLAM_REQUIRE(D % 4 == 0 && D <= MAX_D, SSLAM_EUNSUPPORTED, "match: D=%d (need D%%4==0, D<=256)", D);
  SSLAM_REQUIRE(dtype == SSLAM_SIM_F32 || dtype == SSLAM_SIM_TF32X3 || dtype == SSLAM_SIM_BF16 ||
                    dtype == SSLAM_SIM_F16X3,
                SSLAM_EINVAL, "match: unknown dtype %d", dtype);
  SSLAM_REQUIRE(F1 > 0 && F2 > 0 && (pair_index || (F1 >= P && F2 >= P)), SSLAM_EINVAL,
                "match: banks hold %d / %d sets but %d implicit pairs were requested", F1, F2, P);
  const bool presplit = bank1_lo != nullptr || bank2_lo != nullptr;
  SSLAM_REQUIRE(!presplit || (dtype == SSLAM_SIM_F16X3 && bank1_lo && bank2_lo), SSLAM_EINVAL,
                "match: pre-split banks need SSLAM_SIM_F16X3 and both lo arrays");
  {
    const size_t need = presplit ? align_up((size_t)P * M * sizeof(u64), 256)
                                 : sslam_match_workspace_bytes(F1, F2, P, N, M, D, dtype);
    SSLAM_REQUIRE(ws_bytes >= need, SSLAM_EWORKSPACE, "match: workspace %zu < %zu", ws_bytes, need);
  }
  SSLAM_REQUIRE((reinterpret_cast<uintptr_t>(bank1) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(bank2) & 15) == 0, SSLAM_EINVAL,
                "match: descriptor banks must be 16-byte aligned");
  u64* colkeys = reinterpret_cast<u64*>(ws);
  const size_t colbytes = align_up((size_t)P * M * sizeof(u64), 256);
  SSLAM_CHECK_CUDA(cudaMemsetAsync(colkeys, 0, (size_t)P * M * sizeof(u64), stream));

  if (dtype == SSLAM_SIM_F32) {
    MatchParams mp;
    mp.bank1 = (const float*)bank1; mp.bank2 = (const float*)bank2; mp.pair_index = pair_index;
    mp.P = P; mp.N = N; mp.M = M; mp.D = D;
    mp.nn12 = nn12; mp.best12 = best12; mp.second12 = second12; mp.colkeys = colkeys;
    const int dpad = (D + BK - 1) / BK * BK;
    size_t smem = ((size_t)dpad * LDS_ + 2 * BK * LDS_) * 4 + 8 * BN * sizeof(u64);
    static DeviceOnce once;
    if (once.first_use()) {
      SSLAM_CHECK_CUDA(cudaFuncSetAttribute(match_f32_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    dim3 grid((N + BM - 1) / BM, P);
    SSLAM_LAUNCH(KK_MATCH_F32, stream,
                 match_f32_kernel<<<grid, THREADS, smem, stream>>>(mp));
  } else {
    rc = match_top2_tc(bank1, bank1_lo, F1, bank2, bank2_lo, F2, pair_index, dtype, P, N, M, D, nn12, best12,
                       second12, colkeys,
                       reinterpret_cast<char*>(ws) + colbytes, ws_bytes - colbytes, stream);
    if (rc) return rc;
  }
  const long long total = (long long)P * M;
  SSLAM_LAUNCH(KK_UNPACK, stream,
               unpack_cols_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(colkeys, total, nn21, best21));
  return SSLAM_OK;
}

extern "C" int sslam_match_finalize(int variant, const double* params, const int32_t* pair_index,
                                    int P, int N, int M, const int32_t* nn12, const float* best12,
                                    const float* second12, const int32_t* nn21, const float* best21,
                                    const float* scores1, const float* scores2, const float* inten1,
                                    const float* inten2, int32_t* pairs, float* pair_scores,
                                    int32_t* counts, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(variant >= SSLAM_MATCH_M1 && variant <= SSLAM_MATCH_M5, SSLAM_EINVAL,
                "finalize: unknown variant %d", variant);
  SSLAM_REQUIRE(P >= 0 && N >= 0 && M > 0, SSLAM_EINVAL, "finalize: bad size");
  if (P == 0) return SSLAM_OK;
  SSLAM_REQUIRE(params && nn12 && best12 && second12 && nn21 && best21 && pairs && pair_scores && counts,
                SSLAM_EINVAL, "finalize: null pointer");
  SSLAM_REQUIRE(variant != SSLAM_MATCH_M2 || (scores1 && scores2), SSLAM_EINVAL,
                "finalize: M2 needs saliency scores");
  FinalizeParams f;
  f.variant = variant;
  for (int i = 0; i < 8; ++i) f.prm[i] = (float)params[i];
  f.prm_d[0] = params[0]; f.prm_d[1] = params[1];
  f.pair_index = pair_index; f.P = P; f.N = N; f.M = M;
  f.nn12 = nn12; f.best12 = best12; f.second12 = second12; f.nn21 = nn21; f.best21 = best21;
  f.scores1 = scores1; f.scores2 = scores2; f.inten1 = inten1; f.inten2 = inten2;
  f.pairs = pairs; f.pair_scores = pair_scores; f.counts = counts;
  SSLAM_LAUNCH(KK_FINALIZE, stream,
               finalize_kernel<<<P, 256, 0, stream>>>(f));
  return SSLAM_OK;
}
