// Evaluation adaptors of the match path: nearest warped keypoint, ground-truth
// match lists, match scoring — the N x M pairwise-distance problems of the reference's evaluation
// scripts with the same "best per row" epilogue as the descriptor matcher.
//
//   DescriptorQualityTester.compute_ground_truth_matches   test/test_descriptor_quality.py:144-183
//   DescriptorQualityTester.evaluate_matches               test/test_descriptor_quality.py:185-231
//   RepeatabilityTester.compute_repeatability              test/test_repeatability.py:79-128
//
// Arithmetic follows NumPy's: with a homography the keypoints are promoted to float64 (the
// reference concatenates a float64 ones column, :164 / :100) and everything downstream is double;
// without one (test_repeatability.py:104-105) the float32 keypoints are subtracted and normed in
// float32.  No FMA contraction anywhere: products and sums are rounded separately, as NumPy's
// element-wise loops do.  argmin returns the lowest minimal index.
#include "common.cuh"

namespace sslam {
namespace {

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 1024;            // keypoints of set 2 staged per shared-memory tile

template <typename T> struct Arith;
template <> struct Arith<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dadd_rn(a, -b); }
  static __device__ __forceinline__ double root(double a) { return sqrt(a); }
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
};
template <> struct Arith<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float root(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};

struct NnParams {
  const float* k1;            // [F1,N,2]
  const float* k2;            // [F2,M,2]
  const double* H;            // [P,9] row-major, or null
  const int32_t* pair_index;  // [P,2] or null
  int P, N, M;
  void* min_dist;             // [P,N] double (with H) or float (without)
  int32_t* argmin;            // [P,N]
};

// One thread per keypoint of set 1; set 2 streams through shared memory in tiles.
template <typename T>
__global__ void __launch_bounds__(NN_THREADS) nn_points_kernel(NnParams p) {
  __shared__ float2 tile[NN_TILE];
  const int pair = blockIdx.y;
  int ia = pair, ib = pair;
  if (p.pair_index) { ia = p.pair_index[2 * pair]; ib = p.pair_index[2 * pair + 1]; }
  const int i = blockIdx.x * NN_THREADS + threadIdx.x;
  const bool live = i < p.N;
  T wx = 0, wy = 0;
  if (live) {
    const float2 a = reinterpret_cast<const float2*>(p.k1)[(size_t)ia * p.N + i];
    if (p.H) {
      // (H @ [x, y, 1]^T) then divide by the third component (:165-166 / :101-102), in double
      const double* h = p.H + 9 * (size_t)pair;
      const double x = (double)a.x, y = (double)a.y;
      const double u = __dadd_rn(__dadd_rn(__dmul_rn(h[0], x), __dmul_rn(h[1], y)), h[2]);
      const double v = __dadd_rn(__dadd_rn(__dmul_rn(h[3], x), __dmul_rn(h[4], y)), h[5]);
      const double w = __dadd_rn(__dadd_rn(__dmul_rn(h[6], x), __dmul_rn(h[7], y)), h[8]);
      wx = (T)__ddiv_rn(u, w);
      wy = (T)__ddiv_rn(v, w);
    } else {
      wx = (T)a.x; wy = (T)a.y;
    }
  }
  T best = Arith<T>::inf();
  int best_j = 0;
  const float2* k2 = reinterpret_cast<const float2*>(p.k2) + (size_t)ib * p.M;
  for (int j0 = 0; j0 < p.M; j0 += NN_TILE) {
    const int n = min(NN_TILE, p.M - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += NN_THREADS) tile[t] = k2[j0 + t];
    __syncthreads();
    if (live) {
      for (int t = 0; t < n; ++t) {
        const T dx = Arith<T>::sub(wx, (T)tile[t].x);
        const T dy = Arith<T>::sub(wy, (T)tile[t].y);
        const T d = Arith<T>::root(Arith<T>::add(Arith<T>::mul(dx, dx), Arith<T>::mul(dy, dy)));
        if (d < best) { best = d; best_j = j0 + t; }   // strict: lowest index wins ties
      }
    }
  }
  if (live) {
    const size_t o = (size_t)pair * p.N + i;
    reinterpret_cast<T*>(p.min_dist)[o] = best;
    p.argmin[o] = best_j;
  }
}

// Rows whose nearest distance is below the threshold, in ascending i (:173-181).
struct GtParams {
  const void* min_dist; const int32_t* argmin; int is_double; double threshold;
  int P, N; int32_t* pairs; int32_t* counts;
};

__global__ void __launch_bounds__(256) gt_matches_kernel(GtParams g) {
  __shared__ int scan[256];
  __shared__ int total_s;
  const int pair = blockIdx.x, tid = threadIdx.x;
  const int seg = (g.N + 255) / 256;
  const int beg = min(g.N, tid * seg), end = min(g.N, beg + seg);
  auto keep = [&](int i) {
    const size_t o = (size_t)pair * g.N + i;
    // float32 distances compare against the threshold as float32 (NumPy >= 2 scalar promotion)
    return g.is_double ? reinterpret_cast<const double*>(g.min_dist)[o] < g.threshold
                       : reinterpret_cast<const float*>(g.min_dist)[o] < (float)g.threshold;
  };
  int cnt = 0;
  for (int i = beg; i < end; ++i) cnt += keep(i) ? 1 : 0;
  scan[tid] = cnt;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    int v = (tid >= o) ? scan[tid - o] : 0;
    __syncthreads();
    scan[tid] += v;
    __syncthreads();
  }
  int pos = scan[tid] - cnt;
  if (tid == 255) total_s = scan[255];
  __syncthreads();
  const int total = total_s;
  int32_t* pr = g.pairs ? g.pairs + (size_t)pair * g.N * 2 : nullptr;
  if (pr) {
    for (int i = beg; i < end; ++i)
      if (keep(i)) { pr[2 * pos] = i; pr[2 * pos + 1] = g.argmin[(size_t)pair * g.N + i]; ++pos; }
    for (int i = total + tid; i < g.N; i += 256) { pr[2 * i] = -1; pr[2 * i + 1] = -1; }
  }
  if (tid == 0) g.counts[pair] = total;
}

// tp / fp / fn of predicted against ground-truth pairs (:202-215).  Both lists hold each i at most
// once (mutual-NN lists and nearest-point lists do), so set intersection is a lookup by i.
struct EvalParams {
  const int32_t* pred; const int32_t* pred_counts; int pred_stride;   // [P,pred_stride,2]
  const int32_t* gt; const int32_t* gt_counts; int gt_stride;         // [P,gt_stride,2]
  int P, N; int32_t* lut;   // [P,N] scratch
  int32_t* out;             // [P,3] tp, fp, fn
};

__global__ void __launch_bounds__(256) eval_matches_kernel(EvalParams e) {
  __shared__ int red[256];
  const int pair = blockIdx.x, tid = threadIdx.x;
  int32_t* lut = e.lut + (size_t)pair * e.N;
  for (int i = tid; i < e.N; i += 256) lut[i] = -1;
  __syncthreads();
  const int ng = e.gt_counts[pair], np = e.pred_counts[pair];
  const int32_t* gt = e.gt + (size_t)pair * e.gt_stride * 2;
  const int32_t* pr = e.pred + (size_t)pair * e.pred_stride * 2;
  for (int k = tid; k < ng; k += 256) {
    const int i = gt[2 * k];
    if (i >= 0 && i < e.N) lut[i] = gt[2 * k + 1];
  }
  __syncthreads();
  int tp = 0;
  for (int k = tid; k < np; k += 256) {
    const int i = pr[2 * k];
    if (i >= 0 && i < e.N && lut[i] >= 0 && lut[i] == pr[2 * k + 1]) ++tp;
  }
  red[tid] = tp;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    e.out[3 * pair] = red[0];
    e.out[3 * pair + 1] = np - red[0];
    e.out[3 * pair + 2] = ng - red[0];
  }
}

}  // namespace
}  // namespace sslam

using namespace sslam;

extern "C" int sslam_nn_points(const float* kpts1, int F1, const float* kpts2, int F2, const double* H,
                               const int32_t* pair_index, int P, int N, int M, void* min_dist,
                               int32_t* argmin, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(P >= 0 && N >= 0 && M >= 0, SSLAM_EINVAL, "nn_points: bad size");
  if (P == 0 || N == 0) return SSLAM_OK;
  SSLAM_REQUIRE(M > 0, SSLAM_EINVAL, "nn_points: M must be > 0 (min of an empty row)");
  SSLAM_REQUIRE(kpts1 && kpts2 && min_dist && argmin, SSLAM_EINVAL, "nn_points: null pointer");
  SSLAM_REQUIRE(F1 > 0 && F2 > 0 && (pair_index || (F1 >= P && F2 >= P)), SSLAM_EINVAL,
                "nn_points: banks hold %d / %d sets but %d implicit pairs were requested", F1, F2, P);
  NnParams p{kpts1, kpts2, H, pair_index, P, N, M, min_dist, argmin};
  dim3 grid((N + NN_THREADS - 1) / NN_THREADS, P);
  if (H) SSLAM_LAUNCH(KK_EVAL, stream, nn_points_kernel<double><<<grid, NN_THREADS, 0, stream>>>(p));
  else SSLAM_LAUNCH(KK_EVAL, stream, nn_points_kernel<float><<<grid, NN_THREADS, 0, stream>>>(p));
  return SSLAM_OK;
}

extern "C" int sslam_gt_matches(const void* min_dist, const int32_t* argmin, int is_double, double threshold,
                                int P, int N, int32_t* pairs, int32_t* counts, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(P >= 0 && N >= 0, SSLAM_EINVAL, "gt_matches: bad size");
  if (P == 0) return SSLAM_OK;
  SSLAM_REQUIRE(min_dist && argmin && counts, SSLAM_EINVAL, "gt_matches: null pointer");
  GtParams g{min_dist, argmin, is_double, threshold, P, N, pairs, counts};
  SSLAM_LAUNCH(KK_EVAL, stream, gt_matches_kernel<<<P, 256, 0, stream>>>(g));
  return SSLAM_OK;
}

extern "C" int sslam_eval_matches(const int32_t* pred, const int32_t* pred_counts, int pred_stride,
                                  const int32_t* gt, const int32_t* gt_counts, int gt_stride, int P, int N,
                                  int32_t* scratch, int32_t* out, void* stream_) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  SSLAM_REQUIRE(P >= 0 && N >= 0 && pred_stride >= 0 && gt_stride >= 0, SSLAM_EINVAL, "eval_matches: bad size");
  if (P == 0) return SSLAM_OK;
  SSLAM_REQUIRE(pred && pred_counts && gt && gt_counts && scratch && out, SSLAM_EINVAL,
                "eval_matches: null pointer");
  EvalParams e{pred, pred_counts, pred_stride, gt, gt_counts, gt_stride, P, N, scratch, out};
  SSLAM_LAUNCH(KK_EVAL, stream, eval_matches_kernel<<<P, 256, 0, stream>>>(e));
  return SSLAM_OK;
}
