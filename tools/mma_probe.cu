// Micro-probe: issue rate of tcgen05.mma (kind::f16, SS operands, cta_group::1) on sm_100a for the
// shapes the library uses.  One CTA per SM; thread 0 issues `iters` batches of MMAs on zeroed
// shared memory (no TMA traffic) and waits for their commit; reports cycles per instruction and the
// SM clock implied by globaltimer.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
//   -I../semantic-slam-master_b200/csrc mma_probe.cu -o mma_probe   (tools only; not part of the library)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

using namespace sslam::tc;

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// variant: 0 = N128 x1 accumulator, 1 = N256, 2 = three N128 (two accumulators, as the old 3-MMA step),
//          3 = N256 + N128 (concat step), 4 = N128, A operand re-used at the same address (bank pattern)
__global__ void __launch_bounds__(128, 1) probe(int variant, int iters, long long* out_cycles, long long* out_ns) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t i128 = make_instr_desc(FMT_F16, 128, 128), i256 = make_instr_desc(FMT_F16, 128, 256);
    const uint32_t base = smem_u32(smem);
    uint32_t ph = 0;
    long long t0 = clock64();
    unsigned long long g0 = gtimer();
    for (int it = 0; it < iters; ++it) {
      // one "k-block": 4 k-steps of 32 bytes inside a 128-byte swizzle atom; operands rotate over
      // 4 stages of 64 KB / 4
      const uint32_t st = base + (it & 1) * 32768;
      const uint64_t a_hi = make_smem_desc_sw128(st), a_lo = make_smem_desc_sw128(st + 16384);
      const uint64_t b_hi = make_smem_desc_sw128(st + 65536 + 0), b_lo = make_smem_desc_sw128(st + 65536 + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(k * 32 >> 4);
        if (variant == 0) {
          umma_ss<false>(tm, a_hi + adv, b_hi + adv, i128, 1u);
        } else if (variant == 1) {
          umma_ss<false>(tm, a_hi + adv, b_hi + adv, i256, 1u);
        } else if (variant == 2) {
          umma_ss<false>(tm + 128, a_lo + adv, b_hi + adv, i128, 1u);
          umma_ss<false>(tm + 128, a_hi + adv, b_lo + adv, i128, 1u);
          umma_ss<false>(tm, a_hi + adv, b_hi + adv, i128, 1u);
        } else if (variant == 3) {
          umma_ss<false>(tm, a_hi + adv, b_hi + adv, i256, 1u);
          umma_ss<false>(tm + 128, a_lo + adv, b_hi + adv, i128, 1u);
        } else {
          umma_ss<false>(tm + (k & 1) * 128, a_hi + adv, b_hi + adv, i128, 1u);
        }
      }
    }
    tcgen05_commit(&bar);
    mbar_wait(&bar, ph);
    long long t1 = clock64();
    unsigned long long g1 = gtimer();
    out_cycles[blockIdx.x] = t1 - t0;
    out_ns[blockIdx.x] = (long long)(g1 - g0);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 4096;
  int dev_sms = 0;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 130 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *dc, *dn;
  cudaMalloc(&dc, dev_sms * 8); cudaMalloc(&dn, dev_sms * 8);
  const char* names[] = {"N128", "N256", "3xN128 (2 acc)", "N256+N128 (concat)", "N128 alt acc"};
  const double flop_per_iter[] = {4 * 524288.0, 4 * 1048576.0, 12 * 524288.0, 4 * 1572864.0, 4 * 524288.0};
  const int mma_per_iter[] = {4, 4, 12, 8, 4};
  for (int grid : {1, dev_sms}) {
    for (int v = 0; v < 5; ++v) {
      for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128, smem>>>(v, iters, dc, dn);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> c(grid), n(grid);
      cudaMemcpy(c.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(n.data(), dn, grid * 8, cudaMemcpyDeviceToHost);
      double cyc = 0, ns = 0;
      for (int i = 0; i < grid; ++i) { cyc += c[i]; ns += n[i]; }
      cyc /= grid; ns /= grid;
      printf("grid %3d  %-20s cycles/MMA %7.1f  flop/cycle/SM %7.0f  MHz %6.0f  TFLOP/s(all CTAs) %8.1f\n", grid,
             names[v], cyc / ((double)iters * mma_per_iter[v]), flop_per_iter[v] * iters / cyc, cyc / ns * 1e3,
             flop_per_iter[v] * iters * grid / ns * 1e-3);
    }
  }
  return 0;
}
