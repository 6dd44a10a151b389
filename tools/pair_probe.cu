// Micro-probe: issue rate of tcgen05.mma.cta_group::2 (kind::f16, SS operands, M = 256 over a CTA pair)
// for the accumulator patterns of the f16x3 product  x.w = hi.hi' + 2^-11 (hi.lo' + lo.hi').
// One 2-CTA cluster per TPC; thread 0 of the leader issues `iters` k-blocks (4 k-steps of 16) on zeroed
// shared memory (no TMA traffic) and waits for the commit.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../semantic-slam-master_b200/csrc pair_probe.cu -o pair_probe
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

// variants (S = cross-term accumulator, D = main accumulator, all M = 256):
//  0  S S D      three N128 per k-step, two accumulators (library today)
//  1  S1 S2 D    three N128, three accumulators
//  2  [D|S] + S2 one N256 (A_hi x [B_hi;B_lo]) + one N128 (A_lo x B_hi), three accumulators
//  3  D          N128, one accumulator (dependent chain)
//  4  D          N256, one accumulator (dependent chain)
//  5  S D S      three N128, two accumulators, other order
//  6  S1 S2 D with a fresh set of three accumulators every k-block (no cross-k-block dependency either)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(int variant, int iters, long long* out_cycles, long long* out_ns) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc_pair(&slot, 512);
  fence_proxy_async();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t i128 = make_instr_desc(FMT_F16, 256, 128), i256 = make_instr_desc(FMT_F16, 256, 256);
    const uint32_t base = smem_u32(smem);
    long long t0 = clock64();
    unsigned long long g0 = gtimer();
    for (int it = 0; it < iters; ++it) {
      // A stage: hi 16 KB | lo 16 KB (128 rows x 128 B each), two stages; B: this CTA's half of the rows:
      // hi 8 KB | lo 8 KB per k-block (64 rows each; [hi;lo] = 128 contiguous rows for the N256 form)
      const uint32_t st = base + (it & 1) * 32768;
      const uint32_t bb = base + 65536 + (it % 6) * 16384;
      const uint64_t a_hi = make_smem_desc_sw128(st), a_lo = make_smem_desc_sw128(st + 16384);
      const uint64_t b_hi = make_smem_desc_sw128(bb), b_lo = make_smem_desc_sw128(bb + 8192);
      const uint32_t off = variant == 6 ? 0u : 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(k * 32 >> 4);
        switch (variant) {
          case 0:
            umma_ss_pair(tm + 128, a_lo + adv, b_hi + adv, i128, 1u);
            umma_ss_pair(tm + 128, a_hi + adv, b_lo + adv, i128, 1u);
            umma_ss_pair(tm, a_hi + adv, b_hi + adv, i128, 1u);
            break;
          case 1:
          case 6:
            umma_ss_pair(tm + 128 + off, a_lo + adv, b_hi + adv, i128, 1u);
            umma_ss_pair(tm + 256 + off, a_hi + adv, b_lo + adv, i128, 1u);
            umma_ss_pair(tm + off, a_hi + adv, b_hi + adv, i128, 1u);
            break;
          case 2:
            umma_ss_pair(tm, a_hi + adv, b_hi + adv, i256, 1u);
            umma_ss_pair(tm + 256, a_lo + adv, b_hi + adv, i128, 1u);
            break;
          case 3:
            umma_ss_pair(tm, a_hi + adv, b_hi + adv, i128, 1u);
            break;
          case 4:
            umma_ss_pair(tm, a_hi + adv, b_hi + adv, i256, 1u);
            break;
          default:
            umma_ss_pair(tm + 128, a_lo + adv, b_hi + adv, i128, 1u);
            umma_ss_pair(tm, a_hi + adv, b_hi + adv, i128, 1u);
            umma_ss_pair(tm + 128, a_hi + adv, b_lo + adv, i128, 1u);
            break;
        }
      }
    }
    tcgen05_commit_pair(&bar, (uint16_t)1);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    unsigned long long g1 = gtimer();
    out_cycles[blockIdx.x >> 1] = t1 - t0;
    out_ns[blockIdx.x >> 1] = (long long)(g1 - g0);
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_pair(tm, 512);
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 4096;
  int dev_sms = 0;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 170 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *dc, *dn;
  cudaMalloc(&dc, dev_sms * 8); cudaMalloc(&dn, dev_sms * 8);
  const char* names[] = {"S S D (2 acc)", "S1 S2 D (3 acc)", "N256[D|S] + N128", "N128 chain", "N256 chain", "S D S (2 acc)", "S1 S2 D (again)"};
  const double M128 = 2.0 * 256 * 128 * 16, M256 = 2.0 * 256 * 256 * 16;
  const double flop_per_iter[] = {12 * M128, 12 * M128, 4 * (M256 + M128), 4 * M128, 4 * M256, 12 * M128, 12 * M128};
  const int mma_per_iter[] = {12, 12, 8, 4, 4, 12, 12};
  for (int grid : {2, dev_sms}) {
    for (int v = 0; v < 7; ++v) {
      for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128, smem>>>(v, iters, dc, dn);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      const int np = grid / 2;
      std::vector<long long> c(np), n(np);
      cudaMemcpy(c.data(), dc, np * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(n.data(), dn, np * 8, cudaMemcpyDeviceToHost);
      double cyc = 0, ns = 0;
      for (int i = 0; i < np; ++i) { cyc += c[i]; ns += n[i]; }
      cyc /= np; ns /= np;
      printf("grid %3d  %-20s cycles/MMA %7.1f  cycles/k-step %7.1f (ideal 192)  flop/cycle/SM %7.0f  MHz %6.0f  TFLOP/s(all) %8.1f\n", grid,
             names[v], cyc / ((double)iters * mma_per_iter[v]), cyc / ((double)iters * 4), flop_per_iter[v] * iters / cyc / 2, cyc / ns * 1e3,
             flop_per_iter[v] * iters * np / ns * 1e-3);
    }
  }
  return 0;
}
