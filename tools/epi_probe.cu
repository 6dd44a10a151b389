// Micro-probe of the epilogue building blocks on sm_100a (one warp, idle SM): latency of
// tcgen05.ld (+wait), fence.proxy.async, and the smem -> TMA-store round trip.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../semantic-slam-master_b200/csrc
//        epi_probe.cu ../semantic-slam-master_b200/csrc/match_tc.o ../semantic-slam-master_b200/csrc/api.o -o epi_probe
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace sslam::tc;

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmO, long long* out, float* sink) {
  __shared__ uint32_t slot;
  __shared__ __align__(128) unsigned char stg[2][2048];
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const int IT = 256;
    float acc = 0.f;
    long long t0, t1;
    uint32_t r[16], r2[16], r32[32];
    // 1. x16 + wait
    t0 = clock64();
    for (int i = 0; i < IT; ++i) {
      tmem_ld_32x16(tm + ((i & 7) << 4), r);
      tmem_ld_wait();
      acc += __uint_as_float(r[i & 15]);
    }
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / IT;
    // 2. two x16 then one wait
    t0 = clock64();
    for (int i = 0; i < IT; ++i) {
      tmem_ld_32x16(tm + ((i & 7) << 4), r);
      tmem_ld_32x16(tm + 128 + ((i & 7) << 4), r2);
      tmem_ld_wait();
      acc += __uint_as_float(r[i & 15]) + __uint_as_float(r2[i & 15]);
    }
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / IT;
    // 3. x32 + wait
    t0 = clock64();
    for (int i = 0; i < IT; ++i) {
      tmem_ld_32x32(tm + ((i & 3) << 5), r32);
      tmem_ld_wait();
      acc += __uint_as_float(r32[i & 31]);
    }
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / IT;
    // 4. fence.proxy.async
    t0 = clock64();
    for (int i = 0; i < IT; ++i) {
      stg[0][lane * 4] = (unsigned char)i;
      fence_proxy_async();
    }
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / IT;
    // 5. STS + fence + syncwarp + TMA store (2 KB box) + commit + wait_read<0>
    t0 = clock64();
    for (int i = 0; i < IT; ++i) {
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      float4* d = reinterpret_cast<float4*>(stg[0] + lane * 64);
      d[0] = d[1] = d[2] = d[3] = make_float4(acc, 1.f, 2.f, 3.f);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmO, stg[0], 0, 32 * (i & 15)); tma_store_commit(); }
    }
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / IT;
    if (lane == 0) tma_store_wait_all<0>();
    // 6. same, two staging buffers, wait_read<1>
    t0 = clock64();
    for (int i = 0; i < IT; ++i) {
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      float4* d = reinterpret_cast<float4*>(stg[i & 1] + lane * 64);
      d[0] = d[1] = d[2] = d[3] = make_float4(acc, 1.f, 2.f, 3.f);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmO, stg[i & 1], 0, 32 * (i & 15)); tma_store_commit(); }
    }
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / IT;
    if (lane == 0) tma_store_wait_all<0>();
    // 7. cvt f32->f16 x2 + back, 16 elements (split_f16 of a unit)
    t0 = clock64();
    for (int i = 0; i < IT; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        __half h, l;
        split_f16(acc + (float)j, h, l);
        s += __half2float(h) + __half2float(l);
      }
      acc = s * 1e-3f;
    }
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / IT;
    // 8. shuffle chain of 5 dependent stages
    t0 = clock64();
    for (int i = 0; i < IT; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc = fmaxf(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    }
    t1 = clock64();
    if (lane == 0) out[7] = (t1 - t0) / IT;
    sink[lane] = acc;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  float* obuf; long long* dout; float* sink;
  cudaMalloc(&obuf, 512 * 16 * 4); cudaMalloc(&dout, 64); cudaMalloc(&sink, 128);
  CUtensorMap tm;
  if (make_tensor_map_2d(&tm, obuf, 512, 16, 32, 16, 4, 0)) { printf("tensor map failed\n"); return 1; }
  probe<<<1, 128>>>(tm, dout, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[8];
  cudaMemcpy(h, dout, 64, cudaMemcpyDeviceToHost);
  const char* n[] = {"tcgen05.ld x16 + wait", "2 x (ld x16) + wait", "ld x32 + wait", "STS + fence.proxy.async",
                     "stage+fence+TMA store+wait_read<0>", "same, 2 buffers, wait_read<1>", "split_f16 x16 (+back)",
                     "5-stage shuffle max"};
  for (int i = 0; i < 8; ++i) printf("%-40s %6lld cycles\n", n[i], h[i]);
  return 0;
}
