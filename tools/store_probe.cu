// Micro-probe: throughput of TMA tensor stores of [32 rows][W bytes] boxes (what a GEMM epilogue warp
// produces) for W = 32, 64, 128 (128B-swizzled), issued by 1 or 8 warps of one CTA per SM, against
// plain st.global (row-per-lane 16-byte stores and coalesced 128-byte rows).
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace sslam::tc;

// mode 0..2: TMA boxes of 32 x {32,64,128} bytes; 3: st.global.v4 lane=row (16 B per lane, 2 per 32 B row);
// 4: st.global.v4 coalesced (8 lanes per 128-byte row)
__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ CUtensorMap tm32,
                                                const __grid_constant__ CUtensorMap tm64,
                                                const __grid_constant__ CUtensorMap tm128, unsigned char* gbuf,
                                                size_t row_pitch, int mode, int nwarps, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps) return;
  unsigned char* slab = smem + warp * 8192;                       // two 4 KB buffers per warp
  for (int i = lane; i < 2048; i += 32) reinterpret_cast<uint32_t*>(slab)[i] = i;
  fence_proxy_async();
  __syncwarp();
  const int IT = 512;
  const int W = mode == 0 ? 32 : mode == 1 ? 64 : 128;
  const int row_base = (blockIdx.x * nwarps + warp) * 32;
  long long t0 = clock64();
  for (int i = 0; i < IT; ++i) {
    const int colb = (i * W) % 2048;                              // byte column inside a 2 KB-wide row
    if (mode <= 2) {
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      unsigned char* buf = slab + (i & 1) * 4096;
      *reinterpret_cast<uint4*>(buf + lane * 16) = make_uint4(i, i, i, i);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const CUtensorMap* m = mode == 0 ? &tm32 : mode == 1 ? &tm64 : &tm128;
        tma_store_2d(m, buf, colb / 2, row_base);
        tma_store_commit();
      }
    } else if (mode == 3) {
      unsigned char* g = gbuf + (size_t)(row_base + lane) * row_pitch + (i * 32) % 2048;
      *reinterpret_cast<uint4*>(g) = make_uint4(i, i, i, i);
      *reinterpret_cast<uint4*>(g + 16) = make_uint4(i, i, i, i);
    } else if (mode == 5) {
      unsigned char* g = gbuf + (size_t)(row_base + lane) * row_pitch + (i * 32) % 2048;
      asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(g), "r"(i) : "memory");
    } else {
      // 4 instructions x (4 rows x 128 B) = 32 rows x 64 B ... use 8 instr for 32 rows x 128 B
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        unsigned char* g = gbuf + (size_t)(row_base + 4 * k + (lane >> 3)) * row_pitch + (i * 128) % 2048 + (lane & 7) * 16;
        *reinterpret_cast<uint4*>(g) = make_uint4(i, i, i, i);
      }
    }
  }
  if (mode <= 2 && lane == 0) tma_store_wait_all<0>();
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 8 + warp] = (t1 - t0);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t rows = (size_t)sms * 8 * 32, pitch = 2048;
  unsigned char* gbuf; long long* dout;
  cudaMalloc(&gbuf, rows * pitch); cudaMalloc(&dout, sms * 8 * 8);
  CUtensorMap t32, t64, t128;
  if (make_tensor_map_2d(&t32, gbuf, rows, 1024, 32, 16, 2, 0) || make_tensor_map_2d(&t64, gbuf, rows, 1024, 32, 32, 2, 0) ||
      make_tensor_map_2d(&t128, gbuf, rows, 1024, 32, 64, 2, 128)) { printf("tensor map failed\n"); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const char* names[] = {"TMA box 32x32B", "TMA box 32x64B", "TMA box 32x128B (sw128)", "st.global lane=row 32B", "st.global coalesced 32x128B", "st.global.v8 lane=row 32B"};
  const int bytes[] = {1024, 2048, 4096, 1024, 4096, 1024};
  for (int nw : {1, 8}) {
    for (int mode = 0; mode < 6; ++mode) {
      for (int rep = 0; rep < 2; ++rep) probe<<<sms, 256, 80 * 1024>>>(t32, t64, t128, gbuf, pitch, mode, nw, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      static long long h[148 * 8];
      cudaMemcpy(h, dout, sms * 8 * 8, cudaMemcpyDeviceToHost);
      double c = 0; int n = 0;
      for (int b = 0; b < sms; ++b) for (int w = 0; w < nw; ++w) { c += h[b * 8 + w]; ++n; }
      c /= n;
      printf("warps/SM %d  %-30s %7.1f cycles per box per warp   %6.2f B/cycle/SM\n", nw, names[mode], c / 512,
             bytes[mode] * 512.0 * nw / c);
    }
  }
  return 0;
}
