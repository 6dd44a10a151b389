// Blackwell (sm_100a) building blocks written as raw PTX: mbarrier, TMA tensor loads, TMEM
// allocation, tcgen05.mma / commit / ld, and the UMMA shared-memory / instruction descriptors.
// No CUTLASS/CuTe: the encodings below follow the PTX ISA for tcgen05 (SASS: UTCHMMA/UTCQMMA,
// LDTM, UTMALDG, SYNCS).
#pragma once

#include <cuda.h>   // CUtensorMap
#include <cuda_fp16.h>
// (types only; the driver entry point is fetched at run time)

#include "common.cuh"

namespace sslam {
namespace tc {

// ------------------------------------------------------------------------------------ host side
// Row-major 2-D tensor map [rows, cols] with a {box_cols, box_rows} box and 128-byte swizzle.
// elem_bytes 4 (fp32/tf32) or 2 (bf16).  Returns SSLAM_OK or an error code.
// swizzle_bytes: 128 (default), 64 or 0 (dense box).
int make_tensor_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                       uint32_t box_rows, uint32_t box_cols, int elem_bytes, int swizzle_bytes = 128);

// NHWC [B,H,W,C] tensor of 2-byte elements as a 4-D map, box {box_c, box_w, box_h, 1}, 128-byte swizzle
// (box_c * 2 == 128); out-of-range coordinates are zero-filled.
int make_tensor_map_nhwc(CUtensorMap* map, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C,
                         uint32_t box_h, uint32_t box_w, uint32_t box_c);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------ device side
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps after ~2 s (launch failure reported to the host) instead of
// hanging the GPU.
// Debug aid: when set (tools only), a timed-out wait records {block, thread, barrier address, parity}
// in this host-mapped buffer before trapping.  One copy per translation unit.
static __device__ unsigned long long* g_watchdog_buf = nullptr;

// cold path of mbar_wait, out of line (it used to be inlined at every wait: ~40 instructions per site)
static __device__ __noinline__ void mbar_wait_timeout_check(uint32_t addr, uint32_t parity, unsigned long long& t0) {
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (t0 == 0) { t0 = now; return; }
  if (now - t0 > (threadIdx.x < 64 ? 1500000000ull : 2000000000ull)) {   // control warps report first
    if (g_watchdog_buf) {
      const bool ctl = threadIdx.x < 64;               // producer / MMA warps first
      const unsigned long long slot = atomicAdd(g_watchdog_buf + (ctl ? 1 : 0), 1ull);
      if (slot < (ctl ? 30ull : 32ull)) {
        uint32_t crank;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
        g_watchdog_buf[(ctl ? 2 : 32) + slot] = ((unsigned long long)blockIdx.x << 48) | ((unsigned long long)threadIdx.x << 36) |
                                   ((unsigned long long)(crank & 15) << 32) | ((unsigned long long)(addr & 0xffffff) << 4) |
                                   (parity & 1);
      }
      __threadfence_system();
    }
    __trap();
  }
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    // the suspend-time hint lets the hardware park the thread instead of spinning on issue slots
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(100000u)
        : "memory");
    if (!ok && (spin & 63u) == 63u) mbar_wait_timeout_check(addr, parity, t0);
  }
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// TMA: global (tensor map) -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 4-D tile load (NHWC maps): coordinates (c, x, y, b), signed — out-of-range parts are zero-filled
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store: shared (dense box) -> global (tensor map); completion tracked by bulk async-groups.
// Out-of-range rows / columns of the box are clipped by the hardware.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N groups have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// TMA prefetch of one box into L2 (no smem destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- descriptors
// K-major operand tile in the canonical SWIZZLE_128B layout: rows of 128 bytes, groups of 8 rows
// (1024 bytes) stored back to back.  Exactly what a TMA box {128 bytes, R rows} with
// CU_TENSOR_MAP_SWIZZLE_128B writes.  Fields (PTX "shared memory matrix descriptor"):
//   [0,14)  start address >> 4          [16,30) leading-dim byte offset >> 4 (unused here: 1)
//   [32,46) stride-dim byte offset >> 4 (8 rows * 128 B = 1024)     [46,48) version = 1
//   [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Same for the SWIZZLE_64B layout: rows of 64 bytes, groups of 8 rows (512 bytes) back to back
// (a TMA box {64 bytes, R rows} with CU_TENSOR_MAP_SWIZZLE_64B); layout type 4.
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// K-major operand tile WITHOUT swizzle ("interleaved" canonical layout): 8-row x 16-byte core matrices of
// 128 contiguous bytes; `lbo` = byte distance between the two core matrices of a 32-byte k-step (next 8
// elements of K), `sbo` = byte distance between consecutive 8-row groups.  Layout type 0.
__device__ __forceinline__ uint64_t make_smem_desc_interleave(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor (32 bit) for kind::f16 / kind::tf32, dense, fp32 accumulate, both
// operands K-major:  [4,6) D format 1=f32   [7,10) A format   [10,13) B format
//   (0 f16, 1 bf16, 2 tf32)   [15] A major 0=K   [16] B major 0=K   [17,23) N>>3   [24,29) M>>4
constexpr uint32_t FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2;
__host__ __device__ constexpr uint32_t make_instr_desc(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA
template <bool TF32>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster, on the two SMs of a TPC, execute one MMA with
// M = 256 (each CTA owns 128 rows of A and of the accumulator) and share B (each holds N/2 rows).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope): a cluster-scope release would drain every outstanding
  // memory operation of the thread (ERRBAR) — the data this barrier protects lives in TMEM and is
  // ordered by tcgen05.fence::before_thread_sync
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose byte count is signalled on a barrier that may live in
// the peer CTA (`bar_cluster_addr` is a shared::cluster address, e.g. from mapa_u32)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// Same, multicast: the box lands at the same shared-memory offset in every CTA of `cta_mask` and the
// bytes are signalled on the barrier at `bar_cluster_addr`'s offset in the pair leader (even rank) of
// each destination CTA (bar_cluster_addr must name an even-ranked CTA).
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                    int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
        "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {   // one warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// D[tmem, both CTAs] (+)= A * B^T with M = 256; issued by one thread of the leader CTA (rank 0).  The
// descriptors name the same shared-memory offsets in both CTAs.
__device__ __forceinline__ void umma_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// round-to-nearest fp32 -> tf32 (low 13 mantissa bits cleared)
__device__ __forceinline__ float to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// fp32 -> (fp16 hi, fp16 lo) with x ~= hi + lo * 2^-11 (22 significant bits)
constexpr float F16_LO_SCALE = 2048.0f, F16_LO_INV = 1.0f / 2048.0f;
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(__fmul_rn(__fsub_rn(x, __half2float(hi)), F16_LO_SCALE));
}
__device__ __forceinline__ float join_f16(__half hi, __half lo) {
  return __fmaf_rn(__half2float(lo), F16_LO_INV, __half2float(hi));
}
// ---- two fp32 operations per instruction (sm_100: FFMA2 / FADD2 / FMUL2 on aligned register pairs) and
// the mixed fp16 (+) fp32 forms (FHADD, FHFMA): each lane / operation is the IEEE round-to-nearest one,
// so results are bit-identical to the scalar sequences they replace
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// fp32(h) + c
__device__ __forceinline__ float fhadd(uint16_t h, float c) {
  float r;
  asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(r) : "h"(h), "f"(c));
  return r;
}
// fp32(a) * fp32(b) + c, one rounding
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float c) {
  float r;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(a), "h"(b), "f"(c));
  return r;
}
constexpr uint16_t F16_LO_INV_BITS = 0x1000;   // 2^-11 as fp16
#endif  // __CUDACC__

}  // namespace tc
}  // namespace sslam
