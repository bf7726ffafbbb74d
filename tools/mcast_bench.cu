// Micro-benchmark: what limits the operand stream of the GEMM — the L2 slices (then TMA multicast inside a
// 4-CTA cluster delivers more bytes per clock into every SM) or the SM's own ingest path (then it does not)?
// Every CTA streams 32 KB "stages" from an L2-resident buffer through a 6-stage ring, one thread per CTA.
//   mode 0  unicast, every CTA reads its own rows                                   (what the GEMM does today)
//   mode 1  unicast, the two CTA pairs of a group of four read the SAME rows         (natural sharing through L2)
//   mode 2  preferred cluster 4: half of a stage unicast, half multicast to 2 CTAs   (B shared by two CTA pairs)
//   mode 3  preferred cluster 4: the whole stage multicast to 2 CTAs
//   mode 4  preferred cluster 4: the whole stage multicast to 4 CTAs
// Groups of four that the device launched as two regular 2-CTA clusters fall back to unicast.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mcast_bench mcast_bench.cu ../vit.triton_b200/csrc/tensormap.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../vit.triton_b200/csrc/common.cuh"
#include "../vit.triton_b200/csrc/tensormap.h"
using namespace vt;

#ifndef STOREROWS
#define STOREROWS 32
#endif
#ifndef STAGES
#define STAGES 6
#endif
constexpr int kStages = STAGES;
constexpr int kStageBytes = 32768;
#ifndef BOXROWS
#define BOXROWS 64
#endif
constexpr int kBoxRows = BOXROWS;          // rows x 128 B per TMA box
constexpr int kBoxBytes = kBoxRows * 128;
constexpr int kBoxesPerStage = kStageBytes / kBoxBytes;
// the tensor is [rows, kcols] bf16 (kcols = 64: boxes are contiguous 8 KB; kcols = 768: rows 1536 B apart like a
// GEMM operand); a box = 64 columns x 64 rows at (column block, row block)

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void spin_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_test_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_st, int mode, int iters, int total_rows, int kblocks, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kStages * kStageBytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kStages + s); };
  const uint32_t rank = cl_rank(), csz = cl_size();
  const int group = blockIdx.x >> 2;           // aligned group of four CTAs
  const int in_group = blockIdx.x & 3;

  // who writes into my stages / whose stages I write into (symmetric sets)
  uint16_t mask = static_cast<uint16_t>(1u << rank);
  int fan = 1;
  if (csz == 4) {
    if (mode == 2 || mode == 3) { mask = static_cast<uint16_t>((1u << (rank & 1)) | (1u << ((rank & 1) + 2))); fan = 2; }
    if (mode == 4) { mask = 0xF; fan = 4; }
  }
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile int*>(smem_raw + (base - smem_u32(smem_raw)) + kStages * kStageBytes + 8 * 2 * kStages + 8) = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), fan); }
    fence_barrier_init();
  }
  __syncthreads();
  cl_sync();
  const uint32_t n_boxes = static_cast<uint32_t>(total_rows / kBoxRows);     // kblocks is ignored: contiguous boxes
  const int warp = threadIdx.x >> 5;
  if (warp == 1 && mode != 5) {   // consumer warp (warp-uniform loop)
    for (int j = 0; j < iters; ++j) {
      const int s = j % kStages;
      spin_wait(full(s), (j / kStages) & 1);
      if (elect_one_sync()) {
        if (fan == 1) mbar_arrive(empty(s));
        else
          for (uint32_t c = 0; c < csz; ++c)
            if (mask & (1u << c)) remote_arrive(empty(s), c);
      }
      __syncwarp();
    }
  }
  if (warp == 2 && mode >= 5) {   // store warp: `kblocks` boxes of kStoreRows rows per iteration out of a fixed smem tile
    constexpr int kStoreRows = STOREROWS;
    const uint32_t n_sboxes = static_cast<uint32_t>(total_rows / kStoreRows);
    const long long t0 = clock64();
    uint32_t b = (blockIdx.x * 64u) % n_sboxes;
    for (int i = 0; i < iters; ++i) {
      if (elect_one_sync()) {
        for (int k = 0; k < kblocks; ++k) {
          tma_store_2d(&tmap_st, base, 0, static_cast<int>(b * kStoreRows));
          b += gridDim.x * 64u + 1u;
          while (b >= n_sboxes) b -= n_sboxes;
        }
        tma_store_commit();
        tma_store_wait_read<4>();
      }
      __syncwarp();
    }
    if (elect_one_sync()) tma_store_wait<0>();
    __syncwarp();
    const long long t1 = clock64();
    if (threadIdx.x == 64 && mode == 5) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = csz; }
  }
  if (mode == 7 && warp >= 2) {   // smem readers: conflict-free 16-byte loads, 512 B per warp instruction
    volatile int* flag = reinterpret_cast<volatile int*>(smem_raw + (base - smem_u32(smem_raw)) + kStages * kStageBytes + 8 * 2 * kStages + 8);
    const uint32_t a0 = base + (threadIdx.x & 31) * 16 + (warp - 2) * 8192;
    uint32_t acc = 0;
    long long n = 0;
    const long long t0 = clock64();
    while (*flag == 0) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        uint32_t x, y, z, w;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a0 + u * 512));
        acc += x ^ y ^ z ^ w;
      }
      n += 16;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) {
      out[296 + blockIdx.x * 4 + (warp - 2) * 2] = n * 512;
      out[296 + blockIdx.x * 4 + (warp - 2) * 2 + 1] = (t1 - t0) + (acc == 0x12345 ? 1 : 0);
    }
  }
  if (warp == 0 && mode != 5) {   // producer warp (warp-uniform loop, one elected lane issues)
    const long long t0 = clock64();
    // box index of (iteration i, owner CTA o, k) = (i * grid * BPS + o * BPS + k) mod n_boxes, kept incrementally
    const uint32_t step = (gridDim.x * kBoxesPerStage) % n_boxes;
    uint32_t it_base = 0;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      spin_wait(empty(s), ph ^ 1u);
      const uint32_t dst = base + s * kStageBytes;
      auto row_of = [&](uint32_t owner, uint32_t k) {
        uint32_t b = it_base + owner * kBoxesPerStage + k;
        while (b >= n_boxes) b -= n_boxes;
        return static_cast<int>(b * kBoxRows);
      };
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(full(s), kStageBytes);
        if (csz == 4 && mode == 2 && kBoxesPerStage == 4) {
          tma_load_2d(&tmap, full(s), dst, 0, row_of(blockIdx.x, 0), kEvictNormal);
          tma_load_2d(&tmap, full(s), dst + kBoxBytes, 0, row_of(blockIdx.x, 1), kEvictNormal);
          const uint32_t k = 2 + (rank >> 1);
          tma_load_2d_mc(&tmap, full(s), dst + k * kBoxBytes, 0, row_of(group * 4 + (rank & 1), k), mask);
        } else if (csz == 4 && mode == 3 && kBoxesPerStage == 4) {
          for (uint32_t h = 0; h < 2; ++h) {
            const uint32_t k = 2 * (rank >> 1) + h;
            tma_load_2d_mc(&tmap, full(s), dst + k * kBoxBytes, 0, row_of(group * 4 + (rank & 1), k), mask);
          }
        } else if (csz == 4 && mode == 4 && kBoxesPerStage == 4) {
          tma_load_2d_mc(&tmap, full(s), dst + rank * kBoxBytes, 0, row_of(group * 4, rank), mask);
        } else {
          const uint32_t owner = (mode == 1) ? (group * 4 + (in_group & 1)) : blockIdx.x;
#pragma unroll
          for (uint32_t k = 0; k < kBoxesPerStage; ++k)
            tma_load_2d(&tmap, full(s), dst + k * kBoxBytes, 0, row_of(owner, k), kEvictNormal);
        }
      }
      __syncwarp();
      it_base += step;
      if (it_base >= n_boxes) it_base -= n_boxes;
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
    // all my stages consumed = everything I asked for has arrived
    for (int s2 = 0; s2 < kStages; ++s2) {
      const int uses = (iters - s2 + kStages - 1) / kStages;
      if (uses > 0) spin_wait(empty(s2), (uses - 1) & 1);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = csz;
      *reinterpret_cast<volatile int*>(smem_raw + (base - smem_u32(smem_raw)) + kStages * kStageBytes + 8 * 2 * kStages + 8) = 1;
    }
  }
  __syncthreads();
  cl_sync();
}

// mode 8: the barrier protocol a 2-CTA GEMM with a shared operand would use inside a 4-CTA cluster.  CTA rank r = 2p + h
// (pair p, half h).  Per 32 KB stage a CTA receives a private 16 KB box (unicast, cta_group::2, completion credited to
// its pair LEADER's barrier) and two 8 KB halves of a shared 16 KB box, one issued by itself and one by rank r ^ 2,
// both multicast to {h, h + 2} (cta_group::2 + multicast: each destination's bytes must land on the leader barrier of the
// DESTINATION's pair).  A leader expects 64 KB per stage (its own and its peer's 32 KB); empties need both leaders.
constexpr uint32_t kPeerMaskB = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2cta_b(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerMaskB), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta_mc(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerMaskB), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__global__ void __launch_bounds__(128, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmap128, const __grid_constant__ CUtensorMap tmap64, int iters, int total_rows,
            int asym, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kStages * kStageBytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kStages + s); };
  const uint32_t rank = cl_rank(), csz = cl_size();
  const uint32_t h = rank & 1, p = (rank >> 1) & 1;
  const bool leader = h == 0;
  const int n_lead = csz == 4 ? 2 : 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), n_lead); }
    fence_barrier_init();
  }
  __syncthreads();
  cl_sync();
  const int warp = threadIdx.x >> 5;
  const uint32_t n128 = static_cast<uint32_t>(total_rows / 128);
  if (warp == 1 && leader) {   // consumer = the pair leader: frees the stage in every CTA that writes into this pair
    for (int j = 0; j < iters; ++j) {
      const int s = j % kStages;
      spin_wait(full(s), (j / kStages) & 1);
      if (elect_one_sync())
        for (uint32_t c = 0; c < csz; ++c) remote_arrive(empty(s), c);
      __syncwarp();
    }
  }
  if (warp == 0) {
    const long long t0 = clock64();
    int s = 0; uint32_t ph = 0;
    uint32_t b_priv = (blockIdx.x * 7u) % n128, b_sh = ((blockIdx.x >> 2) * 4u + h + 64u) % n128;
    for (int i = 0; i < iters; ++i) {
      spin_wait(empty(s), ph ^ 1u);
      const uint32_t dst = base + s * kStageBytes;
      if (elect_one_sync()) {
        if (leader) mbar_arrive_expect_tx(full(s), 2 * kStageBytes);
        tma_load_2d_2cta_b(&tmap128, full(s), dst, 0, static_cast<int>(b_priv * 128));
        if (csz == 4 && asym) {
          if (p == 1)
            for (uint32_t q = 0; q < 2; ++q)
              tma_load_2d_2cta_mc(&tmap64, full(s), dst + 16384 + q * 8192, 0, static_cast<int>(b_sh * 128 + q * 64),
                                  static_cast<uint16_t>((1u << h) | (1u << (h + 2))));
        } else if (csz == 4)
          tma_load_2d_2cta_mc(&tmap64, full(s), dst + 16384 + p * 8192, 0, static_cast<int>(b_sh * 128 + p * 64),
                              static_cast<uint16_t>((1u << h) | (1u << (h + 2))));
        else
          tma_load_2d_2cta_b(&tmap128, full(s), dst + 16384, 0, static_cast<int>(b_sh * 128));
      }
      __syncwarp();
      b_priv += 148u * 7u; while (b_priv >= n128) b_priv -= n128;
      b_sh += 37u * 4u + 1u; while (b_sh >= n128) b_sh -= n128;
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
    for (int s2 = 0; s2 < kStages; ++s2) {
      const int uses = (iters - s2 + kStages - 1) / kStages;
      if (uses > 0) spin_wait(empty(s2), (uses - 1) & 1);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = csz; }
  }
  __syncthreads();
  cl_sync();
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  const int kblocks = argc > 2 ? atoi(argv[2]) : 1;   // 64-column blocks per row (12 = a K = 768 operand)
  const int mbytes = argc > 3 ? atoi(argv[3]) : 32;   // buffer size: L2 resident
  const int total_rows = mbytes * 1024 * 1024 / (128 * kblocks) / kBoxRows * kBoxRows;
  void* buf;
  cudaMalloc(&buf, static_cast<size_t>(total_rows) * 128 * kblocks);
  cudaMemset(buf, 1, static_cast<size_t>(total_rows) * 128 * kblocks);
  CUtensorMap tmap;
  if (make_tmap_bf16_2d(&tmap, buf, 64 * kblocks, total_rows, 64 * kblocks, 64, kBoxRows, TMAP_SW_128)) { printf("tmap failed\n"); return 1; }
  void* buf2;
  cudaMalloc(&buf2, static_cast<size_t>(total_rows) * 128);
  CUtensorMap tmap_st;
  if (make_tmap_bf16_2d(&tmap_st, buf2, 64, total_rows, 64, 64, STOREROWS, TMAP_SW_128)) { printf("tmap failed\n"); return 1; }
  long long* out;
  cudaMalloc(&out, 148 * 6 * sizeof(long long));
  const int smem = 1024 + kStages * kStageBytes + 8 * 2 * kStages + 64;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int dev_clock_khz = 0;
  cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
  for (int mode = 0; mode <= 7; ++mode) {
    if (mode >= 2 && mode <= 4 && kBoxesPerStage != 4) continue;
    for (int rep = 0; rep < 3; ++rep) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attrs[2];
      attrs[0].id = cudaLaunchAttributeClusterDimension;
      attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
      attrs[1].id = cudaLaunchAttributePreferredClusterDimension;
      attrs[1].val.preferredClusterDim.x = 4; attrs[1].val.preferredClusterDim.y = 1; attrs[1].val.preferredClusterDim.z = 1;
      cfg.attrs = attrs;
      cfg.numAttrs = mode >= 2 ? 2 : 1;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cudaError_t rc = cudaLaunchKernelEx(&cfg, stream_kernel, tmap, tmap_st, mode, iters, total_rows, kblocks, out);
      cudaEventRecord(e1);
      cudaError_t rc2 = cudaDeviceSynchronize();
      if (rc != cudaSuccess || rc2 != cudaSuccess) {
        printf("mode %d: launch %s / sync %s\n", mode, cudaGetErrorString(rc), cudaGetErrorString(rc2));
        return 1;
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> h(148 * 6);
      cudaMemcpy(h.data(), out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      long long cmin = 1LL << 62, cmax = 0; double csum = 0; int n4 = 0;
      for (int b = 0; b < 148; ++b) { cmin = std::min(cmin, h[2 * b]); cmax = std::max(cmax, h[2 * b]); csum += h[2 * b]; n4 += h[2 * b + 1] == 4; }
      double bytes = static_cast<double>(iters) * kStageBytes;
      if (mode == 5) bytes = static_cast<double>(iters) * kblocks * STOREROWS * 128;
      if (mode >= 5 && rep == 2) printf("  (mode %d: %s, %d store boxes of %d rows per iteration)\n", mode, mode == 5 ? "stores only, B/clk = stored" : "loads + stores, B/clk = loaded", kblocks, STOREROWS);
      if (rep == 2 && mode == 7) {
        double rb = 0, rc = 0;
        for (int b = 0; b < 148; ++b) { rb += h[296 + 4 * b] + h[296 + 4 * b + 2]; rc += 0.5 * (h[296 + 4 * b + 1] + h[296 + 4 * b + 3]); }
        printf("  (mode 7: loads + 2 warps of ld.shared.v4: smem read %.1f B/clk/SM on top of the loads below)\n", rb / rc);
      }
      if (rep == 2)
        printf("mode %d: %d CTAs in 4-clusters | received B/clk/SM avg %.1f (fastest %.1f slowest %.1f) | chip %.2f TB/s received | %.3f ms\n",
               mode, n4, bytes / (csum / 148), bytes / cmin, bytes / cmax, bytes * 148 / (ms * 1e-3) / 1e12, ms);
    }
  }
  {   // mode 8
    CUtensorMap t128, t64;
    if (make_tmap_bf16_2d(&t128, buf, 64, total_rows, 64, 64, 128, TMAP_SW_128) ||
        make_tmap_bf16_2d(&t64, buf, 64, total_rows, 64, 64, 64, TMAP_SW_128)) { printf("tmap failed\n"); return 1; }
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int pref = 0; pref < 3; ++pref) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attrs[2];
      attrs[0].id = cudaLaunchAttributeClusterDimension;
      attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
      attrs[1].id = cudaLaunchAttributePreferredClusterDimension;
      attrs[1].val.preferredClusterDim.x = 4; attrs[1].val.preferredClusterDim.y = 1; attrs[1].val.preferredClusterDim.z = 1;
      cfg.attrs = attrs; cfg.numAttrs = pref ? 2 : 1;
      const int asym = pref == 2;
      for (int rep = 0; rep < 2; ++rep) {
        cudaError_t rc = cudaLaunchKernelEx(&cfg, pair_kernel, t128, t64, iters, total_rows, asym, out);
        cudaError_t rc2 = cudaDeviceSynchronize();
        if (rc != cudaSuccess || rc2 != cudaSuccess) { printf("mode 8: launch %s / sync %s\n", cudaGetErrorString(rc), cudaGetErrorString(rc2)); return 1; }
      }
      std::vector<long long> h(148 * 2);
      cudaMemcpy(h.data(), out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      double c4 = 0, c2 = 0; int n4 = 0, n2 = 0;
      for (int b = 0; b < 148; ++b) { if (h[2 * b + 1] == 4) { c4 += h[2 * b]; ++n4; } else { c2 += h[2 * b]; ++n2; } }
      const double bytes = static_cast<double>(iters) * kStageBytes;
      printf("mode 8 (pair-leader barriers, %s): %d CTAs in 4-clusters %.1f B/clk/SM received, %d CTAs in 2-clusters %.1f B/clk/SM\n",
             pref == 2 ? "preferred cluster 4, BOTH shared halves issued by pair 1 only" : pref ? "preferred cluster 4, shared half multicast" : "clusters of 2, unicast", n4, n4 ? bytes / (c4 / n4) : 0.0, n2,
             n2 ? bytes / (c2 / n2) : 0.0);
    }
  }
  return 0;
}
