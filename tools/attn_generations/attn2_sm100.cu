// K3 (main variant): persistent, software-pipelined fused attention forward on tcgen05.
//
//   ctx[b, i, h*64:(h+1)*64] = softmax_j( scale * q[b,i,h] . k[b,j,h] ) @ v[b,j,h]      (bf16, dh = 64)
//
// Replaces the reference's per-head matmul3 -> softmax -> matmul3 -> slice-assign chain
// (vit/vit.py:60-72,101-108) for all heads at once; the score matrix never leaves the SM.
//
// One CTA per SM hosts TWO independent pipelines ("slots").  Each slot walks over its own list of
// work items (image, head, 128-query tile) and owns a Q/K/V smem buffer set, a 256-column TMEM
// region, a TMA-producer warp, an MMA-issuer warp and a softmax warpgroup, so while one slot's
// softmax occupies the MUFU/FMA pipes the other slot's MMAs occupy the tensor pipe; nothing but the
// hardware arbitrates between them.  (The v1 kernel in attn_sm100.cu ran one tile per CTA and spent
// most of its time in prologue / load latency: 157 us per layer at C2 against ~45 us of work.)
//
//   warps 0-3 / 4-7   softmax warpgroup of slot 0 / 1 (thread per row; running max, sum and fp32
//                     output accumulator in registers; P written back over S in TMEM as bf16x2)
//   warps 8 / 9       TMA producer of slot 0 / 1: Q per item, K_j and V_j per KV block (warp 8 also
//                     allocates TMEM)
//   warps 10 / 11     MMA issuer of slot 0 / 1:  S = Q K_j^T (SS form),  O_j = P V_j (TS form, V as
//                     an MN-major SWIZZLE_128B operand straight from the fused-QKV buffer)
// K is released as soon as S is computed and V as soon as O_j is, so the next block's / item's
// loads fly during the softmax.  The normalised output goes through a swizzled staging tile and a
// TMA store (3-D map: rows >= N of a ragged last tile are clipped).
// TMEM per slot: S fp32 [0, nj) ; P bf16x2 [0, nj/2) ; O_j fp32 [128, 192).
#include "common.cuh"
#include "tensormap.h"

namespace vt {

namespace {

constexpr int kDH = 64;
constexpr int kQTile = 128;
constexpr int kThreads2 = 384;
constexpr int kQBytes = kQTile * kDH * 2;       // 16 KB
constexpr int kOutBytes = kQTile * kDH * 2;     // 16 KB staging per slot
constexpr int kOCol = 128;
constexpr int kSlotCols = 256;
constexpr int kSmemLimit = 232448;

struct Attn2Params {
  int N, H;
  int nqt;            // query tiles per (image, head)
  int bkv, nblk;      // rows per KV block (multiple of 16, <= 256), number of KV blocks
  long long total_items;
  int reverse;        // walk the items from the last to the first (L2 reuse, see api.cu)
  float scale_log2;
  long long* dbg;   // optional cycle counters (developer builds)
};

// per-slot barriers
enum { A_QFULL = 0, A_QEMPTY, A_KFULL, A_KEMPTY, A_VFULL, A_VEMPTY, A_SFULL, A_PFULL, A_OFULL, A_OREAD,
       A_TURN, A_PER_SLOT };
constexpr int A_NBARS = 2 * A_PER_SLOT;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Softmax of one S block held in TMEM: two passes (row max, then exp / sum / P write-back), each
// software-pipelined so the next 32-column TMEM load is in flight while the current one is consumed
// (tcgen05.wait::ld waits for everything outstanding, so the wait sits after the math).
// Returns the block's row sum; m_new_out is the running max including this block (log2 domain).
__device__ __forceinline__ float softmax_block_pipelined(uint32_t t_lane, int nvalid, int nj,
                                                         float scale_log2, float m_run,
                                                         float& m_new_out, uint32_t turn_wait,
                                                         uint32_t turn_parity, uint32_t turn_pass,
                                                         long long* tp = nullptr) {
  const long long tp0 = tp ? clock64() : 0;
  const int nfull = nvalid & ~31;   // columns covered by fully valid 32-wide chunks
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
  uint32_t ra[32], rb[32];
  auto max_chunk = [&](const uint32_t (&r)[32]) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      mx0 = fmax3(mx0, __uint_as_float(r[i + 0]), __uint_as_float(r[i + 1]));
      mx1 = fmax3(mx1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      mx2 = fmax3(mx2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
      mx3 = fmax3(mx3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
    }
  };
  {
    // pass 1 has nothing else live: keep four 32-column loads (128 registers) in flight per wait
    int c = 0;
    for (; c + 128 <= nfull; c += 128) {
      uint32_t rc[32], rd[32];
      tmem_ld_32x32(t_lane + c, ra);
      tmem_ld_32x32(t_lane + c + 32, rb);
      tmem_ld_32x32(t_lane + c + 64, rc);
      tmem_ld_32x32(t_lane + c + 96, rd);
      tmem_ld_wait();
      max_chunk(ra); max_chunk(rb); max_chunk(rc); max_chunk(rd);
    }
    if (c + 64 <= nfull) {
      tmem_ld_32x32(t_lane + c, ra);
      tmem_ld_32x32(t_lane + c + 32, rb);
      tmem_ld_wait();
      max_chunk(ra); max_chunk(rb);
      c += 64;
    }
    if (c + 32 <= nfull) {
      tmem_ld_32x32(t_lane + c, ra);
      tmem_ld_wait();
      max_chunk(ra);
    }
  }
  for (int c = nfull; c < nj; c += 16) {
    uint32_t r[16];
    tmem_ld_32x16(t_lane + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c + i < nvalid) mx0 = fmaxf(mx0, __uint_as_float(r[i]));
  }
  const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2);
  m_new_out = m_new;
  if (tp) tp[0] += clock64() - tp0;
  // The exp pass saturates the MUFU pipe: the two slots take turns so that one slot's exp pass runs
  // while the other does its MUFU-free work (row max, O read, epilogue, MMA waits).
  if (turn_wait) mbar_wait(turn_wait, turn_parity);
  const long long tp1 = tp ? clock64() : 0;

  float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
  auto exp_chunk = [&](const uint32_t (&r)[32], int c) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float p0 = ex2_approx(fmaf(__uint_as_float(r[i + 0]), scale_log2, -m_new));
      const float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m_new));
      const float p2 = ex2_approx(fmaf(__uint_as_float(r[i + 2]), scale_log2, -m_new));
      const float p3 = ex2_approx(fmaf(__uint_as_float(r[i + 3]), scale_log2, -m_new));
      ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
      pk[(i >> 1) + 0] = pack_bf16x2(p0, p1);
      pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
    }
    tmem_st_32x16(t_lane + (c >> 1), pk);
  };
  if (nfull > 0) {
    tmem_ld_32x32(t_lane, ra);
    tmem_ld_wait();
  }
  for (int c = 0; c < nfull; c += 64) {
    // P for columns [c, c+32) overwrites S columns [c/2, c/2+16): always left of (or inside) data
    // this thread has already pulled into registers, including the prefetched chunk c+32.
    if (c + 32 < nfull) tmem_ld_32x32(t_lane + c + 32, rb);
    exp_chunk(ra, c);
    tmem_ld_wait();
    if (c + 32 < nfull) {
      if (c + 64 < nfull) tmem_ld_32x32(t_lane + c + 64, ra);
      exp_chunk(rb, c + 32);
      tmem_ld_wait();
    }
  }
  for (int c = nfull; c < nj; c += 16) {
    uint32_t r[16];
    tmem_ld_32x16(t_lane + c, r);
    tmem_ld_wait();
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), scale_log2, -m_new));
      float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m_new));
      if (c + i >= nvalid) p0 = 0.f;
      if (c + i + 1 >= nvalid) p1 = 0.f;
      ps0 += p0;
      ps1 += p1;
      pk[i >> 1] = pack_bf16x2(p0, p1);
    }
    tmem_st_32x8(t_lane + (c >> 1), pk);
  }
  if (turn_pass) mbar_arrive(turn_pass);
  tmem_st_wait();
  if (tp) tp[1] += clock64() - tp1;
  return (ps0 + ps1) + (ps2 + ps3);
}

// Register-lean variant (one 32-column chunk live at a time) for the multi-block path, where the
// fp32 output accumulator already occupies 64 registers.
__device__ __forceinline__ float softmax_block_lean(uint32_t t_lane, int nvalid, int nj, float scale_log2,
                                                    float m_run, float& m_new_out) {
  const int nfull = nvalid & ~31;
  float mx0 = -INFINITY, mx1 = -INFINITY;
  for (int c = 0; c < nfull; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(t_lane + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      mx0 = fmax3(mx0, __uint_as_float(r[i + 0]), __uint_as_float(r[i + 1]));
      mx1 = fmax3(mx1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    }
  }
  for (int c = nfull; c < nj; c += 16) {
    uint32_t r[16];
    tmem_ld_32x16(t_lane + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c + i < nvalid) mx0 = fmaxf(mx0, __uint_as_float(r[i]));
  }
  const float m_new = fmaxf(m_run, fmaxf(mx0, mx1) * scale_log2);
  m_new_out = m_new;
  float ps0 = 0.f, ps1 = 0.f;
  for (int c = 0; c < nfull; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(t_lane + c, r);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), scale_log2, -m_new));
      const float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m_new));
      ps0 += p0;
      ps1 += p1;
      pk[i >> 1] = pack_bf16x2(p0, p1);
    }
    tmem_st_32x16(t_lane + (c >> 1), pk);
  }
  for (int c = nfull; c < nj; c += 16) {
    uint32_t r[16];
    tmem_ld_32x16(t_lane + c, r);
    tmem_ld_wait();
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), scale_log2, -m_new));
      float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m_new));
      if (c + i >= nvalid) p0 = 0.f;
      if (c + i + 1 >= nvalid) p1 = 0.f;
      ps0 += p0;
      ps1 += p1;
      pk[i >> 1] = pack_bf16x2(p0, p1);
    }
    tmem_st_32x8(t_lane + (c >> 1), pk);
  }
  tmem_st_wait();
  return ps0 + ps1;
}

__global__ void __launch_bounds__(kThreads2, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o,
                 const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int kv_bytes = p.bkv * kDH * 2;                  // one K or V block
  const int slot_bytes = kQBytes + kOutBytes + 2 * kv_bytes;
  // slot layout: [Q 16 KB][O staging 16 KB][K][V]
  const uint32_t bar_base = smem_base + 2 * slot_bytes;
  const uint32_t tmem_slot = bar_base + 8u * A_NBARS;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + 2 * slot_bytes + 8 * A_NBARS);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // role and slot of this warp
  const int g = (warp_idx < 8) ? (warp_idx >> 2) : (warp_idx & 1);
  const bool is_softmax = warp_idx < 8;
  const bool is_producer = warp_idx == 8 || warp_idx == 9;
  const bool is_mma = warp_idx == 10 || warp_idx == 11;

  const uint32_t slot_smem = smem_base + g * slot_bytes;
  const uint32_t q_smem = slot_smem;
  const uint32_t o_smem = slot_smem + kQBytes;
  const uint32_t k_smem = o_smem + kOutBytes;
  const uint32_t v_smem = k_smem + kv_bytes;
  auto bar = [&](int i) { return bar_base + 8u * (g * A_PER_SLOT + i); };

  if (warp_idx == 11 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      const uint32_t b0 = bar_base + 8u * (s * A_PER_SLOT);
      for (int i = 0; i < A_PER_SLOT; ++i)
        mbar_init(b0 + 8u * i, (i == A_PFULL || i == A_OREAD || i == A_TURN) ? 128 : 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 9 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_o);
  }
  if (warp_idx == 8) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint32_t t_slot = tmem_base + g * kSlotCols;

  // Work list of this slot: items w, w + W, w + 2W, ... with w = 2*blockIdx.x + g, W = 2*gridDim.x
  const long long first_item = 2LL * blockIdx.x + g;
  const long long item_step = 2LL * gridDim.x;
  const int n_local = (p.total_items > first_item)
                          ? static_cast<int>((p.total_items - first_item + item_step - 1) / item_step)
                          : 0;
  const int nblk = p.nblk;
  const int bkv = p.bkv;

  auto decode = [&](int local, int& img, int& head, int& qt) {
    long long item = first_item + static_cast<long long>(local) * item_step;
    if (p.reverse) item = p.total_items - 1 - item;
    qt = static_cast<int>(item % p.nqt);
    const long long bh = item / p.nqt;
    head = static_cast<int>(bh % p.H);
    img = static_cast<int>(bh / p.H);
  };

  if (is_producer) {
    // ------------------------------------------------------------------ TMA producer of slot g
    uint32_t step = 0;
    for (int it = 0; it < n_local; ++it) {
      int img, head, qt;
      decode(it, img, head, qt);
      mbar_wait(bar(A_QEMPTY), (static_cast<uint32_t>(it) & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar(A_QFULL), kQBytes);
        tma_load_3d(&tma_q, bar(A_QFULL), q_smem, head * kDH, qt * kQTile, img, kEvictFirst);
      }
      __syncwarp();
      for (int j = 0; j < nblk; ++j, ++step) {
        const uint32_t ph = step & 1u;
        mbar_wait(bar(A_KEMPTY), ph ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar(A_KFULL), kv_bytes);
          tma_load_3d(&tma_k, bar(A_KFULL), k_smem, head * kDH, j * bkv, img, kEvictNormal);
        }
        __syncwarp();
        mbar_wait(bar(A_VEMPTY), ph ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar(A_VFULL), kv_bytes);
          tma_load_3d(&tma_v, bar(A_VFULL), v_smem, head * kDH, j * bkv, img, kEvictNormal);
        }
        __syncwarp();
      }
    }
  } else if (is_mma) {
    // ------------------------------------------------------------------ MMA issuer of slot g
    uint32_t step = 0;
#ifdef VT_ATTN_STAGGER
    if (g == 1) {   // start slot 1 half a period late so the two slots' softmax phases interleave
      const long long t0 = clock64();
      while (clock64() - t0 < VT_ATTN_STAGGER) { }
    }
#endif
    for (int it = 0; it < n_local; ++it) {
      for (int j = 0; j < nblk; ++j, ++step) {
        const uint32_t ph = step & 1u;
        int nj = p.N - j * bkv;
        if (nj > bkv) nj = bkv;
        nj = (nj + 15) & ~15;
        // ---- S = Q K_j^T
        if (j == 0) mbar_wait(bar(A_QFULL), static_cast<uint32_t>(it) & 1u);
        mbar_wait(bar(A_KFULL), ph);
        mbar_wait(bar(A_OREAD), ph ^ 1u);   // TMEM region free: previous step's O has been consumed
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t idesc = make_idesc_bf16(kQTile, nj, 0, 0);
          const uint64_t qd = make_desc_kmajor_sw128(q_smem);
          const uint64_t kd = make_desc_kmajor_sw128(k_smem);
#pragma unroll
          for (int k = 0; k < kDH / 16; ++k) umma_ss(t_slot, qd + 2 * k, kd + 2 * k, idesc, k != 0 ? 1u : 0u);
          umma_commit(bar(A_SFULL));
          umma_commit(bar(A_KEMPTY));
          if (j == nblk - 1) umma_commit(bar(A_QEMPTY));
        }
        __syncwarp();
        // ---- O_j = P V_j
        mbar_wait(bar(A_PFULL), ph);
        mbar_wait(bar(A_VFULL), ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t idesc = make_idesc_bf16(kQTile, kDH, 0, 1);
          const uint64_t vd = make_desc_mnmajor_sw128(v_smem, 1024);
          const int ksteps = nj >> 4;
          for (int k = 0; k < ksteps; ++k)   // 16 kv rows = 2048 bytes = 128 units of the address field
            umma_ts(t_slot + kOCol, t_slot + 8 * k, vd + 128 * k, idesc, k != 0 ? 1u : 0u);
          umma_commit(bar(A_OFULL));
          umma_commit(bar(A_VEMPTY));
        }
        __syncwarp();
      }
    }
  } else if (is_softmax) {
    // ------------------------------------------------------------------ softmax warpgroup of slot g
    const int quarter = warp_idx & 3;
    const uint32_t t_lane = t_slot + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t stage_addr = o_smem + quarter * 4096;
    uint8_t* stage_row = smem_gen + g * slot_bytes + kQBytes + quarter * 4096 + lane * 128;
    const int sw = lane & 7;
    const uint32_t other_turn = bar_base + 8u * ((1 - g) * A_PER_SLOT + A_TURN);
    uint32_t step = 0;
    long long dacc[4] = {0, 0, 0, 0};
    long long pass1[2] = {0, 0};
    const long long dt0 = p.dbg ? clock64() : 0;

    for (int it = 0; it < n_local; ++it) {
      int img, head, qt;
      decode(it, img, head, qt);
      uint32_t packed[kDH / 2];   // normalised bf16x2 output row
      long long e0 = 0;

      if (nblk == 1) {
        // single KV block (N <= 256): no running rescale, O is read once
        const uint32_t ph = step & 1u;
        ++step;
        const int nvalid = p.N;
        const int nj = (nvalid + 15) & ~15;
        long long c0 = 0, c1 = 0;
        if (p.dbg) c0 = clock64();
        mbar_wait(bar(A_SFULL), ph);
        if (p.dbg) { c1 = clock64(); dacc[0] += c1 - c0; c0 = c1; }
        tc_fence_after();
        float m_new;
        const bool wait_turn = !(g == 0 && it == 0);
        const uint32_t turn_par = static_cast<uint32_t>(g == 0 ? it - 1 : it) & 1u;
        const float l = softmax_block_pipelined(t_lane, nvalid, nj, p.scale_log2, -INFINITY, m_new,
                                                wait_turn ? bar(A_TURN) : 0u, turn_par, other_turn,
                                                p.dbg ? pass1 : nullptr);
        tc_fence_before();
        mbar_arrive(bar(A_PFULL));
        if (p.dbg) { c1 = clock64(); dacc[1] += c1 - c0; c0 = c1; }
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(l));
        mbar_wait(bar(A_OFULL), ph);
        if (p.dbg) { c1 = clock64(); dacc[2] += c1 - c0; c0 = c1; }
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kDH; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(t_lane + kOCol + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            packed[(c + i) >> 1] = pack_bf16x2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
        }
        tc_fence_before();
        mbar_arrive(bar(A_OREAD));
      } else {
        float o_acc[kDH];
#pragma unroll
        for (int i = 0; i < kDH; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY;
        float l_run = 0.f;
        for (int j = 0; j < nblk; ++j, ++step) {
          const uint32_t ph = step & 1u;
          int nvalid = p.N - j * bkv;
          if (nvalid > bkv) nvalid = bkv;
          const int nj = (nvalid + 15) & ~15;
          mbar_wait(bar(A_SFULL), ph);
          tc_fence_after();
          float m_new;
          const float psum = softmax_block_lean(t_lane, nvalid, nj, p.scale_log2, m_run, m_new);
          tc_fence_before();
          mbar_arrive(bar(A_PFULL));
          const float alpha = ex2_approx(m_run - m_new);   // first block: exp2(-inf) = 0
          m_run = m_new;
          l_run = l_run * alpha + psum;
          mbar_wait(bar(A_OFULL), ph);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kDH; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(t_lane + kOCol + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[c + i] = fmaf(o_acc[c + i], alpha, __uint_as_float(r[i]));
          }
          tc_fence_before();
          mbar_arrive(bar(A_OREAD));
        }
        const float inv = 1.0f / l_run;
#pragma unroll
        for (int i = 0; i < kDH; i += 2) packed[i >> 1] = pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv);
      }

      if (p.dbg) e0 = clock64();
      // stage as a swizzled [32 rows x 128 B] tile per warp, TMA-store (clipped at N by the 3-D map)
      if (lane == 0) tma_store_wait_read<0>();   // previous store out of this staging tile is done
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const uint4 o4 = make_uint4(packed[4 * jj], packed[4 * jj + 1], packed[4 * jj + 2], packed[4 * jj + 3]);
        *reinterpret_cast<uint4*>(stage_row + ((jj ^ sw) << 4)) = o4;
      }
      fence_proxy_async_smem();
      __syncwarp();
      const int row0 = qt * kQTile + quarter * 32;
      if (lane == 0 && row0 < p.N) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     :
                     : "l"(reinterpret_cast<uint64_t>(&tma_o)), "r"(stage_addr), "r"(head * kDH),
                       "r"(row0), "r"(img)
                     : "memory");
        tma_store_commit();
      }
      if (p.dbg) dacc[3] += clock64() - e0;
    }
    if (lane == 0) tma_store_wait<0>();
    if (p.dbg && (warp_idx & 3) == 0 && lane == 0) {
      long long* d = p.dbg + (2LL * blockIdx.x + g) * 8;
      d[0] = dacc[0]; d[1] = dacc[1]; d[2] = dacc[2]; d[3] = dacc[3];
      d[4] = n_local; d[5] = clock64() - dt0; d[6] = pass1[0]; d[7] = pass1[1];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 8) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

long long* g_attn_dbg = nullptr;

}  // namespace

void attn2_set_debug_buffer(void* ptr) { g_attn_dbg = static_cast<long long*>(ptr); }

int attn2_fwd_tcgen05(const void* q, const void* k, const void* v, void* out, int B, int H, int N,
                      int dh, long long qkv_row_stride, long long qkv_batch_stride,
                      long long out_row_stride, long long out_batch_stride, float scale, int reverse,
                      cudaStream_t stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || N <= 0) return VT_ERR_ARG;
  if (dh != kDH) return VT_ERR_UNSUPPORTED;
  if ((qkv_row_stride % 8) || (qkv_batch_stride % 8) || (out_row_stride % 8) || (out_batch_stride % 8))
    return VT_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return VT_ERR_ALIGN;

  Attn2Params p;
  p.N = N;
  p.H = H;
  p.nqt = (N + kQTile - 1) / kQTile;
  p.nblk = (N + 255) / 256;
  int bkv = (N + p.nblk - 1) / p.nblk;
  p.bkv = (bkv + 15) & ~15;
  p.total_items = static_cast<long long>(B) * H * p.nqt;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.reverse = reverse;
  p.dbg = g_attn_dbg;
  const int kv_bytes = p.bkv * kDH * 2;
  const int smem = 1024 + 2 * (kQBytes + kOutBytes + 2 * kv_bytes) + 8 * A_NBARS + 16;
  if (smem > kSmemLimit) return VT_ERR_UNSUPPORTED;

  CUtensorMap tq, tk, tv, to;
  int rc = make_tmap_bf16_3d(&tq, q, static_cast<uint64_t>(H) * dh, N, B, qkv_row_stride, qkv_batch_stride,
                             dh, kQTile, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tk, k, static_cast<uint64_t>(H) * dh, N, B, qkv_row_stride, qkv_batch_stride, dh,
                         p.bkv, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tv, v, static_cast<uint64_t>(H) * dh, N, B, qkv_row_stride, qkv_batch_stride, dh,
                         p.bkv, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&to, out, static_cast<uint64_t>(H) * dh, N, B, out_row_stride, out_batch_stride, dh,
                         32, TMAP_SW_128);
  if (rc) return rc;

  static int granted[kMaxDevices] = {0};
  if (const int rc_attr = ensure_dynamic_smem(attn2_fwd_kernel, smem, granted)) return rc_attr;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (p.total_items + 1) / 2 < sms ? (p.total_items + 1) / 2 : sms;
  attn2_fwd_kernel<<<static_cast<unsigned>(grid), kThreads2, smem, stream>>>(tq, tk, tv, to, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace vt
