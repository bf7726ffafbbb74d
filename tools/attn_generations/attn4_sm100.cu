// K3 (main variant, head dim 64): persistent fused attention forward on tcgen05 with the score
// tile double-buffered in TMEM and FOUR softmax threads per query row.
//
//   ctx[b, i, h*64:(h+1)*64] = softmax_j( scale * q[b,i,h] . k[b,j,h] ) @ v[b,j,h]      (bf16)
//
// Replaces the reference's per-head matmul3 -> softmax -> matmul3 -> slice-assign chain
// (vit/vit.py:60-72,101-108) for all heads at once; the score matrix never leaves the SM.
//
// Why this shape.  In-kernel cycle counters on attn3 (two independent slots that alias S, P and O in
// 256 TMEM columns each) showed the exponentials are NOT the limit there: with the MUFU removed the
// kernel was 6 % faster.  The limit was the serial chain P -> PV MMA -> O read -> next S MMA -> row
// max inside a slot (~5000 cycles of barrier hand-offs and MMA latency per item against ~1700 cycles
// of MUFU work), forced by S, P and O sharing columns.  Here one CTA works on ONE item at a time
// with all 16 softmax warps, and the MMA issuer runs two score tiles ahead:
//
//   TMEM (512 columns)   S0 [0, 208)   S1 [208, 416)   O [416, 480)   row sums [480, 496)
//   unit u (one KV block of one item) uses S[u & 1]; its probabilities P (bf16x2) overwrite the
//   thread's own already-consumed scores; O_u = P_u V_u goes to the O columns, and a second N = 16
//   MMA per K step against a tile of ones leaves the row sums of the bf16 probabilities next to it
//   (no FADD per element, no sum exchange between the threads of a row).
//
//   softmax warps, per unit u:  wait S_u | row max of my columns | exchange (4 threads per row) |
//                               exp2 -> P_u in place | arrive P_u | read + fold O_{u-1}
//   MMA issuer, per unit u:     wait P_u, V_u, O_{u-1} read | PV_u | then S_{u+2} = Q K^T into S[u & 1]
//
// so every MMA (and its barrier round trip) of unit u+1 and u+2 overlaps the exponentials of unit u:
// the softmax warps never wait for the tensor core in steady state.  The first version of this
// kernel was issue-bound on bookkeeping (13.8 k warp instructions per item, 16 % of them the
// exponentials' FFMA/MUFU/F2FP: profiles/README.md), hence: item coordinates are decoded once by the
// producer and passed through a shared-memory ring, column ranges are computed once per kernel,
// scores are walked in 16-column groups with a warp-uniform "group is partly masked" branch instead
// of per-element predicates, and barrier waits keep their slow path out of line.
//
//   warps 0-15  softmax: warp = 4 * column_quarter + row_quarter (TMEM lanes 32*row_quarter ..)
//   warp 16     TMA producer (also allocates TMEM)      warp 17   MMA issuer
//
// Sequences longer than 208 keys run as several KV blocks per item with the online-softmax
// recurrence (output accumulator, 16 columns per thread, in registers); that path is correct and
// tested but measured slower than attn3 (418 vs 333 us at 577 tokens), so api.cu routes only
// single-block sequences (N <= 208) here.  Head dim 80 stays on attn3_sm100.cu.
// Measured at C2 (256 x 12 heads x 197 tokens): 92.9 us against 99.7 us for attn3.  Tried and
// rejected: keeping the scores of the max pass in registers (64 live registers spill: 101 us) and a
// direct st.global epilogue instead of staging + TMA store (99 us).
#include "attn_softmax.cuh"
#include "common.cuh"
#include "tensormap.h"

namespace vt {

namespace {

constexpr int kDH = 64;
constexpr int kQTile = 128;
constexpr int kSoftmaxWarps4 = 16;
constexpr int kThreads4 = (kSoftmaxWarps4 + 2) * 32;   // 576
constexpr int kQBytes = kQTile * kDH * 2;              // 16 KB
constexpr int kMaxBkv4 = 208;
constexpr int kSCols = 208;                            // columns per score buffer
constexpr int kOCol = 2 * kSCols;                      // 416: O (64 columns) then the row sums (16 columns)
constexpr int kLCol = kOCol + kDH;                     // 480
constexpr int kStageBytes4 = 32 * 16 * 2;              // per warp: 32 rows x 16 bf16
constexpr int kOnesBytes = 16 * 16 * 2;                // [16 keys x 16 columns] of bf16 ones
constexpr int kRing = 8;                               // item descriptors in flight
constexpr int kSmemLimit4 = 232448;

struct Attn4Params {
  int N, H, B;
  int nqt;            // query tiles per (image, head)
  int bkv, nblk;      // rows per KV block (multiple of 16, <= 208), KV blocks per item
  long long total_items;
  int reverse;        // walk the images from the last to the first (L2 reuse, see api.cu)
  float scale_log2;
  long long* dbg;     // optional cycle counters (developer tool tools/attn_dbg.py)
};

enum { B_QFULL = 0, B_QEMPTY = 2, B_KFULL = 4, B_KEMPTY = 6, B_VFULL = 8, B_VEMPTY = 10, B_SFULL = 12,
       B_PFULL = 14, B_OFULL = 16, B_OREAD = 17, B_NBARS = 18 };

// Columns [c0, c1) of a block of nj (multiple of 16) score columns owned by column quarter cq:
// 16-column groups dealt out as evenly as possible, the first quarters take the remainder.
__device__ __forceinline__ void quarter_cols(int nj, int cq, int& c0, int& c1) {
  const int n16 = nj >> 4;
  const int base = n16 >> 2, rem = n16 & 3;
  const int lo = cq * base + (cq < rem ? cq : rem);
  c0 = lo << 4;
  c1 = (lo + base + (cq < rem ? 1 : 0)) << 4;
}

// mbarrier wait whose slow path (sleeping poll loop, watchdog, printf) is NOT inlined: the hot loops
// keep three instructions per wait and no registers for the diagnostics.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

// Row max over NG groups of 16 score columns starting at TMEM address a.  Only the first nv of the
// 16 * NG columns are real keys; nj - nvalid < 16, so at most the LAST group is partly masked.
template <int NG>
__device__ __forceinline__ float max_groups_t(uint32_t a, int nv) {
  uint32_t r[NG][16];
#pragma unroll
  for (int g = 0; g < NG; ++g) tmem_ld_32x16(a + 16 * g, r[g]);
  tmem_ld_wait();
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int g = 0; g < NG - 1; ++g) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      m0 = fmax3(m0, __uint_as_float(r[g][i]), __uint_as_float(r[g][i + 1]));
      m1 = fmax3(m1, __uint_as_float(r[g][i + 2]), __uint_as_float(r[g][i + 3]));
    }
  }
  if (nv >= 16 * NG) {   // warp-uniform
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      m0 = fmax3(m0, __uint_as_float(r[NG - 1][i]), __uint_as_float(r[NG - 1][i + 1]));
      m1 = fmax3(m1, __uint_as_float(r[NG - 1][i + 2]), __uint_as_float(r[NG - 1][i + 3]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (16 * (NG - 1) + i < nv) m0 = fmaxf(m0, __uint_as_float(r[NG - 1][i]));
  }
  return fmaxf(m0, m1);
}

__device__ __forceinline__ float max_groups(uint32_t a, int ng, int nv) {
  switch (ng) {   // warp-uniform
    case 4: return max_groups_t<4>(a, nv);
    case 3: return max_groups_t<3>(a, nv);
    case 2: return max_groups_t<2>(a, nv);
    case 1: return max_groups_t<1>(a, nv);
    default: return -INFINITY;
  }
}

// one group: p = exp2(s * scale - m) -> bf16x2 -> TMEM columns dst .. dst + 7
template <bool MASKED>
__device__ __forceinline__ void exp_group(const uint32_t (&r)[16], uint32_t dst, int nv_in_group, float scale_log2,
                                          float m) {
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), scale_log2, -m));
    float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m));
    if (MASKED) {
      if (i >= nv_in_group) p0 = 0.f;
      if (i + 1 >= nv_in_group) p1 = 0.f;
    }
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
  tmem_st_32x8(dst, pk);
}

// p = exp2(s * scale - m) over the same groups; P (bf16x2) of group g overwrites TMEM columns
// a + 8 * g .. a + 8 * g + 7 (scores this thread has already consumed).  Groups are loaded two at a
// time, the next pair is in flight during the math of the current one.  (Keeping the scores of the
// max pass in registers instead of re-reading them was measured slower: 64 live registers spill.)
template <int NG>
__device__ __forceinline__ void exp_groups_t(uint32_t a, int nv, float scale_log2, float m) {
  uint32_t r[NG][16];
#pragma unroll
  for (int g = 0; g < NG && g < 2; ++g) tmem_ld_32x16(a + 16 * g, r[g]);
  tmem_ld_wait();
#pragma unroll
  for (int g = 2; g < NG; ++g) tmem_ld_32x16(a + 16 * g, r[g]);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g == 2) tmem_ld_wait();
    if (g < NG - 1 || nv >= 16 * NG)
      exp_group<false>(r[g], a + 8 * g, 16, scale_log2, m);
    else
      exp_group<true>(r[g], a + 8 * g, nv - 16 * (NG - 1), scale_log2, m);
  }
  tmem_st_wait();
}

__device__ __forceinline__ void exp_groups(uint32_t a, int ng, int nv, float scale_log2, float m) {
  switch (ng) {   // warp-uniform
    case 4: exp_groups_t<4>(a, nv, scale_log2, m); break;
    case 3: exp_groups_t<3>(a, nv, scale_log2, m); break;
    case 2: exp_groups_t<2>(a, nv, scale_log2, m); break;
    case 1: exp_groups_t<1>(a, nv, scale_log2, m); break;
    default: break;
  }
}

// kSingle: every item is ONE KV block (N <= 208, the ViT-B/L 224-pixel case): no online-softmax
// state (running max, rescale factor, output accumulator) is carried between units.
template <bool kSingle>
__global__ void __launch_bounds__(kThreads4, 1)
attn4_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o,
                 const Attn4Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int kv_bytes = p.bkv * kDH * 2;
  // [Q0][Q1][K0][K1][V0][V1][staging 16 x 1 KB][ones][barriers][tmem slot][item ring][row max exchange]
  const uint32_t q_smem = smem_base;
  const uint32_t k_smem = q_smem + 2 * kQBytes;
  const uint32_t v_smem = k_smem + 2 * kv_bytes;
  const int stage_off = 2 * kQBytes + 4 * kv_bytes;
  const uint32_t stage_smem = smem_base + stage_off;
  const int ones_off = stage_off + kSoftmaxWarps4 * kStageBytes4;
  const uint32_t ones_smem = smem_base + ones_off;
  const int bar_off = ones_off + kOnesBytes;
  const uint32_t bar_base = smem_base + bar_off;
  const uint32_t tmem_slot = bar_base + 8u * B_NBARS;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * B_NBARS);
  int4* ring = reinterpret_cast<int4*>(smem_gen + bar_off + 8 * B_NBARS + 16);              // [kRing]
  float* xm = reinterpret_cast<float*>(smem_gen + bar_off + 8 * B_NBARS + 16 + 16 * kRing);   // [2][4][128]
  auto bar = [&](int i) { return bar_base + 8u * i; };

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 17 && lane == 0) {
    for (int i = 0; i < B_NBARS; ++i)
      mbar_init(bar(i), (i == B_PFULL || i == B_PFULL + 1 || i == B_OREAD) ? kSoftmaxWarps4 : 1);
    fence_barrier_init();
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_o);
  }
  if (warp_idx == 0) {   // the tile of ones behind the row-sum MMAs (read through the async proxy)
    reinterpret_cast<uint4*>(smem_gen + ones_off)[lane] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  if (warp_idx == 16) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // Work list of this CTA: items blockIdx.x, blockIdx.x + grid, ...; a unit = one KV block of an item
  const long long first_item = blockIdx.x;
  const long long item_step = gridDim.x;
  const int n_local = (p.total_items > first_item)
                          ? static_cast<int>((p.total_items - first_item + item_step - 1) / item_step)
                          : 0;
  const int nblk = kSingle ? 1 : p.nblk;
  const int bkv = p.bkv;
  const int n_units = n_local * nblk;
  // the two block shapes: blocks 0 .. nblk-2 have bkv keys, the last one the rest
  const int nv_full = (p.N < bkv) ? p.N : bkv;
  const int nv_last = p.N - (nblk - 1) * bkv;
  const int nj_full = (nv_full + 15) & ~15;
  const int nj_last = (nv_last + 15) & ~15;

  if (warp_idx == 16) {
    // ------------------------------------------------------------------ TMA producer
    int u = 0;
    for (int it = 0; it < n_local; ++it) {
      const unsigned item = static_cast<unsigned>(first_item + static_cast<long long>(it) * item_step);
      const int qt = static_cast<int>(item % static_cast<unsigned>(p.nqt));   // total_items < 2^31 (host)
      const unsigned bh = item / static_cast<unsigned>(p.nqt);
      const int head = static_cast<int>(bh % static_cast<unsigned>(p.H));
      int img = static_cast<int>(bh / static_cast<unsigned>(p.H));
      if (p.reverse) img = p.B - 1 - img;
      const int qb = it & 1;
      mbar_wait(bar(B_QEMPTY + qb), ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u);
      if (elect_one_sync()) {
        // item coordinates for the softmax warps: visible to them through the barrier chain
        // B_QFULL -> (MMA issuer) -> B_SFULL; the ring is deeper than the producer can run ahead
        ring[it & (kRing - 1)] = make_int4(img, head, qt, 0);
        mbar_arrive_expect_tx(bar(B_QFULL + qb), kQBytes);
        tma_load_3d(&tma_q, bar(B_QFULL + qb), q_smem + qb * kQBytes, head * kDH, qt * kQTile, img, kEvictFirst);
      }
      __syncwarp();
      for (int j = 0; j < nblk; ++j, ++u) {
        const int b = u & 1;
        const uint32_t ph = (static_cast<uint32_t>(u) >> 1) & 1u;
        mbar_wait(bar(B_KEMPTY + b), ph ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar(B_KFULL + b), kv_bytes);
          tma_load_3d(&tma_k, bar(B_KFULL + b), k_smem + b * kv_bytes, head * kDH, j * bkv, img, kEvictNormal);
        }
        __syncwarp();
        mbar_wait(bar(B_VEMPTY + b), ph ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar(B_VFULL + b), kv_bytes);
          tma_load_3d(&tma_v, bar(B_VFULL + b), v_smem + b * kv_bytes, head * kDH, j * bkv, img, kEvictNormal);
        }
        __syncwarp();
      }
    }
  } else if (warp_idx == 17) {
    // ------------------------------------------------------------------ MMA issuer
    // S_u = Q K_u^T into score buffer u & 1 (free: PV_{u-2}, issued earlier by this thread, is the
    // last reader of that buffer and tcgen05.mma executes in issue order).
    auto issue_scores = [&](int u, int it, int j) {
      const int b = u & 1;
      const int qb = it & 1;
      const int nj = (j == nblk - 1) ? nj_last : nj_full;
      if (j == 0) mbar_wait(bar(B_QFULL + qb), (static_cast<uint32_t>(it) >> 1) & 1u);
      mbar_wait(bar(B_KFULL + b), (static_cast<uint32_t>(u) >> 1) & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t idesc = make_idesc_bf16(kQTile, nj, 0, 0);
        const uint64_t qd = make_desc_kmajor_sw128(q_smem + qb * kQBytes);
        const uint64_t kd = make_desc_kmajor_sw128(k_smem + b * kv_bytes);
#pragma unroll
        for (int k = 0; k < kDH / 16; ++k)
          umma_ss(tmem_base + b * kSCols, qd + 2 * k, kd + 2 * k, idesc, k != 0 ? 1u : 0u);
        umma_commit(bar(B_SFULL + b));
        umma_commit(bar(B_KEMPTY + b));
        if (j == nblk - 1) umma_commit(bar(B_QEMPTY + qb));
      }
      __syncwarp();
    };
    // (item, block) of unit u + 2, advanced incrementally
    int it2 = 0, j2 = 0;
    auto advance2 = [&]() { if (++j2 == nblk) { j2 = 0; ++it2; } };
    if (n_units > 0) { issue_scores(0, it2, j2); advance2(); }
    if (n_units > 1) { issue_scores(1, it2, j2); advance2(); }
    int j = 0;
    for (int u = 0; u < n_units; ++u) {
      const int b = u & 1;
      const uint32_t ph = (static_cast<uint32_t>(u) >> 1) & 1u;
      const int nj = (j == nblk - 1) ? nj_last : nj_full;
      // ---- O_u = P_u V_u and the row sums P_u 1
      mbar_wait(bar(B_VFULL + b), ph);
      if (u > 0) mbar_wait(bar(B_OREAD), static_cast<uint32_t>(u - 1) & 1u);   // O columns free
      mbar_wait(bar(B_PFULL + b), ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t idesc = make_idesc_bf16(kQTile, kDH, 0, 1);
        const uint32_t idesc_l = make_idesc_bf16(kQTile, 16, 0, 1);
        const uint64_t vd = make_desc_mnmajor_sw128(v_smem + b * kv_bytes, 1024);
        const uint64_t od = make_smem_desc(ones_smem, 256, 256, 6);   // every element is 1: layout is moot
        const uint32_t s_tmem = tmem_base + b * kSCols;
        int k = 0;
#pragma unroll 1
        for (int cq = 0; cq < 4; ++cq) {
          int c0, c1;
          quarter_cols(nj, cq, c0, c1);
          for (int c = c0; c < c1; c += 16, ++k) {   // 16 keys: 8 packed P columns, 2048 B of V
            const uint32_t a_tmem = s_tmem + c0 + ((c - c0) >> 1);
            umma_ts(tmem_base + kOCol, a_tmem, vd + 128 * k, idesc, k != 0 ? 1u : 0u);
            umma_ts(tmem_base + kLCol, a_tmem, od, idesc_l, k != 0 ? 1u : 0u);
          }
        }
        umma_commit(bar(B_OFULL));
        umma_commit(bar(B_VEMPTY + b));
      }
      __syncwarp();
      if (u + 2 < n_units) { issue_scores(u + 2, it2, j2); advance2(); }
      if (++j == nblk) j = 0;
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int cq = warp_idx >> 2;      // column quarter
    const int rq = warp_idx & 3;       // row quarter = TMEM lane group
    const int row_in_tile = rq * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(rq * 32) << 16);
    const uint32_t stage_addr = stage_smem + warp_idx * kStageBytes4;
    uint8_t* stage_row = smem_gen + stage_off + warp_idx * kStageBytes4 + lane * 32;
    const int bar_id = 1 + rq;
    // my columns in the two block shapes: first column, 16-column groups, valid columns from my first
    int c0_full, c1_full, c0_last, c1_last;
    quarter_cols(nj_full, cq, c0_full, c1_full);
    quarter_cols(nj_last, cq, c0_last, c1_last);
    const int ng_full = (c1_full - c0_full) >> 4, ng_last = (c1_last - c0_last) >> 4;
    const int nvr_full = nv_full - c0_full, nvr_last = nv_last - c0_last;

    // Cycle counters exist only in the developer build (make EXTRA=-DVT_ATTN4_DBG, tools/attn_dbg.py):
    // they cost registers the production kernel does not have to spare.
#ifdef VT_ATTN4_DBG
    const bool dbg_on = p.dbg != nullptr;
    unsigned dacc[7] = {0, 0, 0, 0, 0, 0, 0};
    const unsigned dt0 = dbg_on ? static_cast<unsigned>(clock()) : 0u;
    unsigned tc = 0;
#define VT_TICK4(i) if (dbg_on) { const unsigned t_ = static_cast<unsigned>(clock()); dacc[i] += t_ - tc; tc = t_; }
#define VT_TICK4_START() if (dbg_on) tc = static_cast<unsigned>(clock());
#else
#define VT_TICK4(i)
#define VT_TICK4_START()
#endif

    // state of the item whose exponentials are being computed
    int4 desc = make_int4(0, 0, 0, 0);   // (image, head, query tile)
    bool live = false;
    float m_run = -INFINITY;
    // state of the unit whose output block is still to be folded (one unit behind)
    float o_acc[16];
    float l_acc = 0.f;
    float pend_alpha = 0.f;
    int4 pend_desc = make_int4(0, 0, 0, 0);
    bool pend_first = false, pend_last = false, pend_live = false;

    // fold O_v (v = u - 1) into the accumulator; finish the item if v was its last block
    auto fold_output = [&](int v) {
      VT_TICK4_START()
      mbar_wait_lean(bar(B_OFULL), static_cast<uint32_t>(v) & 1u);
      VT_TICK4(4)
      tc_fence_after();
      uint32_t r[16];
      uint32_t rl[8];
      if (pend_live) {
        tmem_ld_32x16(t_lane + kOCol + cq * 16, r);
        tmem_ld_32x8(t_lane + kLCol, rl);   // every one of the 16 sum columns holds the row sum
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_OREAD));
      if (!pend_live) return;
      if (kSingle || pend_first) {
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[i] = __uint_as_float(r[i]);
        l_acc = __uint_as_float(rl[0]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[i] = fmaf(o_acc[i], pend_alpha, __uint_as_float(r[i]));
        l_acc = fmaf(l_acc, pend_alpha, __uint_as_float(rl[0]));
      }
      VT_TICK4(5)
      if (!kSingle && !pend_last) return;
      float inv;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(l_acc));
      // stage this warp's [32 rows x 16 columns] and TMA-store it
      if (lane == 0) tma_store_wait_read<0>();   // previous store out of this staging tile is done
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        uint4 o4;
        o4.x = pack_bf16x2(o_acc[8 * jj + 0] * inv, o_acc[8 * jj + 1] * inv);
        o4.y = pack_bf16x2(o_acc[8 * jj + 2] * inv, o_acc[8 * jj + 3] * inv);
        o4.z = pack_bf16x2(o_acc[8 * jj + 4] * inv, o_acc[8 * jj + 5] * inv);
        o4.w = pack_bf16x2(o_acc[8 * jj + 6] * inv, o_acc[8 * jj + 7] * inv);
        *reinterpret_cast<uint4*>(stage_row + 16 * jj) = o4;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     :
                     : "l"(reinterpret_cast<uint64_t>(&tma_o)), "r"(stage_addr),
                       "r"(pend_desc.y * kDH + cq * 16), "r"(pend_desc.z * kQTile + rq * 32), "r"(pend_desc.x)
                     : "memory");
        tma_store_commit();
      }
      VT_TICK4(6)
    };

    int it = 0, j = 0;
    for (int u = 0; u < n_units; ++u) {
      const int b = u & 1;
      const bool last_blk = (j == nblk - 1);
      const uint32_t t_mine = t_lane + b * kSCols + (last_blk ? c0_last : c0_full);
      const int ng = last_blk ? ng_last : ng_full;
      const int nvr = last_blk ? nvr_last : nvr_full;

      VT_TICK4_START()
      mbar_wait_lean(bar(B_SFULL + b), (static_cast<uint32_t>(u) >> 1) & 1u);
      VT_TICK4(0)
      tc_fence_after();
      if (j == 0) {
        desc = ring[it & (kRing - 1)];
        // warp-uniform: all 32 query rows of this warp lie beyond the sequence (N = 197: the last row
        // quarter of every second tile).  Such a warp keeps the barrier protocol and skips the work;
        // its P rows stay undefined (MMA rows are independent, the rows are never stored).
        live = desc.z * kQTile + rq * 32 < p.N;
        m_run = -INFINITY;
      }
      float alpha = 0.f;
      if (live) {
        // pass 1: row max over my columns, exchanged between the four threads of the row
        float* x = xm + b * 4 * kQTile + row_in_tile;
        x[cq * kQTile] = max_groups(t_mine, ng, nvr);
        VT_TICK4(1)
        named_bar_sync(bar_id, 128);
        const float mx = fmaxf(fmaxf(x[0], x[kQTile]), fmaxf(x[2 * kQTile], x[3 * kQTile]));
        float m_new = mx * p.scale_log2;
        if (!kSingle) {
          m_new = fmaxf(m_run, m_new);
          alpha = ex2_approx(m_run - m_new);   // first block: exp2(-inf) = 0
          m_run = m_new;
        }
        VT_TICK4(2)
        // pass 2: exponentials; P overwrites my own consumed scores
        exp_groups(t_mine, ng, nvr, p.scale_log2, m_new);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_PFULL + b));
      VT_TICK4(3)

      if (u > 0) fold_output(u - 1);
      pend_alpha = alpha;
      pend_first = (j == 0);
      pend_last = last_blk;
      pend_live = live;
      pend_desc = desc;
      if (++j == nblk) { j = 0; ++it; }
    }
    if (n_units > 0) fold_output(n_units - 1);
    if (lane == 0) tma_store_wait<0>();
#ifdef VT_ATTN4_DBG
    if (dbg_on && warp_idx == 0 && lane == 0) {
      long long* d = p.dbg + static_cast<long long>(blockIdx.x) * 8;
      for (int i = 0; i < 7; ++i) d[i] = dacc[i];
      d[7] = static_cast<unsigned>(clock()) - dt0;
    }
#endif
#undef VT_TICK4
#undef VT_TICK4_START
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 16) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

long long* g_attn4_dbg = nullptr;

}  // namespace

void attn4_set_debug_buffer(void* ptr) { g_attn4_dbg = static_cast<long long*>(ptr); }

int attn4_fwd_tcgen05(const void* q, const void* k, const void* v, void* out, int B, int H, int N,
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

  Attn4Params p;
  p.N = N;
  p.H = H;
  p.B = B;
  p.nqt = (N + kQTile - 1) / kQTile;
  p.nblk = (N + kMaxBkv4 - 1) / kMaxBkv4;
  int bkv = (N + p.nblk - 1) / p.nblk;
  p.bkv = (bkv + 15) & ~15;
  p.total_items = static_cast<long long>(B) * H * p.nqt;
  if (p.total_items * p.nblk >= (1LL << 31)) return VT_ERR_UNSUPPORTED;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.reverse = reverse;
  p.dbg = g_attn4_dbg;
  const int smem = 1024 + 2 * kQBytes + 4 * p.bkv * kDH * 2 + kSoftmaxWarps4 * kStageBytes4 + kOnesBytes +
                   8 * B_NBARS + 16 + 16 * kRing + 2 * 4 * kQTile * 4;
  if (smem > kSmemLimit4) return VT_ERR_UNSUPPORTED;

  const uint64_t cols = static_cast<uint64_t>(H) * dh;
  CUtensorMap tq, tk, tv, to;
  int rc = make_tmap_bf16_3d(&tq, q, cols, N, B, qkv_row_stride, qkv_batch_stride, kDH, kQTile, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tk, k, cols, N, B, qkv_row_stride, qkv_batch_stride, kDH, p.bkv, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tv, v, cols, N, B, qkv_row_stride, qkv_batch_stride, kDH, p.bkv, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&to, out, cols, N, B, out_row_stride, out_batch_stride, 16, 32, TMAP_SW_NONE);
  if (rc) return rc;

  static int granted_single[kMaxDevices] = {0}, granted_multi[kMaxDevices] = {0};
  if (const int rc_attr = ensure_dynamic_smem(attn4_fwd_kernel<true>, smem, granted_single)) return rc_attr;
  if (const int rc_attr = ensure_dynamic_smem(attn4_fwd_kernel<false>, smem, granted_multi)) return rc_attr;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long grid = p.total_items < sms ? p.total_items : sms;
  if (p.nblk == 1)
    attn4_fwd_kernel<true><<<static_cast<unsigned>(grid), kThreads4, smem, stream>>>(tq, tk, tv, to, p);
  else
    attn4_fwd_kernel<false><<<static_cast<unsigned>(grid), kThreads4, smem, stream>>>(tq, tk, tv, to, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace vt
