// K3 (main variant): persistent fused attention forward on tcgen05, two slots x two column halves.
//
//   ctx[b, i, h*dh:(h+1)*dh] = softmax_j( scale * q[b,i,h] . k[b,j,h] ) @ v[b,j,h]   (bf16, dh = 64 | 80)
//
// Replaces the reference's per-head matmul3 -> softmax -> matmul3 -> slice-assign chain
// (vit/vit.py:60-72,101-108) for all heads at once; the score matrix never leaves the SM.
//
// Same pipeline as attn2_sm100.cu (one CTA per SM, two independent slots, each with its own TMA
// producer warp, MMA issuer warp, Q/K/V buffers and 256 TMEM columns) but every 128-row score tile
// is handled by TWO softmax warpgroups: thread (row, half) owns the row's columns [0, cs) or
// [cs, nj).  Why: one warp sustains only ~26 B/clk of tcgen05.ld and cannot overlap its own TMEM
// latency with its MUFU work (tools/softmax_bench.cu: 10.8 cycles/element with one warp per
// scheduler, 8.7 = MUFU-bound with two), and the serial S -> P -> PV -> O chain per slot leaves the
// MUFU pipe idle half the time with only one warpgroup per slot (tools/attn_dbg.py).
//
//   warps 0-7 / 8-15   softmax warps of slot 0 / 1: warp = 4*half + quarter
//   warps 16 / 17      TMA producer of slot 0 / 1 (warp 16 also allocates TMEM)
//   warps 18 / 19      MMA issuer of slot 0 / 1
// TMEM per slot (256 columns), cs = column split (<= 96), nj <= 208:
//   S fp32 [0, nj)            scores of one KV block
//   P0 bf16x2 [208, 208+cs/2) probabilities of columns [0, cs)   (free columns: no aliasing hazard
//                             between the two threads of a row)
//   P1 bf16x2 [cs, cs+(nj-cs)/2)  probabilities of columns [cs, nj), aliased over that thread's OWN
//                             already-consumed scores
//   O_j fp32 [0, 64)          written by the PV MMAs after every score has been consumed
//                             ([128, 192) for short blocks, where P1 would overlap [0, 64))
// Row max and row sum are exchanged between the two halves through shared memory (one named
// barrier each); everything else is as in attn2.
//
// Head dim 80 (ViT-H): 160-byte rows do not fit the 128-byte swizzle, so every Q/K/V tile is loaded
// as a 64-column SWIZZLE_128B part plus a 16-column SWIZZLE_32B part; QK^T gets a fifth K step on
// the 32-byte tiles and P.V a second MMA (N = 16) per K step into output columns [64, 80).
#include "attn_softmax.cuh"
#include "common.cuh"
#include "tensormap.h"

namespace vt {

namespace {

constexpr int kDM = 64;                               // columns of the SWIZZLE_128B part of a head
constexpr int kQTile = 128;
constexpr int kSoftmaxWarps = 16;
constexpr int kThreads3 = (kSoftmaxWarps + 4) * 32;   // 640
constexpr int kQMainBytes = kQTile * kDM * 2;         // 16 KB
constexpr int kSlotCols = 256;
constexpr int kP0Col = 208;
constexpr int kMaxBkv = 208;
constexpr int kSmemLimit = 232448;

struct Attn3Params {
  int N, H;
  int nqt;            // query tiles per (image, head)
  int bkv, nblk;      // rows per KV block (multiple of 16, <= 208), number of KV blocks
  long long total_items;
  int reverse;        // walk the items from the last to the first (L2 reuse, see api.cu)
  float scale_log2;
  long long* dbg;     // optional cycle counters (developer tool tools/attn_dbg.py)
};

enum { A_QFULL = 0, A_QEMPTY, A_KFULL, A_KEMPTY, A_VFULL, A_VEMPTY, A_SFULL, A_PFULL, A_PFULL1, A_OFULL,
       A_OREAD, A_TURN, A_PER_SLOT };   // A_PFULL / A_PFULL1: probabilities of column half 0 / 1 are in TMEM
constexpr int A_NBARS = 2 * A_PER_SLOT;

// column split of a block of nj score columns: first half owns [0, cs), second half [cs, nj)
__device__ __forceinline__ int col_split(int nj) {
  int cs = (nj >> 1) & ~31;
  if (cs > 96) cs = 96;
  return cs;
}

// O_j (dh columns) must not overlap P1 = [cs, cs + (nj-cs)/2): columns [128, 128+dh) when P1 ends at or
// before 128, else columns [0, dh) (then cs = 96 >= dh)
__device__ __forceinline__ int out_col(int cs, int nj) { return (cs + ((nj - cs) >> 1) <= 128) ? 128 : 0; }

// shared-memory bytes of one slot: [Q main][K main][V main][Q tail][K tail][V tail][O staging]
__host__ __device__ constexpr int slot_bytes_for(int dh, int bkv) {
  return kQTile * dh * 2 + 2 * bkv * dh * 2 + 8 * 32 * (dh / 2) * 2;
}

template <int DH>
__global__ void __launch_bounds__(kThreads3, 1)
attn3_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o,
                 const __grid_constant__ CUtensorMap tma_qt, const __grid_constant__ CUtensorMap tma_kt,
                 const __grid_constant__ CUtensorMap tma_vt, const Attn3Params p) {
  constexpr int kDH = DH;
  constexpr int kDT = DH - kDM;                            // 0 or 16: columns of the SWIZZLE_32B part
  constexpr int kHalfCols = DH / 2;                        // output columns per softmax thread
  constexpr int kStageTile = 32 * kHalfCols * 2;           // staging bytes per warp
  static_assert(DH == 64 || DH == 80, "head dim 64 or 80");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int kv_main = p.bkv * kDM * 2;                   // SWIZZLE_128B part of one K or V block
  const int kv_tail = p.bkv * kDT * 2;                   // SWIZZLE_32B part
  const int slot_bytes = slot_bytes_for(DH, p.bkv);
  // then barriers, TMEM slot, row-statistics exchange
  const uint32_t bar_base = smem_base + 2 * slot_bytes;
  const uint32_t tmem_slot = bar_base + 8u * A_NBARS;
  const int misc_off = 2 * slot_bytes + 8 * A_NBARS;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + misc_off);
  // xchg_max / xchg_sum [slot][half][row]: partial row max (pass 1) and partial row sum (after the last block)
  float* xchg = reinterpret_cast<float*>(smem_gen + misc_off + 16);
  float* xchg_sum = xchg + 2 * 2 * kQTile;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool is_softmax = warp_idx < kSoftmaxWarps;
  const bool is_producer = warp_idx == 16 || warp_idx == 17;
  const bool is_mma = warp_idx == 18 || warp_idx == 19;
  const int g = is_softmax ? (warp_idx >> 3) : (warp_idx & 1);

  const uint32_t slot_smem = smem_base + g * slot_bytes;
  const uint32_t q_smem = slot_smem;
  const uint32_t k_smem = q_smem + kQMainBytes;
  const uint32_t v_smem = k_smem + kv_main;
  const uint32_t qt_smem = v_smem + kv_main;
  const uint32_t kt_smem = qt_smem + kQTile * kDT * 2;
  const uint32_t vt_smem = kt_smem + kv_tail;
  const uint32_t o_smem = vt_smem + kv_tail;
  const int o_off = g * slot_bytes + kQMainBytes + 2 * kv_main + kQTile * kDT * 2 + 2 * kv_tail;
  auto bar = [&](int i) { return bar_base + 8u * (g * A_PER_SLOT + i); };

  if (warp_idx == 19 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      const uint32_t b0 = bar_base + 8u * (s * A_PER_SLOT);
      for (int i = 0; i < A_PER_SLOT; ++i)
        mbar_init(b0 + 8u * i, (i == A_OREAD || i == A_TURN) ? 256 : (i == A_PFULL || i == A_PFULL1) ? 128 : 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 17 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_o);
    if (kDT) {
      tma_prefetch_desc(&tma_qt);
      tma_prefetch_desc(&tma_kt);
      tma_prefetch_desc(&tma_vt);
    }
  }
  if (warp_idx == 16) {
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
    const unsigned it32 = static_cast<unsigned>(item);          // total_items < 2^31 (checked on host)
    qt = static_cast<int>(it32 % static_cast<unsigned>(p.nqt));
    const unsigned bh = it32 / static_cast<unsigned>(p.nqt);
    head = static_cast<int>(bh % static_cast<unsigned>(p.H));
    img = static_cast<int>(bh / static_cast<unsigned>(p.H));
  };

  if (is_producer) {
    // ------------------------------------------------------------------ TMA producer of slot g
    uint32_t step = 0;
    for (int it = 0; it < n_local; ++it) {
      int img, head, qt;
      decode(it, img, head, qt);
      mbar_wait(bar(A_QEMPTY), (static_cast<uint32_t>(it) & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar(A_QFULL), kQTile * kDH * 2);
        tma_load_3d(&tma_q, bar(A_QFULL), q_smem, head * kDH, qt * kQTile, img, kEvictFirst);
        if (kDT) tma_load_3d(&tma_qt, bar(A_QFULL), qt_smem, head * kDH + kDM, qt * kQTile, img, kEvictFirst);
      }
      __syncwarp();
      for (int j = 0; j < nblk; ++j, ++step) {
        const uint32_t ph = step & 1u;
        mbar_wait(bar(A_KEMPTY), ph ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar(A_KFULL), kv_main + kv_tail);
          tma_load_3d(&tma_k, bar(A_KFULL), k_smem, head * kDH, j * bkv, img, kEvictNormal);
          if (kDT) tma_load_3d(&tma_kt, bar(A_KFULL), kt_smem, head * kDH + kDM, j * bkv, img, kEvictNormal);
        }
        __syncwarp();
        mbar_wait(bar(A_VEMPTY), ph ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar(A_VFULL), kv_main + kv_tail);
          tma_load_3d(&tma_v, bar(A_VFULL), v_smem, head * kDH, j * bkv, img, kEvictNormal);
          if (kDT) tma_load_3d(&tma_vt, bar(A_VFULL), vt_smem, head * kDH + kDM, j * bkv, img, kEvictNormal);
        }
        __syncwarp();
      }
    }
  } else if (is_mma) {
    // ------------------------------------------------------------------ MMA issuer of slot g
    uint32_t step = 0;
    for (int it = 0; it < n_local; ++it) {
      for (int j = 0; j < nblk; ++j, ++step) {
        const uint32_t ph = step & 1u;
        int nj = p.N - j * bkv;
        if (nj > bkv) nj = bkv;
        nj = (nj + 15) & ~15;
        const int cs = col_split(nj);
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
          for (int k = 0; k < kDM / 16; ++k) umma_ss(t_slot, qd + 2 * k, kd + 2 * k, idesc, k != 0 ? 1u : 0u);
          if (kDT)   // fifth K step on the 16-column SWIZZLE_32B tiles (rows of 32 B, 8-row atoms of 256 B)
            umma_ss(t_slot, make_smem_desc(qt_smem, 0, 256, 6), make_smem_desc(kt_smem, 0, 256, 6), idesc, 1u);
          umma_commit(bar(A_SFULL));
          umma_commit(bar(A_KEMPTY));
          if (j == nblk - 1) umma_commit(bar(A_QEMPTY));
        }
        __syncwarp();
        // ---- O_j = P V_j : keys [0, cs) come from P0, keys [cs, nj) from P1.  The two column halves
        // finish their exponentials at different times (96 vs 112 columns): the k steps over P0 are
        // issued as soon as half 0 has arrived.
        const uint32_t idesc_pv = make_idesc_bf16(kQTile, kDM, 0, 1);
        const uint32_t idesc_t = make_idesc_bf16(kQTile, 16, 0, 1);
        const uint64_t vd = make_desc_mnmajor_sw128(v_smem, 1024);
        const uint64_t vtd = make_smem_desc(vt_smem, 256, 256, 6);   // MN-major, SWIZZLE_32B
        const int ksteps = nj >> 4;
        const int k0 = cs >> 4;
        const uint32_t o_tmem = t_slot + out_col(cs, nj);
        // O may only be written early where it cannot overlap scores half 1 has not consumed yet
        const bool early_ok = (out_col(cs, nj) == 0) ? (cs >= kDH) : (nj <= 128);
        mbar_wait(bar(A_VFULL), ph);
        mbar_wait(bar(A_PFULL), ph);
        if (!early_ok) mbar_wait(bar(A_PFULL1), ph);
        tc_fence_after();
        if (elect_one_sync()) {
          for (int k = 0; k < k0; ++k) {   // 16 kv rows: 2048 B of the main tile, 512 B of the tail tile
            const uint32_t a_tmem = t_slot + kP0Col + 8 * k;
            umma_ts(o_tmem, a_tmem, vd + 128 * k, idesc_pv, k != 0 ? 1u : 0u);
            if (kDT) umma_ts(o_tmem + kDM, a_tmem, vtd + 32 * k, idesc_t, k != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
        if (early_ok) mbar_wait(bar(A_PFULL1), ph);
        tc_fence_after();
        if (elect_one_sync()) {
          for (int k = k0; k < ksteps; ++k) {
            const uint32_t a_tmem = t_slot + cs + 8 * (k - k0);
            umma_ts(o_tmem, a_tmem, vd + 128 * k, idesc_pv, k != 0 ? 1u : 0u);
            if (kDT) umma_ts(o_tmem + kDM, a_tmem, vtd + 32 * k, idesc_t, k != 0 ? 1u : 0u);
          }
          umma_commit(bar(A_OFULL));
          umma_commit(bar(A_VEMPTY));
        }
        __syncwarp();
      }
    }
  } else if (is_softmax) {
    // ------------------------------------------------------------------ softmax warps of slot g
    const int half = (warp_idx >> 2) & 1;
    const int quarter = warp_idx & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t t_lane = t_slot + (static_cast<uint32_t>(quarter * 32) << 16);
    const int wslot = (warp_idx & 7);                       // staging tile of this warp within the slot
    // [32 rows x 64 B] SWIZZLE_64B (dh 64) or [32 rows x 80 B] unswizzled (dh 80)
    const uint32_t stage_addr = o_smem + wslot * kStageTile;
    uint8_t* stage_row = smem_gen + o_off + wslot * kStageTile + lane * (kHalfCols * 2);
    const int sw = (kDT == 0) ? ((lane >> 1) & 3) : 0;
    float* my_x = xchg + (g * 2 + half) * kQTile + row_in_tile;
    const float* other_x = xchg + (g * 2 + (half ^ 1)) * kQTile + row_in_tile;
    float* my_l = xchg_sum + (g * 2 + half) * kQTile + row_in_tile;
    const float* other_l = xchg_sum + (g * 2 + (half ^ 1)) * kQTile + row_in_tile;
    const int bar_id = 1 + g;
    const uint32_t other_turn = bar_base + 8u * ((1 - g) * A_PER_SLOT + A_TURN);
    uint32_t step = 0;
    long long dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long dt0 = p.dbg ? clock64() : 0;
    long long tc = 0;
#define VT_TICK(i) if (p.dbg) { const long long t_ = clock64(); dacc[i] += t_ - tc; tc = t_; }

    for (int it = 0; it < n_local; ++it) {
      int img, head, qt;
      decode(it, img, head, qt);

      float o_acc[kHalfCols];
      float m_run = -INFINITY;
      float l_run = 0.f;      // this thread's partial row sum (its columns only)
      // warp-uniform: all 32 query rows of this warp lie beyond the sequence (N = 197: the last quarter
      // of every second tile).  Such a warp keeps every barrier protocol but skips the TMEM reads, the
      // exponentials and the output (its P rows stay undefined; MMA rows are independent).
      const bool live = qt * kQTile + quarter * 32 < p.N;

      for (int j = 0; j < nblk; ++j, ++step) {
        const uint32_t ph = step & 1u;
        int nvalid = p.N - j * bkv;
        if (nvalid > bkv) nvalid = bkv;
        const int nj = (nvalid + 15) & ~15;
        const int cs = col_split(nj);
        const int c0 = half ? cs : 0;
        const int c1 = half ? nj : cs;
        const uint32_t p_col = half ? (t_lane + cs) : (t_lane + kP0Col);

        if (p.dbg) tc = clock64();
        mbar_wait(bar(A_SFULL), ph);
        VT_TICK(0)
        tc_fence_after();

        // pass 1: partial row max, exchanged with the other half of the row
        const float mx = live ? row_max_part(t_lane, c0, c1, nvalid) : 0.f;
        *my_x = mx;
        VT_TICK(1)
        named_bar_sync(bar_id, 256);
        VT_TICK(2)
        const float m_new = fmaxf(m_run, fmaxf(mx, *other_x) * p.scale_log2);
        const float alpha = ex2_approx(m_run - m_new);   // first block: exp2(-inf) = 0
        m_run = m_new;

        // pass 2: exponentials of this thread's columns.  With two warps per scheduler a slot alone
        // saturates the MUFU pipe, so the slots take turns: while one is here the other does its
        // MUFU-free work (row max, O fold, epilogue, MMA waits).  Single-block items only: the item
        // counts of the two slots differ by at most one, which the token protocol tolerates.
        if (nblk == 1 && !(g == 0 && it == 0))
          mbar_wait(bar(A_TURN), static_cast<uint32_t>(g == 0 ? it - 1 : it) & 1u);
        VT_TICK(2)
        const float psum = live ? exp_part(t_lane, c0, c1, nvalid, p_col, p.scale_log2, m_new) : 1.f;
        if (nblk == 1) mbar_arrive(other_turn);
        l_run = l_run * alpha + psum;
        // partial row sum for the other half of the row: published before the arrive below, read
        // after the A_OFULL wait of the last block (PV MMAs are issued only after both halves arrived)
        if (j == nblk - 1) *my_l = l_run;
        tc_fence_before();
        mbar_arrive(bar(half ? A_PFULL1 : A_PFULL));
        VT_TICK(3)

        // O_j = P V_j lands in TMEM columns [0, 64): this thread folds its 32 output columns
        mbar_wait(bar(A_OFULL), ph);
        VT_TICK(4)
        tc_fence_after();
        float l_other = 0.f;
        if (j == nblk - 1) l_other = *other_l;
        if (live) {
          uint32_t r[32];
          uint32_t r8[8];
          const uint32_t o_src = t_lane + out_col(cs, nj) + half * kHalfCols;
          tmem_ld_32x32(o_src, r);
          if (kDT) tmem_ld_32x8(o_src + 32, r8);
          tmem_ld_wait();
          if (j == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[i] = __uint_as_float(r[i]);
            if (kDT) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o_acc[32 + i] = __uint_as_float(r8[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(r[i]));
            if (kDT) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o_acc[32 + i] = fmaf(o_acc[32 + i], alpha, __uint_as_float(r8[i]));
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar(A_OREAD));
        VT_TICK(5)
        if (j == nblk - 1) l_run += l_other;   // total row sum = both halves' partial sums
      }
      if (!live) continue;

      float inv;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(l_run));

      // stage this warp's [32 rows x 32 columns] as a SWIZZLE_64B tile and TMA-store it
      if (lane == 0) tma_store_wait_read<0>();   // previous store out of this staging tile is done
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < kHalfCols / 8; ++jj) {
        uint4 o4;
        o4.x = pack_bf16x2(o_acc[8 * jj + 0] * inv, o_acc[8 * jj + 1] * inv);
        o4.y = pack_bf16x2(o_acc[8 * jj + 2] * inv, o_acc[8 * jj + 3] * inv);
        o4.z = pack_bf16x2(o_acc[8 * jj + 4] * inv, o_acc[8 * jj + 5] * inv);
        o4.w = pack_bf16x2(o_acc[8 * jj + 6] * inv, o_acc[8 * jj + 7] * inv);
        *reinterpret_cast<uint4*>(stage_row + ((jj ^ sw) << 4)) = o4;
      }
      fence_proxy_async_smem();
      __syncwarp();
      const int row0 = qt * kQTile + quarter * 32;
      if (lane == 0 && row0 < p.N) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     :
                     : "l"(reinterpret_cast<uint64_t>(&tma_o)), "r"(stage_addr),
                       "r"(head * kDH + half * kHalfCols), "r"(row0), "r"(img)
                     : "memory");
        tma_store_commit();
      }
      VT_TICK(6)
    }
    if (lane == 0) tma_store_wait<0>();
    if (p.dbg && (warp_idx & 7) == 0 && lane == 0) {
      long long* d = p.dbg + (2LL * blockIdx.x + g) * 8;
      for (int i = 0; i < 7; ++i) d[i] = dacc[i];
      d[7] = clock64() - dt0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 16) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

long long* g_attn3_dbg = nullptr;

template <int DH>
int launch_attn3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                 const CUtensorMap& tqt, const CUtensorMap& tkt, const CUtensorMap& tvt, const Attn3Params& p,
                 int smem, cudaStream_t stream) {
  auto kern = attn3_fwd_kernel<DH>;
  static int granted[kMaxDevices] = {0};
  if (const int rc_attr = ensure_dynamic_smem(kern, smem, granted)) return rc_attr;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (p.total_items + 1) / 2 < sms ? (p.total_items + 1) / 2 : sms;
  kern<<<static_cast<unsigned>(grid), kThreads3, smem, stream>>>(tq, tk, tv, to, tqt, tkt, tvt, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

void attn3_set_debug_buffer(void* ptr) { g_attn3_dbg = static_cast<long long*>(ptr); }

int attn3_fwd_tcgen05(const void* q, const void* k, const void* v, void* out, int B, int H, int N,
                      int dh, long long qkv_row_stride, long long qkv_batch_stride,
                      long long out_row_stride, long long out_batch_stride, float scale, int reverse,
                      cudaStream_t stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || N <= 0) return VT_ERR_ARG;
  if (dh != 64 && dh != 80) return VT_ERR_UNSUPPORTED;
  if ((qkv_row_stride % 8) || (qkv_batch_stride % 8) || (out_row_stride % 8) || (out_batch_stride % 8))
    return VT_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return VT_ERR_ALIGN;

  Attn3Params p;
  p.N = N;
  p.H = H;
  p.nqt = (N + kQTile - 1) / kQTile;
  p.nblk = (N + kMaxBkv - 1) / kMaxBkv;
  int bkv = (N + p.nblk - 1) / p.nblk;
  p.bkv = (bkv + 15) & ~15;
  p.total_items = static_cast<long long>(B) * H * p.nqt;
  if (p.total_items >= (1LL << 31)) return VT_ERR_UNSUPPORTED;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.reverse = reverse;
  p.dbg = g_attn3_dbg;
  const int smem = 1024 + 2 * slot_bytes_for(dh, p.bkv) + 8 * A_NBARS + 16 + 2 * (2 * 2 * kQTile * 4);
  if (smem > kSmemLimit) return VT_ERR_UNSUPPORTED;

  const uint64_t cols = static_cast<uint64_t>(H) * dh;
  CUtensorMap tq, tk, tv, to, tqt, tkt, tvt;
  int rc = make_tmap_bf16_3d(&tq, q, cols, N, B, qkv_row_stride, qkv_batch_stride, kDM, kQTile, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tk, k, cols, N, B, qkv_row_stride, qkv_batch_stride, kDM, p.bkv, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tv, v, cols, N, B, qkv_row_stride, qkv_batch_stride, kDM, p.bkv, TMAP_SW_128);
  if (rc) return rc;
  if (dh == 64) {
    rc = make_tmap_bf16_3d(&to, out, cols, N, B, out_row_stride, out_batch_stride, 32, 32, TMAP_SW_64);
    if (rc) return rc;
    tqt = tq; tkt = tk; tvt = tv;   // unused
    return launch_attn3<64>(tq, tk, tv, to, tqt, tkt, tvt, p, smem, stream);
  }
  rc = make_tmap_bf16_3d(&tqt, q, cols, N, B, qkv_row_stride, qkv_batch_stride, 16, kQTile, TMAP_SW_32);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tkt, k, cols, N, B, qkv_row_stride, qkv_batch_stride, 16, p.bkv, TMAP_SW_32);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tvt, v, cols, N, B, qkv_row_stride, qkv_batch_stride, 16, p.bkv, TMAP_SW_32);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&to, out, cols, N, B, out_row_stride, out_batch_stride, 40, 32, TMAP_SW_NONE);
  if (rc) return rc;
  return launch_attn3<80>(tq, tk, tv, to, tqt, tkt, tvt, p, smem, stream);
}

}  // namespace vt
