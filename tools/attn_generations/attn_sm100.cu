// K3: fused multi-head attention forward on tcgen05 tensor cores (flash style: the N x N score
// matrix never leaves the SM).
//
//   ctx[b, i, h*dh:(h+1)*dh] = softmax_j( scale * q[b,i,h] . k[b,j,h] ) @ v[b,j,h]
//
// Replaces, for all heads at once, the reference's per-head chain
//   matmul3(q, k^T, scale) -> softmax -> matmul3(P, v) -> slice-assign into (B,N,D)
// (vit/vit.py:60-72,101-108; kernels matmul3.py:40-108, softmax.py:9-33), which writes and
// re-reads the (B,N,N) scores four times per head.
//
// One CTA = one (image, head, 128-query tile); 192 threads; two CTAs co-reside per SM so one CTA's
// softmax overlaps the other's MMAs.
//   warp 0      TMA producer (Q once; K_j, V_j per KV block) + TMEM alloc/dealloc
//   warp 1      MMA issuer:  S = Q K_j^T   (SS form, both operands K-major, SWIZZLE_128B)
//                            O_j = P V_j    (TS form: P bf16 in TMEM, V MN-major in smem)
//   warps 2-5   softmax: thread-per-row, S read from TMEM, P written back over S as packed bf16,
//               running (max, sum) and the output accumulator in registers (online softmax).
// TMEM (256 columns): S fp32 [0, nj) ; P bf16x2 [0, nj/2) aliasing S ; O_j fp32 [128, 128+dh).
#include "common.cuh"
#include "tensormap.h"

namespace vt {

namespace {

constexpr int kAttnThreads = 192;
constexpr int kQTile = 128;
constexpr int kTmemCols = 256;
constexpr int kOCol = 128;
constexpr int kMaxBKV = 256;

struct AttnParams {
  int N;        // tokens per image
  int bkv;      // rows per KV block (multiple of 16, <= 256)
  int nblk;     // number of KV blocks
  float scale_log2;
  __nv_bfloat16* out;
  long long out_row_stride, out_batch_stride;
};

// Barrier slots
enum { B_QFULL = 0, B_KFULL, B_VFULL, B_KEMPTY, B_VEMPTY, B_SFULL, B_PFULL, B_OFULL, B_OREAD,
       B_COUNT };

template <int DH>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_q,
                        const __grid_constant__ CUtensorMap tma_k,
                        const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
  static_assert(DH == 64, "SWIZZLE_128B head tiles are 64 bf16 wide");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t q_smem = smem_base;
  const uint32_t k_smem = q_smem + kQTile * DH * 2;
  const uint32_t v_smem = k_smem + p.bkv * DH * 2;
  const uint32_t bar_base = v_smem + p.bkv * DH * 2;
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const uint32_t tmem_slot = bar_base + 8u * B_COUNT;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + (kQTile * DH * 2) + 2 * (p.bkv * DH * 2) + 8 * B_COUNT);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQTile;
  const int head = blockIdx.y;
  const int img = blockIdx.z;

  if (warp_idx == 1 && lane == 0) {
    mbar_init(bar(B_QFULL), 1);
    mbar_init(bar(B_KFULL), 1);
    mbar_init(bar(B_VFULL), 1);
    mbar_init(bar(B_KEMPTY), 1);
    mbar_init(bar(B_VEMPTY), 1);
    mbar_init(bar(B_SFULL), 1);
    mbar_init(bar(B_PFULL), 128);
    mbar_init(bar(B_OFULL), 1);
    mbar_init(bar(B_OREAD), 128);
    fence_barrier_init();
  }
  if (warp_idx == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tma_q);
      tma_prefetch_desc(&tma_k);
      tma_prefetch_desc(&tma_v);
    }
    __syncwarp();
    tmem_alloc<kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int nblk = p.nblk;
  const int bkv = p.bkv;

  if (warp_idx == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(B_QFULL), kQTile * DH * 2);
      tma_load_3d(&tma_q, bar(B_QFULL), q_smem, head * DH, q0, img, kEvictFirst);
      for (int j = 0; j < nblk; ++j) {
        const uint32_t ph = static_cast<uint32_t>(j & 1);
        mbar_wait(bar(B_KEMPTY), ph ^ 1u);
        mbar_arrive_expect_tx(bar(B_KFULL), bkv * DH * 2);
        tma_load_3d(&tma_k, bar(B_KFULL), k_smem, head * DH, j * bkv, img, kEvictNormal);
        mbar_wait(bar(B_VEMPTY), ph ^ 1u);
        mbar_arrive_expect_tx(bar(B_VFULL), bkv * DH * 2);
        tma_load_3d(&tma_v, bar(B_VFULL), v_smem, head * DH, j * bkv, img, kEvictNormal);
      }
    }
  } else if (warp_idx == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      mbar_wait(bar(B_QFULL), 0);
      for (int j = 0; j < nblk; ++j) {
        const uint32_t ph = static_cast<uint32_t>(j & 1);
        int nj = p.N - j * bkv;
        if (nj > bkv) nj = bkv;
        nj = (nj + 15) & ~15;
        if (j > 0) mbar_wait(bar(B_OREAD), ph ^ 1u);
        mbar_wait(bar(B_KFULL), ph);
        tc_fence_after();
        {
          const uint32_t idesc = make_idesc_bf16(kQTile, nj, 0, 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) {
            umma_ss(tmem_base, make_desc_kmajor_sw128(q_smem + k * 32),
                    make_desc_kmajor_sw128(k_smem + k * 32), idesc, k != 0 ? 1u : 0u);
          }
          umma_commit(bar(B_SFULL));
          umma_commit(bar(B_KEMPTY));
        }
        mbar_wait(bar(B_PFULL), ph);
        mbar_wait(bar(B_VFULL), ph);
        tc_fence_after();
        {
          const uint32_t idesc = make_idesc_bf16(kQTile, DH, 0, 1);
          const int ksteps = nj / 16;
          for (int k = 0; k < ksteps; ++k) {
            umma_ts(tmem_base + kOCol, tmem_base + 8 * k,
                    make_desc_mnmajor_sw128(v_smem + k * 2048, 1024), idesc, k != 0 ? 1u : 0u);
          }
          umma_commit(bar(B_OFULL));
          umma_commit(bar(B_VEMPTY));
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + output
    const int quarter = warp_idx & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const int row = q0 + quarter * 32 + lane;

    float o_acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY;
    float l_run = 0.f;

    for (int j = 0; j < nblk; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j & 1);
      int nvalid = p.N - j * bkv;
      if (nvalid > bkv) nvalid = bkv;
      const int nj = (nvalid + 15) & ~15;

      mbar_wait(bar(B_SFULL), ph);
      tc_fence_after();

      // pass 1: row max
      float mx = -INFINITY;
      for (int c = 0; c < nj; c += 32) {
        if (c + 32 <= nj) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < nvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
          uint32_t r[16];
          tmem_ld_32x16(tmem_base + lane_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c + i < nvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = ex2_approx(m_run - m_new);  // first block: exp2(-inf) = 0
      m_run = m_new;

      // pass 2: p = exp2(s*scale - m), row sum, P -> TMEM (bf16x2 packed, aliasing S)
      float psum = 0.f;
      for (int c = 0; c < nj; c += 32) {
        if (c + 32 <= nj) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_addr + c, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_new));
            float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_new));
            if (c + i >= nvalid) p0 = 0.f;
            if (c + i + 1 >= nvalid) p1 = 0.f;
            psum += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x16(tmem_base + lane_addr + (c >> 1), pk);
        } else {
          uint32_t r[16];
          tmem_ld_32x16(tmem_base + lane_addr + c, r);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_new));
            float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_new));
            if (c + i >= nvalid) p0 = 0.f;
            if (c + i + 1 >= nvalid) p1 = 0.f;
            psum += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x8(tmem_base + lane_addr + (c >> 1), pk);
        }
      }
      l_run = l_run * alpha + psum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(B_PFULL));

      // O_j = P V_j arrives in TMEM; fold it into the register accumulator
      mbar_wait(bar(B_OFULL), ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DH; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + lane_addr + kOCol + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c + i] = fmaf(o_acc[c + i], alpha, __uint_as_float(r[i]));
      }
      if (j + 1 < nblk) {
        tc_fence_before();
        mbar_arrive(bar(B_OREAD));
      }
    }

    if (row < p.N) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* o = p.out + static_cast<long long>(img) * p.out_batch_stride +
                         static_cast<long long>(row) * p.out_row_stride + head * DH;
#pragma unroll
      for (int i = 0; i < DH; i += 8) {
        uint4 o4;
        o4.x = pack_bf16x2(o_acc[i + 0] * inv, o_acc[i + 1] * inv);
        o4.y = pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
        o4.z = pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv);
        o4.w = pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
        *reinterpret_cast<uint4*>(o + i) = o4;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 0) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace

// q, k, v: bf16 tensors viewed as [B, N, H*dh] with the given row / batch strides (elements);
// they may be three column slices of one fused-QKV buffer.  out: [B, N, H*dh] bf16.
int attn_fwd_tcgen05(const void* q, const void* k, const void* v, void* out, int B, int H, int N,
                     int dh, long long qkv_row_stride, long long qkv_batch_stride,
                     long long out_row_stride, long long out_batch_stride, float scale,
                     cudaStream_t stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || N <= 0) return VT_ERR_ARG;
  if (dh != 64) return VT_ERR_UNSUPPORTED;
  if ((qkv_row_stride % 8) || (qkv_batch_stride % 8) || (out_row_stride % 8) ||
      (out_batch_stride % 8))
    return VT_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15)
    return VT_ERR_ALIGN;
  if (H > 65535 || B > 65535) return VT_ERR_UNSUPPORTED;

  const int nblk = (N + kMaxBKV - 1) / kMaxBKV;
  int bkv = (N + nblk - 1) / nblk;
  bkv = (bkv + 15) & ~15;

  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bf16_3d(&tq, q, static_cast<uint64_t>(H) * dh, N, B, qkv_row_stride,
                             qkv_batch_stride, dh, kQTile, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tk, k, static_cast<uint64_t>(H) * dh, N, B, qkv_row_stride,
                         qkv_batch_stride, dh, bkv, TMAP_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tv, v, static_cast<uint64_t>(H) * dh, N, B, qkv_row_stride,
                         qkv_batch_stride, dh, bkv, TMAP_SW_128);
  if (rc) return rc;

  AttnParams p;
  p.N = N;
  p.bkv = bkv;
  p.nblk = nblk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_row_stride = out_row_stride;
  p.out_batch_stride = out_batch_stride;

  const int smem = 1024 + kQTile * dh * 2 + 2 * bkv * dh * 2 + 8 * B_COUNT + 16;
  auto kern = attn_fwd_tcgen05_kernel<64>;
  static int granted[kMaxDevices] = {0};
  if (const int rc_attr = ensure_dynamic_smem(kern, smem, granted)) return rc_attr;
  dim3 grid((N + kQTile - 1) / kQTile, H, B);
  kern<<<grid, kAttnThreads, smem, stream>>>(tq, tk, tv, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace vt
