// K2: patch embedding as an im2col-free tcgen05 GEMM with the CLS / position-embedding add fused
// into the epilogue.
//
//   out[b, 0, :]     = cls + pos[0]
//   out[b, 1 + p, :] = sum_k patch(b,p)[k] * W[:, k] + bias + pos[1 + p]      k = (c, i, j)
//
// Replaces conv2d_kernel (reference vit/kernels/conv2d.py:19-97: one program per (image, output
// channel, output row), scalar tl.sum, every patch re-read D times) plus the torch glue after it
// (flatten/transpose, cat(cls), + position_embeddings: vit/vit.py:190-200).
//
// Pixel formats: NCHW fp32 / bf16 with k = (c, i, j) — the reference's layout — and NHWC uint8 with
// k = (i, j, c) (camera / decoder output: the ViTImageProcessor rescale + normalise is folded into
// the packed weights and the bias table on the host, so the kernel only widens bytes to bf16,
// which is exact for 0..255).
//
// A (the patches) is never materialised: four producer warps gather 128 patches x 64 k straight
// from the NCHW pixels into the SWIZZLE_128B K-major smem layout the UMMA descriptor expects
// (TMA cannot express a 14-pixel, 28-byte inner box, and C=3 rules out im2col-mode TMA).
// B (conv weight viewed as [D, 3*P*P], already K-major) arrives by TMA.
//   warps 0-7   A gather producers (256 threads, half a tile row each, software-pipelined through a
//               register ring so several K blocks of pixel loads are in flight), then epilogue
//   warp 8      TMA producer for B + TMEM alloc
//   warp 9      MMA issuer (M=128, N=256, K=16), warp-uniform loop with one elected lane
// Two CTAs are co-resident per SM (2 smem stages each), so one CTA's epilogue overlaps the other's
// gather / MMA phase.
#include "common.cuh"
#include "tensormap.h"

namespace vt {

namespace {

constexpr int PE_BM = 128;
constexpr int PE_BN = 256;
constexpr int PE_BK = 64;
constexpr int PE_STAGES = 2;
constexpr int PE_GATHER_THREADS = 256;
constexpr int PE_THREADS = PE_GATHER_THREADS + 64;
constexpr int PE_WARP_TMA = 8;
constexpr int PE_WARP_MMA = 9;
constexpr int PE_A_BYTES = PE_BM * PE_BK * 2;
constexpr int PE_B_BYTES = PE_BN * PE_BK * 2;
constexpr int PE_STAGE_BYTES = PE_A_BYTES + PE_B_BYTES;
constexpr int PE_SMEM = 1024 + PE_STAGES * PE_STAGE_BYTES + 256;

struct PatchParams {
  const void* pixels;   // [B, C, S, S]
  int B, C, S, P;
  int grid_w;           // S / P
  int n_patches;        // grid_w^2
  int K;                // C * P * P
  int D;
  const float* posb;    // [n_patches + 1, D] fp32: pos + (cls | conv bias)
  void* out;            // [B, n_patches + 1, D]
  int out_f32;
  // optional: (sum, M2 about the group mean) of every 128-column group of every output row, [(B * (n_patches + 1)), D / 128, 2]
  // — the row statistics the LayerNorm folded into the first QKV GEMM consumes (gemm2_sm100.cu)
  float* stats;
};

__device__ __forceinline__ float px_to_f(float v) { return v; }
__device__ __forceinline__ float px_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float px_to_f(uint8_t v) { return static_cast<float>(v); }
// byte b of w as an exact float: 0x4B0000bb is 2^23 + b
__device__ __forceinline__ float byte_to_f(uint32_t w, int b) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 + b)) - 8388608.0f;
}

// Raw (unconverted) pixels of one 8-element K chunk, as loaded: converting at load time would make
// the thread wait for the load and defeat the prefetch ring.
template <typename TPix, bool kVec>
struct RawChunk;
template <>
struct RawChunk<__nv_bfloat16, true> {
  uint4 v;
  __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void load(const __nv_bfloat16* src) { v = __ldg(reinterpret_cast<const uint4*>(src)); }
  __device__ __forceinline__ uint4 packed() const { return v; }
};
template <>
struct RawChunk<float, true> {
  float4 a, b;
  __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
  __device__ __forceinline__ void load(const float* src) {
    a = __ldg(reinterpret_cast<const float4*>(src));
    b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  }
  __device__ __forceinline__ uint4 packed() const {
    return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
};
template <>
struct RawChunk<uint8_t, true> {
  uint2 v;
  __device__ __forceinline__ void zero() { v = make_uint2(0u, 0u); }
  __device__ __forceinline__ void load(const uint8_t* src) { v = __ldg(reinterpret_cast<const uint2*>(src)); }
  __device__ __forceinline__ uint4 packed() const {
    return make_uint4(pack_bf16x2(byte_to_f(v.x, 0), byte_to_f(v.x, 1)), pack_bf16x2(byte_to_f(v.x, 2), byte_to_f(v.x, 3)),
                      pack_bf16x2(byte_to_f(v.y, 0), byte_to_f(v.y, 1)), pack_bf16x2(byte_to_f(v.y, 2), byte_to_f(v.y, 3)));
  }
};
template <typename TPix>
struct RawChunk<TPix, false> {
  TPix e[8];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = TPix(0);
  }
  __device__ __forceinline__ uint4 packed() const {
    return make_uint4(pack_bf16x2(px_to_f(e[0]), px_to_f(e[1])), pack_bf16x2(px_to_f(e[2]), px_to_f(e[3])),
                      pack_bf16x2(px_to_f(e[4]), px_to_f(e[5])), pack_bf16x2(px_to_f(e[6]), px_to_f(e[7])));
  }
};

template <typename TPix, bool kVec>
__global__ void __launch_bounds__(PE_THREADS, 2)
patch_embed_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_w, const PatchParams p) {
  constexpr int NPF = (kVec && sizeof(TPix) <= 2) ? 3 : 2;   // K blocks of pixel loads in flight
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t bar_addr = smem_base + PE_STAGES * PE_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_addr + 8u * s; };
  auto empty_bar = [&](int s) { return bar_addr + 8u * (PE_STAGES + s); };
  const uint32_t acc_bar = bar_addr + 8u * (2 * PE_STAGES);
  const uint32_t tmem_slot = bar_addr + 8u * (2 * PE_STAGES + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + PE_STAGES * PE_STAGE_BYTES + 8 * (2 * PE_STAGES + 1));

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x;
  const int m_blk = blockIdx.y;
  const int num_kb = (p.K + PE_BK - 1) / PE_BK;

  if (warp_idx == PE_WARP_MMA && lane == 0) {
    for (int s = 0; s < PE_STAGES; ++s) {
      mbar_init(full_bar(s), 1 + PE_GATHER_THREADS);  // TMA expect_tx arrive + gather threads
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp_idx == PE_WARP_TMA) {
    if (lane == 0) tma_prefetch_desc(&tma_w);
    __syncwarp();
    tmem_alloc<PE_BN>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp_idx == PE_WARP_TMA) {
    int s = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(empty_bar(s), phase ^ 1u);
      if (elect_one_sync()) {
        const uint32_t b_dst = smem_base + s * PE_STAGE_BYTES + PE_A_BYTES;
        mbar_arrive_expect_tx(full_bar(s), PE_B_BYTES);
        tma_load_2d(&tma_w, full_bar(s), b_dst, kb * PE_BK, n_blk * PE_BN, kEvictLast);
      }
      __syncwarp();
      if (++s == PE_STAGES) { s = 0; phase ^= 1u; }
    }
  } else if (warp_idx == PE_WARP_MMA) {
    constexpr uint32_t idesc = make_idesc_bf16(PE_BM, PE_BN, 0, 0);
    int s = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(full_bar(s), phase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t ad = make_desc_kmajor_sw128(smem_base + s * PE_STAGE_BYTES);
        const uint64_t bd = make_desc_kmajor_sw128(smem_base + s * PE_STAGE_BYTES + PE_A_BYTES);
#pragma unroll
        for (int k = 0; k < PE_BK / 16; ++k)
          umma_ss(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(s));
      }
      __syncwarp();
      if (++s == PE_STAGES) { s = 0; phase ^= 1u; }
    }
    if (elect_one_sync()) umma_commit(acc_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ gather producers (256 threads)
    const int r = threadIdx.x & 127;     // tile row
    const int hh = threadIdx.x >> 7;     // which half of the 64-wide K block: chunks 4*hh .. 4*hh+3
    const long long m = static_cast<long long>(m_blk) * PE_BM + r;
    const bool row_valid = m < static_cast<long long>(p.B) * p.n_patches;
    const int img = row_valid ? static_cast<int>(m / p.n_patches) : 0;
    const int patch = row_valid ? static_cast<int>(m - static_cast<long long>(img) * p.n_patches) : 0;
    const int py = patch / p.grid_w;
    const int px = patch - py * p.grid_w;
    constexpr bool kNHWC = sizeof(TPix) == 1;   // uint8 pixels are [B, S, S, C], k = (i, j, c)
    const TPix* img_base = static_cast<const TPix*>(p.pixels) +
                           static_cast<long long>(img) * p.C * p.S * p.S +
                           (kNHWC ? (static_cast<long long>(py) * p.P * p.S + px * p.P) * p.C
                                  : static_cast<long long>(py) * p.P * p.S + px * p.P);
    const int PP = p.P * p.P;
    const int PC = p.P * p.C;

    RawChunk<TPix, kVec> ring[NPF][4];
    auto issue_loads = [&](RawChunk<TPix, kVec> (&dst)[4], int kb) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int k0 = kb * PE_BK + (hh * 4 + ch) * 8;
        dst[ch].zero();
        if (row_valid && k0 < p.K) {
          if constexpr (kVec && kNHWC) {
            const int i = k0 / PC;          // patch row; the 8 bytes stay inside it: (P*C) % 8 == 0
            const int rem = k0 - i * PC;
            dst[ch].load(img_base + static_cast<long long>(i) * p.S * p.C + rem);
          } else if constexpr (kVec) {
            const int c = k0 / PP;
            const int rem = k0 - c * PP;
            const int i = rem / p.P;
            const int j = rem - i * p.P;
            dst[ch].load(img_base + (static_cast<long long>(c) * p.S + i) * p.S + j);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int k = k0 + e;
              if (k < p.K) {
                if constexpr (kNHWC) {
                  const int i = k / PC;
                  const int rem = k - i * PC;   // j * C + c
                  dst[ch].e[e] = img_base[static_cast<long long>(i) * p.S * p.C + rem];
                } else {
                  const int c = k / PP;
                  const int rem = k - c * PP;
                  const int i = rem / p.P;
                  const int j = rem - i * p.P;
                  dst[ch].e[e] = img_base[(static_cast<long long>(c) * p.S + i) * p.S + j];
                }
              }
            }
          }
        }
      }
    };
    int s = 0;
    uint32_t phase = 0;
    auto publish = [&](const RawChunk<TPix, kVec> (&src)[4]) {
      mbar_wait(empty_bar(s), phase ^ 1u);
      uint8_t* a_row = smem_gen + s * PE_STAGE_BYTES + r * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(a_row + (((hh * 4 + ch) ^ (r & 7)) << 4)) = src[ch].packed();
      fence_proxy_async_smem();  // make generic-proxy smem writes visible to the tensor core
      mbar_arrive(full_bar(s));
      if (++s == PE_STAGES) { s = 0; phase ^= 1u; }
    };

#pragma unroll
    for (int u = 0; u < NPF - 1; ++u)
      if (u < num_kb) issue_loads(ring[u], u);
    for (int kb0 = 0; kb0 < num_kb; kb0 += NPF) {
#pragma unroll
      for (int u = 0; u < NPF; ++u) {
        const int kb = kb0 + u;
        if (kb < num_kb) {
          if (kb + NPF - 1 < num_kb) issue_loads(ring[(u + NPF - 1) % NPF], kb + NPF - 1);
          publish(ring[u]);
        }
      }
    }

    // ------------------------------------------------------------ epilogue (8 warps)
    const int quarter = warp_idx & 3;       // TMEM lane quarter this warp may read
    const int chalf = warp_idx >> 2;        // column half of the tile
    const int er = quarter * 32 + lane;
    const long long em = static_cast<long long>(m_blk) * PE_BM + er;
    const bool e_valid = em < static_cast<long long>(p.B) * p.n_patches;
    const int e_img = e_valid ? static_cast<int>(em / p.n_patches) : 0;
    const int e_patch = e_valid ? static_cast<int>(em - static_cast<long long>(e_img) * p.n_patches) : 0;
    const int n_tok = p.n_patches + 1;
    const long long out_row = static_cast<long long>(e_img) * n_tok + 1 + e_patch;
    const float* posb_row = p.posb + static_cast<long long>(1 + e_patch) * p.D;

    // 16 columns per step; the position/bias values of the NEXT step are loaded before this step's
    // accumulator columns are read and stored (the epilogue was latency-bound on these loads: half
    // of all stall samples, profiles/r01d), and the first step's loads are issued before the
    // accumulator is even complete.
    constexpr int kStep = 16;
    const int col_base = n_blk * PE_BN + chalf * (PE_BN / 2);
    auto load_posb = [&](float4 (&pb)[4], int col) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e_valid) {
          const int c4 = col + 4 * i;
          if (c4 + 3 < p.D) {
            pb[i] = __ldg(reinterpret_cast<const float4*>(posb_row + c4));
          } else {
            if (c4 + 0 < p.D) pb[i].x = posb_row[c4 + 0];
            if (c4 + 1 < p.D) pb[i].y = posb_row[c4 + 1];
            if (c4 + 2 < p.D) pb[i].z = posb_row[c4 + 2];
          }
        }
      }
    };
    float4 pb_cur[4], pb_nxt[4];
    // this warp's 128 columns of the row: sums of (x - pivot) and (x - pivot)^2 (see gemm2_sm100.cu, EPI_STATS)
    float2 st_sum2 = make_float2(0.f, 0.f), st_sq2 = make_float2(0.f, 0.f), st_npiv2 = make_float2(0.f, 0.f);
    load_posb(pb_cur, col_base);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < PE_BN / 2; cc += kStep) {
      const int c = chalf * (PE_BN / 2) + cc;
      const int col = n_blk * PE_BN + c;
      uint32_t rr[16];
      tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c, rr);
      if (cc + kStep < PE_BN / 2) load_posb(pb_nxt, col + kStep);
      tmem_ld_wait();
      if (e_valid && col < p.D) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // packed fp32 pairs: one FADD2 per two columns
          const float2 lo = __fadd2_rn(make_float2(__uint_as_float(rr[4 * i + 0]), __uint_as_float(rr[4 * i + 1])),
                                       make_float2(pb_cur[i].x, pb_cur[i].y));
          const float2 hi = __fadd2_rn(make_float2(__uint_as_float(rr[4 * i + 2]), __uint_as_float(rr[4 * i + 3])),
                                       make_float2(pb_cur[i].z, pb_cur[i].w));
          v[4 * i + 0] = lo.x; v[4 * i + 1] = lo.y;
          v[4 * i + 2] = hi.x; v[4 * i + 3] = hi.y;
          if (cc == 0 && i == 0) st_npiv2 = make_float2(-lo.x, -lo.x);   // pivot: the row's first value in the group
          const float2 dlo = __fadd2_rn(lo, st_npiv2), dhi = __fadd2_rn(hi, st_npiv2);
          st_sum2 = __fadd2_rn(st_sum2, __fadd2_rn(dlo, dhi));
          st_sq2 = __ffma2_rn(dlo, dlo, st_sq2);
          st_sq2 = __ffma2_rn(dhi, dhi, st_sq2);
        }
        const bool full_chunk = col + kStep <= p.D;
        if (p.out_f32) {
          float* o = static_cast<float*>(p.out) + out_row * p.D + col;
          if (full_chunk) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            for (int i = 0; i < 16 && col + i < p.D; ++i) o[i] = v[i];
          }
        } else {
          __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + out_row * p.D + col;
          if (full_chunk) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              uint4 o4;
              o4.x = pack_bf16x2(v[i + 0], v[i + 1]);
              o4.y = pack_bf16x2(v[i + 2], v[i + 3]);
              o4.z = pack_bf16x2(v[i + 4], v[i + 5]);
              o4.w = pack_bf16x2(v[i + 6], v[i + 7]);
              *reinterpret_cast<uint4*>(o + i) = o4;
            }
          } else {
            for (int i = 0; i < 16 && col + i < p.D; ++i) o[i] = __float2bfloat16_rn(v[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) pb_cur[i] = pb_nxt[i];
    }

    if (p.stats != nullptr && e_valid && col_base < p.D) {   // D % 128 == 0 (host): the group is all in or all out
      float2* dst = reinterpret_cast<float2*>(p.stats) + out_row * (p.D >> 7) + (col_base >> 7);
      const float s_sh = st_sum2.x + st_sum2.y, q_sh = st_sq2.x + st_sq2.y;
      *dst = make_float2(fmaf(-128.0f, st_npiv2.x, s_sh), fmaxf(fmaf(-s_sh * (1.0f / 128.0f), s_sh, q_sh), 0.f));
    }

    // CLS rows: the first row-tile of every column block writes cls + pos[0] for all images.
    if (m_blk == 0 && p.stats != nullptr) {
      // their statistics: the same two 128-column groups for every image (one thread per (image, group))
      const int groups = min(PE_BN, p.D - n_blk * PE_BN) >> 7;
      for (int idx = threadIdx.x; idx < p.B * groups; idx += PE_GATHER_THREADS) {
        const int b = idx / groups;
        const int gidx = idx - b * groups;
        const float* src = p.posb + n_blk * PE_BN + gidx * 128;
        float sx = 0.f, sq = 0.f;
        const float piv = __ldg(src);
        for (int c = 0; c < 128; ++c) {
          const float d = __ldg(src + c) - piv;
          sx += d;
          sq = fmaf(d, d, sq);
        }
        float2* dst = reinterpret_cast<float2*>(p.stats) + static_cast<long long>(b) * n_tok * (p.D >> 7) +
                      ((n_blk * PE_BN) >> 7) + gidx;
        *dst = make_float2(fmaf(128.0f, piv, sx), fmaxf(fmaf(-sx * (1.0f / 128.0f), sx, sq), 0.f));
      }
    }
    if (m_blk == 0) {
      const int ncols = min(PE_BN, p.D - n_blk * PE_BN);
      const int total = p.B * ncols;
      for (int idx = threadIdx.x; idx < total; idx += PE_GATHER_THREADS) {
        const int b = idx / ncols;
        const int c = idx - b * ncols;
        const int col = n_blk * PE_BN + c;
        const float val = p.posb[col];
        const long long off = static_cast<long long>(b) * n_tok * p.D + col;
        if (p.out_f32) static_cast<float*>(p.out)[off] = val;
        else static_cast<__nv_bfloat16*>(p.out)[off] = __float2bfloat16_rn(val);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == PE_WARP_TMA) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<PE_BN>(tmem_base);
  }
}

template <typename TPix, bool kVec>
int launch_patch(const CUtensorMap& tw, const PatchParams& p, dim3 grid, cudaStream_t stream) {
  auto kern = patch_embed_tcgen05_kernel<TPix, kVec>;
  static int granted[kMaxDevices] = {0};
  if (const int rc_attr = ensure_dynamic_smem(kern, PE_SMEM, granted)) return rc_attr;
  kern<<<grid, PE_THREADS, PE_SMEM, stream>>>(tw, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

// pixels [B,C,S,S] (fp32 or bf16, k = (c,i,j)) or [B,S,S,C] (uint8, k = (i,j,c)); w [D, C*P*P] bf16 in
// the matching k order with row stride ldw (elements, multiple of 8), posb [n+1, D] fp32,
// out [B, n+1, D] bf16|f32.
int patch_embed_tcgen05(const void* pixels, int pix_dtype, const void* w, long long ldw,
                        const float* posb, void* out, int out_dtype, float* stats, int B, int C, int S, int P,
                        int D, cudaStream_t stream) {
  if (!pixels || !w || !posb || !out || B <= 0 || C <= 0 || S <= 0 || P <= 0 || D <= 0)
    return VT_ERR_ARG;
  if (S % P) return VT_ERR_ARG;
  if ((ldw % 8) || (D % 8)) return VT_ERR_ALIGN;
  if (stats && (D % 128)) return VT_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(stats) & 7) return VT_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(pixels) | reinterpret_cast<uintptr_t>(w) |
       reinterpret_cast<uintptr_t>(posb) | reinterpret_cast<uintptr_t>(out)) & 15)
    return VT_ERR_ALIGN;
  const int K = C * P * P;
  CUtensorMap tw;
  int rc = make_tmap_bf16_2d(&tw, w, K, D, ldw, PE_BK, PE_BN, TMAP_SW_128);
  if (rc) return rc;

  PatchParams p;
  p.pixels = pixels;
  p.B = B; p.C = C; p.S = S; p.P = P;
  p.grid_w = S / P;
  p.n_patches = p.grid_w * p.grid_w;
  p.K = K;
  p.D = D;
  p.posb = posb;
  p.out = out;
  p.out_f32 = (out_dtype == VT_F32);
  p.stats = stats;
  if (out_dtype != VT_F32 && out_dtype != VT_BF16) return VT_ERR_DTYPE;

  const long long M = static_cast<long long>(B) * p.n_patches;
  dim3 grid((D + PE_BN - 1) / PE_BN, static_cast<unsigned>((M + PE_BM - 1) / PE_BM));
  if (grid.y > 65535) return VT_ERR_UNSUPPORTED;
  const bool vec = (P % 8 == 0) && (S % 8 == 0);
  if (pix_dtype == VT_F32)
    return vec ? launch_patch<float, true>(tw, p, grid, stream) : launch_patch<float, false>(tw, p, grid, stream);
  if (pix_dtype == VT_BF16)
    return vec ? launch_patch<__nv_bfloat16, true>(tw, p, grid, stream)
               : launch_patch<__nv_bfloat16, false>(tw, p, grid, stream);
  if (pix_dtype == VT_U8) {
    // 8-byte loads: every patch row (P*C bytes) and every image row (S*C bytes) is a multiple of 8
    const bool vec8 = ((P * C) % 8 == 0) && ((S * C) % 8 == 0);
    return vec8 ? launch_patch<uint8_t, true>(tw, p, grid, stream) : launch_patch<uint8_t, false>(tw, p, grid, stream);
  }
  return VT_ERR_DTYPE;
}

}  // namespace vt
