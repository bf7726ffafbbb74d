// Pack-time kernel: fold a LayerNorm (gamma, beta) into the K-major bf16 weight of the dense layer that
// consumes it (vt_ln_fold).  One CTA per output row n:
//
//   wg[k]     = W[n,k] * gamma[k]                                   (fp32)
//   bias'[n]  = bias[n] + sum_k W[n,k] * beta[k]                    (accumulated in fp64)
//   plain:      W'[n,k] = bf16(wg[k]),             colsum[n] = sum_k W'[n,k]
//   zero-sum:   W'[n,k] = bf16(wg[k] - mean_k wg), then the rounding residue  r = sum_k W'[n,k]  (a few
//               1e-3 after plain rounding of a 768-wide row) is cancelled by re-rounding a few elements
//               the other way, cheapest first: the price of moving element k by one bf16 ulp is
//               ulp_k -/+ 2 |err_k| (err_k = exact - rounded; elements rounded the "wrong" way by almost
//               half an ulp are nearly free).  Per pass the set {price <= T} with the largest T whose ulps
//               sum to <= |r| is found by bisection over the bit pattern of T (prices are >= 0, so the
//               fp32 bit pattern is monotone); every element moves at most once; six passes leave
//               |r| at the 1e-6 level.
//
// With zero-sum rows the mean term of  LN(x) W^T = rstd * (x W'^T) - rstd * mean * colsum  comes out of
// the tensor cores, so the consuming GEMM's epilogue is  rstd * acc + bias'  (gemm2_sm100.cu, EPI_NOCS).
// Reference semantics folded: layernorm_kernel (vit/kernels/layernorm.py:51-85) followed by
// matmul_kernel (vit/kernels/matmul.py:73-108) as called from Transformer.forward (vit/vit.py:133-144).
// Everything is deterministic (fixed-order tree reductions in fp64).
#include "common.cuh"

namespace vt {

namespace {

constexpr int kFoldThreads = 256;

// fixed-order block reduction (sum) of doubles; every thread gets the result
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                       // red[] may still be read from the previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kFoldThreads / 32; ++w) t += red[w];
  return t;
}

__device__ __forceinline__ unsigned block_min_u32(unsigned v, unsigned* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  unsigned t = 0xFFFFFFFFu;
#pragma unroll
  for (int w = 0; w < kFoldThreads / 32; ++w) t = min(t, red[w]);
  return t;
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// one bf16 ulp of f: 2^(exponent - 7), at least the smallest normal
__device__ __forceinline__ float bf16_ulp(float f) {
  int e = static_cast<int>((__float_as_uint(f) >> 23) & 0xFFu) - 7;
  if (e < 1) e = 1;
  return __uint_as_float(static_cast<unsigned>(e) << 23);
}

__global__ void __launch_bounds__(kFoldThreads)
ln_fold_kernel(const __nv_bfloat16* __restrict__ w, long long ldw, const float* __restrict__ bias,
               const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ w_out,
               long long ldo, float* __restrict__ bias_out, float* __restrict__ colsum_out, int K, int zero_sum,
               int passes) {
  extern __shared__ uint8_t fold_smem[];
  float* exact = reinterpret_cast<float*>(fold_smem);                 // [K] target value (fp32)
  float* cur = exact + K;                                             // [K] current bf16 value (as fp32)
  uint8_t* moved = reinterpret_cast<uint8_t*>(cur + K);               // [K]
  __shared__ double red_d[kFoldThreads / 32];
  __shared__ unsigned red_u[kFoldThreads / 32];

  const int n = blockIdx.x;
  const int tid = threadIdx.x;
  const __nv_bfloat16* wr = w + static_cast<long long>(n) * ldw;

  double s_wg = 0.0, s_wb = 0.0;
  for (int k = tid; k < K; k += kFoldThreads) {
    const float wv = __bfloat162float(wr[k]);
    const float wg = wv * gamma[k];
    exact[k] = wg;
    s_wg += static_cast<double>(wg);
    s_wb += static_cast<double>(wv) * static_cast<double>(beta[k]);
  }
  s_wg = block_sum(s_wg, red_d);
  s_wb = block_sum(s_wb, red_d);
  if (tid == 0) bias_out[n] = static_cast<float>((bias ? static_cast<double>(bias[n]) : 0.0) + s_wb);

  const float mean = zero_sum ? static_cast<float>(s_wg / static_cast<double>(K)) : 0.f;
  for (int k = tid; k < K; k += kFoldThreads) {
    const float e = exact[k] - mean;
    exact[k] = e;
    cur[k] = bf16_round(e);
    moved[k] = 0;
  }

  if (zero_sum) {
    for (int pass = 0; pass < passes; ++pass) {
      double part = 0.0;
      for (int k = tid; k < K; k += kFoldThreads) part += static_cast<double>(cur[k]);
      const double resid_d = block_sum(part, red_d);        // exact: bf16 values, < 2^13 of them
      const float resid = static_cast<float>(resid_d);
      const float a = fabsf(resid);
      if (a == 0.f) break;                                    // block-uniform
      // sum of the ulps of the movable elements whose price is <= T (bit pattern), as a function of T
      auto ulps_below = [&](unsigned T, double* count) {
        double s = 0.0, c = 0.0;
        for (int k = tid; k < K; k += kFoldThreads) {
          const float f = cur[k];
          const float ulp = bf16_ulp(f);
          if (moved[k] || !(ulp <= a)) continue;
          const float err = exact[k] - f;
          const float price = fmaxf(ulp + ((err * resid < 0.f) ? -2.f : 2.f) * fabsf(err), 0.f);
          if (__float_as_uint(price) <= T) { s += static_cast<double>(ulp); c += 1.0; }
        }
        const double tot = block_sum(s, red_d);
        if (count) *count = block_sum(c, red_d);
        return tot;
      };
      double movable = 0.0;
      const double all = ulps_below(0x7F7FFFFFu, &movable);
      if (movable == 0.0) break;                             // nothing left that is small enough to help
      unsigned T;
      bool any = true;
      if (all <= static_cast<double>(a)) {
        T = 0x7F7FFFFFu;
      } else {
        // largest T with ulps_below(T) <= a; lo = "take nothing" sentinel handled through `any`
        unsigned lo = 0u, hi = 0x7F7FFFFFu;                  // invariant: ulps_below(hi) > a
        const bool lo_ok = ulps_below(0u, nullptr) <= static_cast<double>(a);
        if (!lo_ok) {
          any = false;
          T = 0u;
        } else {
          while (hi - lo > 1u) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            if (ulps_below(mid, nullptr) <= static_cast<double>(a)) lo = mid; else hi = mid;
          }
          T = lo;
          double cnt = 0.0;
          ulps_below(T, &cnt);
          if (cnt == 0.0) any = false;                       // T lies below the cheapest price
        }
      }
      const float sgn = resid > 0.f ? 1.f : -1.f;
      if (any) {
        for (int k = tid; k < K; k += kFoldThreads) {
          const float f = cur[k];
          const float ulp = bf16_ulp(f);
          if (moved[k] || !(ulp <= a)) continue;
          const float err = exact[k] - f;
          const float price = fmaxf(ulp + ((err * resid < 0.f) ? -2.f : 2.f) * fabsf(err), 0.f);
          if (__float_as_uint(price) <= T) {
            cur[k] = bf16_round(f - sgn * ulp);              // exact: one ulp of f
            moved[k] = 1;
          }
        }
      }
      // the set can be empty although movable elements exist (T below the cheapest price, or a tie group
      // at the cheapest price too heavy as a whole): move the single cheapest element, lowest index first
      if (!any) {
        unsigned best = 0xFFFFFFFFu;
        for (int k = tid; k < K; k += kFoldThreads) {
          const float f = cur[k];
          const float ulp = bf16_ulp(f);
          if (moved[k] || !(ulp <= a)) continue;
          const float err = exact[k] - f;
          const float price = fmaxf(ulp + ((err * resid < 0.f) ? -2.f : 2.f) * fabsf(err), 0.f);
          best = min(best, __float_as_uint(price));
        }
        best = block_min_u32(best, red_u);
        unsigned idx = 0xFFFFFFFFu;
        for (int k = tid; k < K; k += kFoldThreads) {
          const float f = cur[k];
          const float ulp = bf16_ulp(f);
          if (moved[k] || !(ulp <= a)) continue;
          const float err = exact[k] - f;
          const float price = fmaxf(ulp + ((err * resid < 0.f) ? -2.f : 2.f) * fabsf(err), 0.f);
          if (__float_as_uint(price) == best) idx = min(idx, static_cast<unsigned>(k));
        }
        idx = block_min_u32(idx, red_u);
        if (idx != 0xFFFFFFFFu && (idx % kFoldThreads) == static_cast<unsigned>(tid)) {
          const float f = cur[idx];
          cur[idx] = bf16_round(f - sgn * bf16_ulp(f));
          moved[idx] = 1;
        }
      }
      __syncthreads();
    }
  }

  __nv_bfloat16* orow = w_out + static_cast<long long>(n) * ldo;
  double part = 0.0;
  for (int k = tid; k < K; k += kFoldThreads) {
    orow[k] = __float2bfloat16_rn(cur[k]);                   // exact
    part += static_cast<double>(cur[k]);
  }
  if (colsum_out != nullptr) {
    const double cs = block_sum(part, red_d);
    if (tid == 0) colsum_out[n] = static_cast<float>(cs);
  }
}

}  // namespace

int ln_fold(const void* w, long long ldw, const float* bias, const float* gamma, const float* beta, void* w_out,
            long long ldo, float* bias_out, float* colsum_out, int N, int K, int zero_sum, cudaStream_t stream) {
  if (!w || !gamma || !beta || !w_out || !bias_out || N <= 0 || K <= 0 || ldw < K || ldo < K) return VT_ERR_ARG;
  const int smem = K * 9;
  if (smem > 200 * 1024) return VT_ERR_UNSUPPORTED;
  static int granted[kMaxDevices] = {0};
  if (const int rc = ensure_dynamic_smem(ln_fold_kernel, smem, granted)) return rc;
  ln_fold_kernel<<<N, kFoldThreads, smem, stream>>>(static_cast<const __nv_bfloat16*>(w), ldw, bias, gamma, beta,
                                                   static_cast<__nv_bfloat16*>(w_out), ldo, bias_out, colsum_out, K,
                                                   zero_sum, 6);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace vt
