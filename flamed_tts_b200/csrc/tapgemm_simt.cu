// fp32 FMA implicit-conv GEMM (CUDA cores).  The arithmetic of the parity mode (FLM_F32) and of
// every small/skinny projection (time embeddings, speaker projections, duration generators,
// FaCodec encoder): fp32 operands, fp32 FMA accumulation in ascending (tap, k) order.
//
//   out[b,t,n] = epi(sum_tap sum_k A[b, t*stride + off0 + tap*dil, k] * W[tap][n][k] + bias[n])
//
// Replaces the cuBLAS addmm / cuDNN conv1d calls behind nn.Linear / nn.Conv1d / ConvTranspose1d
// on the reference hot path.
#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {

// 128 x 64 block tile, 8 x 4 outputs per thread (two 4-row groups 64 rows apart), BK = 16; the next k-block's global
// loads are issued into registers before the current one is consumed (one __syncthreads pair per k-block, the
// loads' latency overlapped with 512 FMAs per thread).  Every output still accumulates in ascending (tap, k)
// order with one fmaf per term, so results are bit-identical to a plain dot-product loop.
constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;

__global__ void __launch_bounds__(256) tapgemm_simt_kernel(TapGemm p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const float* __restrict__ A = static_cast<const float*>(p.A);
  const float* __restrict__ W = static_cast<const float*>(p.W);
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)p.B * p.T_out;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // loader mapping: k-chunk = (tid % 4) * 4; A rows tid/4 and tid/4 + 64, W row tid/4
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  int lb[2], lt[2];
  bool lm_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t lm = m0 + lrow + 64 * h;
    lm_ok[h] = lm < M;
    lb[h] = lm_ok[h] ? (int)(lm / p.T_out) : 0;
    lt[h] = lm_ok[h] ? (int)(lm % p.T_out) : 0;
  }
  const int ln = n0 + lrow;
  const bool ln_ok = ln < p.N;
  // compute mapping: rows tr..tr+3 and tr+64..tr+67, columns tc..tc+3
  const int tr = (tid >> 4) * 4, tc = (tid & 15) * TN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int kblocks = (p.K + BK - 1) / BK;
  const int total = p.ntaps * kblocks;
  float av[2][4], wv[4];
  auto fetch = [&](int it) {  // global -> registers for iteration `it` = (tap, k-block)
    const int tap = it / kblocks, k = (it % kblocks) * BK + lk;
#pragma unroll
    for (int j = 0; j < 4; ++j) { av[0][j] = 0.f; av[1][j] = 0.f; wv[j] = 0.f; }
    if (k < p.K) {  // K % 4 == 0 is required, so a chunk is all-in or all-out
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int tin = lt[h] * p.stride + p.off0 + tap * p.dil;
        if (lm_ok[h] && tin >= 0 && tin < p.T_in) ld4<float>(A + ((int64_t)lb[h] * p.T_in + tin) * p.lda + k, av[h]);
      }
      if (ln_ok) ld4<float>(W + ((int64_t)tap * p.N + ln) * p.K + k, wv);
    }
  };
  fetch(0);
  for (int it = 0; it < total; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lrow] = av[0][j];
      As[lk + j][lrow + 64] = av[1][j];
      Bs[lk + j][lrow] = wv[j];
    }
    __syncthreads();
    if (it + 1 < total) fetch(it + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tr]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][tr + 64]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tc]);
      const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[TN] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + tr + (i & 3) + 64 * (i >> 2);
    if (m >= M) continue;
    const int b = (int)(m / p.T_out);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tc + j;
      if (n >= p.N) continue;
      float v = acc[i][j] + (p.bias ? p.bias[n] : 0.f);
      if (p.epi == EPI_GATE_RESID) {
        if (p.addend) {
          v += p.addend_bf16 ? ldf<bf16>(static_cast<const bf16*>(p.addend) + m * p.ld_add + n)
                             : ldf<float>(static_cast<const float*>(p.addend) + m * p.ld_add + n);
        }
        float* h = p.hres + m * p.ld_res + n;
        *h = __fadd_rn(*h, __fmul_rn(p.gate[(int64_t)b * p.gate_bstride + n], v));
        continue;
      }
      if (p.epi == EPI_EULER) {
        float* h = p.hres + m * p.ld_res + n;
        *h = __fadd_rn(*h, __fmul_rn(p.alpha, v));
        continue;
      }
      if (p.epi == EPI_RESID) {
        v += p.out_bf16 ? ldf<bf16>(static_cast<const bf16*>(p.resid_in) + m * p.ldc + n)
                        : ldf<float>(static_cast<const float*>(p.resid_in) + m * p.ldc + n);
      } else {
        v = epi_act(p.epi, v);
      }
      if (p.out_bf16)
        stf<bf16>(static_cast<bf16*>(p.out) + m * p.ldc + n, v);
      else
        stf<float>(static_cast<float*>(p.out) + m * p.ldc + n, v);
    }
  }
}

}  // namespace

void launch_tapgemm_simt(const TapGemm& p, cudaStream_t stream) {
  FLM_REQUIRE(p.K % 4 == 0 && p.lda % 4 == 0, "tapgemm_simt: K and lda must be multiples of 4");
  const int64_t M = (int64_t)p.B * p.T_out;
  if (M == 0 || p.N == 0) return;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
  tapgemm_simt_kernel<<<grid, 256, 0, stream>>>(p);
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
