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

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

__global__ void __launch_bounds__(256) tapgemm_simt_kernel(TapGemm p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const float* __restrict__ A = static_cast<const float*>(p.A);
  const float* __restrict__ W = static_cast<const float*>(p.W);
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)p.B * p.T_out;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // loader mapping: row = tid / 4 (0..63), k-chunk = (tid % 4) * 4
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int64_t lm = m0 + lrow;
  const bool lm_ok = lm < M;
  const int lb = lm_ok ? (int)(lm / p.T_out) : 0;
  const int lt = lm_ok ? (int)(lm % p.T_out) : 0;
  const int ln = n0 + lrow;
  const bool ln_ok = ln < p.N;
  // compute mapping
  const int tr = (tid >> 4) * TM, tc = (tid & 15) * TN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < p.ntaps; ++tap) {
    const int tin = lt * p.stride + p.off0 + tap * p.dil;
    const bool a_ok = lm_ok && tin >= 0 && tin < p.T_in;
    const float* arow = A + ((int64_t)lb * p.T_in + (a_ok ? tin : 0)) * p.lda;
    const float* wrow = W + ((int64_t)tap * p.N + (ln_ok ? ln : 0)) * p.K;
    for (int k0 = 0; k0 < p.K; k0 += BK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + lk;
      if (k < p.K) {  // K % 4 == 0 is required, so a chunk is all-in or all-out
        if (a_ok) ld4<float>(arow + k, av);
        if (ln_ok) ld4<float>(wrow + k, wv);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[lk + j][lrow] = av[j];
        Bs[lk + j][lrow] = wv[j];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][tr + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tc + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + tr + i;
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
