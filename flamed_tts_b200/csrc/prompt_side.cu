// Prompt side of the FaCodec decoder: the factorised residual vector quantisers that turn the
// prompt encoder's output into the 6 code streams, and the small timbre transformer that pools it into the 256-d
// speaker vector.  Runs once per distinct prompt; fp32 throughout (the argmax must pick the reference's code).
//
// Reference: flamed/models/facodec/facodec.py:470-507, 521-533 (forward(vq=True)), quantize/fvq.py:35-116
// (FactorizedVectorQuantize: weight-normed in_proj 256->8, cosine-similarity argmax over 1024 codes, out_proj 8->256),
// quantize/rvq.py:27-73 (residual loop), transformer.py:86-234 (pre-LN encoder layers: MHA 4 x 64, conv k=5 FFN).
#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];  // fixed order: deterministic
  return t;
}

// One block (256 threads) per frame runs EVERY quantiser layer of the frame: frames are independent, the residual
// chain is not.  group 0 and 1 quantise x, group 2 quantises x - (q_group0 + q_group1) (facodec.py:476-490).
__global__ void __launch_bounds__(256) vq_frames_kernel(VqPlan plan, const float* __restrict__ x, int64_t rows, int D,
                                                        int64_t* __restrict__ codes, float* __restrict__ qgroups) {
  __shared__ float red[8];
  __shared__ float ze[VQ_MAX_CD], en[VQ_MAX_CD], zq[VQ_MAX_CD];
  __shared__ float best_v[8];
  __shared__ int best_i[8];
  const int64_t row = blockIdx.x;
  const int tid = threadIdx.x;
  constexpr int MAXC = 4;  // D <= 1024
  float xin[MAXC], res[MAXC], gsum[MAXC], prev[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = tid + i * 256;
    xin[i] = c < D ? x[row * D + c] : 0.f;
    prev[i] = 0.f;
  }
  int cur_group = -1;
  for (int l = 0; l < plan.n_layers; ++l) {
    const VqLayer& L = plan.layer[l];
    const int cd = L.cd;
    if (L.group != cur_group) {  // a new residual VQ starts (rvq.py:41-43)
      if (cur_group >= 0) {
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
          const int c = tid + i * 256;
          if (c < D) qgroups[((int64_t)cur_group * rows + row) * D + c] = gsum[i];
          if (cur_group < 2) prev[i] += gsum[i];
        }
      }
      cur_group = L.group;
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        res[i] = L.group == 2 ? xin[i] - prev[i] : xin[i];
        gsum[i] = 0.f;
      }
    }
    // z_e = in_proj(residual)  (fvq.py:60: weight-normed Linear D -> cd)
    for (int j = 0; j < cd; ++j) {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = tid + i * 256;
        if (c < D) part = fmaf(res[i], L.w_in[(int64_t)j * D + c], part);
      }
      const float s = block_sum_256(part, red);
      if (tid == 0) ze[j] = s + L.b_in[j];
    }
    __syncthreads();
    if (tid == 0) {  // F.normalize(z_e): x / max(||x||, 1e-12)  (fvq.py:101-103)
      float n2 = 0.f;
      for (int j = 0; j < cd; ++j) n2 = fmaf(ze[j], ze[j], n2);
      const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
      for (int j = 0; j < cd; ++j) en[j] = ze[j] * inv;
    }
    __syncthreads();
    float esq = 0.f;
    for (int j = 0; j < cd; ++j) esq = fmaf(en[j], en[j], esq);
    // dist = |e|^2 - 2 e.c + |c|^2 on the normalised codebook; index of the first maximum of -dist (fvq.py:105-112)
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int code = tid; code < L.n_codes; code += 256) {
      const float* cn = L.cb_norm + (int64_t)code * cd;
      float dot = 0.f;
      for (int j = 0; j < cd; ++j) dot = fmaf(en[j], cn[j], dot);
      const float nd = -((esq - 2.0f * dot) + L.cb_sq[code]);
      if (nd > bv) { bv = nd; bi = code; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { best_v[tid >> 5] = bv; best_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = best_v[0];
      int ix = best_i[0];
      for (int w = 1; w < 8; ++w)
        if (best_v[w] > v || (best_v[w] == v && best_i[w] < ix)) { v = best_v[w]; ix = best_i[w]; }
      best_i[0] = ix;
      codes[(int64_t)l * rows + row] = ix;
      // straight-through form kept as the reference evaluates it: z_q = z_e + (codebook[idx] - z_e)  (fvq.py:76)
      for (int j = 0; j < cd; ++j) zq[j] = ze[j] + (L.cb[(int64_t)ix * cd + j] - ze[j]);
    }
    __syncthreads();
    // q = out_proj(z_q); residual -= q; group sum += q  (fvq.py:78, rvq.py:52-54)
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = tid + i * 256;
      if (c < D) {
        float q = 0.f;
        for (int j = 0; j < cd; ++j) q = fmaf(zq[j], L.w_out[(int64_t)c * cd + j], q);
        q += L.b_out[c];
        res[i] -= q;
        gsum[i] += q;
      }
    }
    __syncthreads();
  }
  if (cur_group >= 0) {
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = tid + i * 256;
      if (c < D) qgroups[((int64_t)cur_group * rows + row) * D + c] = gsum[i];
    }
  }
}

// (B,C,T) -> (B,T,C), optionally a second copy with a per-sample row vector added (the timbre encoder's positional
// term, which the reference indexes by the BATCH axis: transformer.py:50-52 `x + pe[:x.size(0)]` on a (B,T,d) tensor)
__global__ void transpose_in_kernel(const float* __restrict__ x, int T_, int C, float* __restrict__ y,
                                    const float* __restrict__ pe, float* __restrict__ y_pe) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xb = x + (int64_t)b * T_ * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (t < T_ && c < C) tile[i][threadIdx.x] = xb[(int64_t)c * T_ + t];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T_ && c < C) {
      const float v = tile[threadIdx.x][i];
      const int64_t o = ((int64_t)b * T_ + t) * C + c;
      y[o] = v;
      if (y_pe) y_pe[o] = v + pe[(int64_t)b * C + c];
    }
  }
}

// Multi-head self-attention, fp32, no mask: qkv (B,T,3*H*DH) -> out (B,T,H*DH).  One warp per query; lane = key
// (keys lane, lane+32, ...) with an online softmax per lane, merged across the warp at the end.
template <int DH>
__global__ void __launch_bounds__(256) mha_fp32_kernel(const float* __restrict__ qkv, int T_, int H, float scale,
                                                       float* __restrict__ out) {
  __shared__ float qs[8][DH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int tq = blockIdx.x * 8 + warp;
  const int D = H * DH;
  const float* base = qkv + (int64_t)b * T_ * 3 * D;
  if (tq < T_) {
    for (int j = lane; j < DH; j += 32) qs[warp][j] = base[(int64_t)tq * 3 * D + h * DH + j] * scale;
  }
  __syncwarp();
  if (tq >= T_) return;
  float m = -INFINITY, l = 0.f, acc[DH];
#pragma unroll
  for (int j = 0; j < DH; ++j) acc[j] = 0.f;
  for (int tk = lane; tk < T_; tk += 32) {
    const float4* kp = reinterpret_cast<const float4*>(base + (int64_t)tk * 3 * D + D + h * DH);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < DH / 4; ++j) {
      const float4 k4 = __ldg(kp + j);
      s = fmaf(qs[warp][4 * j], k4.x, s); s = fmaf(qs[warp][4 * j + 1], k4.y, s);
      s = fmaf(qs[warp][4 * j + 2], k4.z, s); s = fmaf(qs[warp][4 * j + 3], k4.w, s);
    }
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn), p = expf(s - mn);
    l = l * corr + p;
    const float4* vp = reinterpret_cast<const float4*>(base + (int64_t)tk * 3 * D + 2 * D + h * DH);
#pragma unroll
    for (int j = 0; j < DH / 4; ++j) {
      const float4 v4 = __ldg(vp + j);
      acc[4 * j] = fmaf(acc[4 * j], corr, p * v4.x); acc[4 * j + 1] = fmaf(acc[4 * j + 1], corr, p * v4.y);
      acc[4 * j + 2] = fmaf(acc[4 * j + 2], corr, p * v4.z); acc[4 * j + 3] = fmaf(acc[4 * j + 3], corr, p * v4.w);
    }
    m = mn;
  }
  float mw = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  const float f = (m == -INFINITY) ? 0.f : expf(m - mw);
  l = warp_sum(l * f);
  const float inv = 1.0f / l;
  float* o = out + ((int64_t)b * T_ + tq) * D + h * DH;
#pragma unroll
  for (int j = 0; j < DH; ++j) {
    const float v = warp_sum(acc[j] * f);
    if (lane == (j & 31)) o[j] = v * inv;
  }
}

// (B,T,C) -> mean over T: (B,C)
__global__ void mean_time_kernel(const float* __restrict__ x, int T_, int C, float* __restrict__ y) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int t = 0; t < T_; ++t) s += x[((int64_t)b * T_ + t) * C + c];
  y[(int64_t)b * C + c] = s / (float)T_;
}

}  // namespace

void launch_vq_frames(const VqPlan& plan, const float* x, int64_t rows, int D, int64_t* codes, float* qgroups,
                      cudaStream_t stream) {
  FLM_REQUIRE(D <= 1024 && plan.n_layers <= VQ_MAX_LAYERS, "vq_frames: D <= 1024, <= 8 layers");
  for (int l = 0; l < plan.n_layers; ++l) FLM_REQUIRE(plan.layer[l].cd <= VQ_MAX_CD, "vq_frames: codebook_dim <= 16");
  if (rows == 0) return;
  vq_frames_kernel<<<(unsigned)rows, 256, 0, stream>>>(plan, x, rows, D, codes, qgroups);
  FLM_LAUNCH_CHECK();
}

void launch_transpose_in(const float* x, int B, int T, int C, float* y, const float* pe, float* y_pe, cudaStream_t stream) {
  if (B == 0 || T == 0) return;
  dim3 grid((T + 31) / 32, (C + 31) / 32, B);
  transpose_in_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, T, C, y, pe, y_pe);
  FLM_LAUNCH_CHECK();
}

void launch_mha_fp32(const float* qkv, int B, int T, int H, int DH, float* out, cudaStream_t stream) {
  FLM_REQUIRE(DH == 64 || DH == 32, "mha_fp32: head dim 32 or 64");
  if (B == 0 || T == 0) return;
  dim3 grid((T + 7) / 8, H, B);
  const float scale = 1.0f / sqrtf((float)DH);
  if (DH == 64) mha_fp32_kernel<64><<<grid, 256, 0, stream>>>(qkv, T, H, scale, out);
  else mha_fp32_kernel<32><<<grid, 256, 0, stream>>>(qkv, T, H, scale, out);
  FLM_LAUNCH_CHECK();
}

void launch_mean_time(const float* x, int B, int T, int C, float* y, cudaStream_t stream) {
  if (B == 0) return;
  dim3 grid((C + 127) / 128, B);
  mean_time_kernel<<<grid, 128, 0, stream>>>(x, T, C, y);
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
