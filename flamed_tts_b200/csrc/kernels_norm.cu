// Memory-bound normalisation / depthwise kernels of the denoiser and the cond down-sampler.
// All tensors are channels-last (rows = frames, C contiguous); storage fp32 or bf16, math fp32.
//
// Reference call sites (flamed/models/synthesizer/prob_generator.py):
//   ln_mod      <- nn.LayerNorm(eps=1e-6) + modulate():        136,146,162-163,229,239,257-259
//   dwconv      <- ConvNeXtBlock.conv_1 (depthwise k=31):      81-88,108-109
//   gn_finalize <- GroupNorm(C,C) / GroupNorm(8,.) statistics: 89,109; 15-17,187
//   gn_apply    <- GroupNorm affine (+ Mish/ReLU, mask, skip): 20-22,30-32,198-205
#include "common.cuh"
#include "kernels.h"

namespace flm {

// ------------------------------------------------------------------------------------ ln_mod
namespace {

constexpr int LN_MAX_V = 8;  // C <= 1024

__global__ void __launch_bounds__(256) ln_mod_kernel(LnMod p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.rows) return;
  const int64_t row = warp;
  const int nv = p.C >> 7;  // float4 per lane
  const float* x = p.x + row * p.ldx;
  float v[LN_MAX_V][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V; ++i) {
    if (i < nv) {
      ld4<float>(x + (i * 32 + lane) * 4, v[i]);
      if (p.relu_in) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = fmaxf(v[i][j], 0.f);
      }
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
  }
  const float mean = warp_sum(s) / (float)p.C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V; ++i) {
    if (i < nv) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.C + p.eps);
  const int bi = (int)(row / p.rows_per_batch);
#pragma unroll
  for (int i = 0; i < LN_MAX_V; ++i) {
    if (i < nv) {
      const int c = (i * 32 + lane) * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mean) * rstd;
      if (p.w) {
        float w[4], b[4];
        ld4<float>(p.w + c, w);
        ld4<float>(p.b + c, b);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf(o[j], w[j], b[j]);
      }
      if (p.scale) {
        float sc[4], sh[4];
        ld4<float>(p.scale + (int64_t)bi * p.mod_bstride + c, sc);
        ld4<float>(p.shift + (int64_t)bi * p.mod_bstride + c, sh);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf(o[j], p.scale_plus_one + sc[j], sh[j]);
      }
      if (p.y_bf16)
        st4<bf16>(static_cast<bf16*>(p.y) + row * p.ldy + c, o);
      else
        st4<float>(static_cast<float*>(p.y) + row * p.ldy + c, o);
    }
  }
}

}  // namespace

void launch_ln_mod(const LnMod& p, cudaStream_t stream) {
  FLM_REQUIRE(p.C % 128 == 0 && p.C <= 128 * LN_MAX_V, "ln_mod: C must be a multiple of 128 and <= 1024");
  if (p.rows == 0) return;
  const int warps_per_block = 8;
  const unsigned grid = (unsigned)((p.rows + warps_per_block - 1) / warps_per_block);
  ln_mod_kernel<<<grid, warps_per_block * 32, 0, stream>>>(p);
  FLM_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ dwconv
namespace {

template <typename T>
__device__ __forceinline__ void ld2(const T* p, float& a, float& b);
template <>
__device__ __forceinline__ void ld2<float>(const float* p, float& a, float& b) {
  float2 t = *reinterpret_cast<const float2*>(p);
  a = t.x; b = t.y;
}
template <>
__device__ __forceinline__ void ld2<bf16>(const bf16* p, float& a, float& b) {
  __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(p);
  a = __low2float(t); b = __high2float(t);
}
template <typename T>
__device__ __forceinline__ void st2(T* p, float a, float b);
template <>
__device__ __forceinline__ void st2<float>(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
template <>
__device__ __forceinline__ void st2<bf16>(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// thread = 2 adjacent channels x DW_TT consecutive frames; block = 128 threads = 256 channels.
// grid = (C/256, nchunk, B).  Zero padding at the two ends of the padded batch row range [0,L).
template <typename T, int KW>
__global__ void __launch_bounds__(128) dwconv_kernel(DwConv p) {
  constexpr int PAD = KW / 2;
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int t0 = chunk * DW_TT;
  if (c >= p.C) return;
  float w0[KW], w1[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    float2 t = *reinterpret_cast<const float2*>(p.w + (int64_t)k * p.C + c);
    w0[k] = t.x; w1[k] = t.y;
  }
  float a0[DW_TT], a1[DW_TT];
#pragma unroll
  for (int j = 0; j < DW_TT; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  const T* xb = static_cast<const T*>(p.x) + (int64_t)b * p.L * p.C + c;
#pragma unroll
  for (int r = 0; r < DW_TT + KW - 1; ++r) {
    const int t = t0 - PAD + r;
    float x0 = 0.f, x1 = 0.f;
    if (t >= 0 && t < p.L) ld2<T>(xb + (int64_t)t * p.C, x0, x1);
#pragma unroll
    for (int j = 0; j < DW_TT; ++j) {
      const int tap = r - j;  // compile-time after unrolling
      if (tap >= 0 && tap < KW) {
        a0[j] = fmaf(w0[tap], x0, a0[j]);
        a1[j] = fmaf(w1[tap], x1, a1[j]);
      }
    }
  }
  const float2 bias = *reinterpret_cast<const float2*>(p.bias + c);
  const int nvalid = min(DW_TT, p.L - t0);
  T* yb = static_cast<T*>(p.y) + (int64_t)b * p.L * p.C + c;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < DW_TT; ++j) {
    a0[j] += bias.x; a1[j] += bias.y;
    if (j < nvalid) {
      // statistics are taken on the values the next kernel will read back (storage precision)
      if (sizeof(T) == 2) {
        a0[j] = __bfloat162float(__float2bfloat16_rn(a0[j]));
        a1[j] = __bfloat162float(__float2bfloat16_rn(a1[j]));
      }
      st2<T>(yb + (int64_t)(t0 + j) * p.C, a0[j], a1[j]);
      s0 += a0[j]; s1 += a1[j];
    }
  }
  const float m0 = s0 / (float)nvalid, m1 = s1 / (float)nvalid;
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int j = 0; j < DW_TT; ++j) {
    if (j < nvalid) {
      const float d0 = a0[j] - m0, d1 = a1[j] - m1;
      q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
    }
  }
  float* part = p.part + (((int64_t)b * gridDim.y + chunk) * p.C + c) * 2;
  *reinterpret_cast<float4*>(part) = make_float4(m0, q0, m1, q1);
}

}  // namespace

void launch_dwconv(const DwConv& p, cudaStream_t stream) {
  FLM_REQUIRE(p.KW == 31, "dwconv: only kernel_size 31 is compiled (configs/prob.yaml convnext.kernel_size)");
  FLM_REQUIRE(p.C % 256 == 0, "dwconv: C must be a multiple of 256");
  if (p.B == 0 || p.L == 0) return;
  dim3 grid(p.C / 256, dw_nchunk(p.L), p.B);
  if (p.io_bf16)
    dwconv_kernel<bf16, 31><<<grid, 128, 0, stream>>>(p);
  else
    dwconv_kernel<float, 31><<<grid, 128, 0, stream>>>(p);
  FLM_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ group stats
namespace {

// grid = (nchunk, B, G); block 256: (mean, M2) of x[b, chunk rows, group channels]
template <typename T>
__global__ void __launch_bounds__(256) group_stats_kernel(const T* x, int L, int C, int G, float* part) {
  __shared__ float red[8];
  __shared__ float bcast;
  const int chunk = blockIdx.x, b = blockIdx.y, g = blockIdx.z;
  const int gs = C / G;
  const int t0 = chunk * GS_ROWS;
  const int nrows = min(GS_ROWS, L - t0);
  const int n = nrows * gs;
  const T* base = x + ((int64_t)b * L + t0) * C + (int64_t)g * gs;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += ldf<T>(base + (int64_t)(i / gs) * C + (i % gs));
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = t / (float)n;
  }
  __syncthreads();
  const float mean = bcast;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = ldf<T>(base + (int64_t)(i / gs) * C + (i % gs)) - mean;
    q = fmaf(d, d, q);
  }
  q = warp_sum(q);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    float* o = part + (((int64_t)b * gridDim.x + chunk) * G + g) * 2;
    o[0] = mean; o[1] = t;
  }
}

// thread per (b, c): merge the chunk partials of c's group with Chan's update in fp64
__global__ void gn_finalize_kernel(const float* part, int B, int L, int C, int G, int nchunk, int chunk_rows,
                                   const float* gamma, const float* beta, float eps, float* scale, float* offset) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx % C;
  const int gs = C / G, g = c / gs;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int k = 0; k < nchunk; ++k) {
    const int rows = min(chunk_rows, L - k * chunk_rows);
    const double nb = (double)rows * gs;
    const float* pp = part + (((int64_t)b * nchunk + k) * G + g) * 2;
    const double mb = pp[0], qb = pp[1];
    const double tot = n + nb;
    const double delta = mb - mean;
    mean += delta * nb / tot;
    m2 += qb + delta * delta * n * nb / tot;
    n = tot;
  }
  const double var = m2 / n;  // biased
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rstd;
  scale[idx] = sc;
  offset[idx] = beta[c] - (float)mean * sc;
}

template <typename TX, typename TY, typename TR>
__global__ void __launch_bounds__(256) gn_apply_kernel(GnApply p) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4n = p.C >> 2;
  const int64_t total = (int64_t)p.B * p.L * c4n;
  if (i4 >= total) return;
  const int c = (int)(i4 % c4n) * 4;
  const int64_t row = i4 / c4n;
  const int b = (int)(row / p.L);
  float v[4], sc[4], of[4];
  ld4<TX>(static_cast<const TX*>(p.x) + row * p.C + c, v);
  ld4<float>(p.scale + (int64_t)b * p.C + c, sc);
  ld4<float>(p.offset + (int64_t)b * p.C + c, of);
  const float m = p.mask ? (float)p.mask[row] : 1.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float y = fmaf(v[j], sc[j], of[j]);
    if (p.act == 1) y = fmaxf(y, 0.f);
    if (p.act == 2) y = mish(y);
    v[j] = y * m;
  }
  if (p.res) {
    float r[4];
    ld4<TR>(static_cast<const TR*>(p.res) + row * p.C + c, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += r[j];
  }
  st4<TY>(static_cast<TY*>(p.y) + row * p.C + c, v);
}

}  // namespace

void launch_group_stats(const void* x, int x_bf16, int B, int L, int C, int G, float* part, cudaStream_t stream) {
  FLM_REQUIRE(C % G == 0, "group_stats: C % G != 0");
  if (B == 0 || L == 0) return;
  dim3 grid(gs_nchunk(L), B, G);
  if (x_bf16)
    group_stats_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(x), L, C, G, part);
  else
    group_stats_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), L, C, G, part);
  FLM_LAUNCH_CHECK();
}

void launch_gn_finalize(const float* part, int B, int L, int C, int G, int nchunk, int chunk_rows, const float* gamma,
                        const float* beta, float eps, float* scale, float* offset, cudaStream_t stream) {
  if (B == 0) return;
  const int n = B * C;
  gn_finalize_kernel<<<(n + 255) / 256, 256, 0, stream>>>(part, B, L, C, G, nchunk, chunk_rows, gamma, beta, eps, scale,
                                                          offset);
  FLM_LAUNCH_CHECK();
}

void launch_gn_apply(const GnApply& p, cudaStream_t stream) {
  FLM_REQUIRE(p.C % 4 == 0, "gn_apply: C % 4 != 0");
  const int64_t total = (int64_t)p.B * p.L * (p.C / 4);
  if (total == 0) return;
  const unsigned grid = (unsigned)((total + 255) / 256);
  const int key = (p.x_bf16 ? 1 : 0) | (p.y_bf16 ? 2 : 0) | ((p.res && p.res_bf16) ? 4 : 0);
  switch (key) {
    case 0: gn_apply_kernel<float, float, float><<<grid, 256, 0, stream>>>(p); break;
    case 1: gn_apply_kernel<bf16, float, float><<<grid, 256, 0, stream>>>(p); break;
    case 2: gn_apply_kernel<float, bf16, float><<<grid, 256, 0, stream>>>(p); break;
    case 3: gn_apply_kernel<bf16, bf16, float><<<grid, 256, 0, stream>>>(p); break;
    case 4: gn_apply_kernel<float, float, bf16><<<grid, 256, 0, stream>>>(p); break;
    case 5: gn_apply_kernel<bf16, float, bf16><<<grid, 256, 0, stream>>>(p); break;
    case 6: gn_apply_kernel<float, bf16, bf16><<<grid, 256, 0, stream>>>(p); break;
    default: gn_apply_kernel<bf16, bf16, bf16><<<grid, 256, 0, stream>>>(p); break;
  }
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
