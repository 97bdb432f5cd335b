// FaCodec memory-bound kernels (channels-last): the anti-aliased Snake activation, the 64->1
// output conv + tanh, the 1->32 input conv of the encoder and the final layout transpose.
//
// Reference: flamed/models/facodec/alias_free_torch/act.py:24-29, resample.py:28-37,54-57,
// filter.py:89-96, facodec.py:105-118 (SnakeBeta); closed form in SURVEY.md Appendix A1:
//   u[2n]   = 2*sum_{j<6} x[clamp(n-3+j)] * fu[11-2j]
//   u[2n+1] = 2*sum_{j<6} x[clamp(n-2+j)] * fu[10-2j]
//   s[m]    = u[m] + sin^2(a*u[m]) * invb
//   y[n]    = sum_{k<12} s[clamp(2n+k-5, 0, 2T-1)] * fd[k]
#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {


template <typename T>
__device__ __forceinline__ void ldpair(const T* p, float& a, float& b);
template <>
__device__ __forceinline__ void ldpair<float>(const float* p, float& a, float& b) {
  float2 t = *reinterpret_cast<const float2*>(p);
  a = t.x; b = t.y;
}
template <>
__device__ __forceinline__ void ldpair<bf16>(const bf16* p, float& a, float& b) {
  __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(p);
  a = __low2float(t); b = __high2float(t);
}
template <typename T>
__device__ __forceinline__ void stpair(T* p, float a, float b);
template <>
__device__ __forceinline__ void stpair<float>(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
template <>
__device__ __forceinline__ void stpair<bf16>(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

template <bool FAST>
__device__ __forceinline__ float snake(float u, float a, float invb) {
  const float sn = FAST ? __sinf(u * a) : sinf(u * a);
  return fmaf(sn * sn, invb, u);
}
// both channels of a thread at once: s = u + sin^2(a u) * invb on a packed pair
template <bool FAST>
__device__ __forceinline__ f32x2 snake2(f32x2 u2, f32x2 a2, f32x2 ib2) {
  float t0, t1;
  unpack2(mul2(u2, a2), t0, t1);
  const float s0 = FAST ? __sinf(t0) : sinf(t0), s1 = FAST ? __sinf(t1) : sinf(t1);
  if (FAST) {  // one packed multiply for both squares (the exact path keeps two scalar multiplies: same values, fp32)
    const f32x2 sn = pack2(s0, s1);
    return fma2(mul2(sn, sn), ib2, u2);
  }
  return fma2(pack2(s0 * s0, s1 * s1), ib2, u2);
}

// u at up-sampled index m (0 <= m < 2T) from the clamped input window xw[i] = x[clamp(n0-5+i)]
// (static indices only).  q = m - (2*n0 - 5) is the local index; fu2[k] = (2 fu[k], 2 fu[k]).
template <int Q, int ACT_TT>
__device__ __forceinline__ f32x2 upsample_at(const f32x2 (&xw)[ACT_TT + 10], const f32x2 (&fu2)[12]) {
  f32x2 acc = 0ull;
  if ((Q & 1) == 0) {
    // odd m = 2j+1, j = n0 - 3 + Q/2: taps x[j-2+i] -> window index Q/2 + i
#pragma unroll
    for (int i = 0; i < 6; ++i) acc = fma2(xw[Q / 2 + i], fu2[10 - 2 * i], acc);
  } else {
    // even m = 2j, j = n0 - 3 + (Q+1)/2: taps x[j-3+i] -> window index (Q+1)/2 - 1 + i
#pragma unroll
    for (int i = 0; i < 6; ++i) acc = fma2(xw[(Q + 1) / 2 - 1 + i], fu2[11 - 2 * i], acc);
  }
  return acc;
}

// EDGE: the thread's window touches the ends of the signal (replicate padding of s); interior threads skip the two
// compare + select pairs per s value (6 ALU instructions each, ~17 % of the instruction stream)
template <bool FAST, bool EDGE, int ACT_TT, int Q>
struct SFill {
  __device__ static __forceinline__ void run(const f32x2 (&xw)[ACT_TT + 10], const f32x2 (&fu2)[12], f32x2 a2, f32x2 ib2,
                                             int mbase, int twoT, f32x2 sf, f32x2 sl, f32x2 (&s)[2 * ACT_TT + 10]) {
    f32x2 v = snake2<FAST>(upsample_at<Q, ACT_TT>(xw, fu2), a2, ib2);
    if (EDGE) {
      const int m = mbase + Q;
      if (m < 0) v = sf;
      if (m > twoT - 1) v = sl;
    }
    s[Q] = v;
    SFill<FAST, EDGE, ACT_TT, Q + 1>::run(xw, fu2, a2, ib2, mbase, twoT, sf, sl, s);
  }
};
template <bool FAST, bool EDGE, int ACT_TT>
struct SFill<FAST, EDGE, ACT_TT, 2 * ACT_TT + 10> {
  __device__ static __forceinline__ void run(const f32x2 (&)[ACT_TT + 10], const f32x2 (&)[12], f32x2, f32x2, int, int,
                                             f32x2, f32x2, f32x2 (&)[2 * ACT_TT + 10]) {}
};

template <typename T>
__device__ __forceinline__ f32x2 ld_pair2(const T* p) {
  float a, b;
  ldpair<T>(p, a, b);
  return pack2(a, b);
}

// blockDim = (bx channel pairs, by time runs); grid = (C/2/bx, ceil(T/(by*ACT_TT)), B).  The two channels of a
// thread travel as packed fp32x2 (FFMA2).
// CC: compile-time channel count (row stride: the 26 loads and 16 stores of the interior path then use immediate
// offsets, no address arithmetic in the instruction stream) or 0 for a run-time C
template <typename T, bool FAST, int ACT_TT, int MINB, int CC>
__global__ void __launch_bounds__(256, MINB) act1d_kernel(Act1d p) {
  pdl_trigger();
  pdl_wait();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int n0 = (blockIdx.y * blockDim.y + threadIdx.y) * ACT_TT;
  const int b = blockIdx.z;
  if (c >= p.C || n0 >= p.T) return;
  const int Tm1 = p.T - 1;
  const T* xb = static_cast<const T*>(p.x) + (int64_t)b * p.T * p.C + c;
  f32x2 fu2[12], fd2[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    fu2[k] = pack2(2.0f * p.fu[k], 2.0f * p.fu[k]);  // up-sampler gain 2 folded in (exact)
    fd2[k] = pack2(p.fd[k], p.fd[k]);
  }
  const f32x2 a2 = *reinterpret_cast<const f32x2*>(p.a + c);
  const f32x2 ib2 = *reinterpret_cast<const f32x2*>(p.invb + c);
  const int mbase = 2 * n0 - 5;
  const int twoT = 2 * p.T;
  T* yb = static_cast<T*>(p.y) + (int64_t)b * p.T * p.C + c;
  f32x2 xw[ACT_TT + 10];
  f32x2 s[2 * ACT_TT + 10];
  // interior: the input window [n0-5, n0+ACT_TT+4] and the s window [2 n0 - 5, 2 n0 + 2 ACT_TT + 4] lie inside the signal
  // (warp-uniform: the lanes of a warp are channel pairs of the same time run)
  if (n0 - 5 >= 0 && n0 + ACT_TT + 4 <= Tm1) {
    const int64_t Cs = CC ? CC : p.C;
    const T* xr = xb + (int64_t)(n0 - 5) * Cs;
#pragma unroll
    for (int i = 0; i < ACT_TT + 10; ++i) xw[i] = ld_pair2<T>(xr + i * Cs);
    SFill<FAST, false, ACT_TT, 0>::run(xw, fu2, a2, ib2, mbase, twoT, 0ull, 0ull, s);
    T* yr = yb + (int64_t)n0 * Cs;
#pragma unroll
    for (int j = 0; j < ACT_TT; ++j) {
      f32x2 y = 0ull;
#pragma unroll
      for (int k = 0; k < 12; ++k) y = fma2(s[2 * j + k], fd2[k], y);
      float y0, y1;
      unpack2(y, y0, y1);
      stpair<T>(yr + j * Cs, y0, y1);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < ACT_TT + 10; ++i) {
    const int t = min(max(n0 - 5 + i, 0), Tm1);
    xw[i] = ld_pair2<T>(xb + (int64_t)t * p.C);
  }
  // replicate padding of the activated up-sampled signal: s[-k] = s[0], s[2T-1+k] = s[2T-1]
  f32x2 sf = 0ull, sl = 0ull;
  if (mbase < 0) {  // s[0] = snake(u[0]), u[0] = 2*sum x[clamp(-3+i)] fu[11-2i]
    f32x2 u = 0ull;
#pragma unroll
    for (int i = 0; i < 6; ++i) u = fma2(ld_pair2<T>(xb + (int64_t)min(max(i - 3, 0), Tm1) * p.C), fu2[11 - 2 * i], u);
    sf = snake2<FAST>(u, a2, ib2);
  }
  if (mbase + 2 * ACT_TT + 9 > twoT - 1) {  // s[2T-1] = snake(u[2(T-1)+1])
    f32x2 u = 0ull;
#pragma unroll
    for (int i = 0; i < 6; ++i) u = fma2(ld_pair2<T>(xb + (int64_t)min(max(Tm1 - 2 + i, 0), Tm1) * p.C), fu2[10 - 2 * i], u);
    sl = snake2<FAST>(u, a2, ib2);
  }
  SFill<FAST, true, ACT_TT, 0>::run(xw, fu2, a2, ib2, mbase, twoT, sf, sl, s);
#pragma unroll
  for (int j = 0; j < ACT_TT; ++j) {
    if (n0 + j < p.T) {
      f32x2 y = 0ull;
#pragma unroll
      for (int k = 0; k < 12; ++k) y = fma2(s[2 * j + k], fd2[k], y);
      float y0, y1;
      unpack2(y, y0, y1);
      stpair<T>(yb + (int64_t)(n0 + j) * p.C, y0, y1);
    }
  }
}

// thread per output sample: tanh(sum_tap sum_c x[t+tap-3,c] w[tap][c] + bias); C % 8 == 0, C <= 128
template <typename T>
__global__ void __launch_bounds__(256) conv_out_tanh_kernel(const T* x, const float* w, float bias, int T_, int C,
                                                            float* wav) {
  __shared__ float ws[7 * 128];
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T_) return;
  const T* xb = x + (int64_t)b * T_ * C;
  float acc = 0.f;
#pragma unroll
  for (int tap = 0; tap < 7; ++tap) {
    const int tt = t + tap - 3;
    if (tt < 0 || tt >= T_) continue;
    const T* row = xb + (int64_t)tt * C;
    for (int c = 0; c < C; c += 4) {
      float v[4];
      ld4<T>(row + c, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc = fmaf(v[j], ws[tap * C + c + j], acc);
    }
  }
  wav[(int64_t)b * T_ + t] = tanhf(acc + bias);
}

// bf16 form, memory-bound: LPR = C/8 lanes share a frame (16 B = 8 channels each, so a warp reads 32/LPR whole
// frames per fully coalesced instruction), every lane keeps its 7 x 8 tap weights in registers, forms the 7
// partial dot products of its frame, the LPR lanes are summed with xor shuffles, and the per-(tap, frame) dots
// go through shared memory: wav[t] = tanh(bias + sum_tap dot[tap][t + tap - 3]).  block = COT_T frames.
constexpr int COT_T = 256;
template <int LPR>
__global__ void __launch_bounds__(256) conv_out_tanh_bf16_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                                 float bias, int T_, float* __restrict__ wav) {
  constexpr int C = LPR * 8;
  constexpr int ROWS = COT_T + 6;
  __shared__ float dots[7][ROWS + 2];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * COT_T;
  const int sub = threadIdx.x % LPR;          // which 8 channels of the frame
  const int rsub = threadIdx.x / LPR;         // frame slot inside one sweep of the block
  constexpr int RPS = 256 / LPR;              // frames per sweep
  float wr[7][8];
#pragma unroll
  for (int tap = 0; tap < 7; ++tap)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[tap][j] = __ldg(w + tap * C + sub * 8 + j);
  const bf16* xb = x + (int64_t)b * T_ * C + sub * 8;
  for (int rbase = 0; rbase < ROWS; rbase += RPS) {  // warp-uniform trip count: the shuffles below need all lanes
    const int r = rbase + rsub;
    const int t = t0 - 3 + r;
    float d[7];
#pragma unroll
    for (int tap = 0; tap < 7; ++tap) d[tap] = 0.f;
    if (r < ROWS && t >= 0 && t < T_) {
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(xb + (int64_t)t * C));
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[2 * j] = __uint_as_float(u[j] << 16);
        f[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
      }
#pragma unroll
      for (int tap = 0; tap < 7; ++tap)
#pragma unroll
        for (int j = 0; j < 8; ++j) d[tap] = fmaf(f[j], wr[tap][j], d[tap]);
    }
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1)
#pragma unroll
      for (int tap = 0; tap < 7; ++tap) d[tap] += __shfl_xor_sync(0xffffffffu, d[tap], o);
#pragma unroll
    for (int tap = 0; tap < 7; ++tap)
      if ((tap % LPR) == sub && r < ROWS) dots[tap][r] = d[tap];
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t < T_) {
    float acc = bias;
#pragma unroll
    for (int tap = 0; tap < 7; ++tap) acc += dots[tap][threadIdx.x + tap];  // frame t + tap - 3 = slot (t - t0) + tap
    wav[(int64_t)b * T_ + t] = tanhf(acc);
  }
}

// y[b,t,c] = sum_tap w[tap][c] * wav[b,t+tap-3] + bias[c]
__global__ void conv_in_wav_kernel(const float* wav, const float* w, const float* bias, int64_t T_, int C, float* y) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= T_ * C) return;
  const int c = (int)(idx % C);
  const int64_t t = idx / C;
  const float* wb = wav + (int64_t)b * T_;
  float acc = 0.f;
#pragma unroll
  for (int tap = 0; tap < 7; ++tap) {
    const int64_t tt = t + tap - 3;
    if (tt >= 0 && tt < T_) acc = fmaf(wb[tt], w[tap * C + c], acc);
  }
  y[((int64_t)b * T_ + t) * C + c] = acc + bias[c];
}

__global__ void transpose_out_kernel(const float* x, int T_, int C, float* y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xb = x + (int64_t)b * T_ * C;
  float* yb = y + (int64_t)b * T_ * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T_ && c < C) tile[i][threadIdx.x] = xb[(int64_t)t * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (t < T_ && c < C) yb[(int64_t)c * T_ + t] = tile[threadIdx.x][i];
  }
}

}  // namespace

template <int ACT_TT, int MINB>
static void launch_act1d_v(const Act1d& p, cudaStream_t stream) {
  int bx = p.C / 2;
  if (bx > 128) bx = 128;
  while ((p.C / 2) % bx) --bx;
  int by = 256 / bx;
  if (by < 1) by = 1;
  dim3 block(bx, by);
  dim3 grid((p.C / 2) / bx, (p.T + by * ACT_TT - 1) / (by * ACT_TT), p.B);
  FLM_REQUIRE(grid.y <= 65535, "act1d: sequence too long");
  if (p.io_bf16 && p.fast_sin) {  // throughput mode: one instantiation per channel count of the FaCodec stacks
    switch (p.C) {
      case 1024: launch_pdl(act1d_kernel<bf16, true, ACT_TT, MINB, 1024>, grid, block, (size_t)0, stream, p); break;
      case 512: launch_pdl(act1d_kernel<bf16, true, ACT_TT, MINB, 512>, grid, block, (size_t)0, stream, p); break;
      case 256: launch_pdl(act1d_kernel<bf16, true, ACT_TT, MINB, 256>, grid, block, (size_t)0, stream, p); break;
      case 128: launch_pdl(act1d_kernel<bf16, true, ACT_TT, MINB, 128>, grid, block, (size_t)0, stream, p); break;
      case 64: launch_pdl(act1d_kernel<bf16, true, ACT_TT, MINB, 64>, grid, block, (size_t)0, stream, p); break;
      default: launch_pdl(act1d_kernel<bf16, true, ACT_TT, MINB, 0>, grid, block, (size_t)0, stream, p); break;
    }
  } else if (p.io_bf16) {
    launch_pdl(act1d_kernel<bf16, false, ACT_TT, MINB, 0>, grid, block, (size_t)0, stream, p);
  } else {
    if (p.fast_sin) launch_pdl(act1d_kernel<float, true, ACT_TT, MINB, 0>, grid, block, (size_t)0, stream, p);
    else launch_pdl(act1d_kernel<float, false, ACT_TT, MINB, 0>, grid, block, (size_t)0, stream, p);
  }
  FLM_LAUNCH_CHECK();
}

void launch_act1d(const Act1d& p, cudaStream_t stream) {
  FLM_REQUIRE(p.C % 2 == 0, "act1d: C must be even");
  if (p.B == 0 || p.T == 0) return;
  // 16 outputs per thread at <= 128 registers (two blocks per SM): measured against 8 / 12 outputs per thread and other
  // register caps on B200 (profiles/r2p/act1d_variants.txt: 16.6 ms per decode of 31.8k frames against 17.1 - 23.6 ms);
  // the kernel is bound by the fp32 pipes (24 FFMA2 + 4 MUFU per element pair), not by occupancy
  launch_act1d_v<16, 2>(p, stream);
}

void launch_conv_out_tanh(const void* x, int x_bf16, const float* w, float bias, int B, int T, int C, float* wav,
                          cudaStream_t stream) {
  FLM_REQUIRE(C % 4 == 0 && C <= 128, "conv_out_tanh: C must be a multiple of 4, <= 128");
  if (B == 0 || T == 0) return;
  if (x_bf16 && (C == 32 || C == 64 || C == 128) && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 g2((T + COT_T - 1) / COT_T, B);
    const bf16* xp = static_cast<const bf16*>(x);
    if (C == 32) conv_out_tanh_bf16_kernel<4><<<g2, 256, 0, stream>>>(xp, w, bias, T, wav);
    else if (C == 64) conv_out_tanh_bf16_kernel<8><<<g2, 256, 0, stream>>>(xp, w, bias, T, wav);
    else conv_out_tanh_bf16_kernel<16><<<g2, 256, 0, stream>>>(xp, w, bias, T, wav);
    FLM_LAUNCH_CHECK();
    return;
  }
  dim3 grid((T + 255) / 256, B);
  if (x_bf16)
    conv_out_tanh_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(x), w, bias, T, C, wav);
  else
    conv_out_tanh_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), w, bias, T, C, wav);
  FLM_LAUNCH_CHECK();
}

void launch_conv_in_wav(const float* wav, const float* w, const float* bias, int B, int64_t T, int C, float* y,
                        cudaStream_t stream) {
  if (B == 0 || T == 0) return;
  dim3 grid((unsigned)((T * C + 255) / 256), B);
  conv_in_wav_kernel<<<grid, 256, 0, stream>>>(wav, w, bias, T, C, y);
  FLM_LAUNCH_CHECK();
}

void launch_transpose_out(const float* x, int B, int T, int C, float* y, cudaStream_t stream) {
  if (B == 0 || T == 0) return;
  dim3 grid((T + 31) / 32, (C + 31) / 32, B);
  transpose_out_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, T, C, y);
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
