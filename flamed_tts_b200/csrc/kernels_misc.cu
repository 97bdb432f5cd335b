// Small kernels: time embeddings, noise init, cond fold, duration-generator pieces and the
// integer length regulator.
#include "common.cuh"
#include "kernels.h"

#include <math.h>

namespace flm {

namespace {

// prob_generator.py:48-67  -  cos || sin, t unscaled, f_k = exp(-ln(max_period) * k / half)
__global__ void timestep_embedding_kernel(const float* ts, int n, int dim, float neg_log_period, float* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= n * half) return;
  const int i = idx / half, k = idx % half;
  // reference: exp(-math.log(max_period) * arange(half, fp32) / half): fp32 mul, fp32 div, exp
  const float freq = expf(__fdiv_rn(__fmul_rn(neg_log_period, (float)k), (float)half));
  const float arg = ts[i] * freq;
  out[(int64_t)i * dim + k] = cosf(arg);
  out[(int64_t)i * dim + half + k] = sinf(arg);
}

// pva.py:9-22  -  sin || cos of 1000 * t * exp(-k * ln(1e4)/(half-1))
__global__ void sinusoidal_pos_emb_kernel(const float* ts, int n, int dim, float neg_emb, float* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= n * half) return;
  const int i = idx / half, k = idx % half;
  // reference: exp(arange(half).float() * -emb), emb = math.log(10000)/(half-1) (python double)
  const float freq = expf(__fmul_rn((float)k, neg_emb));
  const float arg = __fmul_rn(__fmul_rn(1000.0f, ts[i]), freq);
  out[(int64_t)i * dim + k] = sinf(arg);
  out[(int64_t)i * dim + half + k] = cosf(arg);
}

template <typename T>
__global__ void silu_sum_kernel(const float* temb, const float* cvec, int nfe, int B, int C, T* out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)nfe * B * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int64_t r = idx / C;
  const int b = (int)(r % B), i = (int)(r / B);
  stf<T>(out + idx, silu(temb[(int64_t)i * C + c] + cvec[(int64_t)b * C + c]));
}

__global__ void noise_init_kernel(const float* noise, const float* cond, float temperature, int64_t n, float* x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = __fadd_rn(__fmul_rn(noise[i], temperature), cond[i]);
}

// ---- counter-based noise (Philox4x32-10 + Box-Muller).  Documented seed -> tensor map (include/flamed_b200.h):
//   group g = i >> 2 of flat element index i;  (r0,r1,r2,r3) = Philox4x32-10(counter = (g_lo, g_hi, tensor_id, 0),
//   key = (seed_lo, seed_hi));  u_k = ((r_k >> 8) + 0.5) * 2^-24 in (0,1);
//   z0 = sqrt(-2 ln u0) cos(2 pi u1), z1 = sqrt(-2 ln u0) sin(2 pi u1), z2 / z3 the same from (u2, u3);  noise[i] = z[i & 3]
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&r)[4]) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t group, uint32_t tensor_id, float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), tensor_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  float u[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) u[k] = ((float)(r[k] >> 8) + 0.5f) * 5.9604644775390625e-08f;  // exact in fp32
#pragma unroll
  for (int k = 0; k < 4; k += 2) {
    const float rad = sqrtf(-2.0f * logf(u[k]));
    float sn, cs;
    sincospif(2.0f * u[k + 1], &sn, &cs);
    z[k] = rad * cs;
    z[k + 1] = rad * sn;
  }
}
// out[i] = z[i] * scale (+ add[i]); thread = one group of 4 consecutive elements
__global__ void philox_normal_kernel(const uint64_t* seed, uint32_t tensor_id, float scale, const float* add, int64_t n,
                                     float* out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g * 4 >= n) return;
  float z[4];
  philox_normal4(*seed, (uint64_t)g, tensor_id, z);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = g * 4 + k;
    if (i < n) {
      const float v = __fmul_rn(z[k], scale);
      out[i] = add ? __fadd_rn(v, add[i]) : v;
    }
  }
}

__global__ void f32_to_bf16_kernel(const float* x, bf16* y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16_rn(x[i]);
}

__global__ void scale_kernel(const float* x, float s, int64_t n, float* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] * s;
}

// prob_generator.py:375-381 + the `x * mask` of Block1D (18-22)
template <typename T>
__global__ void quantizer_fold_kernel(const float* prior, const float* qemb, const uint8_t* mask, int B, int Q, int L,
                                      int D, T* xq, T* xm) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * L * Q * D;
  if (idx >= total) return;
  const int d = (int)(idx % D);
  int64_t r = idx / D;
  const int q = (int)(r % Q);
  r /= Q;
  const int l = (int)(r % L), b = (int)(r / L);
  const float v = prior[(((int64_t)b * Q + q) * L + l) * D + d] + qemb[q * D + d];
  stf<T>(xq + idx, v);
  stf<T>(xm + idx, mask[(int64_t)b * L + l] ? v : 0.f);
}

__global__ void durgen_input_kernel(const float* encp, const float* xt, const float* w0, const float* temb,
                                    int64_t rows, int C, float* out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  const int c = (int)(idx % C);
  const int64_t r = idx / C;
  out[idx] = fmaf(xt[r], w0[c], encp[idx]) + temb[c];
}

// warp per row: v = LN(relu(x)) . wl + bl ; masked_fill ; xt += dt * v.  C % 128 == 0, C <= 512
__global__ void __launch_bounds__(256) durgen_head_kernel(const float* x, int64_t rows, int C, const float* lnw,
                                                          const float* lnb, const float* wl, const float* bl,
                                                          const uint8_t* mask, float dt, float* xt) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = C >> 7;
  float v[4][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) {
      ld4<float>(x + row * C + (i * 32 + lane) * 4, v[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i][j] = fmaxf(v[i][j], 0.f);
        s += v[i][j];
      }
    }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 4;
      float w[4], b[4], l[4];
      ld4<float>(lnw + c, w);
      ld4<float>(lnb + c, b);
      ld4<float>(wl + c, l);
#pragma unroll
      for (int j = 0; j < 4; ++j) dot = fmaf(fmaf((v[i][j] - mean) * rstd, w[j], b[j]), l[j], dot);
    }
  dot = warp_sum(dot);
  if (lane == 0) {
    float vel = dot + bl[0];
    if (mask && mask[row]) vel = 0.f;
    xt[row] = __fadd_rn(xt[row], __fmul_rn(dt, vel));  // mul then add, as torch (no FMA contraction)
  }
}

// pva.py:111-112: clamp(round(exp(x) - 1), min=0); torch.round = round-half-even = rintf
__global__ void duration_round_kernel(const float* x, int64_t n, float* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaxf(rintf(expf(x[i]) - 1.0f), 0.0f);
}

// ---- length regulator (pva.py:125-166), integer only.  One block per sample.
__global__ void lr_plan_kernel(const float* phone, const float* sil, const int64_t* src_lens, int P, int32_t* cumsum,
                               int64_t* tgt_len) {
  extern __shared__ int32_t rep[];  // 2P
  const int b = blockIdx.x;
  const int64_t sl = src_lens[b];
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const bool valid = i < sl;
    // where(valid, d, 0).round().long() then clamp(min=1) / clamp(min=0)
    long long ph = valid ? (long long)rintf(phone[(int64_t)b * P + i]) : 0;
    long long si = valid ? (long long)rintf(sil[(int64_t)b * P + i]) : 0;
    rep[2 * i] = (int32_t)(ph < 1 ? 1 : ph);
    rep[2 * i + 1] = (int32_t)(si < 0 ? 0 : si);
  }
  __syncthreads();
  // inclusive scan of 2P small integers: warp 0, 32 elements per pass with a running carry
  if (threadIdx.x < 32) {
    int32_t carry = 0;
    for (int base = 0; base < 2 * P; base += 32) {
      const int i = base + threadIdx.x;
      int32_t v = i < 2 * P ? rep[i] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)threadIdx.x >= o) v += u;
      }
      v += carry;
      if (i < 2 * P) cumsum[(int64_t)b * 2 * P + i] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
    if (threadIdx.x == 0) tgt_len[b] = carry;
  }
}

// warp per output frame: searchsorted(cs, f, right=True) -> segment; even segment = phoneme seg/2,
// odd segment = silence = row 0 (pva.py:142); frames >= tgt_len are zero padding (tools.py:299-317)
__global__ void __launch_bounds__(256) lr_expand_kernel(const float* x, const int32_t* cumsum, int P, int H, int Tmax,
                                                        float* out, int32_t* out_index) {
  const int b = blockIdx.y;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= Tmax) return;
  const int32_t* cs = cumsum + (int64_t)b * 2 * P;
  const int32_t total = cs[2 * P - 1];
  int src = -1;
  if (f < total) {
    int lo = 0, hi = 2 * P;  // first index with cs[idx] > f
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cs[mid] > f) hi = mid; else lo = mid + 1;
    }
    src = (lo & 1) ? 0 : (lo >> 1);
  }
  float* o = out + ((int64_t)b * Tmax + f) * H;
  if (src >= 0) {
    const float* s = x + ((int64_t)b * P + src) * H;
    for (int c = lane; c < H; c += 32) o[c] = s[c];
  } else {
    for (int c = lane; c < H; c += 32) o[c] = 0.f;
  }
  if (out_index && lane == 0) out_index[(int64_t)b * Tmax + f] = src;
}

// same expand for a batch gathered from several earlier batches: sample b reads its own source rows xs[b] (P[b], H)
// and its own inclusive cumsum cs[b] (2 P[b]); frames >= its total are zero padding
__global__ void __launch_bounds__(256) lr_expand_gather_kernel(const float* const* xs, const int32_t* const* css,
                                                               const int32_t* Ps, int H, int Tmax, float* out,
                                                               int32_t* out_index) {
  const int b = blockIdx.y;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= Tmax) return;
  const int P = Ps[b];
  const int32_t* cs = css[b];
  const int32_t total = cs[2 * P - 1];
  int src = -1;
  if (f < total) {
    int lo = 0, hi = 2 * P;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cs[mid] > f) hi = mid; else lo = mid + 1;
    }
    src = (lo & 1) ? 0 : (lo >> 1);
  }
  float* o = out + ((int64_t)b * Tmax + f) * H;
  if (src >= 0) {
    const float* s = xs[b] + (int64_t)src * H;
    for (int c = lane; c < H; c += 32) o[c] = s[c];
  } else {
    for (int c = lane; c < H; c += 32) o[c] = 0.f;
  }
  if (out_index && lane == 0) out_index[(int64_t)b * Tmax + f] = src;
}

// deterministic pseudo-random fill in [-scale, scale] (benchmark operands; zeros would under-state power)
template <typename T>
__global__ void fill_random_kernel(T* x, int64_t n, uint32_t seed, float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t h = (uint32_t)i * 2654435761u ^ seed;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  stf<T>(x + i, ((float)(h & 0xFFFF) / 32768.0f - 1.0f) * scale);
}

inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void launch_fill_random(void* x, int is_bf16, int64_t n, uint32_t seed, float scale, cudaStream_t stream) {
  if (n == 0) return;
  if (is_bf16) fill_random_kernel<bf16><<<nblk(n), 256, 0, stream>>>(static_cast<bf16*>(x), n, seed, scale);
  else fill_random_kernel<float><<<nblk(n), 256, 0, stream>>>(static_cast<float*>(x), n, seed, scale);
  FLM_LAUNCH_CHECK();
}

void launch_timestep_embedding(const float* ts, int n, int dim, float* out, cudaStream_t stream) {
  if (n == 0) return;
  timestep_embedding_kernel<<<nblk((int64_t)n * dim / 2), 256, 0, stream>>>(ts, n, dim, (float)(-log(10000.0)), out);
  FLM_LAUNCH_CHECK();
}
void launch_sinusoidal_pos_emb(const float* ts, int n, int dim, float* out, cudaStream_t stream) {
  if (n == 0) return;
  sinusoidal_pos_emb_kernel<<<nblk((int64_t)n * dim / 2), 256, 0, stream>>>(
      ts, n, dim, (float)(-(log(10000.0) / (double)(dim / 2 - 1))), out);
  FLM_LAUNCH_CHECK();
}
void launch_silu_sum(const float* temb, const float* cvec, int nfe, int B, int C, void* out, int out_bf16,
                     cudaStream_t stream) {
  const int64_t total = (int64_t)nfe * B * C;
  if (total == 0) return;
  if (out_bf16)
    silu_sum_kernel<bf16><<<nblk(total), 256, 0, stream>>>(temb, cvec, nfe, B, C, static_cast<bf16*>(out));
  else
    silu_sum_kernel<float><<<nblk(total), 256, 0, stream>>>(temb, cvec, nfe, B, C, static_cast<float*>(out));
  FLM_LAUNCH_CHECK();
}
void launch_noise_init(const float* noise, const float* cond, float temperature, int64_t n, float* x, cudaStream_t s) {
  if (n == 0) return;
  noise_init_kernel<<<nblk(n), 256, 0, s>>>(noise, cond, temperature, n, x);
  FLM_LAUNCH_CHECK();
}
void launch_philox_normal(const uint64_t* seed_dev, uint32_t tensor_id, float scale, const float* add, int64_t n,
                          float* out, cudaStream_t s) {
  if (n == 0) return;
  philox_normal_kernel<<<nblk((n + 3) / 4), 256, 0, s>>>(seed_dev, tensor_id, scale, add, n, out);
  FLM_LAUNCH_CHECK();
}
namespace {
__global__ void ln_affine_rows_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ shift,
                                      const float* __restrict__ scale, int64_t stride, int64_t rows, int C,
                                      bf16* __restrict__ s1, bf16* __restrict__ s2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int64_t r = i / C;
  const int c = (int)(i - r * C);
  const float m = 1.f + scale[r * stride + c];
  s1[i] = __float2bfloat16_rn((w ? w[c] : 1.f) * m);
  s2[i] = __float2bfloat16_rn(fmaf(b ? b[c] : 0.f, m, shift[r * stride + c]));
}
}  // namespace
void launch_ln_affine_rows(const float* w, const float* b, const float* shift, const float* scale, int64_t stride, int64_t rows,
                           int C, bf16* s1, bf16* s2, cudaStream_t stream) {
  if (rows == 0) return;
  ln_affine_rows_kernel<<<nblk(rows * C), 256, 0, stream>>>(w, b, shift, scale, stride, rows, C, s1, s2);
  FLM_LAUNCH_CHECK();
}
void launch_f32_to_bf16(const float* x, bf16* y, int64_t n, cudaStream_t stream) {
  if (n == 0) return;
  f32_to_bf16_kernel<<<nblk(n), 256, 0, stream>>>(x, y, n);
  FLM_LAUNCH_CHECK();
}
void launch_scale(const float* x, float s, int64_t n, float* out, cudaStream_t stream) {
  if (n == 0) return;
  scale_kernel<<<nblk(n), 256, 0, stream>>>(x, s, n, out);
  FLM_LAUNCH_CHECK();
}
void launch_quantizer_fold(const float* prior, const float* qemb, const uint8_t* mask, int B, int Q, int L, int D,
                           void* xq, void* xm, int out_bf16, cudaStream_t stream) {
  const int64_t total = (int64_t)B * L * Q * D;
  if (total == 0) return;
  if (out_bf16)
    quantizer_fold_kernel<bf16><<<nblk(total), 256, 0, stream>>>(prior, qemb, mask, B, Q, L, D, static_cast<bf16*>(xq),
                                                                 static_cast<bf16*>(xm));
  else
    quantizer_fold_kernel<float><<<nblk(total), 256, 0, stream>>>(prior, qemb, mask, B, Q, L, D,
                                                                  static_cast<float*>(xq), static_cast<float*>(xm));
  FLM_LAUNCH_CHECK();
}
void launch_durgen_input(const float* encp, const float* xt, const float* w0, const float* temb, int64_t rows, int C,
                         float* out, cudaStream_t stream) {
  if (rows == 0) return;
  durgen_input_kernel<<<nblk(rows * C), 256, 0, stream>>>(encp, xt, w0, temb, rows, C, out);
  FLM_LAUNCH_CHECK();
}
void launch_durgen_head(const float* x, int64_t rows, int C, const float* lnw, const float* lnb, const float* wl,
                        const float* bl, const uint8_t* mask, float dt, float* xt, cudaStream_t stream) {
  FLM_REQUIRE(C % 128 == 0 && C <= 512, "durgen_head: C must be a multiple of 128, <= 512");
  if (rows == 0) return;
  durgen_head_kernel<<<nblk(rows * 32), 256, 0, stream>>>(x, rows, C, lnw, lnb, wl, bl, mask, dt, xt);
  FLM_LAUNCH_CHECK();
}
void launch_duration_round(const float* x, int64_t n, float* out, cudaStream_t stream) {
  if (n == 0) return;
  duration_round_kernel<<<nblk(n), 256, 0, stream>>>(x, n, out);
  FLM_LAUNCH_CHECK();
}
void launch_lr_plan(const float* phone, const float* sil, const int64_t* src_lens, int B, int P, int32_t* cumsum,
                    int64_t* tgt_len, cudaStream_t stream) {
  if (B == 0) return;
  FLM_REQUIRE(P > 0 && (size_t)P * 2 * sizeof(int32_t) <= 48 * 1024, "lr_plan: P out of range (1..6144)");
  lr_plan_kernel<<<B, 128, (size_t)P * 2 * sizeof(int32_t), stream>>>(phone, sil, src_lens, P, cumsum, tgt_len);
  FLM_LAUNCH_CHECK();
}
void launch_lr_expand(const float* x, const int32_t* cumsum, int B, int P, int H, int Tmax, float* out,
                      int32_t* out_index, cudaStream_t stream) {
  if (B == 0 || Tmax == 0) return;
  dim3 grid((Tmax + 7) / 8, B);
  lr_expand_kernel<<<grid, 256, 0, stream>>>(x, cumsum, P, H, Tmax, out, out_index);
  FLM_LAUNCH_CHECK();
}

void launch_lr_expand_gather(const float* const* xs, const int32_t* const* css, const int32_t* Ps, int B, int H,
                             int Tmax, float* out, int32_t* out_index, cudaStream_t stream) {
  if (B == 0 || Tmax == 0) return;
  dim3 grid((Tmax + 7) / 8, B);
  lr_expand_gather_kernel<<<grid, 256, 0, stream>>>(xs, css, Ps, H, Tmax, out, out_index);
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
