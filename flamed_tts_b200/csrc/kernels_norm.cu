// Memory-bound normalisation / depthwise kernels of the denoiser and the cond down-sampler.
// All tensors are channels-last (rows = frames, C contiguous); storage fp32 or bf16, math fp32.
//
// Reference call sites (flamed/models/synthesizer/prob_generator.py):
//   ln_mod      <- nn.LayerNorm(eps=1e-6) + modulate():        136,146,162-163,229,239,257-259
//   dwconv      <- ConvNeXtBlock.conv_1 (depthwise k=31):      81-88,108-109
//   gn_finalize <- GroupNorm(C,C) / GroupNorm(8,.) statistics: 89,109; 15-17,187
//   gn_apply    <- GroupNorm affine (+ Mish/ReLU, mask, skip): 20-22,30-32,198-205
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace flm {

// ------------------------------------------------------------------------------------ ln_mod
namespace {

constexpr int LN_MAX_V = 8;  // C <= 1024

// Each warp owns a contiguous range of rows and streams them through a private ring of LN_DEPTH
// shared-memory slots filled by 1-D bulk TMA copies (cp.async.bulk, one instruction per 4 KB row, completion
// on an mbarrier), so LN_DEPTH rows per warp are always in flight from HBM regardless of register use.
// The affine and adaLN terms are folded per sample into  y = xhat * A + Bv,  A = w*(spo+scale),
// Bv = b*(spo+scale) + shift, kept in registers while the warp stays inside one sample.
constexpr int LN_DEPTH = 3;
constexpr int LN_WARPS = 16;

__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NV, typename TI>
__global__ void __launch_bounds__(LN_WARPS * 32, 1) ln_mod_kernel(LnMod p, int64_t rows_per_warp) {
  constexpr int C = NV * 128;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TI* ring = reinterpret_cast<TI*>(ln_smem) + (size_t)warp * LN_DEPTH * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + (size_t)LN_WARPS * LN_DEPTH * C * sizeof(TI)) + warp * LN_DEPTH;
  const int64_t gw = (int64_t)blockIdx.x * LN_WARPS + warp;
  const int64_t r_begin = gw * rows_per_warp;
  const int64_t r_end = min(p.rows, r_begin + rows_per_warp);
  if (r_begin >= r_end) return;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < LN_DEPTH; ++d)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ln_smem_u32(&bars[d])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  pdl_trigger();
  pdl_wait();
  auto issue = [&](int64_t r, int slot) {  // lane 0 only
    const uint32_t bar = ln_smem_u32(&bars[slot]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(C * sizeof(TI)))
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ln_smem_u32(ring + (size_t)slot * C)),
                 "l"(static_cast<const TI*>(p.x) + r * p.ldx), "r"((uint32_t)(C * sizeof(TI))), "r"(bar)
                 : "memory");
  };
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < LN_DEPTH; ++d)
      if (r_begin + d < r_end) issue(r_begin + d, d);
  }
  float A[NV][4], Bv[NV][4];
  int cur_b = -1;
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t row = r_begin; row < r_end; ++row) {
    {
      const uint32_t bar = ln_smem_u32(&bars[slot]);
      uint32_t done;
      do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(phase)
            : "memory");
      } while (!done);
    }
    float v[NV][4];
    const TI* xs = ring + (size_t)slot * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) ld4<TI>(xs + (i * 32 + lane) * 4, v[i]);
    __syncwarp();
    if (lane == 0 && row + LN_DEPTH < r_end) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async overwrite
      issue(row + LN_DEPTH, slot);
    }
    const int bi = (int)(row / p.rows_per_batch);
    if (bi != cur_b) {  // fold affine + modulation for this sample (warp-uniform branch)
      cur_b = bi;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        float w[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.w) { ld4<float>(p.w + c, w); ld4<float>(p.b + c, b); }
        if (p.scale) {
          float sc[4], sh[4];
          ld4<float>(p.scale + (int64_t)bi * p.mod_bstride + c, sc);
          ld4<float>(p.shift + (int64_t)bi * p.mod_bstride + c, sh);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float m = p.scale_plus_one + sc[j];
            A[i][j] = w[j] * m;
            Bv[i][j] = fmaf(b[j], m, sh[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) { A[i][j] = w[j]; Bv[i][j] = b[j]; }
        }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (p.relu_in) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = fmaxf(v[i][j], 0.f);
      }
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mean = warp_sum(s) * (1.0f / (float)C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i][j] -= mean;
        q = fmaf(v[i][j], v[i][j], q);
      }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / (float)C) + p.eps);
    const bool zero = p.zero_rows != nullptr && p.zero_rows[row] != 0;  // warp-uniform
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = zero ? 0.f : fmaf(v[i][j] * rstd, A[i][j], Bv[i][j]);
      if (p.y_bf16)
        st4<bf16>(static_cast<bf16*>(p.y) + row * p.ldy + c, o);
      else
        st4<float>(static_cast<float*>(p.y) + row * p.ldy + c, o);
    }
    if (++slot == LN_DEPTH) { slot = 0; phase ^= 1; }
  }
}

// bf16 -> bf16 form with packed fp32x2 arithmetic (ncu on the generic kernel: 56 % issue-slot utilisation at
// 5 TB/s, i.e. instruction-bound before HBM-bound): lane = NCH chunks of 8 contiguous channels (one LDS.128 /
// ST.128 each), statistics, normalisation and the folded affine as FADD2 / FFMA2 on channel pairs - half the
// floating-point instructions per row.  Same ring / row ownership as ln_mod_kernel.
template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32, 1) ln_mod_bf16_kernel(LnMod p, int64_t rows_per_warp) {
  constexpr int C = NCH * 256;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* ring = reinterpret_cast<bf16*>(ln_smem) + (size_t)warp * LN_DEPTH * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + (size_t)LN_WARPS * LN_DEPTH * C * sizeof(bf16)) + warp * LN_DEPTH;
  const int64_t gw = (int64_t)blockIdx.x * LN_WARPS + warp;
  const int64_t r_begin = gw * rows_per_warp;
  const int64_t r_end = min(p.rows, r_begin + rows_per_warp);
  if (r_begin >= r_end) return;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < LN_DEPTH; ++d)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ln_smem_u32(&bars[d])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  pdl_trigger();
  pdl_wait();
  auto issue = [&](int64_t r, int slot) {  // lane 0 only
    const uint32_t bar = ln_smem_u32(&bars[slot]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(C * 2)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ln_smem_u32(ring + (size_t)slot * C)),
                 "l"(static_cast<const bf16*>(p.x) + r * p.ldx), "r"((uint32_t)(C * 2)), "r"(bar)
                 : "memory");
  };
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < LN_DEPTH; ++d)
      if (r_begin + d < r_end) issue(r_begin + d, d);
  }
  f32x2 A2[NCH][4], B2[NCH][4];
  int cur_b = -1;
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t row = r_begin; row < r_end; ++row) {
    {
      const uint32_t bar = ln_smem_u32(&bars[slot]);
      uint32_t done;
      do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(phase)
            : "memory");
      } while (!done);
    }
    f32x2 v[NCH][4];
    const bf16* xs = ring + (size_t)slot * C;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const uint4 q = *reinterpret_cast<const uint4*>(xs + (i * 32 + lane) * 8);
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = pack2(__uint_as_float(u[j] << 16), __uint_as_float(u[j] & 0xffff0000u));
    }
    __syncwarp();
    if (lane == 0 && row + LN_DEPTH < r_end) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async overwrite
      issue(row + LN_DEPTH, slot);
    }
    const int bi = (int)(row / p.rows_per_batch);
    if (bi != cur_b) {  // fold affine + modulation for this sample (warp-uniform branch)
      cur_b = bi;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = (i * 32 + lane) * 8;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float w[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, a[4], bb[4];
          if (p.w) { ld4<float>(p.w + c + 4 * h, w); ld4<float>(p.b + c + 4 * h, b); }
          if (p.scale) {
            float sc[4], sh[4];
            ld4<float>(p.scale + (int64_t)bi * p.mod_bstride + c + 4 * h, sc);
            ld4<float>(p.shift + (int64_t)bi * p.mod_bstride + c + 4 * h, sh);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float m = p.scale_plus_one + sc[j];
              a[j] = w[j] * m;
              bb[j] = fmaf(b[j], m, sh[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { a[j] = w[j]; bb[j] = b[j]; }
          }
          A2[i][2 * h] = pack2(a[0], a[1]); A2[i][2 * h + 1] = pack2(a[2], a[3]);
          B2[i][2 * h] = pack2(bb[0], bb[1]); B2[i][2 * h + 1] = pack2(bb[2], bb[3]);
        }
      }
    }
    f32x2 s2 = 0ull;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s2 = add2(s2, v[i][j]);
    float s_lo, s_hi;
    unpack2(s2, s_lo, s_hi);
    const float mean = warp_sum(s_lo + s_hi) * (1.0f / (float)C);
    const f32x2 negm = pack2(-mean, -mean);
    f32x2 q2 = 0ull;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i][j] = add2(v[i][j], negm);
        q2 = fma2(v[i][j], v[i][j], q2);
      }
    float q_lo, q_hi;
    unpack2(q2, q_lo, q_hi);
    const float rstd = rsqrtf(warp_sum(q_lo + q_hi) * (1.0f / (float)C) + p.eps);
    const f32x2 r2 = pack2(rstd, rstd);
    const bool zero = p.zero_rows != nullptr && p.zero_rows[row] != 0;  // warp-uniform
    bf16* yrow = static_cast<bf16*>(p.y) + row * p.ldy;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo, hi;
        unpack2(fma2(mul2(v[i][j], r2), A2[i][j], B2[i][j]), lo, hi);
        __nv_bfloat162 t = __floats2bfloat162_rn(zero ? 0.f : lo, zero ? 0.f : hi);
        o[j] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(yrow + (i * 32 + lane) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (++slot == LN_DEPTH) { slot = 0; phase ^= 1; }
  }
}

template <int NV, typename TI>
constexpr int ln_smem_bytes() { return LN_WARPS * LN_DEPTH * NV * 128 * (int)sizeof(TI) + LN_WARPS * LN_DEPTH * 8; }
template <int NV>
void ln_set_attr() {
  FLM_CUDA(cudaFuncSetAttribute(ln_mod_kernel<NV, float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ln_smem_bytes<NV, float>()));
  FLM_CUDA(cudaFuncSetAttribute(ln_mod_kernel<NV, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ln_smem_bytes<NV, bf16>()));
  if (NV % 2 == 0)
    FLM_CUDA(cudaFuncSetAttribute(ln_mod_bf16_kernel<(NV >= 2 ? NV / 2 : 1)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ln_smem_bytes<NV, bf16>()));
}

template <int NV>
void ln_launch(const LnMod& p, int sms, cudaStream_t stream) {
  int64_t blocks = (p.rows + LN_WARPS - 1) / LN_WARPS;
  if (blocks > sms) blocks = sms;
  const int64_t total_warps = blocks * LN_WARPS;
  const int64_t rpw = (p.rows + total_warps - 1) / total_warps;
  if (p.x_bf16 && p.y_bf16 && !p.relu_in && NV % 2 == 0 && p.ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0)
    launch_pdl(ln_mod_bf16_kernel<(NV >= 2 ? NV / 2 : 1)>, dim3((unsigned)blocks), dim3(LN_WARPS * 32),
               (size_t)ln_smem_bytes<NV, bf16>(), stream, p, rpw);
  else if (p.x_bf16)
    launch_pdl(ln_mod_kernel<NV, bf16>, dim3((unsigned)blocks), dim3(LN_WARPS * 32), (size_t)ln_smem_bytes<NV, bf16>(), stream, p, rpw);
  else
    launch_pdl(ln_mod_kernel<NV, float>, dim3((unsigned)blocks), dim3(LN_WARPS * 32), (size_t)ln_smem_bytes<NV, float>(), stream, p, rpw);
}

}  // namespace

void launch_ln_mod(const LnMod& p, cudaStream_t stream) {
  FLM_REQUIRE(p.C % 128 == 0 && p.C <= 128 * LN_MAX_V, "ln_mod: C must be a multiple of 128 and <= 1024");
  FLM_REQUIRE(p.ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0, "ln_mod: input rows must be 16-byte aligned");
  if (p.rows == 0) return;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  switch (p.C >> 7) {
    case 1: ln_launch<1>(p, sms, stream); break;
    case 2: ln_launch<2>(p, sms, stream); break;
    case 3: ln_launch<3>(p, sms, stream); break;
    case 4: ln_launch<4>(p, sms, stream); break;
    case 8: ln_launch<8>(p, sms, stream); break;
    default: throw Error(-1, "ln_mod: C must be 128, 256, 384, 512 or 1024");
  }
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

// block = 128 threads = 256 channels x DW_TT frames; thread = 2 adjacent channels x DW_TT frames.
// The (DW_TT + KW - 1) x 256 input tile (with its 15-frame halo on both sides, zero outside [0,L)) is staged
// in shared memory with 16-byte cp.async issued up front by all threads, so every global load of the block
// is in flight at once; the 31-tap sliding window then runs out of shared memory.  The two channels of a
// thread are processed as packed fp32x2 (FFMA2: both IEEE fp32 FMAs in one instruction), tap weights and the
// DW_TT accumulators in 64-bit registers.  grid = (C/256, nchunk, B).
template <typename T>
__device__ __forceinline__ f32x2 lds_pair(const T* p);
template <>
__device__ __forceinline__ f32x2 lds_pair<float>(const float* p) {
  return *reinterpret_cast<const f32x2*>(p);
}
template <>
__device__ __forceinline__ f32x2 lds_pair<bf16>(const bf16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);  // bf16 -> fp32 is a 16-bit shift
  return pack2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

__device__ __forceinline__ void dw_merge(const DwConv& p, int b, int nchunk, int c);

// GroupNorm(C,C) statistics: the last block of (sample b, 256-channel block) to finish merges the chunk partials
// in a fixed order (deterministic, no float atomics) into scale = gamma*rstd, offset = beta - mean*scale.
// Called by all 128 threads of a block after they wrote their partials; c = the thread's first channel.
__device__ __forceinline__ void dw_finalize(const DwConv& p, int b, int cblk, int chunk, int nchunk, int ncblk, int c) {
  if (p.scale == nullptr) return;
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    int* ctr = p.counters + b * ncblk + cblk;
    const int old = atomicAdd(ctr, 1);
    is_last = old == nchunk - 1;
    if (is_last) *ctr = 0;  // re-armed for the next launch
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  dw_merge(p, b, nchunk, c);
}

// merge of the chunk partials of channels (c, c+1) of sample b: two passes, no divisions:
// mean = sum(n_k m_k)/N; M2 = sum(q_k + n_k (m_k - mean)^2); fixed order -> deterministic
__device__ __forceinline__ void dw_merge(const DwConv& p, int b, int nchunk, int c) {
  const float* pbase = p.part + ((int64_t)b * nchunk * p.C + c) * 2;
  double sm0 = 0.0, sm1 = 0.0;
  for (int k = 0; k < nchunk; ++k) {
    const float4 pk = __ldcg(reinterpret_cast<const float4*>(pbase + (int64_t)k * p.C * 2));
    const double nb = (double)min(DW_TT, p.L - k * DW_TT);
    sm0 += nb * pk.x; sm1 += nb * pk.z;
  }
  const double inv_n = 1.0 / (double)p.L;
  const double mean0 = sm0 * inv_n, mean1 = sm1 * inv_n;
  double v0 = 0.0, v1 = 0.0;
  for (int k = 0; k < nchunk; ++k) {
    const float4 pk = __ldcg(reinterpret_cast<const float4*>(pbase + (int64_t)k * p.C * 2));
    const double nb = (double)min(DW_TT, p.L - k * DW_TT);
    const double d0 = pk.x - mean0, d1 = pk.z - mean1;
    v0 += pk.y + nb * d0 * d0; v1 += pk.w + nb * d1 * d1;
  }
  const float2 ga = *reinterpret_cast<const float2*>(p.gamma + c);
  const float2 be = *reinterpret_cast<const float2*>(p.beta + c);
  const float sc0 = ga.x * rsqrtf((float)(v0 * inv_n) + p.eps);
  const float sc1 = ga.y * rsqrtf((float)(v1 * inv_n) + p.eps);
  *reinterpret_cast<float2*>(p.scale + (int64_t)b * p.C + c) = make_float2(sc0, sc1);
  *reinterpret_cast<float2*>(p.offset + (int64_t)b * p.C + c) = make_float2(be.x - (float)mean0 * sc0, be.y - (float)mean1 * sc1);
}

template <typename T, int KW>
__global__ void __launch_bounds__(128) dwconv_kernel(DwConv p) {
  constexpr int PAD = KW / 2;
  constexpr int ROWS = DW_TT + KW - 1;
  constexpr int VPR = 256 * (int)sizeof(T) / 16;  // 16-byte vectors per tile row (32 bf16 / 64 fp32)
  constexpr int RPI = 128 / VPR;                  // tile rows covered per cp.async sweep of the block
  extern __shared__ __align__(16) uint8_t dw_smem[];
  T* tile = reinterpret_cast<T*>(dw_smem);
  const int c0 = blockIdx.x * 256;
  const int c = c0 + threadIdx.x * 2;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int t0 = chunk * DW_TT;
  {
    const int col = threadIdx.x % VPR, row0 = threadIdx.x / VPR;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(static_cast<const T*>(p.x) + ((int64_t)b * p.L + (t0 - PAD + row0)) * p.C + c0) + col * 16;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(dw_smem) + (row0 * VPR + col) * 16;
    const int64_t sstep = (int64_t)RPI * p.C * sizeof(T);
#pragma unroll 4
    for (int row = row0; row < ROWS; row += RPI) {
      const int t = t0 - PAD + row;
      if (t >= 0 && t < p.L) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
      else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0));
      src += sstep;
      dst += RPI * VPR * 16;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  f32x2 w2[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) w2[k] = *reinterpret_cast<const f32x2*>(p.w + (int64_t)k * p.C + c);
  f32x2 acc[DW_TT];
#pragma unroll
  for (int j = 0; j < DW_TT; ++j) acc[j] = 0ull;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const T* xs = tile + threadIdx.x * 2;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const f32x2 x2 = lds_pair<T>(xs + r * 256);
#pragma unroll
    for (int j = 0; j < DW_TT; ++j) {
      const int tap = r - j;  // compile-time after unrolling
      if (tap >= 0 && tap < KW) acc[j] = fma2(w2[tap], x2, acc[j]);
    }
  }
  const f32x2 bias2 = *reinterpret_cast<const f32x2*>(p.bias + c);
  const int nvalid = min(DW_TT, p.L - t0);
  T* yb = static_cast<T*>(p.y) + ((int64_t)b * p.L + t0) * p.C + c;
  f32x2 s2 = 0ull;
#pragma unroll
  for (int j = 0; j < DW_TT; ++j) {
    acc[j] = add2(acc[j], bias2);
    if (j < nvalid) {
      float a0, a1;
      unpack2(acc[j], a0, a1);
      st2<T>(yb + (int64_t)j * p.C, a0, a1);
      s2 = add2(s2, acc[j]);
    }
  }
  // per-chunk (mean, M2) of the fp32 results (two-pass, no cancellation); merged later in fp64
  float s0, s1;
  unpack2(s2, s0, s1);
  const float inv = 1.0f / (float)nvalid;
  const float m0 = s0 * inv, m1 = s1 * inv;
  const f32x2 negm = pack2(-m0, -m1);
  f32x2 q2 = 0ull;
#pragma unroll
  for (int j = 0; j < DW_TT; ++j) {
    if (j < nvalid) {
      const f32x2 d = add2(acc[j], negm);
      q2 = fma2(d, d, q2);
    }
  }
  float q0, q1;
  unpack2(q2, q0, q1);
  float* part = p.part + (((int64_t)b * gridDim.y + chunk) * p.C + c) * 2;
  *reinterpret_cast<float4*>(part) = make_float4(m0, q0, m1, q1);
  dw_finalize(p, b, blockIdx.x, chunk, gridDim.y, gridDim.x, c);
}

// bf16-mode form: one warp per (sample, 64 channels); lane = 2 channels, the chunks are split over the 4 warps of
// a block and combined through shared memory in a fixed order (deterministic); fp32 with the chunk means taken
// relative to the first chunk's mean, which keeps the second moment free of cancellation at bf16-level accuracy
__global__ void __launch_bounds__(128) dw_merge_fast_kernel(DwConv p, int nchunk) {
  __shared__ float red[4][32][4];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cgroups = p.C >> 6;
  const int b = blockIdx.x / cgroups, c = (blockIdx.x % cgroups) * 64 + lane * 2;
  const float* pbase = p.part + ((int64_t)b * nchunk * p.C + c) * 2;
  const float4 first = __ldcg(reinterpret_cast<const float4*>(pbase));
  const float r0 = first.x, r1 = first.z;  // reference means
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (int k = w; k < nchunk; k += 4) {
    const float4 pk = __ldcg(reinterpret_cast<const float4*>(pbase + (int64_t)k * p.C * 2));
    const float nb = (float)min(DW_TT, p.L - k * DW_TT);
    const float d0 = pk.x - r0, d1 = pk.z - r1;
    s0 = fmaf(nb, d0, s0); s1 = fmaf(nb, d1, s1);
    q0 += fmaf(nb * d0, d0, pk.y); q1 += fmaf(nb * d1, d1, pk.w);
  }
  red[w][lane][0] = s0; red[w][lane][1] = s1; red[w][lane][2] = q0; red[w][lane][3] = q1;
  __syncthreads();
  if (w != 0) return;
  s0 = (red[0][lane][0] + red[1][lane][0]) + (red[2][lane][0] + red[3][lane][0]);
  s1 = (red[0][lane][1] + red[1][lane][1]) + (red[2][lane][1] + red[3][lane][1]);
  q0 = (red[0][lane][2] + red[1][lane][2]) + (red[2][lane][2] + red[3][lane][2]);
  q1 = (red[0][lane][3] + red[1][lane][3]) + (red[2][lane][3] + red[3][lane][3]);
  const float inv_n = 1.0f / (float)p.L;
  const float dm0 = s0 * inv_n, dm1 = s1 * inv_n;                    // mean - reference
  const float var0 = fmaxf(q0 * inv_n - dm0 * dm0, 0.f), var1 = fmaxf(q1 * inv_n - dm1 * dm1, 0.f);
  const float2 ga = *reinterpret_cast<const float2*>(p.gamma + c);
  const float2 be = *reinterpret_cast<const float2*>(p.beta + c);
  const float sc0 = ga.x * rsqrtf(var0 + p.eps), sc1 = ga.y * rsqrtf(var1 + p.eps);
  *reinterpret_cast<float2*>(p.scale + (int64_t)b * p.C + c) = make_float2(sc0, sc1);
  *reinterpret_cast<float2*>(p.offset + (int64_t)b * p.C + c) = make_float2(be.x - (r0 + dm0) * sc0, be.y - (r1 + dm1) * sc1);
}

// y[b,t,c] = x[b,t,c]*scale[b,c] + offset[b,c]: pure streaming pass, 16-byte vectors, the per-sample
// coefficients of a thread's channels held in registers.  block = (GNS_ROWS rows of sample b), 8 B200-sized
// row groups per block so that >= 8 independent 16 B loads per thread are in flight.
constexpr int GNS_ROWS = 64;
template <typename T>
__global__ void __launch_bounds__(256) gn_stream_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ offset, int L, int C) {
  constexpr int V = 16 / (int)sizeof(T);  // elements per 16-byte vector
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * GNS_ROWS;
  const int nrows = min(GNS_ROWS, L - r0);
  const int vpr = C / V;                    // vectors per row
  const int rstep = 256 / vpr;              // rows covered by one sweep of the block (C=1024: 2 bf16 / 1 fp32)
  const int col = (threadIdx.x % vpr) * V;
  const int rsub = threadIdx.x / vpr;
  if (rsub >= rstep) return;
  float sc[V], of[V];
#pragma unroll
  for (int j = 0; j < V; j += 4) {
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + (int64_t)b * C + col + j));
    const float4 o4 = __ldg(reinterpret_cast<const float4*>(offset + (int64_t)b * C + col + j));
    sc[j] = s4.x; sc[j + 1] = s4.y; sc[j + 2] = s4.z; sc[j + 3] = s4.w;
    of[j] = o4.x; of[j + 1] = o4.y; of[j + 2] = o4.z; of[j + 3] = o4.w;
  }
  const T* xp = x + ((int64_t)b * L + r0 + rsub) * C + col;
  T* yp = y + ((int64_t)b * L + r0 + rsub) * C + col;
  constexpr int U = 8;
  int r = rsub;
  for (; r + (U - 1) * rstep < nrows; r += U * rstep) {
    uint4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = __ldcs(reinterpret_cast<const uint4*>(xp + (int64_t)u * rstep * C));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (sizeof(T) == 2) {
        uint32_t* w = reinterpret_cast<uint32_t*>(&q[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a0 = fmaf(__uint_as_float(w[j] << 16), sc[2 * j], of[2 * j]);
          const float a1 = fmaf(__uint_as_float(w[j] & 0xffff0000u), sc[2 * j + 1], of[2 * j + 1]);
          __nv_bfloat162 o = __floats2bfloat162_rn(a0, a1);
          w[j] = *reinterpret_cast<uint32_t*>(&o);
        }
      } else {
        float* w = reinterpret_cast<float*>(&q[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = fmaf(w[j], sc[j], of[j]);
      }
      *reinterpret_cast<uint4*>(yp + (int64_t)u * rstep * C) = q[u];
    }
    xp += (int64_t)U * rstep * C;
    yp += (int64_t)U * rstep * C;
  }
  for (; r < nrows; r += rstep) {
    uint4 q = *reinterpret_cast<const uint4*>(xp);
    if (sizeof(T) == 2) {
      uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = fmaf(__uint_as_float(w[j] << 16), sc[2 * j], of[2 * j]);
        const float a1 = fmaf(__uint_as_float(w[j] & 0xffff0000u), sc[2 * j + 1], of[2 * j + 1]);
        __nv_bfloat162 o = __floats2bfloat162_rn(a0, a1);
        w[j] = *reinterpret_cast<uint32_t*>(&o);
      }
    } else {
      float* w = reinterpret_cast<float*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = fmaf(w[j], sc[j], of[j]);
    }
    *reinterpret_cast<uint4*>(yp) = q;
    xp += (int64_t)rstep * C;
    yp += (int64_t)rstep * C;
  }
}

}  // namespace

// fp32 parity mode: the depthwise conv of the ConvNeXt block with the GroupNorm statistics finalised by the last block
// of each (sample, channel block).  The bf16 mode uses dwconv_fused.cu (LayerNorm fused in) + launch_dw_merge.
void launch_dwconv(const DwConv& p, cudaStream_t stream) {
  FLM_REQUIRE(p.KW == 31, "dwconv: only kernel_size 31 is compiled (configs/prob.yaml convnext.kernel_size)");
  FLM_REQUIRE(p.C % 256 == 0, "dwconv: C must be a multiple of 256");
  FLM_REQUIRE(!p.io_bf16, "dwconv: the bf16 mode runs the LayerNorm-fused kernel (launch_dwconv_ln)");
  if (p.B == 0 || p.L == 0) return;
  constexpr int ROWS = DW_TT + 31 - 1;
  dim3 grid(p.C / 256, dw_nchunk(p.L), p.B);
  dwconv_kernel<float, 31><<<grid, 128, ROWS * 256 * 4, stream>>>(p);
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

void launch_dw_merge(const DwConv& p, cudaStream_t stream) {
  if (!p.scale || p.B == 0 || p.L == 0) return;
  const int nchunk = dw_nchunk(p.L);
  FLM_REQUIRE(p.C % 64 == 0, "dw_merge: C must be a multiple of 64");
  launch_pdl(dw_merge_fast_kernel, dim3(p.B * (p.C / 64)), dim3(128), (size_t)0, stream, p, nchunk);
  FLM_LAUNCH_CHECK();
}

void launch_gn_stream(const void* x, void* y, int io_bf16, const float* scale, const float* offset, int B, int L, int C,
                      cudaStream_t stream) {
  const int V = io_bf16 ? 8 : 4;
  FLM_REQUIRE(C % V == 0 && C / V <= 256 && 256 % (C / V) == 0, "gn_stream: unsupported channel count");
  if (B == 0 || L == 0) return;
  dim3 grid((L + GNS_ROWS - 1) / GNS_ROWS, B);
  if (io_bf16)
    launch_pdl(gn_stream_kernel<bf16>, grid, dim3(256), (size_t)0, stream, static_cast<const bf16*>(x), static_cast<bf16*>(y), scale,
               offset, L, C);
  else
    launch_pdl(gn_stream_kernel<float>, grid, dim3(256), (size_t)0, stream, static_cast<const float*>(x), static_cast<float*>(y), scale,
               offset, L, C);
  FLM_LAUNCH_CHECK();
}

// opt-in shared-memory sizes; must run once per process outside any stream capture (flm_ctx_create)
void kernels_norm_init() {
  ln_set_attr<1>(); ln_set_attr<2>(); ln_set_attr<3>(); ln_set_attr<4>(); ln_set_attr<8>();
  constexpr int ROWS = DW_TT + 31 - 1;
  FLM_CUDA(cudaFuncSetAttribute(dwconv_kernel<float, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * 256 * 4));
}

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
