// ConvNeXt front half of the denoiser in ONE kernel (bf16 throughput mode):
//
//     u = LayerNorm(h) * (1 + scale_b) + shift_b          (prob_generator.py:136,162 / 229,257: ln_conv + modulate)
//     d = depthwise_conv31(u) + bias                       (prob_generator.py:81-88,108: conv_1, zero padded)
//     g = GroupNorm(C, C)(d)  over the whole time axis     (prob_generator.py:89,109: ln_1)
//
// Before this kernel the three steps were four launches (ln_mod, dwconv + partial statistics, statistics merge,
// streaming GroupNorm apply) moving 12 B per element through HBM; here h is read once (2 B), u is written once (2 B,
// conv_3's epilogue needs it as the inner residual) and g is written once (2 B).
//
//   * LayerNorm needs full-row statistics (1024 channels) while a block owns 256 channels: the row sums come from the
//     epilogue of the GEMM that produced h (TapGemm::rowstat, (sum, sumsq) partials per row), so normalising is two
//     FFMA2 per loaded element pair:  u = (x * rstd_r - mean_r * rstd_r) * A_c + B_c  with the per-sample affine
//     A = w (1 + scale), B = b (1 + scale) + shift held in registers.
//   * GroupNorm needs per-(sample, channel) statistics over ALL frames before the first output can be normalised.
//     A thread-block cluster of S CTAs owns one (sample, 256-channel block): each CTA convolves a contiguous range of
//     32-frame chunks (3-deep TMA ring, 31-tap FFMA2 sliding window as in the previous kernel), keeps the running
//     (sum, sum of squares) of its channels in registers, writes the un-normalised d; the S partial statistics are
//     exchanged through distributed shared memory (fixed order: deterministic), and every CTA then normalises ITS OWN
//     rows in place.  Those rows were written microseconds earlier by the same SM and are still in L2 (a few hundred
//     KB per CTA), so the second pass costs L2 bandwidth, not HBM bandwidth.
//   * Statistics are taken about the pivot bias_c + B_c * sum_k w[k,c] (the response to the constant part of u), which
//     removes the large per-channel offset from the sums: the accumulators simply start at -B_c * sum_k w[k,c].
//
// Roles: warps 0-3 convolve (thread = 2 channels x 32 frames), warp 4 is the producer: one TMA tile copy per chunk
// (62 frames x 256 channels, out-of-range frames zero-filled) and the per-row LayerNorm constants of that tile.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {

constexpr int KW = 31;
constexpr int PAD = KW / 2;
constexpr int TT = DW_TT;             // 32 output frames per chunk
constexpr int ROWS = TT + KW - 1;     // 62 input frames per chunk
constexpr int CB = 256;               // channels per block
constexpr int STAGES = 3;
constexpr int TILE_BYTES = ROWS * CB * 2;
constexpr int RC_BYTES = 64 * 16;     // per stage: 64 rows x (rstd, rstd, -mean*rstd, -mean*rstd)
constexpr int NTHREADS = 160;
constexpr int SMEM_BYTES = STAGES * TILE_BYTES + STAGES * RC_BYTES + STAGES * 64 /*z flags*/ * 4 + 128 * 16 + CB * 8 + 64;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t a = s32(b);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(a), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote));
  return v;
}
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t u) {
  return pack2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(f32x2 v) {
  float a, b;
  unpack2(v, a, b);
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// one chunk: LayerNorm-modulate on load, 31-tap window, outputs.  MASKED: the tile touches frames outside [0, L)
template <bool MASKED>
__device__ __forceinline__ void conv_tile(const bf16* __restrict__ xs, const float4* __restrict__ rc,
                                          const float* __restrict__ zf, const f32x2 (&w2)[KW], f32x2 A2, f32x2 B2,
                                          f32x2 acc0, f32x2 ob2, int t0, int L, int C, bf16* __restrict__ ub,
                                          bf16* __restrict__ gb, f32x2& S1, f32x2& S2) {
  f32x2 acc[TT];
#pragma unroll
  for (int j = 0; j < TT; ++j) acc[j] = acc0;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const f32x2 x2 = bf16x2_to_f32x2(*reinterpret_cast<const uint32_t*>(xs + r * CB));
    const float4 c4 = rc[r];  // (rstd, rstd, -mean*rstd, -mean*rstd): broadcast read, already packed
    const f32x2 rr = *reinterpret_cast<const f32x2*>(&c4.x), mm = *reinterpret_cast<const f32x2*>(&c4.z);
    f32x2 u2 = fma2(fma2(x2, rr, mm), A2, B2);
    if (MASKED) {
      const float z = zf[r];  // 0 for frames outside the sample: the conv is zero padded
      u2 = mul2(u2, pack2(z, z));
    }
    if (r >= PAD && r < PAD + TT) {  // centre rows: this tile owns them -> inner-residual operand of conv_3
      if (!MASKED || t0 + (r - PAD) < L)
        *reinterpret_cast<uint32_t*>(ub + (int64_t)(r - PAD) * C) = f32x2_to_bf16x2(u2);
    }
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      const int tap = r - j;  // compile-time after unrolling
      if (tap >= 0 && tap < KW) acc[j] = fma2(w2[tap], u2, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < TT; ++j) {
    if (!MASKED || t0 + j < L) {
      S1 = add2(S1, acc[j]);
      S2 = fma2(acc[j], acc[j], S2);
      *reinterpret_cast<uint32_t*>(gb + (int64_t)j * C) = f32x2_to_bf16x2(add2(acc[j], ob2));
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 2) dwconv_fused_kernel(const __grid_constant__ CUtensorMap tmX, DwFused p,
                                                                   int nchunk, int ncblk) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  float4* rcs = reinterpret_cast<float4*>(smem + STAGES * TILE_BYTES);
  float* zfs = reinterpret_cast<float*>(smem + STAGES * TILE_BYTES + STAGES * RC_BYTES);
  float4* cstat = reinterpret_cast<float4*>(zfs + STAGES * 64);        // [128] (S1a, S1b, S2a, S2b) of this CTA
  float2* scof = reinterpret_cast<float2*>(cstat + 128);                // [256] (scale, offset) per channel of the block
  uint64_t* full = reinterpret_cast<uint64_t*>(scof + CB);              // [STAGES]
  uint64_t* empty = full + STAGES;                                      // [STAGES]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank(), S = cluster_size();
  const int unit = blockIdx.x / S;
  const int b = unit / ncblk, cblk = unit % ncblk;
  const int q = (nchunk + (int)S - 1) / (int)S;
  const int c_lo = min((int)rank * q, nchunk), c_hi = min(c_lo + q, nchunk);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);   // expect_tx arrive + row-constants arrive (both by the producer warp)
      mbar_init(&empty[i], 4);  // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 4) {
    // ===================== producer: TMA tile + per-row LayerNorm constants =====================
    int slot = 0;
    uint32_t phase = 0;
    const float inv_c = 1.0f / (float)p.C;
    for (int chunk = c_lo; chunk < c_hi; ++chunk) {
      mbar_wait(&empty[slot], phase ^ 1);
      const int t0 = chunk * TT;
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // consumers' generic reads before the async overwrite
        mbar_expect_tx(&full[slot], TILE_BYTES);
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
                "r"(s32(ring + slot * TILE_BYTES)),
            "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(s32(&full[slot])), "r"(cblk * CB), "r"(t0 - PAD), "r"(b)
            : "memory");
      }
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int i = pass * 32 + lane;
        const int t = t0 - PAD + i;
        float rstd = 0.f, nm = 0.f, z = 0.f;
        if (i < ROWS && t >= 0 && t < p.L) {
          const float2* ps = reinterpret_cast<const float2*>(p.rowstat) + ((int64_t)b * p.L + t) * p.parts;
          float s = 0.f, qq = 0.f;
          for (int k = 0; k < p.parts; k += 2) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(ps + k));
            s += v.x + v.z;
            qq += v.y + v.w;
          }
          const float mean = s * inv_c;
          const float var = fmaxf(qq * inv_c - mean * mean, 0.f);
          rstd = rsqrtf(var + p.ln_eps);
          nm = -mean * rstd;
          z = 1.f;
        }
        if (i < 64) {
          rcs[slot * 64 + i] = make_float4(rstd, rstd, nm, nm);
          zfs[slot * 64 + i] = z;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
      if (++slot == STAGES) { slot = 0; phase ^= 1; }
    }
  } else {
    // ===================== consumers: thread = 2 channels =====================
    const int tid = threadIdx.x;  // 0..127
    const int c = cblk * CB + tid * 2;
    f32x2 w2[KW];
#pragma unroll
    for (int k = 0; k < KW; ++k) w2[k] = *reinterpret_cast<const f32x2*>(p.w + (int64_t)k * p.C + c);
    float a0 = 1.f, a1 = 1.f, b0 = 0.f, b1 = 0.f;
    if (p.ln_w) {
      const float2 w = *reinterpret_cast<const float2*>(p.ln_w + c), bb = *reinterpret_cast<const float2*>(p.ln_b + c);
      a0 = w.x; a1 = w.y; b0 = bb.x; b1 = bb.y;
    }
    if (p.scale) {
      const float2 sc = *reinterpret_cast<const float2*>(p.scale + (int64_t)b * p.mod_bstride + c);
      const float2 sh = *reinterpret_cast<const float2*>(p.shift + (int64_t)b * p.mod_bstride + c);
      const float m0 = 1.f + sc.x, m1 = 1.f + sc.y;
      b0 = fmaf(b0, m0, sh.x); b1 = fmaf(b1, m1, sh.y);
      a0 *= m0; a1 *= m1;
    }
    const f32x2 A2 = pack2(a0, a1), B2 = pack2(b0, b1);
    const float2 ws = *reinterpret_cast<const float2*>(p.wsum + c);
    const float2 bi = *reinterpret_cast<const float2*>(p.bias + c);
    const float pv0 = b0 * ws.x, pv1 = b1 * ws.y;           // response of the conv to the constant part of u
    const f32x2 acc0 = pack2(-pv0, -pv1);                    // accumulators (and statistics) live about that pivot
    const f32x2 ob2 = pack2(bi.x + pv0, bi.y + pv1);         // added back on the way out
    f32x2 S1 = 0ull, S2 = 0ull;
    int slot = 0;
    uint32_t phase = 0;
    for (int chunk = c_lo; chunk < c_hi; ++chunk) {
      mbar_wait(&full[slot], phase);
      const int t0 = chunk * TT;
      const bf16* xs = reinterpret_cast<const bf16*>(ring + slot * TILE_BYTES) + tid * 2;
      bf16* ub = p.u + ((int64_t)b * p.L + t0) * p.C + c;
      bf16* gb = p.g + ((int64_t)b * p.L + t0) * p.C + c;
      const bool interior = (t0 - PAD >= 0) && (t0 + TT + PAD <= p.L);
      if (interior)
        conv_tile<false>(xs, rcs + slot * 64, zfs + slot * 64, w2, A2, B2, acc0, ob2, t0, p.L, p.C, ub, gb, S1, S2);
      else
        conv_tile<true>(xs, rcs + slot * 64, zfs + slot * 64, w2, A2, B2, acc0, ob2, t0, p.L, p.C, ub, gb, S1, S2);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      if (++slot == STAGES) { slot = 0; phase ^= 1; }
    }
    float s0, s1, q0, q1;
    unpack2(S1, s0, s1);
    unpack2(S2, q0, q1);
    cstat[tid] = make_float4(s0, s1, q0, q1);
  }

  // ===================== GroupNorm statistics across the cluster (deterministic order) =====================
  cluster_arrive();
  cluster_wait();
  if (warp < 4) {
    const int tid = threadIdx.x;
    const int c = cblk * CB + tid * 2;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    const uint32_t local = s32(&cstat[tid]);
    for (uint32_t r = 0; r < S; ++r) {
      const float4 v = ld_dsmem_f4(local, r);
      s0 += v.x; s1 += v.y; q0 += v.z; q1 += v.w;
    }
    // the pivot of THIS thread's channels (same expression as above)
    float b0 = 0.f, b1 = 0.f;
    if (p.ln_w) { const float2 bb = *reinterpret_cast<const float2*>(p.ln_b + c); b0 = bb.x; b1 = bb.y; }
    if (p.scale) {
      const float2 sc = *reinterpret_cast<const float2*>(p.scale + (int64_t)b * p.mod_bstride + c);
      const float2 sh = *reinterpret_cast<const float2*>(p.shift + (int64_t)b * p.mod_bstride + c);
      b0 = fmaf(b0, 1.f + sc.x, sh.x); b1 = fmaf(b1, 1.f + sc.y, sh.y);
    }
    const float2 ws = *reinterpret_cast<const float2*>(p.wsum + c);
    const float2 bi = *reinterpret_cast<const float2*>(p.bias + c);
    const float inv_n = 1.0f / (float)p.L;
    const float d0 = s0 * inv_n, d1 = s1 * inv_n;  // mean - pivot
    const float v0 = fmaxf(q0 * inv_n - d0 * d0, 0.f), v1 = fmaxf(q1 * inv_n - d1 * d1, 0.f);
    const float2 ga = *reinterpret_cast<const float2*>(p.gamma + c), be = *reinterpret_cast<const float2*>(p.beta + c);
    const float sc0 = ga.x * rsqrtf(v0 + p.gn_eps), sc1 = ga.y * rsqrtf(v1 + p.gn_eps);
    const float m0 = bi.x + b0 * ws.x + d0, m1 = bi.y + b1 * ws.y + d1;
    scof[tid * 2] = make_float2(sc0, be.x - m0 * sc0);
    scof[tid * 2 + 1] = make_float2(sc1, be.y - m1 * sc1);
  }
  cluster_arrive();  // peers may retire once every CTA has read their statistics (waited on at the very end)
  __syncthreads();

  // ===================== normalise this CTA's own rows in place (they are still in L2) =====================
  if (warp < 4) {
    const int cg = lane * 8;  // 8 channels = 16 bytes; a warp covers one 256-channel row segment
    f32x2 sc2[4], of2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 e0 = scof[cg + 2 * j], e1 = scof[cg + 2 * j + 1];
      sc2[j] = pack2(e0.x, e1.x);
      of2[j] = pack2(e0.y, e1.y);
    }
    const int r_lo = c_lo * TT, r_hi = min(c_hi * TT, p.L);
    bf16* base = p.g + ((int64_t)b * p.L) * p.C + cblk * CB + cg;
    constexpr int U = 8;
    int r = r_lo + warp;
    for (; r + 4 * (U - 1) < r_hi; r += 4 * U) {
      uint4 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) v[k] = __ldcg(reinterpret_cast<const uint4*>(base + (int64_t)(r + 4 * k) * p.C));
#pragma unroll
      for (int k = 0; k < U; ++k) {
        uint32_t* w = reinterpret_cast<uint32_t*>(&v[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = f32x2_to_bf16x2(fma2(bf16x2_to_f32x2(w[j]), sc2[j], of2[j]));
        *reinterpret_cast<uint4*>(base + (int64_t)(r + 4 * k) * p.C) = v[k];
      }
    }
    for (; r < r_hi; r += 4) {
      uint4 v = __ldcg(reinterpret_cast<const uint4*>(base + (int64_t)r * p.C));
      uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = f32x2_to_bf16x2(fma2(bf16x2_to_f32x2(w[j]), sc2[j], of2[j]));
      *reinterpret_cast<uint4*>(base + (int64_t)r * p.C) = v;
    }
  }
  cluster_wait();
}

// Persistent form without the in-kernel GroupNorm: LayerNorm-modulate on load + depthwise conv, per-chunk (mean, M2)
// partial statistics out (same layout as the fp32-mode kernel: (B, nchunk, C, 2)), u and the un-normalised d written;
// the statistics merge (dw_merge_fast_kernel) and an IN-PLACE streaming apply follow as separate launches.  A block
// owns one 256-channel block and a CONTIGUOUS range of (sample, chunk) tiles: it stays inside one or two samples, so
// the per-sample LayerNorm affine is reloaded only at a sample switch and the halo frames of consecutive chunks hit L2.
__global__ void __launch_bounds__(NTHREADS, 2) dwconv_ln_kernel(const __grid_constant__ CUtensorMap tmX, DwFused p,
                                                                float* __restrict__ part, int nchunk, int ncblk,
                                                                int ntiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  float4* rcs = reinterpret_cast<float4*>(smem + STAGES * TILE_BYTES);
  float* zfs = reinterpret_cast<float*>(smem + STAGES * TILE_BYTES + STAGES * RC_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(zfs + STAGES * 64);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = blockIdx.x % ncblk;
  const int slice = blockIdx.x / ncblk, nslices = gridDim.x / ncblk;
  const int t_lo = (int)((int64_t)ntiles * slice / nslices), t_hi = (int)((int64_t)ntiles * (slice + 1) / nslices);
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    int slot = 0;
    uint32_t phase = 0;
    const float inv_c = 1.0f / (float)p.C;
    for (int tile = t_lo; tile < t_hi; ++tile) {
      const int b = tile / nchunk, chunk = tile % nchunk;
      mbar_wait(&empty[slot], phase ^ 1);
      const int t0 = chunk * TT;
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[slot], TILE_BYTES);
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
                "r"(s32(ring + slot * TILE_BYTES)),
            "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(s32(&full[slot])), "r"(cblk * CB), "r"(t0 - PAD), "r"(b)
            : "memory");
      }
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int i = pass * 32 + lane;
        const int t = t0 - PAD + i;
        float rstd = 0.f, nm = 0.f, z = 0.f;
        if (i < ROWS && t >= 0 && t < p.L) {
          const float2* ps = reinterpret_cast<const float2*>(p.rowstat) + ((int64_t)b * p.L + t) * p.parts;
          float s = 0.f, qq = 0.f;
          for (int k = 0; k < p.parts; k += 2) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(ps + k));
            s += v.x + v.z;
            qq += v.y + v.w;
          }
          const float mean = s * inv_c;
          const float var = fmaxf(qq * inv_c - mean * mean, 0.f);
          rstd = rsqrtf(var + p.ln_eps);
          nm = -mean * rstd;
          z = 1.f;
        }
        rcs[slot * 64 + i] = make_float4(rstd, rstd, nm, nm);
        zfs[slot * 64 + i] = z;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
      if (++slot == STAGES) { slot = 0; phase ^= 1; }
    }
    return;
  }
  const int tid = threadIdx.x;
  const int c = cblk * CB + tid * 2;
  f32x2 w2[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) w2[k] = *reinterpret_cast<const f32x2*>(p.w + (int64_t)k * p.C + c);
  const float2 ws = *reinterpret_cast<const float2*>(p.wsum + c);
  const float2 bi = *reinterpret_cast<const float2*>(p.bias + c);
  float lw0 = 1.f, lw1 = 1.f, lb0 = 0.f, lb1 = 0.f;
  if (p.ln_w) {
    const float2 w = *reinterpret_cast<const float2*>(p.ln_w + c), bb = *reinterpret_cast<const float2*>(p.ln_b + c);
    lw0 = w.x; lw1 = w.y; lb0 = bb.x; lb1 = bb.y;
  }
  f32x2 A2 = 0ull, B2 = 0ull, acc0 = 0ull, ob2 = 0ull;
  float piv0 = 0.f, piv1 = 0.f;
  int cur_b = -1;
  int slot = 0;
  uint32_t phase = 0;
  for (int tile = t_lo; tile < t_hi; ++tile) {
    const int b = tile / nchunk, chunk = tile % nchunk;
    if (b != cur_b) {  // per-sample LayerNorm affine + statistics pivot (warp-uniform branch, once or twice per block)
      cur_b = b;
      float a0 = lw0, a1 = lw1, b0 = lb0, b1 = lb1;
      if (p.scale) {
        const float2 sc = *reinterpret_cast<const float2*>(p.scale + (int64_t)b * p.mod_bstride + c);
        const float2 sh = *reinterpret_cast<const float2*>(p.shift + (int64_t)b * p.mod_bstride + c);
        const float m0 = 1.f + sc.x, m1 = 1.f + sc.y;
        b0 = fmaf(b0, m0, sh.x); b1 = fmaf(b1, m1, sh.y);
        a0 *= m0; a1 *= m1;
      }
      A2 = pack2(a0, a1); B2 = pack2(b0, b1);
      const float pv0 = b0 * ws.x, pv1 = b1 * ws.y;
      acc0 = pack2(-pv0, -pv1);
      piv0 = bi.x + pv0; piv1 = bi.y + pv1;
      ob2 = pack2(piv0, piv1);
    }
    mbar_wait(&full[slot], phase);
    const int t0 = chunk * TT;
    const bf16* xs = reinterpret_cast<const bf16*>(ring + slot * TILE_BYTES) + tid * 2;
    bf16* ub = p.u + ((int64_t)b * p.L + t0) * p.C + c;
    bf16* gb = p.g + ((int64_t)b * p.L + t0) * p.C + c;
    const bool interior = (t0 - PAD >= 0) && (t0 + TT + PAD <= p.L);
    f32x2 S1 = 0ull, S2 = 0ull;
    if (interior)
      conv_tile<false>(xs, rcs + slot * 64, zfs + slot * 64, w2, A2, B2, acc0, ob2, t0, p.L, p.C, ub, gb, S1, S2);
    else
      conv_tile<true>(xs, rcs + slot * 64, zfs + slot * 64, w2, A2, B2, acc0, ob2, t0, p.L, p.C, ub, gb, S1, S2);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
    if (++slot == STAGES) { slot = 0; phase ^= 1; }
    // per-chunk (mean, M2) about the pivot: mean = pivot + S/n, M2 = Q - S^2/n
    float s0, s1, q0, q1;
    unpack2(S1, s0, s1);
    unpack2(S2, q0, q1);
    const float inv = 1.0f / (float)min(TT, p.L - t0);
    const float d0 = s0 * inv, d1 = s1 * inv;
    *reinterpret_cast<float4*>(part + (((int64_t)b * nchunk + chunk) * p.C + c) * 2) =
        make_float4(piv0 + d0, fmaxf(q0 - s0 * d0, 0.f), piv1 + d1, fmaxf(q1 - s1 * d1, 0.f));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

void dwconv_fused_init() {
  FLM_CUDA(cudaFuncSetAttribute(dwconv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  FLM_CUDA(cudaFuncSetAttribute(dwconv_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
}

bool dwconv_fused_supported(const DwFused& p) {
  return p.C % CB == 0 && p.rowstat != nullptr && p.parts >= 2 && p.parts % 2 == 0 && p.tma_encode != nullptr &&
         (reinterpret_cast<uintptr_t>(p.h) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.g) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.u) & 3) == 0;
}

// cluster size: the S in {1,2,4,8} that minimises (waves of CTAs) x (chunks per CTA + fixed per-CTA cost)
int dwconv_fused_cluster(int B, int L, int C, int num_sms) {
  const int units = B * (C / CB), nchunk = dw_nchunk(L), slots = 2 * num_sms;
  int best = 1;
  long best_cost = -1;
  for (int S = 1; S <= 8; S *= 2) {
    if (S > 1 && (nchunk + S - 1) / S < 2) break;
    const long waves = ((long)units * S + slots - 1) / slots;
    const long cost = waves * ((nchunk + S - 1) / S + 2);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = S; }
  }
  return best;
}

void launch_dwconv_fused(const DwFused& p, int num_sms, cudaStream_t stream) {
  FLM_REQUIRE(dwconv_fused_supported(p), "dwconv_fused: unsupported problem");
  if (p.B == 0 || p.L == 0) return;
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(p.tma_encode);
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)p.C, (cuuint64_t)p.L, (cuuint64_t)p.B};
  cuuint64_t strides[2] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.C * 2 * (cuuint64_t)p.L};
  cuuint32_t box[3] = {CB, (cuuint32_t)ROWS, 1}, estr[3] = {1, 1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(p.h), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-2, "cuTensorMapEncodeTiled(dwconv_fused h) failed: " + std::to_string((int)r));
  const int ncblk = p.C / CB, nchunk = dw_nchunk(p.L);
  const int S = dwconv_fused_cluster(p.B, p.L, p.C, num_sms);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(p.B * ncblk * S));
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FLM_CUDA(cudaLaunchKernelEx(&cfg, dwconv_fused_kernel, tm, p, nchunk, ncblk));
  FLM_LAUNCH_CHECK();
}

// LayerNorm-on-load + depthwise conv, persistent; writes u, the un-normalised d (into p.g) and the per-chunk partial
// statistics `part` (B, nchunk, C, 2).  Follow with launch_dw_merge + an in-place launch_gn_stream on p.g.
void launch_dwconv_ln(const DwFused& p, float* part, int num_sms, cudaStream_t stream) {
  FLM_REQUIRE(dwconv_fused_supported(p) && part != nullptr, "dwconv_ln: unsupported problem");
  if (p.B == 0 || p.L == 0) return;
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(p.tma_encode);
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)p.C, (cuuint64_t)p.L, (cuuint64_t)p.B};
  cuuint64_t strides[2] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.C * 2 * (cuuint64_t)p.L};
  cuuint32_t box[3] = {CB, (cuuint32_t)ROWS, 1}, estr[3] = {1, 1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(p.h), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-2, "cuTensorMapEncodeTiled(dwconv_ln h) failed: " + std::to_string((int)r));
  const int ncblk = p.C / CB, nchunk = dw_nchunk(p.L);
  const int ntiles = p.B * nchunk;
  int slices = (2 * num_sms) / ncblk;
  if (slices > ntiles) slices = ntiles;
  if (slices < 1) slices = 1;
  dwconv_ln_kernel<<<slices * ncblk, NTHREADS, SMEM_BYTES, stream>>>(tm, p, part, nchunk, ncblk, ntiles);
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
