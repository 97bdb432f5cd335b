// LayerNorm + adaLN modulate fused into the depthwise conv of the ConvNeXt block (bf16 throughput mode):
//
//     u = LayerNorm(h) * (1 + scale_b) + shift_b          (prob_generator.py:136,162 / 229,257: ln_conv + modulate)
//     d = depthwise_conv31(u) + bias                       (prob_generator.py:81-88,108: conv_1, zero padded)
//     + per-(sample, 32-frame chunk, channel) partial statistics of d for GroupNorm(C, C) (prob_generator.py:89,109)
//
// Before this kernel u was produced by a separate LayerNorm pass (read h, write u: 4 B per element through HBM) and read
// again by the conv; here h is read once, u is written once (conv_3's epilogue needs it as the inner residual) and the
// un-normalised d is written once; the statistics merge and the in-place streaming GroupNorm apply stay separate
// launches (kernels_norm.cu).
//
//   * LayerNorm needs full-row statistics (1024 channels) while a block owns 256 channels: the row sums come from the
//     epilogue of the GEMM that produced h (TapGemm::rowstat, (sum, sumsq) partials per row), so normalising is two
//     FFMA2 per loaded element pair:  u = (x * rstd_r - mean_r * rstd_r) * A_c + B_c  with the per-sample affine
//     A = w (1 + scale), B = b (1 + scale) + shift held in registers.
//   * Statistics are taken about the pivot bias_c + B_c * sum_k w[k,c] (the response to the constant part of u), which
//     removes the large per-channel offset from the sums: the accumulators simply start at -B_c * sum_k w[k,c].
//   * A block owns one 256-channel block and a CONTIGUOUS range of (sample, chunk) tiles: it stays inside one or two
//     samples, so the per-sample affine is reloaded only at a sample switch and the halo frames of consecutive chunks
//     hit L2.
//
// What was tried and measured on B200 (profiles/r2h): a thread-block-cluster form that also exchanged the GroupNorm
// statistics through distributed shared memory and normalised its own rows in place out of L2 (one kernel for the whole
// front half) cost 0.517 ms per velocity evaluation against 0.414 + 0.106 ms for this kernel + the streaming apply (the
// conv is bound by the FMA pipe and the issue slots, so every phase that keeps its warps away from FFMA2 shows); a
// dedicated producer warp (160 threads) capped the registers at 168 and spilled; global loads of the row statistics
// were sunk by the compiler next to their use (barrier stalls 0.64 per issue) until they were moved to cp.async; three
// resident blocks per SM (168 registers, 2-deep rings) were slower than two (0.409 vs 0.377 ms per velocity, profiles/r2z).
//
// 128 threads (thread = 2 channels x 32 frames); thread 0 issues one TMA tile copy per chunk two tiles ahead (62 frames
// x 256 channels, out-of-range frames zero-filled), threads 0..61 prepare the per-row LayerNorm constants of the next
// tile around the convolution of the current one.
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
constexpr int NTHREADS = 128;
constexpr int SMEM_BYTES = STAGES * TILE_BYTES + 2 * RC_BYTES + 64 + 2 * 64 * 16 * 4;

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
// One chunk: LayerNorm-modulate on load, 31-tap window, outputs.
// MASKED: the tile touches frames outside [0, L) (first / last chunks of a sample): their u is forced to zero (the conv
// is zero padded) and stores are predicated.  Interior tiles run the variant without any of that: measured 0.414 ms
// per velocity against 0.486 ms for a single always-masked variant (B=26, L=1225; profiles/r2h).
// CC: compile-time channel count (row stride of the outputs: the 64 stores of a tile then use immediate offsets from two
// base registers, no pointer arithmetic on the issue slots), or 0 for a run-time C.
template <bool MASKED, bool WRITE_U, int CC, typename Mid>
__device__ __forceinline__ void conv_tile(const bf16* __restrict__ xs, const float4* __restrict__ rc,
                                          const f32x2 (&w2)[KW], f32x2 A2, f32x2 B2, f32x2 acc0, f32x2 ob2, int nvalid,
                                          int Crt, bf16* __restrict__ ub, bf16* __restrict__ gb, f32x2& S1, f32x2& S2,
                                          Mid&& mid) {
  const int64_t C = CC ? CC : Crt;
  float b0, b1;
  unpack2(B2, b0, b1);
  f32x2 acc[TT];
#pragma unroll
  for (int j = 0; j < TT; ++j) acc[j] = acc0;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const f32x2 x2 = bf16x2_to_f32x2(*reinterpret_cast<const uint32_t*>(xs + r * CB));
    const float4 c4 = rc[r];  // (rstd, rstd, -mean*rstd, -mean*rstd): broadcast read, already packed
    const f32x2 rr = *reinterpret_cast<const f32x2*>(&c4.x), mm = *reinterpret_cast<const f32x2*>(&c4.z);
    f32x2 Bz = B2;
    if (MASKED) {  // frames outside the sample carry rstd = -mean*rstd = 0: select a zero B term as well -> u = 0
      const bool inside = c4.x != 0.f;
      Bz = pack2(inside ? b0 : 0.f, inside ? b1 : 0.f);
    }
    const f32x2 u2 = fma2(fma2(x2, rr, mm), A2, Bz);
    if (WRITE_U && r >= PAD && r < PAD + TT) {  // centre rows: this tile owns them -> inner-residual operand of conv_3
      if (!MASKED || r - PAD < nvalid) *reinterpret_cast<uint32_t*>(ub + (r - PAD) * C) = f32x2_to_bf16x2(u2);
    }
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      const int tap = r - j;  // compile-time after unrolling
      if (tap >= 0 && tap < KW) acc[j] = fma2(w2[tap], u2, acc[j]);
    }
  }
  mid();  // the LayerNorm constants of the next tile: their inputs landed long ago, the output stores below hide them
#pragma unroll
  for (int j = 0; j < TT; ++j) {
    if (!MASKED || j < nvalid) {
      S1 = add2(S1, acc[j]);
      S2 = fma2(acc[j], acc[j], S2);
      *reinterpret_cast<uint32_t*>(gb + j * C) = f32x2_to_bf16x2(add2(acc[j], ob2));
    }
  }
}

// ---- shared machinery of the two kernels: 128 threads, thread 0 issues the TMA tile copies STAGES-1 tiles ahead,
// threads 0..61 turn the (sum, sumsq) partials of the NEXT tile's rows into its LayerNorm constants while the current
// tile is convolved (the partial loads are issued before the conv and consumed after it).
// The (sum, sumsq) partials of the next tile's rows travel global -> shared with cp.async (no registers, nothing the
// compiler can sink next to the use): issued before the convolution of the current tile, reduced after it.
constexpr int RAW_FLOATS = 16;  // up to 8 partials per row
__device__ __forceinline__ void rc_prefetch(const DwFused& p, int b, int t, float* raw_row) {
  if (t >= 0 && t < p.L && p.parts <= 8) {
    const float* src = p.rowstat + ((int64_t)b * p.L + t) * p.parts * 2;
    const uint32_t dst = s32(raw_row);
    for (int k = 0; 2 * k < p.parts; ++k)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * k), "l"(src + 4 * k) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void rc_finish(const DwFused& p, int b, int t, const float* raw_row, float4* rc) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float rstd = 0.f, nm = 0.f;
  if (t >= 0 && t < p.L) {
    float s = 0.f, qq = 0.f;
    if (p.parts <= 8) {
      for (int k = 0; 2 * k < p.parts; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(raw_row + 4 * k);
        s += v.x + v.z;
        qq += v.y + v.w;
      }
    } else {  // small problems (narrow N tiles): more partials than the staging row holds, read them directly
      const float4* ps = reinterpret_cast<const float4*>(p.rowstat) + (((int64_t)b * p.L + t) * p.parts >> 1);
      for (int k = 0; 2 * k < p.parts; ++k) {
        const float4 v = __ldg(ps + k);
        s += v.x + v.z;
        qq += v.y + v.w;
      }
    }
    const float inv_c = 1.0f / (float)p.C;
    const float mean = s * inv_c;
    const float var = fmaxf(qq * inv_c - mean * mean, 0.f);
    rstd = rsqrtf(var + p.ln_eps);
    nm = -mean * rstd;
  }
  *rc = make_float4(rstd, rstd, nm, nm);  // frames outside the sample: zeros (conv_tile<MASKED> keys on rstd == 0)
}
__device__ __forceinline__ void tma_tile(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int t0, int b) {
  mbar_expect_tx(bar, TILE_BYTES);
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(s32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(c0), "r"(t0), "r"(b)
      : "memory");
}
// per-sample LayerNorm affine of this thread's two channels and the statistics pivot derived from it
struct Affine {
  f32x2 A2, B2, acc0, ob2;
  float piv0, piv1;
};
__device__ __forceinline__ Affine make_affine(const DwFused& p, int b, int c) {
  const float2 ws = *reinterpret_cast<const float2*>(p.wsum + c);
  const float2 bi = *reinterpret_cast<const float2*>(p.bias + c);
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
  Affine a;
  a.A2 = pack2(a0, a1); a.B2 = pack2(b0, b1);
  const float pv0 = b0 * ws.x, pv1 = b1 * ws.y;  // response of the conv to the constant part of u
  a.acc0 = pack2(-pv0, -pv1);                     // accumulators (and statistics) live about that pivot
  a.piv0 = bi.x + pv0; a.piv1 = bi.y + pv1;
  a.ob2 = pack2(a.piv0, a.piv1);                  // added back on the way out
  return a;
}

// persistent blocks over contiguous (sample, chunk) tile ranges; per-chunk (mean, M2) partial statistics out (layout
// (B, nchunk, C, 2), the layout launch_dw_merge consumes)
template <int CC, bool WU>
__global__ void __launch_bounds__(NTHREADS, 2) dwconv_ln_kernel(const __grid_constant__ CUtensorMap tmX, DwFused p,
                                                                float* __restrict__ part, int nchunk, int ncblk,
                                                                int ntiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  float4* rcs = reinterpret_cast<float4*>(smem + STAGES * TILE_BYTES);   // [2][64] LayerNorm constants per row
  uint64_t* full = reinterpret_cast<uint64_t*>(rcs + 2 * 64);            // [STAGES]
  float* praw = reinterpret_cast<float*>(full + 8);                      // [2][64][RAW_FLOATS] staged row partials

  const int tid = threadIdx.x;
  const int cblk = blockIdx.x % ncblk;
  const int slice = blockIdx.x / ncblk, nslices = gridDim.x / ncblk;
  const int t_lo = (int)((int64_t)ntiles * slice / nslices), t_hi = (int)((int64_t)ntiles * (slice + 1) / nslices);
  if (t_lo >= t_hi) return;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();  // h and its row statistics come from the previous kernel
  if (tid == 0) {
    for (int d = 0; d < STAGES - 1; ++d)
      if (t_lo + d < t_hi)
        tma_tile(&tmX, &full[d], ring + d * TILE_BYTES, cblk * CB, ((t_lo + d) % nchunk) * TT - PAD, (t_lo + d) / nchunk);
  }
  const int c = cblk * CB + tid * 2;
  f32x2 w2[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) w2[k] = *reinterpret_cast<const f32x2*>(p.w + (int64_t)k * p.C + c);
  if (tid < ROWS) {  // LayerNorm constants of the first tile (latency exposed once per block)
    const int t = (t_lo % nchunk) * TT - PAD + tid;
    rc_prefetch(p, t_lo / nchunk, t, praw + tid * RAW_FLOATS);
    rc_finish(p, t_lo / nchunk, t, praw + tid * RAW_FLOATS, &rcs[tid]);
  }
  __syncthreads();
  // (sample, chunk) of the current tile, of the next one and of the tile the TMA ring prefetches: kept incrementally
  // (integer divisions by the run-time chunk count cost ~20 instructions each, six of them per tile)
  int b = t_lo / nchunk, chunk = t_lo % nchunk;
  int fb = (t_lo + STAGES - 1) / nchunk, fchunk = (t_lo + STAGES - 1) % nchunk;
  int cur_b = b;
  Affine af = make_affine(p, cur_b, c);
  int slot = 0;
  uint32_t phase = 0;
  for (int tile = t_lo; tile < t_hi; ++tile) {
    const int it = tile - t_lo;
    if (tid == 0) {  // refill the slot drained in the previous iteration (all threads passed its __syncthreads)
      if (tile + STAGES - 1 < t_hi) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int ns = (slot + STAGES - 1) % STAGES;
        tma_tile(&tmX, &full[ns], ring + ns * TILE_BYTES, cblk * CB, fchunk * TT - PAD, fb);
      }
    }
    if (++fchunk == nchunk) { fchunk = 0; ++fb; }
    const bool have_next = tile + 1 < t_hi;
    const int nchunk_next = chunk + 1 == nchunk ? 0 : chunk + 1;
    const int nb = chunk + 1 == nchunk ? b + 1 : b, nt = nchunk_next * TT - PAD + tid;
    float* raw_next = praw + (((it + 1) & 1) * 64 + tid) * RAW_FLOATS;
    if (have_next && tid < ROWS) rc_prefetch(p, nb, nt, raw_next);
    if (b != cur_b) {  // sample switch (warp-uniform, once or twice per block)
      cur_b = b;
      af = make_affine(p, b, c);
    }
    mbar_wait(&full[slot], phase);
    const int t0 = chunk * TT;
    const bf16* xs = reinterpret_cast<const bf16*>(ring + slot * TILE_BYTES) + tid * 2;
    bf16* ub = WU ? p.u + ((int64_t)b * p.L + t0) * p.C + c : nullptr;
    bf16* gb = p.g + ((int64_t)b * p.L + t0) * p.C + c;
    f32x2 T1 = 0ull, T2 = 0ull;
    const float4* rc = rcs + (it & 1) * 64;
    auto mid = [&]() {
      if (have_next && tid < ROWS) rc_finish(p, nb, nt, raw_next, &rcs[((it + 1) & 1) * 64 + tid]);
    };
    const bool interior = (t0 - PAD >= 0) && (t0 + TT + PAD <= p.L);
    if (WU) {
      if (interior) conv_tile<false, true, CC>(xs, rc, w2, af.A2, af.B2, af.acc0, af.ob2, TT, p.C, ub, gb, T1, T2, mid);
      else conv_tile<true, true, CC>(xs, rc, w2, af.A2, af.B2, af.acc0, af.ob2, min(TT, p.L - t0), p.C, ub, gb, T1, T2, mid);
    } else {
      if (interior) conv_tile<false, false, CC>(xs, rc, w2, af.A2, af.B2, af.acc0, af.ob2, TT, p.C, ub, gb, T1, T2, mid);
      else conv_tile<true, false, CC>(xs, rc, w2, af.A2, af.B2, af.acc0, af.ob2, min(TT, p.L - t0), p.C, ub, gb, T1, T2, mid);
      // by-products for conv_3's epilogue: the LayerNorm constants of the rows this tile owns (one channel block writes
      // them) and, on the first chunk of a sample, the (gate, gate A, gate (bias3 + B)) entries of this thread's channels
      if (cblk == 0 && tid >= PAD && tid < PAD + TT && t0 + tid - PAD < p.L) {
        const float4 q = rc[tid];
        reinterpret_cast<float2*>(p.rowconst_out)[(int64_t)b * p.L + t0 + tid - PAD] = make_float2(q.x, q.z);
      }
      if (chunk == 0) {
        const float2 g = *reinterpret_cast<const float2*>(p.gate + (int64_t)b * p.mod_bstride + c);
        const float2 b3 = *reinterpret_cast<const float2*>(p.bias3 + c);
        float a0, a1, b0, b1;
        unpack2(af.A2, a0, a1);
        unpack2(af.B2, b0, b1);
        float* t = p.lnu_out + (int64_t)b * p.lnu_vecs * p.C + c;
        *reinterpret_cast<float2*>(t) = g;
        *reinterpret_cast<float2*>(t + p.C) = make_float2(g.x * a0, g.y * a1);
        *reinterpret_cast<float2*>(t + 2 * p.C) = make_float2(g.x * (b3.x + b0), g.y * (b3.y + b1));
        if (p.lnu_vecs == 4) {
          const float2 sc = *reinterpret_cast<const float2*>(p.scale2 + (int64_t)b * p.mod_bstride + c);
          float2 w2v = make_float2(1.f, 1.f);
          if (p.ln2_w) w2v = *reinterpret_cast<const float2*>(p.ln2_w + c);
          *reinterpret_cast<float2*>(t + 3 * p.C) = make_float2(w2v.x * (1.f + sc.x), w2v.y * (1.f + sc.y));
        }
      }
    }
    {  // per-chunk (mean, M2) about the pivot: mean = pivot + S/n, M2 = Q - S^2/n
      float s0, s1, q0, q1;
      unpack2(T1, s0, s1);
      unpack2(T2, q0, q1);
      const float inv = 1.0f / (float)min(TT, p.L - t0);
      const float d0 = s0 * inv, d1 = s1 * inv;
      *reinterpret_cast<float4*>(part + (((int64_t)b * nchunk + chunk) * p.C + c) * 2) =
          make_float4(af.piv0 + d0, fmaxf(q0 - s0 * d0, 0.f), af.piv1 + d1, fmaxf(q1 - s1 * d1, 0.f));
    }
    __syncthreads();  // the tile slot and the other constants buffer may be overwritten from here on
    if (++slot == STAGES) { slot = 0; phase ^= 1; }
    b = nb; chunk = nchunk_next;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

void dwconv_fused_init() {
  FLM_CUDA(cudaFuncSetAttribute(dwconv_ln_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  FLM_CUDA(cudaFuncSetAttribute(dwconv_ln_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  FLM_CUDA(cudaFuncSetAttribute(dwconv_ln_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  FLM_CUDA(cudaFuncSetAttribute(dwconv_ln_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
}

bool dwconv_fused_supported(const DwFused& p) {
  return p.C % CB == 0 && p.rowstat != nullptr && p.parts >= 2 && p.parts % 2 == 0 && p.tma_encode != nullptr &&
         (reinterpret_cast<uintptr_t>(p.h) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.g) & 3) == 0 &&
         (reinterpret_cast<uintptr_t>(p.u) & 3) == 0 &&
         (p.u != nullptr || (p.rowconst_out && p.lnu_out && p.gate && p.bias3 && p.scale && (p.lnu_vecs == 3 || (p.lnu_vecs == 4 && p.scale2))));
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
  auto go = [&](auto kernel) {
    launch_pdl(kernel, dim3(slices * ncblk), dim3(NTHREADS), (size_t)SMEM_BYTES, stream, tm, p, part, nchunk, ncblk, ntiles);
  };
  if (p.C == 1024) { if (p.u) go(dwconv_ln_kernel<1024, true>); else go(dwconv_ln_kernel<1024, false>); }
  else { if (p.u) go(dwconv_ln_kernel<0, true>); else go(dwconv_ln_kernel<0, false>); }
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
