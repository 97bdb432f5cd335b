// Depthwise conv k=31 of the ConvNeXt blocks on the tensor cores, LayerNorm + adaLN modulate applied on load
// (bf16 throughput mode):
//
//     u = LayerNorm(h) * (1 + scale_b) + shift_b          (prob_generator.py:136,162 / 229,257: ln_conv + modulate)
//     d = depthwise_conv31(u) + bias                       (prob_generator.py:81-88,108: conv_1, zero padded)
//     + the GroupNorm(C, C) statistics of d over the padded time axis (prob_generator.py:89,109)
//
// Why tensor cores for a depthwise op.  The FMA form (dwconv_fused.cu) costs 31 FMA + ~25 other instructions per
// output pair and is bound by the FMA pipe / the issue slots at 3x its HBM floor (profiles/r2j: 17 % of a bench step).
// A first tensor-core form (round 1, 16x16 diagonal weight blocks over 16 channels, one MMA per tap) wasted 16x of the
// MACs AND re-read its A operand per tap, which made it shared-memory bound and slower than the FMA kernel.  This
// kernel has no such waste: ONE channel per MMA, the time axis on both M and K.
//
//   * Per channel the modulated input lives in shared memory as a plain time series X[pos] (bf16).  The canonical
//     no-swizzle K-major operand layout puts the rows of an 8-row core matrix 16 B apart; with the descriptor start at
//     X, SBO = 128 B and LBO = 16 B the tensor core therefore reads the HANKEL matrix A[m][k] = X[8 m + k]
//     (m < 128 rows, k < 16) straight out of the series - no im2col, rows simply overlap.
//   * d[8 m + s] = sum_q X[8 m + q] * w[q - s],  q in [s, s + 30]  =>  D (128 x 16) = sum_{r<3} A_r (128 x 16) . T_r (16 x 16)
//     with A_r = the same series advanced by 16 r positions (descriptor start + 32 r bytes) and T_r the Toeplitz
//     block T_r[k][s] = w[16 r + k - s] (zero outside the 31 taps; columns s >= 8 are not used and point at a shared
//     zero block).  3 MMAs of 128x16x16 produce 1024 outputs of one channel: 48 MACs per output against 31 useful.
//   * Samples are laid end to end on a virtual position axis with V = roundup32(L + 30) positions per sample (15 zero
//     positions in front = the conv padding), so that a 992-position tile may straddle samples; rows that fall into the
//     padding are computed and dropped.
//
// STATUS: opt-in experiment (FLAMED_B200_DWCONV=tc), parity-tested, 18-27 % SLOWER than the FMA kernel as measured
// (profiles/r2w/SUMMARY.md): the MMAs are free (tensor pipe 3-4 % busy) but conversion, statistics and addressing still
// cost ~28 thread instructions per output at ~31 % issue-slot utilisation; the floor of the formulation is ~8.
//
// Per tile (32 channels x 992 owned positions), one CTA per SM = four independent pipelines of 4 warps + 4 issue warps.
// Warp group j owns the 8-channel chunk j of the tile end to end:
//   A. the tile's h rows arrive by TMA (32 boxes of 32 positions x 32 channels, SWIZZLE_64B, issued one tile ahead after
//      an L2 prefetch; positions outside a sample are out of bounds for the tensor map = zeros).  ldmatrix.x4.trans reads
//      32 positions x 8 channels per warp instruction and hands every thread (channel, two consecutive positions) pairs -
//      the transposition is free.  LayerNorm + modulate are TWO packed-bf16 FMAs per pair, (x rstd + nm) A + B, with the
//      row constants (launch_dwconv_tc's first kernel reduces the GEMM epilogue's row partials) held as bf16 pairs in
//      shared memory; the pairs go to the channel series X (4-byte stores, conflict free because the series pitch is an
//      odd multiple of 16 B).  u itself is NOT written: conv_3's epilogue recomputes the inner residual (TapGemm::lnu_*);
//   B. the group's issue warp issues 8 channels x 3 MMAs into the group's 128 TMEM columns (16 per channel);
//   C. the group reads TMEM thread-per-row (8 positions x 8 channels per thread), adds the bias, stores d, and reduces
//      (sum, sum of squares about the per-channel pivot) over the warp's rows with a shuffle butterfly: one (mean, M2)
//      partial per (sample, 256-position chunk, channel), merged in a fixed order by the last kernel.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {

constexpr int KW = 31;
constexpr int PADW = KW / 2;
constexpr int TC = 32;                 // channels per tile
constexpr int NPOS = 1024;             // positions converted per tile (128 MMA rows x 8)
constexpr int OWN_ROWS = 124;          // rows whose 8 outputs the tile owns
constexpr int TSTRIDE = OWN_ROWS * 8;  // 992
constexpr int XALLOC = 1064;           // positions the MMAs may touch (row 127, k 47)
constexpr int XPITCH = XALLOC * 2;     // bytes per channel series (multiple of 16)
constexpr int X_BYTES = TC * XPITCH;
constexpr int WBLK = 256;              // one Toeplitz block: 8 (n) x 16 (k) bf16 as two k-halves of 128 B
constexpr int WB_BYTES = TC * 3 * WBLK;
constexpr int NWORK = 512;              // worker threads (16 warps)
constexpr int NTHREADS = NWORK + 128;   // + four issue warps (20 warps: 5 per SM sub-partition)
constexpr int TMEM_COLS = TC * 16;     // 512: 16 columns per channel
constexpr int ST_BYTES = NPOS * TC * 2;  // staged h tile
constexpr int SMEM_BYTES = 1024 + ST_BYTES + X_BYTES + WB_BYTES + 256 /* zero block */ + 4 * 2 * (NPOS / 2) * 8 /* row constants: per warp group, double-buffered */ +
                           TC * 2 * 4 /* bias, wsum */ + 256;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
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
// wait of the issue warps: a long suspend-time hint, so that the polling loop does not take issue slots from the workers
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* b, uint32_t parity) {
  const uint32_t a = s32(b);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(a), "r"(parity), "r"(20000u)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// no-swizzle K-major matrix descriptor: rows of an 8-row core matrix 16 B apart, 8-row groups SBO apart, the two
// 8-element K halves LBO apart
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t hfma2_bf16(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// geometry of the virtual position axis (shared by the conv and the merge kernel)
struct Geo {
  int V;        // positions per sample: roundup8(L + 30)
  int total;    // B * V (the launcher checks that it fits 31 bits)
  int npt;      // position tiles
  int kmax;     // statistics pieces per sample (bound)
};
__host__ __device__ inline Geo make_geo(int B, int L) {
  Geo g;
  g.V = (L + 2 * PADW + 31) & ~31;  // a multiple of the TMA box height: boxes never straddle samples
  g.total = B * g.V;
  g.npt = (g.total + TSTRIDE - 1) / TSTRIDE;
  g.kmax = g.V / 224 + 3;
  return g;
}
// statistics chunks: (tile, quarter) with quarter q covering positions [256 q, min(256 (q + 1), 992)) of the tile
__host__ __device__ inline int chunk_of(int pos) {
  const int pt = pos / TSTRIDE;
  int q = (pos - pt * TSTRIDE) >> 8;
  if (q > 3) q = 3;
  return 4 * pt + q;
}
__host__ __device__ inline void chunk_range(int gchunk, int total, int& lo, int& hi) {
  const int pt = gchunk >> 2, q = gchunk & 3;
  lo = pt * TSTRIDE + 256 * q;
  hi = pt * TSTRIDE + (q == 3 ? TSTRIDE : 256 * (q + 1));
  if (hi > total) hi = total;
}
// valid output frames of sample b inside [lo, hi)
__host__ __device__ inline int valid_in(int b, int V, int L, int lo, int hi) {
  const int a = b * V, e = a + L;
  const int x = lo > a ? lo : a, y = hi < e ? hi : e;
  return y > x ? y - x : 0;
}

// ---- kernel 1: LayerNorm row constants + the per-sample affine
//   rowconst[row] = (rstd, -mean * rstd)  from the (sum, sumsq) partials the producing GEMM's epilogue wrote
//   ab[b][c]      = (A, B) with u = xhat * A + B:  A = w (1 + scale_b), B = b (1 + scale_b) + shift_b
//   lnu[b][0..2][c] = gate, gate * A, gate * (bias3 + B): what conv_3's epilogue needs to recompute u (TapGemm::lnu_table)
__global__ void __launch_bounds__(256) ln_consts_kernel(DwFused p, float2* __restrict__ rowconst, float2* __restrict__ ab,
                                                        const float* __restrict__ gate, const float* __restrict__ bias3,
                                                        float* __restrict__ lnu) {
  pdl_trigger();
  pdl_wait();
  const int64_t rows = (int64_t)p.B * p.L;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) {
    const float4* ps = reinterpret_cast<const float4*>(p.rowstat) + (i * p.parts >> 1);
    float s = 0.f, q = 0.f;
    for (int k = 0; 2 * k < p.parts; ++k) {
      const float4 v = __ldg(ps + k);
      s += v.x + v.z;
      q += v.y + v.w;
    }
    const float inv_c = 1.0f / (float)p.C;
    const float mean = s * inv_c;
    const float var = fmaxf(q * inv_c - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.ln_eps);
    rowconst[i] = make_float2(rstd, -mean * rstd);
  } else if (i < rows + (int64_t)p.B * p.C) {
    const int64_t j = i - rows;
    const int b = (int)(j / p.C), c = (int)(j % p.C);
    float a = 1.f, bb = 0.f;
    if (p.ln_w) { a = p.ln_w[c]; bb = p.ln_b[c]; }
    if (p.scale) {
      const float m = 1.f + p.scale[(int64_t)b * p.mod_bstride + c];
      bb = fmaf(bb, m, p.shift[(int64_t)b * p.mod_bstride + c]);
      a *= m;
    }
    ab[j] = make_float2(a, bb);
    if (lnu) {
      const float g = gate[(int64_t)b * p.mod_bstride + c];
      float* t = lnu + (int64_t)b * p.lnu_vecs * p.C + c;
      t[0] = g;
      t[p.C] = g * a;
      t[2 * p.C] = g * (bias3[c] + bb);
      if (p.lnu_vecs == 4) t[3 * p.C] = (p.ln2_w ? p.ln2_w[c] : 1.f) * (1.f + p.scale2[(int64_t)b * p.mod_bstride + c]);
    }
  }
}

// ---- kernel 2: the conv
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
// the 32 boxes (32 positions x 32 channels, 2 KB, SWIZZLE_64B) of one tile; positions outside a sample's frames (the conv
// padding, the alignment gap, samples beyond the batch) are out of bounds for the tensor map and arrive as zeros
__device__ __forceinline__ void tma_tile(const CUtensorMap* tm, uint64_t* bar, uint8_t* st, int c0, int P0, int V) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(NPOS * TC * 2) : "memory");
  int b = P0 / V, f = P0 - b * V - PADW;
#pragma unroll 1
  for (int k = 0; k < NPOS / 32; ++k) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(s32(st + k * 2048)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(c0), "r"(f), "r"(b)
        : "memory");
    f += 32;
    if (f + PADW >= V) { f -= V; ++b; }
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// Four independent pipelines of 4 warps each + 1 copy warp.  Warp group j owns the 8-channel chunk j of the tile end to
// end: it converts its 16-byte column of the staged rows into its 8 channel series, issues its 24 MMAs into its 128 TMEM
// columns and runs their epilogue - while it waits for its MMAs or the next tile's copies the other groups compute.
// Shared between the groups: the staged tile only.
//   full_bar     copy warp -> groups   the staged tile has landed
//   st_free      groups -> copy warp   the staged tile has been consumed (16 warp arrivals): next tile's copies may start
//   mma_bar[j]   tensor core -> group j   the 24 MMAs of chunk j have completed (tcgen05.commit)
__global__ void __launch_bounds__(NTHREADS, 1)  // 20 warps: 5 per SM sub-partition (16 K registers) -> 96 per thread
dwconv_tc_kernel(const __grid_constant__ CUtensorMap tmH, DwFused p, const float2* __restrict__ rowconst,
                 const float2* __restrict__ ab, float2* __restrict__ part, Geo geo, int ntiles) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET on the shared-space pointer (a round trip through uintptr_t would turn every access
  // below into a generic-address load / store)
  uint8_t* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  uint8_t* st = smem;                                    // [1024 positions][32 ch] bf16, SWIZZLE_64B (TMA)
  uint8_t* xs = st + ST_BYTES;                           // [TC][XPITCH] channel series
  uint8_t* wb = xs + X_BYTES;                            // [TC][3][256] Toeplitz blocks
  uint8_t* zb = wb + WB_BYTES;                           // 256 B of zeros (columns 8..15 of every block)
  uint2* rc_all = reinterpret_cast<uint2*>(zb + 256);    // [4 groups][2 (tile parity)][512] per position pair: bf16x2 rstd, bf16x2 -mean * rstd (0 = padding)
  float* s_bias = reinterpret_cast<float*>(rc_all + 4 * 2 * (NPOS / 2));  // [TC]
  float* s_wsum = s_bias + TC;                           // [TC]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(s_wsum + TC);  // [4]
  uint64_t* xready = mma_bar + 4;                                // [4]
  uint64_t* full_bar = xready + 4;
  uint64_t* st_free = full_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(st_free + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t_lo = (int)((int64_t)ntiles * blockIdx.x / gridDim.x), t_hi = (int)((int64_t)ntiles * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmH)) : "memory");
    for (int j = 0; j < 4; ++j) { mbar_init(&mma_bar[j], 1); mbar_init(&xready[j], 4); }
    mbar_init(full_bar, 1);
    mbar_init(st_free, NWORK / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_ptr)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 16) {
    reinterpret_cast<uint4*>(zb)[tid] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // read by every group's MMAs
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();
  pdl_wait();  // h, the row constants and the affine come from the previous kernels

  const int V = geo.V, L = p.L, C = p.C;

  if (warp >= NWORK / 32) {
    // ======================================= issue warps: one per warp group =======================================
    // warp NWORK / 32 + j issues the MMAs of group j (a single thread needs ~15 instructions per MMA: four issuers in
    // parallel keep that off every group's critical path); the first one also owns the copies of the staged tile.
    const int j = warp - NWORK / 32;
    if (lane == 0 && t_lo < t_hi) {
      constexpr uint32_t idesc = make_idesc(128, 16);
      constexpr uint64_t kDbStep = (uint64_t)(WBLK >> 4) - ((uint64_t)(WBLK >> 4) << 32);
      if (j == 0) tma_tile(&tmH, full_bar, st, (t_lo / geo.npt) * TC, (t_lo % geo.npt) * TSTRIDE, V);
      const uint32_t blk0 = s32(wb) + (uint32_t)(j * 8 * 3 * WBLK);
      const uint64_t da0 = make_desc_noswz(s32(xs) + (uint32_t)(j * 8 * XPITCH), 16, 128);
      const uint64_t db0 = make_desc_noswz(blk0, 128, s32(zb) - blk0);
      const uint32_t td = tmem_base + (uint32_t)(j * 128);
      uint32_t ph = 0;
      for (int tile = t_lo; tile < t_hi; ++tile) {
        const bool have_next = tile + 1 < t_hi;
        if (j == 0 && have_next) {  // pull the next tile into L2 now: its copies (issued once this tile is consumed) then hit
          int P1 = ((tile + 1) % geo.npt) * TSTRIDE;
          int b = P1 / V, f = P1 - b * V - PADW;
          const int c1 = ((tile + 1) / geo.npt) * TC;
#pragma unroll 1
          for (int k = 0; k < NPOS / 32; ++k) {
            asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(&tmH)),
                         "r"(c1), "r"(f), "r"(b)
                         : "memory");
            f += 32;
            if (f + PADW >= V) { f -= V; ++b; }
          }
        }
        // 8 channels x 3 MMAs once the group's series are written.  The descriptors advance linearly: A by one channel
        // series per channel and 32 B per k-chunk, the Toeplitz block by 256 B per MMA with its SBO (distance to the zero
        // block) shrinking in step
        mbar_wait_relaxed(&xready[j], ph);
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
#pragma unroll
          for (int r3 = 0; r3 < 3; ++r3)
            umma_f16(td + (uint32_t)(ch * 16), da0 + (uint64_t)(ch * (XPITCH >> 4) + 2 * r3),
                     db0 + (uint64_t)(ch * 3 + r3) * kDbStep, idesc, r3 > 0 ? 1u : 0u);
        }
        umma_commit(&mma_bar[j]);
        if (j == 0 && have_next) {
          mbar_wait_relaxed(st_free, ph);
          tma_tile(&tmH, full_bar, st, ((tile + 1) / geo.npt) * TC, ((tile + 1) % geo.npt) * TSTRIDE, V);
        }
        ph ^= 1;
      }
    }
  } else {
    // ======================================= warp group `grp`: channels 8 grp .. 8 grp + 7 of the tile =======================
    const int grp = warp >> 2, quarter = warp & 3, gt = tid & 127;  // gt: thread inside the group
    const int cl = grp * 8;
    uint2* rcp_base = rc_all + grp * 2 * (NPOS / 2);
    uint8_t* xg = xs + cl * XPITCH;          // the group's 8 channel series
    uint8_t* wg = wb + cl * 3 * WBLK;        // and their 24 Toeplitz blocks
    uint32_t ph = 0;
    int cur_cb = -1;
    // LayerNorm constants of the 4 position pairs (gt + 128 k) of a tile as bf16x2 (rstd, rstd') and (nm, nm'); padding
    // positions get zeros.  bf16 constants: the conversion below runs in packed bf16 arithmetic (two instructions per
    // position pair instead of six); their rounding is a per-row factor 1 +- 2^-9 on u, of the size of the bf16 rounding u
    // gets anyway (the inner residual of conv_3 is NOT taken from here: its epilogue recomputes it in fp32).
    auto load_rc = [&](int tile, uint2 (&rc)[4]) {
      const int P0 = (tile % geo.npt) * TSTRIDE;
      int b = (P0 + 2 * gt) / V, f = P0 + 2 * gt - b * V - PADW;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 r0 = make_float2(0.f, 0.f), r1 = r0;
        if (b < p.B) {
          const float2* rp = rowconst + (int64_t)b * L + f;
          if ((unsigned)f < (unsigned)L) r0 = __ldg(rp);
          if ((unsigned)(f + 1) < (unsigned)L) r1 = __ldg(rp + 1);
        }
        rc[k] = make_uint2(pack_bf16(r0.x, r1.x), pack_bf16(r0.y, r1.y));
        f += 256;
        while (f + PADW >= V) { f -= V; ++b; }
      }
    };
    uint2 rc[4];
    if (t_lo < t_hi) load_rc(t_lo, rc);
    const int em = quarter * 32 + lane;  // epilogue row of this thread
    int e_pt = -2, e_b = 0, e_j = 0;

    for (int tile = t_lo; tile < t_hi; ++tile) {
      const int cb = tile / geo.npt, pt = tile - cb * geo.npt;
      const int c0 = cb * TC;
      const int cg = c0 + cl;
      // (double-buffered by tile parity: a warp may start the next tile while another one of its group still converts)
      uint2* rcp_s = rcp_base + ((tile - t_lo) & 1) * (NPOS / 2);
      if (cb != cur_cb) {  // new channel block (once or twice per CTA): Toeplitz blocks of the group's 8 channels
        cur_cb = cb;
        group_sync(grp);  // every warp of the group is done with the previous block's bias / tap sums
        float* wst = reinterpret_cast<float*>(xg);  // (31, 8) staging, aliases the group's series (idle: its MMAs are done)
        for (int i = gt; i < KW * 8; i += 128) wst[i] = p.w[(int64_t)(i >> 3) * C + cg + (i & 7)];
        if (gt < 8) { s_bias[cl + gt] = p.bias[cg + gt]; s_wsum[cl + gt] = p.wsum[cg + gt]; }
        group_sync(grp);
        for (int e = gt; e < 8 * 3 * 16; e += 128) {
          const int ch = e / 48, rem = e - ch * 48, r = rem >> 4, khalf = (rem >> 3) & 1, n = rem & 7;
          uint32_t wd[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int tap = 16 * r + khalf * 8 + 2 * j + hh - n;
              v[hh] = (tap >= 0 && tap < KW) ? wst[tap * 8 + ch] : 0.f;
            }
            wd[j] = pack_bf16(v[0], v[1]);
          }
          *reinterpret_cast<uint4*>(wg + (ch * 3 + r) * WBLK + khalf * 128 + n * 16) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
        group_sync(grp);
        // positions [1024, 1064) of every series are read (times zero weights) but never written by the conversion
        if (gt < 8 * 5) *reinterpret_cast<uint4*>(xg + (gt / 5) * XPITCH + NPOS * 2 + (gt % 5) * 16) = make_uint4(0, 0, 0, 0);
      }
      const int P0 = pt * TSTRIDE;
#pragma unroll
      for (int k = 0; k < 4; ++k) rcp_s[gt + 128 * k] = rc[k];
      group_sync(grp);
      mbar_wait(full_bar, ph);

      // ============ A: LayerNorm-modulate the group's column of the staged tile, transposed into its 8 channel series ============
      {
        // One ldmatrix.x4.trans reads 32 positions x 8 channels: lanes 8 mi .. 8 mi + 7 address the rows of the 8-position
        // group mi; result register mi holds channel lane / 4 at positions 2 (lane % 4), + 1 of that group - already a
        // transposed bf16 pair, which two packed bf16 FMAs turn into u:  (x rstd + nm) A + B.
        // The SWIZZLE_64B term of the staged address depends on the low row bits only.
        const int chl = lane >> 2;
        const uint32_t sa0 = s32(st) + quarter * 16384 + ((lane >> 3) * 8 + (lane & 7)) * 64 + ((grp ^ (((lane & 7) >> 1) & 3)) << 4);
        const uint32_t xd0 = s32(xg) + chl * XPITCH + quarter * 512 + (lane & 3) * 4;
        const uint2* rcq = rcp_s + quarter * 128 + (lane & 3);
        const int pos0 = P0 + quarter * 256;  // the warp converts positions pos0 .. pos0 + 255
        int b = pos0 / V;
        const int rem0 = pos0 - b * V;
        if (b < p.B && rem0 >= PADW && rem0 + 256 <= PADW + L) {
          // every position is a frame of sample b
          const float2 abv = __ldg(ab + (int64_t)b * C + cg + chl);
          const uint32_t A2 = pack_bf16(abv.x, abv.x), B2 = pack_bf16(abv.y, abv.y);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            uint32_t r[4];
            ldmatrix_x4_trans(sa0 + it * 2048, r);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
              const uint2 c2 = rcq[it * 16 + mi * 4];
              const uint32_t u2 = hfma2_bf16(hfma2_bf16(r[mi], c2.x, c2.y), A2, B2);
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(xd0 + it * 64 + mi * 16), "r"(u2) : "memory");
            }
          }
        } else {
          // the range touches conv padding / a sample boundary: per 32 positions (never straddle samples), B masked to zero
          // on padding positions (their x, rstd and nm are zero: u = 0, the conv is zero padded)
          int rem = rem0;
#pragma unroll 1
          for (int it = 0; it < 8; ++it) {
            const float2 abv = __ldg(ab + (int64_t)(b < p.B ? b : 0) * C + cg + chl);
            const uint32_t A2 = pack_bf16(abv.x, abv.x), B2 = pack_bf16(abv.y, abv.y);
            uint32_t r[4];
            ldmatrix_x4_trans(sa0 + it * 2048, r);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
              const uint2 c2 = rcq[it * 16 + mi * 4];
              const uint32_t mask = ((c2.x & 0xffffu) ? 0xffffu : 0u) | ((c2.x >> 16) ? 0xffff0000u : 0u);
              const uint32_t u2 = hfma2_bf16(hfma2_bf16(r[mi], c2.x, c2.y), A2, B2 & mask);
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(xd0 + it * 64 + mi * 16), "r"(u2) : "memory");
            }
            rem += 32;
            if (rem >= V) { rem -= V; ++b; }
          }
        }
        // this warp is done with the staged tile (generic-proxy reads -> the next tile's TMA writes) and has written its
        // part of the series (generic-proxy writes -> tensor-core reads)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();  // (orders this thread's TMEM reads of the previous tile before the hand-off)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(st_free);
          mbar_arrive(&xready[grp]);  // 4 warp arrivals: the group's series are complete, its TMEM columns drained
        }
      }
      if (tile + 1 < t_hi) load_rc(tile + 1, rc);  // in flight across the epilogue

      // ============ C: epilogue of the group's 8 channels, thread = 8 consecutive output positions ============
      // row geometry: sample / first frame of this thread's 8 positions, kept incrementally from tile to tile
      if (pt != e_pt + 1) {
        e_b = (P0 + 8 * em) / V;
        e_j = P0 + 8 * em - e_b * V;
      } else {
        e_j += TSTRIDE;
        while (e_j >= V) { e_j -= V; ++e_b; }
      }
      e_pt = pt;
      const bool active = em < OWN_ROWS && P0 + 8 * em < geo.total;
      const int eb = active ? e_b : 0;
      const int j0 = active ? e_j : 0;
      int nval = active ? L - j0 : 0;  // valid frames of this row: j0 .. j0 + nval - 1
      nval = nval < 0 ? 0 : (nval > 8 ? 8 : nval);
      const unsigned act = __ballot_sync(0xffffffffu, active);
      const int b_lo = __shfl_sync(0xffffffffu, eb, 0), b_hi = __shfl_sync(0xffffffffu, eb, act ? 31 - __clz(act) : 0);
      const int gchunk = 4 * pt + quarter;
      const int clo = P0 + 256 * quarter;
      const int chi = min(P0 + (quarter == 3 ? TSTRIDE : 256 * (quarter + 1)), geo.total);
      const bool full_warp = __all_sync(0xffffffffu, nval == 8) && b_lo == b_hi;  // every row complete, one sample
      // d = acc - pivot (the response of the conv to the constant part of u): statistics are taken about the pivot, bias +
      // pivot is added back on the way out
      float piv[8], ob[8];
      {
        const float4* ap = reinterpret_cast<const float4*>(ab + (int64_t)eb * C + cg);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 a = __ldg(ap + j);
          piv[2 * j] = a.y * s_wsum[cl + 2 * j];
          piv[2 * j + 1] = a.w * s_wsum[cl + 2 * j + 1];
          ob[2 * j] = s_bias[cl + 2 * j] + piv[2 * j];
          ob[2 * j + 1] = s_bias[cl + 2 * j + 1] + piv[2 * j + 1];
        }
      }
      mbar_wait(&mma_bar[grp], ph);  // the MMAs of this warp group's 8 channels
      tc_fence_after();
      float st[16];  // [0..7] sums, [8..15] sums of squares
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cl * 16);
      bf16* gp = p.g + ((int64_t)eb * L + j0) * C + cg;
      {
        // packed arithmetic, pairs = two consecutive positions of a channel.  Rows that are not complete (the end of a
        // sample, rows the tile does not own) take part with their statistics masked and their stores predicated; the
        // branch is warp-uniform (tcgen05.ld is a warp-collective instruction).
        const bool partial = __any_sync(0xffffffffu, nval != 8);
        f32x2 s1[8], s2[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) s1[cc] = s2[cc] = 0ull;
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {  // positions 4 hs .. 4 hs + 3 of the row
          float v[8][4];
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) tmem_ld4(taddr + (uint32_t)(cc * 16 + hs * 4), v[cc]);
          tmem_ld_wait();
          if (!partial) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const f32x2 np = pack2(-piv[cc], -piv[cc]), o2 = pack2(ob[cc], ob[cc]);
#pragma unroll
              for (int s = 0; s < 4; s += 2) {
                const f32x2 d = add2(pack2(v[cc][s], v[cc][s + 1]), np);
                s1[cc] = add2(s1[cc], d);
                s2[cc] = fma2(d, d, s2[cc]);
                unpack2(add2(d, o2), v[cc][s], v[cc][s + 1]);
              }
            }
#pragma unroll
            for (int s = 0; s < 4; ++s)
              *reinterpret_cast<uint4*>(gp + (int64_t)(hs * 4 + s) * C) =
                  make_uint4(pack_bf16(v[0][s], v[1][s]), pack_bf16(v[2][s], v[3][s]), pack_bf16(v[4][s], v[5][s]),
                             pack_bf16(v[6][s], v[7][s]));
          } else {
            // 1.0 / 0.0 per position: masks the statistics of the positions beyond the row's valid frames
            const f32x2 k01 = pack2(hs * 4 + 0 < nval ? 1.f : 0.f, hs * 4 + 1 < nval ? 1.f : 0.f);
            const f32x2 k23 = pack2(hs * 4 + 2 < nval ? 1.f : 0.f, hs * 4 + 3 < nval ? 1.f : 0.f);
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const f32x2 np = pack2(-piv[cc], -piv[cc]), o2 = pack2(ob[cc], ob[cc]);
#pragma unroll
              for (int s = 0; s < 4; s += 2) {
                const f32x2 d = add2(pack2(v[cc][s], v[cc][s + 1]), np);
                const f32x2 dm = mul2(d, s == 0 ? k01 : k23);
                s1[cc] = add2(s1[cc], dm);
                s2[cc] = fma2(dm, dm, s2[cc]);
                unpack2(add2(d, o2), v[cc][s], v[cc][s + 1]);
              }
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              if (hs * 4 + s < nval)
                *reinterpret_cast<uint4*>(gp + (int64_t)(hs * 4 + s) * C) =
                    make_uint4(pack_bf16(v[0][s], v[1][s]), pack_bf16(v[2][s], v[3][s]), pack_bf16(v[4][s], v[5][s]),
                               pack_bf16(v[6][s], v[7][s]));
            }
          }
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          float a0, a1, q0, q1;
          unpack2(s1[cc], a0, a1);
          unpack2(s2[cc], q0, q1);
          st[cc] = a0 + a1;
          st[8 + cc] = q0 + q1;
        }
      }
      tc_fence_before();
      // statistics of the warp's chunk, one pass per sample present in it (a chunk straddles samples when V is small)
      if (act) {  // warp-uniform
        for (int bb = b_lo; bb <= b_hi; ++bb) {
          float w[16];
          if (full_warp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = st[i];
          } else {
            const bool sel = active && eb == bb;
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = sel ? st[i] : 0.f;
          }
#pragma unroll
          for (int o = 16, n = 8; o >= 2; o >>= 1, n >>= 1) {
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              const float send = upper ? w[i] : w[i + n];
              const float keep = upper ? w[i + n] : w[i];
              w[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);  // lane l now holds the total of entry l >> 1
          const float q = __shfl_down_sync(0xffffffffu, w[0], 16);  // the matching sum of squares
          if (lane < 16 && !(lane & 1)) {
            const int cc = lane >> 1;
            const int n = valid_in(bb, V, L, clo, chi);
            float mean = 0.f, m2 = 0.f;
            if (n > 0) {
              const float dm = w[0] / (float)n;
              // pivot of sample bb for this channel (the lane's own piv[] belongs to its own row's sample)
              const float pv = __ldg(reinterpret_cast<const float*>(ab + (int64_t)bb * C + cg + cc) + 1) * s_wsum[cl + cc];
              mean = pv + s_bias[cl + cc] + dm;
              m2 = fmaxf(q - w[0] * dm, 0.f);
            }
            const int k = gchunk - chunk_of(bb * V);
            part[((int64_t)bb * geo.kmax + k) * C + cg + cc] = make_float2(mean, m2);
          }
        }
      }
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- kernel 3: merge the (mean, M2) pieces of every (sample, channel) in a fixed order (Chan) -> GroupNorm scale / offset
__global__ void __launch_bounds__(256) dwtc_merge_kernel(DwFused p, const float2* __restrict__ part, Geo geo,
                                                         float* __restrict__ scale, float* __restrict__ offset) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)p.B * p.C) return;
  const int b = (int)(i / p.C), c = (int)(i % p.C);
  const int g0 = chunk_of(b * geo.V), g1 = chunk_of((b + 1) * geo.V - 1);
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int g = g0; g <= g1; ++g) {
    int lo, hi;
    chunk_range(g, geo.total, lo, hi);
    const int nk = valid_in(b, geo.V, p.L, lo, hi);
    if (nk == 0) continue;
    const float2 v = part[((int64_t)b * geo.kmax + (g - g0)) * p.C + c];
    const float nn = n + (float)nk;
    const float d = v.x - mean;
    mean += d * ((float)nk / nn);
    m2 += v.y + d * d * (n * (float)nk / nn);
    n = nn;
  }
  const float var = m2 / n;
  const float sc = p.gamma[c] * rsqrtf(var + p.gn_eps);
  scale[i] = sc;
  offset[i] = p.beta[c] - mean * sc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

void dwconv_tc_init() {
  FLM_CUDA(cudaFuncSetAttribute(dwconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
}

bool dwconv_tc_supported(const DwFused& p) {
  return p.C % TC == 0 && p.rowstat != nullptr && p.parts >= 2 && p.parts % 2 == 0 && p.tma_encode != nullptr &&
         (reinterpret_cast<uintptr_t>(p.h) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.g) & 15) == 0 &&
         p.u == nullptr && (int64_t)p.B * (p.L + 64) < (1ll << 30);
}

size_t dwconv_tc_part_bytes(int B, int L, int C) {
  const Geo g = make_geo(B, L);
  return (size_t)B * g.kmax * C * sizeof(float2);
}

// LayerNorm-on-load + depthwise conv on the tensor cores + GroupNorm statistics (p.u must be null: see lnu below): writes the
// un-normalised d (into p.g) and the GroupNorm scale / offset (B, C).  Follow with an in-place launch_gn_stream on p.g.
// Scratch: rowconst (B*L float2), ab (B*C float2), part (dwconv_tc_part_bytes).
// gate / bias3 / lnu (all nullable together): adaLN gate of the block (per sample, stride p.mod_bstride), conv_3's bias and the
// (B, 3, C) table conv_3's epilogue uses to recompute u - pass p.u = nullptr with it.
void launch_dwconv_tc(const DwFused& p, float* rowconst, float* ab, float* part, float* scale, float* offset,
                      const float* gate, const float* bias3, float* lnu, int num_sms, cudaStream_t stream) {
  FLM_REQUIRE(dwconv_tc_supported(p) && rowconst && ab && part && scale && offset, "dwconv_tc: unsupported problem");
  if (p.B == 0 || p.L == 0) return;
  const Geo geo = make_geo(p.B, p.L);
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(p.tma_encode);
  CUtensorMap tm;
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.C, (cuuint64_t)p.L, (cuuint64_t)p.B};
    cuuint64_t strides[2] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.C * 2 * (cuuint64_t)p.L};
    cuuint32_t box[3] = {TC, 32, 1}, estr[3] = {1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(p.h), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-2, "cuTensorMapEncodeTiled(dwconv_tc h) failed: " + std::to_string((int)r));
  }
  const int64_t n1 = (int64_t)p.B * p.L + (int64_t)p.B * p.C;
  launch_pdl(ln_consts_kernel, dim3((unsigned)((n1 + 255) / 256)), dim3(256), (size_t)0, stream, p,
             reinterpret_cast<float2*>(rowconst), reinterpret_cast<float2*>(ab), gate, bias3, lnu);
  FLM_LAUNCH_CHECK();
  const int ntiles = (p.C / TC) * geo.npt;
  int grid = num_sms;
  if (grid > ntiles) grid = ntiles;
  launch_pdl(dwconv_tc_kernel, dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, stream, tm, p,
             reinterpret_cast<const float2*>(rowconst), reinterpret_cast<const float2*>(ab), reinterpret_cast<float2*>(part), geo,
             ntiles);
  FLM_LAUNCH_CHECK();
  const int64_t n3 = (int64_t)p.B * p.C;
  launch_pdl(dwtc_merge_kernel, dim3((unsigned)((n3 + 255) / 256)), dim3(256), (size_t)0, stream, p,
             reinterpret_cast<const float2*>(part), geo, scale, offset);
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
