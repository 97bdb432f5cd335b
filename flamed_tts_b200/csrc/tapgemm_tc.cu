// tcgen05 / TMEM / TMA implicit-conv GEMM for sm_100a (bf16 operands, fp32 accumulate).
//
//   out[b,t,n] = epi(sum_tap sum_k A[b, t + off0 + tap*dil, k] * W[tap][n][k] + bias[n])
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  - A tile (128 frames x 64 ch) through a 3-D tensor map (K, T, B) whose
//                               out-of-bounds rows are zero-filled by the TMA unit: that IS the conv
//                               zero padding and the ragged tail; W tile (BLOCK_N x 64) through a 2-D map
//   warp 1      MMA issuer    - tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16 per
//                               instruction, accumulators in TMEM (2 stages x BLOCK_N columns), smem
//                               slots released with tcgen05.commit
//   warps 2..9  epilogue      - two warps per TMEM lane quarter, each taking every other 32-column chunk:
//                               tcgen05.ld (32 lanes x 32 columns), bias + activation fused and stored
//                               straight from registers (64-128 B per thread); the residual epilogues
//                               (adaLN-gated residual, Euler update, codec skip) transpose the chunk
//                               through shared memory so that every global load/store is a coalesced
//                               128 B row segment and all loads of 8 rows are in flight together
// smem ring: STAGES x (16 KiB A + BLOCK_N*128 B W), SWIZZLE_128B on both sides (TMA writes it, the UMMA
// shared-memory descriptor reads it).
//
// Replaces the cuBLAS/cuDNN calls behind every dense projection of the denoiser, the cond
// down-sampler and the FaCodec decoder convolutions.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int STAGE_LD = 33;  // padded row of the epilogue transpose buffer (floats)

template <int BLOCK_N>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages; 128/256/512: power of two
  static constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 32 * STAGE_LD * 4;
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/ +
                                    EPI_STAGE_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B (8 rows x 128 B)
// | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major both,
// N>>3 [17,23), M>>4 [24,29)
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
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct Sched {
  int num_tiles, num_n_tiles, tiles_m_per_b, flatten;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// first row / sample / frame / valid-row count of epilogue quarter `quarter` of M tile `mt`
__device__ __forceinline__ void warp_rows(const TapGemm& p, const Sched& sch, int mt, int quarter, int64_t& m_w, int& b_w,
                                          int& t_w, int& rows_valid) {
  if (sch.flatten) {
    m_w = (int64_t)mt * BLOCK_M + quarter * 32;
    const int64_t total = (int64_t)p.B * p.T_out;
    rows_valid = (int)max((int64_t)0, min((int64_t)32, total - m_w));
    b_w = (int)(m_w / p.T_out);
    t_w = (int)(m_w % p.T_out);
  } else {
    b_w = mt / sch.tiles_m_per_b;
    t_w = (mt % sch.tiles_m_per_b) * BLOCK_M + quarter * 32;
    rows_valid = max(0, min(32, p.T_out - t_w));
    m_w = (int64_t)b_w * p.T_out + t_w;
  }
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// pull the residual-stream / addend rows the epilogue of `tile` will read into L2 while the tensor core is
// still busy with it: lane = row, one bulk prefetch per row covering all BLOCK_N columns of the tile
// (issued by the half-0 warp of each TMEM lane quarter)
template <int BLOCK_N, int EPI>
__device__ __forceinline__ void prefetch_epilogue_operands(const TapGemm& p, const Sched& sch, int tile, int quarter,
                                                           int half, int lane) {
  if (half != 0) return;
  const int nt = tile % sch.num_n_tiles, mt = tile / sch.num_n_tiles;
  int64_t m_w;
  int b_w, t_w, rows_valid;
  warp_rows(p, sch, mt, quarter, m_w, b_w, t_w, rows_valid);
  if (lane >= rows_valid) return;
  const int64_t m = m_w + lane;
  const int n = nt * BLOCK_N;
  if (EPI == EPI_RESID) {
    if (p.out_bf16) prefetch_l2_bulk(static_cast<const bf16*>(p.resid_in) + m * p.ldc + n, BLOCK_N * 2);
    else prefetch_l2_bulk(static_cast<const float*>(p.resid_in) + m * p.ldc + n, BLOCK_N * 4);
  } else if (EPI == EPI_GATE_RESID && p.hres_bf16) {
    prefetch_l2_bulk(reinterpret_cast<const bf16*>(p.hres) + m * p.ld_res + n, BLOCK_N * 2);
  } else {
    prefetch_l2_bulk(p.hres + m * p.ld_res + n, BLOCK_N * 4);
  }
  if (EPI == EPI_GATE_RESID && p.addend) {
    if (p.addend_bf16) prefetch_l2_bulk(static_cast<const bf16*>(p.addend) + m * p.ld_add + n, BLOCK_N * 2);
    else prefetch_l2_bulk(static_cast<const float*>(p.addend) + m * p.ld_add + n, BLOCK_N * 4);
  }
}

// ---------------------------------------------------------------- epilogues
// fast forms for the bf16 mode (results are rounded to bf16 or feed a bf16 GEMM): erf by
// Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), exp by MUFU.EX2
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_approx(z * z * -1.4426950408889634f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float silu_fast(float x) {
  return x * rcp_approx(1.0f + ex2_approx(x * -1.4426950408889634f));
}

template <int EPI>
__device__ __forceinline__ float act_fast(float v) {
  if (EPI == EPI_GELU) return gelu_fast(v);
  if (EPI == EPI_SILU) return silu_fast(v);
  if (EPI == EPI_RELU) return fmaxf(v, 0.0f);
  return v;
}

// (a) activation epilogues: thread = row, 32 consecutive columns from registers
template <int EPI>
__device__ __forceinline__ void epilogue_direct(const TapGemm& p, float (&v)[32], int64_t m, int n) {
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
      v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = act_fast<EPI>(v[j]);
  if (p.out_bf16) {
    bf16* o = static_cast<bf16*>(p.out) + m * p.ldc + n;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      __nv_bfloat162 q0 = __floats2bfloat162_rn(v[j], v[j + 1]);
      __nv_bfloat162 q1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
      __nv_bfloat162 q2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
      __nv_bfloat162 q3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&q0); u.y = *reinterpret_cast<uint32_t*>(&q1);
      u.z = *reinterpret_cast<uint32_t*>(&q2); u.w = *reinterpret_cast<uint32_t*>(&q3);
      *reinterpret_cast<uint4*>(o + j) = u;
    }
  } else {
    float* o = static_cast<float*>(p.out) + m * p.ldc + n;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
}

// (b) residual epilogues: the warp's 32x32 chunk is transposed through `stage` (32 x STAGE_LD floats) so
// that lane = column: every global access of a row is one coalesced 128 B (fp32) / 64 B (bf16) segment.
// All loads of the chunk are issued before the first use.  FULL: all 32 rows valid and in one sample (the
// common case, no predicates, running pointers); otherwise the valid rows are a prefix of length rows_valid.
template <int EPI, bool FULL, typename TRES, typename TADD>
__device__ __forceinline__ void epilogue_rows(const TapGemm& p, const float* stage, int lane, int64_t m_base,
                                              int rows_valid, int b0, int t0, int col, float bias) {
  // residual source / destination for this lane's column
  const TRES* rsrc;
  TRES* rdst;
  int64_t rld;
  if (EPI == EPI_RESID) {
    rsrc = static_cast<const TRES*>(p.resid_in) + m_base * p.ldc + col;
    rdst = static_cast<TRES*>(p.out) + m_base * p.ldc + col;
    rld = p.ldc;
  } else {
    rsrc = reinterpret_cast<const TRES*>(p.hres) + m_base * p.ld_res + col;
    rdst = reinterpret_cast<TRES*>(p.hres) + m_base * p.ld_res + col;
    rld = p.ld_res;
  }
  const TADD* asrc = (EPI == EPI_GATE_RESID && p.addend) ? static_cast<const TADD*>(p.addend) + m_base * p.ld_add + col : nullptr;
  float hv[32], av[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    hv[r] = 0.f; av[r] = 0.f;
    if (FULL || r < rows_valid) {
      hv[r] = ldf<TRES>(rsrc + r * rld);
      if (EPI == EPI_GATE_RESID && asrc) av[r] = ldf<TADD>(asrc + r * p.ld_add);
    }
  }
  float g = 0.f;
  int b = b0, t = t0;
  if (EPI == EPI_GATE_RESID) g = __ldg(p.gate + (int64_t)b0 * p.gate_bstride + col);
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    if (FULL || r < rows_valid) {
      const float val = stage[r * STAGE_LD + lane] + bias;
      float o;
      if (EPI == EPI_GATE_RESID) o = __fadd_rn(hv[r], __fmul_rn(g, val + av[r]));
      else if (EPI == EPI_EULER) o = __fadd_rn(hv[r], __fmul_rn(p.alpha, val));
      else o = hv[r] + val;
      stf<TRES>(rdst + r * rld, o);
      // Euler update: also hand the next step's proj_in GEMM its bf16 operand (p.out, row stride ldc)
      if (EPI == EPI_EULER && p.out) static_cast<bf16*>(p.out)[(m_base + r) * p.ldc + col] = __float2bfloat16(o);
    }
    if (!FULL && EPI == EPI_GATE_RESID) {  // rows may cross into the next sample: reload the gate there
      if (++t == p.T_out) {
        t = 0; ++b;
        if (r + 1 < rows_valid) g = __ldg(p.gate + (int64_t)b * p.gate_bstride + col);
      }
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_transposed(const TapGemm& p, float (&v)[32], float* stage, int lane,
                                                    int64_t m_base, int rows_valid, int b0, int t0, int n) {
  // a quarter that lies entirely beyond the last row (ragged last tile) has nothing to do; its sample index
  // b0 would be out of range, so no operand (gate, bias, stream) may be touched.  rows_valid is warp-uniform.
  if (rows_valid <= 0) return;
#pragma unroll
  for (int j = 0; j < 32; ++j) stage[lane * STAGE_LD + j] = v[j];
  __syncwarp();
  const int col = n + lane;
  const float bias = p.bias ? __ldg(p.bias + col) : 0.f;
  const bool full = rows_valid == 32 && t0 + 32 <= p.T_out;
  const bool res_bf16 = (EPI == EPI_RESID && p.out_bf16) || (EPI == EPI_GATE_RESID && p.hres_bf16);
  const bool add_bf16 = EPI == EPI_GATE_RESID && p.addend_bf16 != 0;
  if (res_bf16) {
    if (full) epilogue_rows<EPI, true, bf16, bf16>(p, stage, lane, m_base, rows_valid, b0, t0, col, bias);
    else epilogue_rows<EPI, false, bf16, bf16>(p, stage, lane, m_base, rows_valid, b0, t0, col, bias);
  } else if (add_bf16) {
    if (full) epilogue_rows<EPI, true, float, bf16>(p, stage, lane, m_base, rows_valid, b0, t0, col, bias);
    else epilogue_rows<EPI, false, float, bf16>(p, stage, lane, m_base, rows_valid, b0, t0, col, bias);
  } else {
    if (full) epilogue_rows<EPI, true, float, float>(p, stage, lane, m_base, rows_valid, b0, t0, col, bias);
    else epilogue_rows<EPI, false, float, float>(p, stage, lane, m_base, rows_valid, b0, t0, col, bias);
  }
  __syncwarp();
}

// ---------------------------------------------------------------- the kernel
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tapgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TapGemm p,
                  const Sched sch) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + C::STAGES * C::B_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_trigger();  // programmatic dependent launch: the set-up above overlapped the previous kernel's tail
  pdl_wait();

  const int kblocks = p.K / BLOCK_K;
  const int iters_per_tile = p.ntaps * kblocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < sch.num_tiles; tile += gridDim.x) {
      const int nt = tile % sch.num_n_tiles, mt = tile / sch.num_n_tiles;
      int b, t0;
      if (sch.flatten) { b = 0; t0 = mt * BLOCK_M; }
      else { b = mt / sch.tiles_m_per_b; t0 = (mt % sch.tiles_m_per_b) * BLOCK_M; }
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int trow = t0 + p.off0 + tap * p.dil;
        const int wrow = tap * p.N + nt * BLOCK_N;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (lane == 0) {
            mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + C::B_STAGE_BYTES);
            tma_load_3d(&tmA, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, trow, b);
            tma_load_2d(&tmB, &full_bar[stage], smem_b + stage * C::B_STAGE_BYTES, kb * BLOCK_K, wrow);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < sch.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
      for (int i = 0; i < iters_per_tile; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * C::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) address field
            umma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (i == iters_per_tile - 1) umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9: TMEM lane quarter = warp % 4, column half = (warp-2)/4) =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    float* stage = epi_stage + (warp - 2) * 32 * STAGE_LD;
    constexpr bool kTransposed = (EPI == EPI_RESID || EPI == EPI_GATE_RESID || EPI == EPI_EULER);
    int it = 0;
    if (kTransposed && (int)blockIdx.x < sch.num_tiles)
      prefetch_epilogue_operands<BLOCK_N, EPI>(p, sch, blockIdx.x, quarter, half, lane);
    for (int tile = blockIdx.x; tile < sch.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int nt = tile % sch.num_n_tiles, mt = tile / sch.num_n_tiles;
      if (kTransposed && tile + (int)gridDim.x < sch.num_tiles)
        prefetch_epilogue_operands<BLOCK_N, EPI>(p, sch, tile + gridDim.x, quarter, half, lane);
      // the warp's first row: global row m_w, sample b_w, frame t_w; its valid rows are a prefix
      int64_t m_w;
      int b_w, t_w, rows_valid;
      warp_rows(p, sch, mt, quarter, m_w, b_w, t_w, rows_valid);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N);
#pragma unroll 1
      for (int ch = half; ch < BLOCK_N / 32; ch += 2) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)(ch * 32), v);
        const int n = nt * BLOCK_N + ch * 32;
        if (kTransposed) {
          epilogue_transposed<EPI>(p, v, stage, lane, m_w, rows_valid, b_w, t_w, n);
        } else {
          if (lane < rows_valid) epilogue_direct<EPI>(p, v, m_w + lane, n);
        }
      }
      (void)row;
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BLOCK_N, int EPI>
void launch_one(const TapGemm& p, const CUtensorMap& tmA, const CUtensorMap& tmB, const Sched& sch, int num_sms,
                cudaStream_t stream) {
  using C = Cfg<BLOCK_N>;
  const int grid = sch.num_tiles < num_sms ? sch.num_tiles : num_sms;
  launch_pdl(tapgemm_tc_kernel<BLOCK_N, EPI>, dim3(grid), dim3(NUM_THREADS), (size_t)C::SMEM_BYTES, stream, tmA, tmB, p, sch);
  FLM_LAUNCH_CHECK();
}

template <int BLOCK_N>
void launch_cfg(const TapGemm& p, const CUtensorMap& tmA, const CUtensorMap& tmB, const Sched& sch, int num_sms,
                cudaStream_t stream) {
  switch (p.epi) {
    case EPI_NONE: launch_one<BLOCK_N, EPI_NONE>(p, tmA, tmB, sch, num_sms, stream); break;
    case EPI_GELU: launch_one<BLOCK_N, EPI_GELU>(p, tmA, tmB, sch, num_sms, stream); break;
    case EPI_SILU: launch_one<BLOCK_N, EPI_SILU>(p, tmA, tmB, sch, num_sms, stream); break;
    case EPI_RELU: launch_one<BLOCK_N, EPI_RELU>(p, tmA, tmB, sch, num_sms, stream); break;
    case EPI_RESID: launch_one<BLOCK_N, EPI_RESID>(p, tmA, tmB, sch, num_sms, stream); break;
    case EPI_GATE_RESID: launch_one<BLOCK_N, EPI_GATE_RESID>(p, tmA, tmB, sch, num_sms, stream); break;
    case EPI_EULER: launch_one<BLOCK_N, EPI_EULER>(p, tmA, tmB, sch, num_sms, stream); break;
    default: throw Error(-1, "tapgemm_tc: unknown epilogue");
  }
}

template <int BLOCK_N, int EPI>
void set_attr_one() {
  FLM_CUDA(cudaFuncSetAttribute(tapgemm_tc_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg<BLOCK_N>::SMEM_BYTES));
}
template <int BLOCK_N>
void set_attr_all() {
  set_attr_one<BLOCK_N, EPI_NONE>(); set_attr_one<BLOCK_N, EPI_GELU>(); set_attr_one<BLOCK_N, EPI_SILU>();
  set_attr_one<BLOCK_N, EPI_RELU>(); set_attr_one<BLOCK_N, EPI_RESID>(); set_attr_one<BLOCK_N, EPI_GATE_RESID>();
  set_attr_one<BLOCK_N, EPI_EULER>();
}

}  // namespace

// must run once per process before the first launch (and outside any stream capture)
void tapgemm_tc_init() {
  set_attr_all<256>();
  set_attr_all<128>();
  set_attr_all<64>();
}

static int64_t sch_rows(const TapGemm& p) {
  const bool flatten = p.ntaps == 1 && p.off0 == 0;
  return flatten ? (int64_t)p.B * p.T_in : (int64_t)p.B * (((int64_t)p.T_out + BLOCK_M - 1) / BLOCK_M) * BLOCK_M;
}

bool tapgemm_tc_supported(const TapGemm& p) {
  return p.stride == 1 && p.T_in == p.T_out && p.K % BLOCK_K == 0 && p.N % 64 == 0 && p.lda % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(p.A) % 16 == 0) && (reinterpret_cast<uintptr_t>(p.W) % 16 == 0);
}

void launch_tapgemm_tc(const TapGemm& p, void* tma_encode, int num_sms, cudaStream_t stream) {
  FLM_REQUIRE(tapgemm_tc_supported(p), "tapgemm_tc: unsupported shape (need stride 1, K%64==0, N%64==0)");
  FLM_REQUIRE(tma_encode != nullptr, "tapgemm_tc: cuTensorMapEncodeTiled entry point not resolved");
  if ((int64_t)p.B * p.T_out == 0) return;
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(tma_encode);
  // widest N tile that still gives every SM a tile (small-M problems are latency bound: prefer more CTAs)
  int BN = (p.N % 256 == 0) ? 256 : (p.N % 128 == 0 ? 128 : 64);
  {
    const int64_t rows = sch_rows(p);
    while (BN > 64 && p.N % (BN / 2) == 0 && ((rows + BLOCK_M - 1) / BLOCK_M) * (p.N / BN) < num_sms) BN /= 2;
  }
  Sched sch;
  sch.flatten = (p.ntaps == 1 && p.off0 == 0) ? 1 : 0;
  sch.num_n_tiles = p.N / BN;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BLOCK_M, 1}, estr[3] = {1, 1, 1};
    if (sch.flatten) {
      dims[0] = (cuuint64_t)p.K; dims[1] = (cuuint64_t)p.B * p.T_in; dims[2] = 1;
      sch.tiles_m_per_b = (int)((dims[1] + BLOCK_M - 1) / BLOCK_M);
      sch.num_tiles = sch.tiles_m_per_b * sch.num_n_tiles;
    } else {
      dims[0] = (cuuint64_t)p.K; dims[1] = (cuuint64_t)p.T_in; dims[2] = (cuuint64_t)p.B;
      sch.tiles_m_per_b = (p.T_out + BLOCK_M - 1) / BLOCK_M;
      sch.num_tiles = sch.tiles_m_per_b * p.B * sch.num_n_tiles;
    }
    strides[0] = (cuuint64_t)p.lda * 2;
    strides[1] = (cuuint64_t)p.lda * 2 * (sch.flatten ? dims[1] : (cuuint64_t)p.T_in);
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p.A), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-2, "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.ntaps * p.N};
    cuuint64_t strides[1] = {(cuuint64_t)p.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BN}, estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p.W), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-2, "cuTensorMapEncodeTiled(W) failed: " + std::to_string((int)r));
  }
  if (BN == 256) launch_cfg<256>(p, tmA, tmB, sch, num_sms, stream);
  else if (BN == 128) launch_cfg<128>(p, tmA, tmB, sch, num_sms, stream);
  else launch_cfg<64>(p, tmA, tmB, sch, num_sms, stream);
}

}  // namespace flm
