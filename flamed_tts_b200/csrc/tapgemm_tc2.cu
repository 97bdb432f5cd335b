// tcgen05 implicit-conv GEMM, second generation: CTA pairs (cta_group::2) + all epilogue traffic through TMA.
//
//   out[b,t,n] = epi(sum_tap sum_k A[b, t + off0 + tap*dil, k] * W[tap][n][k] + bias[n])      bf16 in / bf16 out
//
// Why a second kernel (measured on B200, profiles/r01b): with 128x256 tiles per single CTA the K=1024 GEMMs
// pull 48 KB per k-block per SM out of L2 - 14.5 TB/s chip-wide at the mainloop's own speed, which is the L2
// slice throughput limit, and the residual epilogues (thread-per-column loads after a shared-memory
// transpose) never had more than ~48 KB per SM in flight, so they ran at 13.8 us per tile against 4.3 us of
// tensor work.  Here
//   * two CTAs of one TPC form a pair that computes a 256 x BLOCK_N tile with tcgen05.mma.cta_group::2: each
//     CTA loads its own 128 rows of A and HALF of the W tile, so the L2 -> SM traffic per FLOP drops by 1/3
//     and the operand ring costs 32 KB instead of 48 KB per stage;
//   * the epilogue moves whole 128-row x 64-column bf16 slabs (16 KB, SWIZZLE_128B) with TMA: a loader warp
//     streams the residual stream (and the ConvNeXt inner-residual addend) slabs in ahead of the accumulator, the
//     8 epilogue warps work thread-per-row straight out of TMEM on the swizzled slab (conflict-free 16 B
//     shared-memory accesses, no transpose), and a store warp writes the finished slab back with one bulk
//     tensor store.  Rows beyond the end of a sample / of the problem are clipped by the TMA unit.
//
// Roles (12 warps per CTA): 0 TMA producer (A/W ring), 1 MMA issuer (leader CTA only), 2..9 epilogue,
// 10 epilogue-operand loader, 11 slab store.  Handles the bf16-output epilogues (none/GELU/SiLU/ReLU), the
// codec skip connection (EPI_RESID, bf16) and the adaLN-gated residual with a bf16 residual stream; fp32
// outputs / fp32 residual streams / the Euler update stay on the first-generation kernel (tapgemm_tc.cu).
#include <cuda.h>

#include <unordered_map>

#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {

constexpr int BLOCK_M = 128;   // rows per CTA (256 per pair)
constexpr int BLOCK_K = 64;    // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int SLAB_COLS = 64;
constexpr int SLAB_BYTES = BLOCK_M * SLAB_COLS * 2;  // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 12 * 32;
constexpr int MAX_STAGES = 12;
constexpr int MAX_SLAB_BUFS = 8;
constexpr int BAR_BYTES = 1024;
constexpr int SMEM_LIMIT = 232448;  // 227 KB

struct Sched2 {
  int num_tiles;      // pair tiles = ceil(mt_count / 2) * num_n_tiles
  int num_n_tiles;
  int tiles_m_per_b;  // 128-row tiles per sample (non-flattened) / in total (flattened)
  int mt_count;       // 128-row tiles in the problem
  int flatten;
  int stages;         // operand ring depth
  int nbuf;           // epilogue slab buffers
  int store_depth;    // slab stores whose shared-memory read may still be pending when the next one is issued
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on a barrier of the leader CTA.  Default (.release.cta) semantics on purpose: what the barrier orders is
// TMEM traffic (tcgen05.fence::before_thread_sync precedes it); a .release.cluster here costs an ERRBAR per arrive
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
// Waits on barriers that the peer CTA / the multicast tcgen05.commit arrive on use the same CTA-scope wait: the
// data they guard moves through the async proxy (TMA -> smem, tcgen05 -> TMEM), never through generic-proxy
// global/shared accesses of the peer.  An .acquire.cluster wait makes ptxas emit CCTL.IVALL after every wait: the L1
// invalidation evicted the bias / gate vectors every tile (ncu: 30 % of the epilogue's stall samples, profiles/r01k).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// operand loads of a CTA pair: the data lands in the issuing CTA's shared memory, the bytes are counted on the
// LEADER CTA's barrier (`bar_cluster` = mapa(bar, 0)), which is the one the MMA issuer waits on
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// CTA-local slab load / store
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// returns once at most `pending` (0..2) of the committed slab stores have not finished reading shared memory
__device__ __forceinline__ void tma_store_wait_read_upto(int pending) {
  if (pending <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (pending == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (same encoding as tapgemm_tc.cu)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs of the pair once the MMAs issued so far
// have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
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

// ---------------------------------------------------------------- epilogue math (bf16 mode: fast forms)
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
// GELU(x) = x * Phi(x) without MUFU: t = clamp(x, +-4.5), Phi(t) = 0.5 + t * P(t^2) with a degree-8 minimax P
// (|Phi error| <= 1.7e-5), tails restored exactly by max(x - t, 0); max |error| of the whole form 1.2e-4 (at
// x = -4.4), i.e. below the bf16 rounding the result gets anyway.  Two elements per instruction (FFMA2): the
// previous erf/exp form cost 2 MUFU + 16 ALU per element and made the GELU GEMM epilogue-bound.
__device__ __forceinline__ void gelu_pair(float& a, float& b) {
  const float ta = fminf(fmaxf(a, -4.5f), 4.5f), tb = fminf(fmaxf(b, -4.5f), 4.5f);
  const f32x2 t = pack2(ta, tb);
  const f32x2 u = mul2(t, t);
  f32x2 pl = pack2(3.805699895e-11f, 3.805699895e-11f);
  pl = fma2(pl, u, pack2(-4.002322479e-09f, -4.002322479e-09f));
  pl = fma2(pl, u, pack2(1.846168244e-07f, 1.846168244e-07f));
  pl = fma2(pl, u, pack2(-4.959740917e-06f, -4.959740917e-06f));
  pl = fma2(pl, u, pack2(8.727668674e-05f, 8.727668674e-05f));
  pl = fma2(pl, u, pack2(-1.076739452e-03f, -1.076739452e-03f));
  pl = fma2(pl, u, pack2(9.729491531e-03f, 9.729491531e-03f));
  pl = fma2(pl, u, pack2(-6.624043811e-02f, -6.624043811e-02f));
  pl = fma2(pl, u, pack2(3.988664886e-01f, 3.988664886e-01f));
  const f32x2 phi = fma2(t, pl, pack2(0.5f, 0.5f));
  const f32x2 g = fma2(t, phi, pack2(fmaxf(a - ta, 0.0f), fmaxf(b - tb, 0.0f)));
  unpack2(g, a, b);
}
// SiLU(x) = x * sigmoid(x) = x * (0.5 + 0.5 * tanh(x / 2)): ONE MUFU op per element (tanh.approx, max relative error
// 2^-11, i.e. below the bf16 rounding of the result) instead of ex2 + rcp - the SiLU GEMM's epilogue was MUFU-bound
__device__ __forceinline__ float silu_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
template <int EPI>
__device__ __forceinline__ void act_fast32(float (&v)[32]) {
  if (EPI == EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_pair(v[j], v[j + 1]);
  } else if (EPI == EPI_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = silu_fast(v[j]);
  } else if (EPI == EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float* f) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 o = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&o);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// tile -> first row coordinates of this CTA's 128-row half: TMA coordinates (row, sample) and the flat row index
struct TileRows {
  int row;     // TMA coordinate 1 (frame inside the sample, or flat row when flattened)
  int b;       // TMA coordinate 2 (sample, 0 when flattened)
};
__device__ __forceinline__ TileRows tile_rows(const Sched2& sch, int mt) {
  TileRows r;
  if (sch.flatten) { r.b = 0; r.row = mt * BLOCK_M; }
  else { r.b = mt / sch.tiles_m_per_b; r.row = (mt % sch.tiles_m_per_b) * BLOCK_M; }
  return r;
}

constexpr bool epi_loads_residual(int epi) { return epi == EPI_RESID || epi == EPI_GATE_RESID; }

// ---------------------------------------------------------------- the kernel
template <int BLOCK_N, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
tapgemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                   const __grid_constant__ CUtensorMap tmD, const TapGemm p, const Sched2 sch) {
  constexpr int B_STAGE_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;  // this CTA's half of the W tile
  constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  constexpr int SLABS = BLOCK_N / SLAB_COLS;
  constexpr bool kResid = epi_loads_residual(EPI);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const bool has_addend = kResid && EPI == EPI_GATE_RESID && p.addend != nullptr;
  const bool has_out2 = kResid && EPI == EPI_GATE_RESID && p.out2 != nullptr;  // second output slab ring (lnu path)
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + sch.stages * A_STAGE_BYTES;
  uint8_t* smem_r = smem_b + sch.stages * B_STAGE_BYTES;
  uint8_t* smem_d = smem_r + sch.nbuf * SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_d + ((has_addend || has_out2) ? sch.nbuf * SLAB_BYTES : 0));
  uint64_t* full_bar = bars;                       // [MAX_STAGES] operand stage filled (leader CTA's is the one used)
  uint64_t* empty_bar = bars + MAX_STAGES;         // [MAX_STAGES] operand stage consumed (multicast commit)
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;     // [2] accumulator stage complete (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;            // [2] accumulator stage drained by both CTAs (leader's is used)
  uint64_t* r_full = tmem_empty + 2;               // [MAX_SLAB_BUFS] slab operands landed
  uint64_t* r_free = r_full + MAX_SLAB_BUFS;       // [MAX_SLAB_BUFS] slab buffer reusable (its store has read it)
  uint64_t* r_ready = r_free + MAX_SLAB_BUFS;      // [MAX_SLAB_BUFS] slab results written by the 8 epilogue warps
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(r_ready + MAX_SLAB_BUFS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vw = warp;  // role index: 0 producer, 1 MMA issuer, 2..9 epilogue, 10 slab loader, 11 slab store
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (vw == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (kResid) tma_prefetch_desc(&tmR);
    if (has_addend || has_out2) tma_prefetch_desc(&tmD);
    for (int i = 0; i < sch.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * NUM_EPI_WARPS);  // one arrival per epilogue warp of both CTAs
    }
    for (int i = 0; i < sch.nbuf; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_free[i], 1);
      mbar_init(&r_ready[i], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (vw == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above overlapped the tail of the previous kernel (programmatic dependent launch); its results are needed now
  pdl_trigger();
  pdl_wait();

  const int kblocks = p.K / BLOCK_K;
  const int iters_per_tile = p.ntaps * kblocks;

  if (vw == 0) {
    // ===================== operand producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < sch.num_tiles; tile += npairs) {
      const int nt = tile % sch.num_n_tiles, mp = tile / sch.num_n_tiles;
      const TileRows tr = tile_rows(sch, 2 * mp + (int)rank);
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int trow = tr.row + p.off0 + tap * p.dil;
        const int wrow = tap * p.N + nt * BLOCK_N + (int)rank * (BLOCK_N / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
          if (lane == 0) {
            const uint32_t leader_full = mapa(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES));
            tma_load_3d_pair(&tmA, leader_full, smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, trow, tr.b);
            tma_load_2d_pair(&tmB, leader_full, smem_b + stage * B_STAGE_BYTES, kb * BLOCK_K, wrow);
          }
          __syncwarp();
          if (++stage == sch.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (vw == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(2 * BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < sch.num_tiles; tile += npairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait_cluster(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int i = 0; i < iters_per_tile; ++i) {
          mbar_wait_cluster(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
            const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * B_STAGE_BYTES));
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_f16_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);  // frees the slot in both CTAs
            if (i == iters_per_tile - 1) umma_commit_pair(&tmem_full[as]);
          }
          __syncwarp();
          if (++stage == sch.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (vw < 2 + NUM_EPI_WARPS) {
    // ===================== epilogue: thread = row, 32 columns of each 64-column slab =====================
    const int quarter = warp & 3;         // TMEM lane quarter this warp may read
    const int half = (vw - 2) >> 2;       // which 32 columns of the slab
    const int rrow = quarter * 32 + lane; // row inside the CTA's 128-row tile
    const uint32_t leader_tmem_empty0 = mapa(smem_u32(&tmem_empty[0]), 0);
    int it = 0;
    int buf = 0;
    uint32_t bphase = 0;
    for (int tile = pair; tile < sch.num_tiles; tile += npairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int nt = tile % sch.num_n_tiles, mp = tile / sch.num_n_tiles;
      const int mt = 2 * mp + (int)rank;
      // sample of this thread's row (gate index); clamped for rows the TMA store will clip
      int bidx = 0;
      if (EPI == EPI_GATE_RESID || (EPI == EPI_SILU && p.raff_c1 != nullptr)) {
        if (sch.flatten) bidx = (int)(((unsigned)mt * BLOCK_M + (unsigned)rrow) / (unsigned)p.T_out);
        else bidx = mt / sch.tiles_m_per_b;
        bidx = min(bidx, p.B - 1);
      }
      f32x2 raff_rs = 0ull, raff_nm = 0ull;  // (rstd, -mean * rstd) of this thread's row (row-affine SiLU epilogue)
      if (EPI == EPI_SILU && p.raff_c1 != nullptr) {
        int64_t grow;
        bool ok;
        if (sch.flatten) {
          grow = (int64_t)mt * BLOCK_M + rrow;
          ok = grow < (int64_t)p.B * p.T_out;
        } else {
          const int tb = mt / sch.tiles_m_per_b, tr = (mt % sch.tiles_m_per_b) * BLOCK_M + rrow;
          grow = (int64_t)tb * p.T_out + tr;
          ok = tb < p.B && tr < p.T_out;
        }
        if (ok) {
          const float4* ps = reinterpret_cast<const float4*>(p.raff_rowstat) + (grow * p.raff_parts >> 1);
          float sm = 0.f, sq = 0.f;
          for (int k = 0; 2 * k < p.raff_parts; ++k) {
            const float4 v4 = __ldg(ps + k);
            sm += v4.x + v4.z;
            sq += v4.y + v4.w;
          }
          const float inv_c = 1.0f / (float)p.raff_ln_dim;
          const float mean = sm * inv_c;
          const float rstd = rsqrtf(fmaxf(sq * inv_c - mean * mean, 0.f) + p.raff_eps);
          raff_rs = pack2(rstd, rstd);
          raff_nm = pack2(-mean * rstd, -mean * rstd);
        }
      }
      mbar_wait_cluster(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N);
      f32x2 st_s = 0ull, st_q = 0ull;  // LayerNorm row statistics of this thread's columns of the tile (p.rowstat)
      if (EPI == EPI_GATE_RESID && p.lnu_table != nullptr) {
        // inner residual recomputed from the residual-stream slab:  h += G acc + GA (h rstd + nm) + GB
        f32x2 rs2 = 0ull, nm2 = 0ull;
        {
          int64_t grow;
          bool ok;
          if (sch.flatten) {
            grow = (int64_t)mt * BLOCK_M + rrow;
            ok = grow < (int64_t)p.B * p.T_out;
          } else {
            const int tb = mt / sch.tiles_m_per_b, tr = (mt % sch.tiles_m_per_b) * BLOCK_M + rrow;
            grow = (int64_t)tb * p.T_out + tr;
            ok = tb < p.B && tr < p.T_out;
          }
          if (ok) {
            const float2 rc = __ldg(reinterpret_cast<const float2*>(p.lnu_rowconst) + grow);
            rs2 = pack2(rc.x, rc.x);
            nm2 = pack2(rc.y, rc.y);
          }
        }
        const float* tab = p.lnu_table + (int64_t)bidx * p.lnu_vecs * p.N;
#pragma unroll 1
        for (int s = 0; s < SLABS; ++s) {
          uint8_t* rrow_ptr = smem_r + buf * SLAB_BYTES + rrow * 128;
#pragma unroll 1
          for (int q = 0; q < 2; ++q) {
            const int n = nt * BLOCK_N + s * SLAB_COLS + half * 32 + q * 16;
            float4 G[4], GA[4], GB[4], A2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              G[j] = __ldg(reinterpret_cast<const float4*>(tab + n) + j);
              GA[j] = __ldg(reinterpret_cast<const float4*>(tab + p.N + n) + j);
              GB[j] = __ldg(reinterpret_cast<const float4*>(tab + 2 * p.N + n) + j);
              if (has_out2) A2[j] = __ldg(reinterpret_cast<const float4*>(tab + 3 * p.N + n) + j);
            }
            float v[16];
            tmem_ld16(taddr + (uint32_t)(s * SLAB_COLS + half * 32 + q * 16), v);
            if (q == 0) mbar_wait(&r_full[buf], bphase);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint4* slot = reinterpret_cast<uint4*>(rrow_ptr + (((half * 4 + q * 2 + c) ^ (rrow & 7)) << 4));
              const uint4 hq = *slot;
              const uint32_t w[4] = {hq.x, hq.y, hq.z, hq.w};
              uint32_t o[4], o2[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const f32x2 h2 = pack2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u));
                const f32x2 acc2 = pack2(v[c * 8 + 2 * j], v[c * 8 + 2 * j + 1]);
                const float4 g = G[c * 2 + (j >> 1)], ga = GA[c * 2 + (j >> 1)], gb = GB[c * 2 + (j >> 1)];
                const f32x2 t2 = fma2(h2, rs2, nm2);
                f32x2 r2 = fma2((j & 1) ? pack2(g.z, g.w) : pack2(g.x, g.y), acc2, h2);
                r2 = fma2((j & 1) ? pack2(ga.z, ga.w) : pack2(ga.x, ga.y), t2, r2);
                r2 = add2(r2, (j & 1) ? pack2(gb.z, gb.w) : pack2(gb.x, gb.y));
                if (p.rowstat) {
                  st_s = add2(st_s, r2);
                  st_q = fma2(r2, r2, st_q);
                }
                float lo, hi;
                unpack2(r2, lo, hi);
                __nv_bfloat162 tt = __floats2bfloat162_rn(lo, hi);
                o[j] = *reinterpret_cast<uint32_t*>(&tt);
                if (has_out2) {  // the column-scaled copy for the next LayerNorm's GEMM
                  const float4 a = A2[c * 2 + (j >> 1)];
                  unpack2(mul2(r2, (j & 1) ? pack2(a.z, a.w) : pack2(a.x, a.y)), lo, hi);
                  tt = __floats2bfloat162_rn(lo, hi);
                  o2[j] = *reinterpret_cast<uint32_t*>(&tt);
                }
              }
              *slot = make_uint4(o[0], o[1], o[2], o[3]);
              if (has_out2)
                *reinterpret_cast<uint4*>(smem_d + buf * SLAB_BYTES + rrow * 128 + (((half * 4 + q * 2 + c) ^ (rrow & 7)) << 4)) =
                    make_uint4(o2[0], o2[1], o2[2], o2[3]);
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&r_ready[buf]);
          if (++buf == sch.nbuf) { buf = 0; bphase ^= 1; }
        }
      } else {
#pragma unroll 1
      for (int s = 0; s < SLABS; ++s) {
        const int n = nt * BLOCK_N + s * SLAB_COLS + half * 32;
        // per-column operands first: their (L1) latency hides behind the TMEM load and the slab barrier
        float4 bq[8], gq[8];
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
        }
        if (EPI == EPI_GATE_RESID) {
          const float4* gp = reinterpret_cast<const float4*>(p.gate + (int64_t)bidx * p.gate_bstride + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) gq[j] = __ldg(gp + j);
        }
        float v[32];
        tmem_ld32(taddr + (uint32_t)(s * SLAB_COLS + half * 32), v);
        // packed fp32x2 from here on: a2[i] = columns (2i, 2i+1)
        f32x2 a2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a2[i] = pack2(v[2 * i], v[2 * i + 1]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a2[2 * j] = add2(a2[2 * j], pack2(bq[j].x, bq[j].y));
            a2[2 * j + 1] = add2(a2[2 * j + 1], pack2(bq[j].z, bq[j].w));
          }
        }
        uint8_t* rrow_ptr = smem_r + buf * SLAB_BYTES + rrow * 128;
        if (kResid) {
          mbar_wait(&r_full[buf], bphase);
          if (EPI == EPI_GATE_RESID && has_addend) {
            const uint8_t* drow_ptr = smem_d + buf * SLAB_BYTES + rrow * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 q = *reinterpret_cast<const uint4*>(drow_ptr + (((half * 4 + c) ^ (rrow & 7)) << 4));
              const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
                a2[c * 4 + j] = add2(a2[c * 4 + j], pack2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u)));
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4* slot = reinterpret_cast<uint4*>(rrow_ptr + (((half * 4 + c) ^ (rrow & 7)) << 4));
            const uint4 q = *slot;
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const f32x2 h2 = pack2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u));
              f32x2 r2;
              if (EPI == EPI_GATE_RESID) {  // h + g * (acc + bias + u): one fused multiply-add per element (bf16 result)
                const float4 g = gq[c * 2 + (j >> 1)];
                r2 = fma2((j & 1) ? pack2(g.z, g.w) : pack2(g.x, g.y), a2[c * 4 + j], h2);
              } else {
                r2 = add2(h2, a2[c * 4 + j]);
              }
              if (EPI == EPI_GATE_RESID && p.rowstat) {
                st_s = add2(st_s, r2);
                st_q = fma2(r2, r2, st_q);
              }
              float lo, hi;
              unpack2(r2, lo, hi);
              __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
              o[j] = *reinterpret_cast<uint32_t*>(&t);
            }
            *slot = make_uint4(o[0], o[1], o[2], o[3]);
          }
        } else {
          if (EPI == EPI_SILU && p.raff_c1 != nullptr) {
            // LayerNorm of the previous layer applied algebraically: value = rstd_r * acc + (nm_r * c1 + c2)
            const float4* c1p = reinterpret_cast<const float4*>(p.raff_c1 + (int64_t)bidx * p.N + n);
            const float4* c2p = reinterpret_cast<const float4*>(p.raff_c2 + (int64_t)bidx * p.N + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 c1 = __ldg(c1p + j), c2 = __ldg(c2p + j);
              a2[2 * j] = fma2(a2[2 * j], raff_rs, fma2(pack2(c1.x, c1.y), raff_nm, pack2(c2.x, c2.y)));
              a2[2 * j + 1] = fma2(a2[2 * j + 1], raff_rs, fma2(pack2(c1.z, c1.w), raff_nm, pack2(c2.z, c2.w)));
            }
          }
#pragma unroll
          if (EPI == EPI_NONE && p.rowstat) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              st_s = add2(st_s, a2[i]);
              st_q = fma2(a2[i], a2[i], st_q);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) unpack2(a2[i], v[2 * i], v[2 * i + 1]);
          act_fast32<EPI>(v);
          mbar_wait(&r_free[buf], bphase ^ 1);  // the previous store out of this buffer has read it
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(rrow_ptr + (((half * 4 + c) ^ (rrow & 7)) << 4)) = pack_bf16x8(&v[c * 8]);
        }
        fence_async_smem();  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(&r_ready[buf]);
        if (++buf == sch.nbuf) { buf = 0; bphase ^= 1; }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + (uint32_t)(as * 8));
      if ((EPI == EPI_NONE || EPI == EPI_GATE_RESID) && p.rowstat) {
        // global row of this thread (rows the TMA store clips are skipped)
        int64_t grow;
        bool ok;
        if (sch.flatten) {
          grow = (int64_t)mt * BLOCK_M + rrow;
          ok = grow < (int64_t)p.B * p.T_out;
        } else {
          const int tb = mt / sch.tiles_m_per_b, tr = (mt % sch.tiles_m_per_b) * BLOCK_M + rrow;
          grow = (int64_t)tb * p.T_out + tr;
          ok = tb < p.B && tr < p.T_out;
        }
        if (ok) {
          float s0, s1, q0, q1;
          unpack2(st_s, s0, s1);
          unpack2(st_q, q0, q1);
          *reinterpret_cast<float2*>(p.rowstat + (grow * p.rowstat_parts + nt * 2 + half) * 2) = make_float2(s0 + s1, q0 + q1);
        }
      }
    }
  } else if (vw == 2 + NUM_EPI_WARPS) {
    // ===================== epilogue operand loader =====================
    if (kResid && lane == 0) {
      int buf = 0;
      uint32_t bphase = 0;
      for (int tile = pair; tile < sch.num_tiles; tile += npairs) {
        const int nt = tile % sch.num_n_tiles, mp = tile / sch.num_n_tiles;
        const TileRows tr = tile_rows(sch, 2 * mp + (int)rank);
        for (int s = 0; s < SLABS; ++s) {
          mbar_wait(&r_free[buf], bphase ^ 1);
          mbar_expect_tx(&r_full[buf], has_addend ? 2 * SLAB_BYTES : SLAB_BYTES);
          const int col = nt * BLOCK_N + s * SLAB_COLS;
          tma_load_3d(&tmR, &r_full[buf], smem_r + buf * SLAB_BYTES, col, tr.row, tr.b);
          if (has_addend) tma_load_3d(&tmD, &r_full[buf], smem_d + buf * SLAB_BYTES, col, tr.row, tr.b);
          if (++buf == sch.nbuf) { buf = 0; bphase ^= 1; }
        }
      }
    }
  } else {
    // ===================== slab store =====================
    // Up to `store_depth` stores stay in flight: waiting for each store's shared-memory read before looking at the next
    // slab serialised the four slabs of a tile on the store latency.  Bulk groups complete in order, so once at most
    // `depth` groups are pending the buffer of the store issued `depth` slabs ago is reusable.
    if (lane == 0) {
      const int depth = sch.store_depth;
      int buf = 0, fbuf = 0, inflight = 0;
      uint32_t bphase = 0;
      for (int tile = pair; tile < sch.num_tiles; tile += npairs) {
        const int nt = tile % sch.num_n_tiles, mp = tile / sch.num_n_tiles;
        const TileRows tr = tile_rows(sch, 2 * mp + (int)rank);
        for (int s = 0; s < SLABS; ++s) {
          mbar_wait(&r_ready[buf], bphase);
          tma_store_3d(&tmO, smem_r + buf * SLAB_BYTES, nt * BLOCK_N + s * SLAB_COLS, tr.row, tr.b);
          if (has_out2) tma_store_3d(&tmD, smem_d + buf * SLAB_BYTES, nt * BLOCK_N + s * SLAB_COLS, tr.row, tr.b);
          tma_store_commit();
          if (++buf == sch.nbuf) { buf = 0; bphase ^= 1; }
          if (++inflight > depth) {
            tma_store_wait_read_upto(depth);
            mbar_arrive(&r_free[fbuf]);
            if (++fbuf == sch.nbuf) fbuf = 0;
            --inflight;
          }
        }
      }
      tma_store_wait_read();
      for (; inflight > 0; --inflight) {
        mbar_arrive(&r_free[fbuf]);
        if (++fbuf == sch.nbuf) fbuf = 0;
      }
      tma_store_wait_all();  // global writes complete before the CTA retires
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer may still multicast into / read from this CTA's shared memory until here
  if (vw == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BLOCK_N>
constexpr int stage_bytes() { return A_STAGE_BYTES + (BLOCK_N / 2) * BLOCK_K * 2; }

template <int BLOCK_N, int EPI>
void launch_one(const TapGemm& p, const CUtensorMap* tm, Sched2 sch, int num_sms, cudaStream_t stream) {
  const bool addend = EPI == EPI_GATE_RESID && (p.addend != nullptr || p.out2 != nullptr);  // two slab rings
  const bool resid = epi_loads_residual(EPI);
  // slab buffers: enough loads in flight to cover HBM latency at the tensor-core rate; the rest goes to the ring
  sch.nbuf = addend ? 3 : (resid ? 4 : 3);
  // (two rings: 3 slab pairs measured best - 2 pairs 0.175 ms, 3 pairs 0.149 ms, 4 pairs 0.163 ms at M = 76 800; 4 residual +
  //  2 second-output buffers with separate hand-offs: GEMM class 1.455 vs 1.423 ms per velocity, profiles/r2z)
  sch.store_depth = 1;  // (measured: 0, 1 and 2 are within noise of each other, profiles/r2w)
  const int epi_bytes = sch.nbuf * SLAB_BYTES * (addend ? 2 : 1);
  int stages = (SMEM_LIMIT - 1024 - BAR_BYTES - epi_bytes) / stage_bytes<BLOCK_N>();
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  sch.stages = stages;
  const int smem = 1024 + stages * stage_bytes<BLOCK_N>() + epi_bytes + BAR_BYTES;
  // the dynamic shared-memory opt-in of every instantiation is done per device in tapgemm_tc2_init (flm_ctx_create)
  const int pairs = sch.num_tiles < num_sms / 2 ? sch.num_tiles : num_sms / 2;
  launch_pdl(tapgemm_tc2_kernel<BLOCK_N, EPI>, dim3(2 * pairs), dim3(NUM_THREADS), (size_t)smem, stream, tm[0], tm[1], tm[2],
             tm[3], tm[4], p, sch);
  FLM_LAUNCH_CHECK();
}

template <int BLOCK_N>
void launch_cfg(const TapGemm& p, const CUtensorMap* tm, const Sched2& sch, int num_sms, cudaStream_t stream) {
  switch (p.epi) {
    case EPI_NONE: launch_one<BLOCK_N, EPI_NONE>(p, tm, sch, num_sms, stream); break;
    case EPI_GELU: launch_one<BLOCK_N, EPI_GELU>(p, tm, sch, num_sms, stream); break;
    case EPI_SILU: launch_one<BLOCK_N, EPI_SILU>(p, tm, sch, num_sms, stream); break;
    case EPI_RELU: launch_one<BLOCK_N, EPI_RELU>(p, tm, sch, num_sms, stream); break;
    case EPI_RESID: launch_one<BLOCK_N, EPI_RESID>(p, tm, sch, num_sms, stream); break;
    case EPI_GATE_RESID: launch_one<BLOCK_N, EPI_GATE_RESID>(p, tm, sch, num_sms, stream); break;
    default: throw Error(-1, "tapgemm_tc2: unsupported epilogue");
  }
}

template <int BLOCK_N, int EPI>
void set_attr_one() {
  FLM_CUDA(cudaFuncSetAttribute(tapgemm_tc2_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
}
template <int BLOCK_N>
void set_attr_all() {
  set_attr_one<BLOCK_N, EPI_NONE>(); set_attr_one<BLOCK_N, EPI_GELU>(); set_attr_one<BLOCK_N, EPI_SILU>();
  set_attr_one<BLOCK_N, EPI_RELU>(); set_attr_one<BLOCK_N, EPI_RESID>(); set_attr_one<BLOCK_N, EPI_GATE_RESID>();
}

// A tensor map is a pure function of (base pointer, dims, strides, box, swizzle, L2 promotion): the loop re-launches
// the same few dozen (pointer, shape) combinations tens of thousands of times per step, so the encoded maps are kept
// (per host thread; bounded) instead of calling cuTensorMapEncodeTiled five times per launch.
struct MapKey {
  const void* base;
  uint64_t d0, d1, d2, s0, s1;
  uint32_t b0, b1, b2, rank, promo;
  bool operator==(const MapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s0 == o.s0 && s1 == o.s1 && b0 == o.b0 &&
           b1 == o.b1 && b2 == o.b2 && rank == o.rank && promo == o.promo;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {k.d0, k.d1, k.d2, k.s0, k.s1, (uint64_t)k.b0 << 32 | k.b1, (uint64_t)k.b2 << 40 | (uint64_t)k.rank << 8 | k.promo})
      h = (h ^ v) * 0x100000001B3ull + (h >> 29);
    return (size_t)h;
  }
};
CUtensorMap cached_map(EncodeTiledFn encode, uint32_t rank, const void* base, const cuuint64_t* dims,
                       const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapL2promotion promo, const char* what) {
  static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey k{base, dims[0], dims[1], rank > 2 ? dims[2] : 1, strides[0], rank > 2 ? strides[1] : 0, box[0], box[1],
           rank > 2 ? box[2] : 1, rank, (uint32_t)promo};
  auto it = cache.find(k);
  if (it != cache.end()) return it->second;
  CUtensorMap tm;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-2, std::string("cuTensorMapEncodeTiled(") + what + ") failed: " + std::to_string((int)r));
  if (cache.size() >= 4096) cache.clear();
  cache.emplace(k, tm);
  return tm;
}

// 3-D bf16 map (cols, rows-per-sample, samples) with a (64 x 128 x 1) SWIZZLE_128B box over a row-major tensor
void encode_slab_map(EncodeTiledFn encode, CUtensorMap* tm, const void* base, int64_t ld, int N, int rows, int nb,
                     const char* what) {
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)rows, (cuuint64_t)nb};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)SLAB_COLS, (cuuint32_t)BLOCK_M, 1};
  *tm = cached_map(encode, 3, base, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what);
}

}  // namespace

// must run once per process before the first launch (and outside any stream capture)
void tapgemm_tc2_init() {
  set_attr_all<256>();
  set_attr_all<128>();
  set_attr_all<64>();
}

bool tapgemm_tc2_supported(const TapGemm& p) {
  if (!tapgemm_tc_supported(p)) return false;
  auto aligned = [](const void* q, int64_t ld) { return (reinterpret_cast<uintptr_t>(q) % 16 == 0) && ld % 8 == 0; };
  switch (p.epi) {
    case EPI_NONE: case EPI_GELU: case EPI_SILU: case EPI_RELU:
      return p.out_bf16 && p.out && aligned(p.out, p.ldc);
    case EPI_RESID:
      return p.out_bf16 && p.out && p.resid_in && aligned(p.out, p.ldc) && aligned(p.resid_in, p.ldc);
    case EPI_GATE_RESID:
      return p.hres_bf16 && p.hres && aligned(p.hres, p.ld_res) &&
             (!p.addend || (p.addend_bf16 && aligned(p.addend, p.ld_add))) &&
             (!p.out2 || (p.lnu_table && p.lnu_vecs == 4 && !p.addend && aligned(p.out2, p.ld_out2))) &&
             (!p.lnu_table || p.lnu_vecs == 3 || p.lnu_vecs == 4);
    default: return false;
  }
}

// number of (sum, sumsq) partials per row the epilogue of this problem writes: 2 per N tile
int tapgemm_tc2_rowstat_parts(const TapGemm& p, int num_sms) {
  const bool flatten = p.ntaps == 1 && p.off0 == 0;
  const int64_t rows_total = (int64_t)p.B * p.T_out;
  const int mt = flatten ? (int)((rows_total + BLOCK_M - 1) / BLOCK_M) : ((p.T_out + BLOCK_M - 1) / BLOCK_M) * p.B;
  const int mpairs = (mt + 1) / 2;
  int BN = (p.N % 256 == 0) ? 256 : (p.N % 128 == 0 ? 128 : 64);
  while (BN > 64 && p.N % (BN / 2) == 0 && (int64_t)mpairs * (p.N / BN) < num_sms / 2) BN /= 2;
  return 2 * (p.N / BN);
}

void launch_tapgemm_tc2(const TapGemm& p, void* tma_encode, int num_sms, cudaStream_t stream) {
  FLM_REQUIRE(tapgemm_tc2_supported(p), "tapgemm_tc2: unsupported problem");
  FLM_REQUIRE(tma_encode != nullptr, "tapgemm_tc2: cuTensorMapEncodeTiled entry point not resolved");
  if ((int64_t)p.B * p.T_out == 0) return;
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(tma_encode);
  Sched2 sch;
  sch.flatten = (p.ntaps == 1 && p.off0 == 0) ? 1 : 0;
  const int64_t rows_total = (int64_t)p.B * p.T_out;
  if (sch.flatten) {
    sch.tiles_m_per_b = (int)((rows_total + BLOCK_M - 1) / BLOCK_M);
    sch.mt_count = sch.tiles_m_per_b;
  } else {
    sch.tiles_m_per_b = (p.T_out + BLOCK_M - 1) / BLOCK_M;
    sch.mt_count = sch.tiles_m_per_b * p.B;
  }
  const int mpairs = (sch.mt_count + 1) / 2;
  // widest N tile that still gives every SM pair a tile (small problems are latency bound: prefer more CTAs)
  int BN = (p.N % 256 == 0) ? 256 : (p.N % 128 == 0 ? 128 : 64);
  while (BN > 64 && p.N % (BN / 2) == 0 && (int64_t)mpairs * (p.N / BN) < num_sms / 2) BN /= 2;
  sch.num_n_tiles = p.N / BN;
  sch.num_tiles = mpairs * sch.num_n_tiles;
  FLM_REQUIRE(!p.rowstat || p.rowstat_parts == 2 * sch.num_n_tiles, "tapgemm_tc2: rowstat_parts must be tapgemm_tc2_rowstat_parts(p)");
  sch.stages = 0; sch.nbuf = 0;
  CUtensorMap tm[5];
  memset(tm, 0, sizeof(tm));
  {
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BLOCK_M, 1};
    if (sch.flatten) { dims[0] = (cuuint64_t)p.K; dims[1] = (cuuint64_t)p.B * p.T_in; dims[2] = 1; }
    else { dims[0] = (cuuint64_t)p.K; dims[1] = (cuuint64_t)p.T_in; dims[2] = (cuuint64_t)p.B; }
    strides[0] = (cuuint64_t)p.lda * 2;
    strides[1] = (cuuint64_t)p.lda * 2 * dims[1];
    tm[0] = cached_map(encode, 3, p.A, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "A");
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.ntaps * p.N};
    cuuint64_t strides[1] = {(cuuint64_t)p.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(BN / 2)};
    tm[1] = cached_map(encode, 2, p.W, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "W");
  }
  const int rows = sch.flatten ? (int)rows_total : p.T_out, nb = sch.flatten ? 1 : p.B;
  if (p.epi == EPI_GATE_RESID) {
    encode_slab_map(encode, &tm[2], p.hres, p.ld_res, p.N, rows, nb, "hres");
    tm[3] = tm[2];
    if (p.addend) encode_slab_map(encode, &tm[4], p.addend, p.ld_add, p.N, rows, nb, "addend");
    else if (p.out2) encode_slab_map(encode, &tm[4], p.out2, p.ld_out2, p.N, rows, nb, "out2");
    else tm[4] = tm[2];
  } else {
    encode_slab_map(encode, &tm[3], p.out, p.ldc, p.N, rows, nb, "out");
    if (p.epi == EPI_RESID) encode_slab_map(encode, &tm[2], p.resid_in, p.ldc, p.N, rows, nb, "resid_in");
    else tm[2] = tm[3];
    tm[4] = tm[3];
  }
  if (BN == 256) launch_cfg<256>(p, tm, sch, num_sms, stream);
  else if (BN == 128) launch_cfg<128>(p, tm, sch, num_sms, stream);
  else launch_cfg<64>(p, tm, sch, num_sms, stream);
}

}  // namespace flm
