// Depthwise conv k=31 of the ConvNeXt blocks on the tensor cores (bf16 storage, fp32 accumulate).
//
//   y[b,t,c] = bias[c] + sum_j w[j,c] * x[b, t + j - 15, c]          (prob_generator.py:81-88, 108)
//
// Why tensor cores for a depthwise op: the fp32 FMA form costs 31 FMA per element and the SM issues 128 FMA
// per clock, i.e. at most 4.1 elements/clk/SM = 4.6 TB/s of (read+write) traffic at 100 % pipe utilisation;
// measured (profiles/r01b, r01f) both FMA kernels sit at 63 % FMA-pipe utilisation = 1.9 TB/s, 29 % of the HBM
// roofline, and no scheduling change moved that.  A 16-channel group treated as a 16x16 DIAGONAL weight matrix
// per tap is a legal tcgen05.mma (M=128 frames, N=16, K=16): 16x wasted MACs, but the tensor pipe retires one
// such MMA in 8 clocks (floor 128*N/256), so 124 MMAs cover a 128-frame x 64-channel tile in ~1000 clocks =
// 8 elements/clk/SM, twice what HBM can feed.  The op becomes memory-bound, which is where it belongs.
//
// Data path per 128 x 64 tile:
//   * one 4-D TMA copy brings the (128+30)-frame x 64-channel input window into shared memory as 8 column
//     blocks of [158 frames][8 channels] (16 B per frame): the canonical no-swizzle K-major UMMA layout with
//     consecutive frames 16 B apart, so the A operand of tap j is the SAME buffer with the descriptor start
//     address advanced by j*16 B - no per-tap reload, no im2col; frames outside [0,L) are zero-filled by the
//     TMA unit (the conv's zero padding);
//   * the 4 x 31 diagonal 16x16 bf16 weight blocks of the CTA's 64 channels live in shared memory for the
//     whole kernel (the grid is a multiple of the number of channel blocks);
//   * accumulators (4 groups x 16 columns) are double-buffered in TMEM; 4 epilogue warps read them
//     thread-per-frame, reduce the per-(32-frame chunk, channel) GroupNorm partials (sum, sum of squares about
//     the bias) with a shuffle butterfly, add the bias and hand the bf16 tile to a bulk tensor store through a
//     SWIZZLE_128B staging buffer (frames beyond L are clipped by the TMA unit).
// The chunk partials have the format of the FMA kernels (kernels_norm.cu) and are merged by dw_merge_kernel.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {

constexpr int KW = 31;
constexpr int PAD = KW / 2;
constexpr int TILE_T = 128;                // frames per tile = MMA M
constexpr int TILE_C = 64;                 // channels per tile
constexpr int ROWS = TILE_T + KW - 1;      // 158 input frames per tile
constexpr int KCH = TILE_C / 8;            // 8-channel (16 B) column blocks per tile
constexpr int A_STAGE_BYTES = KCH * ROWS * 16;  // 20224
constexpr int STAGES = 4;
constexpr int GROUPS = TILE_C / 16;        // 16-channel MMA groups per tile
constexpr int WBLK_BYTES = 16 * 16 * 2;    // one diagonal 16x16 bf16 block
constexpr int W_BYTES = GROUPS * KW * WBLK_BYTES;  // 63488
constexpr int OUT_BYTES = TILE_T * TILE_C * 2;     // 16384, SWIZZLE_128B slab
constexpr int NUM_THREADS = 6 * 32;        // 0 producer, 1 MMA, 2..5 epilogue
constexpr int SMEM_BYTES = 1024 + 2 * OUT_BYTES + STAGES * A_STAGE_BYTES + W_BYTES + TILE_C * 4 + 256;

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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// no-swizzle K-major matrix descriptor: rows of an 8-row core matrix 16 B apart, 8-row groups SBO apart, the two
// 8-element K halves LBO apart (cute::UMMA::make_umma_desc<Major::K>, LayoutType::INTERLEAVE)
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
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

// butterfly transpose-reduce over the 32 lanes: on entry every lane holds 64 values (its frame, 64 channels); on
// exit v[0], v[1] are the sums over the warp's 32 frames of channels 2*lane and 2*lane+1
__device__ __forceinline__ void butterfly64(float (&v)[64], int lane) {
#pragma unroll
  for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
dwconv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const DwConv p,
                 int ncblk, int nchunk_t, int ntiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_out = smem;                                   // 2 x 16 KB, 1024-aligned (SWIZZLE_128B)
  uint8_t* smem_a = smem_out + 2 * OUT_BYTES;                 // STAGES x 20224 (multiple of 128)
  uint8_t* smem_w = smem_a + STAGES * A_STAGE_BYTES;          // diagonal weight blocks
  float* smem_bias = reinterpret_cast<float*>(smem_w + W_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bias + TILE_C);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = blockIdx.x % ncblk;  // gridDim.x % ncblk == 0: fixed for the whole kernel
  const int c0 = cblk * TILE_C;
  const int tstride = gridDim.x / ncblk;
  const int tfirst = blockIdx.x / ncblk;

  // diagonal weight blocks: block (g, j) at (g*KW + j)*512 B; element (n, k) of a block at
  // (k/8)*256 + (n/8)*128 + (n%8)*16 + (k%8)*2; only n == k is non-zero
  for (int i = threadIdx.x; i < W_BYTES / 16; i += NUM_THREADS) reinterpret_cast<uint4*>(smem_w)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < TILE_C) smem_bias[threadIdx.x] = p.bias[c0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.x; i < GROUPS * KW * 16; i += NUM_THREADS) {
    const int n = i & 15, blk = i >> 4;  // blk = g*KW + j
    const int g = blk / KW, j = blk % KW;
    const float w = p.w[(int64_t)j * p.C + c0 + g * 16 + n];
    bf16* dst = reinterpret_cast<bf16*>(smem_w + blk * WBLK_BYTES + (n >> 3) * 256 + (n >> 3) * 128 + (n & 7) * 16 + (n & 7) * 2);
    *dst = __float2bfloat16(w);
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // weight blocks (generic writes) -> tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tfirst; tile < ntiles; tile += tstride) {
        const int b = tile / nchunk_t, ck = tile % nchunk_t;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                "r"(smem_u32(smem_a + stage * A_STAGE_BYTES)),
            "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(smem_u32(&full_bar[stage])), "r"(0), "r"(ck * TILE_T - PAD),
            "r"(cblk * KCH), "r"(b)
            : "memory");
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(TILE_T, 16);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = tfirst; tile < ntiles; tile += tstride, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_base = smem_u32(smem_a + stage * A_STAGE_BYTES);
        const uint32_t w_base = smem_u32(smem_w);
#pragma unroll 1
        for (int g = 0; g < GROUPS; ++g) {
          // K = 16 channels of group g = column blocks 2g, 2g+1 (LBO = one column block); tap j = +j frames = +j*16 B
          uint64_t da = make_desc_noswz(a_base + (2 * g) * (ROWS * 16), ROWS * 16, 128);
          uint64_t db = make_desc_noswz(w_base + g * KW * WBLK_BYTES, 256, 128);
          const uint32_t tmem_d = tmem_base + (uint32_t)(as * TILE_C + g * 16);
#pragma unroll
          for (int j = 0; j < KW; ++j) {
            umma_f16(tmem_d, da, db, idesc, j > 0 ? 1u : 0u);
            da += 1;                       // 16 B
            db += WBLK_BYTES >> 4;         // next diagonal block
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[as]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== epilogue: thread = frame =====================
    const int quarter = warp & 3;  // TMEM lane quarter
    const int rrow = quarter * 32 + lane;
    int it = 0;
    for (int tile = tfirst; tile < ntiles; tile += tstride, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int b = tile / nchunk_t, ck = tile % nchunk_t;
      const int t0 = ck * TILE_T;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      float v[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TILE_C);
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);  // accumulator stage drained: the next tile's MMAs may start
      // ---- output tile (bias added), staged for the bulk store
      uint8_t* obuf = smem_out + (it & 1) * OUT_BYTES;
      if (it >= 2) {  // the store issued two tiles ago from this buffer must have read it
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      {
        uint8_t* orow = obuf + rrow * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = c * 8 + 2 * j;
            __nv_bfloat162 o = __floats2bfloat162_rn(v[ch] + smem_bias[ch], v[ch + 1] + smem_bias[ch + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&o);
          }
          *reinterpret_cast<uint4*>(orow + ((c ^ (rrow & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (warp == 2 && lane == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tmY)),
                     "r"(smem_u32(obuf)), "r"(c0), "r"(t0), "r"(b)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      // ---- GroupNorm partials of this warp's 32-frame chunk: sums of (y - bias) and (y - bias)^2 over valid frames
      const int tw = t0 + quarter * 32;  // first frame of the warp's chunk
      const int nvalid = min(32, p.L - tw);
      if (nvalid > 0) {                  // warp-uniform
        const bool valid = lane < nvalid;
        float q[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          v[i] = valid ? v[i] : 0.f;
          q[i] = v[i] * v[i];
        }
        butterfly64(v, lane);
        butterfly64(q, lane);
        const float inv = 1.0f / (float)nvalid;
        const int ch = 2 * lane;
        const float m0 = v[0] * inv, m1 = v[1] * inv;
        const float q0 = fmaxf(q[0] - v[0] * m0, 0.f), q1 = fmaxf(q[1] - v[1] * m1, 0.f);
        const int chunk32 = tw / DW_TT;
        float* part = p.part + (((int64_t)b * ((p.L + DW_TT - 1) / DW_TT) + chunk32) * p.C + c0 + ch) * 2;
        *reinterpret_cast<float4*>(part) = make_float4(m0 + smem_bias[ch], q0, m1 + smem_bias[ch + 1], q1);
      }
    }
    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

void dwconv_tc_init() {
  FLM_CUDA(cudaFuncSetAttribute(dwconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
}

bool dwconv_tc_supported(const DwConv& p) {
  return p.io_bf16 && p.KW == KW && p.C % TILE_C == 0 && p.tma_encode != nullptr &&
         (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0;
}

// writes y and the per-(sample, 32-frame chunk, channel) partials; the caller merges them (dw_merge)
void launch_dwconv_tc(const DwConv& p, int num_sms, cudaStream_t stream) {
  FLM_REQUIRE(dwconv_tc_supported(p), "dwconv_tc: unsupported problem");
  if (p.B == 0 || p.L == 0) return;
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(p.tma_encode);
  CUtensorMap tmX, tmY;
  {
    // (8 channels, frames, 8-channel column blocks, samples): box (8, 158, 8, 1) lands as [column block][frame][8 ch]
    cuuint64_t dims[4] = {8, (cuuint64_t)p.L, (cuuint64_t)(p.C / 8), (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)p.C * 2, 16, (cuuint64_t)p.C * 2 * (cuuint64_t)p.L};
    cuuint32_t box[4] = {8, (cuuint32_t)ROWS, (cuuint32_t)KCH, 1}, estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-2, "cuTensorMapEncodeTiled(dwconv_tc x) failed: " + std::to_string((int)r));
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.C, (cuuint64_t)p.L, (cuuint64_t)p.B};
    cuuint64_t strides[2] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.C * 2 * (cuuint64_t)p.L};
    cuuint32_t box[3] = {(cuuint32_t)TILE_C, (cuuint32_t)TILE_T, 1}, estr[3] = {1, 1, 1};
    CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p.y, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-2, "cuTensorMapEncodeTiled(dwconv_tc y) failed: " + std::to_string((int)r));
  }
  const int ncblk = p.C / TILE_C;
  const int nchunk_t = (p.L + TILE_T - 1) / TILE_T;
  const int ntiles = p.B * nchunk_t;  // per channel block
  int per_cblk = num_sms / ncblk;
  if (per_cblk < 1) per_cblk = 1;
  if (per_cblk > ntiles) per_cblk = ntiles;
  dwconv_tc_kernel<<<per_cblk * ncblk, NUM_THREADS, SMEM_BYTES, stream>>>(tmX, tmY, p, ncblk, nchunk_t, ntiles);
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
