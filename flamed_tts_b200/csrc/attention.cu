// Self-attention of the prior generator's FFT decoder blocks in the bf16 mode:
//
//     o[b,s,h,:] = softmax_k( q[b,s,h,:] . k[b,k,h,:] / sqrt(d) ,  k < key_len[b] ) @ v[b,:,h,:]
//
// Reference: flamed/models/module/transformer/SubLayers.py:29-57 (MultiHeadAttention: 12 heads x 32), Modules.py:14-25
// (ScaledDotProductAttention: scores / sqrt(d_k), masked_fill(mask, -inf), softmax, @ v).  The key-padding masks of
// this model are PREFIX masks (get_mask_from_lengths), so a sample is described by its number of valid keys; key
// tiles beyond it are skipped altogether.
//
// Flash-attention dataflow on the warp-level tensor-core path (mma.sync m16n8k16, bf16 in / fp32 accumulate): d = 32
// makes the op exponent- and softmax-bound rather than MMA-bound (64 MACs per score against one ex2), so the 5th-gen
// tensor path (tcgen05 / TMEM, used by every GEMM of the library) would buy nothing here; what matters is that scores
// never leave registers.  A block of 4 warps owns 64 queries of one (sample, head); K / V tiles of 64 keys stream
// through a double-buffered cp.async ring (XOR-swizzled 64-byte rows: conflict-free ldmatrix); each warp keeps its
// 16 x 32 Q fragments, the running row max / sum and the 16 x 32 output accumulator in registers.
#include "common.cuh"
#include "kernels.h"

namespace flm {

namespace {

constexpr int DH = 32;        // head dim
constexpr int BQ = 64;        // queries per block (16 per warp)
constexpr int BK = 64;        // keys per tile
constexpr int ROW_BYTES = DH * 2;  // 64

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// byte offset of 16-byte chunk c (0..3) of row r inside a [rows][64 B] tile, XOR-swizzled
__device__ __forceinline__ uint32_t sw(int r, int c) { return (uint32_t)(r * ROW_BYTES + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;  // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// qkv: (B, S, 3, H, 32) bf16 (q | k | v of every head, as the fused QKV projection writes them); out: (B, S, H*32) bf16
__global__ void __launch_bounds__(128) attn_prefix_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ key_lens,
                                                          int S, int H, float scale_log2e, bf16* __restrict__ out) {
  __shared__ __align__(128) uint8_t qs[BQ * ROW_BYTES];
  __shared__ __align__(128) uint8_t ks[2][BK * ROW_BYTES];
  __shared__ __align__(128) uint8_t vs[2][BK * ROW_BYTES];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int klen = min(max(key_lens[b], 1), S);
  const int64_t row_stride = (int64_t)3 * H * DH;  // elements between consecutive positions
  const bf16* base = qkv + (int64_t)b * S * row_stride + h * DH;
  const int ntiles = (klen + BK - 1) / BK;

  // loader mapping: 128 threads x 16 B = 32 rows per pass, 2 passes per 64-row tile
  const int lrow = tid >> 2, lch = tid & 3;
  auto load_kv = [&](int tile, int buf) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int r = lrow + p * 32, key = tile * BK + r;
      const bool ok = key < S;
      const bf16* src = base + (int64_t)(ok ? key : 0) * row_stride + lch * 8;
      cp_async16(smem_addr(ks[buf]) + sw(r, lch), src + H * DH, ok);
      cp_async16(smem_addr(vs[buf]) + sw(r, lch), src + 2 * H * DH, ok);
    }
  };
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int r = lrow + p * 32, q = q0 + r;
    const bool ok = q < S;
    cp_async16(smem_addr(qs) + sw(r, lch), base + (int64_t)(ok ? q : 0) * row_stride + lch * 8, ok);
  }
  load_kv(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");

  // Q fragments of this warp's 16 rows (2 k-steps of 16)
  uint32_t qf[2][4];
  float o[4][4];  // 16 x 32 output accumulator: 4 n-tiles of 8 columns
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows lane/4 and lane/4 + 8

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) load_kv(t + 1, buf ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    if (t == 0) {
      // ldmatrix x4: matrices (rows 0-7, cols 0-7), (rows 8-15, cols 0-7), (rows 0-7, cols 8-15), (rows 8-15, cols 8-15)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = kk * 2 + (lane >> 4);
        ldmatrix_x4(smem_addr(qs) + sw(r, c), qf[kk]);
      }
    }
    // ---- S = Q K^T (16 x 64 per warp): 8 n-tiles of 8 keys, 2 k-steps
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[j][i] = 0.f;
      // K rows j*8..+7; the 4 matrices are the 4 16-byte chunks (d 0-7, 8-15, 16-23, 24-31) = b0,b1 of k-steps 0,1
      uint32_t kf[4];
      const int r = j * 8 + (lane & 7), c = lane >> 3;
      ldmatrix_x4(smem_addr(ks[buf]) + sw(r, c), kf);
      mma_bf16(sc[j], qf[0], kf[0], kf[1]);
      mma_bf16(sc[j], qf[1], kf[2], kf[3]);
    }
    // ---- scale, mask the tail of the key prefix, online softmax (base-2 exponent)
    const int kbase = t * BK + (lane & 3) * 2;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kbase + j * 8;
      const bool v0 = key < klen, v1 = key + 1 < klen;
      sc[j][0] = v0 ? sc[j][0] * scale_log2e : -INFINITY;
      sc[j][1] = v1 ? sc[j][1] * scale_log2e : -INFINITY;
      sc[j][2] = v0 ? sc[j][2] * scale_log2e : -INFINITY;
      sc[j][3] = v1 ? sc[j][3] * scale_log2e : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile visited has at least one valid key
    const float corr0 = fast_exp2(m0 - mn0), corr1 = fast_exp2(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];  // P as the A operand of the 4 k-steps (16 keys each) of P @ V
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = fast_exp2(sc[j][0] - mn0), p1 = fast_exp2(sc[j][1] - mn0);
      const float p2 = fast_exp2(sc[j][2] - mn1), p3 = fast_exp2(sc[j][3] - mn1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[j >> 1][(j & 1) * 2] = pack_bf16(p0, p1);      // a0 / a2: row lane/4
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);  // a1 / a3: row lane/4 + 8
    }
    l0 = l0 * corr0 + rs0;
    l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j][0] *= corr0; o[j][1] *= corr0;
      o[j][2] *= corr1; o[j][3] *= corr1;
    }
    // ---- O += P V: 4 k-steps of 16 keys, 4 n-tiles of 8 output columns
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        // V rows kk*16 .. +15, d-chunks 2*jj and 2*jj+1, transposed: matrices (keys 0-7, chunk a), (keys 8-15, chunk a),
        // (keys 0-7, chunk a+1), (keys 8-15, chunk a+1) = b0,b1 of n-tile 2*jj and b0,b1 of n-tile 2*jj+1
        uint32_t vf[4];
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = jj * 2 + (lane >> 4);
        ldmatrix_x4_trans(smem_addr(vs[buf]) + sw(r, c), vf);
        mma_bf16(o[jj * 2], pf[kk], vf[0], vf[1]);
        mma_bf16(o[jj * 2 + 1], pf[kk], vf[2], vf[3]);
      }
    }
    __syncthreads();  // every warp is done with this buffer before the next prefetch overwrites it
  }
  // ---- normalise and store
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  bf16* ob = out + (int64_t)b * S * H * DH + h * DH + (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (r0 < S) *reinterpret_cast<uint32_t*>(ob + (int64_t)r0 * H * DH + j * 8) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
    if (r1 < S) *reinterpret_cast<uint32_t*>(ob + (int64_t)r1 * H * DH + j * 8) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
  }
}

}  // namespace

void launch_attn_prefix(const bf16* qkv, const int32_t* key_lens, int B, int S, int H, int dh, bf16* out, cudaStream_t stream) {
  FLM_REQUIRE(dh == DH, "attn_prefix: head dim must be 32");
  FLM_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
              "attn_prefix: qkv must be 16-byte aligned");
  if (B == 0 || S == 0) return;
  dim3 grid((S + BQ - 1) / BQ, H, B);
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)dh);
  launch_pdl(attn_prefix_kernel, grid, dim3(128), (size_t)0, stream, qkv, key_lens, S, H, scale_log2e, out);
  FLM_LAUNCH_CHECK();
}

}  // namespace flm
