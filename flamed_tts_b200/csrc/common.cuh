// Shared device/host helpers for the flamed_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

typedef __nv_bfloat16 bf16;

namespace flm {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FLM_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      cudaGetLastError(); /* clear the sticky last-error so later launch checks are not poisoned */  \
      throw flm::Error(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                               std::to_string(__LINE__) + ")");                                     \
    }                                                                                               \
  } while (0)

#define FLM_REQUIRE(cond, msg)                                                  \
  do {                                                                          \
    if (!(cond)) throw flm::Error(-1, std::string("argument error: ") + (msg)); \
  } while (0)

// every kernel launcher ends with this: checks the launch and counts it (flm_launch_count)
extern std::atomic<unsigned long long> g_launch_count;
#define FLM_LAUNCH_CHECK()                                          \
  do {                                                              \
    flm::g_launch_count.fetch_add(1, std::memory_order_relaxed);    \
    FLM_CUDA(cudaGetLastError());                                   \
  } while (0)

// ---- element access in storage type T (float or bf16), arithmetic always fp32
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) {
  return *p;
}
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) {
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void stf<bf16>(bf16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// 4 consecutive elements (16 B fp32 / 8 B bf16), pointer must be aligned accordingly
template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ld4<bf16>(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void st4<bf16>(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2: two IEEE fp32 operations per instruction)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- programmatic dependent launch: a kernel launched with launch_pdl() may start while the previous kernel in the
// stream is still draining its last wave; everything it does before pdl_wait() (barrier / TMEM / descriptor set-up,
// constant-weight loads) overlaps that tail, pdl_wait() returns once the previous grid has completed and its writes
// are visible.  pdl_trigger() (at the top of a kernel) lets ITS successor do the same.  Both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FLM_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- activations (exact forms used by the reference)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float mish(float x) {
  float sp = x > 20.0f ? x : log1pf(expf(x));  // F.softplus threshold 20
  return x * tanhf(sp);
}

enum Epi : int {
  EPI_NONE = 0,
  EPI_GELU = 1,
  EPI_SILU = 2,
  EPI_RELU = 3,
  EPI_RESID = 4,       // out = resid_in + v                      (codec ResidualUnit skip)
  EPI_GATE_RESID = 5,  // hres = hres + gate[b,n] * (v + addend)  (adaLN-gated residual, fp32 stream)
  EPI_EULER = 6,       // hres = hres + alpha * v                 (x_t += dt * v)
};

// Implicit-conv GEMM problem (see include/flamed_b200.h: flm_tapgemm_test)
struct TapGemm {
  const void* A;   // (B, T_in, K) storage type TA, row stride lda elements
  const void* W;   // (ntaps, N, K) storage type TA
  const float* bias;  // (N) or null
  void* out;       // (B, T_out, N) row stride ldc elements; may be null for GATE_RESID / EULER.  EULER on the tcgen05
                   // path: if non-null, receives a bf16 copy of the updated state (operand of the next step's proj_in)
  int64_t lda, ldc;
  int B, T_in, T_out, K, N;
  int ntaps, off0, dil, stride;
  int epi;
  int out_bf16;    // 0: out is float, 1: out is bf16
  // epilogue operands
  const float* gate;     // gate[b * gate_bstride + n]
  int64_t gate_bstride;
  const void* addend;    // (B,T_out,N) storage type = A's, row stride ld_add (inner ConvNeXt residual u)
  int64_t ld_add;
  int addend_bf16;
  float* hres;           // residual stream (B,T_out,N), row stride ld_res, updated in place: fp32, or bf16 when
  int64_t ld_res;        //   hres_bf16 != 0 (EPI_GATE_RESID on the tcgen05 path only; EPI_EULER is always fp32)
  int hres_bf16;
  const void* resid_in;  // EPI_RESID: (B,T_out,N) same storage type as out, row stride ldc
  float alpha;
  // optional LayerNorm row statistics of the result (tcgen05 CTA-pair kernel, EPI_NONE / EPI_GATE_RESID): every
  // epilogue thread owns a row and writes (sum, sum of squares) of its share of the columns to
  // rowstat[row * rowstat_parts + part] (float2; part = n_tile * 2 + column half, fixed order -> deterministic);
  // the consumer (dwconv_fused) adds the rowstat_parts partials.  *rowstat_parts_out receives the number of parts
  // of this launch (2 * N / BLOCK_N) on the host.
  float* rowstat;
  int rowstat_parts;
  // EPI_GATE_RESID on the CTA-pair kernel, alternative to `addend`: the inner residual u = LayerNorm(h) * A + B is
  // recomputed in the epilogue from the residual-stream slab it holds anyway (h is unchanged since the LayerNorm), so u
  // is neither written by the depthwise kernel nor read back here.  lnu_rowconst: (rows) float2 (rstd, -mean * rstd);
  // lnu_table: (B, 3, N) = gate, gate * A, gate * (bias + B) per sample (written by launch_dwconv_tc); `bias` must be null.
  const float* lnu_rowconst;
  const float* lnu_table;
  // ... and, with lnu_table holding a 4th vector A2 (lnu_vecs == 4), a second bf16 output out2 = A2[b,n] * (updated h):
  // the column-scaled copy of the residual stream the NEXT LayerNorm's GEMM runs on (see raff_*)
  void* out2;
  int64_t ld_out2;
  int lnu_vecs;  // 3 or 4
  // EPI_SILU on the CTA-pair kernel: LayerNorm of the PREVIOUS layer applied algebraically.  A = A2 * h (out2 above),
  //   value = rstd_r * acc + nm_r * c1[b,n] + c2[b,n],   c1 = W . A2_b,  c2 = W . B2_b + bias  (`bias` must be null),
  // with (rstd_r, nm_r = -mean_r * rstd_r) of row r reduced in the epilogue from the (sum, sumsq) partials of h
  // (raff_rowstat: rows x raff_parts float2, the rowstat output of the GEMM that produced h); raff_ln_dim = channels of h.
  const float* raff_rowstat;
  int raff_parts;
  const float* raff_c1;  // (B, N)
  const float* raff_c2;  // (B, N)
  float raff_eps;
  int raff_ln_dim;
};

// value after bias -> final value; handles every epilogue except the memory side effects
__device__ __forceinline__ float epi_act(int epi, float v) {
  switch (epi) {
    case EPI_GELU: return gelu_erf(v);
    case EPI_SILU: return silu(v);
    case EPI_RELU: return fmaxf(v, 0.0f);
    default: return v;
  }
}

}  // namespace flm
