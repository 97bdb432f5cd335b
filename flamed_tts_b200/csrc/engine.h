// Host-side runtime pieces shared by the module handles: context, device buffers, weight
// lookup / packing, GEMM dispatch (fp32 FMA vs tcgen05) and the CUDA-graph cache.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/flamed_b200.h"
#include "common.cuh"
#include "kernels.h"

// kernel classes of the per-launch profiler (flm_profile_*): CUDA events around every launch
enum KClass : int {
  KC_GEMM_TC = 0,   // tcgen05 / TMA tap-GEMM
  KC_GEMM_FMA,      // fp32 FMA tap-GEMM
  KC_LN_MOD,        // LayerNorm + adaLN modulate
  KC_DWCONV,        // depthwise k=31 + GroupNorm partial statistics (fp32 mode) / the fused LN + dwconv + GN kernel
  KC_GN_FINALIZE,
  KC_GN_APPLY,
  KC_ACT1D,         // anti-aliased Snake activation
  KC_OTHER,
  KC_COUNT
};

struct flm_ctx {
  int device = 0;
  int num_sms = 0;
  void* tma_encode = nullptr;  // cuTensorMapEncodeTiled
  // profiler state
  bool prof_on = false;
  struct Rec { int kc; cudaEvent_t a, b; double flops, bytes; std::string tag; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  ~flm_ctx() {
    for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : pool) cudaEventDestroy(e);
  }
};

namespace flm {

// records a CUDA-event pair around the launches issued during its lifetime (profiling mode only,
// never while the stream is being captured into a graph)
struct ProfScope {
  flm_ctx* c; cudaStream_t s; cudaEvent_t b = nullptr;
  ProfScope(flm_ctx* ctx, int kc, cudaStream_t st, double flops, double bytes, const char* tag = nullptr)
      : c(ctx), s(st) {
    if (!c->prof_on) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    flm_ctx::Rec r;
    r.kc = kc; r.a = c->get_event(); r.b = c->get_event(); r.flops = flops; r.bytes = bytes;
    if (tag) r.tag = tag;
    cudaEventRecord(r.a, s);
    b = r.b;
    c->recs.push_back(r);
  }
  ~ProfScope() { if (b) cudaEventRecord(b, s); }
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  // grow-only; returns true if the pointer changed (callers drop cached graphs then)
  bool ensure(size_t n) {
    if (n <= bytes) return false;
    if (p) FLM_CUDA(cudaFree(p));
    p = nullptr; bytes = 0;
    FLM_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return true;
  }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

// owns every packed weight of a module
struct WeightStore {
  std::vector<void*> ptrs;
  ~WeightStore() { for (void* p : ptrs) cudaFree(p); }
  float* upload(const std::vector<float>& v) {
    void* d = nullptr;
    FLM_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)));
    ptrs.push_back(d);
    if (!v.empty()) FLM_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return static_cast<float*>(d);
  }
  bf16* upload_bf16(const std::vector<float>& v) {
    std::vector<bf16> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16_rn(v[i]);
    void* d = nullptr;
    FLM_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(bf16)));
    ptrs.push_back(d);
    if (!v.empty()) FLM_CUDA(cudaMemcpy(d, h.data(), v.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    return static_cast<bf16*>(d);
  }
};

struct WeightMap {
  std::unordered_map<std::string, const flm_tensor*> m;
  WeightMap(const flm_tensor* w, int n) {
    for (int i = 0; i < n; ++i) m[w[i].name] = &w[i];
  }
  bool has(const std::string& k) const { return m.count(k) != 0; }
  const flm_tensor& get(const std::string& k) const {
    auto it = m.find(k);
    if (it == m.end()) throw Error(FLM_ERR_WEIGHT, "missing weight: " + k);
    return *it->second;
  }
  static int64_t numel(const flm_tensor& t) {
    int64_t n = 1;
    for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
    return n;
  }
  // host copy, checking the expected shape
  std::vector<float> vec(const std::string& k, std::initializer_list<int64_t> shape) const {
    const flm_tensor& t = get(k);
    if ((size_t)t.ndim != shape.size()) throw Error(FLM_ERR_WEIGHT, "bad rank for " + k);
    int i = 0;
    for (int64_t s : shape) {
      if (s >= 0 && t.shape[i] != s)
        throw Error(FLM_ERR_WEIGHT, "bad shape for " + k + ": dim " + std::to_string(i) + " is " +
                                        std::to_string(t.shape[i]) + ", expected " + std::to_string(s));
      ++i;
    }
    return std::vector<float>(t.data, t.data + numel(t));
  }
};

// a dense / conv layer in tap-GEMM form
struct Layer {
  int K = 0, N = 0, ntaps = 1, off0 = 0, dil = 1, stride = 1;
  float alg_scale = 1.f;  // algorithmic / executed FLOPs (2/3 for the zero-padded transposed-conv taps)
  float* w32 = nullptr;  // (ntaps, N, K)
  bf16* w16 = nullptr;   // same, bf16 (only when the module runs in FLM_BF16)
  float* bias = nullptr; // (N)
};

// torch weight_norm (old style): w = g * v / ||v||, norm over all dims but 0 (facodec.py:27-32)
inline std::vector<float> fold_weight_norm(const std::vector<float>& g, const std::vector<float>& v, int64_t dim0) {
  const int64_t inner = (int64_t)v.size() / dim0;
  std::vector<float> w(v.size());
  for (int64_t i = 0; i < dim0; ++i) {
    double s = 0;
    for (int64_t j = 0; j < inner; ++j) s += (double)v[i * inner + j] * v[i * inner + j];
    const float scale = (float)((double)g[i] / std::sqrt(s));
    for (int64_t j = 0; j < inner; ++j) w[i * inner + j] = v[i * inner + j] * scale;
  }
  return w;
}

// Conv1d weight (N, K, k) -> tap-major (k, N, K)
inline std::vector<float> pack_conv(const std::vector<float>& w, int N, int K, int k) {
  std::vector<float> o((size_t)k * N * K);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < K; ++c)
      for (int t = 0; t < k; ++t) o[((size_t)t * N + n) * K + c] = w[((size_t)n * K + c) * k + t];
  return o;
}

// ConvTranspose1d weight (Cin, Cout, 2s), stride s, padding p = s/2 + s%2, output_padding s%2
// (facodec.py:246-265) as a 3-tap conv with s*Cout output columns: out row q holds the s output
// frames q*s + r;  out[q*s+r, co] = sum_{d in {1,0,-1}} sum_ci x[q-d, ci] * w[ci, co, s*d + r + p]
inline std::vector<float> pack_conv_transpose(const std::vector<float>& w, int Cin, int Cout, int s) {
  const int k = 2 * s, p = s / 2 + s % 2;
  std::vector<float> o((size_t)3 * s * Cout * Cin, 0.f);
  for (int tap = 0; tap < 3; ++tap) {
    const int d = 1 - tap;  // tap 0 reads x[q-1]
    for (int r = 0; r < s; ++r) {
      const int kk = s * d + r + p;
      if (kk < 0 || kk >= k) continue;
      for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
          o[(((size_t)tap * s + r) * Cout + co) * Cin + ci] = w[((size_t)ci * Cout + co) * k + kk];
    }
  }
  return o;
}

struct Engine {
  flm_ctx* ctx;
  int mode;
  WeightStore store;
  Engine(flm_ctx* c, int m) : ctx(c), mode(m) {}
  bool bf() const { return mode == FLM_BF16; }
  size_t esize() const { return bf() ? 2 : 4; }

  Layer make_layer(const std::vector<float>& w_packed, const std::vector<float>& bias, int K, int N, int ntaps,
                   int off0, int dil, int stride, bool want_bf16) {
    Layer l;
    l.K = K; l.N = N; l.ntaps = ntaps; l.off0 = off0; l.dil = dil; l.stride = stride;
    l.w32 = store.upload(w_packed);
    if (want_bf16) l.w16 = store.upload_bf16(w_packed);
    l.bias = store.upload(bias);
    return l;
  }

  // base problem for a layer: A (B,T_in,K) -> out (B,T_out,N); caller fills the epilogue
  TapGemm problem(const Layer& l, const void* A, int64_t lda, int B, int T_in, int T_out, void* out, int64_t ldc,
                  int out_bf16, int epi) const {
    TapGemm p;
    memset(&p, 0, sizeof(p));
    p.A = A; p.lda = lda; p.W = nullptr; p.bias = l.bias; p.out = out; p.ldc = ldc;
    p.B = B; p.T_in = T_in; p.T_out = T_out; p.K = l.K; p.N = l.N;
    p.ntaps = l.ntaps; p.off0 = l.off0; p.dil = l.dil; p.stride = l.stride;
    p.epi = epi; p.out_bf16 = out_bf16;
    return p;
  }
  // fp32 FMA kernel (A fp32) or tcgen05 kernel (A bf16), chosen by `a_bf16`
  void gemm(TapGemm p, const Layer& l, bool a_bf16, cudaStream_t s) const {
    const double M = (double)p.B * p.T_out;
    const double flops = 2.0 * M * p.N * p.K * p.ntaps * l.alg_scale;
    const double bytes = M * p.K * (a_bf16 ? 2 : 4) + (double)p.ntaps * p.N * p.K * (a_bf16 ? 2 : 4) +
                         M * p.N * ((p.epi == EPI_GATE_RESID || p.epi == EPI_EULER) ? (p.hres_bf16 ? 4 : 8) : (p.out_bf16 ? 2 : 4));
    char tag[96];
    tag[0] = 0;
    if (ctx->prof_on)
      snprintf(tag, sizeof(tag), "K%d N%d taps%d epi%d B%d T%d", p.K, p.N, p.ntaps, p.epi, p.B, p.T_out);
    ProfScope ps(ctx, a_bf16 ? KC_GEMM_TC : KC_GEMM_FMA, s, flops, bytes, tag);
    if (a_bf16) {
      if (!l.w16) throw Error(FLM_ERR_ARG, "layer has no bf16 weights");
      p.W = l.w16;
      if (tapgemm_tc2_supported(p)) launch_tapgemm_tc2(p, ctx->tma_encode, ctx->num_sms, s);
      else launch_tapgemm_tc(p, ctx->tma_encode, ctx->num_sms, s);
    } else {
      p.W = l.w32;
      launch_tapgemm_simt(p, s);
    }
  }
};

// restores the caller's current CUDA device when an API call returns (the library switches to the handle's device)
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    FLM_CUDA(cudaGetDevice(&prev));
    if (prev != device) FLM_CUDA(cudaSetDevice(device));
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// CUDA-graph cache: body(stream) is captured on a private stream (capture is not allowed on the legacy default
// stream, which is what torch hands us by default) and the instantiated graph is then launched on the caller's
// stream.  Keys are data dependent (B, P | L, nfe, temperature bits), so the cache is bounded: a key is launched
// directly the first time it is seen and only captured + instantiated when it comes back (one-shot shapes never pay
// for a capture), and at most `capacity` executables are kept, least recently used first out.
struct GraphCache {
  typedef std::tuple<int, int, int, int> Key;
  struct Entry { cudaGraphExec_t exec; unsigned long long kernels; unsigned long long stamp; };
  std::map<Key, Entry> execs;
  std::map<Key, unsigned long long> seen;  // keys launched directly so far -> last-use stamp (bounded below)
  size_t capacity = 8;
  unsigned long long clock = 0;
  cudaStream_t capture_stream = nullptr;
  ~GraphCache() {
    clear();
    if (capture_stream) cudaStreamDestroy(capture_stream);
  }
  void clear() {
    for (auto& kv : execs) cudaGraphExecDestroy(kv.second.exec);
    execs.clear();
    seen.clear();
  }
  size_t size() const { return execs.size(); }
  void run(Key key, cudaStream_t stream, const std::function<void(cudaStream_t)>& body) {
    ++clock;
    auto it = execs.find(key);
    if (it == execs.end()) {
      auto sit = seen.find(key);
      if (sit == seen.end()) {  // first sighting: direct launches, remember the key
        if (seen.size() >= 4 * capacity) {
          auto old = seen.begin();
          for (auto j = seen.begin(); j != seen.end(); ++j)
            if (j->second < old->second) old = j;
          seen.erase(old);
        }
        seen[key] = clock;
        body(stream);
        return;
      }
      seen.erase(sit);
      if (execs.size() >= capacity) {  // evict the least recently used executable
        auto old = execs.begin();
        for (auto j = execs.begin(); j != execs.end(); ++j)
          if (j->second.stamp < old->second.stamp) old = j;
        cudaGraphExecDestroy(old->second.exec);
        execs.erase(old);
      }
      if (!capture_stream) FLM_CUDA(cudaStreamCreateWithFlags(&capture_stream, cudaStreamNonBlocking));
      cudaGraph_t graph = nullptr;
      const unsigned long long before = g_launch_count.load();
      FLM_CUDA(cudaStreamBeginCapture(capture_stream, cudaStreamCaptureModeRelaxed));
      try {
        body(capture_stream);
      } catch (...) {
        cudaStreamEndCapture(capture_stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        throw;
      }
      FLM_CUDA(cudaStreamEndCapture(capture_stream, &graph));
      cudaGraphExec_t exec = nullptr;
      cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      FLM_CUDA(e);
      Entry en;
      en.exec = exec;
      en.kernels = g_launch_count.load() - before;
      en.stamp = clock;
      g_launch_count.fetch_sub(en.kernels);  // captured, not launched: counted at each replay below
      it = execs.emplace(key, en).first;
    }
    it->second.stamp = clock;
    g_launch_count.fetch_add(it->second.kernels);
    FLM_CUDA(cudaGraphLaunch(it->second.exec, stream));
  }
};

}  // namespace flm
