// The hot path's only collective, behind the C ABI: a variable-size gather of the generated waveforms (int16 PCM, the
// sample format the reference stores with soundfile) to one rank over NCCL / NVLink (SURVEY.md section 8 b / e).
// There is no reduction and nothing inside the sampling loops: every rank sends once, the root posts one receive per
// peer, all inside one NCCL group on the caller's stream.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 that PyTorch already has in the process), so the library
// has no link-time dependency on it; the handful of declarations below mirror nccl.h (2.2x ABI: ncclUniqueId is 128
// bytes passed by value, ncclInt8 = 0).
#include <dlfcn.h>

#include <mutex>

#include "engine.h"

using namespace flm;

extern thread_local std::string flm_g_last_error;

namespace {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
typedef int NcclResult;  // ncclSuccess == 0
constexpr int kNcclInt8 = 0;

struct NcclApi {
  NcclResult (*GetUniqueId)(NcclUniqueId*) = nullptr;
  NcclResult (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  NcclResult (*CommDestroy)(NcclComm) = nullptr;
  NcclResult (*GroupStart)() = nullptr;
  NcclResult (*GroupEnd)() = nullptr;
  NcclResult (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  NcclResult (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(NcclResult) = nullptr;
  void* lib = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy torch.distributed already loaded
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
    if (!lib) return;
    api.lib = lib;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(lib, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(dlsym(lib, "ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(dlsym(lib, "ncclRecv"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  });
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.GroupStart || !api.GroupEnd ||
      !api.Send || !api.Recv)
    throw Error(FLM_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded: multi-GPU gather unavailable");
  return api;
}

void nccl_check(NcclResult r, const char* what) {
  if (r == 0) return;
  NcclApi& a = nccl();
  throw Error(FLM_ERR_CUDA, std::string(what) + " failed: " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"));
}

__global__ void wav_to_pcm16_kernel(const float* __restrict__ x, int64_t n, int16_t* __restrict__ y) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    short4 o;
    // libsndfile float -> PCM_16 (normalised input): lrintf(x * 32767), saturated here (tanh output is inside (-1, 1))
    o.x = (short)__float2int_rn(fminf(fmaxf(v.x * 32767.0f, -32768.0f), 32767.0f));
    o.y = (short)__float2int_rn(fminf(fmaxf(v.y * 32767.0f, -32768.0f), 32767.0f));
    o.z = (short)__float2int_rn(fminf(fmaxf(v.z * 32767.0f, -32768.0f), 32767.0f));
    o.w = (short)__float2int_rn(fminf(fmaxf(v.w * 32767.0f, -32768.0f), 32767.0f));
    *reinterpret_cast<short4*>(y + i) = o;
  } else {
    for (int64_t k = i; k < n; ++k) y[k] = (short)__float2int_rn(fminf(fmaxf(x[k] * 32767.0f, -32768.0f), 32767.0f));
  }
}

}  // namespace

struct flm_comm {
  flm_ctx* ctx;
  NcclComm comm = nullptr;
  int world = 1, rank = 0;
};

#define FLM_COMM_BEGIN try {
#define FLM_COMM_END                 \
  return FLM_OK;                     \
  }                                  \
  catch (const flm::Error& e) {      \
    flm_g_last_error = e.what();     \
    return e.code;                   \
  }                                  \
  catch (const std::exception& e) {  \
    flm_g_last_error = e.what();     \
    return FLM_ERR_CUDA;             \
  }

extern "C" int flm_wav_to_pcm16(flm_ctx* ctx, const float* wav, int64_t n, int16_t* out, flm_stream stream) {
  FLM_COMM_BEGIN
  FLM_REQUIRE(ctx && wav && out && n >= 0, "bad arguments");
  FLM_REQUIRE((reinterpret_cast<uintptr_t>(wav) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
              "wav_to_pcm16: pointers must be 16-byte (wav) / 8-byte (out) aligned");
  DeviceGuard dguard(ctx->device);
  if (n == 0) return FLM_OK;
  const int64_t groups = (n + 3) / 4;
  wav_to_pcm16_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(wav, n, out);
  FLM_LAUNCH_CHECK();
  FLM_COMM_END
}

extern "C" int flm_comm_unique_id(unsigned char* out128) {
  FLM_COMM_BEGIN
  FLM_REQUIRE(out128 != nullptr, "null argument");
  NcclUniqueId id;
  nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, id.internal, 128);
  FLM_COMM_END
}

extern "C" int flm_comm_create(flm_ctx* ctx, const unsigned char* id128, int world, int rank, flm_comm** out) {
  FLM_COMM_BEGIN
  FLM_REQUIRE(ctx && id128 && out && world >= 1 && rank >= 0 && rank < world, "bad arguments");
  DeviceGuard dguard(ctx->device);
  std::unique_ptr<flm_comm> c(new flm_comm);
  c->ctx = ctx; c->world = world; c->rank = rank;
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  nccl_check(nccl().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
  *out = c.release();
  FLM_COMM_END
}

extern "C" void flm_comm_destroy(flm_comm* c) {
  if (!c) return;
  try {
    if (c->comm) nccl().CommDestroy(c->comm);
  } catch (...) {
  }
  delete c;
}

// every rank sends `n_send` int16 samples; the root receives rank r's block at recv + sum(counts[0..r)) (its own block
// is a device-to-device copy).  counts: HOST array of world entries, identical on every rank.
extern "C" int flm_gather_wav(flm_comm* c, const int16_t* send, int64_t n_send, int16_t* recv, const int64_t* counts,
                              int root, flm_stream stream) {
  FLM_COMM_BEGIN
  FLM_REQUIRE(c && counts && root >= 0 && root < c->world, "bad arguments");
  FLM_REQUIRE(counts[c->rank] == n_send, "counts[rank] must equal n_send");
  FLM_REQUIRE(n_send == 0 || send != nullptr, "null send buffer");
  FLM_REQUIRE(c->rank != root || recv != nullptr, "the root needs a receive buffer");
  DeviceGuard dguard(c->ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  NcclApi& a = nccl();
  if (c->rank == root) {
    int64_t off = 0;
    nccl_check(a.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < c->world; ++r) {
      if (r != root && counts[r] > 0)
        nccl_check(a.Recv(recv + off, (size_t)counts[r] * 2, kNcclInt8, r, c->comm, s), "ncclRecv");
      if (r == root && n_send > 0)
        FLM_CUDA(cudaMemcpyAsync(recv + off, send, (size_t)n_send * 2, cudaMemcpyDeviceToDevice, s));
      off += counts[r];
    }
    nccl_check(a.GroupEnd(), "ncclGroupEnd");
  } else if (n_send > 0) {
    nccl_check(a.GroupStart(), "ncclGroupStart");
    nccl_check(a.Send(send, (size_t)n_send * 2, kNcclInt8, root, c->comm, s), "ncclSend");
    nccl_check(a.GroupEnd(), "ncclGroupEnd");
  }
  FLM_COMM_END
}
