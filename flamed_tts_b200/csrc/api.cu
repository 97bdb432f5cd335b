// C ABI of libflamed_b200.so (see include/flamed_b200.h): context, weight packing and the host
// orchestration of the four hot-path modules.  All compute is in the CUDA kernels of this
// directory; there is no CPU fallback.
#include <cuda.h>

#include "engine.h"

using namespace flm;

thread_local std::string flm_g_last_error;  // shared with comm.cu
#define g_last_error flm_g_last_error
namespace flm { std::atomic<unsigned long long> g_launch_count{0}; }

#define FLM_API_BEGIN try {
#define FLM_API_END                          \
  return FLM_OK;                             \
  }                                          \
  catch (const flm::Error& e) {              \
    g_last_error = e.what();                 \
    return e.code;                           \
  }                                          \
  catch (const std::exception& e) {          \
    g_last_error = e.what();                 \
    return FLM_ERR_CUDA;                     \
  }

extern "C" const char* flm_last_error(void) { return g_last_error.c_str(); }
extern "C" int flm_version(void) { return 100; }
extern "C" unsigned long long flm_launch_count(void) { return flm::g_launch_count.load(); }

// ===================================================================================== context
extern "C" int flm_ctx_create(int device, flm_ctx** out) {
  FLM_API_BEGIN
  FLM_REQUIRE(out != nullptr, "out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(FLM_ERR_CUDA, "no CUDA device: flamed_b200 has no CPU fallback (needs an sm_100 GPU)");
  FLM_REQUIRE(device >= 0 && device < count, "bad device index");
  DeviceGuard dguard(device);
  cudaDeviceProp prop;
  FLM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    throw Error(FLM_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                  ", the kernels are built for sm_100a only");
  std::unique_ptr<flm_ctx> c(new flm_ctx);
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  cudaDriverEntryPointQueryResult qres;
  FLM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &c->tma_encode, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !c->tma_encode)
    throw Error(FLM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  tapgemm_tc_init();
  tapgemm_tc2_init();
  dwconv_fused_init();
  dwconv_tc_init();
  kernels_norm_init();
  *out = c.release();
  FLM_API_END
}
extern "C" void flm_ctx_destroy(flm_ctx* ctx) { delete ctx; }

static inline cudaStream_t S(flm_stream s) { return static_cast<cudaStream_t>(s); }

static std::vector<float> replicate(const std::vector<float>& v, int times) {
  std::vector<float> o;
  o.reserve(v.size() * times);
  for (int i = 0; i < times; ++i) o.insert(o.end(), v.begin(), v.end());
  return o;
}

// ===================================================================================== durgen
namespace {
struct DurNet {
  Layer encproj, te1, te3, c1, c2;
  float *w0, *ln1w, *ln1b, *ln2w, *ln2b, *wl, *bl;
  int D, F;
  // per-call buffers
  DevBuf encp, semb, teh, temb, a0, r1, l1, r2, xt;
};
}  // namespace

struct flm_durgen : Engine {
  DurNet net[2];
  DevBuf enc_s, noise_s[2], mask_s, ts_s, out_s[2], seed_s;
  GraphCache graphs;
  flm_durgen(flm_ctx* c) : Engine(c, FLM_F32) {}

  void load_net(const WeightMap& wm, const std::string& p, DurNet& n) {
    const flm_tensor& pw = wm.get(p + ".proj.weight");
    const int D = (int)pw.shape[0];
    FLM_REQUIRE(pw.ndim == 2 && pw.shape[1] == D + 1, "proj.weight must be (D, D+1)");
    std::vector<float> w = wm.vec(p + ".proj.weight", {D, D + 1});
    std::vector<float> w0(D), wenc((size_t)D * D);
    for (int c = 0; c < D; ++c) {
      w0[c] = w[(size_t)c * (D + 1)];
      for (int k = 0; k < D; ++k) wenc[(size_t)c * D + k] = w[(size_t)c * (D + 1) + 1 + k];
    }
    n.D = D;
    n.w0 = store.upload(w0);
    n.encproj = make_layer(wenc, wm.vec(p + ".proj.bias", {D}), D, D, 1, 0, 1, 1, false);
    const int TS = (int)wm.get(p + ".time_emb.time_emb.1.weight").shape[0];
    n.te1 = make_layer(wm.vec(p + ".time_emb.time_emb.1.weight", {TS, D}), wm.vec(p + ".time_emb.time_emb.1.bias", {TS}),
                       D, TS, 1, 0, 1, 1, false);
    n.te3 = make_layer(wm.vec(p + ".time_emb.time_emb.3.weight", {D, TS}), wm.vec(p + ".time_emb.time_emb.3.bias", {D}),
                       TS, D, 1, 0, 1, 1, false);
    const flm_tensor& cw = wm.get(p + ".conv_layer.conv1d_1.conv.weight");
    const int F = (int)cw.shape[0], k = (int)cw.shape[2];
    FLM_REQUIRE(k == 3, "duration generator kernel_size must be 3");
    n.F = F;
    n.c1 = make_layer(pack_conv(wm.vec(p + ".conv_layer.conv1d_1.conv.weight", {F, D, k}), F, D, k),
                      wm.vec(p + ".conv_layer.conv1d_1.conv.bias", {F}), D, F, k, -(k - 1) / 2, 1, 1, false);
    n.c2 = make_layer(pack_conv(wm.vec(p + ".conv_layer.conv1d_2.conv.weight", {F, F, k}), F, F, k),
                      wm.vec(p + ".conv_layer.conv1d_2.conv.bias", {F}), F, F, k, -1, 1, 1, false);
    n.ln1w = store.upload(wm.vec(p + ".conv_layer.layer_norm_1.weight", {F}));
    n.ln1b = store.upload(wm.vec(p + ".conv_layer.layer_norm_1.bias", {F}));
    n.ln2w = store.upload(wm.vec(p + ".conv_layer.layer_norm_2.weight", {F}));
    n.ln2b = store.upload(wm.vec(p + ".conv_layer.layer_norm_2.bias", {F}));
    n.wl = store.upload(wm.vec(p + ".linear_layer.weight", {1, F}));
    n.bl = store.upload(wm.vec(p + ".linear_layer.bias", {1}));
  }

  // one velocity evaluation of generator g at time-table row i, accumulated into xt: xt += dt * v
  void velocity(int g, int B, int P, int i, float dt, cudaStream_t s) {
    trunk(g, B, P, i, s);
    DurNet& n = net[g];
    launch_durgen_head(n.r2.as<float>(), (int64_t)B * P, n.F, n.ln2w, n.ln2b, n.wl, n.bl, mask_s.as<uint8_t>(), dt,
                       n.xt.as<float>(), s);
  }
  // everything of ProbabilisticModule.forward up to the second conv (pva.py:222-233): r2 = conv2(LN(relu(conv1(...))))
  void trunk(int g, int B, int P, int i, cudaStream_t s) {
    const int64_t rows = (int64_t)B * P;
    DurNet& n = net[g];
    launch_durgen_input(n.encp.as<float>(), n.xt.as<float>(), n.w0, n.temb.as<float>() + (int64_t)i * n.D, rows, n.D,
                        n.a0.as<float>(), s);
    gemm(problem(n.c1, n.a0.p, n.D, B, P, P, n.r1.p, n.F, 0, EPI_NONE), n.c1, false, s);
    LnMod ln;
    memset(&ln, 0, sizeof(ln));
    ln.x = n.r1.p; ln.ldx = n.F; ln.y = n.l1.p; ln.ldy = n.F; ln.y_bf16 = 0;
    ln.w = n.ln1w; ln.b = n.ln1b; ln.eps = 1e-5f; ln.rows = rows; ln.rows_per_batch = P; ln.C = n.F;
    ln.relu_in = 1;
    launch_ln_mod(ln, s);
    gemm(problem(n.c2, n.l1.p, n.F, B, P, P, n.r2.p, n.F, 0, EPI_NONE), n.c2, false, s);
  }
  void prologue(int g, int B, int P, int nfe, cudaStream_t s) {
    DurNet& n = net[g];
    gemm(problem(n.encproj, enc_s.p, n.D, B, P, P, n.encp.p, n.D, 0, EPI_NONE), n.encproj, false, s);
    launch_sinusoidal_pos_emb(ts_s.as<float>(), nfe, n.D, n.semb.as<float>(), s);
    gemm(problem(n.te1, n.semb.p, n.D, 1, nfe, nfe, n.teh.p, n.te1.N, 0, EPI_SILU), n.te1, false, s);
    gemm(problem(n.te3, n.teh.p, n.te1.N, 1, nfe, nfe, n.temb.p, n.D, 0, EPI_NONE), n.te3, false, s);
  }
  bool ensure(int B, int P, int nfe) {
    const int64_t rows = (int64_t)B * P;
    bool moved = false;
    moved |= enc_s.ensure(rows * net[0].D * 4);
    moved |= mask_s.ensure(rows);
    moved |= ts_s.ensure((nfe + 1) * 4);
    moved |= seed_s.ensure(8);
    for (int g = 0; g < 2; ++g) {
      DurNet& n = net[g];
      moved |= noise_s[g].ensure(rows * 4);
      moved |= out_s[g].ensure(rows * 4);
      moved |= n.encp.ensure(rows * n.D * 4);
      moved |= n.semb.ensure((size_t)nfe * n.D * 4);
      moved |= n.teh.ensure((size_t)nfe * n.te1.N * 4);
      moved |= n.temb.ensure((size_t)nfe * n.D * 4);
      moved |= n.a0.ensure(rows * n.D * 4);
      moved |= n.r1.ensure(rows * n.F * 4);
      moved |= n.l1.ensure(rows * n.F * 4);
      moved |= n.r2.ensure(rows * n.F * 4);
      moved |= n.xt.ensure(rows * 4);
    }
    if (moved) graphs.clear();
    return moved;
  }

  void body(int B, int P, int nfe, float temperature, bool philox, cudaStream_t s) {
    const int64_t rows = (int64_t)B * P;
    const float dt = (float)(1.0 / nfe);
    for (int g = 0; g < 2; ++g) {
      DurNet& n = net[g];
      if (philox) launch_philox_normal(seed_s.as<uint64_t>(), (uint32_t)g, temperature, nullptr, rows, n.xt.as<float>(), s);
      else launch_scale(noise_s[g].as<float>(), temperature, rows, n.xt.as<float>(), s);
      prologue(g, B, P, nfe, s);
    }
    for (int i = 0; i < nfe; ++i)
      for (int g = 0; g < 2; ++g) velocity(g, B, P, i, dt, s);
    for (int g = 0; g < 2; ++g) launch_duration_round(net[g].xt.as<float>(), rows, out_s[g].as<float>(), s);
  }
};

extern "C" int flm_durgen_load(flm_ctx* ctx, const flm_tensor* weights, int n, flm_durgen** out) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && weights && out, "null argument");
  DeviceGuard dguard(ctx->device);
  WeightMap wm(weights, n);
  std::unique_ptr<flm_durgen> h(new flm_durgen(ctx));
  h->load_net(wm, "duration_generator", h->net[0]);
  h->load_net(wm, "sil_generator", h->net[1]);
  *out = h.release();
  FLM_API_END
}
extern "C" void flm_durgen_destroy(flm_durgen* h) { delete h; }

extern "C" int flm_durgen_sample(flm_durgen* h, const float* enc, const float* noise_dur, const float* noise_sil,
                                 uint64_t seed, const uint8_t* src_mask, const float* ts_host, int nfe, float temperature,
                                 int B, int P, float* out_phone, float* out_sil, float* out_dur_t, float* out_sil_t,
                                 flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && enc && src_mask && ts_host && out_phone && out_sil, "null argument");
  FLM_REQUIRE((noise_dur == nullptr) == (noise_sil == nullptr), "noise_dur and noise_sil must both be given or both be NULL");
  FLM_REQUIRE(nfe >= 1 && B >= 0 && P >= 1, "bad sizes");
  DeviceGuard dguard(h->ctx->device);
  if (B == 0) return FLM_OK;
  cudaStream_t s = S(stream);
  const int64_t rows = (int64_t)B * P;
  const bool philox = noise_dur == nullptr;
  h->ensure(B, P, nfe);
  FLM_CUDA(cudaMemcpyAsync(h->enc_s.p, enc, rows * h->net[0].D * 4, cudaMemcpyDeviceToDevice, s));
  if (philox) {
    FLM_CUDA(cudaMemcpyAsync(h->seed_s.p, &seed, 8, cudaMemcpyHostToDevice, s));
  } else {
    FLM_CUDA(cudaMemcpyAsync(h->noise_s[0].p, noise_dur, rows * 4, cudaMemcpyDeviceToDevice, s));
    FLM_CUDA(cudaMemcpyAsync(h->noise_s[1].p, noise_sil, rows * 4, cudaMemcpyDeviceToDevice, s));
  }
  FLM_CUDA(cudaMemcpyAsync(h->mask_s.p, src_mask, rows, cudaMemcpyDeviceToDevice, s));
  FLM_CUDA(cudaMemcpyAsync(h->ts_s.p, ts_host, (nfe + 1) * 4, cudaMemcpyHostToDevice, s));
  int tbits;
  memcpy(&tbits, &temperature, 4);
  h->graphs.run(std::make_tuple(B, P, philox ? -nfe : nfe, tbits), s,
                [&](cudaStream_t cs) { h->body(B, P, nfe, temperature, philox, cs); });
  FLM_CUDA(cudaMemcpyAsync(out_phone, h->out_s[0].p, rows * 4, cudaMemcpyDeviceToDevice, s));
  FLM_CUDA(cudaMemcpyAsync(out_sil, h->out_s[1].p, rows * 4, cudaMemcpyDeviceToDevice, s));
  if (out_dur_t) FLM_CUDA(cudaMemcpyAsync(out_dur_t, h->net[0].xt.p, rows * 4, cudaMemcpyDeviceToDevice, s));
  if (out_sil_t) FLM_CUDA(cudaMemcpyAsync(out_sil_t, h->net[1].xt.p, rows * 4, cudaMemcpyDeviceToDevice, s));
  FLM_API_END
}

// one vector-field evaluation v = ProbabilisticModule.forward(x, enc, t, mask) (pva.py:221-238) of generator
// `which` (0 duration, 1 silence): the same kernels as one step of the loop, with x as the state and dt = 1
extern "C" int flm_durgen_forward(flm_durgen* h, int which, const float* x, const float* enc, float t,
                                  const uint8_t* src_mask, int B, int P, float* out_v, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && x && enc && out_v, "null argument");
  FLM_REQUIRE(which == 0 || which == 1, "which must be 0 (duration) or 1 (silence)");
  FLM_REQUIRE(B >= 0 && P >= 1, "bad sizes");
  DeviceGuard dguard(h->ctx->device);
  if (B == 0) return FLM_OK;
  cudaStream_t s = S(stream);
  const int64_t rows = (int64_t)B * P;
  h->ensure(B, P, 1);
  DurNet& n = h->net[which];
  const float ts[2] = {t, 1.0f};
  FLM_CUDA(cudaMemcpyAsync(h->enc_s.p, enc, rows * n.D * 4, cudaMemcpyDeviceToDevice, s));
  if (src_mask) FLM_CUDA(cudaMemcpyAsync(h->mask_s.p, src_mask, rows, cudaMemcpyDeviceToDevice, s));
  else FLM_CUDA(cudaMemsetAsync(h->mask_s.p, 0, rows, s));
  FLM_CUDA(cudaMemcpyAsync(h->ts_s.p, ts, 8, cudaMemcpyHostToDevice, s));
  FLM_CUDA(cudaMemcpyAsync(n.xt.p, x, rows * 4, cudaMemcpyDeviceToDevice, s));
  h->prologue(which, B, P, 1, s);
  h->trunk(which, B, P, 0, s);
  // the head accumulates xt += dt * v: on a zero state with dt = 1 that is exactly v
  FLM_CUDA(cudaMemsetAsync(n.xt.p, 0, rows * 4, s));
  launch_durgen_head(n.r2.as<float>(), rows, n.F, n.ln2w, n.ln2b, n.wl, n.bl, h->mask_s.as<uint8_t>(), 1.0f,
                     n.xt.as<float>(), s);
  FLM_CUDA(cudaMemcpyAsync(out_v, n.xt.p, rows * 4, cudaMemcpyDeviceToDevice, s));
  FLM_API_END
}

// ===================================================================================== length regulator
extern "C" int flm_lr_plan(flm_ctx* ctx, const float* phone_dur, const float* sil_dur, const int64_t* src_lens, int B,
                           int P, int32_t* out_cumsum, int64_t* out_tgt_len, int64_t* out_tmax_host, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && phone_dur && sil_dur && src_lens && out_cumsum && out_tgt_len, "null argument");
  DeviceGuard dguard(ctx->device);
  if (out_tmax_host) *out_tmax_host = 0;
  if (B == 0) return FLM_OK;
  launch_lr_plan(phone_dur, sil_dur, src_lens, B, P, out_cumsum, out_tgt_len, S(stream));
  if (!out_tmax_host) return FLM_OK;  // asynchronous form: the caller reads tgt_len later (batched re-bucketing)
  std::vector<int64_t> tl(B);
  FLM_CUDA(cudaMemcpyAsync(tl.data(), out_tgt_len, (size_t)B * 8, cudaMemcpyDeviceToHost, S(stream)));
  FLM_CUDA(cudaStreamSynchronize(S(stream)));  // the one documented sync (reference: pva.py:158 .tolist())
  int64_t mx = 0;
  for (int64_t v : tl) mx = std::max(mx, v);
  *out_tmax_host = mx;
  FLM_API_END
}

extern "C" int flm_lr_expand(flm_ctx* ctx, const float* x, const int32_t* cumsum, int B, int P, int H, int Tmax,
                             float* out, int32_t* out_index, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && x && cumsum && out, "null argument");
  DeviceGuard dguard(ctx->device);
  launch_lr_expand(x, cumsum, B, P, H, Tmax, out, out_index, S(stream));
  FLM_API_END
}

extern "C" int flm_lr_expand_gather(flm_ctx* ctx, const float* const* x_rows, const int32_t* const* cumsums,
                                    const int32_t* P_per_sample, int B, int H, int Tmax, float* out, int32_t* out_index,
                                    flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && x_rows && cumsums && P_per_sample && out, "null argument");
  DeviceGuard dguard(ctx->device);
  launch_lr_expand_gather(x_rows, cumsums, P_per_sample, B, H, Tmax, out, out_index, S(stream));
  FLM_API_END
}

// ===================================================================================== denoiser
namespace {
struct ConvNeXtW {
  float *dw_w, *dw_b, *gn_w, *gn_b, *dw_wsum;
  Layer conv2, conv3;
};
struct ResBlockW {
  float *lnc_w, *lnc_b, *lnm_w, *lnm_b;
  ConvNeXtW cn;
  Layer mlp0, mlp2;
};
}  // namespace

struct flm_denoiser : Engine {
  flm_prob_cfg cfg;
  int H, D, ada_n;
  bool h16 = false;  // bf16 residual stream inside a step (FLM_BF16 mode); fp32 stream in the FLM_F32 parity mode
  size_t hsize() const { return h16 ? 2 : 4; }
  // denoiser weights
  Layer time0, time2, cond_embed, proj_in, ada_all, conv_out;
  std::vector<ResBlockW> blocks;
  ConvNeXtW fin;
  // cond down-sampler weights
  float* qemb;
  struct Stage { Layer conv_a, conv_b; float *gna_w, *gna_b, *gnb_w, *gnb_b; };
  std::vector<Stage> stages;
  Layer proj_out;
  // buffers
  DevBuf seed_s, rowstat, rowconst, lnab, lnu, s1buf, s2buf, c1tab, c2tab;
  DevBuf cond_s, spk_s, noise_s, ts_s, x, xb, h, bufU, bufD, bufG, bufA, part, gsc, gof, gctr, ada, sbuf, temb, tfreq, teh,
      cvec, vout;
  DevBuf c_prior, c_mask, c_xq, c_xm, c_h, c_part, c_sc, c_of, c_out;
  GraphCache graphs;
  flm_denoiser(flm_ctx* c, int mode) : Engine(c, mode) {}

  ConvNeXtW load_convnext(const WeightMap& wm, const std::string& p) {
    ConvNeXtW c;
    const int k = cfg.kernel_size;
    std::vector<float> w = wm.vec(p + ".conv_1.weight", {H, 1, k});
    std::vector<float> wt((size_t)k * H);
    for (int ch = 0; ch < H; ++ch)
      for (int t = 0; t < k; ++t) wt[(size_t)t * H + ch] = w[(size_t)ch * k + t];
    c.dw_w = store.upload(wt);
    {
      std::vector<float> ws(H, 0.f);  // response of the conv to a constant input (statistics pivot of the fused kernel)
      for (int ch = 0; ch < H; ++ch) {
        double a = 0;
        for (int t = 0; t < k; ++t) a += w[(size_t)ch * k + t];
        ws[ch] = (float)a;
      }
      c.dw_wsum = store.upload(ws);
    }
    c.dw_b = store.upload(wm.vec(p + ".conv_1.bias", {H}));
    c.gn_w = store.upload(wm.vec(p + ".ln_1.weight", {H}));
    c.gn_b = store.upload(wm.vec(p + ".ln_1.bias", {H}));
    c.conv2 = make_layer(wm.vec(p + ".conv_2.weight", {H, H, 1}), wm.vec(p + ".conv_2.bias", {H}), H, H, 1, 0, 1, 1, bf());
    c.conv3 = make_layer(wm.vec(p + ".conv_3.weight", {H, H, 1}), wm.vec(p + ".conv_3.bias", {H}), H, H, 1, 0, 1, 1, bf());
    return c;
  }

  void load(const WeightMap& wm) {
    H = cfg.hidden_dim; D = cfg.target_dim;
    const std::string dn = "denoiser";
    time0 = make_layer(wm.vec(dn + ".time_embed.mlp.0.weight", {H, 256}), wm.vec(dn + ".time_embed.mlp.0.bias", {H}), 256, H, 1, 0, 1, 1, false);
    time2 = make_layer(wm.vec(dn + ".time_embed.mlp.2.weight", {H, H}), wm.vec(dn + ".time_embed.mlp.2.bias", {H}), H, H, 1, 0, 1, 1, false);
    cond_embed = make_layer(wm.vec(dn + ".cond_embed.weight", {H, cfg.spk_dim}), wm.vec(dn + ".cond_embed.bias", {H}), cfg.spk_dim, H, 1, 0, 1, 1, false);
    proj_in = make_layer(wm.vec(dn + ".proj_in.weight", {H, D}), wm.vec(dn + ".proj_in.bias", {H}), D, H, 1, 0, 1, 1, bf());
    std::vector<float> ada_w, ada_b;
    for (int i = 0; i < cfg.n_layers; ++i) {
      const std::string p = dn + ".res_blocks." + std::to_string(i);
      ResBlockW b;
      b.lnc_w = store.upload(wm.vec(p + ".ln_conv.weight", {H}));
      b.lnc_b = store.upload(wm.vec(p + ".ln_conv.bias", {H}));
      b.lnm_w = store.upload(wm.vec(p + ".ln_mlp.weight", {H}));
      b.lnm_b = store.upload(wm.vec(p + ".ln_mlp.bias", {H}));
      b.cn = load_convnext(wm, p + ".conv_in");
      b.mlp0 = make_layer(wm.vec(p + ".mlp.0.weight", {H, H}), wm.vec(p + ".mlp.0.bias", {H}), H, H, 1, 0, 1, 1, bf());
      b.mlp2 = make_layer(wm.vec(p + ".mlp.2.weight", {H, H}), wm.vec(p + ".mlp.2.bias", {H}), H, H, 1, 0, 1, 1, bf());
      blocks.push_back(b);
      std::vector<float> w = wm.vec(p + ".adaLN_modulation.1.weight", {6 * H, H});
      std::vector<float> bb = wm.vec(p + ".adaLN_modulation.1.bias", {6 * H});
      ada_w.insert(ada_w.end(), w.begin(), w.end());
      ada_b.insert(ada_b.end(), bb.begin(), bb.end());
    }
    {
      const std::string p = dn + ".final_layer";
      std::vector<float> w = wm.vec(p + ".adaLN_modulation.1.weight", {5 * H, H});
      std::vector<float> bb = wm.vec(p + ".adaLN_modulation.1.bias", {5 * H});
      ada_w.insert(ada_w.end(), w.begin(), w.end());
      ada_b.insert(ada_b.end(), bb.begin(), bb.end());
      fin = load_convnext(wm, p + ".conv_in");
      conv_out = make_layer(pack_conv(wm.vec(p + ".conv_out.weight", {D, H, 3}), D, H, 3), wm.vec(p + ".conv_out.bias", {D}),
                            H, D, 3, -1, 1, 1, bf());
    }
    ada_n = (int)ada_b.size();  // 4*6H + 5H
    // one projection for every adaLN of every block: rows = (step, sample)
    ada_all = make_layer(ada_w, ada_b, H, ada_n, 1, 0, 1, 1, bf());
    // cond down-sampler
    const int Q = cfg.n_quantizers, CD = cfg.cond_dim;
    qemb = store.upload(wm.vec("quantizer_encoding.quantizer_emb.weight", {Q, CD}));
    int cin = Q * CD;
    for (int s = 0; s < cfg.downsampling_stages; ++s) {
      Stage st;
      const std::string rp = "cond_downsampling.resblocks." + std::to_string(s) + ".block.block";
      const std::string dp = "cond_downsampling.downblocks." + std::to_string(s);
      st.conv_a = make_layer(wm.vec(rp + ".0.weight", {cin, cin, 1}), wm.vec(rp + ".0.bias", {cin}), cin, cin, 1, 0, 1, 1, bf());
      st.gna_w = store.upload(wm.vec(rp + ".1.weight", {cin}));
      st.gna_b = store.upload(wm.vec(rp + ".1.bias", {cin}));
      st.conv_b = make_layer(wm.vec(dp + ".0.weight", {cin / 2, cin, 1}), wm.vec(dp + ".0.bias", {cin / 2}), cin, cin / 2, 1, 0, 1, 1, bf());
      st.gnb_w = store.upload(wm.vec(dp + ".1.weight", {cin / 2}));
      st.gnb_b = store.upload(wm.vec(dp + ".1.bias", {cin / 2}));
      stages.push_back(st);
      cin /= 2;
    }
    proj_out = make_layer(wm.vec("cond_downsampling.proj_out.0.weight", {D, cin}), wm.vec("cond_downsampling.proj_out.0.bias", {D}), cin, D, 1, 0, 1, 1, bf());
  }

  // bf16 mode with a bf16 residual stream: LayerNorm+modulate, depthwise conv and GroupNorm in one kernel, fed by
  // the row statistics the previous GEMM's epilogue left in `rowstat` (dwconv_fused.cu)
  bool fused() const { return bf(); }
  // depthwise conv of the fused front half on the tensor cores (dwconv_tc.cu) instead of the FMA pipe (dwconv_fused.cu)
  bool dw_tensor = [] { const char* e = getenv("FLAMED_B200_DWCONV"); return e && e[0] == 't'; }();
  // MLP-branch LayerNorm applied algebraically (no pass of its own): conv_3's epilogue also writes A2 * h', mlp.0 runs on
  // it and its epilogue applies the per-row part (TapGemm::raff_*)
  bool mlp_ln_fused = [] { const char* e = getenv("FLAMED_B200_MLPLN"); return !(e && e[0] == '0'); }();
  int rowstat3_parts = 0;  // parts written by the last conv_3
  int rowstat_parts = 0;  // parts written by the last GEMM that produced h

  void ln_dwconv_gn(const ConvNeXtW& c, const float* lnw, const float* lnb, const float* shift, const float* scale,
                    const float* gate, const float* ln2_w, const float* scale2, int B, int L, cudaStream_t s) {
    DwFused f;
    memset(&f, 0, sizeof(f));
    f.h = h.as<bf16>(); f.u = bufU.as<bf16>(); f.g = bufG.as<bf16>();
    f.rowstat = rowstat.as<float>(); f.parts = rowstat_parts;
    f.ln_w = lnw; f.ln_b = lnb; f.shift = shift; f.scale = scale; f.mod_bstride = ada_n; f.ln_eps = 1e-6f;
    f.w = c.dw_w; f.wsum = c.dw_wsum; f.bias = c.dw_b; f.gamma = c.gn_w; f.beta = c.gn_b; f.gn_eps = 1e-5f;
    f.B = B; f.L = L; f.C = H; f.tma_encode = ctx->tma_encode;
    // the inner residual u is not written: conv_3's epilogue recomputes it from h (TapGemm::lnu_*)
    f.u = nullptr;
    f.lnu_vecs = scale2 ? 4 : 3; f.ln2_w = ln2_w; f.scale2 = scale2;
    const double elems = (double)B * L * H;
    if (dw_tensor) {  // tensor-core form: statistics merged into gsc / gof by its own last kernel
      {
        ProfScope ps(ctx, KC_DWCONV, s, elems * 2 * (cfg.kernel_size + 2), elems * 4);
        launch_dwconv_tc(f, rowconst.as<float>(), lnab.as<float>(), part.as<float>(), gsc.as<float>(), gof.as<float>(),
                         gate, c.conv3.bias, lnu.as<float>(), ctx->num_sms, s);
      }
      ProfScope ps(ctx, KC_GN_APPLY, s, elems * 2, elems * 2 * 2);
      launch_gn_stream(bufG.p, bufG.p, 1, gsc.as<float>(), gof.as<float>(), B, L, H, s);
      return;
    }
    {
      f.rowconst_out = rowconst.as<float>(); f.lnu_out = lnu.as<float>(); f.gate = gate; f.bias3 = c.conv3.bias;
      ProfScope ps(ctx, KC_DWCONV, s, elems * 2 * (cfg.kernel_size + 2), elems * 4);
      launch_dwconv_ln(f, part.as<float>(), ctx->num_sms, s);
      DwConv dw;
      memset(&dw, 0, sizeof(dw));
      dw.io_bf16 = 1; dw.part = part.as<float>(); dw.B = B; dw.L = L; dw.C = H; dw.KW = cfg.kernel_size;
      dw.gamma = c.gn_w; dw.beta = c.gn_b; dw.eps = 1e-5f; dw.scale = gsc.as<float>(); dw.offset = gof.as<float>();
      launch_dw_merge(dw, s);
    }
    {
      ProfScope ps(ctx, KC_GN_APPLY, s, elems * 2, elems * 2 * 2);
      launch_gn_stream(bufG.p, bufG.p, 1, gsc.as<float>(), gof.as<float>(), B, L, H, s);  // in place: d is still in L2
    }
  }

  // conv_2 (GELU) + conv_3 (gated residual with the inner residual u): hres += gate * (u + conv3(gelu(conv2(g))))
  void convnext_tail(const ConvNeXtW& c, int B, int L, const float* gate, cudaStream_t s, bool scaled_copy = false) {
    const int b16 = bf() ? 1 : 0;
    gemm(problem(c.conv2, bufG.p, H, B, L, L, bufA.p, H, b16, EPI_GELU), c.conv2, bf(), s);
    TapGemm p = problem(c.conv3, bufA.p, H, B, L, L, nullptr, H, 0, EPI_GATE_RESID);
    p.gate = gate; p.gate_bstride = ada_n; p.addend = bufU.p; p.ld_add = H; p.addend_bf16 = b16;
    if (fused()) {  // u recomputed in the epilogue (row constants and table written by the depthwise kernel)
      p.addend = nullptr; p.bias = nullptr;
      p.lnu_rowconst = rowconst.as<float>(); p.lnu_table = lnu.as<float>(); p.lnu_vecs = scaled_copy ? 4 : 3;
      if (scaled_copy) {  // A2 * h' for mlp.0 and the row statistics of h' for its epilogue
        p.out2 = bufU.p; p.ld_out2 = H;
        p.rowstat = rowstat.as<float>();
        p.rowstat_parts = rowstat3_parts = tapgemm_tc2_rowstat_parts(p, ctx->num_sms);
      }
    }
    p.hres = h.as<float>(); p.ld_res = H; p.hres_bf16 = h16 ? 1 : 0;
    gemm(p, c.conv3, bf(), s);
  }

  // ---- one ConvNeXt in the fp32 parity mode: hres += gate * (u + conv3(gelu(conv2(GN(dwconv(u))))))
  //      (prob_generator.py:107-111,162); the depthwise kernel finalises the GroupNorm statistics itself
  void convnext(const ConvNeXtW& c, int B, int L, const float* gate, cudaStream_t s) {
    DwConv dw;
    memset(&dw, 0, sizeof(dw));
    dw.x = bufU.p; dw.y = bufD.p; dw.io_bf16 = 0; dw.w = c.dw_w; dw.bias = c.dw_b; dw.part = part.as<float>();
    dw.B = B; dw.L = L; dw.C = H; dw.KW = cfg.kernel_size;
    dw.gamma = c.gn_w; dw.beta = c.gn_b; dw.eps = 1e-5f; dw.scale = gsc.as<float>(); dw.offset = gof.as<float>();
    dw.counters = gctr.as<int>();
    const double elems = (double)B * L * H, eb = (double)esize();
    {
      ProfScope ps(ctx, KC_DWCONV, s, elems * 2 * cfg.kernel_size, elems * 2 * eb);
      launch_dwconv(dw, s);
    }
    {
      ProfScope ps(ctx, KC_GN_APPLY, s, elems * 2, elems * 2 * eb);
      launch_gn_stream(bufD.p, bufG.p, 0, gsc.as<float>(), gof.as<float>(), B, L, H, s);
    }
    convnext_tail(c, B, L, gate, s);
  }

  void ln_modulate(const float* w, const float* b, const float* shift, const float* scale, int B, int L, void* y,
                   cudaStream_t s) {
    LnMod ln;
    memset(&ln, 0, sizeof(ln));
    ln.x = h.p; ln.ldx = H; ln.x_bf16 = h16 ? 1 : 0; ln.y = y; ln.ldy = H; ln.y_bf16 = bf() ? 1 : 0;
    ln.w = w; ln.b = b; ln.shift = shift; ln.scale = scale; ln.mod_bstride = ada_n; ln.scale_plus_one = 1.f;
    ln.eps = 1e-6f; ln.rows = (int64_t)B * L; ln.rows_per_batch = L; ln.C = H;
    ProfScope ps(ctx, KC_LN_MOD, s, (double)B * L * H * 8, (double)B * L * H * (hsize() + esize()));
    launch_ln_mod(ln, s);
  }

  // adaLN table for `nfe` time points: ada[(i*B+b), :]   (hoisted out of the loop)
  void modulation_table(int B, int nfe, cudaStream_t s) {
    table_rows = nfe * B;
    launch_timestep_embedding(ts_s.as<float>(), nfe, 256, tfreq.as<float>(), s);
    gemm(problem(time0, tfreq.p, 256, 1, nfe, nfe, teh.p, H, 0, EPI_SILU), time0, false, s);
    gemm(problem(time2, teh.p, H, 1, nfe, nfe, temb.p, H, 0, EPI_NONE), time2, false, s);
    gemm(problem(cond_embed, spk_s.p, cfg.spk_dim, 1, B, B, cvec.p, H, 0, EPI_NONE), cond_embed, false, s);
    launch_silu_sum(temb.as<float>(), cvec.as<float>(), nfe, B, H, sbuf.p, bf() ? 1 : 0, s);
    gemm(problem(ada_all, sbuf.p, H, 1, nfe * B, nfe * B, ada.p, ada_n, 0, EPI_NONE), ada_all, bf(), s);
    if (fused() && mlp_ln_fused) {
      // c1 = W0 . A2, c2 = W0 . B2 + b0 for every (step, sample) of every block: the column part of the MLP-branch LayerNorm
      const int64_t rows = (int64_t)nfe * B;
      for (size_t k = 0; k < blocks.size(); ++k) {
        const ResBlockW& rb = blocks[k];
        const float* a = ada.as<float>() + k * 6 * H;
        launch_ln_affine_rows(rb.lnm_w, rb.lnm_b, a + 3 * H, a + 4 * H, ada_n, rows, H, s1buf.as<bf16>(), s2buf.as<bf16>(), s);
        TapGemm p1 = problem(rb.mlp0, s1buf.p, H, 1, (int)rows, (int)rows, c1tab.as<float>() + k * rows * H, H, 0, EPI_NONE);
        p1.bias = nullptr;
        gemm(p1, rb.mlp0, true, s);
        gemm(problem(rb.mlp0, s2buf.p, H, 1, (int)rows, (int)rows, c2tab.as<float>() + k * rows * H, H, 0, EPI_NONE), rb.mlp0, true, s);
      }
    }
  }
  int table_rows = 0;  // nfe * B of the current modulation table

  // one velocity evaluation at table row `i`, accumulated as target += alpha * v
  void step(int B, int L, int i, float* target, float alpha, cudaStream_t s) {
    const int64_t M = (int64_t)B * L;
    const float* xin = x.as<float>();
    const bool fz = fused();
    if (bf()) {
      // xb = bf16(x): written by the previous step's Euler epilogue when the loop accumulates into x itself
      if (!xb_fresh) launch_f32_to_bf16(xin, xb.as<bf16>(), M * D, s);
      TapGemm pi = problem(proj_in, xb.p, D, B, L, L, h.p, H, h16 ? 1 : 0, EPI_NONE);
      if (fz) {  // the epilogue also leaves the LayerNorm row statistics of h for the first block's fused kernel
        pi.rowstat = rowstat.as<float>();
        pi.rowstat_parts = rowstat_parts = tapgemm_tc2_rowstat_parts(pi, ctx->num_sms);
      }
      gemm(pi, proj_in, true, s);
    } else {
      gemm(problem(proj_in, xin, D, B, L, L, h.p, H, 0, EPI_NONE), proj_in, false, s);
    }
    const float* arow = ada.as<float>() + (int64_t)i * B * ada_n;
    const int b16 = bf() ? 1 : 0;
    for (size_t k = 0; k < blocks.size(); ++k) {
      const ResBlockW& rb = blocks[k];
      const float* a = arow + k * 6 * H;  // shift_c, scale_c, gate_c, shift_m, scale_m, gate_m
      const bool mf = fz && mlp_ln_fused;
      if (fz) {
        ln_dwconv_gn(rb.cn, rb.lnc_w, rb.lnc_b, a, a + H, a + 2 * H, mf ? rb.lnm_w : nullptr, mf ? a + 4 * H : nullptr, B, L, s);
        convnext_tail(rb.cn, B, L, a + 2 * H, s, mf);
      } else {
        ln_modulate(rb.lnc_w, rb.lnc_b, a, a + H, B, L, bufU.p, s);
        convnext(rb.cn, B, L, a + 2 * H, s);
      }
      TapGemm p0 = problem(rb.mlp0, bufU.p, H, B, L, L, bufA.p, H, b16, EPI_SILU);
      if (mf) {  // LayerNorm + modulate of the MLP branch applied algebraically in mlp.0's epilogue
        const int64_t trow = ((int64_t)k * table_rows + (int64_t)i * B) * H;
        p0.bias = nullptr;
        p0.raff_rowstat = rowstat.as<float>(); p0.raff_parts = rowstat3_parts;
        p0.raff_c1 = c1tab.as<float>() + trow; p0.raff_c2 = c2tab.as<float>() + trow;
        p0.raff_eps = 1e-6f; p0.raff_ln_dim = H;
      } else {
        ln_modulate(rb.lnm_w, rb.lnm_b, a + 3 * H, a + 4 * H, B, L, bufU.p, s);
      }
      gemm(p0, rb.mlp0, bf(), s);
      TapGemm p = problem(rb.mlp2, bufA.p, H, B, L, L, nullptr, H, 0, EPI_GATE_RESID);
      p.gate = a + 5 * H; p.gate_bstride = ada_n; p.hres = h.as<float>(); p.ld_res = H; p.hres_bf16 = h16 ? 1 : 0;
      if (fz) {  // row statistics of the updated h for the next block's (or the final layer's) fused kernel
        p.rowstat = rowstat.as<float>();
        p.rowstat_parts = rowstat_parts = tapgemm_tc2_rowstat_parts(p, ctx->num_sms);
      }
      gemm(p, rb.mlp2, bf(), s);
    }
    const float* a = arow + blocks.size() * 6 * H;  // shift_c, scale_c, gate_c, shift_m, scale_m
    if (fz) {
      ln_dwconv_gn(fin, nullptr, nullptr, a, a + H, a + 2 * H, nullptr, nullptr, B, L, s);
      convnext_tail(fin, B, L, a + 2 * H, s);
    } else {
      ln_modulate(nullptr, nullptr, a, a + H, B, L, bufU.p, s);
      convnext(fin, B, L, a + 2 * H, s);
    }
    ln_modulate(nullptr, nullptr, a + 3 * H, a + 4 * H, B, L, bufU.p, s);
    TapGemm p = problem(conv_out, bufU.p, H, B, L, L, nullptr, D, 0, EPI_EULER);
    p.hres = target; p.ld_res = D; p.alpha = alpha;
    xb_fresh = bf() && target == x.as<float>();
    if (xb_fresh) p.out = xb.p;
    gemm(p, conv_out, bf(), s);
  }
  bool xb_fresh = false;  // xb holds bf16(x) of the current state
  int launches_per_step() const {  // steady state (the first step of a loop adds the f32 -> bf16 conversion of x0)
    const int nb = (int)blocks.size();
    // bf16: proj_in + per block (LayerNorm-fused depthwise conv [tensor-core form: + 2 helper kernels], statistics merge,
    // GroupNorm apply, conv_2, conv_3, [LayerNorm unless applied algebraically], mlp.0, mlp.2) + the FinalLayer's
    // (depthwise, merge, GroupNorm apply, conv_2, conv_3, LayerNorm, conv_out)
    if (fused()) return 1 + nb * (mlp_ln_fused ? 7 : 8) + 7 + (dw_tensor ? nb + 1 : 0);
    return 1 + nb * 8 + 7;  // fp32: proj_in + per block (LN, dwconv, GN apply, 2 GEMMs, LN, 2 GEMMs) + the FinalLayer's 7
  }

  bool ensure(int B, int L, int nfe) {
    const int64_t M = (int64_t)B * L;
    const size_t e = esize();
    bool moved = false;
    moved |= cond_s.ensure(M * D * 4); moved |= spk_s.ensure((size_t)B * cfg.spk_dim * 4);
    moved |= noise_s.ensure(M * D * 4); moved |= ts_s.ensure((size_t)(nfe + 1) * 4); moved |= seed_s.ensure(8);
    moved |= x.ensure(M * D * 4); moved |= xb.ensure(M * D * 2); moved |= vout.ensure(M * D * 4);
    moved |= h.ensure(M * H * 4);
    moved |= bufU.ensure(M * H * e); moved |= bufD.ensure(M * H * e); moved |= bufG.ensure(M * H * e);
    moved |= bufA.ensure(M * H * e);
    moved |= part.ensure(std::max((size_t)B * dw_nchunk(L) * H * 2 * 4, dwconv_tc_part_bytes(B, L, H)));
    moved |= rowconst.ensure((size_t)M * 8); moved |= lnab.ensure((size_t)B * H * 8); moved |= lnu.ensure((size_t)B * H * 16);
    if (fused() && mlp_ln_fused) {
      moved |= s1buf.ensure((size_t)nfe * B * H * 2); moved |= s2buf.ensure((size_t)nfe * B * H * 2);
      moved |= c1tab.ensure(blocks.size() * (size_t)nfe * B * H * 4); moved |= c2tab.ensure(blocks.size() * (size_t)nfe * B * H * 4);
    }
    moved |= gsc.ensure((size_t)B * H * 4); moved |= gof.ensure((size_t)B * H * 4);
    if (gctr.ensure((size_t)B * (H / 256) * 4)) {  // arrival tickets of the depthwise kernel start at zero
      moved = true;
      FLM_CUDA(cudaMemset(gctr.p, 0, gctr.bytes));
    }
    moved |= rowstat.ensure((size_t)M * 32 * 8);  // (sum, sumsq) partials of the rows of h: <= 32 parts (N tile 64)
    moved |= ada.ensure((size_t)nfe * B * ada_n * 4); moved |= sbuf.ensure((size_t)nfe * B * H * e);
    moved |= temb.ensure((size_t)nfe * H * 4); moved |= tfreq.ensure((size_t)nfe * 256 * 4);
    moved |= teh.ensure((size_t)nfe * H * 4); moved |= cvec.ensure((size_t)B * H * 4);
    if (moved) graphs.clear();
    return moved;
  }
};

extern "C" int flm_denoiser_load(flm_ctx* ctx, const flm_tensor* weights, int n, const flm_prob_cfg* cfg, int mode,
                                 flm_denoiser** out) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && weights && cfg && out, "null argument");
  FLM_REQUIRE(mode == FLM_F32 || mode == FLM_BF16, "bad mode");
  FLM_REQUIRE(cfg->hidden_dim % 256 == 0 && cfg->hidden_dim <= 1024, "hidden_dim must be a multiple of 256, <= 1024");
  FLM_REQUIRE(cfg->target_dim % 64 == 0, "target_dim must be a multiple of 64");
  DeviceGuard dguard(ctx->device);
  WeightMap wm(weights, n);
  std::unique_ptr<flm_denoiser> h(new flm_denoiser(ctx, mode));
  h->cfg = *cfg;
  // the throughput mode keeps the residual stream of a step in bf16 (128-step latents stay at 2.8e-3 rel-L2 of the
  // fp32 reference, 2.5e-3 with an fp32 stream); the ODE state x_t itself is fp32 in both modes
  h->h16 = mode == FLM_BF16;
  h->load(wm);
  *out = h.release();
  FLM_API_END
}
extern "C" void flm_denoiser_destroy(flm_denoiser* h) { delete h; }
extern "C" int flm_denoiser_launches_per_step(flm_denoiser* h) { return h ? h->launches_per_step() : 0; }

extern "C" int flm_cond_prepare(flm_denoiser* h, const float* prior_embs, const uint8_t* mask, int B, int L,
                                float* out_cond, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && prior_embs && mask && out_cond, "null argument");
  DeviceGuard dguard(h->ctx->device);
  if (B == 0 || L == 0) return FLM_OK;
  cudaStream_t s = S(stream);
  const int Q = h->cfg.n_quantizers, CD = h->cfg.cond_dim;
  const int64_t M = (int64_t)B * L;
  int cin = Q * CD;
  const size_t e = h->esize();
  const int b16 = h->bf() ? 1 : 0;
  h->c_xq.ensure(M * cin * e); h->c_xm.ensure(M * cin * e); h->c_h.ensure(M * cin * e);
  h->c_part.ensure((size_t)B * gs_nchunk(L) * 8 * 2 * 4);
  h->c_sc.ensure((size_t)B * cin * 4); h->c_of.ensure((size_t)B * cin * 4);
  launch_quantizer_fold(prior_embs, h->qemb, mask, B, Q, L, CD, h->c_xq.p, h->c_xm.p, b16, s);
  void* xq = h->c_xq.p;  // current x (unmasked)
  void* xm = h->c_xm.p;  // GEMM input
  void* hb = h->c_h.p;
  for (size_t st = 0; st < h->stages.size(); ++st) {
    const flm_denoiser::Stage& sg = h->stages[st];
    if (st > 0)  // configs/prob.yaml uses 1 stage; a second stage needs x*mask of the previous output
      throw Error(FLM_ERR_UNSUPPORTED, "downsampling_stages > 1 is not implemented");
    // ResnetBlock1D: x + Mish(GN8(conv(x*m))) * m   (prob_generator.py:11-32)
    h->gemm(h->problem(sg.conv_a, xm, cin, B, L, L, hb, cin, b16, EPI_NONE), sg.conv_a, h->bf(), s);
    launch_group_stats(hb, b16, B, L, cin, 8, h->c_part.as<float>(), s);
    launch_gn_finalize(h->c_part.as<float>(), B, L, cin, 8, gs_nchunk(L), GS_ROWS, sg.gna_w, sg.gna_b, 1e-5f,
                       h->c_sc.as<float>(), h->c_of.as<float>(), s);
    GnApply ga;
    memset(&ga, 0, sizeof(ga));
    ga.x = hb; ga.x_bf16 = b16; ga.y = xm; ga.y_bf16 = b16; ga.scale = h->c_sc.as<float>(); ga.offset = h->c_of.as<float>();
    ga.mask = mask; ga.res = xq; ga.res_bf16 = b16; ga.act = 2; ga.B = B; ga.L = L; ga.C = cin;
    launch_gn_apply(ga, s);
    // down block: ReLU(GN8(conv(x)))   (prob_generator.py:181-192)
    h->gemm(h->problem(sg.conv_b, xm, cin, B, L, L, hb, cin / 2, b16, EPI_NONE), sg.conv_b, h->bf(), s);
    launch_group_stats(hb, b16, B, L, cin / 2, 8, h->c_part.as<float>(), s);
    launch_gn_finalize(h->c_part.as<float>(), B, L, cin / 2, 8, gs_nchunk(L), GS_ROWS, sg.gnb_w, sg.gnb_b, 1e-5f,
                       h->c_sc.as<float>(), h->c_of.as<float>(), s);
    memset(&ga, 0, sizeof(ga));
    ga.x = hb; ga.x_bf16 = b16; ga.y = xq; ga.y_bf16 = b16; ga.scale = h->c_sc.as<float>(); ga.offset = h->c_of.as<float>();
    ga.act = 1; ga.B = B; ga.L = L; ga.C = cin / 2;
    launch_gn_apply(ga, s);
    cin /= 2;
  }
  // proj_out: ReLU(Linear)   (prob_generator.py:194-197)
  h->gemm(h->problem(h->proj_out, xq, cin, B, L, L, out_cond, h->D, 0, EPI_RELU), h->proj_out, h->bf(), s);
  FLM_API_END
}

extern "C" int flm_denoiser_sample(flm_denoiser* h, const float* cond, const float* spk, const float* noise,
                                   uint64_t seed, const float* ts_host, int B, int L, int nfe, float temperature,
                                   float* out_latents, int use_graph, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && cond && spk && ts_host && out_latents, "null argument");
  FLM_REQUIRE(nfe >= 1, "nfe must be >= 1");
  DeviceGuard dguard(h->ctx->device);
  if (B == 0 || L == 0) return FLM_OK;
  cudaStream_t s = S(stream);
  const int64_t M = (int64_t)B * L;
  const bool philox = noise == nullptr;  // x0 = temperature * N(0,1) + cond drawn in the init kernel (documented map)
  h->ensure(B, L, nfe);
  FLM_CUDA(cudaMemcpyAsync(h->cond_s.p, cond, M * h->D * 4, cudaMemcpyDeviceToDevice, s));
  FLM_CUDA(cudaMemcpyAsync(h->spk_s.p, spk, (size_t)B * h->cfg.spk_dim * 4, cudaMemcpyDeviceToDevice, s));
  if (philox) FLM_CUDA(cudaMemcpyAsync(h->seed_s.p, &seed, 8, cudaMemcpyHostToDevice, s));
  else FLM_CUDA(cudaMemcpyAsync(h->noise_s.p, noise, M * h->D * 4, cudaMemcpyDeviceToDevice, s));
  FLM_CUDA(cudaMemcpyAsync(h->ts_s.p, ts_host, (size_t)(nfe + 1) * 4, cudaMemcpyHostToDevice, s));
  const float dt = (float)(1.0 / nfe);
  auto body = [&](cudaStream_t cs) {
    h->xb_fresh = false;
    h->modulation_table(B, nfe, cs);
    if (philox)
      launch_philox_normal(h->seed_s.as<uint64_t>(), 2u, temperature, h->cond_s.as<float>(), M * h->D, h->x.as<float>(), cs);
    else
      launch_noise_init(h->noise_s.as<float>(), h->cond_s.as<float>(), temperature, M * h->D, h->x.as<float>(), cs);
    for (int i = 0; i < nfe; ++i) h->step(B, L, i, h->x.as<float>(), dt, cs);
  };
  if (use_graph) {
    int tbits;
    memcpy(&tbits, &temperature, 4);
    h->graphs.run(std::make_tuple(B, L, philox ? -nfe : nfe, tbits), s, body);
  } else {
    body(s);
  }
  FLM_CUDA(cudaMemcpyAsync(out_latents, h->x.p, M * h->D * 4, cudaMemcpyDeviceToDevice, s));
  FLM_API_END
}

// the (n) standard-normal values of the documented Philox map for tensor `tensor_id` (0 duration, 1 silence, 2 latent):
// what flm_durgen_sample / flm_denoiser_sample draw internally when their noise pointers are NULL.  For tests.
extern "C" int flm_philox_normal(flm_ctx* ctx, uint64_t seed, int tensor_id, int64_t n, float* out, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && out && n >= 0 && tensor_id >= 0, "bad arguments");
  DeviceGuard dguard(ctx->device);
  DevBuf sd;
  sd.ensure(8);
  FLM_CUDA(cudaMemcpyAsync(sd.p, &seed, 8, cudaMemcpyHostToDevice, S(stream)));
  launch_philox_normal(sd.as<uint64_t>(), (uint32_t)tensor_id, 1.0f, nullptr, n, out, S(stream));
  FLM_CUDA(cudaStreamSynchronize(S(stream)));  // sd is freed on return
  FLM_API_END
}

extern "C" int flm_denoiser_forward(flm_denoiser* h, const float* x, const float* spk, float t, int B, int L,
                                    float* out_v, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && x && spk && out_v, "null argument");
  DeviceGuard dguard(h->ctx->device);
  if (B == 0 || L == 0) return FLM_OK;
  cudaStream_t s = S(stream);
  const int64_t M = (int64_t)B * L;
  h->ensure(B, L, 1);
  const float ts[2] = {t, 1.0f};
  FLM_CUDA(cudaMemcpyAsync(h->x.p, x, M * h->D * 4, cudaMemcpyDeviceToDevice, s));
  FLM_CUDA(cudaMemcpyAsync(h->spk_s.p, spk, (size_t)B * h->cfg.spk_dim * 4, cudaMemcpyDeviceToDevice, s));
  FLM_CUDA(cudaMemcpyAsync(h->ts_s.p, ts, 8, cudaMemcpyHostToDevice, s));
  FLM_CUDA(cudaMemsetAsync(h->vout.p, 0, M * h->D * 4, s));
  h->modulation_table(B, 1, s);
  h->xb_fresh = false;
  h->step(B, L, 0, h->vout.as<float>(), 1.0f, s);
  FLM_CUDA(cudaMemcpyAsync(out_v, h->vout.p, M * h->D * 4, cudaMemcpyDeviceToDevice, s));
  FLM_API_END
}

// ===================================================================================== codec
namespace {
struct ActW {
  float *a, *invb;
  float fu[12], fd[12];
  int C;
};
struct ResUnitW {
  ActW act1, act2;
  Layer conv7, conv1;
};

ActW load_act(WeightStore& store, const WeightMap& wm, const std::string& p, int C) {
  ActW a;
  a.C = C;
  std::vector<float> al = wm.vec(p + ".act.alpha", {C}), be = wm.vec(p + ".act.beta", {C});
  std::vector<float> ea(C), ib(C);
  for (int i = 0; i < C; ++i) {
    ea[i] = expf(al[i]);                          // alpha_logscale (facodec.py:113-114)
    ib[i] = 1.0f / (expf(be[i]) + 0.000000001f);  // facodec.py:116
  }
  a.a = store.upload(ea);
  a.invb = store.upload(ib);
  std::vector<float> fu = wm.vec(p + ".upsample.filter", {1, 1, 12});
  std::vector<float> fd = wm.vec(p + ".downsample.lowpass.filter", {1, 1, 12});
  for (int i = 0; i < 12; ++i) { a.fu[i] = fu[i]; a.fd[i] = fd[i]; }
  return a;
}

Layer load_wn_conv(Engine& e, const WeightMap& wm, const std::string& p, int N, int K, int k, int off0, int dil,
                   int stride, bool want_bf16) {
  std::vector<float> w = fold_weight_norm(wm.vec(p + ".weight_g", {N, 1, 1}), wm.vec(p + ".weight_v", {N, K, k}), N);
  return e.make_layer(pack_conv(w, N, K, k), wm.vec(p + ".bias", {N}), K, N, k, off0, dil, stride, want_bf16);
}

ResUnitW load_res_unit(Engine& e, const WeightMap& wm, const std::string& p, int C, int dil, bool want_bf16) {
  ResUnitW r;
  r.act1 = load_act(e.store, wm, p + ".block.0", C);
  r.conv7 = load_wn_conv(e, wm, p + ".block.1", C, C, 7, -3 * dil, dil, 1, want_bf16);
  r.act2 = load_act(e.store, wm, p + ".block.2", C);
  r.conv1 = load_wn_conv(e, wm, p + ".block.3", C, C, 1, 0, 1, 1, want_bf16);
  return r;
}

void run_act(const Engine& e, const ActW& a, const void* x, void* y, int B, int T, cudaStream_t s) {
  Act1d p;
  p.x = x; p.y = y; p.io_bf16 = e.bf() ? 1 : 0; p.a = a.a; p.invb = a.invb;
  memcpy(p.fu, a.fu, sizeof(p.fu));
  memcpy(p.fd, a.fd, sizeof(p.fd));
  p.B = B; p.T = T; p.C = a.C; p.fast_sin = e.bf() ? 1 : 0;
  const double elems = (double)B * T * a.C;
  ProfScope ps(e.ctx, KC_ACT1D, s, elems * 52, elems * 2 * e.esize());
  launch_act1d(p, s);
}

// x <- x + conv1(act(conv7(act(x))))   (facodec.py:121-133); t1,t2 scratch
void run_res_unit(const Engine& e, const ResUnitW& r, void* x, void* t1, void* t2, int B, int T, cudaStream_t s) {
  const int C = r.act1.C, b16 = e.bf() ? 1 : 0;
  run_act(e, r.act1, x, t1, B, T, s);
  e.gemm(e.problem(r.conv7, t1, C, B, T, T, t2, C, b16, EPI_NONE), r.conv7, e.bf(), s);
  run_act(e, r.act2, t2, t1, B, T, s);
  TapGemm p = e.problem(r.conv1, t1, C, B, T, T, x, C, b16, EPI_RESID);
  p.resid_in = x;
  e.gemm(p, r.conv1, e.bf(), s);
}
}  // namespace

// prompt side: residual VQs + timbre transformer (facodec.py:470-507; fvq.py; transformer.py:86-234)
struct PromptSide {
  bool present = false;
  VqPlan plan;
  int D = 0;
  struct TLayer { float *ln1w, *ln1b, *ln2w, *ln2b; Layer qkv, out, ffn1, ffn2; };
  std::vector<TLayer> layers;
  float *last_w = nullptr, *last_b = nullptr, *pe = nullptr;
  int pe_rows = 0, heads = 4;
  DevBuf xt, xa, xn, qkv, att, ffh, qg, pooled_in;
};

struct flm_codec_dec : Engine {
  PromptSide prompt;
  Layer timbre_linear, conv0;
  struct Block { ActW act; Layer up; int stride, cin, cout; ResUnitW ru[3]; };
  std::vector<Block> blocks;
  ActW act_out;
  float* wout; float bout; int cout_final;
  std::map<std::string, ActW> acts_by_name;
  DevBuf style, buf[3], tbuf;
  flm_codec_dec(flm_ctx* c, int mode) : Engine(c, mode) {}
};

extern "C" int flm_codec_dec_load(flm_ctx* ctx, const flm_tensor* weights, int n, int mode, flm_codec_dec** out) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && weights && out, "null argument");
  FLM_REQUIRE(mode == FLM_F32 || mode == FLM_BF16, "bad mode");
  DeviceGuard dguard(ctx->device);
  WeightMap wm(weights, n);
  std::unique_ptr<flm_codec_dec> h(new flm_codec_dec(ctx, mode));
  const bool b = h->bf();
  const flm_tensor& w0 = wm.get("model.0.weight_v");
  const int C0 = (int)w0.shape[0], Cin = (int)w0.shape[1];
  FLM_REQUIRE(w0.ndim == 3 && w0.shape[2] == 7, "model.0 must be a k=7 conv");
  h->timbre_linear = h->make_layer(wm.vec("timbre_linear.weight", {2 * Cin, Cin}), wm.vec("timbre_linear.bias", {2 * Cin}),
                                   Cin, 2 * Cin, 1, 0, 1, 1, false);
  h->conv0 = load_wn_conv(*h, wm, "model.0", C0, Cin, 7, -3, 1, 1, b);
  int c = C0, i = 1;
  while (wm.has("model." + std::to_string(i) + ".block.1.weight_v")) {
    const std::string p = "model." + std::to_string(i);
    const flm_tensor& wv = wm.get(p + ".block.1.weight_v");  // (Cin, Cout, 2s)
    flm_codec_dec::Block blk;
    blk.cin = (int)wv.shape[0]; blk.cout = (int)wv.shape[1]; blk.stride = (int)wv.shape[2] / 2;
    FLM_REQUIRE(blk.cin == c, "decoder block channel mismatch");
    blk.act = load_act(h->store, wm, p + ".block.0", blk.cin);
    h->acts_by_name[p + ".block.0"] = blk.act;
    std::vector<float> w = fold_weight_norm(wm.vec(p + ".block.1.weight_g", {blk.cin, 1, 1}),
                                            wm.vec(p + ".block.1.weight_v", {blk.cin, blk.cout, 2 * blk.stride}), blk.cin);
    blk.up = h->make_layer(pack_conv_transpose(w, blk.cin, blk.cout, blk.stride),
                           replicate(wm.vec(p + ".block.1.bias", {blk.cout}), blk.stride), blk.cin,
                           blk.stride * blk.cout, 3, -1, 1, 1, b);
    blk.up.alg_scale = 2.0f / 3.0f;  // each output frame really uses 2 of the 3 zero-padded taps
    const int dils[3] = {1, 3, 9};
    for (int j = 0; j < 3; ++j) {
      const std::string rp = p + ".block." + std::to_string(j + 2);
      blk.ru[j] = load_res_unit(*h, wm, rp, blk.cout, dils[j], b);
      h->acts_by_name[rp + ".block.0"] = blk.ru[j].act1;
      h->acts_by_name[rp + ".block.2"] = blk.ru[j].act2;
    }
    h->blocks.push_back(blk);
    c = blk.cout;
    ++i;
  }
  FLM_REQUIRE(!h->blocks.empty(), "no decoder blocks found (model.1.block.1.weight_v missing)");
  const std::string pa = "model." + std::to_string(i), pc = "model." + std::to_string(i + 1);
  h->act_out = load_act(h->store, wm, pa, c);
  h->acts_by_name[pa] = h->act_out;
  std::vector<float> w = fold_weight_norm(wm.vec(pc + ".weight_g", {1, 1, 1}), wm.vec(pc + ".weight_v", {1, c, 7}), 1);
  std::vector<float> wt((size_t)7 * c);
  for (int ch = 0; ch < c; ++ch)
    for (int t = 0; t < 7; ++t) wt[(size_t)t * c + ch] = w[(size_t)ch * 7 + t];
  h->wout = h->store.upload(wt);
  h->bout = wm.vec(pc + ".bias", {1})[0];
  h->cout_final = c;
  // ---- prompt side (optional: present in FACodecDecoder.state_dict(), not needed by .inference())
  if (wm.has("quantizer.0.layers.0.in_proj.weight_v") && wm.has("timbre_encoder.last_ln.weight")) {
    PromptSide& ps = h->prompt;
    const flm_tensor& wi = wm.get("quantizer.0.layers.0.in_proj.weight_v");  // (cd, D)
    ps.D = (int)wi.shape[1];
    ps.plan.n_layers = 0;
    for (int g = 0; g < 3; ++g) {
      for (int l = 0;; ++l) {
        const std::string p = "quantizer." + std::to_string(g) + ".layers." + std::to_string(l);
        if (!wm.has(p + ".in_proj.weight_v")) break;
        FLM_REQUIRE(ps.plan.n_layers < VQ_MAX_LAYERS, "too many quantiser layers");
        const flm_tensor& cbk = wm.get(p + "._codebook.weight");
        const int ncode = (int)cbk.shape[0], cd = (int)cbk.shape[1];
        VqLayer& L = ps.plan.layer[ps.plan.n_layers++];
        L.cd = cd; L.n_codes = ncode; L.group = g;
        // weight_norm of nn.Linear: norm over dim 1 per output row
        L.w_in = h->store.upload(fold_weight_norm(wm.vec(p + ".in_proj.weight_g", {cd, 1}), wm.vec(p + ".in_proj.weight_v", {cd, ps.D}), cd));
        L.b_in = h->store.upload(wm.vec(p + ".in_proj.bias", {cd}));
        L.w_out = h->store.upload(fold_weight_norm(wm.vec(p + ".out_proj.weight_g", {ps.D, 1}), wm.vec(p + ".out_proj.weight_v", {ps.D, cd}), ps.D));
        L.b_out = h->store.upload(wm.vec(p + ".out_proj.bias", {ps.D}));
        std::vector<float> cb = wm.vec(p + "._codebook.weight", {ncode, cd}), cn(cb.size()), sq(ncode);
        for (int i = 0; i < ncode; ++i) {  // F.normalize(codebook) and its squared norms, in fp32 like the reference
          float n2 = 0.f;
          for (int j = 0; j < cd; ++j) n2 += cb[(size_t)i * cd + j] * cb[(size_t)i * cd + j];
          const float inv = 1.0f / std::max(std::sqrt(n2), 1e-12f);
          float s2 = 0.f;
          for (int j = 0; j < cd; ++j) {
            cn[(size_t)i * cd + j] = cb[(size_t)i * cd + j] * inv;
            s2 += cn[(size_t)i * cd + j] * cn[(size_t)i * cd + j];
          }
          sq[i] = s2;
        }
        L.cb = h->store.upload(cb); L.cb_norm = h->store.upload(cn); L.cb_sq = h->store.upload(sq);
      }
    }
    const int D = ps.D;
    for (int l = 0;; ++l) {
      const std::string p = "timbre_encoder.layers." + std::to_string(l);
      if (!wm.has(p + ".ln_1.weight")) break;
      PromptSide::TLayer t;
      t.ln1w = h->store.upload(wm.vec(p + ".ln_1.weight", {D})); t.ln1b = h->store.upload(wm.vec(p + ".ln_1.bias", {D}));
      t.ln2w = h->store.upload(wm.vec(p + ".ln_2.weight", {D})); t.ln2b = h->store.upload(wm.vec(p + ".ln_2.bias", {D}));
      t.qkv = h->make_layer(wm.vec(p + ".self_attn.in_proj_weight", {3 * D, D}), wm.vec(p + ".self_attn.in_proj_bias", {3 * D}), D, 3 * D, 1, 0, 1, 1, false);
      t.out = h->make_layer(wm.vec(p + ".self_attn.out_proj.weight", {D, D}), wm.vec(p + ".self_attn.out_proj.bias", {D}), D, D, 1, 0, 1, 1, false);
      const flm_tensor& f1 = wm.get(p + ".ffn.ffn_1.weight");  // (F, D, k)
      const int F = (int)f1.shape[0], k = (int)f1.shape[2];
      t.ffn1 = h->make_layer(pack_conv(wm.vec(p + ".ffn.ffn_1.weight", {F, D, k}), F, D, k), wm.vec(p + ".ffn.ffn_1.bias", {F}), D, F, k, -(k / 2), 1, 1, false);
      t.ffn2 = h->make_layer(wm.vec(p + ".ffn.ffn_2.weight", {D, F}), wm.vec(p + ".ffn.ffn_2.bias", {D}), F, D, 1, 0, 1, 1, false);
      ps.layers.push_back(t);
    }
    ps.last_w = h->store.upload(wm.vec("timbre_encoder.last_ln.weight", {D}));
    ps.last_b = h->store.upload(wm.vec("timbre_encoder.last_ln.bias", {D}));
    const flm_tensor& pe = wm.get("timbre_encoder.position_emb.pe");  // (max_len, 1, D)
    ps.pe_rows = (int)pe.shape[0];
    ps.pe = h->store.upload(wm.vec("timbre_encoder.position_emb.pe", {ps.pe_rows, 1, D}));
    ps.present = true;
  }
  *out = h.release();
  FLM_API_END
}
extern "C" void flm_codec_dec_destroy(flm_codec_dec* h) { delete h; }

// replaces: FACodecDecoder.forward(vq=True) (facodec.py:509-533): quantise the prompt encoder output and pool the timbre.
// enc_out (B,D,T) f32 in the reference layout -> codes (n_q, B, T) i64, quantized (3, B, D, T) f32 (per quantiser
// group, nullable), spk (B, D) f32
extern "C" int flm_codec_dec_prompt(flm_codec_dec* h, const float* enc_out, int B, int T, int64_t* out_codes,
                                    float* out_quantized, float* out_spk, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && enc_out && out_codes && out_spk, "null argument");
  PromptSide& ps = h->prompt;
  if (!ps.present) throw Error(FLM_ERR_WEIGHT, "the decoder was loaded without quantizer.* / timbre_encoder.* weights");
  FLM_REQUIRE(B <= ps.pe_rows, "batch larger than the positional table (the reference indexes it by the batch axis)");
  DeviceGuard dguard(h->ctx->device);
  if (B == 0 || T == 0) return FLM_OK;
  cudaStream_t s = S(stream);
  const int D = ps.D;
  const int64_t rows = (int64_t)B * T;
  const size_t act = (size_t)rows * D * 4;
  ps.xt.ensure(act); ps.xa.ensure(act); ps.xn.ensure(act); ps.att.ensure(act); ps.qkv.ensure(act * 3);
  ps.qg.ensure(act * 3);
  int F = 0;
  for (auto& t : ps.layers) F = std::max(F, t.ffn1.N);
  ps.ffh.ensure((size_t)rows * std::max(F, 1) * 4);
  // x (B,T,D) channels-last, and xa = x + pe[b] (transformer.py:50-52 indexes the table by the BATCH axis)
  launch_transpose_in(enc_out, B, T, D, ps.xt.as<float>(), ps.pe, ps.xa.as<float>(), s);
  launch_vq_frames(ps.plan, ps.xt.as<float>(), rows, D, out_codes, ps.qg.as<float>(), s);
  if (out_quantized)
    for (int g = 0; g < 3; ++g)
      launch_transpose_out(ps.qg.as<float>() + (size_t)g * rows * D, B, T, D, out_quantized + (size_t)g * rows * D, s);
  // pre-LN transformer layers (transformer.py:86-151): x += MHA(LN1(x)); x += ffn_2(relu(conv_k5(LN2(x))))
  auto ln = [&](const float* w, const float* b, const void* x, void* y) {
    LnMod l;
    memset(&l, 0, sizeof(l));
    l.x = x; l.ldx = D; l.y = y; l.ldy = D; l.w = w; l.b = b; l.eps = 1e-5f; l.rows = rows; l.rows_per_batch = T; l.C = D;
    l.scale_plus_one = 1.f;
    launch_ln_mod(l, s);
  };
  for (auto& t : ps.layers) {
    ln(t.ln1w, t.ln1b, ps.xa.p, ps.xn.p);
    h->gemm(h->problem(t.qkv, ps.xn.p, D, B, T, T, ps.qkv.p, 3 * D, 0, EPI_NONE), t.qkv, false, s);
    launch_mha_fp32(ps.qkv.as<float>(), B, T, ps.heads, D / ps.heads, ps.att.as<float>(), s);
    TapGemm po = h->problem(t.out, ps.att.p, D, B, T, T, ps.xa.p, D, 0, EPI_RESID);
    po.resid_in = ps.xa.p;
    h->gemm(po, t.out, false, s);
    ln(t.ln2w, t.ln2b, ps.xa.p, ps.xn.p);
    h->gemm(h->problem(t.ffn1, ps.xn.p, D, B, T, T, ps.ffh.p, t.ffn1.N, 0, EPI_RELU), t.ffn1, false, s);
    TapGemm pf = h->problem(t.ffn2, ps.ffh.p, t.ffn1.N, B, T, T, ps.xa.p, D, 0, EPI_RESID);
    pf.resid_in = ps.xa.p;
    h->gemm(pf, t.ffn2, false, s);
  }
  ln(ps.last_w, ps.last_b, ps.xa.p, ps.xn.p);
  launch_mean_time(ps.xn.as<float>(), B, T, D, out_spk, s);
  FLM_API_END
}

extern "C" int flm_codec_decode(flm_codec_dec* h, const float* latents, const float* spk, int B, int L, float* out_wav,
                                flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && latents && spk && out_wav, "null argument");
  DeviceGuard dguard(h->ctx->device);
  if (B == 0 || L == 0) return FLM_OK;
  cudaStream_t s = S(stream);
  const int Cin = h->conv0.K, b16 = h->bf() ? 1 : 0;
  const size_t e = h->esize();
  // largest activation: elements per latent frame
  int64_t per_frame = std::max<int64_t>(h->conv0.N, Cin);
  {
    int64_t rate = 1;
    for (auto& blk : h->blocks) { rate *= blk.stride; per_frame = std::max<int64_t>(per_frame, rate * blk.cout); }
  }
  for (int k = 0; k < 3; ++k) h->buf[k].ensure((size_t)B * L * per_frame * e);
  h->style.ensure((size_t)B * 2 * Cin * 4);
  void *x = h->buf[0].p, *t1 = h->buf[1].p, *t2 = h->buf[2].p;
  // style = timbre_linear(spk); x = LN_noaffine(latents) * gamma + beta   (facodec.py:631-636)
  h->gemm(h->problem(h->timbre_linear, spk, Cin, 1, B, B, h->style.p, 2 * Cin, 0, EPI_NONE), h->timbre_linear, false, s);
  LnMod ln;
  memset(&ln, 0, sizeof(ln));
  ln.x = latents; ln.ldx = Cin; ln.y = t1; ln.ldy = Cin; ln.y_bf16 = b16;
  ln.scale = h->style.as<float>(); ln.shift = h->style.as<float>() + Cin; ln.mod_bstride = 2 * Cin; ln.scale_plus_one = 0.f;
  ln.eps = 1e-5f; ln.rows = (int64_t)B * L; ln.rows_per_batch = L; ln.C = Cin;
  launch_ln_mod(ln, s);
  h->gemm(h->problem(h->conv0, t1, Cin, B, L, L, x, h->conv0.N, b16, EPI_NONE), h->conv0, h->bf(), s);
  int T = L;
  for (auto& blk : h->blocks) {
    run_act(*h, blk.act, x, t1, B, T, s);
    // transposed conv: (B,T,cin) -> (B,T,s*cout) == (B,T*s,cout)
    h->gemm(h->problem(blk.up, t1, blk.cin, B, T, T, t2, blk.up.N, b16, EPI_NONE), blk.up, h->bf(), s);
    std::swap(x, t2);
    T *= blk.stride;
    for (int j = 0; j < 3; ++j) run_res_unit(*h, blk.ru[j], x, t1, t2, B, T, s);
  }
  run_act(*h, h->act_out, x, t1, B, T, s);
  launch_conv_out_tanh(t1, b16, h->wout, h->bout, B, T, h->cout_final, out_wav, s);
  FLM_API_END
}

extern "C" int flm_codec_dec_activation(flm_codec_dec* h, const char* prefix, const float* x, int B, int T, int C,
                                        float* y, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && prefix && x && y, "null argument");
  DeviceGuard dguard(h->ctx->device);
  auto it = h->acts_by_name.find(prefix);
  if (it == h->acts_by_name.end()) throw Error(FLM_ERR_ARG, std::string("no activation named ") + prefix);
  FLM_REQUIRE(it->second.C == C, "channel count mismatch");
  Act1d p;
  p.x = x; p.y = y; p.io_bf16 = 0; p.a = it->second.a; p.invb = it->second.invb;
  memcpy(p.fu, it->second.fu, sizeof(p.fu));
  memcpy(p.fd, it->second.fd, sizeof(p.fd));
  p.B = B; p.T = T; p.C = C; p.fast_sin = 0;
  launch_act1d(p, S(stream));
  FLM_API_END
}

// ---- encoder (fp32 FMA; runs once per distinct prompt)
struct flm_codec_enc : Engine {
  float *w_in, *b_in; int c_in;
  struct Block { ResUnitW ru[3]; ActW act; Layer down; int stride, pad; };
  std::vector<Block> blocks;
  ActW act_out;
  Layer conv_out;
  DevBuf buf[3];
  flm_codec_enc(flm_ctx* c) : Engine(c, FLM_F32) {}
};

extern "C" int flm_codec_enc_load(flm_ctx* ctx, const flm_tensor* weights, int n, flm_codec_enc** out) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && weights && out, "null argument");
  DeviceGuard dguard(ctx->device);
  WeightMap wm(weights, n);
  std::unique_ptr<flm_codec_enc> h(new flm_codec_enc(ctx));
  const flm_tensor& w0 = wm.get("block.0.weight_v");
  const int C0 = (int)w0.shape[0];
  FLM_REQUIRE(w0.ndim == 3 && w0.shape[1] == 1 && w0.shape[2] == 7, "block.0 must be a 1->C k=7 conv");
  {
    std::vector<float> w = fold_weight_norm(wm.vec("block.0.weight_g", {C0, 1, 1}), wm.vec("block.0.weight_v", {C0, 1, 7}), C0);
    std::vector<float> wt((size_t)7 * C0);
    for (int c = 0; c < C0; ++c)
      for (int t = 0; t < 7; ++t) wt[(size_t)t * C0 + c] = w[(size_t)c * 7 + t];
    h->w_in = h->store.upload(wt);
    h->b_in = h->store.upload(wm.vec("block.0.bias", {C0}));
    h->c_in = C0;
  }
  int c = C0, i = 1;
  while (wm.has("block." + std::to_string(i) + ".block.4.weight_v")) {
    const std::string p = "block." + std::to_string(i);
    const flm_tensor& wv = wm.get(p + ".block.4.weight_v");  // (2c, c, 2s)
    flm_codec_enc::Block blk;
    blk.stride = (int)wv.shape[2] / 2;
    blk.pad = blk.stride / 2 + blk.stride % 2;
    FLM_REQUIRE((int)wv.shape[1] == c, "encoder block channel mismatch");
    const int dils[3] = {1, 3, 9};
    for (int j = 0; j < 3; ++j) blk.ru[j] = load_res_unit(*h, wm, p + ".block." + std::to_string(j), c, dils[j], false);
    blk.act = load_act(h->store, wm, p + ".block.3", c);
    const int cn = (int)wv.shape[0];
    blk.down = load_wn_conv(*h, wm, p + ".block.4", cn, c, 2 * blk.stride, -blk.pad, 1, blk.stride, false);
    h->blocks.push_back(blk);
    c = cn;
    ++i;
  }
  FLM_REQUIRE(!h->blocks.empty(), "no encoder blocks found");
  const std::string pa = "block." + std::to_string(i), pc = "block." + std::to_string(i + 1);
  h->act_out = load_act(h->store, wm, pa, c);
  const flm_tensor& wo = wm.get(pc + ".weight_v");
  h->conv_out = load_wn_conv(*h, wm, pc, (int)wo.shape[0], c, 3, -1, 1, 1, false);
  *out = h.release();
  FLM_API_END
}
extern "C" void flm_codec_enc_destroy(flm_codec_enc* h) { delete h; }

extern "C" int64_t flm_codec_enc_frames(flm_codec_enc* h, int64_t S_) {
  if (!h) return -1;
  int64_t T = S_;
  for (auto& blk : h->blocks) {
    const int64_t num = T + 2 * blk.pad - 2 * blk.stride;
    if (num < 0) return 0;
    T = num / blk.stride + 1;
  }
  return T;
}

extern "C" int flm_codec_encode(flm_codec_enc* h, const float* wav, int B, int64_t S_, float* out, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(h && wav && out, "null argument");
  DeviceGuard dguard(h->ctx->device);
  const int64_t Tout = flm_codec_enc_frames(h, S_);
  FLM_REQUIRE(Tout > 0, "prompt too short for the encoder");
  FLM_REQUIRE(S_ < (1ll << 30), "prompt too long");
  if (B == 0) return FLM_OK;
  cudaStream_t s = S(stream);
  // largest activation: first stage S x C0 (channels double while T shrinks by >= 2)
  size_t mx = (size_t)S_ * h->c_in;
  {
    int64_t T = S_;
    for (auto& blk : h->blocks) {
      T = (T + 2 * blk.pad - 2 * blk.stride) / blk.stride + 1;
      mx = std::max(mx, (size_t)T * blk.down.N);
    }
  }
  for (int k = 0; k < 3; ++k) h->buf[k].ensure((size_t)B * mx * 4);
  void *x = h->buf[0].p, *t1 = h->buf[1].p, *t2 = h->buf[2].p;
  launch_conv_in_wav(wav, h->w_in, h->b_in, B, S_, h->c_in, static_cast<float*>(x), s);
  int T = (int)S_;
  for (auto& blk : h->blocks) {
    for (int j = 0; j < 3; ++j) run_res_unit(*h, blk.ru[j], x, t1, t2, B, T, s);
    run_act(*h, blk.act, x, t1, B, T, s);
    const int Tn = (T + 2 * blk.pad - 2 * blk.stride) / blk.stride + 1;
    h->gemm(h->problem(blk.down, t1, blk.down.K, B, T, Tn, t2, blk.down.N, 0, EPI_NONE), blk.down, false, s);
    std::swap(x, t2);
    T = Tn;
  }
  run_act(*h, h->act_out, x, t1, B, T, s);
  h->gemm(h->problem(h->conv_out, t1, h->conv_out.K, B, T, T, t2, h->conv_out.N, 0, EPI_NONE), h->conv_out, false, s);
  launch_transpose_out(static_cast<const float*>(t2), B, T, h->conv_out.N, out, s);
  FLM_API_END
}

// ===================================================================================== profiler
extern "C" const char* flm_profile_class_name(int kc);
extern "C" int flm_profile_enable(flm_ctx* ctx, int on) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx != nullptr, "null ctx");
  DeviceGuard dguard(ctx->device);
  FLM_CUDA(cudaDeviceSynchronize());
  for (auto& r : ctx->recs) { ctx->pool.push_back(r.a); ctx->pool.push_back(r.b); }
  ctx->recs.clear();
  ctx->prof_on = on != 0;
  FLM_API_END
}

extern "C" int flm_profile_read(flm_ctx* ctx, double* out, int n_classes) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && out && n_classes >= KC_COUNT, "bad arguments (need room for 8 classes x 4 doubles)");
  DeviceGuard dguard(ctx->device);
  FLM_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < n_classes * 4; ++i) out[i] = 0.0;
  for (auto& r : ctx->recs) {
    float ms = 0.f;
    FLM_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    out[r.kc * 4 + 0] += 1.0;
    out[r.kc * 4 + 1] += ms;
    out[r.kc * 4 + 2] += r.flops;
    out[r.kc * 4 + 3] += r.bytes;
  }
  FLM_API_END
}

// per-shape detail of the same records: one text line per (class, tag): "class|tag|launches|ms|flops|bytes"
extern "C" int flm_profile_detail(flm_ctx* ctx, char* out, int cap) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && out && cap > 0, "bad arguments");
  DeviceGuard dguard(ctx->device);
  FLM_CUDA(cudaDeviceSynchronize());
  struct Agg { double n = 0, ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : ctx->recs) {
    float ms = 0.f;
    FLM_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    Agg& a = agg[std::string(flm_profile_class_name(r.kc)) + "|" + r.tag];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
  }
  std::string s;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s|%.0f|%.4f|%.6e|%.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops,
             kv.second.bytes);
    s += line;
  }
  if ((int)s.size() + 1 > cap) throw Error(FLM_ERR_ARG, "flm_profile_detail: buffer too small");
  memcpy(out, s.c_str(), s.size() + 1);
  FLM_API_END
}

extern "C" const char* flm_profile_class_name(int kc) {
  static const char* names[KC_COUNT] = {"tapgemm_tcgen05", "tapgemm_fp32_fma", "ln_modulate", "dwconv31_stats",
                                        "groupnorm_finalize", "groupnorm_apply", "snake_act1d", "other"};
  return (kc >= 0 && kc < KC_COUNT) ? names[kc] : "?";
}

// ===================================================================================== test hook
extern "C" int flm_tapgemm_test(flm_ctx* ctx, int mode, const float* A, const float* W, const float* bias, int B,
                                int T_in, int T_out, int K, int N, int ntaps, int off0, int dil, int stride, int epi,
                                float* out, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && A && W && out, "null argument");
  FLM_REQUIRE(epi >= 0 && epi <= EPI_RELU, "epi must be 0..3");
  DeviceGuard dguard(ctx->device);
  cudaStream_t s = S(stream);
  TapGemm p;
  memset(&p, 0, sizeof(p));
  p.bias = bias; p.out = out; p.lda = K; p.ldc = N; p.B = B; p.T_in = T_in; p.T_out = T_out; p.K = K; p.N = N;
  p.ntaps = ntaps; p.off0 = off0; p.dil = dil; p.stride = stride; p.epi = epi; p.out_bf16 = 0;
  if (mode == FLM_F32) {
    p.A = A; p.W = W;
    launch_tapgemm_simt(p, s);
  } else {
    DevBuf a16, w16;
    const int64_t na = (int64_t)B * T_in * K, nw = (int64_t)ntaps * N * K;
    a16.ensure(na * 2); w16.ensure(nw * 2);
    launch_f32_to_bf16(A, a16.as<bf16>(), na, s);
    launch_f32_to_bf16(W, w16.as<bf16>(), nw, s);
    p.A = a16.p; p.W = w16.p;
    launch_tapgemm_tc(p, ctx->tma_encode, ctx->num_sms, s);
    FLM_CUDA(cudaStreamSynchronize(s));  // a16/w16 are freed on return
  }
  FLM_API_END
}

// bf16-in / bf16-out test hook for the tcgen05 kernels (gen 1 = tapgemm_tc.cu, 2 = tapgemm_tc2.cu).  All tensor
// pointers are device bf16 except bias / gate (fp32).  epi 0..3: out = act(v); 4: out = resid + v;
// 5: resid (in place) += gate[b,:] * (v + addend)   (addend nullable)
extern "C" int flm_tapgemm_test_bf16(flm_ctx* ctx, int gen, const void* A, const void* W, const float* bias, int B,
                                     int T_in, int T_out, int K, int N, int ntaps, int off0, int dil, int epi, void* out,
                                     void* resid, const void* addend, const float* gate, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && A && W, "null argument");
  DeviceGuard dguard(ctx->device);
  cudaStream_t s = S(stream);
  TapGemm p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.W = W; p.bias = bias; p.out = out; p.lda = K; p.ldc = N; p.B = B; p.T_in = T_in; p.T_out = T_out; p.K = K;
  p.N = N; p.ntaps = ntaps; p.off0 = off0; p.dil = dil; p.stride = 1; p.epi = epi; p.out_bf16 = 1;
  if (epi == EPI_RESID) p.resid_in = resid;
  if (epi == EPI_GATE_RESID) {
    p.hres = static_cast<float*>(resid); p.ld_res = N; p.hres_bf16 = 1; p.gate = gate; p.gate_bstride = N;
    if (addend) { p.addend = addend; p.ld_add = N; p.addend_bf16 = 1; }
  }
  if (gen >= 2) {
    FLM_REQUIRE(tapgemm_tc2_supported(p), "problem not supported by the CTA-pair kernel");
    launch_tapgemm_tc2(p, ctx->tma_encode, ctx->num_sms, s);
  } else {
    launch_tapgemm_tc(p, ctx->tma_encode, ctx->num_sms, s);
  }
  FLM_API_END
}

// ===================================================================================== generic bf16 ops
// Building blocks for callers next to the hot path (the prior generator's FFT stacks): the same
// tcgen05 implicit-conv GEMM and row LayerNorm the denoiser uses, on caller-owned bf16 tensors.
extern "C" int flm_conv1d_bf16(flm_ctx* ctx, const void* A, const void* W, const float* bias, int B, int T, int K, int N,
                               int ntaps, int off0, int dil, int epi, void* out, const void* resid, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && A && W && out, "null argument");
  FLM_REQUIRE(epi >= EPI_NONE && epi <= EPI_RESID, "epi must be 0..4");
  FLM_REQUIRE(epi != EPI_RESID || resid != nullptr, "epi 4 needs resid");
  DeviceGuard dguard(ctx->device);
  if ((int64_t)B * T == 0) return FLM_OK;
  TapGemm p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.W = W; p.bias = bias; p.out = out; p.lda = K; p.ldc = N; p.B = B; p.T_in = T; p.T_out = T; p.K = K; p.N = N;
  p.ntaps = ntaps; p.off0 = off0; p.dil = dil; p.stride = 1; p.epi = epi; p.out_bf16 = 1; p.resid_in = resid;
  FLM_REQUIRE(tapgemm_tc_supported(p), "conv1d_bf16: need K % 64 == 0, N % 64 == 0 and 16-byte aligned operands");
  const double M = (double)B * T;
  char tag[96];
  tag[0] = 0;
  if (ctx->prof_on) snprintf(tag, sizeof(tag), "K%d N%d taps%d epi%d B%d T%d", K, N, ntaps, epi, B, T);
  ProfScope ps(ctx, KC_GEMM_TC, S(stream), 2.0 * M * N * K * ntaps, M * K * 2 + (double)ntaps * N * K * 2 + M * N * 2, tag);
  if (tapgemm_tc2_supported(p)) launch_tapgemm_tc2(p, ctx->tma_encode, ctx->num_sms, S(stream));
  else launch_tapgemm_tc(p, ctx->tma_encode, ctx->num_sms, S(stream));
  FLM_API_END
}

// y = LayerNorm(x; w, b, eps) row-wise over C, rows with zero_rows[r] != 0 written as zeros; x, y bf16 (rows, C)
extern "C" int flm_layernorm_bf16(flm_ctx* ctx, const void* x, const float* w, const float* b, float eps, int64_t rows,
                                  int C, const uint8_t* zero_rows, void* y, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && x && y, "null argument");
  DeviceGuard dguard(ctx->device);
  if (rows == 0) return FLM_OK;
  LnMod ln;
  memset(&ln, 0, sizeof(ln));
  ln.x = x; ln.ldx = C; ln.x_bf16 = 1; ln.y = y; ln.ldy = C; ln.y_bf16 = 1; ln.w = w; ln.b = b; ln.scale_plus_one = 1.f;
  ln.eps = eps; ln.rows = rows; ln.rows_per_batch = (int)std::min<int64_t>(rows, 1 << 30); ln.C = C; ln.zero_rows = zero_rows;
  ProfScope ps(ctx, KC_LN_MOD, S(stream), (double)rows * C * 8, (double)rows * C * 4);
  launch_ln_mod(ln, S(stream));
  FLM_API_END
}

// o = softmax(q k^T / sqrt(d), keys < key_lens[b]) v per head: the self-attention of the FFT decoder blocks
// (SubLayers.py:29-57, Modules.py:14-25) on the fused QKV projection's output
extern "C" int flm_attention_bf16(flm_ctx* ctx, const void* qkv, const int32_t* key_lens, int B, int seq, int H, int dh,
                                  void* out, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && qkv && key_lens && out, "null argument");
  FLM_REQUIRE(dh == 32, "attention_bf16: head dim must be 32");
  DeviceGuard dguard(ctx->device);
  // 4 B S^2 H d FLOP (QK^T and PV); the op is softmax-bound, profiled under `other`
  ProfScope ps(ctx, KC_OTHER, S(stream), 4.0 * B * (double)seq * seq * H * dh, 4.0 * B * (double)seq * H * dh * 2, "attention");
  launch_attn_prefix(static_cast<const bf16*>(qkv), key_lens, B, seq, H, dh, static_cast<bf16*>(out), S(stream));
  FLM_API_END
}

// ---- micro-benchmark hook: `reps` back-to-back launches of one tap-GEMM on pseudo-random operands,
// timed with CUDA events on `stream`; *out_ms = average ms per launch.  epi 0..3, or 5 (gated residual).
// Operands are pseudo-random (zeros would under-state the power draw and over-state the clocks).
extern "C" int flm_tapgemm_bench(flm_ctx* ctx, int mode, int B, int T, int K, int N, int ntaps, int dil, int epi,
                                 int out_bf16, int reps, float* out_ms, flm_stream stream) {
  FLM_API_BEGIN
  FLM_REQUIRE(ctx && out_ms && reps > 0, "bad arguments");
  DeviceGuard dguard(ctx->device);
  cudaStream_t s = S(stream);
  const size_t e = mode == FLM_BF16 ? 2 : 4;
  const int64_t M = (int64_t)B * T;
  // test-only flag bits above the epilogue id: 0x100 = with addend (bf16 in bf16 mode), 0x200 = bf16 residual stream
  const int with_addend = (epi >> 8) & 1, hres16 = (epi >> 9) & 1;
  epi &= 0xff;
  DevBuf a, w, o, h, g, bias, add;
  if (with_addend) {
    add.ensure(M * N * e);
    launch_fill_random(add.p, mode == FLM_BF16, M * N, 7u, 1.0f, s);
  }
  a.ensure(M * K * e); w.ensure((size_t)ntaps * N * K * e); o.ensure(M * N * 4); h.ensure(M * N * 4);
  g.ensure((size_t)B * N * 4); bias.ensure((size_t)N * 4);
  launch_fill_random(a.p, mode == FLM_BF16, M * K, 1u, 1.0f, s);
  launch_fill_random(w.p, mode == FLM_BF16, (int64_t)ntaps * N * K, 2u, 0.03f, s);
  launch_fill_random(h.p, hres16, M * N, 3u, 1.0f, s);
  launch_fill_random(g.p, 0, (int64_t)B * N, 4u, 0.1f, s);
  launch_fill_random(bias.p, 0, N, 5u, 0.1f, s);
  TapGemm p;
  memset(&p, 0, sizeof(p));
  p.A = a.p; p.W = w.p; p.bias = bias.as<float>(); p.out = o.p; p.lda = K; p.ldc = N; p.B = B; p.T_in = T; p.T_out = T;
  p.K = K; p.N = N; p.ntaps = ntaps; p.off0 = -(ntaps / 2) * dil; p.dil = dil; p.stride = 1; p.epi = epi;
  p.out_bf16 = out_bf16;
  p.gate = g.as<float>(); p.gate_bstride = N; p.hres = h.as<float>(); p.ld_res = N;
  p.hres_bf16 = hres16;
  if (with_addend) { p.addend = add.p; p.ld_add = N; p.addend_bf16 = mode == FLM_BF16; }
  if (epi == EPI_RESID) { p.resid_in = o.p; }
  cudaEvent_t e0, e1;
  FLM_CUDA(cudaEventCreate(&e0));
  FLM_CUDA(cudaEventCreate(&e1));
  auto launch = [&]() {
    if (mode == FLM_BF16) {
      if (tapgemm_tc2_supported(p)) launch_tapgemm_tc2(p, ctx->tma_encode, ctx->num_sms, s);
      else launch_tapgemm_tc(p, ctx->tma_encode, ctx->num_sms, s);
    } else {
      launch_tapgemm_simt(p, s);
    }
  };
  for (int i = 0; i < 3; ++i) launch();
  FLM_CUDA(cudaEventRecord(e0, s));
  for (int i = 0; i < reps; ++i) launch();
  FLM_CUDA(cudaEventRecord(e1, s));
  FLM_CUDA(cudaStreamSynchronize(s));
  float ms = 0.f;
  FLM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *out_ms = ms / reps;
  FLM_API_END
}
