/*
 * flamed_b200.h - C ABI of the B200-native Flamed-TTS inference hot path.
 *
 * The reference (nghiahuynh-ai/Flamed-TTS) is pure Python/PyTorch and has no FFI of
 * its own; the boundary a maintainer would bind is the set of Python methods listed
 * below.  Each entry point replaces the body of one of them ("replaces:" lines cite
 * /root/reference paths).  The Python side (ctypes stub in INTEGRATION.md, product
 * binding in flamed_tts_b200/_lib.py) passes raw device pointers of torch tensors and
 * torch's current CUDA stream.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ or torch types.
 *   - every function returns 0 on success, <0 on error; flm_last_error() returns a
 *     thread-local message.  No exceptions cross the boundary.
 *   - `*_load` takes HOST fp32 tensors named by the reference's state-dict keys
 *     (relative to the module), copies and re-packs them into library-owned device
 *     memory (weight-norm folded, conv taps split, bf16 copies for the tensor-core
 *     path, adaLN projections concatenated); the caller may free its tensors after.
 *   - run functions take DEVICE pointers owned by the caller, only enqueue work on
 *     `stream` and never synchronise, except flm_lr_plan (one documented sync; the
 *     reference syncs at the same place, pva.py:158).
 *   - activations are channels-last: (B, T, C) with C contiguous.
 *   - a handle is NOT thread-safe and is bound to one stream at a time: its workspaces and CUDA-graph
 *     cache (keyed by (B, L, nfe), bounded, LRU) are reused by every call, so two calls on the same
 *     handle must be ordered on one stream (or by events).  Different handles may be driven from
 *     different threads / streams.  Every entry point restores the caller's current CUDA device.
 *   - there is no CPU fallback: every entry point fails with FLM_ERR_CUDA if no
 *     sm_100 device is present.
 */
#ifndef FLAMED_B200_H
#define FLAMED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLM_OK 0
#define FLM_ERR_ARG (-1)
#define FLM_ERR_CUDA (-2)
#define FLM_ERR_WEIGHT (-3)
#define FLM_ERR_UNSUPPORTED (-4)

/* arithmetic mode of a module handle */
#define FLM_F32 0  /* fp32 storage, fp32 FMA everywhere (parity mode)                    */
#define FLM_BF16 1 /* bf16 GEMM operands on tcgen05 tensor cores, fp32 accumulate/state   */

typedef struct flm_ctx flm_ctx;
typedef struct flm_durgen flm_durgen;
typedef struct flm_denoiser flm_denoiser;
typedef struct flm_codec_dec flm_codec_dec;
typedef struct flm_codec_enc flm_codec_enc;
typedef struct flm_comm flm_comm;
typedef void* flm_stream; /* cudaStream_t */

typedef struct {
  const char* name;  /* reference state-dict key, relative to the module            */
  const float* data; /* HOST pointer, fp32, contiguous                              */
  int32_t ndim;
  int64_t shape[4];
} flm_tensor;

const char* flm_last_error(void);
int flm_version(void);
/* kernels launched by this library in this process so far (graph replays count their kernel nodes) */
unsigned long long flm_launch_count(void);

int flm_ctx_create(int device, flm_ctx** out);
void flm_ctx_destroy(flm_ctx* ctx);

/* ---------------------------------------------------------------- duration / silence generators
 * replaces: PVA.sample loop + rounding, flamed/models/synthesizer/pva.py:88-112
 *           (ProbabilisticModule.forward pva.py:221-238, TimeEmbedding pva.py:9-41).
 * weights: keys relative to `prior_generator.pva.` ("duration_generator.proj.weight", ...).
 * Always fp32 FMA (durations must round identically to the reference). */
int flm_durgen_load(flm_ctx* ctx, const flm_tensor* weights, int n, flm_durgen** out);
void flm_durgen_destroy(flm_durgen* h);
/* enc (B,P,192) f32; noise_* (B,P) f32 standard normal (dur first, pva.py:101-102), or BOTH NULL: the two draws are
 * then made inside the library from `seed` with the documented counter-based map below (tensor ids 0 and 1);
 * src_mask (B,P) u8, 1 = padding; ts (nfe+1) f32 HOST = torch.linspace(0,1,nfe+1);
 * out_phone/out_sil (B,P) f32 holding clamp(round(exp(x)-1),0); out_dur_t/out_sil_t (B,P) f32
 * = the ODE state before rounding (may be NULL).
 *
 * Seed -> tensor map (flm_philox_normal returns exactly these values): for flat element index i of a tensor,
 *   g = i >> 2;  (r0,r1,r2,r3) = Philox4x32-10(counter = (g & 0xffffffff, g >> 32, tensor_id, 0),
 *                                              key = (seed & 0xffffffff, seed >> 32));
 *   u_k = ((r_k >> 8) + 0.5) * 2^-24;  z0 = sqrt(-2 ln u0) cos(2 pi u1), z1 = sqrt(-2 ln u0) sin(2 pi u1),
 *   z2, z3 likewise from (u2, u3);  noise[i] = z[i & 3].
 * tensor_id: 0 = duration noise (B,P), 1 = silence noise (B,P), 2 = latent noise (B,L,D). */
int flm_durgen_sample(flm_durgen* h, const float* enc, const float* noise_dur, const float* noise_sil, uint64_t seed,
                      const uint8_t* src_mask, const float* ts_host, int nfe, float temperature, int B, int P,
                      float* out_phone, float* out_sil, float* out_dur_t, float* out_sil_t, flm_stream stream);
/* replaces: ProbabilisticModule.forward, pva.py:221-238 - one vector-field evaluation of generator `which`
 * (0 = duration_generator, 1 = sil_generator): x (B,P) f32, enc (B,P,192) f32, scalar t, src_mask (B,P) u8 1 = padding
 * (nullable) -> out_v (B,P) f32 (masked positions 0). */
int flm_durgen_forward(flm_durgen* h, int which, const float* x, const float* enc, float t, const uint8_t* src_mask,
                       int B, int P, float* out_v, flm_stream stream);
/* n values of the map above (tests / documentation of the map) */
int flm_philox_normal(flm_ctx* ctx, uint64_t seed, int tensor_id, int64_t n, float* out, flm_stream stream);

/* ---------------------------------------------------------------- length regulator
 * replaces: LengthRegulator.LR, pva.py:125-166 (+ pad, flamed/utils/tools.py:299-317).
 * plan: integer repeats -> inclusive cumsum (B,2P) i32 + tgt_len (B) i64 on device, and the
 * batch maximum on the host (one stream sync, as pva.py:158 `.tolist()`).  out_tmax_host == NULL: no
 * synchronisation, the caller reads tgt_len itself later (metadata path: every front batch is planned first,
 * one sync for all of them, then the utterances are re-bucketed by their real frame counts). */
int flm_lr_plan(flm_ctx* ctx, const float* phone_dur, const float* sil_dur, const int64_t* src_lens, int B, int P,
                int32_t* out_cumsum, int64_t* out_tgt_len, int64_t* out_tmax_host, flm_stream stream);
/* expand: out (B,Tmax,H) f32 = gather of x (B,P,H) f32, zero beyond tgt_len;
 * out_index (B,Tmax) i32 = source phoneme row, -1 for padding (may be NULL). */
int flm_lr_expand(flm_ctx* ctx, const float* x, const int32_t* cumsum, int B, int P, int H, int Tmax, float* out,
                  int32_t* out_index, flm_stream stream);

/* the same expand for a batch whose B samples come from different planned batches: x_rows[b] -> (P[b],H) f32 rows,
 * cumsums[b] -> (2 P[b]) i32 of sample b (device arrays of device pointers); frames >= the sample's total are zero
 * (the re-padding of tools.py:299-317 to the new batch maximum Tmax). */
int flm_lr_expand_gather(flm_ctx* ctx, const float* const* x_rows, const int32_t* const* cumsums,
                         const int32_t* P_per_sample, int B, int H, int Tmax, float* out, int32_t* out_index,
                         flm_stream stream);

/* ---------------------------------------------------------------- code-decoder denoiser
 * replaces: ProbGenerator.sample, flamed/models/synthesizer/prob_generator.py:434-446
 *           (QuantizerEncoding 375-381, ConditionDownSampler 198-205, SimpleMLPAdaLN 349-365).
 * weights: keys relative to `prob_generator.`; cfg taken from configs/prob.yaml. */
typedef struct {
  int32_t target_dim, spk_dim, cond_dim, hidden_dim, n_layers, n_quantizers, kernel_size, downsampling_stages;
} flm_prob_cfg;
int flm_denoiser_load(flm_ctx* ctx, const flm_tensor* weights, int n, const flm_prob_cfg* cfg, int mode,
                      flm_denoiser** out);
void flm_denoiser_destroy(flm_denoiser* h);
/* prior_embs (B,Q,L,cond_dim) f32; mask (B,L) u8, 1 = valid frame; out_cond (B,L,target_dim) f32 */
int flm_cond_prepare(flm_denoiser* h, const float* prior_embs, const uint8_t* mask, int B, int L, float* out_cond,
                     flm_stream stream);
/* cond (B,L,D) f32; spk (B,spk_dim) f32; noise (B,L,D) f32 standard normal (prob_generator.py:440) or NULL: the
 * draw is then fused into the x0 = temperature * noise + cond kernel from `seed` (tensor id 2 of the map above);
 * ts (nfe+1) f32 HOST; out_latents (B,L,D) f32 channels-last: the reference returns its
 * transpose(1,2) VIEW (prob_generator.py:446).  use_graph != 0: the whole nfe-step loop runs as one CUDA graph;
 * a (B,L,nfe) key is launched directly the first time it is seen, captured when it comes back, and at most 8
 * executables are kept per handle (least recently used first out). */
int flm_denoiser_sample(flm_denoiser* h, const float* cond, const float* spk, const float* noise, uint64_t seed,
                        const float* ts_host, int B, int L, int nfe, float temperature, float* out_latents,
                        int use_graph, flm_stream stream);
/* one velocity evaluation v = denoiser(x, t, spk) (SimpleMLPAdaLN.forward), for parity tests */
int flm_denoiser_forward(flm_denoiser* h, const float* x, const float* spk, float t, int B, int L, float* out_v,
                         flm_stream stream);
/* number of kernels launched per Euler step (for bench.py's gpu_launches) */
int flm_denoiser_launches_per_step(flm_denoiser* h);

/* ---------------------------------------------------------------- FaCodec decoder
 * replaces: FACodecDecoder.inference, flamed/models/facodec/facodec.py:630-638 (model stack
 *           400-415, DecoderBlock 246-265, ResidualUnit 121-133, Activation1d act.py:24-29).
 * weights: keys of FACodecDecoder.state_dict() (`model.*`, `timbre_linear.*`; others ignored). */
int flm_codec_dec_load(flm_ctx* ctx, const flm_tensor* weights, int n, int mode, flm_codec_dec** out);
void flm_codec_dec_destroy(flm_codec_dec* h);
/* latents (B,L,256) f32 channels-last; spk (B,256) f32; out_wav (B,1,200*L) f32 */
int flm_codec_decode(flm_codec_dec* h, const float* latents, const float* spk, int B, int L, float* out_wav,
                     flm_stream stream);
/* one anti-aliased Snake activation (Activation1d) of the decoder, by state-dict prefix
 * (e.g. "model.5"); x,y (B,T,C) f32 channels-last.  For parity tests. */
int flm_codec_dec_activation(flm_codec_dec* h, const char* prefix, const float* x, int B, int T, int C, float* y,
                             flm_stream stream);

/* replaces: FACodecDecoder.forward(vq=True), facodec.py:509-533 (quantize 470-507: three residual VQs of factorised
 *           layers, quantize/fvq.py:35-116, quantize/rvq.py:27-73; timbre transformer transformer.py:86-234) - the
 *           prompt side.  Needs the `quantizer.*` and `timbre_encoder.*` keys at load time.
 * enc_out (B,256,T) f32 in the REFERENCE layout (the encoder's output); out_codes (n_q = 6, B, T) i64;
 * out_quantized (3, B, 256, T) f32 = summed quantised vectors of the prosody / content / residual groups (nullable);
 * out_spk (B,256) f32 = mean over time of the timbre encoder's output.  fp32 FMA throughout. */
int flm_codec_dec_prompt(flm_codec_dec* h, const float* enc_out, int B, int T, int64_t* out_codes, float* out_quantized,
                         float* out_spk, flm_stream stream);

/* ---------------------------------------------------------------- FaCodec encoder (prompt)
 * replaces: FACodecEncoder.forward, facodec.py:215-217 (ctor 183-213, EncoderBlock 136-155). */
int flm_codec_enc_load(flm_ctx* ctx, const flm_tensor* weights, int n, flm_codec_enc** out);
void flm_codec_enc_destroy(flm_codec_enc* h);
/* returns the number of output frames for S input samples (<=0: too short) */
int64_t flm_codec_enc_frames(flm_codec_enc* h, int64_t S);
/* wav (B,1,S) f32; out (B,256,T') f32 in the REFERENCE layout (channels-first) */
int flm_codec_encode(flm_codec_enc* h, const float* wav, int B, int64_t S, float* out, flm_stream stream);

/* ---------------------------------------------------------------- waveform write-out and the final gather
 * replaces: the per-item `wav_tensor[0].detach().cpu().numpy()` + `sf.write(path, wav, 16000)` of
 *           /root/reference/synthesize.py:293-298 (soundfile stores PCM_16: lrintf(x * 32767)).
 * flm_wav_to_pcm16: n fp32 samples -> n int16 samples on the device, so that half the bytes cross PCIe / NVLink. */
int flm_wav_to_pcm16(flm_ctx* ctx, const float* wav, int64_t n, int16_t* out, flm_stream stream);
/* The path's only collective (one process per GPU, utterances sharded, no exchange inside the loops): gather of the
 * PCM waveforms to `root` over NCCL send/recv (NVLink).  flm_comm_unique_id fills 128 bytes (ncclUniqueId) on one rank;
 * the caller ships them to every rank (torch.distributed broadcast) and each rank calls flm_comm_create.
 * flm_gather_wav: rank r contributes counts[r] int16 samples (counts: HOST array, world entries, same on all ranks;
 * counts[rank] == n_send); on the root they land back to back in rank order in `recv` (>= sum(counts) samples; NULL
 * elsewhere).  Enqueues on `stream`, never synchronises. */
int flm_comm_unique_id(unsigned char* out128);
int flm_comm_create(flm_ctx* ctx, const unsigned char* id128, int world, int rank, flm_comm** out);
void flm_comm_destroy(flm_comm* c);
int flm_gather_wav(flm_comm* c, const int16_t* send, int64_t n_send, int16_t* recv, const int64_t* counts, int root,
                   flm_stream stream);

/* ---------------------------------------------------------------- per-launch profiler (bench.py roofline)
 * While enabled, every kernel launch made outside a CUDA-graph capture is bracketed by CUDA events on
 * the launching stream.  flm_profile_read synchronises and fills out[class*4 + {0,1,2,3}] =
 * {launch count, total ms, algorithmic FLOPs, algorithmic bytes} for the 8 kernel classes. */
int flm_profile_enable(flm_ctx* ctx, int on);
int flm_profile_read(flm_ctx* ctx, double* out, int n_classes);
const char* flm_profile_class_name(int kclass);
/* per-shape detail of the same records, one text line per (class, tag):
 * "class|tag|launches|ms|flops|bytes\n" (tag of a tap-GEMM: "K.. N.. taps.. epi.. B.. T..") */
int flm_profile_detail(flm_ctx* ctx, char* out, int cap);

/* ---------------------------------------------------------------- generic kernels exposed for tests
 * out[b,t,n] = epi(sum_tap sum_k A[b, t*stride + off0 + tap*dil, k] * W[tap][n][k] + bias[n]),
 * rows outside [0,T_in) read as zero.  A (B,T_in,K) f32, W (ntaps,N,K) f32, out (B,T_out,N) f32.
 * mode FLM_F32 runs the fp32 FMA kernel, FLM_BF16 converts operands to bf16 and runs the
 * tcgen05/TMA kernel (stride must be 1).  epi: 0 none, 1 gelu(erf), 2 silu, 3 relu. */
int flm_tapgemm_test(flm_ctx* ctx, int mode, const float* A, const float* W, const float* bias, int B, int T_in,
                     int T_out, int K, int N, int ntaps, int off0, int dil, int stride, int epi, float* out,
                     flm_stream stream);

/* ---------------------------------------------------------------- generic bf16 building blocks
 * The denoiser's tcgen05 implicit-conv GEMM and row LayerNorm on caller-owned device bf16 tensors, for the callers
 * either side of the hot path (the prior generator's FFT decoder stacks, Models.py:107-171 / SubLayers.py:8-93).
 * flm_conv1d_bf16: out[b,t,n] = epi(sum_tap sum_k A[b,t+off0+tap*dil,k] W[tap][n][k] + bias[n]); A (B,T,K),
 * W (ntaps,N,K), out/resid (B,T,N) bf16; bias (N) f32; K % 64 == 0, N % 64 == 0; epi 0 none, 1 gelu, 2 silu, 3 relu,
 * 4 out = resid + v.  flm_layernorm_bf16: y = LN(x; w, b, eps) over the last dim (C % 128 == 0, C <= 1024), rows with
 * zero_rows[r] != 0 (nullable) are written as zeros - the masked_fill(pad, 0) that follows every FFT sub-layer. */
int flm_conv1d_bf16(flm_ctx* ctx, const void* A, const void* W, const float* bias, int B, int T, int K, int N, int ntaps,
                    int off0, int dil, int epi, void* out, const void* resid, flm_stream stream);
int flm_layernorm_bf16(flm_ctx* ctx, const void* x, const float* w, const float* b, float eps, int64_t rows, int C,
                       const uint8_t* zero_rows, void* y, flm_stream stream);
/* replaces: MultiHeadAttention / ScaledDotProductAttention of the FFT decoder blocks (SubLayers.py:29-57,
 * Modules.py:14-25): out[b,s,h,:] = softmax_k(q.k / sqrt(dh), k < key_lens[b]) v.  qkv (B,S,3,H,dh) bf16 = the
 * output of the fused q|k|v projection, key_lens (B) i32 device (the key-padding masks of this model are prefix
 * masks: get_mask_from_lengths), out (B,S,H*dh) bf16; dh must be 32.  Scores never leave registers
 * (flash-attention dataflow); key tiles beyond a sample's prefix are skipped. */
int flm_attention_bf16(flm_ctx* ctx, const void* qkv, const int32_t* key_lens, int B, int S, int H, int dh, void* out,
                       flm_stream stream);

/* bf16-in / bf16-out form of the same problem on the tcgen05 kernels: gen 1 = single-CTA kernel, 2 = CTA-pair
 * (cta_group::2) kernel with the TMA epilogue.  A, W, out, resid, addend are device bf16; bias (N) and gate (B,N)
 * are f32.  epi 0..3 as above; 4: out = resid + v (codec skip connection, facodec.py:131-133);
 * 5: resid += gate[b,:] * (v + addend) in place (adaLN-gated residual, prob_generator.py:162-163; addend nullable). */
int flm_tapgemm_test_bf16(flm_ctx* ctx, int gen, const void* A, const void* W, const float* bias, int B, int T_in,
                          int T_out, int K, int N, int ntaps, int off0, int dil, int epi, void* out, void* resid,
                          const void* addend, const float* gate, flm_stream stream);

/* micro-benchmark hook: `reps` launches of one tap-GEMM (pseudo-random operands) timed with CUDA events on
 * `stream`; *out_ms = average ms per launch.  A (B,T,K), W (ntaps,N,K); epi 0..3 or 5 (gated residual). */
int flm_tapgemm_bench(flm_ctx* ctx, int mode, int B, int T, int K, int N, int ntaps, int dil, int epi, int out_bf16,
                      int reps, float* out_ms, flm_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* FLAMED_B200_H */
