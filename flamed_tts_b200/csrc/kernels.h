// Host-side launchers of every kernel in the library (all enqueue on `stream`, never sync).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace flm {

// ---- GEMMs
void launch_tapgemm_simt(const TapGemm& p, cudaStream_t stream);
// tcgen05 + TMA path: A/W bf16, K % 64 == 0, N % 64 == 0, stride == 1.  `tma_encode` is the
// driver's cuTensorMapEncodeTiled entry point (resolved once per context).
void launch_tapgemm_tc(const TapGemm& p, void* tma_encode, int num_sms, cudaStream_t stream);
bool tapgemm_tc_supported(const TapGemm& p);
void tapgemm_tc_init();
// second generation: CTA pairs (cta_group::2) + TMA epilogue, bf16 outputs / bf16 residual stream only
void launch_tapgemm_tc2(const TapGemm& p, void* tma_encode, int num_sms, cudaStream_t stream);
bool tapgemm_tc2_supported(const TapGemm& p);
void tapgemm_tc2_init();
int tapgemm_tc2_rowstat_parts(const TapGemm& p, int num_sms);  // partials per row written when p.rowstat != null
void kernels_norm_init();

// ---- row-wise LayerNorm (+ adaLN modulate): one warp per row, C % 128 == 0, C <= 1024
// y = LN(x; w, b, eps) * (scale_plus_one + scale[bi]) + shift[bi],  bi = row / rows_per_batch
struct LnMod {
  const void* x; int64_t ldx; int x_bf16;  // input rows (fp32, or bf16 when x_bf16)
  void* y; int64_t ldy; int y_bf16;
  const float* w; const float* b;  // affine (nullable)
  const float* shift; const float* scale; int64_t mod_bstride;  // per-sample modulation (nullable)
  float scale_plus_one;  // 1 for adaLN `x*(1+scale)+shift`, 0 for plain `x*gamma+beta`
  float eps;
  int64_t rows; int rows_per_batch; int C;
  int relu_in;  // apply ReLU to x before the statistics (durgen: LN(ReLU(conv)))
  const uint8_t* zero_rows;  // nullable (rows): non-zero -> the output row is all zeros (FFT blocks: masked_fill of padding)
};
void launch_ln_mod(const LnMod& p, cudaStream_t stream);

// ---- depthwise conv k=31 over time (channels-last) + per-(b,chunk,c) partial statistics: fp32 parity mode (the bf16
//      mode runs the LayerNorm-fused kernel of dwconv_fused.cu); also the parameter block of launch_dw_merge
struct DwConv {
  const void* x; void* y; int io_bf16;  // (B,L,C)
  const float* w;   // (KW, C) tap-major
  const float* bias;  // (C)
  float* part;      // (B, nchunk, C, 2) = (mean, M2) of each chunk of DW_TT outputs
  int B, L, C, KW;
  // optional fused GroupNorm(C,C) finalisation (scale != null): the last block of each (sample, 256-channel
  // block) merges the partials into scale = gamma*rstd, offset = beta - mean*scale, both (B, C)
  const float* gamma; const float* beta; float eps;
  float* scale; float* offset;
  int* counters;    // (B, C/256) arrival tickets, zero before the first launch, re-armed by the kernel
};
constexpr int DW_TT = 32;
inline int dw_nchunk(int L) { return (L + DW_TT - 1) / DW_TT; }
void launch_dwconv(const DwConv& p, cudaStream_t stream);
// merge of the per-chunk partials (B, nchunk, C, 2) into the GroupNorm(C,C) scale / offset (bf16 mode, after launch_dwconv_ln)
void launch_dw_merge(const DwConv& p, cudaStream_t stream);

// ---- LayerNorm + modulate fused into the depthwise conv (dwconv_fused.cu, bf16 mode with a bf16 residual stream):
//      u = LN(h)*(1+scale)+shift, d = dwconv31(u)+bias (+ per-chunk statistics of d for the GroupNorm that follows)
struct DwFused {
  const bf16* h;  // (B,L,C) residual stream = LayerNorm input
  bf16* u;        // (B,L,C) out: modulated LayerNorm output (inner residual of the ConvNeXt block)
  bf16* g;        // (B,L,C) out: depthwise-conv output d (normalised in place afterwards: input of conv_2)
  const float* rowstat; int parts;  // (B*L, parts) float2 (sum, sumsq) partials of the rows of h (TapGemm::rowstat)
  const float* ln_w; const float* ln_b;  // LayerNorm affine (nullable: FinalLayer's LN has none)
  const float* shift; const float* scale; int64_t mod_bstride;  // adaLN modulation of sample b (nullable)
  float ln_eps;
  const float* w;     // (31, C) tap-major depthwise weights
  const float* wsum;  // (C) sum over the taps
  const float* bias;  // (C)
  const float* gamma; const float* beta; float gn_eps;  // GroupNorm affine
  int B, L, C;
  void* tma_encode;
  // optional by-products for conv_3's epilogue (TapGemm::lnu_rowconst / lnu_table), all null or all set together with
  // u == nullptr: the inner residual u is then recomputed there instead of being written here
  float* rowconst_out;   // (B*L) float2 (rstd, -mean * rstd)
  float* lnu_out;        // (B, lnu_vecs, C): gate, gate * A, gate * (bias3 + B) [, A2]
  const float* gate;     // adaLN gate of the block, sample stride mod_bstride
  const float* bias3;    // conv_3's bias (C)
  // optional 4th vector A2 = ln2_w * (1 + scale2_b): column scale of the NEXT LayerNorm (the MLP branch's), which conv_3's
  // epilogue applies to its second output (TapGemm::out2); lnu_vecs = 4 then, else 3
  int lnu_vecs;
  const float* ln2_w;    // (C), nullable (no affine)
  const float* scale2;   // adaLN scale of the next LayerNorm, sample stride mod_bstride
};
void launch_dwconv_ln(const DwFused& p, float* part, int num_sms, cudaStream_t stream);
bool dwconv_fused_supported(const DwFused& p);
void dwconv_fused_init();

// tensor-core form of the same front half (dwconv_tc.cu): also merges the GroupNorm statistics into scale / offset (B, C).
// Scratch: rowconst (B*L float2), ab (B*C float2), part (dwconv_tc_part_bytes)
void launch_dwconv_tc(const DwFused& p, float* rowconst, float* ab, float* part, float* scale, float* offset,
                      const float* gate, const float* bias3, float* lnu, int num_sms, cudaStream_t stream);
bool dwconv_tc_supported(const DwFused& p);
size_t dwconv_tc_part_bytes(int B, int L, int C);
void dwconv_tc_init();

// ---- generic grouped statistics over (rows x channels-in-group) for GroupNorm(G) in the cond
//      down-sampler: partials (B, nchunk, G, 2) from chunks of GS_ROWS rows
constexpr int GS_ROWS = 32;
inline int gs_nchunk(int L) { return (L + GS_ROWS - 1) / GS_ROWS; }
void launch_group_stats(const void* x, int x_bf16, int B, int L, int C, int G, float* part, cudaStream_t stream);

// ---- merge partials (Chan) -> per-(b,c) scale = gamma*rstd, offset = beta - mean*scale
// chunk_rows = rows per chunk, group_size = channels per group (1 for GroupNorm(C,C))
void launch_gn_finalize(const float* part, int B, int L, int C, int G, int nchunk, int chunk_rows,
                        const float* gamma, const float* beta, float eps, float* scale, float* offset,
                        cudaStream_t stream);

// ---- y = x*scale[b,c] + offset[b,c], streaming (ConvNeXt GroupNorm apply after the fused finalisation)
void launch_gn_stream(const void* x, void* y, int io_bf16, const float* scale, const float* offset, int B, int L, int C,
                      cudaStream_t stream);

// ---- y = act(x*scale[b,c] + offset[b,c]) (*mask[b,t]) (+ res)   act: 0 none, 1 relu, 2 mish
struct GnApply {
  const void* x; int x_bf16;
  void* y; int y_bf16;
  const float* scale; const float* offset;  // (B,C)
  const uint8_t* mask;   // (B,L) nullable, 1 = keep
  const void* res; int res_bf16;  // nullable residual added after masking
  int act;
  int B, L, C;
};
void launch_gn_apply(const GnApply& p, cudaStream_t stream);

// ---- denoiser small pieces
// emb (n, 256) = [cos(t_i f_k), sin(t_i f_k)], f_k = exp(-ln(1e4) k / 128)   (prob_generator.py:48-67)
void launch_timestep_embedding(const float* ts, int n, int dim, float* out, cudaStream_t stream);
// s[(i*B+b), c] = silu(temb[i,c] + cvec[b,c])
void launch_silu_sum(const float* temb, const float* cvec, int nfe, int B, int C, void* out, int out_bf16,
                     cudaStream_t stream);
// x0 = noise * temperature + cond  (prob_generator.py:440); also writes a bf16 copy if xb != null
void launch_noise_init(const float* noise, const float* cond, float temperature, int64_t n, float* x, cudaStream_t s);
// out[i] = N(0,1)[i] * scale (+ add[i]): Philox4x32-10 + Box-Muller, seed read from device memory (graph-safe);
// tensor_id: 0 duration noise, 1 silence noise, 2 latent noise (see include/flamed_b200.h for the exact map)
void launch_philox_normal(const uint64_t* seed_dev, uint32_t tensor_id, float scale, const float* add, int64_t n,
                          float* out, cudaStream_t s);
void launch_f32_to_bf16(const float* x, bf16* y, int64_t n, cudaStream_t stream);
// rows (r, c) of the two operand matrices of the algebraic LayerNorm (TapGemm::raff_*), bf16:
//   s1[r, c] = w[c] * (1 + scale[r * stride + c]),  s2[r, c] = b[c] * (1 + scale[...]) + shift[r * stride + c]   (w, b nullable: 1, 0)
void launch_ln_affine_rows(const float* w, const float* b, const float* shift, const float* scale, int64_t stride, int64_t rows,
                           int C, bf16* s1, bf16* s2, cudaStream_t stream);
void launch_fill_random(void* x, int is_bf16, int64_t n, uint32_t seed, float scale, cudaStream_t stream);

// ---- cond down-sampler front end: xq[b,l,q*D+d] = prior[b,q,l,d] + qemb[q,d]; xm = xq * mask
void launch_quantizer_fold(const float* prior, const float* qemb, const uint8_t* mask, int B, int Q, int L, int D,
                           void* xq, void* xm, int out_bf16, cudaStream_t stream);

// ---- duration generator pieces
// emb (n, dim) = [sin(1000 t_i f_k), cos(...)], f_k = exp(-k ln(1e4)/(dim/2-1))   (pva.py:9-22)
void launch_sinusoidal_pos_emb(const float* ts, int n, int dim, float* out, cudaStream_t stream);
// a0[r,c] = encp[r,c] + xt[r]*w0[c] + temb[c]
void launch_durgen_input(const float* encp, const float* xt, const float* w0, const float* temb, int64_t rows, int C,
                         float* out, cudaStream_t stream);
// v = LN(relu(x); w,b) . wl + bl ; masked -> 0 ; xt += dt * v   (pva.py:217-236, 106/109)
void launch_durgen_head(const float* x, int64_t rows, int C, const float* lnw, const float* lnb, const float* wl,
                        const float* bl, const uint8_t* mask, float dt, float* xt, cudaStream_t stream);
// out = clamp(round(exp(x) - 1), 0)   (pva.py:111-112)
void launch_duration_round(const float* x, int64_t n, float* out, cudaStream_t stream);
void launch_scale(const float* x, float s, int64_t n, float* out, cudaStream_t stream);

// ---- length regulator
void launch_lr_plan(const float* phone, const float* sil, const int64_t* src_lens, int B, int P, int32_t* cumsum,
                    int64_t* tgt_len, cudaStream_t stream);
void launch_lr_expand(const float* x, const int32_t* cumsum, int B, int P, int H, int Tmax, float* out,
                      int32_t* out_index, cudaStream_t stream);

// expand into a batch whose samples come from different earlier batches: per-sample source rows / cumsum / P
void launch_lr_expand_gather(const float* const* xs, const int32_t* const* css, const int32_t* Ps, int B, int H,
                             int Tmax, float* out, int32_t* out_index, cudaStream_t stream);

// ---- codec
struct Act1d {
  const void* x; void* y; int io_bf16;  // (B,T,C)
  const float* a;     // exp(alpha) (C)
  const float* invb;  // 1/(exp(beta)+1e-9) (C)
  float fu[12], fd[12];  // up / down 12-tap filters
  int B, T, C;
  int fast_sin;  // 1: __sinf (bf16 mode), 0: sinf
};
void launch_act1d(const Act1d& p, cudaStream_t stream);
// wav[b,t] = tanh(sum_tap sum_c x[b,t+tap-3,c] w[tap][c] + bias)
void launch_conv_out_tanh(const void* x, int x_bf16, const float* w, float bias, int B, int T, int C, float* wav,
                          cudaStream_t stream);
// y[b,t,c] = sum_tap w[tap][c] * wav[b,t+tap-3] + bias[c]   (encoder first conv, 1 -> C)
void launch_conv_in_wav(const float* wav, const float* w, const float* bias, int B, int64_t T, int C, float* y,
                        cudaStream_t stream);
// (B,T,C) -> (B,C,T)
void launch_transpose_out(const float* x, int B, int T, int C, float* y, cudaStream_t stream);

// ---- self-attention with per-sample key prefixes (attention.cu): qkv (B,S,3,H,32) bf16 -> out (B,S,H*32) bf16
void launch_attn_prefix(const bf16* qkv, const int32_t* key_lens, int B, int S, int H, int dh, bf16* out, cudaStream_t stream);

// ---- prompt side of the FaCodec decoder (prompt_side.cu): residual vector quantisers + timbre transformer pieces
constexpr int VQ_MAX_LAYERS = 8;
constexpr int VQ_MAX_CD = 16;
struct VqLayer {
  const float* w_in;     // (cd, D) weight-norm folded in_proj
  const float* b_in;     // (cd)
  const float* cb;       // (n_codes, cd) codebook
  const float* cb_norm;  // (n_codes, cd) F.normalize(codebook)
  const float* cb_sq;    // (n_codes) |normalised code|^2
  const float* w_out;    // (D, cd) weight-norm folded out_proj
  const float* b_out;    // (D)
  int cd, n_codes, group;  // group 0 / 1 quantise x, group 2 quantises x - (q0 + q1)
};
struct VqPlan {
  VqLayer layer[VQ_MAX_LAYERS];
  int n_layers;
};
// x (rows, D) channels-last -> codes (n_layers, rows) int64, qgroups (3, rows, D) summed quantised vectors per group
void launch_vq_frames(const VqPlan& plan, const float* x, int64_t rows, int D, int64_t* codes, float* qgroups,
                      cudaStream_t stream);
// (B,C,T) -> (B,T,C); y_pe (nullable) = the same + pe[b, :] broadcast over T
void launch_transpose_in(const float* x, int B, int T, int C, float* y, const float* pe, float* y_pe, cudaStream_t stream);
// fp32 multi-head self-attention without a mask: qkv (B,T,3*H*DH) -> out (B,T,H*DH), DH 32 or 64
void launch_mha_fp32(const float* qkv, int B, int T, int H, int DH, float* out, cudaStream_t stream);
void launch_mean_time(const float* x, int B, int T, int C, float* y, cudaStream_t stream);

}  // namespace flm
