"""Seeded random-init weights in the reference's state-dict key layout + the synthetic workloads of
BASELINE.json (used by tests, the oracle, smoke() and bench.py; there is no network for real
checkpoints or datasets).  No pretrained checkpoint is
available offline, and the reference's default init makes the denoiser output
exactly zero (prob_generator.py:338-347 zero-inits every adaLN projection and
final_layer.conv_out), so parity runs use these weights instead:

* every tensor is drawn from its own generator seeded by crc32(name) ^ seed, so
  the result is independent of enumeration order and identical on every box
  with the same torch build (CPU generator);
* dense / conv weights ~ U(+-1/sqrt(fan_in)); tensors that the reference
  initialises to a constant (norm affine, biases, Snake alpha/beta, adaLN,
  conv_out) are perturbed so that no term of the arithmetic is trivially 0 or 1;
* `pva.duration_generator.linear_layer.bias` = 2.0 and `sil_generator` = 0.1 (a few 1-frame silences)
  give LibriSpeech-like durations (~6.7 frames/phoneme at temperature 0.3,
  SURVEY.md Appendix C).

The shapes below restate the constructors of the reference:
  flamed/models/synthesizer/prior_generator.py:28-61, pva.py:173-219,
  prob_generator.py:267-412, module/transformer/Models.py:33-104,
  facodec/facodec.py:121-155,183-213,246-265,400-431, facodec/quantize/fvq.py:16-29,
  facodec/transformer.py:86-206.
`tests/golden/state_dict_keys.json` (written by oracle/make_golden.py from the live
reference) pins names and shapes.
"""
import math
import zlib

import numpy as np
import torch

N_SYMBOLS = 360  # len(flamed.text.symbols.symbols), SURVEY.md section 1
# output biases of the duration / silence generators that give LibriSpeech-like speaking rates with these
# random weights (measured: ~6.7 frames per phoneme = 12 phonemes/s at 80 frames/s, temperature 0.3)
BENCH_DUR_BIAS, BENCH_SIL_BIAS = 1.35, -0.5


def _gen(name, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def _uniform(name, shape, seed, fan_in):
    bound = 1.0 / math.sqrt(max(fan_in, 1))
    return (torch.rand(shape, generator=_gen(name, seed)) * 2 - 1) * bound


def _normal(name, shape, seed, std, mean=0.0):
    return torch.randn(shape, generator=_gen(name, seed)) * std + mean


def sinusoid_table(n_position, d_hid):
    """module/transformer/Models.py:10-30 (float64 numpy, then FloatTensor)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    idx = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000.0, 2 * (idx // 2) / d_hid)
    tab = np.empty_like(ang)
    tab[:, 0::2] = np.sin(ang[:, 0::2])
    tab[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.FloatTensor(tab)


def kaiser_sinc_filter(cutoff=0.25, half_width=0.3, kernel_size=12):
    """facodec/alias_free_torch/filter.py:27-58 -> (1,1,K)."""
    half = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    f = f / f.sum()
    return f.view(1, 1, kernel_size)


class _Builder:
    def __init__(self, seed):
        self.seed = seed
        self.sd = {}

    def linear(self, name, out_f, in_f, w_std=None, b_std=0.1):
        if w_std is None:
            self.sd[name + ".weight"] = _uniform(name + ".weight", (out_f, in_f), self.seed, in_f)
        else:
            self.sd[name + ".weight"] = _normal(name + ".weight", (out_f, in_f), self.seed, w_std)
        self.sd[name + ".bias"] = _normal(name + ".bias", (out_f,), self.seed, b_std)

    def conv(self, name, out_c, in_c, k, w_std=None, b_std=0.1):
        if w_std is None:
            self.sd[name + ".weight"] = _uniform(name + ".weight", (out_c, in_c, k), self.seed, in_c * k)
        else:
            self.sd[name + ".weight"] = _normal(name + ".weight", (out_c, in_c, k), self.seed, w_std)
        self.sd[name + ".bias"] = _normal(name + ".bias", (out_c,), self.seed, b_std)

    def norm(self, name, c):
        self.sd[name + ".weight"] = _normal(name + ".weight", (c,), self.seed, 0.1, 1.0)
        self.sd[name + ".bias"] = _normal(name + ".bias", (c,), self.seed, 0.1)

    def emb(self, name, n, d, std=1.0, padding_idx=None):
        w = _normal(name, (n, d), self.seed, std)
        if padding_idx is not None:
            w[padding_idx] = 0
        self.sd[name] = w

    def wn_conv(self, name, dim0, dim1, k, fan_in, bias_c):
        """old-style torch weight_norm: weight_g (dim0,1,1), weight_v (dim0,dim1,k)."""
        v = _uniform(name + ".weight_v", (dim0, dim1, k), self.seed, fan_in)
        g = v.flatten(1).norm(dim=1).view(dim0, 1, 1)
        g = g * _normal(name + ".weight_g", (dim0, 1, 1), self.seed, 0.1, 1.0)
        self.sd[name + ".bias"] = _normal(name + ".bias", (bias_c,), self.seed, 0.05)
        self.sd[name + ".weight_g"] = g
        self.sd[name + ".weight_v"] = v

    def wn_linear(self, name, out_f, in_f):
        v = _uniform(name + ".weight_v", (out_f, in_f), self.seed, in_f)
        g = v.norm(dim=1, keepdim=True) * _normal(name + ".weight_g", (out_f, 1), self.seed, 0.1, 1.0)
        self.sd[name + ".bias"] = _normal(name + ".bias", (out_f,), self.seed, 0.05)
        self.sd[name + ".weight_g"] = g
        self.sd[name + ".weight_v"] = v

    def act(self, name, c):
        """Activation1d(SnakeBeta): alpha, beta (log-scale) + two 12-tap filter buffers."""
        self.sd[name + ".act.alpha"] = _normal(name + ".act.alpha", (c,), self.seed, 0.3)
        self.sd[name + ".act.beta"] = _normal(name + ".act.beta", (c,), self.seed, 0.3)
        self.sd[name + ".upsample.filter"] = kaiser_sinc_filter()
        self.sd[name + ".downsample.lowpass.filter"] = kaiser_sinc_filter()


def _fft_block(b, p, d, d_inner, ks):
    for nm in ("w_qs", "w_ks", "w_vs", "fc"):
        b.linear(f"{p}.slf_attn.{nm}", d, d)
    b.norm(f"{p}.slf_attn.layer_norm", d)
    b.conv(f"{p}.pos_ffn.w_1", d_inner, d, ks[0])
    b.conv(f"{p}.pos_ffn.w_2", d, d_inner, ks[1])
    b.norm(f"{p}.pos_ffn.layer_norm", d)


def _prob_module(b, p, cfg, out_bias):
    d, f, k, ts = cfg["input_size"], cfg["filter_size"], cfg["kernel_size"], cfg["time_scale"]
    b.linear(f"{p}.proj", d, d + 1)
    b.linear(f"{p}.time_emb.time_emb.1", d * ts, d)
    b.linear(f"{p}.time_emb.time_emb.3", d, d * ts)
    b.conv(f"{p}.conv_layer.conv1d_1.conv", f, d, k)
    b.norm(f"{p}.conv_layer.layer_norm_1", f)
    b.conv(f"{p}.conv_layer.conv1d_2.conv", f, f, k)
    b.norm(f"{p}.conv_layer.layer_norm_2", f)
    b.linear(f"{p}.linear_layer", 1, f)
    b.sd[f"{p}.linear_layer.bias"] = torch.full((1,), float(out_bias))


def _convnext(b, p, c, k):
    b.conv(f"{p}.conv_1", c, 1, k)
    b.norm(f"{p}.ln_1", c)
    b.conv(f"{p}.conv_2", c, c, 1)
    b.conv(f"{p}.conv_3", c, c, 1)


def make_flamed_state_dict(prior_cfg, prob_cfg, seed=0, dur_bias=2.0, sil_bias=0.1):
    """504 tensors, keys as `Flamed(cfg).state_dict()` in the reference."""
    b = _Builder(seed)
    tr = prior_cfg["transformer"]
    # ---- prior generator (prior_generator.py:28-61)
    P = "prior_generator"
    de, dd = tr["encoder_hidden"], tr["decoder_hidden"]
    b.sd[f"{P}.encoder.position_enc"] = sinusoid_table(tr["encoder_max_seq_len"] + 1, de).unsqueeze(0)
    b.emb(f"{P}.encoder.src_word_emb.weight", N_SYMBOLS + 1, de, padding_idx=0)
    for i in range(tr["encoder_layer"]):
        _fft_block(b, f"{P}.encoder.layer_stack.{i}", de, tr["encoder_conv_filter_size"], tr["encoder_conv_kernel_size"])
    va = prior_cfg["variance_adaptor"]
    _prob_module(b, f"{P}.pva.duration_generator", va["duration_generator"], dur_bias)
    _prob_module(b, f"{P}.pva.sil_generator", va["sil_generator"], sil_bias)
    b.linear(f"{P}.bridge", dd, de)
    vocab, nq = prior_cfg["codec"]["vocab_size"], prior_cfg["codec"]["n_quantizers"]
    b.emb(f"{P}.code_embedding.weight", vocab + 1, dd, padding_idx=vocab)

    def decoder(p, n_layers):
        b.sd[f"{p}.position_enc"] = sinusoid_table(tr["decoder_max_seq_len"] + 1, dd).unsqueeze(0)
        for i in range(n_layers):
            _fft_block(b, f"{p}.layer_stack.{i}", dd, tr["decoder_conv_filter_size"], tr["decoder_conv_kernel_size"])

    decoder(f"{P}.shared_decoder", tr["decoder_shared_layers"])
    b.sd[f"{P}.pre_encode.prompt_emb"] = torch.rand((1, 1, dd), generator=_gen("prompt_emb", seed))
    b.sd[f"{P}.pre_encode.target_emb"] = torch.rand((1, 1, dd), generator=_gen("target_emb", seed))
    b.emb(f"{P}.pre_encode.quantizer_emb.weight", nq, dd)
    for q in range(nq):
        decoder(f"{P}.prior_decoder.{q}", tr["decoder_layers"][q])
    b.linear(f"{P}.head", vocab + 1, dd)
    # ---- prob generator (prob_generator.py:384-412)
    G = "prob_generator"
    H, D, S = prob_cfg["hidden_dim"], prob_cfg["target_dim"], prob_cfg["spk_dim"]
    cin = prob_cfg["n_quantizers"] * prob_cfg["cond_dim"]
    b.emb(f"{G}.quantizer_encoding.quantizer_emb.weight", prob_cfg["n_quantizers"], prob_cfg["cond_dim"])
    for s in range(prob_cfg["downsampling_stages"]):
        b.conv(f"{G}.cond_downsampling.resblocks.{s}.block.block.0", cin, cin, 1)
        b.norm(f"{G}.cond_downsampling.resblocks.{s}.block.block.1", cin)
        b.conv(f"{G}.cond_downsampling.downblocks.{s}.0", cin // 2, cin, 1)
        b.norm(f"{G}.cond_downsampling.downblocks.{s}.1", cin // 2)
        cin //= 2
    b.linear(f"{G}.cond_downsampling.proj_out.0", D, cin)
    dn = f"{G}.denoiser"
    b.linear(f"{dn}.time_embed.mlp.0", H, 256, w_std=0.02)
    b.linear(f"{dn}.time_embed.mlp.2", H, H, w_std=0.02)
    b.linear(f"{dn}.cond_embed", H, S)
    b.linear(f"{dn}.proj_in", H, D)
    k = prob_cfg["convnext"]["kernel_size"]
    for i in range(prob_cfg["n_layers"]):
        p = f"{dn}.res_blocks.{i}"
        b.linear(f"{p}.adaLN_modulation.1", 6 * H, H, w_std=0.02, b_std=0.02)
        b.norm(f"{p}.ln_conv", H)
        _convnext(b, f"{p}.conv_in", H, k)
        b.norm(f"{p}.ln_mlp", H)
        b.linear(f"{p}.mlp.0", H, H)
        b.linear(f"{p}.mlp.2", H, H)
    p = f"{dn}.final_layer"
    b.linear(f"{p}.adaLN_modulation.1", 5 * H, H, w_std=0.02, b_std=0.02)
    _convnext(b, f"{p}.conv_in", H, k)
    b.conv(f"{p}.conv_out", D, H, 3, w_std=0.02, b_std=0.02)
    return b.sd


def _residual_unit(b, p, c):
    b.act(f"{p}.block.0", c)
    b.wn_conv(f"{p}.block.1", c, c, 7, c * 7, c)
    b.act(f"{p}.block.2", c)
    b.wn_conv(f"{p}.block.3", c, c, 1, c, c)


def make_codec_encoder_state_dict(seed=0, ngf=32, up_ratios=(2, 4, 5, 5), out_channels=256):
    """206 tensors, keys as FACodecEncoder.state_dict() (facodec.py:183-213)."""
    b = _Builder(seed + 101)
    d = ngf
    b.wn_conv("block.0", d, 1, 7, 7, d)
    for i, s in enumerate(up_ratios):
        d *= 2
        p = f"block.{i + 1}"
        for j in range(3):
            _residual_unit(b, f"{p}.block.{j}", d // 2)
        b.act(f"{p}.block.3", d // 2)
        b.wn_conv(f"{p}.block.4", d, d // 2, 2 * s, (d // 2) * 2 * s, d)
    n = len(up_ratios) + 1
    b.act(f"block.{n}", d)
    b.wn_conv(f"block.{n + 1}", out_channels, d, 3, d * 3, out_channels)
    return b.sd


def make_codec_decoder_state_dict(seed=0, in_channels=256, channels=1024, up_ratios=(5, 5, 4, 2),
                                  n_q=(1, 2, 3), codebook_dim=8, codebook_size=1024):
    """The 301 hot-path + prompt-side tensors of FACodecDecoder.state_dict()
    (`model.*`, `quantizer.*`, `timbre_encoder.*`, `timbre_linear.*`;
    facodec.py:340-431).  The 244 tensors of the training-only heads
    (f0/phone/x_timbre predictors) are not generated: inference never reads them."""
    b = _Builder(seed + 202)
    for gi, n in enumerate(n_q):
        for li in range(n):
            p = f"quantizer.{gi}.layers.{li}"
            b.wn_linear(f"{p}.in_proj", codebook_dim, in_channels)
            b.wn_linear(f"{p}.out_proj", in_channels, codebook_dim)
            b.emb(f"{p}._codebook.weight", codebook_size, codebook_dim)
    b.wn_conv("model.0", channels, in_channels, 7, in_channels * 7, channels)
    out_dim = channels
    for i, s in enumerate(up_ratios):
        in_dim, out_dim = channels // 2 ** i, channels // 2 ** (i + 1)
        p = f"model.{i + 1}"
        b.act(f"{p}.block.0", in_dim)
        # ConvTranspose1d weight is (in, out, k); weight_norm dim 0 = IN channels
        b.wn_conv(f"{p}.block.1", in_dim, out_dim, 2 * s, in_dim * 2, out_dim)
        for j in range(3):
            _residual_unit(b, f"{p}.block.{j + 2}", out_dim)
    n = len(up_ratios) + 1
    b.act(f"model.{n}", out_dim)
    b.wn_conv(f"model.{n + 1}", 1, out_dim, 7, out_dim * 7, 1)
    # timbre transformer (facodec/transformer.py:154-234) + timbre_linear (facodec.py:428-430)
    pe = torch.zeros(5000, 1, in_channels)
    position = torch.arange(5000).unsqueeze(1)
    div = torch.exp(torch.arange(0, in_channels, 2) * (-math.log(10000.0) / in_channels))
    pe[:, 0, 0::2] = torch.sin(position * div)
    pe[:, 0, 1::2] = torch.cos(position * div)
    b.sd["timbre_encoder.position_emb.pe"] = pe
    for i in range(4):
        p = f"timbre_encoder.layers.{i}"
        b.norm(f"{p}.ln_1", in_channels)
        b.norm(f"{p}.ln_2", in_channels)
        b.sd[f"{p}.self_attn.in_proj_weight"] = _uniform(f"{p}.in_proj_weight", (3 * in_channels, in_channels), b.seed, in_channels)
        b.sd[f"{p}.self_attn.in_proj_bias"] = _normal(f"{p}.in_proj_bias", (3 * in_channels,), b.seed, 0.05)
        b.linear(f"{p}.self_attn.out_proj", in_channels, in_channels)
        b.conv(f"{p}.ffn.ffn_1", 1024, in_channels, 5, w_std=0.02)
        b.linear(f"{p}.ffn.ffn_2", in_channels, 1024, w_std=0.02)
    b.norm("timbre_encoder.last_ln", in_channels)
    b.linear("timbre_linear", 2 * in_channels, in_channels)
    bias = b.sd["timbre_linear.bias"]
    bias[:in_channels] += 1.0  # reference init: gamma bias 1, beta bias 0 (facodec.py:429-430)
    return b.sd


# ----------------------------------------------------------------------------- synthetic workloads
def metadata_workload(n_utterances=256, n_prompts=64, seed=0, dur_range=(2.0, 15.0), phonemes_per_second=12.0):
    """BASELINE.json config 3 (SURVEY.md section 8 d): `n` utterances with LibriSpeech-like durations
    d_i ~ U(2,15) s, P_i = round(12 d_i) phoneme ids uniform in the ARPAbet range of the symbol table,
    utterance i speaks with prompt i mod n_prompts; prompts are 3 s of seeded noise * 0.1 at 16 kHz."""
    rng = np.random.default_rng(seed)
    dur = rng.uniform(dur_range[0], dur_range[1], size=n_utterances)
    n_ph = np.maximum(1, np.rint(phonemes_per_second * dur).astype(np.int64))
    lo, hi = 64, 64 + 84  # '@AA' (64) ... '@ZH' (147): ids of the 84 ARPAbet symbols in flamed/text/symbol_table.json
    phonemes = [torch.from_numpy(rng.integers(lo, hi, size=int(p))) for p in n_ph]
    prng = np.random.default_rng(seed * 7919 + 1)
    prompts = torch.from_numpy((prng.standard_normal((n_prompts, 1, 48000)) * 0.1).astype(np.float32))
    prompt_of = np.arange(n_utterances) % n_prompts
    return dict(durations_s=dur, phonemes=phonemes, prompts_wav=prompts, prompt_of=prompt_of)
