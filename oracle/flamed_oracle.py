"""CPU restatement (plain PyTorch, fp32 or fp64) of the Flamed-TTS inference hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and only
as the checker or as the timed CPU baseline - never as a product code path.

Every function works on a flat state-dict `sd` (reference key layout, see
oracle/weights.py) and cites the reference lines it restates.  It is pinned against
the live reference by `oracle/make_golden.py` (run in the build container, where
/root/reference can be imported) and against the committed vectors under
`tests/golden/` by `tests/test_oracle_golden.py` (runs anywhere).

Layout conventions: the reference's own - (B, L, C) for the generators, (B, C, T)
for the codec.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- helpers


def get_mask_from_lengths(lengths, max_len=None):
    """flamed/utils/tools.py:91-99 - True = padding."""
    if max_len is None:
        max_len = int(lengths.max().item())
    ids = torch.arange(0, max_len, device=lengths.device).unsqueeze(0)
    return ids >= lengths.unsqueeze(1)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _conv(sd, p, x, padding=0, dilation=1, stride=1, groups=1):
    return F.conv1d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding, dilation=dilation, groups=groups)


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


# ----------------------------------------------------------------------------- a2: durgen


def time_embedding(sd, p, t, dim):
    """pva.py:9-41: sin||cos of 1000*t*exp(-k*ln(1e4)/(half-1)), then Linear-SiLU-Linear."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1))).to(t.dtype)
    if t.ndim < 1:
        t = t.unsqueeze(0)
    emb = 1000 * t.unsqueeze(1) * freq.unsqueeze(0)
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    h = _lin(sd, p + ".time_emb.1", emb)
    return _lin(sd, p + ".time_emb.3", F.silu(h))


def prob_module_forward(sd, p, xt, enc, t, mask):
    """ProbabilisticModule.forward, pva.py:221-238."""
    out = _lin(sd, p + ".proj", torch.cat([xt.unsqueeze(-1), enc], dim=-1))
    temb = time_embedding(sd, p + ".time_emb", t, enc.shape[-1])
    out = out + temb.unsqueeze(1)
    out = _conv(sd, p + ".conv_layer.conv1d_1.conv", out.transpose(1, 2), padding=1).transpose(1, 2)
    out = _ln(sd, p + ".conv_layer.layer_norm_1", F.relu(out))
    out = _conv(sd, p + ".conv_layer.conv1d_2.conv", out.transpose(1, 2), padding=1).transpose(1, 2)
    out = _ln(sd, p + ".conv_layer.layer_norm_2", F.relu(out))
    out = _lin(sd, p + ".linear_layer", out).squeeze(-1)
    if mask is not None:
        out = out.masked_fill(mask, 0.0)
    return out


def durgen_sample(sd, p, enc, src_mask, noise_dur, noise_sil, nfe, temperature, trace=None):
    """PVA.sample up to the rounding, pva.py:88-112.  `noise_*` are the two (B,P)
    standard-normal draws (dur first).  Returns float tensors holding integers."""
    ts = torch.linspace(0, 1, nfe + 1).to(enc.dtype)
    delta_t = 1 / nfe
    dur_t = noise_dur * temperature
    sil_t = noise_sil * temperature
    for i in range(1, nfe + 1):
        dur_t = dur_t + delta_t * prob_module_forward(sd, p + ".duration_generator", dur_t, enc, ts[i - 1], src_mask)
        sil_t = sil_t + delta_t * prob_module_forward(sd, p + ".sil_generator", sil_t, enc, ts[i - 1], src_mask)
        if trace is not None:
            trace.append((dur_t.clone(), sil_t.clone()))
    phone = torch.clamp(torch.round(torch.exp(dur_t) - 1), min=0)
    sil = torch.clamp(torch.round(torch.exp(sil_t) - 1), min=0)
    return phone, sil, dur_t, sil_t


# ----------------------------------------------------------------------------- a4: length regulator


def length_regulator_index(phone_dur, sil_dur, src_lens):
    """Closed form of LengthRegulator.LR (pva.py:125-166), integer only.
    Returns (index (B,Tmax) int64 with -1 = zero padding, tgt_len (B,) int64).
    Padded phonemes get 1 frame each (pva.py:136-137); silence frames copy row 0
    (pva.py:142); segments interleave phone0, sil0, phone1, sil1, ... (pva.py:144-145)."""
    B, P = phone_dur.shape
    valid = torch.arange(P).unsqueeze(0) < src_lens.unsqueeze(1)
    ph = torch.clamp(torch.where(valid, phone_dur, torch.zeros_like(phone_dur)).round().long(), min=1)
    si = torch.clamp(torch.where(valid, sil_dur, torch.zeros_like(sil_dur)).round().long(), min=0)
    rep = torch.stack((ph, si), dim=2).reshape(B, 2 * P)
    cs = rep.cumsum(1)
    tgt_len = cs[:, -1]
    tmax = int(tgt_len.max().item())
    f = torch.arange(tmax).unsqueeze(0).expand(B, -1).contiguous()
    seg = torch.searchsorted(cs, f, right=True)
    src = torch.where(seg % 2 == 0, seg // 2, torch.zeros_like(seg))
    src = torch.where(f < tgt_len.unsqueeze(1), src, torch.full_like(src, -1))
    return src, tgt_len


def length_regulator(x, phone_dur, sil_dur, src_lens):
    idx, tgt_len = length_regulator_index(phone_dur, sil_dur, src_lens)
    g = torch.gather(x, 1, idx.clamp(min=0).unsqueeze(-1).expand(-1, -1, x.shape[-1]))
    return g * (idx >= 0).unsqueeze(-1).to(x.dtype), tgt_len


# ----------------------------------------------------------------------------- FFT glue (rows f1)


def fft_block(sd, p, x, mask, n_head):
    """Layers.py:11-30, SubLayers.py:8-93, Modules.py:6-25 (eval mode: dropout = identity)."""
    B, L, D = x.shape
    dk = D // n_head
    q = _lin(sd, p + ".slf_attn.w_qs", x).view(B, L, n_head, dk).permute(0, 2, 1, 3)
    k = _lin(sd, p + ".slf_attn.w_ks", x).view(B, L, n_head, dk).permute(0, 2, 1, 3)
    v = _lin(sd, p + ".slf_attn.w_vs", x).view(B, L, n_head, dk).permute(0, 2, 1, 3)
    attn = torch.matmul(q, k.transpose(-1, -2)) / (dk ** 0.5)
    attn = attn.masked_fill(mask[:, None, None, :], float("-inf"))
    attn = torch.softmax(attn, dim=-1)
    o = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(B, L, D)
    o = _ln(sd, p + ".slf_attn.layer_norm", _lin(sd, p + ".slf_attn.fc", o) + x)
    o = o.masked_fill(mask.unsqueeze(-1), 0)
    w1, w2 = sd[p + ".pos_ffn.w_1.weight"], sd[p + ".pos_ffn.w_2.weight"]
    h = F.conv1d(o.transpose(1, 2), w1, sd[p + ".pos_ffn.w_1.bias"], padding=(w1.shape[-1] - 1) // 2)
    h = F.conv1d(F.relu(h), w2, sd[p + ".pos_ffn.w_2.bias"], padding=(w2.shape[-1] - 1) // 2).transpose(1, 2)
    o = _ln(sd, p + ".pos_ffn.layer_norm", h + o)
    return o.masked_fill(mask.unsqueeze(-1), 0)


def _n_layers(sd, p):
    n = 0
    while f"{p}.layer_stack.{n}.slf_attn.fc.weight" in sd:
        n += 1
    return n


def phoneme_encoder(sd, p, phonemes, mask, n_head):
    """Models.py:76-104."""
    x = F.embedding(phonemes, sd[p + ".src_word_emb.weight"]) + sd[p + ".position_enc"][:, : phonemes.shape[1]]
    for i in range(_n_layers(sd, p)):
        x = fft_block(sd, f"{p}.layer_stack.{i}", x, mask, n_head)
    return x


def fft_decoder(sd, p, x, mask, n_head):
    """Models.py:139-171 (sequence shorter than decoder_max_seq_len)."""
    x = x + sd[p + ".position_enc"][:, : x.shape[1]]
    for i in range(_n_layers(sd, p)):
        x = fft_block(sd, f"{p}.layer_stack.{i}", x, mask, n_head)
    return x


def prior_after_pva(sd, p, x, tgt_lens, prompts, cfg):
    """PriorGenerator.sample after pva.sample, prior_generator.py:162-181."""
    n_head = cfg["transformer"]["decoder_head"]
    x = _lin(sd, p + ".bridge", x)
    tgt_mask = get_mask_from_lengths(tgt_lens, x.shape[1])
    x = fft_decoder(sd, p + ".shared_decoder", x, tgt_mask, n_head)
    lp = prompts.shape[-1]
    dec_mask = get_mask_from_lengths(lp + tgt_lens, lp + x.shape[1])
    prompt_embs = F.embedding(prompts, sd[p + ".code_embedding.weight"])
    hiddens = []
    for q in range(cfg["codec"]["n_quantizers"]):
        inp = torch.cat([prompt_embs[:, q], x], dim=1).clone()
        inp[:, :lp] = inp[:, :lp] + sd[p + ".pre_encode.prompt_emb"]
        inp[:, lp:] = inp[:, lp:] + sd[p + ".pre_encode.target_emb"]
        inp = inp + sd[p + ".pre_encode.quantizer_emb.weight"][q]
        x = fft_decoder(sd, f"{p}.prior_decoder.{q}", inp, dec_mask, n_head)[:, lp:]
        hiddens.append(x.unsqueeze(1))
    out = torch.cat(hiddens, dim=1)
    logits = _lin(sd, p + ".head", out)
    logits = logits * (~tgt_mask)[:, None, :, None]
    return out, logits.permute(0, 3, 1, 2).contiguous(), tgt_mask


# ----------------------------------------------------------------------------- a5/a6: prob generator


def cond_prepare(sd, p, prior_embs, mask):
    """QuantizerEncoding + ConditionDownSampler, prob_generator.py:375-381, 198-205.
    mask: (B,L,1) bool True = valid."""
    B, Q, L, D = prior_embs.shape
    x = prior_embs + sd[p + ".quantizer_encoding.quantizer_emb.weight"][None, :, None, :]
    x = x.permute(0, 2, 1, 3).reshape(B, L, Q * D).transpose(1, 2)
    m = mask.transpose(1, 2).to(x.dtype)
    s = 0
    while f"{p}.cond_downsampling.resblocks.{s}.block.block.0.weight" in sd:
        rp = f"{p}.cond_downsampling.resblocks.{s}.block.block"
        h = _conv(sd, rp + ".0", x * m)
        h = F.group_norm(h, 8, sd[rp + ".1.weight"], sd[rp + ".1.bias"], 1e-5)
        x = x + F.mish(h) * m
        dp = f"{p}.cond_downsampling.downblocks.{s}"
        h = _conv(sd, dp + ".0", x)
        x = F.relu(F.group_norm(h, 8, sd[dp + ".1.weight"], sd[dp + ".1.bias"], 1e-5))
        s += 1
    return F.relu(_lin(sd, p + ".cond_downsampling.proj_out.0", x.transpose(1, 2)))


def timestep_embedding(t, dim=256, max_period=10000):
    """prob_generator.py:48-67: cos||sin, t NOT scaled."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.dtype)
    args = t[:, :, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _modulate(x, shift, scale):
    return x * (1 + scale) + shift


def _convnext(sd, p, x):
    """ConvNeXtBlock.forward, prob_generator.py:107-111 (x: (B,L,C))."""
    xt = x.transpose(1, 2)
    k = sd[p + ".conv_1.weight"].shape[-1]
    h = F.conv1d(xt, sd[p + ".conv_1.weight"], sd[p + ".conv_1.bias"], padding=k // 2, groups=xt.shape[1])
    h = F.group_norm(h, h.shape[1], sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"], 1e-5)
    h = _conv(sd, p + ".conv_3", F.gelu(_conv(sd, p + ".conv_2", h)))
    return (xt + h).transpose(1, 2)


def denoiser_forward(sd, p, x, t, spk):
    """SimpleMLPAdaLN.forward, prob_generator.py:349-365.  t: (1,1)."""
    temb = _lin(sd, p + ".time_embed.mlp.2", F.silu(_lin(sd, p + ".time_embed.mlp.0", timestep_embedding(t))))
    y = temb + _lin(sd, p + ".cond_embed", spk).unsqueeze(1)
    h = _lin(sd, p + ".proj_in", x)
    H = h.shape[-1]
    i = 0
    while f"{p}.res_blocks.{i}.mlp.0.weight" in sd:
        bp = f"{p}.res_blocks.{i}"
        sc, cc, gc, sm, cm, gm = _lin(sd, bp + ".adaLN_modulation.1", F.silu(y)).chunk(6, dim=-1)
        h = h + gc * _convnext(sd, bp + ".conv_in", _modulate(_ln(sd, bp + ".ln_conv", h, 1e-6), sc, cc))
        w = _modulate(_ln(sd, bp + ".ln_mlp", h, 1e-6), sm, cm)
        h = h + gm * _lin(sd, bp + ".mlp.2", F.silu(_lin(sd, bp + ".mlp.0", w)))
        i += 1
    fp = p + ".final_layer"
    sc, cc, gc, sm, cm = _lin(sd, fp + ".adaLN_modulation.1", F.silu(y)).chunk(5, dim=-1)
    h = h + gc * _convnext(sd, fp + ".conv_in", _modulate(F.layer_norm(h, (H,), None, None, 1e-6), sc, cc))
    h = _modulate(F.layer_norm(h, (H,), None, None, 1e-6), sm, cm)
    return _conv(sd, fp + ".conv_out", h.transpose(1, 2), padding=1).transpose(1, 2)


def denoiser_sample(sd, p, cond, spk, noise, nfe, temperature, trace_steps=None):
    """ProbGenerator.sample loop, prob_generator.py:439-446.  cond: (B,L,256) already
    down-sampled; noise: the (B,L,256) standard-normal draw.  Returns (B,256,L)."""
    ts = torch.linspace(0, 1, nfe + 1).to(cond.dtype)
    xt = noise * temperature + cond
    delta_t = 1 / nfe
    trace = {}
    for i in range(1, nfe + 1):
        vt = denoiser_forward(sd, p + ".denoiser", xt, ts[i - 1].unsqueeze(0).unsqueeze(1), spk)
        xt = xt + delta_t * vt
        if trace_steps is not None and i in trace_steps:
            trace[i] = xt.clone()
    if trace_steps is not None:
        return xt.transpose(1, -1), trace
    return xt.transpose(1, -1)


# ----------------------------------------------------------------------------- a7-a10: codec


def wn_weight(sd, p):
    """torch.nn.utils.weight_norm (old style): w = g * v / ||v||, norm over all dims
    but 0 (facodec.py:27-32).  For ConvTranspose1d dim 0 is IN channels."""
    v, g = sd[p + ".weight_v"], sd[p + ".weight_g"]
    n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.ndim - 1)))
    return v * (g / n)


def activation1d(sd, p, x):
    """Activation1d = UpSample1d x2 -> SnakeBeta(log-scale) -> DownSample1d /2
    (act.py:24-29, resample.py:28-37, filter.py:89-96, facodec.py:105-118)."""
    C = x.shape[1]
    fu = sd[p + ".upsample.filter"].to(x.dtype)
    fd = sd[p + ".downsample.lowpass.filter"].to(x.dtype)
    K = fu.shape[-1]
    ratio = 2
    pad = K // ratio - 1
    pad_left = pad * ratio + (K - ratio) // 2
    pad_right = pad * ratio + (K - ratio + 1) // 2
    u = F.pad(x, (pad, pad), mode="replicate")
    u = ratio * F.conv_transpose1d(u, fu.expand(C, -1, -1), stride=ratio, groups=C)
    u = u[..., pad_left:-pad_right]
    a = torch.exp(sd[p + ".act.alpha"])[None, :, None]
    b = torch.exp(sd[p + ".act.beta"])[None, :, None]
    s = u + (1.0 / (b + 0.000000001)) * torch.pow(torch.sin(u * a), 2)
    s = F.pad(s, (K // 2 - 1, K // 2), mode="replicate")
    return F.conv1d(s, fd.expand(C, -1, -1), stride=ratio, groups=C)


def residual_unit(sd, p, x, dilation):
    """facodec.py:121-133."""
    h = activation1d(sd, p + ".block.0", x)
    h = F.conv1d(h, wn_weight(sd, p + ".block.1"), sd[p + ".block.1.bias"], dilation=dilation, padding=3 * dilation)
    h = activation1d(sd, p + ".block.2", h)
    h = F.conv1d(h, wn_weight(sd, p + ".block.3"), sd[p + ".block.3.bias"])
    return x + h


def codec_decode(sd, latents, spk, up_ratios=(5, 5, 4, 2)):
    """FACodecDecoder.inference, facodec.py:630-638 + model stack 400-415."""
    style = _lin(sd, "timbre_linear", spk).unsqueeze(2)
    gamma, beta = style.chunk(2, 1)
    x = F.layer_norm(latents.transpose(1, 2), (latents.shape[1],)).transpose(1, 2)
    x = x * gamma + beta
    x = F.conv1d(x, wn_weight(sd, "model.0"), sd["model.0.bias"], padding=3)
    for i, s in enumerate(up_ratios):
        p = f"model.{i + 1}"
        x = activation1d(sd, p + ".block.0", x)
        x = F.conv_transpose1d(x, wn_weight(sd, p + ".block.1"), sd[p + ".block.1.bias"], stride=s,
                               padding=s // 2 + s % 2, output_padding=s % 2)
        for j, d in enumerate((1, 3, 9)):
            x = residual_unit(sd, f"{p}.block.{j + 2}", x, d)
    n = len(up_ratios) + 1
    x = activation1d(sd, f"model.{n}", x)
    x = F.conv1d(x, wn_weight(sd, f"model.{n + 1}"), sd[f"model.{n + 1}.bias"], padding=3)
    return torch.tanh(x)


def codec_encode(sd, wav, up_ratios=(2, 4, 5, 5)):
    """FACodecEncoder.forward, facodec.py:183-217."""
    x = F.conv1d(wav, wn_weight(sd, "block.0"), sd["block.0.bias"], padding=3)
    for i, s in enumerate(up_ratios):
        p = f"block.{i + 1}"
        for j, d in enumerate((1, 3, 9)):
            x = residual_unit(sd, f"{p}.block.{j}", x, d)
        x = activation1d(sd, p + ".block.3", x)
        x = F.conv1d(x, wn_weight(sd, p + ".block.4"), sd[p + ".block.4.bias"], stride=s, padding=s // 2 + s % 2)
    n = len(up_ratios) + 1
    x = activation1d(sd, f"block.{n}", x)
    return F.conv1d(x, wn_weight(sd, f"block.{n + 1}"), sd[f"block.{n + 1}.bias"], padding=1)


# ---- prompt-side VQ + timbre transformer (row f3; PyTorch glue in the product as well)


def _wn_linear(sd, p, x):
    v, g = sd[p + ".weight_v"], sd[p + ".weight_g"]
    return F.linear(x, v * (g / v.norm(dim=1, keepdim=True)), sd[p + ".bias"])


def _fvq(sd, p, z):
    """FactorizedVectorQuantize.forward (eval), fvq.py:37-116.  z: (B,D,T)."""
    z_e = _wn_linear(sd, p + ".in_proj", z.transpose(1, 2))  # (B,T,8)
    cb = sd[p + "._codebook.weight"]
    enc = F.normalize(z_e.reshape(-1, z_e.shape[-1]))
    cbn = F.normalize(cb)
    dist = enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cbn.t() + cbn.pow(2).sum(1, keepdim=True).t()
    idx = (-dist).max(1)[1].view(z.shape[0], -1)
    z_q = F.embedding(idx, cb)  # (B,T,8)
    z_q = z_e + (z_q - z_e)
    return _wn_linear(sd, p + ".out_proj", z_q).transpose(1, 2), idx


def _rvq(sd, p, x, n):
    """ResidualVQ.forward (eval), rvq.py:27-73."""
    residual, out, idxs = x, 0.0, []
    for i in range(n):
        q, idx = _fvq(sd, f"{p}.layers.{i}", residual)
        residual = residual - q
        out = out + q
        idxs.append(idx)
    return out, torch.stack(idxs)


def codec_prompt_features(sd, enc_out, n_q=(1, 2, 3)):
    """FACodecDecoder.forward(vq=True): codes (6,B,T) + timbre (B,256);
    facodec.py:470-507, 521-533; timbre transformer transformer.py:86-234."""
    q0, i0 = _rvq(sd, "quantizer.0", enc_out, n_q[0])
    q1, i1 = _rvq(sd, "quantizer.1", enc_out, n_q[1])
    _, i2 = _rvq(sd, "quantizer.2", enc_out - (q0 + q1), n_q[2])
    codes = torch.cat([i0, i1, i2], dim=0)
    x = enc_out.transpose(1, 2)
    # PositionalEncoding indexes pe by x.size(0) (= batch!) - transformer.py:50-52
    x = x + sd["timbre_encoder.position_emb.pe"][: x.shape[0]]
    i = 0
    while f"timbre_encoder.layers.{i}.ln_1.weight" in sd:
        p = f"timbre_encoder.layers.{i}"
        h = _ln(sd, p + ".ln_1", x)
        B, T, D = h.shape
        qkv = F.linear(h, sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"])
        q, k, v = [t.view(B, T, 4, D // 4).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(D // 4), dim=-1) @ v
        x = x + _lin(sd, p + ".self_attn.out_proj", a.transpose(1, 2).reshape(B, T, D))
        h = _ln(sd, p + ".ln_2", x)
        h = F.relu(_conv(sd, p + ".ffn.ffn_1", h.transpose(1, 2), padding=2).transpose(1, 2))
        x = x + _lin(sd, p + ".ffn.ffn_2", h)
        i += 1
    x = _ln(sd, "timbre_encoder.last_ln", x)
    return codes, x.mean(dim=1)


# ----------------------------------------------------------------------------- whole path


def front_stage(sd, cfg, phonemes, src_lens, noise_dur, noise_sil, nfe_dur, temp_dur):
    """first half of Flamed.sample_batch (flamed.py:183-196 -> prior_generator.py:141-161): phoneme encoder, the two
    duration ODEs + rounding, length regulator.  Returns a dict with x (B,Tmax,192) zero-padded and tgt_len."""
    P = "prior_generator"
    src_mask = get_mask_from_lengths(src_lens, phonemes.shape[1])
    enc = phoneme_encoder(sd, P + ".encoder", phonemes, src_mask, cfg["prior_generator"]["transformer"]["encoder_head"])
    phone, sil, dur_t, sil_t = durgen_sample(sd, P + ".pva", enc, src_mask, noise_dur, noise_sil, nfe_dur, temp_dur)
    x, tgt_len = length_regulator(enc, phone, sil, src_lens)
    return dict(enc=enc, phone_dur=phone, sil_dur=sil, dur_t=dur_t, sil_t=sil_t, x=x, tgt_len=tgt_len)


def back_stage(sd, cfg, x, tgt_len, prompts, timbres, noise_lat_fn, nfe_den, temp_den, codec_sd=None):
    """second half of Flamed.sample_batch (prior_generator.py:162-181, flamed.py:198-215) on a length-regulated,
    zero-padded batch x (B,Tmax,192): prior decoders, cond fold, denoiser loop, codec."""
    prior_embs, logits, tgt_mask = prior_after_pva(sd, "prior_generator", x, tgt_len, prompts, cfg["prior_generator"])
    cond = cond_prepare(sd, "prob_generator", prior_embs, ~tgt_mask.unsqueeze(-1))
    noise = noise_lat_fn(cond.shape[0], cond.shape[1])
    latents = denoiser_sample(sd, "prob_generator", cond, timbres, noise, nfe_den, temp_den)
    out = dict(prior_embs=prior_embs, prior_logits=logits, tgt_mask=tgt_mask, cond=cond, latents=latents)
    if codec_sd is not None:
        out["wav"] = codec_decode(codec_sd, latents, timbres)
    return out


def repad(xs, tgt_lens):
    """utterances (each (Tmax_i,192) from its own front batch, with its frame count) -> one zero-padded batch: what
    tools.py:299-317 `pad` does inside the length regulator, applied to a different grouping of the same utterances"""
    T = int(max(tgt_lens))
    out = torch.zeros((len(xs), T, xs[0].shape[-1]), dtype=xs[0].dtype)
    for i, (x, n) in enumerate(zip(xs, tgt_lens)):
        out[i, : int(n)] = x[: int(n)]
    return out


def sample_batch(sd, cfg, phonemes, src_lens, prompts, timbres, noise_dur, noise_sil, noise_lat_fn,
                 nfe_dur, nfe_den, temp_dur, temp_den, codec_sd=None):
    """Flamed.sample_batch, flamed.py:168-217, with the three CPU randn draws
    (pva.py:101-102, prob_generator.py:440) injected.  `noise_lat_fn(B, L)` returns
    the (B,L,256) draw once L is known."""
    out = front_stage(sd, cfg, phonemes, src_lens, noise_dur, noise_sil, nfe_dur, temp_dur)
    out.update(back_stage(sd, cfg, out["x"], out["tgt_len"], prompts, timbres, noise_lat_fn, nfe_den, temp_den, codec_sd))
    return out
