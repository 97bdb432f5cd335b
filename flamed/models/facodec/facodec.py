"""FaCodec encoder / decoder drop-ins.

Same constructors, `from_pretrained`, `forward` / `inference` signatures and state-dict key
layout (`block.N...`, `model.N...weight_g/weight_v/bias`, `...act.alpha/beta`,
`...upsample.filter`, `...downsample.lowpass.filter`, `quantizer.*`, `timbre_encoder.*`,
`timbre_linear.*`) as the reference's flamed/models/facodec/facodec.py.  The convolution /
anti-aliased Snake stacks run in sm_100a kernels (flm_codec_encode / flm_codec_decode), and so
do the prompt-side vector quantisers and the timbre transformer (flm_codec_dec_prompt, SURVEY.md
section 8 f3); the nn.Modules below only hold the parameters in the reference's key layout.
Training-only heads of the released decoder checkpoint (f0 / phone / x_timbre predictors) are
accepted and ignored by `load_state_dict`.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from flamed.models.synthesizer._engine import EngineOwner

_HERE = os.path.dirname(__file__)
_TRAINING_ONLY = ("f0_predictor", "phone_predictor", "res_f0_predictor", "res_phone_predictor",
                  "content_f0_predictor", "prosody_phone_predictor", "x_timbre_predictor")


def _kaiser_sinc(cutoff=0.25, half_width=0.3, k=12):
    """12-tap low-pass of the alias-free resamplers (reference alias_free_torch/filter.py:27-58)"""
    half = k // 2
    a = 2.285 * (half - 1) * math.pi * 4 * half_width + 7.95
    beta = 0.1102 * (a - 8.7) if a > 50 else (0.5842 * (a - 21) ** 0.4 + 0.07886 * (a - 21) if a >= 21 else 0.0)
    t = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * torch.kaiser_window(k, beta=beta, periodic=False) * torch.sinc(2 * cutoff * t)
    return (f / f.sum()).view(1, 1, k)


class _WN(nn.Module):
    """weight-normed conv parameters in torch's old-style layout: weight_g, weight_v, bias"""

    def __init__(self, dim0, dim1, k, bias_c, linear=False):
        super().__init__()
        v = torch.empty((dim0, dim1) if linear else (dim0, dim1, k))
        nn.init.kaiming_uniform_(v, a=math.sqrt(5))
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(dim0, *([1] * (v.dim() - 1))))
        self.weight_v = nn.Parameter(v)
        self.bias = nn.Parameter(torch.zeros(bias_c))

    def weight(self):
        v = self.weight_v
        return v * (self.weight_g / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1))))


class _Act(nn.Module):
    """Activation1d(SnakeBeta(log-scale)) parameters"""

    def __init__(self, c):
        super().__init__()
        self.act = nn.Module()
        self.act.alpha = nn.Parameter(torch.zeros(c))
        self.act.beta = nn.Parameter(torch.zeros(c))
        self.upsample = nn.Module()
        self.upsample.register_buffer("filter", _kaiser_sinc())
        self.downsample = nn.Module()
        self.downsample.add_module("lowpass", nn.Module())
        self.downsample.lowpass.register_buffer("filter", _kaiser_sinc())


def _seq(mods):
    m = nn.Module()
    for i, v in enumerate(mods):
        m.add_module(str(i), v)
    return m


def _block(mods):
    m = nn.Module()
    m.add_module("block", _seq(mods))
    return m


def _residual_unit(c):
    return _block([_Act(c), _WN(c, c, 7, c), _Act(c), _WN(c, c, 1, c)])


class FACodecEncoder(EngineOwner):
    default_ckpt = os.path.join(_HERE, "checkpoints", "ns3_facodec_encoder.bin")

    def __init__(self, ngf=32, up_ratios=(2, 4, 5, 5), out_channels=1024):
        super().__init__()
        self.up_ratios = list(up_ratios)
        self.hop_length = int(np.prod(up_ratios))
        d = ngf
        mods = [_WN(d, 1, 7, d)]
        for s in up_ratios:
            d *= 2
            mods.append(_block([_residual_unit(d // 2) for _ in range(3)] + [_Act(d // 2), _WN(d, d // 2, 2 * s, d)]))
        mods += [_Act(d), _WN(out_channels, d, 3, out_channels)]
        self.block = _seq(mods)
        self.enc_dim = d
        self.precision = "fp32"  # the prompt encoder always runs the fp32 FMA kernels

    @classmethod
    def from_pretrained(cls, cfg, ckpt_path=None):
        enc = cls(ngf=cfg["ngf"], up_ratios=cfg["up_ratios"], out_channels=cfg["out_channels"])
        enc.load_state_dict(torch.load(ckpt_path or cls.default_ckpt, map_location=cfg.get("device", "cpu")))
        return enc.eval()

    def _build_engine(self, ctx):
        from flamed_tts_b200.engines import CodecEncoderEngine
        return CodecEncoderEngine(ctx, self.state_dict())

    @torch.inference_mode()
    def forward(self, x):
        """wav (B,1,S) -> (B,out_channels,S/hop); reference facodec.py:215-217"""
        return self.engine().encode(x)

    inference = forward


class _FVQ(nn.Module):
    """parameters of one factorised VQ layer (reference quantize/fvq.py:16-116)"""

    def __init__(self, dim, codebook_size, codebook_dim):
        super().__init__()
        self.in_proj = _WN(codebook_dim, dim, 1, codebook_dim, linear=True)
        self.out_proj = _WN(dim, codebook_dim, 1, dim, linear=True)
        self._codebook = nn.Embedding(codebook_size, codebook_dim)


class _RVQ(nn.Module):
    """parameters of one residual VQ (reference quantize/rvq.py:14-73)"""

    def __init__(self, n, dim, codebook_size, codebook_dim):
        super().__init__()
        self.layers = nn.ModuleList(_FVQ(dim, 2 ** codebook_size, codebook_dim) for _ in range(n))


class _TimbreLayer(nn.Module):
    def __init__(self, d=256, heads=4, filt=1024, k=5):
        super().__init__()
        self.ln_1, self.ln_2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.self_attn = nn.MultiheadAttention(d, heads, batch_first=True)
        self.ffn = nn.Module()
        self.ffn.ffn_1 = nn.Conv1d(d, filt, k, padding=k // 2)
        self.ffn.ffn_2 = nn.Linear(filt, d)


class _TimbreEncoder(nn.Module):
    """parameters of the timbre transformer (reference facodec/transformer.py:154-234: use_cln=False, no token
    embedding; the positional table is indexed by the batch axis, transformer.py:50-52)"""

    def __init__(self, d=256, n_layers=4):
        super().__init__()
        pe = torch.zeros(5000, 1, d)
        pos = torch.arange(5000).unsqueeze(1)
        div = torch.exp(torch.arange(0, d, 2) * (-math.log(10000.0) / d))
        pe[:, 0, 0::2], pe[:, 0, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
        self.position_emb = nn.Module()
        self.position_emb.register_buffer("pe", pe)
        self.layers = nn.ModuleList(_TimbreLayer(d) for _ in range(n_layers))
        self.last_ln = nn.LayerNorm(d)


class FACodecDecoder(EngineOwner):
    default_ckpt = os.path.join(_HERE, "checkpoints", "ns3_facodec_decoder.bin")

    def __init__(self, in_channels=256, upsample_initial_channel=1536, ngf=32, up_ratios=(5, 5, 4, 2), vq_num_q_c=2,
                 vq_num_q_p=1, vq_num_q_r=3, vq_dim=1024, codebook_dim=8, codebook_size_prosody=10,
                 codebook_size_content=10, codebook_size_residual=10, **unused_training_options):
        super().__init__()
        self.up_ratios = list(up_ratios)
        self.hop_length = int(np.prod(up_ratios))
        self.vq_num_q_p, self.vq_num_q_c, self.vq_num_q_r = vq_num_q_p, vq_num_q_c, vq_num_q_r
        self.quantizer = nn.ModuleList([_RVQ(vq_num_q_p, vq_dim, codebook_size_prosody, codebook_dim),
                                        _RVQ(vq_num_q_c, vq_dim, codebook_size_content, codebook_dim)])
        if vq_num_q_r > 0:
            self.quantizer.append(_RVQ(vq_num_q_r, vq_dim, codebook_size_residual, codebook_dim))
        c = upsample_initial_channel
        mods = [_WN(c, in_channels, 7, c)]
        out_dim = c
        for i, s in enumerate(up_ratios):
            in_dim, out_dim = c // 2 ** i, c // 2 ** (i + 1)
            # ConvTranspose1d weight is (in, out, k); weight_norm's dim 0 is the IN channel axis
            mods.append(_block([_Act(in_dim), _WN(in_dim, out_dim, 2 * s, out_dim)] +
                               [_residual_unit(out_dim) for _ in range(3)]))
        mods += [_Act(out_dim), _WN(1, out_dim, 7, 1)]
        self.model = _seq(mods)
        self.timbre_encoder = _TimbreEncoder(in_channels)
        self.timbre_linear = nn.Linear(in_channels, in_channels * 2)
        with torch.no_grad():
            self.timbre_linear.bias[:in_channels] = 1
            self.timbre_linear.bias[in_channels:] = 0

    @classmethod
    def from_pretrained(cls, cfg, ckpt_path=None):
        dec = cls(**{k: v for k, v in dict(cfg).items() if k not in ("ckpt_filename", "device", "checkpoint")})
        dec.load_state_dict(torch.load(ckpt_path or cls.default_ckpt, map_location=cfg.get("device", "cpu")))
        return dec.eval()

    def load_state_dict(self, state_dict, strict=True, **kw):
        kept = {k: v for k, v in state_dict.items() if k.split(".")[0] not in _TRAINING_ONLY}
        return super().load_state_dict(kept, strict=strict, **kw)

    def _build_engine(self, ctx):
        from flamed_tts_b200.engines import CodecDecoderEngine
        return CodecDecoderEngine(ctx, self.state_dict(), precision=self.precision)

    @torch.inference_mode()
    def forward(self, x, vq=True, get_vq=False, eval_vq=True, speaker_embedding=None, n_quantizers=None,
                quantized=None):
        """prompt side (vq=True): enc_out (B,256,T) -> (outs, codes (6,B,T) int64, commit, quantized_buf,
        timbre (B,256)); reference facodec.py:509-533.  Runs in flm_codec_dec_prompt (fp32 kernels)."""
        if get_vq:
            return [layer._codebook.weight for q in self.quantizer for layer in q.layers]
        if not vq:
            raise NotImplementedError("FACodecDecoder.forward(vq=False) is a training path; use .inference()")
        codes, quant, spk = self.engine().prompt(x)
        bufs = [quant[g] for g in range(len(self.quantizer))]
        outs = bufs[0] + bufs[1]
        if len(bufs) > 2:
            outs = outs + bufs[2]
        return outs, codes, torch.zeros(codes.shape[0], x.shape[0], device=codes.device), bufs, spk

    @torch.inference_mode()
    def inference(self, x, speaker_embedding):
        """latents (B,256,T) + timbre (B,256) -> wav (B,1,hop*T); reference facodec.py:630-638"""
        lat = x.transpose(1, 2)  # channels-last; free when x is the view ProbGenerator.sample returns
        return self.engine().decode(lat, speaker_embedding)
