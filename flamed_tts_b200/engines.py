"""Torch-facing wrappers around the C ABI handles.

PyTorch is used here only for device memory, streams and dtype plumbing: every tensor
that reaches the library is a raw device pointer, every result is written by the
library's kernels into a torch-allocated output.  One Context per device.
"""
import ctypes
import threading
from ctypes import byref, c_int64, c_void_p

import torch

from . import _lib
from ._lib import FLM_BF16, FLM_F32, check, flm_prob_cfg, pack_weights

_contexts = {}
_lock = threading.Lock()


def _mode(precision):
    if precision in ("fp32", "f32", FLM_F32):
        return FLM_F32
    if precision in ("bf16", FLM_BF16):
        return FLM_BF16
    raise ValueError("precision must be 'fp32' or 'bf16', got %r" % (precision,))


class Context:
    def __init__(self, device):
        self.lib = _lib.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("flamed_b200 runs on CUDA devices only (no CPU fallback); got %s" % (device,))
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.handle = c_void_p()
        check(self.lib.flm_ctx_create(idx, byref(self.handle)))

    @staticmethod
    def get(device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("flamed_b200 runs on CUDA devices only (no CPU fallback); got %s" % (device,))
        idx = device.index if device.index is not None else torch.cuda.current_device()
        with _lock:
            if idx not in _contexts:
                _contexts[idx] = Context(torch.device("cuda", idx))
            return _contexts[idx]

    def stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def profile(self, on):
        """per-launch CUDA-event profiler of the library (launches inside CUDA graphs are not recorded)"""
        check(self.lib.flm_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        n = 8
        buf = (ctypes.c_double * (n * 4))()
        check(self.lib.flm_profile_read(self.handle, buf, n))
        out = {}
        for k in range(n):
            cnt, ms, flops, byts = buf[k * 4:k * 4 + 4]
            if cnt:
                out[self.lib.flm_profile_class_name(k).decode()] = dict(launches=int(cnt), ms=ms, flops=flops, bytes=byts)
        return out

    def profile_detail(self):
        """same records keyed by (class, shape tag) -> list of dict(cls, tag, launches, ms, flops, bytes)"""
        cap = 1 << 20
        buf = ctypes.create_string_buffer(cap)
        check(self.lib.flm_profile_detail(self.handle, buf, cap))
        rows = []
        for line in buf.value.decode().splitlines():
            cls, tag, n, ms, fl, by = line.split("|")
            rows.append(dict(cls=cls, tag=tag, launches=int(float(n)), ms=float(ms), flops=float(fl), bytes=float(by)))
        return rows


def _f32(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p()


def _strip(sd, prefix):
    return [(k[len(prefix):], v) for k, v in sd.items() if k.startswith(prefix)]


class DurationEngine:
    """PVA.sample loop + rounding (pva.py:88-112) and the length regulator (pva.py:125-166)."""

    def __init__(self, ctx, pva_state_dict):
        self.ctx, self.lib = ctx, ctx.lib
        named = [(k, v) for k, v in pva_state_dict.items()
                 if k.startswith("duration_generator.") or k.startswith("sil_generator.")]
        arr, n, keep = pack_weights(named)
        self.handle = c_void_p()
        check(self.lib.flm_durgen_load(ctx.handle, arr, n, byref(self.handle)))

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.flm_durgen_destroy(self.handle)
            self.handle = None

    def sample(self, enc, src_mask, noise_dur, noise_sil, ts, temperature, seed=0):
        """noise_dur / noise_sil: (B,P) standard-normal draws, or both None: drawn inside the library from `seed`
        (documented Philox map, include/flamed_b200.h)"""
        dev = self.ctx.device
        enc = _f32(enc, dev)
        B, P, _ = enc.shape
        if noise_dur is not None:
            noise_dur, noise_sil = _f32(noise_dur, dev), _f32(noise_sil, dev)
        mask = src_mask.to(device=dev, dtype=torch.uint8).contiguous()
        ts = ts.detach().to(device="cpu", dtype=torch.float32).contiguous()
        nfe = ts.numel() - 1
        phone = torch.empty((B, P), device=dev, dtype=torch.float32)
        sil = torch.empty_like(phone)
        dur_t = torch.empty_like(phone)
        sil_t = torch.empty_like(phone)
        check(self.lib.flm_durgen_sample(self.handle, _ptr(enc), _ptr(noise_dur), _ptr(noise_sil), int(seed), _ptr(mask),
                                         c_void_p(ts.data_ptr()), nfe, float(temperature), B, P, _ptr(phone),
                                         _ptr(sil), _ptr(dur_t), _ptr(sil_t), self.ctx.stream()))
        return phone, sil, dur_t, sil_t

    def forward(self, which, x, enc, t, src_mask=None):
        """one ProbabilisticModule.forward evaluation (pva.py:221-238): which 0 = duration, 1 = silence generator"""
        dev = self.ctx.device
        x, enc = _f32(x, dev), _f32(enc, dev)
        B, P, _ = enc.shape
        mask = src_mask.to(device=dev, dtype=torch.uint8).contiguous() if src_mask is not None else None
        out = torch.empty((B, P), device=dev, dtype=torch.float32)
        check(self.lib.flm_durgen_forward(self.handle, int(which), _ptr(x), _ptr(enc), float(t), _ptr(mask), B, P,
                                          _ptr(out), self.ctx.stream()))
        return out

    def plan(self, phone_dur, sil_dur, src_lens, sync=True):
        """integer plan of the length regulator -> (cumsum (B,2P) i32, tgt_len (B,) i64 on device, Tmax or None)"""
        dev = self.ctx.device
        phone_dur, sil_dur = _f32(phone_dur, dev), _f32(sil_dur, dev)
        B, P = phone_dur.shape
        src_lens = src_lens.to(device=dev, dtype=torch.int64).contiguous()
        cumsum = torch.empty((B, 2 * P), device=dev, dtype=torch.int32)
        tgt_len = torch.empty((B,), device=dev, dtype=torch.int64)
        tmax = c_int64(0)
        check(self.lib.flm_lr_plan(self.ctx.handle, _ptr(phone_dur), _ptr(sil_dur), _ptr(src_lens), B, P,
                                   _ptr(cumsum), _ptr(tgt_len), byref(tmax) if sync else None, self.ctx.stream()))
        return cumsum, tgt_len, (int(tmax.value) if sync else None)

    def length_regulate(self, x, phone_dur, sil_dur, src_lens, return_index=False):
        dev = self.ctx.device
        x = _f32(x, dev)
        B, P, H = x.shape
        cumsum, tgt_len, T = self.plan(phone_dur, sil_dur, src_lens, sync=True)
        out = torch.empty((B, T, H), device=dev, dtype=torch.float32)
        index = torch.empty((B, T), device=dev, dtype=torch.int32) if return_index else None
        check(self.lib.flm_lr_expand(self.ctx.handle, _ptr(x), _ptr(cumsum), B, P, H, T, _ptr(out), _ptr(index),
                                     self.ctx.stream()))
        if return_index:
            return out, tgt_len, index
        return out, tgt_len

    def expand_gather(self, sources, Tmax):
        """sources: list over the B samples of the new batch of (x (Pmax,H) f32 row block of that sample, cumsum (2 Pmax)
        i32 of that sample), both views into tensors of earlier planned batches -> (B,Tmax,H) f32, zero-padded"""
        dev = self.ctx.device
        B = len(sources)
        H = sources[0][0].shape[-1]
        table = torch.tensor([[x.data_ptr() for x, _ in sources], [c.data_ptr() for _, c in sources]], dtype=torch.int64)
        ps = torch.tensor([x.shape[0] for x, _ in sources], dtype=torch.int32)
        table, ps = table.to(dev, non_blocking=True), ps.to(dev, non_blocking=True)
        out = torch.empty((B, Tmax, H), device=dev, dtype=torch.float32)
        check(self.lib.flm_lr_expand_gather(self.ctx.handle, _ptr(table[0]), _ptr(table[1]), _ptr(ps), B, H, int(Tmax),
                                            _ptr(out), None, self.ctx.stream()))
        return out


class DenoiserEngine:
    """ProbGenerator.sample (prob_generator.py:434-446) behind flm_cond_prepare / flm_denoiser_sample."""

    def __init__(self, ctx, prob_state_dict, cfg, precision="bf16"):
        self.ctx, self.lib = ctx, ctx.lib
        self.precision = precision
        c = flm_prob_cfg()
        c.target_dim, c.spk_dim, c.cond_dim = int(cfg["target_dim"]), int(cfg["spk_dim"]), int(cfg["cond_dim"])
        c.hidden_dim, c.n_layers, c.n_quantizers = int(cfg["hidden_dim"]), int(cfg["n_layers"]), int(cfg["n_quantizers"])
        c.kernel_size = int(cfg["convnext"]["kernel_size"])
        c.downsampling_stages = int(cfg["downsampling_stages"])
        if cfg["convnext"].get("groups") is not None or int(cfg["convnext"].get("expand", 1)) != 1 \
                or int(cfg["convnext"].get("stride", 1)) != 1:
            raise RuntimeError("only depthwise ConvNeXt (groups: null, expand 1, stride 1) is implemented")
        self.cfg = c
        arr, n, keep = pack_weights(list(prob_state_dict.items()))
        self.handle = c_void_p()
        check(self.lib.flm_denoiser_load(ctx.handle, arr, n, byref(c), _mode(precision), byref(self.handle)))

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.flm_denoiser_destroy(self.handle)
            self.handle = None

    @property
    def launches_per_step(self):
        return int(self.lib.flm_denoiser_launches_per_step(self.handle))

    def cond_prepare(self, prior_embs, mask):
        """prior_embs (B,Q,L,cond_dim); mask (B,L[,1]) bool True = valid -> cond (B,L,target_dim) fp32"""
        dev = self.ctx.device
        prior_embs = _f32(prior_embs, dev)
        B, Q, L, _ = prior_embs.shape
        mask = mask.reshape(B, L).to(device=dev, dtype=torch.uint8).contiguous()
        out = torch.empty((B, L, self.cfg.target_dim), device=dev, dtype=torch.float32)
        check(self.lib.flm_cond_prepare(self.handle, _ptr(prior_embs), _ptr(mask), B, L, _ptr(out), self.ctx.stream()))
        return out

    def sample(self, cond, spk, noise, ts, temperature, use_graph=True, seed=0):
        """returns latents channels-last (B,L,D); the reference's (B,D,L) is `.transpose(1,2)` of it.
        noise None: x0 = temperature * N(0,1) + cond is drawn inside the init kernel from `seed` (Philox map)"""
        dev = self.ctx.device
        cond, spk = _f32(cond, dev), _f32(spk, dev)
        noise = _f32(noise, dev) if noise is not None else None
        B, L, D = cond.shape
        ts = ts.detach().to(device="cpu", dtype=torch.float32).contiguous()
        nfe = ts.numel() - 1
        out = torch.empty((B, L, D), device=dev, dtype=torch.float32)
        check(self.lib.flm_denoiser_sample(self.handle, _ptr(cond), _ptr(spk), _ptr(noise), int(seed),
                                           c_void_p(ts.data_ptr()), B, L, nfe, float(temperature), _ptr(out),
                                           1 if use_graph else 0, self.ctx.stream()))
        return out

    def forward(self, x, t, spk):
        dev = self.ctx.device
        x, spk = _f32(x, dev), _f32(spk, dev)
        B, L, D = x.shape
        out = torch.empty_like(x)
        check(self.lib.flm_denoiser_forward(self.handle, _ptr(x), _ptr(spk), float(t), B, L, _ptr(out),
                                            self.ctx.stream()))
        return out


class CodecDecoderEngine:
    """FACodecDecoder.inference (facodec.py:630-638)."""

    def __init__(self, ctx, state_dict, precision="bf16"):
        self.ctx, self.lib = ctx, ctx.lib
        named = [(k, v) for k, v in state_dict.items() if k.split(".")[0] in ("model", "timbre_linear", "quantizer",
                                                                              "timbre_encoder")]
        arr, n, keep = pack_weights(named)
        self.handle = c_void_p()
        check(self.lib.flm_codec_dec_load(ctx.handle, arr, n, _mode(precision), byref(self.handle)))
        self.n_q = len([k for k in state_dict if k.startswith("quantizer.") and k.endswith("._codebook.weight")])

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.flm_codec_dec_destroy(self.handle)
            self.handle = None

    def decode(self, latents_bld, spk):
        """latents (B,L,256) channels-last, spk (B,256) -> wav (B,1,200L)"""
        dev = self.ctx.device
        latents_bld, spk = _f32(latents_bld, dev), _f32(spk, dev)
        B, L, _ = latents_bld.shape
        hop = 200
        wav = torch.empty((B, 1, L * hop), device=dev, dtype=torch.float32)
        check(self.lib.flm_codec_decode(self.handle, _ptr(latents_bld), _ptr(spk), B, L, _ptr(wav), self.ctx.stream()))
        return wav

    def prompt(self, enc_out):
        """FACodecDecoder.forward(vq=True): enc_out (B,256,T) -> codes (n_q,B,T) i64, quantized (3,B,256,T), spk (B,256)"""
        dev = self.ctx.device
        enc_out = _f32(enc_out, dev)
        B, D, T = enc_out.shape
        codes = torch.empty((self.n_q, B, T), device=dev, dtype=torch.int64)
        quant = torch.empty((3, B, D, T), device=dev, dtype=torch.float32)
        spk = torch.empty((B, D), device=dev, dtype=torch.float32)
        check(self.lib.flm_codec_dec_prompt(self.handle, _ptr(enc_out), B, T, _ptr(codes), _ptr(quant), _ptr(spk),
                                            self.ctx.stream()))
        return codes, quant, spk

    def activation(self, prefix, x_btc):
        dev = self.ctx.device
        x_btc = _f32(x_btc, dev)
        B, T, C = x_btc.shape
        y = torch.empty_like(x_btc)
        check(self.lib.flm_codec_dec_activation(self.handle, prefix.encode(), _ptr(x_btc), B, T, C, _ptr(y),
                                                self.ctx.stream()))
        return y


class CodecEncoderEngine:
    """FACodecEncoder.forward (facodec.py:215-217)."""

    def __init__(self, ctx, state_dict):
        self.ctx, self.lib = ctx, ctx.lib
        arr, n, keep = pack_weights(list(state_dict.items()))
        self.handle = c_void_p()
        check(self.lib.flm_codec_enc_load(ctx.handle, arr, n, byref(self.handle)))

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.flm_codec_enc_destroy(self.handle)
            self.handle = None

    def encode(self, wav):
        """wav (B,1,S) -> (B,256,T') in the reference layout"""
        dev = self.ctx.device
        wav = _f32(wav, dev)
        B, _, S = wav.shape
        T = int(self.lib.flm_codec_enc_frames(self.handle, S))
        if T <= 0:
            raise ValueError("prompt of %d samples is too short for the FaCodec encoder" % S)
        out = torch.empty((B, 256, T), device=dev, dtype=torch.float32)
        check(self.lib.flm_codec_encode(self.handle, _ptr(wav), B, S, _ptr(out), self.ctx.stream()))
        return out


def wav_to_pcm16(ctx, wav):
    """fp32 waveform tensor (any shape, contiguous) -> int16 PCM of the same shape on the device: lrintf(x * 32767),
    what soundfile stores for the reference's sf.write(path, wav, 16000) (synthesize.py:296)"""
    wav = _f32(wav, ctx.device)
    out = torch.empty(wav.shape, device=ctx.device, dtype=torch.int16)
    check(ctx.lib.flm_wav_to_pcm16(ctx.handle, _ptr(wav), wav.numel(), _ptr(out), ctx.stream()))
    return out


def philox_normal(ctx, seed, tensor_id, n):
    """the library's documented seed -> N(0,1) map (tests)"""
    out = torch.empty((n,), device=ctx.device, dtype=torch.float32)
    check(ctx.lib.flm_philox_normal(ctx.handle, int(seed), int(tensor_id), int(n), _ptr(out), ctx.stream()))
    return out


def tapgemm(ctx, precision, A, W, bias, T_out, ntaps, off0, dil, stride, epi):
    """test hook: A (B,T_in,K), W (ntaps,N,K), bias (N) or None -> (B,T_out,N)"""
    dev = ctx.device
    A, W = _f32(A, dev), _f32(W, dev)
    bias = _f32(bias, dev) if bias is not None else None
    B, T_in, K = A.shape
    N = W.shape[1]
    out = torch.empty((B, T_out, N), device=dev, dtype=torch.float32)
    check(ctx.lib.flm_tapgemm_test(ctx.handle, _mode(precision), _ptr(A), _ptr(W), _ptr(bias), B, T_in, T_out, K, N,
                                   ntaps, off0, dil, stride, epi, _ptr(out), ctx.stream()))
    return out


# ------------------------------------------------------------------------------------------------ generic bf16 ops
def conv1d_bf16(ctx, a, w, bias, epi=0, off0=0, dil=1, resid=None, out=None):
    """a (B,T,K) bf16, w (ntaps,N,K) bf16 packed tap-major, bias (N) f32 or None -> (B,T,N) bf16 through the tcgen05
    implicit-conv GEMM (flm_conv1d_bf16).  epi: 0 none, 1 gelu, 2 silu, 3 relu, 4 out = resid + v."""
    B, T, K = a.shape
    ntaps, N, _ = w.shape
    if out is None:
        out = torch.empty((B, T, N), device=a.device, dtype=torch.bfloat16)
    check(ctx.lib.flm_conv1d_bf16(ctx.handle, _ptr(a), _ptr(w), _ptr(bias), B, T, K, N, ntaps, off0, dil, epi, _ptr(out),
                                  _ptr(resid), ctx.stream()))
    return out


def attention_bf16(ctx, qkv, key_lens):
    """qkv (B,S,3,H,32) bf16 contiguous, key_lens (B,) int32 -> (B,S,H*32) bf16: softmax(q k^T / sqrt(32)) v over the
    first key_lens[b] keys of every sample (flm_attention_bf16)"""
    B, S, three, H, dh = qkv.shape
    out = torch.empty((B, S, H * dh), device=qkv.device, dtype=torch.bfloat16)
    check(ctx.lib.flm_attention_bf16(ctx.handle, _ptr(qkv), _ptr(key_lens), B, S, H, dh, _ptr(out), ctx.stream()))
    return out


def layernorm_bf16(ctx, x, w, b, eps, zero_rows=None, out=None):
    """row LayerNorm of x (..., C) bf16 with fp32 affine; rows flagged in zero_rows (uint8/bool, one per row) -> 0"""
    C = x.shape[-1]
    rows = x.numel() // C
    if out is None:
        out = torch.empty_like(x)
    if zero_rows is not None and zero_rows.dtype != torch.uint8:
        zero_rows = zero_rows.to(torch.uint8)
    check(ctx.lib.flm_layernorm_bf16(ctx.handle, _ptr(x), _ptr(w), _ptr(b), float(eps), rows, C, _ptr(zero_rows), _ptr(out),
                                     ctx.stream()))
    return out
