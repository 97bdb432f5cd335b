"""FFT (feed-forward transformer) stacks of the prior generator - host-side PyTorch glue.

Not part of the B200 kernel scope: attention-based, runs once per
utterance between the two sampling loops.  Parameter names follow the reference so that its
checkpoints load unchanged (flamed/models/module/transformer/{Models,Layers,SubLayers}.py);
the math is restated with fused attention (F.scaled_dot_product_attention) instead of
materialised (heads*B, L, L) score tensors.
"""

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from flamed.text.symbols import symbols


def sinusoid_table(n_position, d_hid):
    """pos / 10000^(2*(j//2)/d): sin on even columns, cos on odd (Models.py:10-30)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000.0, 2 * (j // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.FloatTensor(ang)


class _SelfAttention(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.n_head = n_head
        self.w_qs = nn.Linear(d_model, d_model)
        self.w_ks = nn.Linear(d_model, d_model)
        self.w_vs = nn.Linear(d_model, d_model)
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(d_model, d_model)

    def forward(self, x, key_pad):
        B, L, D = x.shape
        split = lambda t: t.view(B, L, self.n_head, D // self.n_head).transpose(1, 2)
        bias = torch.zeros((B, 1, 1, L), dtype=x.dtype, device=x.device).masked_fill(key_pad[:, None, None, :], float("-inf"))
        o = F.scaled_dot_product_attention(split(self.w_qs(x)), split(self.w_ks(x)), split(self.w_vs(x)), attn_mask=bias)
        o = self.fc(o.transpose(1, 2).reshape(B, L, D))
        return self.layer_norm(o + x)


class _ConvFFN(nn.Module):
    def __init__(self, d_in, d_hid, kernel_size):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, kernel_size[0], padding=(kernel_size[0] - 1) // 2)
        self.w_2 = nn.Conv1d(d_hid, d_in, kernel_size[1], padding=(kernel_size[1] - 1) // 2)
        self.layer_norm = nn.LayerNorm(d_in)

    def forward(self, x):
        y = self.w_2(F.relu(self.w_1(x.transpose(1, 2)))).transpose(1, 2)
        return self.layer_norm(y + x)


class FFTBlock(nn.Module):
    def __init__(self, d_model, n_head, d_inner, kernel_size):
        super().__init__()
        self.slf_attn = _SelfAttention(d_model, n_head)
        self.pos_ffn = _ConvFFN(d_model, d_inner, kernel_size)
        self._packed = None

    def forward(self, x, pad_mask):
        x = self.slf_attn(x, pad_mask).masked_fill(pad_mask.unsqueeze(-1), 0)
        return self.pos_ffn(x).masked_fill(pad_mask.unsqueeze(-1), 0)

    # ---- B200 path (bf16 mode): projections / conv-FFN / LayerNorm on the library's tcgen05 GEMM and row-LN
    # kernels (flm_conv1d_bf16, flm_layernorm_bf16), attention through torch's fused SDPA (library kernel)
    def _pack(self, device):
        a, f = self.slf_attn, self.pos_ffn
        key = (str(device), a.w_qs.weight._version, f.w_1.weight._version, a.w_qs.weight.data_ptr())
        if self._packed is None or self._packed["key"] != key:
            bf = lambda t: t.detach().to(device=device, dtype=torch.bfloat16).contiguous()
            f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
            self._packed = dict(
                key=key,
                wqkv=bf(torch.cat([a.w_qs.weight, a.w_ks.weight, a.w_vs.weight], 0)).unsqueeze(0),  # (1, 3D, D)
                bqkv=f32(torch.cat([a.w_qs.bias, a.w_ks.bias, a.w_vs.bias], 0)),
                wfc=bf(a.fc.weight).unsqueeze(0), bfc=f32(a.fc.bias),
                ln1w=f32(a.layer_norm.weight), ln1b=f32(a.layer_norm.bias),
                w1=bf(f.w_1.weight.permute(2, 0, 1)), b1=f32(f.w_1.bias),     # (k, d_hid, d_in) tap-major
                w2=bf(f.w_2.weight.permute(2, 0, 1)), b2=f32(f.w_2.bias),
                ln2w=f32(f.layer_norm.weight), ln2b=f32(f.layer_norm.bias),
                k1=f.w_1.kernel_size[0], k2=f.w_2.kernel_size[0])
        return self._packed

    def forward_b200(self, ctx, x, pad_u8, key_lens):
        """x (B,S,D) bf16 contiguous, pad_u8 (B,S) uint8 1 = padding, key_lens (B,) int32 valid prefix length.
        Every op is one of the library's kernels: fused q|k|v projection and the other projections / conv-FFN on the
        tcgen05 implicit-conv GEMM, attention with per-sample key prefixes (flm_attention_bf16), row LayerNorm with
        the masked_fill of the padding folded in."""
        from flamed_tts_b200.engines import attention_bf16, conv1d_bf16, layernorm_bf16
        w = self._pack(x.device)
        B, S, D = x.shape
        H = self.slf_attn.n_head
        qkv = conv1d_bf16(ctx, x, w["wqkv"], w["bqkv"]).view(B, S, 3, H, D // H)
        o = attention_bf16(ctx, qkv, key_lens)
        y = conv1d_bf16(ctx, o, w["wfc"], w["bfc"], epi=4, resid=x)                # fc(o) + x
        x = layernorm_bf16(ctx, y, w["ln1w"], w["ln1b"], self.slf_attn.layer_norm.eps, zero_rows=pad_u8, out=y)
        h = conv1d_bf16(ctx, x, w["w1"], w["b1"], epi=3, off0=-(w["k1"] // 2))      # relu(conv k)
        y = conv1d_bf16(ctx, h, w["w2"], w["b2"], epi=4, off0=-(w["k2"] // 2), resid=x)
        return layernorm_bf16(ctx, y, w["ln2w"], w["ln2b"], self.pos_ffn.layer_norm.eps, zero_rows=pad_u8, out=y)


class _Stack(nn.Module):
    def __init__(self, max_len, d_model, n_layers, n_head, d_inner, kernel_size):
        super().__init__()
        self.max_seq_len, self.d_model = max_len, d_model
        self.position_enc = nn.Parameter(sinusoid_table(max_len + 1, d_model).unsqueeze(0), requires_grad=False)
        self.layer_stack = nn.ModuleList(FFTBlock(d_model, n_head, d_inner, kernel_size) for _ in range(n_layers))

    def _positions(self, L, device):
        if L > self.max_seq_len:  # rebuild the table for over-long sequences (Models.py:82-87,145-152)
            return sinusoid_table(L, self.d_model).unsqueeze(0).to(device)
        return self.position_enc[:, :L]

    b200 = False  # set by the owner (PriorGenerator) in bf16 mode: run the blocks on the library's kernels

    def _run(self, x, pad_mask):
        x = x + self._positions(x.shape[1], x.device).to(x.dtype)
        if self.b200 and x.is_cuda and x.shape[-1] % 128 == 0 and x.shape[-1] // self.layer_stack[0].slf_attn.n_head == 32:
            from flamed_tts_b200.engines import Context
            ctx = Context.get(x.device)
            out_dtype = x.dtype
            x = x.to(torch.bfloat16).contiguous()
            pad_u8 = pad_mask.to(torch.uint8).contiguous()
            key_lens = (~pad_mask).sum(1).to(torch.int32)  # get_mask_from_lengths masks are prefix masks
            for blk in self.layer_stack:
                x = blk.forward_b200(ctx, x, pad_u8, key_lens)
            return x.to(out_dtype) if out_dtype != torch.bfloat16 else x
        for blk in self.layer_stack:
            x = blk(x, pad_mask)
        return x


class Encoder(_Stack):
    """phoneme encoder (Models.py:33-104)"""

    def __init__(self, config):
        t = config["transformer"]
        super().__init__(t["encoder_max_seq_len"], t["encoder_hidden"], t["encoder_layer"], t["encoder_head"],
                         t["encoder_conv_filter_size"], t["encoder_conv_kernel_size"])
        self.src_word_emb = nn.Embedding(len(symbols) + 1, t["encoder_hidden"], padding_idx=0)

    def forward(self, src_seq, mask):
        return self._run(self.src_word_emb(src_seq), mask)


class Decoder(_Stack):
    """frame-level FFT decoder (Models.py:107-171)"""

    def __init__(self, config, n_layers):
        t = config["transformer"]
        super().__init__(t["decoder_max_seq_len"], t["decoder_hidden"], n_layers, t["decoder_head"],
                         t["decoder_conv_filter_size"], t["decoder_conv_kernel_size"])

    def forward(self, enc_seq, mask):
        return self._run(enc_seq, mask), mask
