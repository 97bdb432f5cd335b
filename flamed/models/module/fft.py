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

    def forward(self, x, pad_mask):
        x = self.slf_attn(x, pad_mask).masked_fill(pad_mask.unsqueeze(-1), 0)
        return self.pos_ffn(x).masked_fill(pad_mask.unsqueeze(-1), 0)


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

    def _run(self, x, pad_mask):
        x = x + self._positions(x.shape[1], x.device).to(x.dtype)
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
