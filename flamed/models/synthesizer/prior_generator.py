"""Prior generator: phoneme encoder -> PVA (B200 kernels) -> shared + per-quantizer FFT decoders.

Drop-in for the reference's flamed/models/synthesizer/prior_generator.py (`sample` signature,
parameter names).  Only `pva` is on the B200 kernel path; the FFT stacks are PyTorch glue
 and run under bf16 autocast when the model precision is 'bf16'
(the phoneme encoder always stays fp32: rounded durations must match the reference).
"""
import torch
import torch.nn as nn

from flamed.models.module import Decoder, Encoder
from flamed.utils.tools import get_mask_from_lengths

from .pva import PVA


class _PreEncoding(nn.Module):
    def __init__(self, hidden_dim, n_quantizer):
        super().__init__()
        self.prompt_emb = nn.Parameter(torch.rand(1, 1, hidden_dim))
        self.target_emb = nn.Parameter(torch.rand(1, 1, hidden_dim))
        self.quantizer_emb = nn.Embedding(n_quantizer, hidden_dim)

    def forward(self, prompt, target, q_idx):
        q = self.quantizer_emb.weight[q_idx]
        return torch.cat([prompt + self.prompt_emb + q, target + self.target_emb + q], dim=1)


class PriorGenerator(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        t = config["transformer"]
        vocab, nq = config["codec"]["vocab_size"], config["codec"]["n_quantizers"]
        self.encoder = Encoder(config)
        self.pva = PVA(config["variance_adaptor"])
        self.bridge = nn.Linear(t["encoder_hidden"], t["decoder_hidden"])
        self.code_embedding = nn.Embedding(vocab + 1, t["decoder_hidden"], padding_idx=vocab)
        self.shared_decoder = Decoder(config, t["decoder_shared_layers"])
        self.pre_encode = _PreEncoding(t["decoder_hidden"], nq)
        self.prior_decoder = nn.ModuleList(Decoder(config, t["decoder_layers"][i]) for i in range(nq))
        self.head = nn.Linear(t["decoder_hidden"], vocab + 1)

    def compute_loss(self, *a, **k):
        raise NotImplementedError("training is out of scope of the B200 inference hot path")

    @torch.inference_mode()
    def sample(self, texts, src_lens, max_src_len, prompts, prompts_len, nfe=4, temperature=1.0):
        """reference prior_generator.py:141-196 -> (embs (B,6,L,384), logits (B,1025,6,L), tgt_mask (B,L))"""
        src_mask = get_mask_from_lengths(src_lens, max_src_len)
        # cuDNN would run the conv-FFNs in TF32 by default; the encoder feeds the duration ODE whose
        # rounded output must match the reference bit for bit, so it always runs IEEE fp32
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            enc = self.encoder(texts, src_mask)                   # fp32 glue
        x, tgt_lens = self.pva.sample(enc, src_lens, src_mask, nfe=nfe, temperature=temperature)  # B200 kernels
        return self.decode_priors(x, tgt_lens, prompts, prompts_len, bf16=self.pva.precision == "bf16" and x.is_cuda)

    @torch.inference_mode()
    def front(self, texts, src_lens, max_src_len, nfe=4, temperature=1.0):
        """first half of `sample`: phoneme encoder -> duration / silence ODEs -> length regulator.
        Returns (x (B,L,192), tgt_lens (B,)); ends with the path's one host synchronisation (L is data dependent)."""
        src_mask = get_mask_from_lengths(src_lens, max_src_len)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            enc = self.encoder(texts, src_mask)
        return self.pva.sample(enc, src_lens, src_mask, nfe=nfe, temperature=temperature)

    @torch.inference_mode()
    def front_plan(self, texts, src_lens, max_src_len, nfe=4, temperature=1.0):
        """`front` without the expand and without any host synchronisation (metadata path with re-bucketing):
        returns (enc (B,P,192), cumsum (B,2P) i32, tgt_lens (B,) i64), all on the device."""
        src_mask = get_mask_from_lengths(src_lens, max_src_len)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            enc = self.encoder(texts, src_mask)
        cumsum, tgt_lens = self.pva.sample_plan(enc, src_lens, src_mask, nfe=nfe, temperature=temperature)
        return enc.float().contiguous(), cumsum, tgt_lens

    @torch.inference_mode()
    def logits_from(self, embs, tgt_mask, bf16=False):
        """the `head` projection of prior_generator.py:179-181 on its own: (B,6,L,384) -> (B,1025,6,L).  The sampling
        path never reads the logits, so Flamed.sample_batch evaluates this lazily (7.5 GB at B=256)."""
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16 and embs.is_cuda):
            logits = self.head(embs) * (~tgt_mask)[:, None, :, None]
        return logits.float().permute(0, 3, 1, 2).contiguous()

    @torch.inference_mode()
    def decode_priors(self, x, tgt_lens, prompts, prompts_len, bf16=False, want_logits=True):
        """length-regulated encoder output (B,L,192) -> (embs, logits, tgt_mask); prior_generator.py:162-181.
        want_logits=False returns None for the logits (see logits_from)."""
        fast = bf16 and x.is_cuda
        self.shared_decoder.b200 = fast
        for d in self.prior_decoder:
            d.b200 = fast
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16), \
                torch.backends.cudnn.flags(enabled=True, allow_tf32=bf16):
            x = self.bridge(x)
            tgt_mask = get_mask_from_lengths(tgt_lens, x.size(1))
            x, _ = self.shared_decoder(x, tgt_mask)
            dec_mask = get_mask_from_lengths(prompts_len + tgt_lens, prompts_len + x.size(1))
            prompt_embs = self.code_embedding(prompts)
            hiddens = []
            for q, layer in enumerate(self.prior_decoder):
                x, _ = layer(self.pre_encode(prompt_embs[:, q], x, q), dec_mask)
                x = x[:, prompts_len:]
                hiddens.append(x)
            out = torch.stack(hiddens, dim=1).float()             # (B, 6, L, 384)
        logits = self.logits_from(out, tgt_mask, bf16) if want_logits else None
        return out, logits, tgt_mask
