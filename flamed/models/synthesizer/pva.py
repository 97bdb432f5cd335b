"""Probabilistic variance adaptor: duration / silence flow-matching generators + length regulator.

Drop-in for the reference's flamed/models/synthesizer/pva.py (same module tree, parameter
names and `sample` / `LR` signatures).  The sampling loops and the integer length regulator
run in hand-written sm_100a kernels behind the C ABI (flm_durgen_sample, flm_lr_plan,
flm_lr_expand); this file only holds the parameters and moves tensors.
"""
import torch
import torch.nn as nn

from ._engine import EngineOwner


def _children(**mods):
    m = nn.Module()
    for k, v in mods.items():
        m.add_module(k, v)
    return m


class ProbabilisticModule(nn.Module):
    """Parameter holder for one vector-field network (reference pva.py:173-238):
    proj(193->192) + time MLP, two k=3 convs each followed by ReLU+LayerNorm, Linear(384->1)."""

    def __init__(self, cfg):
        super().__init__()
        d, f, k, ts = cfg["input_size"], cfg["filter_size"], cfg["kernel_size"], cfg["time_scale"]
        self.proj = nn.Linear(d + 1, d)
        self.time_emb = _children(time_emb=_children(**{"1": nn.Linear(d, d * ts), "3": nn.Linear(d * ts, d)}))
        self.conv_layer = _children(
            conv1d_1=_children(conv=nn.Conv1d(d, f, k, padding=(k - 1) // 2)), layer_norm_1=nn.LayerNorm(f),
            conv1d_2=_children(conv=nn.Conv1d(f, f, k, padding=1)), layer_norm_2=nn.LayerNorm(f))
        self.linear_layer = nn.Linear(f, 1)

    def forward(self, *a, **k):
        raise NotImplementedError("single vector-field evaluations are fused into PVA.sample on the B200 path")


class LengthRegulator(nn.Module):
    def __init__(self, owner=None):
        super().__init__()
        self.__dict__["_owner"] = owner  # not a sub-module: avoids a reference cycle in the module tree

    def LR(self, x, phone_duration, sil_duration, src_lens, max_len=None):
        """(B,P,H), (B,P), (B,P), (B,) -> ((B,Tmax,H), tgt_len (B,) int64); reference pva.py:125-166."""
        out, tgt_len = self._owner.engine().length_regulate(x, phone_duration, sil_duration, src_lens)
        if max_len is not None and out.shape[1] != max_len:
            if out.shape[1] > max_len:
                raise ValueError("max_len %d is shorter than the regulated length %d" % (max_len, out.shape[1]))
            out = torch.nn.functional.pad(out, (0, 0, 0, max_len - out.shape[1]))
        return out, tgt_len

    def forward(self, x, phone_duration, sil_duration, src_lens, max_len=None):
        return self.LR(x, phone_duration, sil_duration, src_lens, max_len)


class PVA(EngineOwner):
    def __init__(self, model_config):
        super().__init__()
        self.sigma_min = float(model_config["sigma_min"])
        self.duration_generator = ProbabilisticModule(model_config["duration_generator"])
        self.sil_generator = ProbabilisticModule(model_config["sil_generator"])
        self.length_regulator = LengthRegulator(self)
        self.noise_device = "cpu"  # 'cpu' = the reference's CPU default-generator draws; 'cuda' = on-device

    def _build_engine(self, ctx):
        from flamed_tts_b200.engines import DurationEngine
        return DurationEngine(ctx, self.state_dict())

    def compute_loss(self, *a, **k):
        raise NotImplementedError("training is out of scope of the B200 inference hot path")

    @torch.inference_mode()
    def sample(self, x, src_len, src_mask, max_tgt_len=None, nfe=32, temperature=1.0, return_durations=False):
        """reference pva.py:88-116.  Noise: two (B,P) standard-normal draws, duration first."""
        b, l, _ = x.shape
        ts = torch.linspace(0, 1, nfe + 1)
        ndev = x.device if self.noise_device == "cuda" else "cpu"
        n_dur = torch.randn((b, l), device=ndev)
        n_sil = torch.randn((b, l), device=ndev)
        eng = self.engine()
        phone, sil, dur_t, sil_t = eng.sample(x, src_mask, n_dur, n_sil, ts, temperature)
        out, tgt_len = self.length_regulator(x, phone, sil, src_len, max_tgt_len)
        if return_durations:
            return out, tgt_len, dict(phone=phone, sil=sil, dur_t=dur_t, sil_t=sil_t)
        return out, tgt_len
