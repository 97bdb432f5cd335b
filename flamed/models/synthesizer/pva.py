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

    def forward(self, xt, encoder_output, t, mask=None):
        """one vector-field evaluation (reference pva.py:221-238): xt (B,P), encoder_output (B,P,192), scalar t,
        mask (B,P) bool True = padding -> v (B,P).  Runs the kernels of one step of the fused sampling loop."""
        owner = self.__dict__.get("_owner")
        if owner is None:
            raise RuntimeError("ProbabilisticModule.forward needs its PVA owner (no CPU/PyTorch fallback on the B200 path)")
        return owner.engine().forward(self.__dict__["_which"], xt, encoder_output, float(t), mask)


class LengthRegulator(nn.Module):
    def __init__(self, owner=None):
        super().__init__()
        self.__dict__["_owner"] = owner  # not a sub-module: avoids a reference cycle in the module tree

    def LR(self, x, phone_duration, sil_duration, src_lens, max_len=None):
        """(B,P,H), (B,P), (B,P), (B,) -> ((B,Tmax,H), tgt_len (B,) int64); reference pva.py:125-166."""
        out, tgt_len = self._owner.engine().length_regulate(x, phone_duration, sil_duration, src_lens)
        if max_len is not None and out.shape[1] != max_len:
            # the reference pads with F.pad(batch, (0, 0, 0, max_len - len)) (tools.py:299-317): a negative amount
            # TRUNCATES to max_len; tgt_len is returned unclipped, as the reference does
            if out.shape[1] > max_len:
                out = out[:, :max_len].contiguous()
            else:
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
        for which, m in enumerate((self.duration_generator, self.sil_generator)):
            m.__dict__["_owner"], m.__dict__["_which"] = self, which  # not sub-modules: no cycle in the module tree
        self.length_regulator = LengthRegulator(self)
        # 'cpu' = the reference's CPU default-generator draws; 'cuda' = torch's device generator; 'philox' = drawn
        # inside the library's own init kernels from a 63-bit seed taken from torch's CPU generator (documented map)
        self.noise_device = "cpu"

    def _build_engine(self, ctx):
        from flamed_tts_b200.engines import DurationEngine
        return DurationEngine(ctx, self.state_dict())

    def compute_loss(self, *a, **k):
        raise NotImplementedError("training is out of scope of the B200 inference hot path")

    def _draw(self, x):
        """the two (B,P) standard-normal draws (duration first, pva.py:101-102) -> (n_dur, n_sil, seed)"""
        b, l, _ = x.shape
        if self.noise_device == "philox":
            return None, None, int(torch.randint(0, 2 ** 62, (1,)).item())
        ndev = x.device if self.noise_device == "cuda" else "cpu"
        return torch.randn((b, l), device=ndev), torch.randn((b, l), device=ndev), 0

    @torch.inference_mode()
    def sample(self, x, src_len, src_mask, max_tgt_len=None, nfe=32, temperature=1.0, return_durations=False):
        """reference pva.py:88-116.  Noise: two (B,P) standard-normal draws, duration first."""
        ts = torch.linspace(0, 1, nfe + 1)
        n_dur, n_sil, seed = self._draw(x)
        eng = self.engine()
        phone, sil, dur_t, sil_t = eng.sample(x, src_mask, n_dur, n_sil, ts, temperature, seed=seed)
        out, tgt_len = self.length_regulator(x, phone, sil, src_len, max_tgt_len)
        if return_durations:
            return out, tgt_len, dict(phone=phone, sil=sil, dur_t=dur_t, sil_t=sil_t)
        return out, tgt_len

    @torch.inference_mode()
    def sample_plan(self, x, src_len, src_mask, nfe=32, temperature=1.0):
        """`sample` without the expand and WITHOUT a host synchronisation: duration / silence ODEs, rounding and the
        integer plan of the length regulator.  Returns (cumsum (B,2P) i32, tgt_len (B,) i64), both on the device; the
        caller expands later into batches of its own choosing (LengthRegulator.expand_gather)."""
        ts = torch.linspace(0, 1, nfe + 1)
        n_dur, n_sil, seed = self._draw(x)
        eng = self.engine()
        phone, sil, _, _ = eng.sample(x, src_mask, n_dur, n_sil, ts, temperature, seed=seed)
        cumsum, tgt_len, _ = eng.plan(phone, sil, src_len, sync=False)
        return cumsum, tgt_len
