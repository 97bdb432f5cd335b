"""Import the UNMODIFIED reference: /root/reference in the build container, or its verbatim git-ignored copy
`oracle/_ref/` (made by oracle/make_ref.py, travels to the GPU box with the gpurun snapshot).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python/PyTorch but imports a
dozen packages that are absent here (librosa, lightning, g2p_en, ...); none of
them is touched by the inference hot path, so they are replaced by empty stub
modules before `import flamed` (recipe: SURVEY.md Appendix B).  Users: `oracle/make_golden.py`,
tests that pin the oracle port against the live reference, and bench.py's `--impl reference` (CPU) and
`--impl eager` (the reference's own PyTorch code on the same B200) arms.  Never a product path.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("FLAMED_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/flamed") else os.path.join(_HERE, "_ref"))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "flamed"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference `flamed` package (imported from REFERENCE_ROOT)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch.nn as nn

    class _LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                import torch
                return torch.device("cpu")

    _stub("librosa", load=None)
    _stub("librosa.filters", mel=None)
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    _stub("soundfile")
    _stub("pyworld")
    _stub("g2p_en", G2p=object)
    _stub("tgt")
    _stub("lightning", LightningModule=_LightningModule, LightningDataModule=object)
    _stub("pytorch_lightning")
    _stub("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
    _stub("matplotlib", use=lambda *a, **k: None)
    _stub("matplotlib.pyplot")
    _stub("unidecode", unidecode=lambda s: s)
    _stub("inflect", engine=lambda: None)
    _stub("omegaconf", DictConfig=dict, OmegaConf=object)
    _stub("wandb")
    _stub("transformers", get_cosine_schedule_with_warmup=None)
    # our own drop-in package is also called `flamed`: make sure the reference wins
    for k in [k for k in sys.modules if k == "flamed" or k.startswith("flamed.")]:
        del sys.modules[k]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    mods = {}
    try:
        import flamed  # noqa
        import flamed.models.facodec  # noqa
        mods = {k: v for k, v in sys.modules.items() if k == "flamed" or k.startswith("flamed.")}
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # leave the name `flamed` free for the drop-in package
        for k in list(mods):
            del sys.modules[k]
    return mods


def build_reference_models(flamed_sd, codec_dec_sd=None, codec_enc_sd=None, device="cpu"):
    """The reference's own modules with the given (seeded) weights loaded: (cfg, Flamed, FACodecEncoder|None,
    FACodecDecoder|None), in eval mode on `device`.  Construction arguments of the codec: synthesize.py:46-69."""
    import torch
    import yaml
    mods = import_reference()
    with open(os.path.join(REFERENCE_ROOT, "configs", "prior.yaml")) as f:
        prior = yaml.safe_load(f)
    with open(os.path.join(REFERENCE_ROOT, "configs", "prob.yaml")) as f:
        prob = yaml.safe_load(f)
    cfg = {"prior_generator": prior, "prob_generator": prob}
    model = mods["flamed.models.flamed"].Flamed(cfg).eval()
    model.load_state_dict(flamed_sd, strict=True)
    fm = mods["flamed.models.facodec.facodec"]
    enc = dec = None
    if codec_enc_sd is not None:
        enc = fm.FACodecEncoder(ngf=32, up_ratios=[2, 4, 5, 5], out_channels=256).eval()
        enc.load_state_dict(codec_enc_sd, strict=True)
        enc = enc.to(device)
    if codec_dec_sd is not None:
        dec = fm.FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2],
                                vq_num_q_c=2, vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8,
                                codebook_size_prosody=10, codebook_size_content=10, codebook_size_residual=10,
                                use_gr_x_timbre=True, use_gr_residual_f0=True, use_gr_residual_phone=True).eval()
        res = dec.load_state_dict(codec_dec_sd, strict=False)  # training-only heads keep their own init
        assert not res.unexpected_keys, res.unexpected_keys
        dec = dec.to(device)
    return cfg, model.to(device), enc, dec
