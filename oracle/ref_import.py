"""Import the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python/PyTorch but imports a
dozen packages that are absent here (librosa, lightning, g2p_en, ...); none of
them is touched by the inference hot path, so they are replaced by empty stub
modules before `import flamed` (recipe: SURVEY.md Appendix B).  This module is
used only by `oracle/make_golden.py` and by tests that pin the oracle port
against the live reference; `/root/reference` does not exist on the GPU box, so
nothing under `-m gpu`, `smoke()` or `bench.py` may import this file.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FLAMED_REFERENCE_ROOT", "/root/reference")


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
