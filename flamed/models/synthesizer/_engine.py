"""Lazy ownership of B200 engine handles by the drop-in nn.Modules."""
import os

import torch
import torch.nn as nn


def default_precision():
    return os.environ.get("FLAMED_B200_PRECISION", "bf16")


class EngineOwner(nn.Module):
    """An nn.Module whose inference runs in a flamed_tts_b200 engine built from its own
    state_dict.  The engine is rebuilt when the weights may have changed (load_state_dict,
    .to(), precision switch)."""

    def __init__(self):
        super().__init__()
        self._engine = None
        self._engine_key = None
        self._weights_version = 0
        self.precision = default_precision()

    def _apply(self, fn, *a, **k):
        self._weights_version += 1
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._weights_version += 1
        return super().load_state_dict(*a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._weights_version += 1
        return super()._load_from_state_dict(*a, **k)

    def set_precision(self, precision):
        self.precision = precision
        for m in self.children():
            if isinstance(m, EngineOwner):
                m.set_precision(precision)
        return self

    def _device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    def _build_engine(self, ctx):
        raise NotImplementedError

    def engine(self):
        from flamed_tts_b200.engines import Context
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError(
                "%s: the Flamed-TTS hot path runs on a B200 (sm_100a) through libflamed_b200.so; there is no "
                "CPU/PyTorch fallback. Move the model to a CUDA device." % type(self).__name__)
        key = (dev.index if dev.index is not None else torch.cuda.current_device(), self.precision,
               self._weights_version)
        if self._engine is None or self._engine_key != key:
            self._engine = self._build_engine(Context.get(dev))
            self._engine_key = key
        return self._engine
