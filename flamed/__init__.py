"""Drop-in `flamed` package: the reference's Python entry points (Flamed.from_pretrained /
sample / sample_batch, FACodecEncoder / FACodecDecoder) backed by the B200-native library
`flamed_tts_b200` (hand-written sm_100a kernels behind a C ABI)."""
from .models import Flamed  # noqa: F401
