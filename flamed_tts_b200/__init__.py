"""flamed_tts_b200 - B200-native (sm_100a) implementation of the Flamed-TTS inference hot path.

csrc/      hand-written CUDA kernels + the C ABI (include/flamed_b200.h) -> libflamed_b200.so
_lib.py    ctypes binding
engines.py torch-facing handles (device memory / streams only)
The reference-compatible Python API lives in the sibling `flamed` package.
"""
from ._lib import FLM_BF16, FLM_F32, LIB_PATH, load_library  # noqa: F401
