from .fft import Decoder, Encoder, FFTBlock  # noqa: F401
