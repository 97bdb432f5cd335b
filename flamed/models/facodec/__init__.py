from .facodec import FACodecDecoder, FACodecEncoder  # noqa: F401
