"""ctypes binding of libflamed_b200.so (C ABI: include/flamed_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a
call fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libflamed_b200.so")

FLM_F32 = 0
FLM_BF16 = 1


class flm_tensor(ctypes.Structure):
    _fields_ = [("name", c_char_p), ("data", c_void_p), ("ndim", c_int32), ("shape", c_int64 * 4)]


class flm_prob_cfg(ctypes.Structure):
    _fields_ = [(k, c_int32) for k in ("target_dim", "spk_dim", "cond_dim", "hidden_dim", "n_layers", "n_quantizers",
                                        "kernel_size", "downsampling_stages")]


# every symbol declared in include/flamed_b200.h: name -> (restype, argtypes)
SIGNATURES = {
    "flm_last_error": (c_char_p, []),
    "flm_version": (c_int, []),
    "flm_launch_count": (ctypes.c_ulonglong, []),
    "flm_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "flm_ctx_destroy": (None, [c_void_p]),
    "flm_durgen_load": (c_int, [c_void_p, POINTER(flm_tensor), c_int, POINTER(c_void_p)]),
    "flm_durgen_destroy": (None, [c_void_p]),
    "flm_durgen_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_int, c_float,
                                  c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "flm_durgen_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p,
                                   c_void_p]),
    "flm_philox_normal": (c_int, [c_void_p, c_uint64, c_int, c_int64, c_void_p, c_void_p]),
    "flm_lr_expand_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p]),
    "flm_lr_plan": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                            POINTER(c_int64), c_void_p]),
    "flm_lr_expand": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "flm_denoiser_load": (c_int, [c_void_p, POINTER(flm_tensor), c_int, POINTER(flm_prob_cfg), c_int,
                                  POINTER(c_void_p)]),
    "flm_denoiser_destroy": (None, [c_void_p]),
    "flm_cond_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "flm_denoiser_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_int, c_int, c_int,
                                    c_float, c_void_p, c_int, c_void_p]),
    "flm_denoiser_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p]),
    "flm_denoiser_launches_per_step": (c_int, [c_void_p]),
    "flm_codec_dec_load": (c_int, [c_void_p, POINTER(flm_tensor), c_int, c_int, POINTER(c_void_p)]),
    "flm_codec_dec_destroy": (None, [c_void_p]),
    "flm_codec_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "flm_codec_dec_activation": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "flm_codec_dec_prompt": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "flm_codec_enc_load": (c_int, [c_void_p, POINTER(flm_tensor), c_int, POINTER(c_void_p)]),
    "flm_codec_enc_destroy": (None, [c_void_p]),
    "flm_codec_enc_frames": (c_int64, [c_void_p, c_int64]),
    "flm_codec_encode": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "flm_wav_to_pcm16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "flm_comm_unique_id": (c_int, [c_void_p]),
    "flm_comm_create": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "flm_comm_destroy": (None, [c_void_p]),
    "flm_gather_wav": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, POINTER(c_int64), c_int, c_void_p]),
    "flm_profile_enable": (c_int, [c_void_p, c_int]),
    "flm_profile_read": (c_int, [c_void_p, POINTER(ctypes.c_double), c_int]),
    "flm_profile_class_name": (c_char_p, [c_int]),
    "flm_profile_detail": (c_int, [c_void_p, c_char_p, c_int]),
    "flm_tapgemm_test": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "flm_conv1d_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_void_p, c_void_p, c_void_p]),
    "flm_layernorm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int64, c_int, c_void_p, c_void_p,
                                   c_void_p]),
    "flm_attention_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "flm_tapgemm_test_bf16": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "flm_tapgemm_bench": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  POINTER(c_float), c_void_p]),
}

_lib = None


def load_library():
    """dlopen libflamed_b200.so and bind every declared symbol (no compute is issued)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libflamed_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`python flamed_tts_b200/build.py`. There is no CPU/PyTorch fallback for the hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load_library().flm_last_error()
        raise RuntimeError("flamed_b200 error %d: %s" % (code, msg.decode() if msg else "?"))


def pack_weights(named_tensors):
    """[(name, torch.Tensor)] -> (flm_tensor array, n, keep-alive list).  Tensors are converted to
    contiguous fp32 on the host; the library copies them during *_load."""
    import torch
    keep, arr = [], (flm_tensor * len(named_tensors))()
    for i, (name, t) in enumerate(named_tensors):
        t = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
        if t.ndim > 4:
            raise ValueError("tensor %s has rank %d > 4" % (name, t.ndim))
        bname = name.encode()
        keep.append((t, bname))
        arr[i].name = bname
        arr[i].data = t.data_ptr()
        arr[i].ndim = t.ndim
        for d in range(t.ndim):
            arr[i].shape[d] = t.shape[d]
    return arr, len(named_tensors), keep
