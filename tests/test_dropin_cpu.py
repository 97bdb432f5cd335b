"""CPU tests of the drop-in boundary: state-dict layout, PyTorch glue (FFT stacks, VQ, timbre
transformer) against the oracle / golden vectors, host logic and the C-ABI export table."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import flamed_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def model(cfg, flamed_sd):
    from flamed import Flamed
    m = Flamed(cfg).eval()
    m.load_state_dict(flamed_sd, strict=True)
    return m


def test_state_dict_layouts_match_reference(cfg, golden_dir):
    from flamed import Flamed
    from flamed.models.facodec import FACodecDecoder, FACodecEncoder
    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    assert {k: list(v.shape) for k, v in Flamed(cfg).state_dict().items()} == keys["flamed"]
    enc = FACodecEncoder(ngf=32, up_ratios=[2, 4, 5, 5], out_channels=256)
    assert {k: list(v.shape) for k, v in enc.state_dict().items()} == keys["codec_encoder"]
    dec = FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2], vq_num_q_c=2,
                         vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8, codebook_size_prosody=10,
                         codebook_size_content=10, codebook_size_residual=10, use_gr_x_timbre=True,
                         use_gr_residual_f0=True, use_gr_residual_phone=True)
    mine = {k: list(v.shape) for k, v in dec.state_dict().items()}
    assert all(keys["codec_decoder"][k] == v for k, v in mine.items())
    # the released checkpoint also carries training-only heads: accepted and ignored
    full = {k: torch.zeros(1) if k not in mine else torch.zeros(v) for k, v in keys["codec_decoder"].items()}
    res = dec.load_state_dict(full, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_symbols():
    from flamed.text import text_to_sequence
    from flamed.text.symbols import symbols
    assert len(symbols) == 360 and symbols[0] == "_" and symbols[-1] == "@sil"
    seq = text_to_sequence("{sp HH AH0 L OW1}", ["english_cleaners"])
    assert seq == [symbols.index("@sp"), symbols.index("@HH"), symbols.index("@AH0"), symbols.index("@L"),
                   symbols.index("@OW1")]


def test_phoneme_encoder_glue(model, flamed_sd, golden_dir):
    g = np.load(os.path.join(golden_dir, "sample_batch.npz"))
    phon, sl = torch.from_numpy(g["phonemes"]), torch.from_numpy(g["src_lens"])
    mask = O.get_mask_from_lengths(sl, phon.shape[1])
    with torch.inference_mode():
        enc = model.prior_generator.encoder(phon, mask)
    assert _rel(enc, torch.from_numpy(g["enc"])) < 2e-5


def test_prior_decoders_glue(model, cfg, flamed_sd, golden_dir):
    g = np.load(os.path.join(golden_dir, "sample_batch.npz"))
    enc = torch.from_numpy(g["enc"])
    x, tgt_len = O.length_regulator(enc, torch.from_numpy(g["phone_dur"]), torch.from_numpy(g["sil_dur"]),
                                    torch.from_numpy(g["src_lens"]))
    prompts = torch.from_numpy(g["prompts"])
    with torch.inference_mode():
        embs, logits, mask = model.prior_generator.decode_priors(x, tgt_len, prompts, prompts.shape[-1])
        o_embs, o_logits, o_mask = O.prior_after_pva(flamed_sd, "prior_generator", x, tgt_len, prompts,
                                                     cfg["prior_generator"])
    assert torch.equal(mask, o_mask)
    assert _rel(embs, o_embs) < 2e-5 and _rel(logits, o_logits) < 2e-5
    sub = lambda t: t.reshape(-1)[:: max(1, t.numel() // 4096)]
    assert _rel(sub(embs), torch.from_numpy(g["prior_embs_sub"])) < 2e-5


def test_prompt_side_has_no_cpu_fallback(codec_dec_sd, golden_dir):
    """FACodecDecoder.forward(vq=True) (quantisers + timbre transformer) runs in flm_codec_dec_prompt; on CPU it refuses"""
    from flamed.models.facodec import FACodecDecoder
    dec = FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2], vq_num_q_c=2,
                         vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8).eval()
    dec.load_state_dict(codec_dec_sd)
    g = np.load(os.path.join(golden_dir, "codec_encode.npz"))
    with pytest.raises(RuntimeError, match=r"no\s+CPU/PyTorch fallback"):
        dec(torch.from_numpy(g["enc_out"]), eval_vq=False, vq=True)
    assert len(dec(None, get_vq=True)) == 6


def test_sample_argument_errors(model):
    with pytest.raises(ValueError):
        model.sample(text="a", phonemes=torch.zeros(3, dtype=torch.long), prompt_raw=np.zeros(10), codec_encoder=1,
                     codec_decoder=1)
    with pytest.raises(ValueError):
        model.sample(phonemes=torch.zeros(3, dtype=torch.long), codec_encoder=1, codec_decoder=1)
    with pytest.raises(ValueError):
        model.sample(phonemes=torch.zeros(3, dtype=torch.long), prompt_processed=torch.zeros(6, 4), codec_encoder=1,
                     codec_decoder=1)
    with pytest.raises(ValueError):
        model.sample(phonemes=torch.zeros(3, dtype=torch.long), prompt_raw=np.zeros(10))


def test_no_cpu_fallback(model):
    """the hot path must fail loudly without a CUDA device - never silently run elsewhere"""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError, match="no CPU"):
        model.prob_generator.sample(torch.zeros(1, 6, 4, 384), torch.zeros(1, 256), torch.ones(1, 4, 1, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU"):
        model.prior_generator.pva.sample(torch.zeros(1, 3, 192), torch.tensor([3]), torch.zeros(1, 3, dtype=torch.bool))


def test_sample_batches_has_no_cpu_fallback(model):
    """the pipelined metadata entry point (Flamed.sample_batches) fails loudly on a CPU model, like sample_batch"""
    b = dict(phonemes=torch.ones(1, 4, dtype=torch.long), src_lens=torch.tensor([4]),
             prompts=torch.zeros(1, 6, 8, dtype=torch.long), timbres=torch.zeros(1, 256))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        model.sample_batches([b], nsteps_durgen=2, nsteps_denoiser=2)


def test_c_abi_exports_every_declared_symbol():
    """libflamed_b200.so loads and exports each function include/flamed_b200.h declares (no compute)."""
    from flamed_tts_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    header = open(os.path.join(ROOT, "include", "flamed_b200.h")).read()
    declared = set(re.findall(r"\b(flm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load_library().flm_version() == 100
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        assert lib.flm_ctx_create(0, ctypes.byref(h)) != 0  # fails loudly without a GPU
        lib.flm_last_error.restype = ctypes.c_char_p
        assert b"no CUDA device" in lib.flm_last_error() or b"CPU fallback" in lib.flm_last_error()
