"""Host-side pieces that need no GPU: the documented Philox map (oracle pinned on Random123's known answers), the
PCM_16 WAV writer (byte-identical to scipy's / soundfile's files), the lazy result dict."""
import io
import os

import numpy as np
import pytest
import torch


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 with 10 rounds: counter, key -> output"""
    from oracle import philox as P
    kats = [
        ([0, 0, 0, 0], (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, (0xffffffff, 0xffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0),
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        got = P.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert [int(x) for x in got] == want


def test_philox_normal_map_properties():
    from oracle import philox as P
    z = P.normal(2 ** 40 + 17, 2, 1 << 18)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert np.array_equal(P.normal(2 ** 40 + 17, 2, 1001), z[:1001])           # prefix property: index -> value
    assert not np.array_equal(P.normal(2 ** 40 + 17, 1, 1001), z[:1001])       # tensor id separates the draws
    assert not np.array_equal(P.normal(2 ** 40 + 18, 2, 1001), z[:1001])


def test_wav_writer_matches_scipy_bytes(tmp_path):
    from scipy.io import wavfile
    from flamed_tts_b200.wavio import WavWriter, pcm16_from_float, wav_header, write_wav_pcm16
    rng = np.random.default_rng(0)
    wav = np.tanh(rng.standard_normal(16000 * 2 + 37)).astype(np.float32)
    pcm = pcm16_from_float(wav)
    assert pcm.dtype == np.int16 and np.array_equal(pcm, np.rint(wav * np.float32(32767)).astype(np.int16))
    a, b = str(tmp_path / "a.wav"), str(tmp_path / "b.wav")
    write_wav_pcm16(a, pcm, 16000)
    wavfile.write(b, 16000, pcm)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert len(wav_header(10)) == 44
    paths = [str(tmp_path / ("w%03d.wav" % i)) for i in range(40)]
    with WavWriter(workers=4, sr=16000, max_pending=8) as w:
        for i, p in enumerate(paths):
            w.submit(p, pcm[: 1000 + i] if i % 2 else wav[: 1000 + i])   # int16 and float inputs
    for i, p in enumerate(paths):
        sr, data = wavfile.read(p)
        assert sr == 16000 and np.array_equal(data, pcm[: 1000 + i])
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".part")]
    with pytest.raises(OSError):
        with WavWriter(workers=1) as w:
            w.submit(str(tmp_path / "no_such_dir" / "x.wav"), pcm[:10])


def test_lazy_outputs_behave_like_the_reference_dict():
    from flamed.models.flamed import LazyOutputs
    calls = []
    out = LazyOutputs(prior_embs=1)
    out.lazy("prior_logits", lambda: calls.append(1) or "LOGITS")
    out.update(tgt_mask=2, latents=3, time=0.0)
    assert list(out.keys()) == ["prior_embs", "prior_logits", "tgt_mask", "latents", "time"]  # flamed.py:205-211 order
    assert "prior_logits" in out and not calls
    assert out["latents"] == 3 and not calls
    assert out["prior_logits"] == "LOGITS" and calls == [1]
    assert out["prior_logits"] == "LOGITS" and calls == [1]  # evaluated once
    out2 = LazyOutputs()
    out2.lazy("x", lambda: 5)
    assert dict(out2.items()) == {"x": 5} and out2.get("y", 7) == 7
