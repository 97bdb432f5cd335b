"""Pins the oracle port (oracle/flamed_oracle.py) against the committed outputs of the
UNMODIFIED reference (tests/golden/, produced by oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import torch

from oracle import flamed_oracle as O
from oracle import weights as W


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _sub(t, n=4096):
    f = t.reshape(-1)
    return f[:: max(1, f.numel() // n)]


def test_state_dict_layout(flamed_sd, codec_dec_sd, codec_enc_sd, golden_dir):
    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    assert len(keys["flamed"]) == 504 and len(keys["codec_encoder"]) == 206 and len(keys["codec_decoder"]) == 545
    assert {k: list(v.shape) for k, v in flamed_sd.items()} == keys["flamed"]
    assert {k: list(v.shape) for k, v in codec_enc_sd.items()} == keys["codec_encoder"]
    for k, v in codec_dec_sd.items():
        assert keys["codec_decoder"][k] == list(v.shape), k
    heads = ("f0_predictor", "phone_predictor", "res_f0_predictor", "res_phone_predictor", "x_timbre_predictor")
    assert all(k.split(".")[0] in heads for k in keys["codec_decoder"] if k not in codec_dec_sd)


def test_sample_batch_matches_reference(cfg, flamed_sd, codec_dec_sd, golden_dir):
    g = np.load(os.path.join(golden_dir, "sample_batch.npz"))
    T = lambda k: torch.from_numpy(g[k])
    nfe_dur, nfe_den = [int(v) for v in g["nfe"]]
    t_dur, t_den = [float(v) for v in g["temps"]]
    B, P = g["phonemes"].shape
    torch.manual_seed(int(g["noise_seed"]))
    n_dur, n_sil = torch.randn((B, P)), torch.randn((B, P))
    L = g["latents"].shape[-1]
    n_lat = torch.randn((B, L, 256))
    with torch.inference_mode():
        out = O.sample_batch(flamed_sd, cfg, T("phonemes"), T("src_lens"), T("prompts"), T("timbres"), n_dur, n_sil,
                             lambda b, l: n_lat, nfe_dur, nfe_den, t_dur, t_den, codec_sd=codec_dec_sd)
    # integer-valued results: bit exact
    assert torch.equal(out["phone_dur"], T("phone_dur")) and torch.equal(out["sil_dur"], T("sil_dur"))
    assert torch.equal(out["tgt_len"], T("tgt_len"))
    idx, _ = O.length_regulator_index(out["phone_dur"], out["sil_dur"], T("src_lens"))
    assert torch.equal(idx, T("lr_index"))
    # floating point: fp32 CPU vs fp32 CPU, tolerance 2e-5 relative L2
    for k in ("enc", "dur_t", "sil_t", "cond", "latents"):
        assert _rel(out[k], T(k)) < 2e-5, k
    assert _rel(_sub(out["prior_embs"]), T("prior_embs_sub")) < 2e-5
    assert _rel(_sub(out["prior_logits"]), T("prior_logits_sub")) < 2e-5
    assert list(out["wav"].shape) == list(g["wav_shape"])
    assert _rel(_sub(out["wav"], 16384), T("wav_sub")) < 2e-5


def test_length_regulator_edge_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "length_regulator.npz"))
    for i in range(int(g["n"])):
        ph, si, sl = (torch.from_numpy(g[f"{k}{i}"]) for k in ("phone", "sil", "src_lens"))
        idx, tl = O.length_regulator_index(ph, si, sl)
        assert torch.equal(idx, torch.from_numpy(g[f"index{i}"]))
        assert torch.equal(tl, torch.from_numpy(g[f"tgt_len{i}"]))


def test_activation1d(codec_dec_sd, golden_dir):
    g = np.load(os.path.join(golden_dir, "activation1d.npz"))
    y = O.activation1d(codec_dec_sd, "model.5", torch.from_numpy(g["x"]))
    assert _rel(y, torch.from_numpy(g["y"])) < 1e-6


def test_activation1d_closed_form(codec_dec_sd):
    """the polyphase closed form the CUDA kernel implements."""
    torch.manual_seed(3)
    x = torch.randn(1, 64, 23) * 2
    y = O.activation1d(codec_dec_sd, "model.5", x)
    f = codec_dec_sd["model.5.upsample.filter"].view(-1).double()
    a = torch.exp(codec_dec_sd["model.5.act.alpha"]).double()
    b = torch.exp(codec_dec_sd["model.5.act.beta"]).double()
    xd = x[0].double()
    T_ = xd.shape[1]
    cl = lambda i: min(max(i, 0), T_ - 1)
    u = torch.zeros(64, 2 * T_, dtype=torch.float64)
    for n in range(T_):
        for j in range(6):
            u[:, 2 * n] += 2 * xd[:, cl(n - 3 + j)] * f[11 - 2 * j]
            u[:, 2 * n + 1] += 2 * xd[:, cl(n - 2 + j)] * f[10 - 2 * j]
    s = u + torch.sin(u * a[:, None]) ** 2 / (b[:, None] + 1e-9)
    yy = torch.zeros(64, T_, dtype=torch.float64)
    for n in range(T_):
        for k in range(12):
            yy[:, n] += s[:, min(max(2 * n + k - 5, 0), 2 * T_ - 1)] * f[k]
    assert _rel(yy.float(), y[0]) < 1e-6


def test_codec_encode_and_prompt_features(codec_enc_sd, codec_dec_sd, golden_dir):
    g = np.load(os.path.join(golden_dir, "codec_encode.npz"))
    with torch.inference_mode():
        e = O.codec_encode(codec_enc_sd, torch.from_numpy(g["wav"]))
        assert _rel(e, torch.from_numpy(g["enc_out"])) < 2e-5
        codes, spk = O.codec_prompt_features(codec_dec_sd, torch.from_numpy(g["enc_out"]))
    assert torch.equal(codes, torch.from_numpy(g["codes"]))
    assert _rel(spk, torch.from_numpy(g["timbre"])) < 2e-5


def test_codec_decode(codec_dec_sd, golden_dir):
    g = np.load(os.path.join(golden_dir, "codec_decode.npz"))
    with torch.inference_mode():
        w = O.codec_decode(codec_dec_sd, torch.from_numpy(g["latents"]), torch.from_numpy(g["spk"]))
    assert _rel(w, torch.from_numpy(g["wav"])) < 2e-5
