"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and
the committed reference outputs (tests/golden/).  Tolerances (stated per test):
  integer-valued results (durations, tgt_len, LR indices): bit exact
  fp32 mode  (FLM_F32, fp32 FMA):        relative L2 <= 1e-5 per kernel, <= 5e-5 end to end
  bf16 mode  (FLM_BF16, tcgen05 bf16):   relative L2 <= 1e-2 latents, <= 3e-2 waveform
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import flamed_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def ctx():
    from flamed_tts_b200.engines import Context
    return Context.get(DEV)


def _ref_tapgemm(A, W, bias, T_out, off0, dil, stride, epi):
    """torch CPU reference of the tap-GEMM definition (fp64)"""
    B, T_in, K = A.shape
    ntaps, N, _ = W.shape
    out = torch.zeros(B, T_out, N, dtype=torch.float64)
    Ad, Wd = A.double(), W.double()
    for tap in range(ntaps):
        for t in range(T_out):
            ti = t * stride + off0 + tap * dil
            if 0 <= ti < T_in:
                out[:, t] += Ad[:, ti] @ Wd[tap].t()
    if bias is not None:
        out += bias.double()
    if epi == 1:
        out = F.gelu(out)
    elif epi == 2:
        out = F.silu(out)
    elif epi == 3:
        out = F.relu(out)
    return out


CASES = [  # B, T_in, K, N, ntaps, off0, dil, stride, epi
    (2, 37, 64, 64, 1, 0, 1, 1, 0),
    (1, 130, 192, 384, 3, -1, 1, 1, 3),
    (3, 50, 128, 128, 7, -9, 3, 1, 1),
    (2, 41, 32, 64, 4, -1, 1, 2, 2),
    (1, 200, 256, 320, 3, -1, 1, 1, 0),
]


@pytest.mark.parametrize("case", CASES)
def test_tapgemm_fp32(ctx, case):
    from flamed_tts_b200.engines import tapgemm
    B, T_in, K, N, ntaps, off0, dil, stride, epi = case
    g = torch.Generator().manual_seed(1)
    A = torch.randn(B, T_in, K, generator=g)
    W = torch.randn(ntaps, N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    T_out = T_in if stride == 1 else (T_in + 2 * 1 - ntaps) // stride + 1
    out = tapgemm(ctx, "fp32", A, W, bias, T_out, ntaps, off0, dil, stride, epi)
    ref = _ref_tapgemm(A, W, bias, T_out, off0, dil, stride, epi)
    assert _rel(out, ref) < 2e-6


_TC_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from flamed_tts_b200.engines import Context, tapgemm
ctx = Context.get("cuda:0")
worst = 0.0
for (B, T, K, N, ntaps, off0, dil, epi) in [(1, 128, 64, 64, 1, 0, 1, 0), (2, 300, 256, 256, 1, 0, 1, 1),
                                             (3, 333, 1024, 1024, 1, 0, 1, 2), (2, 150, 128, 128, 7, -9, 3, 0),
                                             (2, 77, 512, 2560, 3, -1, 1, 0), (1, 1000, 64, 64, 7, -3, 1, 3),
                                             (4, 260, 1024, 256, 3, -1, 1, 0), (1, 5000, 1024, 1024, 1, 0, 1, 0)]:
    g = torch.Generator().manual_seed(2)
    A = torch.randn(B, T, K, generator=g).bfloat16().float()
    W = (torch.randn(ntaps, N, K, generator=g) / K ** 0.5).bfloat16().float()
    bias = torch.randn(N, generator=g)
    o_tc = tapgemm(ctx, "bf16", A, W, bias, T, ntaps, off0, dil, 1, epi)
    o_32 = tapgemm(ctx, "fp32", A, W, bias, T, ntaps, off0, dil, 1, epi)
    torch.cuda.synchronize()
    rel = float((o_tc.double() - o_32.double()).norm() / o_32.double().norm())
    print("tc-vs-fp32", (B, T, K, N, ntaps, off0, dil, epi), "rel %%.3e" %% rel, flush=True)
    worst = max(worst, rel)
assert worst < 1e-5, worst
print("OK")
"""


def test_tapgemm_tcgen05_matches_fp32_kernel():
    """tcgen05/TMA kernel vs the fp32 FMA kernel on identical bf16-representable operands (products are
    exact in fp32, so only the accumulation order differs: <= 1e-5).  Runs in a subprocess with a timeout
    so that a pipeline dead-lock cannot hang the session."""
    r = subprocess.run([sys.executable, "-c", _TC_SCRIPT % ROOT], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "OK" in r.stdout


def test_tapgemm_cta_pair_kernel_vs_torch_and_first_generation():
    """tapgemm_tc2.cu (cta_group::2 + TMA slab epilogue): every epilogue (none/GELU/SiLU/ReLU, codec skip, adaLN-gated
    residual with and without the ConvNeXt addend), flattened and per-sample tilings, ragged tails, odd tile counts,
    dilated taps - against a torch fp32 restatement on the same bf16 operands (rel-L2 <= 4e-3 = bf16 output
    rounding) and bit-for-bit against the single-CTA kernel.  Subprocess + timeout: a dead-lock must not hang."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tc2_check.py")], capture_output=True, text=True,
                       timeout=240)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0 and "ALL OK" in r.stdout
    assert "gen1 vs gen2 max-abs 0.000e+00" in r.stdout


def test_tiny_and_ragged_shapes_bf16():
    """L = 2 ... 129 frames, B = 1 ... 5 through the bf16 denoiser velocity and the codec decoder (tiles smaller than a
    CTA pair, chunks shorter than the conv window, dummy half-tiles): <= 2e-2 / 5e-2 of the oracle, no NaNs"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "edge_shapes.py")], capture_output=True, text=True,
                       timeout=240)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "ALL OK" in r.stdout


def test_depthwise_conv_one_kernel_per_dtype(ctx, flamed_sd, cfg):
    """one depthwise-conv kernel per arithmetic mode: fp32 = dwconv_kernel<float> (statistics finalised in the kernel),
    bf16 = the LayerNorm-fused persistent kernel (dwconv_fused.cu) fed by row statistics from the GEMM epilogue.
    Ragged batch (333 = 10 full 32-frame chunks + 13 frames, last 128-row tile partially empty): velocity within the
    stated tolerances of the oracle, and bit-reproducible."""
    from flamed_tts_b200.engines import DenoiserEngine
    prob = {k[len("prob_generator."):]: v for k, v in flamed_sd.items() if k.startswith("prob_generator.")}
    g = torch.Generator().manual_seed(5)
    B, L = 3, 333
    x, spk = torch.randn(B, L, 256, generator=g), torch.randn(B, 256, generator=g)
    with torch.inference_mode():
        ref = O.denoiser_forward(flamed_sd, "prob_generator.denoiser", x, torch.full((1, 1), 0.37), spk)
    for mode, tol in (("fp32", 1e-5), ("bf16", 1e-2)):
        den = DenoiserEngine(ctx, prob, cfg["prob_generator"], mode)
        v = den.forward(x, 0.37, spk)
        e = _rel(v, ref)
        print("%s velocity rel-L2 %.3e" % (mode, e))
        assert e < tol
        assert torch.equal(v, den.forward(x, 0.37, spk))


# ------------------------------------------------------------------------------------------------ modules
@pytest.fixture(scope="module")
def engines(ctx, cfg, flamed_sd, codec_dec_sd, codec_enc_sd):
    from flamed_tts_b200.engines import CodecDecoderEngine, CodecEncoderEngine, DenoiserEngine, DurationEngine
    pva = {k[len("prior_generator.pva."):]: v for k, v in flamed_sd.items() if k.startswith("prior_generator.pva.")}
    prob = {k[len("prob_generator."):]: v for k, v in flamed_sd.items() if k.startswith("prob_generator.")}
    return dict(
        dur=DurationEngine(ctx, pva),
        den32=DenoiserEngine(ctx, prob, cfg["prob_generator"], "fp32"),
        den16=DenoiserEngine(ctx, prob, cfg["prob_generator"], "bf16"),
        dec32=CodecDecoderEngine(ctx, codec_dec_sd, "fp32"),
        dec16=CodecDecoderEngine(ctx, codec_dec_sd, "bf16"),
        enc=CodecEncoderEngine(ctx, codec_enc_sd),
    )


@pytest.fixture(scope="module")
def gold(golden_dir):
    return {k: np.load(os.path.join(golden_dir, k + ".npz")) for k in
            ("sample_batch", "length_regulator", "activation1d", "codec_encode", "codec_decode")}


def _noise(g):
    B, P = g["phonemes"].shape
    torch.manual_seed(int(g["noise_seed"]))
    n_dur, n_sil = torch.randn((B, P)), torch.randn((B, P))
    n_lat = torch.randn((B, g["latents"].shape[-1], 256))
    return n_dur, n_sil, n_lat


def test_durgen_and_rounding(engines, gold):
    """a1-a3: ODE state within 1e-5, rounded durations bit exact (ties within 2e-5*(1+v) of k+.5 reported)"""
    g = gold["sample_batch"]
    enc, sl = torch.from_numpy(g["enc"]), torch.from_numpy(g["src_lens"])
    mask = O.get_mask_from_lengths(sl, enc.shape[1])
    n_dur, n_sil, _ = _noise(g)
    nfe = int(g["nfe"][0])
    ts = torch.linspace(0, 1, nfe + 1)
    phone, sil, dur_t, sil_t = engines["dur"].sample(enc, mask, n_dur, n_sil, ts, float(g["temps"][0]))
    assert _rel(dur_t, torch.from_numpy(g["dur_t"])) < 1e-5 and _rel(sil_t, torch.from_numpy(g["sil_t"])) < 1e-5
    for got, ref, x in ((phone, g["phone_dur"], g["dur_t"]), (sil, g["sil_dur"], g["sil_t"])):
        got = got.cpu().numpy()
        diff = got != ref
        if diff.any():
            v = np.exp(x.astype(np.float64)) - 1
            tie = np.abs(v - np.floor(v) - 0.5) < 2e-5 * (1 + np.abs(v))
            assert (diff & ~tie).sum() == 0, "durations differ away from rounding ties"
            print("rounding ties:", int(diff.sum()))


def test_durgen_larger_batch_vs_oracle(engines, flamed_sd):
    torch.manual_seed(5)
    B, P, nfe = 5, 70, 16
    enc = torch.randn(B, P, 192)
    sl = torch.tensor([70, 33, 1, 64, 70])
    mask = O.get_mask_from_lengths(sl, P)
    n_dur, n_sil = torch.randn(B, P), torch.randn(B, P)
    ts = torch.linspace(0, 1, nfe + 1)
    phone, sil, dur_t, sil_t = engines["dur"].sample(enc, mask, n_dur, n_sil, ts, 0.3)
    with torch.inference_mode():
        o_phone, o_sil, o_dur_t, o_sil_t = O.durgen_sample(flamed_sd, "prior_generator.pva", enc, mask, n_dur, n_sil, nfe, 0.3)
    assert _rel(dur_t, o_dur_t) < 1e-5 and _rel(sil_t, o_sil_t) < 1e-5
    v = torch.exp(o_dur_t.double()) - 1
    tie = (v - v.floor() - 0.5).abs() < 2e-5 * (1 + v.abs())
    assert int(((phone.cpu() != o_phone) & ~tie).sum()) == 0
    assert int((sil.cpu() != o_sil).sum()) <= int(((torch.exp(o_sil_t.double()) - 1 - 0.5).abs() < 1e-4).sum())


def test_length_regulator_bit_exact(engines, gold):
    """a4: indices, tgt_len and gathered rows identical to the reference (incl. padded phonemes -> 1 frame,
    zero durations, silence frames copying row 0)"""
    g = gold["length_regulator"]
    for i in range(int(g["n"])):
        x, ph, si, sl = (torch.from_numpy(g[f"{k}{i}"]) for k in ("x", "phone", "sil", "src_lens"))
        out, tgt_len, idx = engines["dur"].length_regulate(x, ph, si, sl, return_index=True)
        ref_idx = torch.from_numpy(g[f"index{i}"])
        assert torch.equal(tgt_len.cpu(), torch.from_numpy(g[f"tgt_len{i}"]))
        assert torch.equal(idx.cpu().long(), ref_idx)
        ref, _ = O.length_regulator(x, ph, si, sl)
        assert torch.equal(out.cpu(), ref)
    g = gold["sample_batch"]
    out, tgt_len, idx = engines["dur"].length_regulate(torch.from_numpy(g["enc"]), torch.from_numpy(g["phone_dur"]),
                                                       torch.from_numpy(g["sil_dur"]), torch.from_numpy(g["src_lens"]),
                                                       return_index=True)
    assert torch.equal(idx.cpu().long(), torch.from_numpy(g["lr_index"]))
    assert torch.equal(tgt_len.cpu(), torch.from_numpy(g["tgt_len"]))


def test_length_regulator_full_size_properties(engines):
    """BASELINE config sizes: B=64, P=180: tgt_len = sum of clamped repeats, index non-decreasing over phoneme
    segments, every phoneme i appears exactly phone_rep[i] times"""
    g = torch.Generator().manual_seed(9)
    B, P = 64, 180
    x = torch.randn(B, P, 192, generator=g)
    ph = torch.randint(0, 20, (B, P), generator=g).float()
    si = torch.zeros(B, P)
    sl = torch.randint(100, P + 1, (B,), generator=g)
    out, tgt_len, idx = engines["dur"].length_regulate(x, ph, si, sl, return_index=True)
    valid = torch.arange(P)[None] < sl[:, None]
    rep = torch.where(valid, ph, torch.zeros_like(ph)).long().clamp(min=1)
    assert torch.equal(tgt_len.cpu(), rep.sum(1))
    idx = idx.cpu().long()
    for b in range(0, B, 7):
        row = idx[b, : int(tgt_len[b])]
        assert bool((row[1:] >= row[:-1]).all())
        assert torch.equal(torch.bincount(row, minlength=P), rep[b])
        assert bool((idx[b, int(tgt_len[b]):] == -1).all())
        assert torch.equal(out[b, : int(tgt_len[b])].cpu(), x[b][row])


def test_cond_prepare(engines, flamed_sd):
    """a5 (first half): fold + down-sampler; fp32 <= 1e-5, bf16 <= 1e-2"""
    torch.manual_seed(4)
    B, L = 2, 70
    prior = torch.randn(B, 6, L, 384)
    mask = torch.ones(B, L, 1, dtype=torch.bool)
    mask[1, 50:] = False
    with torch.inference_mode():
        ref = O.cond_prepare(flamed_sd, "prob_generator", prior, mask)
    assert _rel(engines["den32"].cond_prepare(prior, mask), ref) < 1e-5
    assert _rel(engines["den16"].cond_prepare(prior, mask), ref) < 1e-2


def test_denoiser_velocity(engines, flamed_sd):
    """a6: one SimpleMLPAdaLN forward; fp32 <= 1e-5, bf16 <= 1e-2"""
    torch.manual_seed(6)
    B, L = 2, 77
    x, spk = torch.randn(B, L, 256), torch.randn(B, 256)
    t = 0.375
    with torch.inference_mode():
        ref = O.denoiser_forward(flamed_sd, "prob_generator.denoiser", x, torch.tensor([[t]]), spk)
    e32 = _rel(engines["den32"].forward(x, t, spk), ref)
    e16 = _rel(engines["den16"].forward(x, t, spk), ref)
    print("velocity rel-L2 fp32 %.3e bf16 %.3e" % (e32, e16))
    assert e32 < 1e-5 and e16 < 1e-2


def test_denoiser_sample_matches_reference(engines, gold):
    """a5/a6: nfe-step Euler loop (CUDA graph) on the reference's cond / noise; latents vs the reference"""
    g = gold["sample_batch"]
    _, _, n_lat = _noise(g)
    cond, spk = torch.from_numpy(g["cond"]), torch.from_numpy(g["timbres"])
    nfe = int(g["nfe"][1])
    ts = torch.linspace(0, 1, nfe + 1)
    ref = torch.from_numpy(g["latents"]).transpose(1, 2)
    for name, tol in (("den32", 1e-5), ("den16", 1e-2)):
        lat = engines[name].sample(cond, spk, n_lat, ts, float(g["temps"][1]), use_graph=True)
        lat2 = engines[name].sample(cond, spk, n_lat, ts, float(g["temps"][1]), use_graph=False)
        lat3 = engines[name].sample(cond, spk, n_lat, ts, float(g["temps"][1]), use_graph=True)  # replay
        e = _rel(lat, ref)
        print(name, "latents rel-L2 %.3e max-abs %.3e" % (e, float((lat.cpu() - ref).abs().max())))
        assert e < tol
        assert torch.equal(lat, lat2) and torch.equal(lat, lat3), "graph replay must be bit-identical to eager launch"


def test_activation1d(engines, gold):
    """a8: anti-aliased SnakeBeta, fp32, <= 1e-6 (edges: replicate padding on both signals)"""
    g = gold["activation1d"]
    x = torch.from_numpy(g["x"])
    y = engines["dec32"].activation("model.5", x.transpose(1, 2).contiguous())
    assert _rel(y.transpose(1, 2), torch.from_numpy(g["y"])) < 1e-6


def test_activation1d_short_sequences(engines, codec_dec_sd):
    torch.manual_seed(8)
    for T in (1, 2, 3, 5, 11, 16, 17, 33):
        x = torch.randn(2, 64, T) * 2
        y = engines["dec32"].activation("model.5", x.transpose(1, 2).contiguous())
        assert _rel(y.transpose(1, 2), O.activation1d(codec_dec_sd, "model.5", x)) < 1e-6, T


def test_codec_decode(engines, gold, codec_dec_sd):
    """a7: FACodecDecoder.inference; fp32 <= 2e-5 vs the reference, bf16 <= 3e-2"""
    g = gold["codec_decode"]
    lat, spk = torch.from_numpy(g["latents"]), torch.from_numpy(g["spk"])
    ref = torch.from_numpy(g["wav"])
    w32 = engines["dec32"].decode(lat.transpose(1, 2).contiguous(), spk)
    w16 = engines["dec16"].decode(lat.transpose(1, 2).contiguous(), spk)
    e32, e16 = _rel(w32, ref), _rel(w16, ref)
    print("codec decode rel-L2 fp32 %.3e bf16 %.3e" % (e32, e16))
    assert w32.shape == ref.shape and e32 < 2e-5 and e16 < 3e-2
    # batch of ragged content: a longer, batched case against the oracle
    torch.manual_seed(10)
    lat, spk = torch.randn(3, 256, 45), torch.randn(3, 256)
    with torch.inference_mode():
        ref = O.codec_decode(codec_dec_sd, lat, spk)
    assert _rel(engines["dec32"].decode(lat.transpose(1, 2).contiguous(), spk), ref) < 2e-5


def test_codec_encode(engines, gold):
    """a9: FACodecEncoder.forward, fp32 <= 2e-5"""
    g = gold["codec_encode"]
    out = engines["enc"].encode(torch.from_numpy(g["wav"]))
    ref = torch.from_numpy(g["enc_out"])
    assert out.shape == ref.shape and _rel(out, ref) < 2e-5


def test_codec_encode_ragged_length(engines, codec_enc_sd):
    torch.manual_seed(12)
    wav = torch.randn(2, 1, 3333) * 0.1
    with torch.inference_mode():
        ref = O.codec_encode(codec_enc_sd, wav)
    out = engines["enc"].encode(wav)
    assert out.shape == ref.shape and _rel(out, ref) < 2e-5


# ------------------------------------------------------------------------------------------------ drop-in API
@pytest.fixture(scope="module")
def dropin(cfg, flamed_sd, codec_dec_sd, codec_enc_sd):
    from flamed import Flamed
    from flamed.models.facodec import FACodecDecoder, FACodecEncoder
    model = Flamed(cfg).eval()
    model.load_state_dict(flamed_sd)
    model.to(DEV)
    dec = FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2], vq_num_q_c=2,
                         vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8).eval()
    dec.load_state_dict(codec_dec_sd)
    dec.to(DEV)
    enc = FACodecEncoder(ngf=32, up_ratios=[2, 4, 5, 5], out_channels=256).eval()
    enc.load_state_dict(codec_enc_sd)
    enc.to(DEV)
    return model, enc, dec


def _sub(t, n=4096):
    f = t.reshape(-1)
    return f[:: max(1, f.numel() // n)]


def test_sample_batch_fp32_matches_reference(dropin, gold):
    """Flamed.sample_batch on the reference's inputs and seed: integers exact, latents <= 5e-5, wav <= 2e-4"""
    model, _, dec = dropin
    model.set_precision("fp32").set_noise_device("cpu")
    dec.set_precision("fp32")
    g = gold["sample_batch"]
    torch.manual_seed(int(g["noise_seed"]))
    out = model.sample_batch(torch.from_numpy(g["phonemes"]), torch.from_numpy(g["src_lens"]),
                             torch.from_numpy(g["prompts"]), torch.from_numpy(g["timbres"]), codec_decoder=dec,
                             temp_durgen=float(g["temps"][0]), temp_denoiser=float(g["temps"][1]),
                             nsteps_durgen=int(g["nfe"][0]), nsteps_denoiser=int(g["nfe"][1]))
    assert list(out["latents"].shape) == list(g["latents"].shape)
    assert torch.equal((~out["tgt_mask"]).sum(1).cpu(), torch.from_numpy(g["tgt_len"]))
    e_emb = _rel(_sub(out["prior_embs"]), torch.from_numpy(g["prior_embs_sub"]))
    e_lat = _rel(out["latents"], torch.from_numpy(g["latents"]))
    e_wav = _rel(_sub(out["wav"], 16384), torch.from_numpy(g["wav_sub"]))
    print("fp32 end-to-end rel-L2: prior_embs %.3e latents %.3e wav %.3e" % (e_emb, e_lat, e_wav))
    assert list(out["wav"].shape) == list(g["wav_shape"])
    assert e_emb < 2e-5 and e_lat < 5e-5 and e_wav < 2e-4


def test_sample_batch_bf16_tolerance(dropin, gold):
    """bf16 mode (tcgen05): same inputs; durations still exact (fp32 path), latents <= 1e-2, wav <= 3e-2"""
    model, _, dec = dropin
    model.set_precision("bf16").set_noise_device("cpu")
    dec.set_precision("bf16")
    g = gold["sample_batch"]
    torch.manual_seed(int(g["noise_seed"]))
    out = model.sample_batch(torch.from_numpy(g["phonemes"]), torch.from_numpy(g["src_lens"]),
                             torch.from_numpy(g["prompts"]), torch.from_numpy(g["timbres"]), codec_decoder=dec,
                             temp_durgen=float(g["temps"][0]), temp_denoiser=float(g["temps"][1]),
                             nsteps_durgen=int(g["nfe"][0]), nsteps_denoiser=int(g["nfe"][1]))
    assert torch.equal((~out["tgt_mask"]).sum(1).cpu(), torch.from_numpy(g["tgt_len"]))
    e_lat = _rel(out["latents"], torch.from_numpy(g["latents"]))
    e_wav = _rel(_sub(out["wav"], 16384), torch.from_numpy(g["wav_sub"]))
    print("bf16 end-to-end rel-L2: latents %.3e wav %.3e" % (e_lat, e_wav))
    assert e_lat < 1e-2 and e_wav < 3e-2
    model.set_precision("fp32")
    dec.set_precision("fp32")


def test_sample_batches_pipelined_equals_sequential(dropin):
    """Flamed.sample_batches (front stage of bucket i+1 on a side stream) returns exactly what a loop of
    sample_batch calls returns: same draws (device generator, same order), same kernels, bit-identical tensors."""
    model, enc, dec = dropin
    model.set_precision("bf16").set_noise_device("cuda")
    dec.set_precision("bf16")
    g = torch.Generator().manual_seed(11)
    batches = []
    for B, P in ((3, 17), (2, 40), (4, 9)):
        lens = torch.randint(max(2, P - 6), P + 1, (B,), generator=g)
        lens[0] = P
        ph = torch.randint(1, 70, (B, P), generator=g)
        for i in range(B):
            ph[i, lens[i]:] = 0
        batches.append(dict(phonemes=ph, src_lens=lens, prompts=torch.randint(0, 1024, (B, 6, 30), generator=g),
                            timbres=torch.randn(B, 256, generator=g)))
    kw = dict(codec_decoder=dec, temp_durgen=0.3, temp_denoiser=0.3, nsteps_durgen=4, nsteps_denoiser=4)
    try:
        torch.manual_seed(3)
        seq = [model.sample_batch(b["phonemes"], b["src_lens"], b["prompts"], b["timbres"], **kw) for b in batches]
        torch.cuda.synchronize()
        torch.manual_seed(3)
        pipe = model.sample_batches(batches, **kw)
        torch.cuda.synchronize()
        assert len(pipe) == len(seq)
        for a, b in zip(seq, pipe):
            assert torch.equal(a["tgt_mask"], b["tgt_mask"])
            assert torch.equal(a["latents"], b["latents"]) and torch.equal(a["wav"], b["wav"])
    finally:
        model.set_precision("fp32").set_noise_device("cpu")
        dec.set_precision("fp32")


def test_sample_single_utterance_with_raw_prompt(dropin):
    """Flamed.sample(phonemes=..., prompt_raw=...) end to end incl. the prompt encoder; shapes + finiteness,
    and idempotence for a fixed seed"""
    model, enc, dec = dropin
    model.set_precision("fp32")
    g = torch.Generator().manual_seed(21)
    phon = torch.randint(1, 300, (20,), generator=g)
    prompt = (torch.randn(1, 1, 8000, generator=g) * 0.1)
    outs = []
    for _ in range(2):
        torch.manual_seed(99)
        r = model.sample(phonemes=phon, prompt_raw=prompt, codec_encoder=enc, codec_decoder=dec, nsteps_durgen=4,
                         nsteps_denoiser=4)
        assert r["wav"].dtype == np.float32 and r["wav"].ndim == 1 and r["wav"].size % 200 == 0
        assert np.isfinite(r["wav"]).all() and np.abs(r["wav"]).max() <= 1.0
        outs.append(r["wav"])
    assert np.array_equal(outs[0], outs[1])


def test_denoiser_ragged_last_tile_many_steps(engines, flamed_sd):
    """regression: B*L with a partially empty last 128-row tile (rows % 128 in (0, 96]) at the LAST Euler step used
    to read the adaLN gate of a non-existent sample (out-of-bounds).  Also checks the loop against the oracle."""
    torch.manual_seed(17)
    B, L, nfe = 3, 53, 6  # 159 rows: last tile has 31 valid rows -> three empty quarters
    cond, spk, noise = torch.relu(torch.randn(B, L, 256)), torch.randn(B, 256), torch.randn(B, L, 256)
    with torch.inference_mode():
        ref = O.denoiser_sample(flamed_sd, "prob_generator", cond, spk, noise, nfe, 0.3).transpose(1, 2)
    ts = torch.linspace(0, 1, nfe + 1)
    for rep in range(3):
        lat = engines["den16"].sample(cond, spk, noise, ts, 0.3, use_graph=False)
    assert _rel(lat, ref) < 1e-2
    assert _rel(engines["den32"].sample(cond, spk, noise, ts, 0.3, use_graph=False), ref) < 1e-5
