"""GPU parity tests (-m gpu), round 2: parity AT THE SIZES THE BENCH RUNS (direct-launch path, L = 1200 / 2400),
the prior FFT decoders on the library kernels, the device-side noise map, the re-bucketed metadata path, the PCM
write-out and the boundary fixes (ProbabilisticModule.forward, LengthRegulator max_len, bounded graph cache).

Tolerances (stated per test): integers bit exact; fp32 mode rel-L2 <= 1e-5 per kernel; bf16 mode rel-L2 <= 1e-2
(waveform <= 3e-2)."""
import os

import numpy as np
import pytest
import torch

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


@pytest.fixture(scope="module")
def engines(ctx, cfg, flamed_sd, codec_dec_sd):
    from flamed_tts_b200.engines import CodecDecoderEngine, DenoiserEngine, DurationEngine
    pva = {k[len("prior_generator.pva."):]: v for k, v in flamed_sd.items() if k.startswith("prior_generator.pva.")}
    prob = {k[len("prob_generator."):]: v for k, v in flamed_sd.items() if k.startswith("prob_generator.")}
    return dict(dur=DurationEngine(ctx, pva),
                den32=DenoiserEngine(ctx, prob, cfg["prob_generator"], "fp32"),
                den16=DenoiserEngine(ctx, prob, cfg["prob_generator"], "bf16"),
                dec32=CodecDecoderEngine(ctx, codec_dec_sd, "fp32"),
                dec16=CodecDecoderEngine(ctx, codec_dec_sd, "bf16"))


@pytest.fixture(scope="module")
def dropin(cfg, flamed_sd, codec_dec_sd):
    from flamed import Flamed
    from flamed.models.facodec import FACodecDecoder
    model = Flamed(cfg).eval()
    model.load_state_dict(flamed_sd)
    model.to(DEV)
    dec = FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2], vq_num_q_c=2,
                         vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8).eval()
    dec.load_state_dict(codec_dec_sd)
    dec.to(DEV)
    return model, dec


# ------------------------------------------------------------------------------------------------ bench-scale parity
@pytest.mark.parametrize("B,L", [(4, 1200), (2, 2400)])
def test_velocity_at_bench_scale(engines, flamed_sd, B, L):
    """one SimpleMLPAdaLN evaluation through the direct-launch path at the frame counts the bench / config 5 use
    (M = 4800 rows: 38 row tiles, 38 / 75 depthwise chunks per sample, ragged last tile): fp32 <= 1e-5, bf16 <= 1e-2"""
    torch.manual_seed(60 + B)
    x, spk = torch.randn(B, L, 256), torch.randn(B, 256)
    t = 0.625
    with torch.inference_mode():
        ref = O.denoiser_forward(flamed_sd, "prob_generator.denoiser", x, torch.tensor([[t]]), spk)
    e32 = _rel(engines["den32"].forward(x, t, spk), ref)
    e16 = _rel(engines["den16"].forward(x, t, spk), ref)
    print("B=%d L=%d velocity rel-L2 fp32 %.3e bf16 %.3e" % (B, L, e32, e16))
    assert e32 < 1e-5 and e16 < 1e-2


def test_denoiser_loop_at_bench_scale(engines, flamed_sd):
    """8 Euler steps at B=3, L=1230 (3690 rows > the depth of every ring, direct launches) against the oracle, in the
    bf16 mode with both noise forms; the graph path must agree bit for bit with direct launches at this size too"""
    torch.manual_seed(71)
    B, L, nfe = 3, 1230, 8
    cond, spk, noise = torch.relu(torch.randn(B, L, 256)), torch.randn(B, 256), torch.randn(B, L, 256)
    with torch.inference_mode():
        ref = O.denoiser_sample(flamed_sd, "prob_generator", cond, spk, noise, nfe, 0.3).transpose(1, 2)
    ts = torch.linspace(0, 1, nfe + 1)
    lat = engines["den16"].sample(cond, spk, noise, ts, 0.3, use_graph=False)
    e16 = _rel(lat, ref)
    e32 = _rel(engines["den32"].sample(cond, spk, noise, ts, 0.3, use_graph=False), ref)
    print("B=3 L=1230 nfe=8 latents rel-L2 fp32 %.3e bf16 %.3e" % (e32, e16))
    assert e16 < 1e-2 and e32 < 1e-5
    g1 = engines["den16"].sample(cond, spk, noise, ts, 0.3, use_graph=True)   # first sighting: direct
    g2 = engines["den16"].sample(cond, spk, noise, ts, 0.3, use_graph=True)   # second: captured + replayed
    g3 = engines["den16"].sample(cond, spk, noise, ts, 0.3, use_graph=True)   # replay
    assert torch.equal(lat, g1) and torch.equal(lat, g2) and torch.equal(lat, g3)


@pytest.mark.parametrize("B,L", [(2, 1200), (1, 2400)])
def test_codec_decode_at_bench_scale(engines, codec_dec_sd, B, L):
    """FACodecDecoder.inference at T = 240 000 / 480 000 samples per utterance: fp32 <= 5e-5, bf16 <= 3e-2"""
    torch.manual_seed(80 + B)
    lat, spk = torch.randn(B, L, 256), torch.randn(B, 256)
    with torch.inference_mode():
        ref = O.codec_decode(codec_dec_sd, lat.transpose(1, 2), spk)
    w32 = engines["dec32"].decode(lat, spk)
    w16 = engines["dec16"].decode(lat, spk)
    e32, e16 = _rel(w32, ref), _rel(w16, ref)
    print("B=%d L=%d wav rel-L2 fp32 %.3e bf16 %.3e" % (B, L, e32, e16))
    assert list(w32.shape) == [B, 1, 200 * L] and e32 < 5e-5 and e16 < 3e-2


def test_prior_decoders_b200_path_vs_oracle(dropin, cfg, flamed_sd):
    """f1: PriorGenerator.decode_priors in the bf16 mode (FFT blocks on the tcgen05 conv GEMM / row-LN kernels +
    fused attention with per-sample key lengths) against the fp32 oracle on a ragged batch: prior_embs rel-L2 <= 2e-2
    (12 + 6 x 4 bf16 transformer layers), padded frames exactly zero, lazy logits consistent with the embeddings"""
    model, _ = dropin
    model.set_precision("bf16")
    try:
        torch.manual_seed(31)
        B, L, Lp = 3, 150, 40
        tgt = torch.tensor([150, 97, 128])
        x = torch.randn(B, L, 192)
        for b in range(B):
            x[b, tgt[b]:] = 0
        prompts = torch.randint(0, 1024, (B, 6, Lp))
        with torch.inference_mode():
            r_emb, r_log, r_mask = O.prior_after_pva(flamed_sd, "prior_generator", x, tgt, prompts, cfg["prior_generator"])
        pg = model.prior_generator
        emb, logits, mask = pg.decode_priors(x.to(DEV), tgt.to(DEV), prompts.to(DEV), Lp, bf16=True)
        assert torch.equal(mask.cpu(), r_mask)
        e = _rel(emb, r_emb)
        print("prior_embs (bf16 kernels) rel-L2 %.3e" % e)
        assert e < 2e-2
        for b in range(B):
            assert float(emb[b, :, tgt[b]:].abs().max()) == 0.0 if tgt[b] < L else True
        assert _rel(logits, r_log) < 3e-2
        lazy = pg.logits_from(emb, mask, bf16=True)
        assert torch.equal(lazy, logits)
    finally:
        model.set_precision("fp32")


# ------------------------------------------------------------------------------------------------ device noise (f2)
def test_philox_map_matches_the_oracle(ctx):
    """the documented seed -> N(0,1) map: device values against the numpy restatement (float64 Box-Muller): <= 2e-6
    absolute on |z| <= 5.3; and the integer core (same uniforms) shows as an exact match of the 24-bit mantissas"""
    from flamed_tts_b200.engines import philox_normal
    from oracle import philox as P
    for seed, tid, n in ((0, 0, 1000), (2 ** 61 + 12345, 2, 100003), (987654321, 1, 4097)):
        z = philox_normal(ctx, seed, tid, n).cpu().numpy()
        ref = P.normal(seed, tid, n)
        assert z.shape == ref.shape
        assert float(np.abs(z - ref).max()) < 2e-6 * (1 + float(np.abs(ref).max()))
    z = philox_normal(ctx, 7, 2, 1 << 20).cpu().numpy()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1) < 5e-3


def test_null_noise_equals_explicit_philox_noise(ctx, engines):
    """flm_denoiser_sample / flm_durgen_sample with NULL noise pointers + seed draw exactly the documented map: the
    result is bit-identical to passing that noise tensor explicitly (fp32 mode: mul then add, no contraction)"""
    from flamed_tts_b200.engines import philox_normal
    torch.manual_seed(5)
    B, L, nfe, seed = 2, 61, 3, 2 ** 40 + 99
    cond, spk = torch.relu(torch.randn(B, L, 256)), torch.randn(B, 256)
    ts = torch.linspace(0, 1, nfe + 1)
    noise = philox_normal(ctx, seed, 2, B * L * 256).view(B, L, 256)
    for name in ("den32", "den16"):
        a = engines[name].sample(cond, spk, noise, ts, 0.3, use_graph=False)
        b = engines[name].sample(cond, spk, None, ts, 0.3, use_graph=False, seed=seed)
        c = engines[name].sample(cond, spk, None, ts, 0.3, use_graph=False, seed=seed + 1)
        assert torch.equal(a, b) and not torch.equal(a, c)
    P = 23
    enc = torch.randn(B, P, 192)
    mask = O.get_mask_from_lengths(torch.tensor([23, 11]), P)
    nd, ns = philox_normal(ctx, seed, 0, B * P).view(B, P), philox_normal(ctx, seed, 1, B * P).view(B, P)
    ts = torch.linspace(0, 1, 5)
    a = engines["dur"].sample(enc, mask, nd, ns, ts, 0.3)
    for rep in range(3):  # first sighting direct, then captured, then replayed: the seed lives in device memory
        b = engines["dur"].sample(enc, mask, None, None, ts, 0.3, seed=seed)
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    c = engines["dur"].sample(enc, mask, None, None, ts, 0.3, seed=seed + 5)
    assert not torch.equal(a[2], c[2])


# ------------------------------------------------------------------------------------------------ boundary fixes
def test_probabilistic_module_forward_is_callable(dropin, flamed_sd):
    """pva.duration_generator(xt, enc, t, mask) - the call oracle/make_golden.py makes on the reference
    (pva.py:221-238) - exists on the drop-in and matches the oracle: <= 1e-5"""
    model, _ = dropin
    torch.manual_seed(8)
    B, P = 3, 29
    enc, xt = torch.randn(B, P, 192), torch.randn(B, P)
    mask = O.get_mask_from_lengths(torch.tensor([29, 5, 17]), P)
    pva = model.prior_generator.pva
    for name, mod in (("duration_generator", pva.duration_generator), ("sil_generator", pva.sil_generator)):
        for t in (0.0, 0.4375):
            with torch.inference_mode():
                ref = O.prob_module_forward(flamed_sd, "prior_generator.pva." + name, xt, enc, torch.tensor(t), mask)
            v = mod(xt.to(DEV), enc.to(DEV), torch.tensor(t), mask.to(DEV))
            assert _rel(v, ref) < 1e-5
            assert float(v.cpu()[mask].abs().max()) == 0.0
    v = pva.duration_generator(xt.to(DEV), enc.to(DEV), 0.25, None)
    with torch.inference_mode():
        ref = O.prob_module_forward(flamed_sd, "prior_generator.pva.duration_generator", xt, enc, torch.tensor(0.25), None)
    assert _rel(v, ref) < 1e-5


def test_length_regulator_max_len_pads_and_truncates(dropin):
    """LengthRegulator.LR(max_len): the reference's pad() uses F.pad with max_len - len, which TRUNCATES when negative
    (tools.py:299-317); tgt_len stays unclipped"""
    model, _ = dropin
    lr = model.prior_generator.pva.length_regulator
    torch.manual_seed(2)
    x = torch.randn(2, 5, 192)
    ph = torch.tensor([[2., 3., 1., 4., 2.], [1., 1., 2., 0., 0.]])
    si = torch.tensor([[0., 1., 0., 0., 2.], [0., 0., 1., 0., 0.]])
    sl = torch.tensor([5, 3])
    ref, ref_len = O.length_regulator(x, ph, si, sl)
    full, tl = lr(x.to(DEV), ph.to(DEV), si.to(DEV), sl.to(DEV), None)
    assert torch.equal(full.cpu(), ref) and torch.equal(tl.cpu(), ref_len)
    T = ref.shape[1]
    longer, tl2 = lr(x.to(DEV), ph.to(DEV), si.to(DEV), sl.to(DEV), T + 7)
    assert longer.shape[1] == T + 7 and torch.equal(longer[:, :T].cpu(), ref) and float(longer[:, T:].abs().max()) == 0
    shorter, tl3 = lr(x.to(DEV), ph.to(DEV), si.to(DEV), sl.to(DEV), T - 4)
    assert shorter.shape[1] == T - 4 and torch.equal(shorter.cpu(), ref[:, : T - 4])
    assert torch.equal(tl2.cpu(), ref_len) and torch.equal(tl3.cpu(), ref_len)


def test_graph_cache_is_bounded(engines):
    """keys are data dependent; 40 distinct (B, L) shapes, each seen three times, must neither fail nor
    change results (first sighting direct, second captured, LRU of 8 executables)"""
    torch.manual_seed(3)
    spk = torch.randn(1, 256)
    ts = torch.linspace(0, 1, 3)
    first = {}
    for rep in range(3):
        for L in range(20, 60):
            g = torch.Generator().manual_seed(L)
            cond, noise = torch.randn(1, L, 256, generator=g), torch.randn(1, L, 256, generator=g)
            out = engines["den16"].sample(cond, spk, noise, ts, 0.3, use_graph=True)
            if rep == 0:
                first[L] = out.clone()
            else:
                assert torch.equal(out, first[L])


# ------------------------------------------------------------------------------------------------ metadata path
def _utterances(n, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        P = int(torch.randint(5, 40, (1,), generator=g))
        out.append(dict(ph=torch.randint(1, 70, (P,), generator=g), prompts=torch.randint(0, 1024, (6, 30), generator=g),
                        timbre=torch.randn(256, generator=g)))
    return out


def _front_batches(utts, size):
    order = sorted(range(len(utts)), key=lambda i: -utts[i]["ph"].numel())
    batches = []
    for i in range(0, len(order), size):
        idx = order[i:i + size]
        ph = torch.nn.utils.rnn.pad_sequence([utts[j]["ph"] for j in idx], batch_first=True, padding_value=0)
        batches.append(dict(phonemes=ph, src_lens=torch.tensor([utts[j]["ph"].numel() for j in idx]),
                            prompts=torch.stack([utts[j]["prompts"] for j in idx]),
                            timbres=torch.stack([utts[j]["timbre"] for j in idx]), idx=idx))
    return batches


def test_rebucketed_metadata_path_matches_the_oracle(dropin, cfg, flamed_sd, codec_dec_sd):
    """Flamed.sample_batches(rebucket=True): an utterance's front result is the reference's on its front batch, its back
    result the reference's on its re-padded back batch.  fp32 mode, CPU draws in the documented order (all duration
    draws, then one latent draw per back batch): frame counts exact, latents <= 5e-5, wav <= 2e-4; every utterance
    delivered exactly once; PCM host copies = rint(wav * 32767)."""
    from flamed_tts_b200.parallel import bucket_by_rows
    model, dec = dropin
    model.set_precision("fp32").set_noise_device("cpu")
    dec.set_precision("fp32")
    utts = _utterances(11, 5)
    batches = _front_batches(utts, 4)
    kw = dict(temp_durgen=0.3, temp_denoiser=0.3, nsteps_durgen=4, nsteps_denoiser=3)
    torch.manual_seed(77)
    outs = model.sample_batches(batches, codec_decoder=dec, rebucket=True, row_budget=900, max_batch=5,
                                wav_to_host="pcm16", **kw)
    torch.cuda.synchronize()
    # ---- oracle: same draws in the same order
    torch.manual_seed(77)
    fronts = []
    with torch.inference_mode():
        for b in batches:
            B, P = b["phonemes"].shape
            nd, ns = torch.randn((B, P)), torch.randn((B, P))
            fronts.append(O.front_stage(flamed_sd, cfg, b["phonemes"], b["src_lens"], nd, ns, 4, 0.3))
        lens = [int(v) for f in fronts for v in f["tgt_len"]]
        owner = [(fi, r) for fi, f in enumerate(fronts) for r in range(f["tgt_len"].numel())]
        buckets = bucket_by_rows(lens, 900, 5)
        assert len(outs) == len(buckets) and len(buckets) >= 3
        seen = []
        for out, idx in zip(outs, buckets):
            src = [owner[j] for j in idx]
            assert out["index"] == src and out["tgt_lens"] == [lens[j] for j in idx]
            seen += src
            x = O.repad([fronts[fi]["x"][r] for fi, r in src], out["tgt_lens"])
            prompts = torch.stack([batches[fi]["prompts"][r] for fi, r in src])
            timbres = torch.stack([batches[fi]["timbres"][r] for fi, r in src])
            ref = O.back_stage(flamed_sd, cfg, x, torch.tensor(out["tgt_lens"]), prompts, timbres,
                               lambda b, l: torch.randn((b, l, 256)), 3, 0.3, codec_sd=codec_dec_sd)
            assert torch.equal(out["tgt_mask"].cpu(), ref["tgt_mask"])
            e_lat, e_wav = _rel(out["latents"], ref["latents"]), _rel(out["wav"], ref["wav"])
            assert e_lat < 5e-5 and e_wav < 2e-4, (e_lat, e_wav)
            out["wav_ready"].synchronize()
            want = torch.round(out["wav"].cpu() * 32767.0).to(torch.int16)
            assert out["wav_host"].dtype == torch.int16 and torch.equal(out["wav_host"], want)
            assert out["time"] > 0
        assert sorted(seen) == sorted(owner)


def test_rebucketed_path_bf16_philox_runs_and_is_reproducible(dropin):
    """throughput configuration of bench.py (bf16 kernels, Philox noise fused into the init kernels): two runs from the
    same torch seed are bit-identical, a different seed differs, tgt_lens agree with the masks"""
    model, dec = dropin
    model.set_precision("bf16").set_noise_device("philox")
    dec.set_precision("bf16")
    try:
        batches = _front_batches(_utterances(9, 6), 3)
        kw = dict(codec_decoder=dec, temp_durgen=0.3, temp_denoiser=0.3, nsteps_durgen=4, nsteps_denoiser=4, rebucket=True,
                  row_budget=700, max_batch=4)
        runs = []
        for seed in (1, 1, 2):
            torch.manual_seed(seed)
            runs.append(model.sample_batches(batches, **kw))
            torch.cuda.synchronize()
        for a, b in zip(runs[0], runs[1]):
            assert a["index"] == b["index"] and torch.equal(a["wav"], b["wav"])
            assert (~a["tgt_mask"]).sum(1).cpu().tolist() == a["tgt_lens"]
            assert bool(torch.isfinite(a["wav"]).all())
        assert any(a["tgt_lens"] != c["tgt_lens"] or not torch.equal(a["wav"], c["wav"]) for a, c in zip(runs[0], runs[2]))
    finally:
        model.set_precision("fp32").set_noise_device("cpu")
        dec.set_precision("fp32")


def test_wav_to_pcm16(ctx):
    from flamed_tts_b200.engines import wav_to_pcm16
    torch.manual_seed(0)
    for n in (1, 7, 4096, 100003):
        w = torch.tanh(torch.randn(n) * 2)
        w[:: max(1, n // 5)] = 1.0
        got = wav_to_pcm16(ctx, w.to(DEV)).cpu()
        want = torch.clamp(torch.round(w * 32767.0), -32768, 32767).to(torch.int16)
        assert torch.equal(got, want)


# ------------------------------------------------------------------------------------------------ prompt side (f3)
def test_prompt_side_codes_bit_exact_and_timbre(dropin, codec_dec_sd, golden_dir):
    """FACodecDecoder.forward(vq=True) through flm_codec_dec_prompt: the six code streams of the reference fixture bit
    exact (a code may differ only where the two best cosine similarities are within 1e-6: reported, none expected),
    timbre vector <= 2e-5, quantised sums <= 1e-5; and against the oracle on a larger seeded batch (B=3, T=173)"""
    _, dec = dropin
    dec.set_precision("fp32")
    g = np.load(os.path.join(golden_dir, "codec_encode.npz"))
    enc_out = torch.from_numpy(g["enc_out"])
    outs, codes, commit, bufs, spk = dec(enc_out.to(DEV), eval_vq=False, vq=True)
    assert codes.dtype == torch.int64 and list(codes.shape) == list(g["codes"].shape)
    assert torch.equal(codes.cpu(), torch.from_numpy(g["codes"]))
    assert _rel(spk, torch.from_numpy(g["timbre"])) < 2e-5
    assert len(bufs) == 3 and list(outs.shape) == [1, 256, enc_out.shape[-1]] and list(commit.shape) == [6, 1]
    torch.manual_seed(12)
    x = torch.randn(3, 256, 173)
    with torch.inference_mode():
        o_codes, o_spk = O.codec_prompt_features(codec_dec_sd, x)
        q0, _ = O._rvq(codec_dec_sd, "quantizer.0", x, 1)
        q1, _ = O._rvq(codec_dec_sd, "quantizer.1", x, 2)
        q2, _ = O._rvq(codec_dec_sd, "quantizer.2", x - (q0 + q1), 3)
    outs, codes, _, bufs, spk = dec(x.to(DEV), eval_vq=False, vq=True)
    diff = codes.cpu() != o_codes
    print("prompt codes differing from the oracle: %d of %d" % (int(diff.sum()), diff.numel()))
    assert int(diff.sum()) <= 2  # near-ties of the cosine argmax only
    if not diff.any():
        assert _rel(bufs[0], q0) < 1e-5 and _rel(bufs[1], q1) < 1e-5 and _rel(bufs[2], q2) < 1e-5
        assert _rel(outs, q0 + q1 + q2) < 1e-5
    assert _rel(spk, o_spk) < 2e-5


# ------------------------------------------------------------------------------------------------ attention (f1)
@pytest.mark.parametrize("B,S,lens", [(3, 200, [200, 1, 65]), (2, 64, [64, 63]), (4, 333, [333, 128, 129, 7]), (2, 1465, [1465, 900])])
def test_attention_prefix_kernel(ctx, B, S, lens):
    """flm_attention_bf16 (own flash-style kernel, 12 heads x 32, per-sample key prefixes) against fp64 softmax
    attention on the same bf16 operands: rel-L2 <= 6e-3 over the valid query rows (probabilities and the output are
    rounded to bf16), every output finite, masked keys have no influence"""
    from flamed_tts_b200.engines import attention_bf16
    H, dh = 12, 32
    g = torch.Generator().manual_seed(S + B)
    qkv = (torch.randn(B, S, 3, H, dh, generator=g) * 1.5).to(torch.bfloat16)
    kl = torch.tensor(lens, dtype=torch.int32)
    out = attention_bf16(ctx, qkv.to(DEV).contiguous(), kl.to(DEV)).float().cpu()
    q, k, v = (qkv[:, :, i].double().permute(0, 2, 1, 3) for i in range(3))  # (B,H,S,dh)
    sc = q @ k.transpose(-1, -2) / dh ** 0.5
    mask = torch.arange(S)[None, :] >= kl[:, None].long()
    sc = sc.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B, S, H * dh)
    assert bool(torch.isfinite(out).all())
    e = _rel(out, ref)
    print("attention B=%d S=%d rel-L2 %.3e" % (B, S, e))
    assert e < 6e-3
    # keys beyond the prefix must not matter: scramble them and compare bit for bit
    qkv2 = qkv.clone()
    for b in range(B):
        qkv2[b, lens[b]:, 1:] = torch.randn(S - lens[b], 2, H, dh, generator=g).to(torch.bfloat16) * 50
    out2 = attention_bf16(ctx, qkv2.to(DEV).contiguous(), kl.to(DEV)).float().cpu()
    assert torch.equal(out, out2)


# ------------------------------------------------------------------------------------------------ opt-in depthwise kernel
_DWTC_SCRIPT = r"""
import os, sys, torch, yaml
sys.path.insert(0, %(root)r)
from flamed_tts_b200 import synthetic as W
from flamed_tts_b200.engines import Context, DenoiserEngine
from oracle import flamed_oracle as O
root = %(root)r
prior = yaml.safe_load(open(os.path.join(root, "configs", "prior.yaml")))
prob = yaml.safe_load(open(os.path.join(root, "configs", "prob.yaml")))
sd = W.make_flamed_state_dict(prior, prob, 0)
psd = {k[len("prob_generator."):]: v for k, v in sd.items() if k.startswith("prob_generator.")}
den = DenoiserEngine(Context.get("cuda:0"), psd, prob, "bf16")
worst = 0.0
for B, L in ((1, 7), (3, 97), (2, 1200), (5, 64)):
    g = torch.Generator().manual_seed(B * 1000 + L)
    x, spk = torch.randn(B, L, 256, generator=g), torch.randn(B, 256, generator=g)
    with torch.inference_mode():
        ref = O.denoiser_forward(psd, "denoiser", x, torch.full((1, 1), 0.37), spk)
    v = den.forward(x.cuda(), 0.37, spk.cuda()).float().cpu()
    v2 = den.forward(x.cuda(), 0.37, spk.cuda()).float().cpu()
    assert torch.equal(v, v2), "not reproducible"
    e = float((v.double() - ref.double()).norm() / ref.double().norm())
    worst = max(worst, e)
print("WORST %%.4e" %% worst)
"""


@pytest.mark.parametrize("var,val", [("FLAMED_B200_DWCONV", "tc"), ("FLAMED_B200_MLPLN", "0")])
def test_alternative_denoiser_paths_match_the_oracle(var, val):
    """The two switches the denoiser reads when a handle is created (hence the subprocess), same bf16 tolerance as the
    production path (1e-2), at edge lengths (tiles that straddle samples, rows shorter than the conv window) and at a
    bench-sized frame count:
      FLAMED_B200_DWCONV=tc  the depthwise conv of the ConvNeXt blocks on the tensor cores (dwconv_tc.cu: Hankel operand out
                             of a channel time series, TMA-staged input, ldmatrix transposition, packed-bf16 LayerNorm)
                             instead of the FMA kernel - an experiment, slower as measured (profiles/r2w);
      FLAMED_B200_MLPLN=0    the MLP-branch LayerNorm as a pass of its own (ln_mod kernel) instead of the algebraic form
                             inside conv_3's / mlp.0's epilogues (the default)."""
    import subprocess
    import sys
    env = dict(os.environ)
    env[var] = val
    r = subprocess.run([sys.executable, "-c", _DWTC_SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    worst = float(r.stdout.strip().splitlines()[-1].split()[1])
    print("%s=%s: worst velocity rel-L2 %.3e" % (var, val, worst))
    assert worst < 1e-2
