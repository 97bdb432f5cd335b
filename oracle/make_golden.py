#!/usr/bin/env python3
"""Generate tests/golden/*.npz + state_dict_keys.json from the LIVE reference.

Runs only in the build container (needs /root/reference).  For each case it
  1. builds the reference modules (Flamed, FACodecEncoder, FACodecDecoder) and loads
     the seeded weights of oracle/weights.py with load_state_dict (strict for Flamed
     and the encoder; the codec decoder's training-only heads keep their own init);
  2. runs the reference's own entry points (prior_generator.sample,
     prob_generator.sample, codec_decoder.inference, codec_encoder.forward, ...);
  3. runs the oracle port (oracle/flamed_oracle.py) on the same inputs and the same
     three CPU randn draws and asserts agreement (exact for integers, <=1e-5 rel for
     floats: both are fp32 PyTorch-CPU, only op grouping differs);
  4. stores inputs + REFERENCE outputs as small fixtures.
tests/test_oracle_golden.py replays step 3 anywhere against the stored outputs.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
import json
import os
import sys
import warnings

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import flamed_oracle as O  # noqa: E402
from oracle import weights as W  # noqa: E402
from oracle.ref_import import REFERENCE_ROOT, import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEED = 0


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def load_cfg(root):
    with open(os.path.join(root, "configs", "prior.yaml")) as f:
        prior = yaml.safe_load(f)
    with open(os.path.join(root, "configs", "prob.yaml")) as f:
        prob = yaml.safe_load(f)
    return {"prior_generator": prior, "prob_generator": prob}


def build_reference(mods):
    cfg = load_cfg(REFERENCE_ROOT)
    torch.manual_seed(SEED)
    model = mods["flamed.models.flamed"].Flamed(cfg).eval()
    sd = W.make_flamed_state_dict(cfg["prior_generator"], cfg["prob_generator"], SEED)
    model.load_state_dict(sd, strict=True)
    fm = mods["flamed.models.facodec.facodec"]
    enc = fm.FACodecEncoder(ngf=32, up_ratios=[2, 4, 5, 5], out_channels=256).eval()
    dec = fm.FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2],
                            vq_num_q_c=2, vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8,
                            codebook_size_prosody=10, codebook_size_content=10, codebook_size_residual=10,
                            use_gr_x_timbre=True, use_gr_residual_f0=True, use_gr_residual_phone=True).eval()
    keys = {
        "flamed": {k: list(v.shape) for k, v in model.state_dict().items()},
        "codec_encoder": {k: list(v.shape) for k, v in enc.state_dict().items()},
        "codec_decoder": {k: list(v.shape) for k, v in dec.state_dict().items()},
    }
    esd = W.make_codec_encoder_state_dict(SEED)
    enc.load_state_dict(esd, strict=True)
    dsd = W.make_codec_decoder_state_dict(SEED)
    res = dec.load_state_dict(dsd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    heads = ("f0_predictor", "phone_predictor", "res_f0_predictor", "res_phone_predictor", "x_timbre_predictor")
    assert all(k.split(".")[0] in heads for k in res.missing_keys), res.missing_keys
    return cfg, model, enc, dec, sd, esd, dsd, keys


def sub(t, n=4096):
    """deterministic sub-sample of a big tensor (flattened, fixed stride) + its sum."""
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].clone()


def main():
    os.makedirs(GOLD, exist_ok=True)
    mods = import_reference()
    cfg, model, enc, dec, sd, esd, dsd, keys = build_reference(mods)
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    report = {}
    with torch.inference_mode():
        # ------------------------------------------------------------ case: full sample_batch (a1-a7)
        B, P, Lp = 2, 8, 12
        g = torch.Generator().manual_seed(11)
        phon = torch.randint(1, W.N_SYMBOLS, (B, P), generator=g)
        src_lens = torch.tensor([P, P - 2])
        phon[1, P - 2:] = 0
        prompts = torch.randint(0, 1024, (B, 6, Lp), generator=g)
        prompts[1, :, Lp - 3:] = 1024
        timbres = torch.randn(B, 256, generator=g)
        nfe_dur, nfe_den, t_dur, t_den = 4, 4, 0.3, 0.3
        torch.manual_seed(1234)
        ref = model.sample_batch(phonemes=phon, src_lens=src_lens, prompts=prompts, timbres=timbres,
                                 codec_decoder=dec, temp_durgen=t_dur, temp_denoiser=t_den,
                                 nsteps_durgen=nfe_dur, nsteps_denoiser=nfe_den)
        # intermediate reference values through the reference's own sub-module entry points
        src_mask = O.get_mask_from_lengths(src_lens, P)
        ref_enc = model.prior_generator.encoder(phon, src_mask)
        torch.manual_seed(1234)
        n_dur = torch.randn((B, P))
        n_sil = torch.randn((B, P))
        L = ref["latents"].shape[-1]
        n_lat = torch.randn((B, L, 256))
        # reference durations: replay the loop with the reference module forward
        pva = model.prior_generator.pva
        ts = torch.linspace(0, 1, nfe_dur + 1)
        d, s = n_dur * t_dur, n_sil * t_dur
        for i in range(1, nfe_dur + 1):
            d = d + (1 / nfe_dur) * pva.duration_generator(d, ref_enc, ts[i - 1], src_mask)
            s = s + (1 / nfe_dur) * pva.sil_generator(s, ref_enc, ts[i - 1], src_mask)
        ref_phone = torch.clamp(torch.round(torch.exp(d) - 1), min=0)
        ref_sil = torch.clamp(torch.round(torch.exp(s) - 1), min=0)
        ref_lr, ref_tgt_len = pva.length_regulator(ref_enc, ref_phone, ref_sil, src_lens, None)
        ref_cond = model.prob_generator.cond_downsampling(
            model.prob_generator.quantizer_encoding(ref["prior_embs"]), ~ref["tgt_mask"].unsqueeze(-1))
        ref_v0 = model.prob_generator.denoiser(n_lat * t_den + ref_cond, ts[0].unsqueeze(0).unsqueeze(1), timbres)

        orc = O.sample_batch(sd, cfg, phon, src_lens, prompts, timbres, n_dur, n_sil, lambda b, l: n_lat,
                             nfe_dur, nfe_den, t_dur, t_den, codec_sd=dsd)
        idx, tl = O.length_regulator_index(ref_phone, ref_sil, src_lens)
        o_lr, _ = O.length_regulator(ref_enc, ref_phone, ref_sil, src_lens)
        o_v0 = O.denoiser_forward(sd, "prob_generator.denoiser", n_lat * t_den + ref_cond, ts[0].view(1, 1), timbres)
        assert torch.equal(orc["phone_dur"], ref_phone) and torch.equal(orc["sil_dur"], ref_sil)
        assert torch.equal(tl, ref_tgt_len) and torch.equal(orc["tgt_len"], ref_tgt_len)
        assert torch.equal(o_lr, ref_lr), "LR closed form differs from reference"
        assert torch.equal(orc["tgt_mask"], ref["tgt_mask"])
        report["sample_batch"] = {
            "enc": rel_l2(orc["enc"], ref_enc), "dur_t": rel_l2(orc["dur_t"], d),
            "prior_embs": rel_l2(orc["prior_embs"], ref["prior_embs"]),
            "prior_logits": rel_l2(orc["prior_logits"], ref["prior_logits"]),
            "cond": rel_l2(orc["cond"], ref_cond), "v0": rel_l2(o_v0, ref_v0),
            "latents": rel_l2(orc["latents"], ref["latents"]), "wav": rel_l2(orc["wav"], ref["wav"]),
        }
        for k, v in report["sample_batch"].items():
            assert v < 2e-5, (k, v)
        np.savez(os.path.join(GOLD, "sample_batch.npz"),
                 phonemes=phon.numpy(), src_lens=src_lens.numpy(), prompts=prompts.numpy(), timbres=timbres.numpy(),
                 nfe=np.array([nfe_dur, nfe_den]), temps=np.array([t_dur, t_den]), noise_seed=np.array(1234),
                 enc=ref_enc.numpy(), dur_t=d.numpy(), sil_t=s.numpy(), phone_dur=ref_phone.numpy(),
                 sil_dur=ref_sil.numpy(), tgt_len=ref_tgt_len.numpy(), lr_index=idx.numpy(),
                 prior_embs_sub=sub(ref["prior_embs"]).numpy(), prior_logits_sub=sub(ref["prior_logits"]).numpy(),
                 cond=ref_cond.numpy(), v0_sub=sub(ref_v0).numpy(), latents=ref["latents"].numpy(),
                 wav_sub=sub(ref["wav"], 16384).numpy(), wav_shape=np.array(ref["wav"].shape))

        # ------------------------------------------------------------ case: length regulator edge cases (a4)
        lr_cases = []
        g = torch.Generator().manual_seed(5)
        for (b, p) in [(1, 1), (3, 9), (4, 17)]:
            x = torch.randn(b, p, 192, generator=g)
            ph = torch.randint(0, 6, (b, p), generator=g).float()
            si = (torch.randint(0, 4, (b, p), generator=g) * (torch.rand(b, p, generator=g) < 0.4)).float()
            sl = torch.randint(1, p + 1, (b,), generator=g)
            if b > 1:
                sl[0] = p
                ph[1] = 0  # all-zero phone durations -> clamp(min=1)
            out, tl = pva.length_regulator(x, ph, si, sl, None)
            idx, tl2 = O.length_regulator_index(ph, si, sl)
            o, _ = O.length_regulator(x, ph, si, sl)
            assert torch.equal(tl, tl2) and torch.equal(o, out)
            lr_cases.append((x, ph, si, sl, idx, tl))
        np.savez(os.path.join(GOLD, "length_regulator.npz"), n=np.array(len(lr_cases)),
                 **{f"{nm}{i}": c[j].numpy() for i, c in enumerate(lr_cases)
                    for j, nm in enumerate(["x", "phone", "sil", "src_lens", "index", "tgt_len"])})

        # ------------------------------------------------------------ case: Activation1d (a8)
        x = torch.randn(2, 64, 37, generator=g) * 2
        act = dec.model[5]  # final Activation1d, 64 channels
        y_ref = act(x)
        y_o = O.activation1d(dsd, "model.5", x)
        report["activation1d"] = rel_l2(y_o, y_ref)
        assert report["activation1d"] < 1e-6
        np.savez(os.path.join(GOLD, "activation1d.npz"), x=x.numpy(), y=y_ref.numpy())

        # ------------------------------------------------------------ case: codec encoder + prompt features (a9, f3)
        wav = torch.randn(1, 1, 3200, generator=g) * 0.1
        e_ref = enc(wav)
        e_o = O.codec_encode(esd, wav)
        report["codec_encode"] = rel_l2(e_o, e_ref)
        assert report["codec_encode"] < 2e-5
        _, qs, _, _, spk = dec(e_ref, eval_vq=False, vq=True)
        codes_o, spk_o = O.codec_prompt_features(dsd, e_ref)
        assert torch.equal(codes_o, qs)
        report["timbre"] = rel_l2(spk_o, spk)
        assert report["timbre"] < 2e-5
        np.savez(os.path.join(GOLD, "codec_encode.npz"), wav=wav.numpy(), enc_out=e_ref.numpy(),
                 codes=qs.numpy(), timbre=spk.numpy())

        # ------------------------------------------------------------ case: codec decoder alone (a7)
        lat = torch.randn(1, 256, 9, generator=g)
        spk1 = torch.randn(1, 256, generator=g)
        w_ref = dec.inference(lat, spk1)
        w_o = O.codec_decode(dsd, lat, spk1)
        report["codec_decode"] = rel_l2(w_o, w_ref)
        assert report["codec_decode"] < 2e-5
        np.savez(os.path.join(GOLD, "codec_decode.npz"), latents=lat.numpy(), spk=spk1.numpy(), wav=w_ref.numpy())

    with open(os.path.join(GOLD, "oracle_vs_reference.json"), "w") as f:
        json.dump({"torch": torch.__version__, "seed": SEED, "rel_l2_oracle_vs_reference": report}, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
