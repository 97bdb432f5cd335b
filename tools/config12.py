#!/usr/bin/env python3
"""BASELINE.json configs[0] and configs[1]: ONE utterance (80 phonemes, one 3 s prompt, random-init weights).

  config 1: the unmodified reference's `Flamed.sample` on the host cores (nsteps-durgen 16, nsteps-denoiser 64): latency.
  config 2: the same utterance through the drop-in on one B200 at nsteps-denoiser 128, temp 0.3 - fp32 mode against the
            reference's output for the same torch seed (durations bit exact, waveform rel-L2 <= 2e-4), then the bf16
            mode (waveform rel-L2 <= 3e-2) - with the latency of each.
One JSON line per measurement.  Needs oracle/_ref (python oracle/make_ref.py in the build container)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from flamed_tts_b200 import synthetic as W  # noqa: E402
from oracle import ref_import  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
cfg = bench.load_cfg()
sd, dsd, esd = bench.make_weights(cfg)
rng = np.random.default_rng(0)
phon = torch.from_numpy(rng.integers(64, 148, size=80))
prompt = torch.from_numpy((np.random.default_rng(1).standard_normal((1, 1, 48000)) * 0.1).astype(np.float32))


def rel(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


# ---------------------------------------------------------------- the reference on the CPU
_, rmodel, renc, rdec = ref_import.build_reference_models(sd, codec_dec_sd=dsd, codec_enc_sd=esd, device="cpu")
ref_out = {}
for nd, nn in ((16, 64), (16, 128)):
    torch.manual_seed(5)
    t0 = time.perf_counter()
    with torch.inference_mode():
        r = rmodel.sample(phonemes=phon, prompt_raw=prompt, sr=16000, codec_encoder=renc, codec_decoder=rdec,
                          nsteps_durgen=nd, nsteps_denoiser=nn, temp_durgen=0.3, temp_denoiser=0.3)
    dt = time.perf_counter() - t0
    ref_out[nn] = r["wav"]
    print(json.dumps({"config": "1 (reference, CPU)" if nn == 64 else "2 (reference side, CPU)", "impl": "reference",
                      "nsteps_durgen": nd, "nsteps_denoiser": nn, "latency_s": dt, "audio_s": len(r["wav"]) / 16000,
                      "audio_s_per_s": len(r["wav"]) / 16000 / dt, "cores": torch.get_num_threads()}), flush=True)

if not torch.cuda.is_available():
    sys.exit(0)
# ---------------------------------------------------------------- the drop-in on one B200
from flamed import Flamed  # noqa: E402
from flamed.models.facodec import FACodecDecoder, FACodecEncoder  # noqa: E402

dev = torch.device("cuda:0")
model = Flamed(cfg).eval()
model.load_state_dict(sd)
model.to(dev)
dec = FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2], vq_num_q_c=2, vq_num_q_p=1,
                     vq_num_q_r=3, vq_dim=256, codebook_dim=8).eval()
dec.load_state_dict(dsd)
dec.to(dev)
enc = FACodecEncoder(ngf=32, up_ratios=[2, 4, 5, 5], out_channels=256).eval()
enc.load_state_dict(esd)
enc.to(dev)
for prec, tol in (("fp32", 2e-4), ("bf16", 3e-2)):
    model.set_precision(prec).set_noise_device("cpu")
    dec.set_precision(prec)
    for nn in (64, 128):
        lat = []
        for rep in range(3):
            torch.manual_seed(5)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = model.sample(phonemes=phon, prompt_raw=prompt, sr=16000, codec_encoder=enc, codec_decoder=dec,
                             nsteps_durgen=16, nsteps_denoiser=nn, temp_durgen=0.3, temp_denoiser=0.3)
            torch.cuda.synchronize()
            lat.append(time.perf_counter() - t0)
        same_len = len(r["wav"]) == len(ref_out[nn])
        e = rel(r["wav"], ref_out[nn]) if same_len else float("nan")
        ok = same_len and e < tol
        print(json.dumps({"config": "2 (drop-in, 1 x B200)", "impl": "flamed_b200", "precision": prec, "nsteps_durgen": 16,
                          "nsteps_denoiser": nn, "latency_s_first_call": lat[0], "latency_s": min(lat[1:]),
                          "audio_s": len(r["wav"]) / 16000, "audio_s_per_s": len(r["wav"]) / 16000 / min(lat[1:]),
                          "same_frame_count_as_reference": same_len, "wav_rel_l2_vs_reference": e, "tolerance": tol,
                          "ok": ok}), flush=True)
        if not ok:
            sys.exit(1)
