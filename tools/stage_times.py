#!/usr/bin/env python3
"""Per-stage device time of one sample_batch (B=64 bucket of the bench workload), CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from flamed.utils.tools import get_mask_from_lengths  # noqa: E402


class A:
    utterances, max_batch = 256, 64
    nsteps_durgen, nsteps_denoiser, temp_durgen, temp_denoiser = 16, int(os.environ.get("NFE", 128)), 0.3, 0.3


dev = torch.device("cuda:0")
cfg, model, enc, dec = bench.build_models(dev, "bf16")
model.set_noise_device("cuda")
wl, batches = bench.make_batches(A, 0, model, enc, dec, dev)
b = batches[int(os.environ.get("BATCH", 1))]
ph, sl, pr, tb = (b[k].to(dev) for k in ("phonemes", "src_lens", "prompts", "timbres"))
pg, pb = model.prior_generator, model.prob_generator


def run(record):
    ev = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append((name, e))

    with torch.inference_mode():
        mark("start")
        src_mask = get_mask_from_lengths(sl, ph.size(-1))
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            e = pg.encoder(ph, src_mask)
        mark("phoneme_encoder(glue)")
        x, tgt = pg.pva.sample(e, sl, src_mask, nfe=A.nsteps_durgen, temperature=A.temp_durgen)
        mark("durgen+LR")
        embs, logits, tmask = pg.decode_priors(x, tgt, pr, pr.size(-1), bf16=True)
        mark("prior_decoders(glue)")
        eng = pb.engine()
        c = eng.cond_prepare(embs, ~tmask.unsqueeze(-1))
        mark("cond_prepare")
        ts = torch.linspace(0, 1, A.nsteps_denoiser + 1)
        noise = torch.randn((c.shape[0], c.shape[1], 256), device=dev)
        mark("noise")
        lat = eng.sample(c, tb, noise, ts, A.temp_denoiser, use_graph=False)
        mark("denoiser_loop")
        wav = dec.inference(lat.transpose(1, 2), tb)
        mark("codec_decode")
    torch.cuda.synchronize()
    if record:
        print("B=%d P=%d L=%d audio=%.1f s" % (ph.shape[0], ph.shape[1], c.shape[1], float(tgt.sum()) * 200 / 16000))
        tot = ev[0][1].elapsed_time(ev[-1][1])
        for (n0, e0), (n1, e1) in zip(ev, ev[1:]):
            ms = e0.elapsed_time(e1)
            print("%-26s %9.2f ms  %5.1f %%" % (n1, ms, 100 * ms / tot))
        print("%-26s %9.2f ms" % ("total", tot))


run(False)
run(True)
