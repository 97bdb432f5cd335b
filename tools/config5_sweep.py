#!/usr/bin/env python3
"""BASELINE.json configs[4]: step-count sweep on long-form utterances - 64 utterances of 30 s (P = 360 phonemes, about
2400 latent frames) per GPU, nsteps-denoiser in {8, 32, 64, 128} x nsteps-durgen in {4, 16, 64}, bf16 mode.
One JSON line per combination (same fields as bench.py's line, device-resident leg only).

    python tools/config5_sweep.py                                            # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/config5_sweep.py   # 8 GPUs, weak scaling + PCM gather
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
device = torch.device("cuda", local_rank)
torch.cuda.set_device(device)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=device)
args = bench.parse_args(["--workload", "config5"] + sys.argv[1:])
cfg, model, enc, dec = bench.build_models(device, "bf16")
model.set_noise_device(args.noise)
wl, n_pool = bench.global_workload(args, world)
codes, timbres = bench.prompt_codes(wl, enc, dec, device)
batches = bench.host_batches(wl, bench.rank_share(args, wl, rank, world), codes, timbres)
for b in batches:
    b["dev"] = {k: b[k].to(device) for k in ("phonemes", "src_lens", "prompts", "timbres")}
gatherer = None
if world > 1:
    from flamed_tts_b200.parallel import WavGather
    gatherer = WavGather(device, rank, world)


def step():
    torch.manual_seed(1234 + rank)
    st = bench.run_step(model, dec, batches, args, device, False, collect_pcm=gatherer is not None)
    done = None
    if gatherer is not None:
        _, _, done = gatherer.gather(st["pcm"])
    return st, done


for nd in (4, 16, 64):
    for nn in (8, 32, 64, 128):
        args.nsteps_durgen, args.nsteps_denoiser = nd, nn
        for _ in range(2):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 3
        e0.record()
        pend = []
        for _ in range(K):
            st, done = step()
            if done is not None:
                pend.append(done)
        for d in pend:
            torch.cuda.current_stream(device).wait_event(d)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / K, st["valid"] * bench.HOP / bench.SR, st["padded"] * bench.HOP / bench.SR],
                         device=device, dtype=torch.float64)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tm[0]
        if rank == 0:
            print(json.dumps({"metric": "audio_seconds_per_second", "value": float(t[1]) / (float(t[0]) / 1000), "unit": "audio_s/s",
                              "n_gpus": world, "steps": K, "warmup": 2, "ms_per_step": float(t[0]), "nsteps_denoiser": nn,
                              "nsteps_durgen": nd, "valid_audio_s": float(t[1]), "padded_audio_s": float(t[2]),
                              "config": bench.workload_config(args, world)["workload"], "dtype": "bf16"}), flush=True)
if world > 1:
    dist.destroy_process_group()
