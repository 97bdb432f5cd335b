#!/usr/bin/env python3
"""Benchmark of the Flamed-TTS inference hot path on B200 (metric of BASELINE.json:
generated audio-seconds per wall-second at 128 denoiser steps).

    python bench.py --gpus 1 --steps K --warmup W          # this build (one rank per GPU under torchrun)
    python bench.py --impl reference ...                   # the reference algorithm's CPU path (oracle port)

A step = one pass of the hot path (durgen loop -> length regulator -> prior FFT glue -> cond fold ->
128-step denoiser loop -> FaCodec decode) over the `synthesize_via_metadata` workload of
BASELINE.json config 3: 256 synthetic LibriSpeech-length utterances (2-15 s) per GPU, 64 distinct 3 s
prompts, length-bucketed into batches of <= 64, bf16 tensor-core mode, random-init weights of the
configured architecture (no checkpoints offline).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402

SR, HOP = 16000, 200
METRIC = "audio_seconds_per_second_at_128_denoiser_steps"


def load_cfg():
    with open(os.path.join(ROOT, "configs", "prior.yaml")) as f:
        prior = yaml.safe_load(f)
    with open(os.path.join(ROOT, "configs", "prob.yaml")) as f:
        prob = yaml.safe_load(f)
    return {"prior_generator": prior, "prob_generator": prob}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def build_models(device, precision):
    from flamed import Flamed
    from flamed.models.facodec import FACodecDecoder, FACodecEncoder
    from flamed_tts_b200 import synthetic as W
    cfg = load_cfg()
    model = Flamed(cfg).eval()
    model.load_state_dict(W.make_flamed_state_dict(cfg["prior_generator"], cfg["prob_generator"], 0,
                                                   dur_bias=W.BENCH_DUR_BIAS, sil_bias=W.BENCH_SIL_BIAS))
    dec = FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2], vq_num_q_c=2,
                         vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8).eval()
    dec.load_state_dict(W.make_codec_decoder_state_dict(0))
    enc = FACodecEncoder(ngf=32, up_ratios=[2, 4, 5, 5], out_channels=256).eval()
    enc.load_state_dict(W.make_codec_encoder_state_dict(0))
    if device is not None:
        model.to(device).set_precision(precision)
        dec.to(device).set_precision(precision)
        enc.to(device)
    return cfg, model, enc, dec


def make_batches(args, rank, model, enc, dec, device):
    """per-rank workload -> list of host-side (pinned) batches; prompt features are computed once per distinct
    prompt (the reference caches them the same way, synthesize.py:108-125) and are INPUTS of sample_batch."""
    from flamed_tts_b200 import synthetic as W
    wl = W.metadata_workload(args.utterances, 64, seed=rank)
    with torch.inference_mode():
        codes, timbres = [], []
        for i in range(0, wl["prompts_wav"].shape[0], 16):
            e = enc(wl["prompts_wav"][i:i + 16].to(device))
            _, q, _, _, spk = dec(e, eval_vq=False, vq=True)
            codes.append(q.permute(1, 0, 2).cpu())
            timbres.append(spk.cpu())
        codes, timbres = torch.cat(codes), torch.cat(timbres)
    batches = []
    from flamed_tts_b200.parallel import bucket_by_length
    for idx in bucket_by_length([p.numel() for p in wl["phonemes"]], args.max_batch):
        ph = torch.nn.utils.rnn.pad_sequence([wl["phonemes"][i] for i in idx], batch_first=True, padding_value=0)
        sl = torch.tensor([wl["phonemes"][i].numel() for i in idx], dtype=torch.long)
        pi = [int(wl["prompt_of"][i]) for i in idx]
        batches.append(dict(phonemes=ph.pin_memory(), src_lens=sl.pin_memory(), prompts=codes[pi].contiguous().pin_memory(),
                            timbres=timbres[pi].contiguous().pin_memory(), idx=idx))
    return wl, batches


def run_step(model, dec, batches, args, device, from_host, host_out=None):
    """one pass over all batches.  from_host: inputs are copied from pinned host memory inside the step and
    the waveforms are read back into pinned host buffers (the e2e leg)."""
    tgt_lens, wavs = [], []
    if getattr(args, "pipelined", True):
        # the metadata entry point: Flamed.sample_batches overlaps the front stage (duration ODEs + the path's one host
        # sync) of bucket i+1 with the denoiser / codec kernels of bucket i
        def on_result(bi, out):
            tgt_lens.append((~out["tgt_mask"]).sum(1))
            if from_host:
                w = out["wav"]
                if host_out[bi] is None or host_out[bi].shape != w.shape:
                    host_out[bi] = torch.empty(w.shape, dtype=w.dtype).pin_memory()
                host_out[bi].copy_(w, non_blocking=True)
            wavs.append(out["wav"])
        model.sample_batches([b if from_host else b["dev"] for b in batches], codec_decoder=dec,
                             temp_durgen=args.temp_durgen, temp_denoiser=args.temp_denoiser,
                             nsteps_durgen=args.nsteps_durgen, nsteps_denoiser=args.nsteps_denoiser, on_result=on_result)
        return tgt_lens, wavs
    for bi, b in enumerate(batches):
        src = b if from_host else b["dev"]
        out = model.sample_batch(src["phonemes"].to(device, non_blocking=True), src["src_lens"].to(device, non_blocking=True),
                                 src["prompts"].to(device, non_blocking=True), src["timbres"].to(device, non_blocking=True),
                                 codec_decoder=dec, temp_durgen=args.temp_durgen, temp_denoiser=args.temp_denoiser,
                                 nsteps_durgen=args.nsteps_durgen, nsteps_denoiser=args.nsteps_denoiser)
        tgt_lens.append((~out["tgt_mask"]).sum(1))
        if from_host:
            w = out["wav"]
            if host_out[bi] is None or host_out[bi].shape != w.shape:
                host_out[bi] = torch.empty(w.shape, dtype=w.dtype).pin_memory()
            host_out[bi].copy_(w, non_blocking=True)
        wavs.append(out["wav"])
    return tgt_lens, wavs


def gather_wavs(wavs, rank, world):
    """the path's only collective: final NCCL gather of the waveforms to rank 0"""
    from flamed_tts_b200.parallel import gather_waveforms
    return gather_waveforms(wavs, rank, world)


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(args, n_utts=2, reps=1):
    """the oracle port (reference algorithm, PyTorch CPU fp32, all host threads) on a bounded sample of the
    same workload: `n_utts` utterances around the median length, same nfe / temperatures."""
    from oracle import flamed_oracle as O
    from flamed_tts_b200 import synthetic as W
    cfg = load_cfg()
    torch.set_num_threads(os.cpu_count() or 1)
    sd = W.make_flamed_state_dict(cfg["prior_generator"], cfg["prob_generator"], 0, dur_bias=W.BENCH_DUR_BIAS,
                                  sil_bias=W.BENCH_SIL_BIAS)
    dsd = W.make_codec_decoder_state_dict(0)
    wl = W.metadata_workload(args.utterances, 64, seed=0)
    order = sorted(range(len(wl["phonemes"])), key=lambda i: wl["phonemes"][i].numel())
    mid = len(order) // 2
    pick = order[mid - n_utts // 2: mid - n_utts // 2 + n_utts]
    ph = torch.nn.utils.rnn.pad_sequence([wl["phonemes"][i] for i in pick], batch_first=True, padding_value=0)
    sl = torch.tensor([wl["phonemes"][i].numel() for i in pick])
    g = torch.Generator().manual_seed(0)
    prompts = torch.randint(0, 1024, (len(pick), 6, 240), generator=g)  # prompt codes are inputs of the path
    timbres = torch.randn(len(pick), 256, generator=g)
    B, P = ph.shape
    best, audio = None, 0.0
    for _ in range(reps):
        torch.manual_seed(1)
        n_dur, n_sil = torch.randn((B, P)), torch.randn((B, P))
        t0 = time.perf_counter()
        with torch.inference_mode():
            out = O.sample_batch(sd, cfg, ph, sl, prompts, timbres, n_dur, n_sil, lambda b, l: torch.randn((b, l, 256)),
                                 args.nsteps_durgen, args.nsteps_denoiser, args.temp_durgen, args.temp_denoiser,
                                 codec_sd=dsd)
        dt = time.perf_counter() - t0
        audio = float(out["tgt_len"].sum()) * HOP / SR
        best = dt if best is None else min(best, dt)
    return dict(value=audio / best, unit="audio_s/s", cores=torch.get_num_threads(), kind="port",
                sample="%d median-length utterances of the workload (%.1f audio-s, padded batch), nfe %d/%d, fp32, "
                       "%.1f s wall" % (n_utts, audio, args.nsteps_durgen, args.nsteps_denoiser, best)), audio, best


def run_reference_arm(args, rank):
    if rank != 0:
        return
    times, audio = [], 0.0
    for i in range(args.warmup + args.steps):
        cb, audio, dt = cpu_baseline(args, n_utts=args.ref_utts)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000 * float(np.mean(times))
    v = audio / (ms / 1000)
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "audio_s/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "audio_s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "synthesize_via_metadata: %d synthetic LibriSpeech-length utterances (2-15 s) per GPU, 64 "
                        "distinct 3 s prompts, length-bucketed batches <= %d (BASELINE.json configs[2])" %
                        (args.utterances, args.max_batch),
            "nsteps_denoiser": args.nsteps_denoiser, "nsteps_durgen": args.nsteps_durgen,
            "temp_denoiser": args.temp_denoiser, "temp_durgen": args.temp_durgen, "precision": args.precision,
            "weights": "random-init of configs/{prior,prob,codec}.yaml (seeded; duration bias calibrated to 12 phonemes/s)", "noise": "device (torch cuda generator)",
            "l2": "per-step working set (>10 GB activations per batch) exceeds the 126 MB L2; no flush needed",
            "parallelism": "dp%d, one process per GPU, no collective in the loops, final NCCL waveform gather" % args.gpus}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utterances", type=int, default=256, help="utterances per GPU (weak scaling)")
    ap.add_argument("--max-batch", type=int, default=64)
    ap.add_argument("--nsteps-denoiser", type=int, default=128)
    ap.add_argument("--nsteps-durgen", type=int, default=16)
    ap.add_argument("--temp-denoiser", type=float, default=0.3)
    ap.add_argument("--temp-durgen", type=float, default=0.3)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-utts", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--sequential", dest="pipelined", action="store_false",
                    help="loop over Flamed.sample_batch instead of the pipelined Flamed.sample_batches")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 hot path has no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    import __graft_entry__
    __graft_entry__.build()
    from flamed_tts_b200 import _lib
    from flamed_tts_b200.engines import Context
    cfg, model, enc, dec = build_models(device, args.precision)
    model.set_noise_device("cuda")
    model.prob_generator.use_cuda_graph = "auto"
    wl, batches = make_batches(args, rank, model, enc, dec, device)
    for b in batches:
        b["dev"] = {k: b[k].to(device) for k in ("phonemes", "src_lens", "prompts", "timbres")}
    h2d = sum(sum(b[k].numel() * b[k].element_size() for k in ("phonemes", "src_lens", "prompts", "timbres")) for b in batches)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host, steps, warmup):
        host_out = [None] * len(batches)
        audio_s, d2h = 0.0, 0
        for _ in range(warmup):
            torch.manual_seed(1234 + rank)
            tl, wavs = run_step(model, dec, batches, args, device, from_host, host_out)
            if world > 1:
                gather_wavs(wavs, rank, world)
            del wavs
        barrier()
        lib = _lib.load_library()
        n0 = lib.flm_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            torch.manual_seed(1234 + rank)
            tl, wavs = run_step(model, dec, batches, args, device, from_host, host_out)
            if world > 1:
                gather_wavs(wavs, rank, world)
            d2h = sum(w.numel() * w.element_size() for w in wavs)
            del wavs
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.flm_launch_count() - n0
        audio_s = float(sum(int(t.sum()) for t in tl)) * HOP / SR  # valid frames only
        t = torch.tensor([ms, audio_s], device=device, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, audio_s = float(tm[0]), float(t[1])
        return ms / steps, audio_s, launches // max(steps, 1), d2h

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, audio_total, launches, _ = timed(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, audio_e2e, _, d2h = timed(True, args.steps, 1)

    roof, kernels = None, None
    if not args.no_profile:
        # one extra profiled pass (direct launches, CUDA events around every kernel of the library)
        ctx = Context.get(device)
        model.prob_generator.use_cuda_graph = False
        ctx.profile(True)
        torch.manual_seed(1234 + rank)
        t0 = time.perf_counter()
        run_step(model, dec, batches, args, device, False)
        torch.cuda.synchronize()
        prof_wall = (time.perf_counter() - t0) * 1000
        prof = ctx.profile_read()
        ctx.profile(False)
        pk = peaks()
        kernels = {}
        for name, r in prof.items():
            tensor_bound = name.startswith("tapgemm_tc")
            ach = (r["flops"] / (r["ms"] * 1e-3) / 1e12) if tensor_bound else (r["bytes"] / (r["ms"] * 1e-3) / 1e9)
            peak = pk["bf16_tflops_sustained"] if tensor_bound else pk["hbm_gbs"]
            kernels[name] = {"launches": r["launches"], "ms": round(r["ms"], 3), "share_of_step": round(r["ms"] / prof_wall, 4),
                             "bound": "tensor" if tensor_bound else "hbm", "achieved": round(ach, 2),
                             "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": round(ach / peak, 4)}
            if name in ("dwconv31_stats", "snake_act1d") and r["flops"] > 0:
                # these two are bound by the fp32 FMA pipes before HBM: 31 / ~26 FMA per element.
                # fp32 peak = SMs x 128 lanes x 2 FLOP x the SM clock measured during the run
                sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
                fpeak = torch.cuda.get_device_properties(device).multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
                fach = r["flops"] / (r["ms"] * 1e-3) / 1e12
                kernels[name].update({"fp32_tflops": round(fach, 2), "fp32_peak_tflops": round(fpeak, 2),
                                      "fp32_frac": round(fach / fpeak, 4), "binding": "fp32 FMA pipe"})
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])[0]
        k = kernels[top]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        traffic_detail = None
        if os.path.exists(tpath):  # DRAM bytes per launch of the top kernel from the committed ncu --set full capture
            traffic_detail = json.load(open(tpath)).get(top)
            traffic = traffic_detail.get("dram_bytes_read_plus_write_per_launch") if isinstance(traffic_detail, dict) else traffic_detail
        roof = {"kernel": top, "bound": k["bound"], "achieved": k["achieved"],
                "peak": pk["bf16_tflops_sustained"] if k["bound"] == "tensor" else pk["hbm_gbs"], "unit": k["unit"],
                "frac": k["frac"], "traffic": traffic, "peak_source": pk["source"] + (" (sustained bf16)" if k["bound"] == "tensor" else " (copy)"),
                "avg_launch_ms": round(prof[top]["ms"] / prof[top]["launches"], 4), "profiled_step_ms": round(prof_wall, 1),
                "traffic_detail": traffic_detail}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _, _ = cpu_baseline(args, n_utts=args.ref_utts)

    if rank == 0:
        line = {"metric": METRIC, "value": audio_total / (ms_dev / 1000), "unit": "audio_s/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": audio_e2e / (ms_e2e / 1000), "unit": "audio_s/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
                "gpu_launches": int(launches), "audio_seconds_per_step": audio_total, "clocks": clocks,
                "roofline": roof, "kernels": kernels, "cpu_baseline": cb}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
