#!/usr/bin/env python3
"""Benchmark of the Flamed-TTS inference hot path on B200 (metric of BASELINE.json:
generated audio-seconds per wall-second at 128 denoiser steps).

    python bench.py --gpus 1 --steps K --warmup W          # this build (one rank per GPU under torchrun)
    python bench.py --impl reference ...                   # the reference's own CPU path (oracle/_ref, else the port)
    python bench.py --impl eager ...                       # the reference's own PyTorch code on the same B200
    python bench.py --workload config4|config5 ...         # BASELINE.json configs[3] / configs[4]

A step = one pass of the hot path (phoneme encoder -> durgen loops -> length regulator -> prior FFT decoders -> cond
fold -> 128-step denoiser loop -> FaCodec decode) over a `synthesize_via_metadata` workload:

  config3 (default)  BASELINE.json configs[2]: 256 synthetic LibriSpeech-length utterances (2-15 s) per GPU, 64 distinct
                     3 s prompts.  With N GPUs ONE global pool of 256*N utterances is length-bucketed and its buckets are
                     dealt to the ranks by descending cost (flamed_tts_b200.parallel.deal_buckets): weak scaling.
  config4            BASELINE.json configs[3]: one fixed pool of 4096 utterances, same partitioner: strong scaling.
  config5            BASELINE.json configs[4]: 64 long-form 30 s utterances (P=360) per GPU.

bf16 tensor-core mode, random-init weights of the configured architecture (no checkpoints offline).  Inside a rank
the utterances are re-grouped after the duration stage by their real frame counts into row-budgeted batches
(Flamed.sample_batches(rebucket=True)).  Prints ONE JSON line (rank 0).
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
def make_weights(cfg):
    from flamed_tts_b200 import synthetic as W
    return (W.make_flamed_state_dict(cfg["prior_generator"], cfg["prob_generator"], 0, dur_bias=W.BENCH_DUR_BIAS,
                                     sil_bias=W.BENCH_SIL_BIAS),
            W.make_codec_decoder_state_dict(0), W.make_codec_encoder_state_dict(0))


def build_models(device, precision):
    from flamed import Flamed
    from flamed.models.facodec import FACodecDecoder, FACodecEncoder
    cfg = load_cfg()
    sd, dsd, esd = make_weights(cfg)
    model = Flamed(cfg).eval()
    model.load_state_dict(sd)
    dec = FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2], vq_num_q_c=2,
                         vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8).eval()
    dec.load_state_dict(dsd)
    enc = FACodecEncoder(ngf=32, up_ratios=[2, 4, 5, 5], out_channels=256).eval()
    enc.load_state_dict(esd)
    if device is not None:
        model.to(device).set_precision(precision)
        dec.to(device).set_precision(precision)
        enc.to(device)
    return cfg, model, enc, dec


def global_workload(args, world):
    """the global utterance pool of the run -> (workload dict, number of utterances)"""
    from flamed_tts_b200 import synthetic as W
    if args.workload == "config4":
        n = args.pool
        return W.metadata_workload(n, 64, seed=0), n
    if args.workload == "config5":
        n = args.utterances_long * world
        return W.metadata_workload(n, 64, seed=0, dur_range=(30.0, 30.0)), n
    n = args.utterances * world
    return W.metadata_workload(n, 64, seed=0), n


def rank_share(args, wl, rank, world):
    """front buckets (<= max_batch utterances, sorted by phoneme count) of the global pool, dealt to the ranks by
    descending cost; returns this rank's list of index lists.  No data-path collective: every rank derives the same
    deal from the same seeded pool."""
    from flamed_tts_b200.parallel import bucket_by_length, deal_buckets
    lens = [p.numel() for p in wl["phonemes"]]
    buckets = bucket_by_length(lens, args.front_batch)
    return deal_buckets(lens, buckets, world)[rank]


def prompt_codes(wl, enc, dec, device):
    """prompt features are computed once per distinct prompt (the reference caches them the same way,
    synthesize.py:108-125) and are INPUTS of sample_batch"""
    with torch.inference_mode():
        codes, timbres = [], []
        for i in range(0, wl["prompts_wav"].shape[0], 16):
            e = enc(wl["prompts_wav"][i:i + 16].to(device))
            _, q, _, _, spk = dec(e, eval_vq=False, vq=True)
            codes.append(q.permute(1, 0, 2).cpu())
            timbres.append(spk.cpu())
    return torch.cat(codes), torch.cat(timbres)


def host_batches(wl, share, codes, timbres, pin=True):
    out = []
    pm = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    for idx in share:
        ph = torch.nn.utils.rnn.pad_sequence([wl["phonemes"][i] for i in idx], batch_first=True, padding_value=0)
        sl = torch.tensor([wl["phonemes"][i].numel() for i in idx], dtype=torch.long)
        pi = [int(wl["prompt_of"][i]) for i in idx]
        out.append(dict(phonemes=pm(ph), src_lens=pm(sl), prompts=pm(codes[pi].contiguous()),
                        timbres=pm(timbres[pi].contiguous()), idx=idx))
    return out


def run_step(model, dec, batches, args, device, from_host, collect_pcm=False):
    """one pass over this rank's batches -> dict(valid_frames, padded_frames, d2h_bytes, pcm).  from_host: inputs are
    copied from pinned host memory inside the step and every waveform is read back as PCM_16 into pinned host memory
    (the e2e leg)."""
    st = dict(valid=0, padded=0, d2h=0, pcm=[], masks=[], ready=[])

    def on_result(bi, out):
        w = out["wav"]
        st["padded"] += w.shape[0] * (w.shape[-1] // HOP)
        if "tgt_lens" in out:
            st["valid"] += int(sum(out["tgt_lens"]))
        else:
            st["masks"].append((~out["tgt_mask"]).sum())
        if from_host:
            st["d2h"] += out["wav_host"].numel() * out["wav_host"].element_size()
            st["ready"].append(out["wav_ready"])
        if collect_pcm:
            from flamed_tts_b200.engines import Context, wav_to_pcm16
            st["pcm"].append(wav_to_pcm16(Context.get(device), w))

    model.sample_batches([b if from_host else b["dev"] for b in batches], codec_decoder=dec,
                         temp_durgen=args.temp_durgen, temp_denoiser=args.temp_denoiser,
                         nsteps_durgen=args.nsteps_durgen, nsteps_denoiser=args.nsteps_denoiser, on_result=on_result,
                         rebucket=args.rebucket, row_budget=args.row_budget, max_batch=args.max_batch,
                         batch_overhead_rows=args.batch_overhead_rows, wave_rows=args.wave_rows,
                         wav_to_host="pcm16" if from_host else None)
    for ev in st["ready"]:
        ev.synchronize()  # the PCM of every batch has landed in host memory
    if st["masks"]:
        st["valid"] += int(torch.stack(st["masks"]).sum())
    return st


# ------------------------------------------------------------------------------------------------ reference arms
def reference_models(device):
    """the UNMODIFIED reference (oracle/_ref or /root/reference) with the bench weights, or None if not importable"""
    try:
        from oracle import ref_import
        if not ref_import.reference_available():
            return None
        cfg = load_cfg()
        sd, dsd, _ = make_weights(cfg)
        _, model, _, dec = ref_import.build_reference_models(sd, codec_dec_sd=dsd, device=device)
        return model, dec
    except Exception as e:  # noqa: BLE001 - the arm falls back to the port and says so
        sys.stderr.write("bench.py: reference import failed (%s: %s); using the oracle port\n" % (type(e).__name__, e))
        return None


def median_sample(args, n_utts):
    """`n_utts` utterances around the median length of the config-3 pool + seeded prompt codes / timbres"""
    from flamed_tts_b200 import synthetic as W
    wl = W.metadata_workload(args.utterances, 64, seed=0)
    order = sorted(range(len(wl["phonemes"])), key=lambda i: wl["phonemes"][i].numel())
    mid = len(order) // 2
    pick = order[mid - n_utts // 2: mid - n_utts // 2 + n_utts]
    ph = torch.nn.utils.rnn.pad_sequence([wl["phonemes"][i] for i in pick], batch_first=True, padding_value=0)
    sl = torch.tensor([wl["phonemes"][i].numel() for i in pick])
    g = torch.Generator().manual_seed(0)
    prompts = torch.randint(0, 1024, (len(pick), 6, 240), generator=g)  # prompt codes are inputs of the path
    timbres = torch.randn(len(pick), 256, generator=g)
    return ph, sl, prompts, timbres


def cpu_baseline(args, n_utts=2, ref=None):
    """the reference's CPU path (fp32, all host threads) on a bounded sample of the same workload: `n_utts` utterances
    around the median length, same nfe / temperatures.  kind "reference": the unmodified reference code imported
    from oracle/_ref; kind "port": the oracle restatement (same PyTorch CPU ops) when that copy is absent."""
    torch.set_num_threads(os.cpu_count() or 1)
    ph, sl, prompts, timbres = median_sample(args, n_utts)
    B, P = ph.shape
    if ref is None:
        ref = reference_models("cpu")
    torch.manual_seed(1)
    t0 = time.perf_counter()
    with torch.inference_mode():
        if ref is not None:
            model, dec = ref
            out = model.sample_batch(phonemes=ph, src_lens=sl, prompts=prompts, timbres=timbres, codec_decoder=dec,
                                     temp_durgen=args.temp_durgen, temp_denoiser=args.temp_denoiser,
                                     nsteps_durgen=args.nsteps_durgen, nsteps_denoiser=args.nsteps_denoiser)
            frames = int((~out["tgt_mask"]).sum())
            kind = "reference"
        else:
            from oracle import flamed_oracle as O
            cfg = load_cfg()
            sd, dsd, _ = make_weights(cfg)
            n_dur, n_sil = torch.randn((B, P)), torch.randn((B, P))
            out = O.sample_batch(sd, cfg, ph, sl, prompts, timbres, n_dur, n_sil, lambda b, l: torch.randn((b, l, 256)),
                                 args.nsteps_durgen, args.nsteps_denoiser, args.temp_durgen, args.temp_denoiser,
                                 codec_sd=dsd)
            frames = int(out["tgt_len"].sum())
            kind = "port"
    dt = time.perf_counter() - t0
    audio = frames * HOP / SR
    return dict(value=audio / dt, unit="audio_s/s", cores=torch.get_num_threads(), kind=kind,
                sample="%d median-length utterances of the workload (%.1f audio-s, one padded batch), nfe %d/%d, fp32, "
                       "%.1f s wall" % (n_utts, audio, args.nsteps_durgen, args.nsteps_denoiser, dt)), audio, dt, ref


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation on the box's host cores (rank 0 only)"""
    if rank != 0:
        return
    times, audio, ref, cb = [], 0.0, None, None
    for i in range(args.warmup + args.steps):
        cb, audio, dt, ref = cpu_baseline(args, n_utts=args.ref_utts, ref=ref)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000 * float(np.mean(times))
    v = audio / (ms / 1000)
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "audio_s/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1, arm="reference"), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "audio_s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def eager_pass(ref, batches, args, device, autocast):
    """one pass of the reference's own sample_batch loop (synthesize.py:270-299) on `device` -> (ms, valid frames)"""
    model, dec = ref
    frames = []
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        for b in batches:
            out = model.sample_batch(phonemes=b["phonemes"].to(device), src_lens=b["src_lens"].to(device),
                                     prompts=b["prompts"].to(device), timbres=b["timbres"].to(device), codec_decoder=dec,
                                     temp_durgen=args.temp_durgen, temp_denoiser=args.temp_denoiser,
                                     nsteps_durgen=args.nsteps_durgen, nsteps_denoiser=args.nsteps_denoiser)
            frames.append((~out["tgt_mask"]).sum())
            del out
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1), int(torch.stack(frames).sum())


def gpu_eager_baseline(args, device, batches, passes=1):
    """PyTorch-eager on the same B200: the unmodified reference with device='cuda' on `batches` (host dicts), torch
    defaults (nn.Linear in IEEE fp32, cuDNN convs in TF32) and under bf16 autocast.  Returns a dict or None."""
    ref = reference_models(device)
    if ref is None:
        return None
    res = {"impl": "unmodified reference (oracle/_ref), PyTorch eager, device=cuda", "batches": len(batches),
           "utterances": int(sum(b["phonemes"].shape[0] for b in batches))}
    for name, ac in (("fp32_tf32_defaults", False), ("bf16_autocast", True)):
        torch.manual_seed(1)
        eager_pass(ref, batches[:1], args, device, ac)  # warm-up: cuDNN heuristics, allocator
        best, frames = None, 0
        for _ in range(passes):
            ms, frames = eager_pass(ref, batches, args, device, ac)
            best = ms if best is None else min(best, ms)
        res[name] = {"value": frames * HOP / SR / (best / 1000), "unit": "audio_s/s", "ms": best,
                     "audio_s": frames * HOP / SR}
    del ref
    torch.cuda.empty_cache()
    return res


def run_eager_arm(args, rank, local_rank, world):
    """--impl eager: every rank runs the reference's loop over its share of the same pool; rank 0 prints"""
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    import __graft_entry__
    __graft_entry__.build()
    cfg, model, enc, dec = build_models(device, "fp32")
    wl, n_pool = global_workload(args, world)
    codes, timbres = prompt_codes(wl, enc, dec, device)
    del model, enc, dec
    batches = host_batches(wl, rank_share(args, wl, rank, world), codes, timbres, pin=False)
    ref = reference_models(device)
    if ref is None:
        if rank == 0:
            print(json.dumps({"impl": "eager", "unavailable": "oracle/_ref (copy of the reference) is not present"}), flush=True)
        return
    ac = args.precision == "bf16"
    for _ in range(args.warmup):
        eager_pass(ref, batches[:1], args, device, ac)
    ms, frames = 0.0, 0
    for _ in range(args.steps):
        m, frames = eager_pass(ref, batches, args, device, ac)
        ms += m
    t = torch.tensor([ms / max(args.steps, 1), frames * HOP / SR], device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t[0] = tm[0]
    if rank == 0:
        v = float(t[1]) / (float(t[0]) / 1000)
        print(json.dumps({"impl": "eager", "metric": METRIC, "value": v, "unit": "audio_s/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t[0]), "higher_is_better": True,
                          "scaling": "strong" if args.workload == "config4" else "weak", "vs_baseline": None,
                          "dtype": "bf16 autocast" if ac else "f32 (Linear IEEE fp32, cuDNN conv TF32: torch defaults)",
                          "data": "synthetic", "config": workload_config(args, world, arm="eager"),
                          "e2e": {"value": v, "unit": "audio_s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def workload_config(args, world, arm="ours"):
    if args.workload == "config4":
        wl = ("synthesize_via_metadata: ONE pool of %d synthetic LibriSpeech-length utterances (2-15 s), 64 distinct 3 s "
              "prompts, length-bucketed (<= %d per front bucket) and dealt to %d rank(s) by descending cost "
              "(BASELINE.json configs[3])" % (args.pool, args.front_batch, world))
    elif args.workload == "config5":
        wl = ("long-form: %d utterances of 30 s (P=360) per GPU, 64 distinct 3 s prompts (BASELINE.json configs[4])"
              % args.utterances_long)
    else:
        wl = ("synthesize_via_metadata: one pool of %d x %d synthetic LibriSpeech-length utterances (2-15 s), 64 distinct "
              "3 s prompts, length-bucketed (<= %d per front bucket) and dealt to the ranks by descending cost "
              "(BASELINE.json configs[2] per GPU)" % (args.utterances, world, args.front_batch))
    c = {"workload": wl, "nsteps_denoiser": args.nsteps_denoiser, "nsteps_durgen": args.nsteps_durgen,
         "temp_denoiser": args.temp_denoiser, "temp_durgen": args.temp_durgen,
         "weights": "random-init of configs/{prior,prob,codec}.yaml (seeded; duration bias calibrated to 12 phonemes/s)"}
    if arm == "reference":
        c.update(precision="fp32 (PyTorch CPU)", noise="CPU torch.randn (the reference's own draws)",
                 batching="one padded batch of the sampled utterances")
    elif arm == "eager":
        c.update(precision=("bf16 autocast" if args.precision == "bf16" else "torch defaults: fp32 Linear, TF32 cuDNN conv"),
                 noise="CPU torch.randn + H2D (the reference's own draws)",
                 batching="the reference's loop: one sample_batch call per front bucket, no re-bucketing")
    else:
        c.update(precision=args.precision,
                 noise="device: Philox4x32-10 fused into the init kernels (seed from torch's generator)" if args.noise == "philox" else "device (torch cuda generator)",
                 batching=("re-bucketed by real frame count after the duration stage: <= %d samples and <= %d padded rows "
                           "per back batch, cuts minimising padded rows + %d rows per batch"
                           % (args.max_batch, args.row_budget, args.batch_overhead_rows)) if args.rebucket else
                          "one sample_batch per front bucket (pipelined)",
                 l2="per-step working set (>5 GB activations per batch) exceeds the 126 MB L2; no flush needed",
                 parallelism="dp%d, one process per GPU, no collective in the loops, final NCCL gather of the PCM_16 "
                             "waveforms (flm_gather_wav)" % world)
    return c


# ------------------------------------------------------------------------------------------------ main
def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--workload", default="config3", choices=["config3", "config4", "config5"])
    ap.add_argument("--utterances", type=int, default=256, help="config3: utterances per GPU (weak scaling)")
    ap.add_argument("--pool", type=int, default=4096, help="config4: size of the global pool (strong scaling)")
    ap.add_argument("--utterances-long", type=int, default=64, help="config5: 30 s utterances per GPU")
    ap.add_argument("--front-batch", type=int, default=64, help="utterances per front (duration-stage) bucket")
    ap.add_argument("--max-batch", type=int, default=64, help="samples per back (denoiser / codec) batch")
    ap.add_argument("--wave-rows", type=int, default=4736,
                    help="rows of one GEMM wave (74 CTA pairs x 256 rows / 4 N tiles) in the batch cost model; 0 = ignore waves")
    ap.add_argument("--batch-overhead-rows", type=int, default=2500,
                    help="cost of one more back batch in padded rows (parallel.bucket_by_rows)")
    ap.add_argument("--row-budget", type=int, default=32768, help="padded rows (B x L) per back batch")
    ap.add_argument("--no-rebucket", dest="rebucket", action="store_false")
    ap.add_argument("--nsteps-denoiser", type=int, default=128)
    ap.add_argument("--nsteps-durgen", type=int, default=16)
    ap.add_argument("--temp-denoiser", type=float, default=0.3)
    ap.add_argument("--temp-durgen", type=float, default=0.3)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--noise", default="philox", choices=["philox", "cuda"])
    ap.add_argument("--ref-utts", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 hot path has no CPU fallback); use --impl reference for the CPU arm")
    if args.impl == "eager":
        run_eager_arm(args, rank, local_rank, world)
        return
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
    model.set_noise_device(args.noise)
    model.prob_generator.use_cuda_graph = "auto"
    wl, n_pool = global_workload(args, world)
    codes, timbres = prompt_codes(wl, enc, dec, device)
    batches = host_batches(wl, rank_share(args, wl, rank, world), codes, timbres)
    for b in batches:
        b["dev"] = {k: b[k].to(device) for k in ("phonemes", "src_lens", "prompts", "timbres")}
    h2d = sum(sum(b[k].numel() * b[k].element_size() for k in ("phonemes", "src_lens", "prompts", "timbres")) for b in batches)
    gatherer = None
    if world > 1:
        from flamed_tts_b200.parallel import WavGather
        gatherer = WavGather(device, rank, world)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def one(from_host):
        torch.manual_seed(1234 + rank)
        st = run_step(model, dec, batches, args, device, from_host, collect_pcm=gatherer is not None)
        done = None
        if gatherer is not None:  # the path's only collective; runs on a side stream under the next step's kernels
            _, _, done = gatherer.gather(st["pcm"])
            st["pcm"] = None
        return st, done

    def timed(from_host, steps, warmup):
        for _ in range(warmup):
            one(from_host)
        barrier()
        lib = _lib.load_library()
        n0 = lib.flm_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st, pend = None, []
        for _ in range(steps):
            st, done = one(from_host)
            if done is not None:
                pend.append(done)
        for d in pend:
            torch.cuda.current_stream(device).wait_event(d)  # every gather has landed before the clock stops
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.flm_launch_count() - n0
        t = torch.tensor([ms, st["valid"] * HOP / SR, st["padded"] * HOP / SR, float(st["d2h"])], device=device,
                         dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tm[0]
        return dict(ms=float(t[0]) / steps, valid_s=float(t[1]), padded_s=float(t[2]), d2h=int(t[3]),
                    launches=launches // max(steps, 1))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_leg = timed(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_leg = None if args.no_e2e else timed(True, args.steps, 1)

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
        model.prob_generator.use_cuda_graph = "auto"
        pk = peaks()
        kernels = {}
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fpeak = torch.cuda.get_device_properties(device).multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
        for name, r in prof.items():
            tensor_bound = name.startswith("tapgemm_tc")
            fma_bound = name == "tapgemm_fp32_fma"
            if tensor_bound:
                ach, peak, unit, bound = r["flops"] / (r["ms"] * 1e-3) / 1e12, pk["bf16_tflops_sustained"], "TFLOP/s", "tensor"
            elif fma_bound:  # an FMA GEMM: bound by the fp32 pipes (SMs x 128 lanes x 2 FLOP x measured SM clock)
                ach, peak, unit, bound = r["flops"] / (r["ms"] * 1e-3) / 1e12, fpeak, "TFLOP/s", "fp32"
            else:
                ach, peak, unit, bound = r["bytes"] / (r["ms"] * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s", "hbm"
            kernels[name] = {"launches": r["launches"], "ms": round(r["ms"], 3), "share_of_step": round(r["ms"] / prof_wall, 4),
                             "bound": bound, "achieved": round(ach, 2), "unit": unit, "peak": round(peak, 1),
                             "frac": round(ach / peak, 4)}
            if name in ("dwconv31_stats", "dwconv31_fused", "snake_act1d") and r["flops"] > 0:
                # bound by the fp32 FMA pipes before HBM
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
        roof = {"kernel": top, "bound": k["bound"], "achieved": k["achieved"], "peak": k["peak"], "unit": k["unit"],
                "frac": k["frac"], "traffic": traffic,
                "peak_source": pk["source"] + (" (sustained bf16)" if k["bound"] == "tensor" else " (copy)"),
                "avg_launch_ms": round(prof[top]["ms"] / prof[top]["launches"], 4), "profiled_step_ms": round(prof_wall, 1),
                "traffic_detail": traffic_detail}

    eager = None
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        # bounded sample: the median-cost front bucket of this rank through the unmodified reference on the same GPU
        order = sorted(range(len(batches)), key=lambda i: batches[i]["phonemes"].shape[0] * batches[i]["phonemes"].shape[1])
        mid = batches[order[len(order) // 2]]
        del model, dec, enc
        torch.cuda.empty_cache()
        try:
            eager = gpu_eager_baseline(args, device, [mid])
            if eager is not None:
                eager["sample"] = "the median-cost front bucket (%d utterances) of the %d of this run" % (
                    mid["phonemes"].shape[0], len(batches))
        except Exception as e:  # noqa: BLE001 - a failing baseline must not lose the measured line
            eager = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _, _, _ = cpu_baseline(args, n_utts=args.ref_utts)

    if rank == 0:
        line = {"metric": METRIC, "value": dev_leg["valid_s"] / (dev_leg["ms"] / 1000), "unit": "audio_s/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_leg["ms"], "higher_is_better": True,
                "scaling": "strong" if args.workload == "config4" else "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": None if e2e_leg is None else {
                    "value": e2e_leg["valid_s"] / (e2e_leg["ms"] / 1000), "unit": "audio_s/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": e2e_leg["d2h"], "ms_per_step": e2e_leg["ms"],
                    "d2h_format": "PCM_16 (converted on the device; what the reference's sf.write stores)"},
                "gpu_launches": int(dev_leg["launches"]), "audio_seconds_per_step": dev_leg["valid_s"],
                "padding": {"valid_audio_s": dev_leg["valid_s"], "padded_audio_s": dev_leg["padded_s"],
                            "ratio": round(dev_leg["padded_s"] / max(dev_leg["valid_s"], 1e-9), 4),
                            "value_on_padded_audio": dev_leg["padded_s"] / (dev_leg["ms"] / 1000)},
                "utterances": n_pool, "clocks": clocks, "roofline": roof, "kernels": kernels,
                "gpu_eager_baseline": eager, "cpu_baseline": cb}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
