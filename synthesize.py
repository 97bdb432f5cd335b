#!/usr/bin/env python3
"""Flamed-TTS synthesis CLI on the B200-native hot path.

Same flags, modes, output naming and RTF report as the reference's synthesize.py (328-404):
  direct mode   --text ... --prompt-list a.wav b.wav --prompt-dir DIR
  batch mode    --metadata-file meta.txt (lines `target|prompt|text`) --prompt-dir DIR [--batch-size N]
plus `--precision {bf16,fp32}`, `--bucket` (sort the pending metadata entries by text length before
batching, so that a batch pads little; the reference batches in file order), `--rebucket` (after the duration
stage, re-group ALL pending utterances by their real frame count into row-budgeted batches and write each
waveform trimmed to its own length) and `--writers N` (PCM_16 files are written by a thread pool from int16
samples converted on the device).  Launched under torchrun (one process per GPU) the pending entries are dealt
across the ranks by cost; every rank writes its own files, `--gather true` ships the PCM to rank 0 over NCCL
first.  WAV files are written directly (canonical PCM_16, what soundfile produces); the model config is read
with OmegaConf when installed and PyYAML otherwise.
"""
import argparse
import math
import os

import numpy as np
import torch

from flamed import Flamed
from flamed.models.facodec import FACodecDecoder, FACodecEncoder

SR = 16000
HERE = os.path.dirname(os.path.abspath(__file__))


def str2bool(v):
    if isinstance(v, bool):
        return v
    s = str(v).strip().lower()
    if s in ("true", "1", "yes", "y"):
        return True
    if s in ("false", "0", "no", "n"):
        return False
    raise argparse.ArgumentTypeError("Cannot interpret %r as boolean." % (v,))


def build_arg_parser():
    p = argparse.ArgumentParser(description="Flamed-TTS synthesis (B200-native hot path).")
    p.add_argument("--ckpt-path", required=True)
    p.add_argument("--cfg-path", required=True)
    p.add_argument("--text", default=None)
    p.add_argument("--prompt-list", nargs="+", default=None)
    p.add_argument("--prompt-dir", "--input-dir", dest="prompt_dir", default=None)
    p.add_argument("--metadata-file", "--text-file", dest="metadata_file", default=None)
    p.add_argument("--output-dir", default=".")
    p.add_argument("--weights-only", type=str2bool, default=True)
    p.add_argument("--nsteps-durgen", type=int, default=64)
    p.add_argument("--nsteps-denoiser", type=int, default=64)
    p.add_argument("--temp-durgen", type=float, default=0.3)
    p.add_argument("--temp-denoiser", type=float, default=0.3)
    p.add_argument("--device", default="cuda:0")
    p.add_argument("--skip-existing", type=str2bool, default=True)
    p.add_argument("--batch-size", type=int, default=4)
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--bucket", type=str2bool, default=False)
    p.add_argument("--rebucket", type=str2bool, default=False)
    p.add_argument("--row-budget", type=int, default=32768)
    p.add_argument("--max-batch", type=int, default=64)
    p.add_argument("--writers", type=int, default=8)
    p.add_argument("--gather", type=str2bool, default=False)
    p.add_argument("--codec-encoder-ckpt", default=os.path.join(HERE, "flamed", "models", "facodec", "checkpoints", "ns3_facodec_encoder.bin"))
    p.add_argument("--codec-decoder-ckpt", default=os.path.join(HERE, "flamed", "models", "facodec", "checkpoints", "ns3_facodec_decoder.bin"))
    return p


def load_cfg(path):
    try:
        from omegaconf import OmegaConf
        return OmegaConf.to_container(OmegaConf.load(path), resolve=True)
    except ImportError:
        import yaml
        with open(path) as f:
            return yaml.safe_load(f)


def write_wav(path, wav):
    """float waveform -> PCM_16 WAV, the bytes `sf.write(path, wav, SR)` produces (reference synthesize.py:211,296)"""
    from flamed_tts_b200.wavio import pcm16_from_float, write_wav_pcm16
    write_wav_pcm16(path, pcm16_from_float(wav), SR)


def get_codec(device, enc_ckpt, dec_ckpt):
    enc = FACodecEncoder(ngf=32, up_ratios=[2, 4, 5, 5], out_channels=256)
    dec = FACodecDecoder(in_channels=256, upsample_initial_channel=1024, ngf=32, up_ratios=[5, 5, 4, 2], vq_num_q_c=2,
                         vq_num_q_p=1, vq_num_q_r=3, vq_dim=256, codebook_dim=8, codebook_size_prosody=10,
                         codebook_size_content=10, codebook_size_residual=10, use_gr_x_timbre=True,
                         use_gr_residual_f0=True, use_gr_residual_phone=True)
    enc.load_state_dict(torch.load(enc_ckpt, map_location="cpu"))
    dec.load_state_dict(torch.load(dec_ckpt, map_location="cpu"))
    return enc.to(device).eval(), dec.to(device).eval()


def prompt_features(model, enc, dec, path, cache):
    if path not in cache:
        with torch.inference_mode():
            e = enc(model._preprocess_acoustic_prompt(path, sr=SR))
            _, codes, _, _, timbre = dec(e, eval_vq=False, vq=True)
        cache[path] = (codes.permute(1, 0, 2)[0].cpu(), timbre[0].cpu())
    return cache[path]


def run_prompts(model, enc, dec, a):
    os.makedirs(a.output_dir, exist_ok=True)
    rtfs = []
    for name in a.prompt_list:
        path = name if os.path.isabs(name) else os.path.join(a.prompt_dir, name)
        r = model.sample(text=a.text, prompt_raw=path, sr=SR, codec_encoder=enc, codec_decoder=dec,
                         nsteps_durgen=a.nsteps_durgen, nsteps_denoiser=a.nsteps_denoiser, temp_durgen=a.temp_durgen,
                         temp_denoiser=a.temp_denoiser)
        stem = os.path.splitext(os.path.basename(name))[0]
        out = "%s-%s-%s-%s-%s.wav" % (stem, a.nsteps_durgen, a.nsteps_denoiser, a.temp_durgen, a.temp_denoiser)
        write_wav(os.path.join(a.output_dir, out), r["wav"])
        rtfs.append(r["time"] / (len(r["wav"]) / SR))
    return sum(rtfs) / len(rtfs) if rtfs else None


def _dist():
    """(rank, world) of a torchrun launch; initialises the NCCL process group on first use"""
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world == 1:
        return 0, 1
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    return dist.get_rank(), world


def run_metadata(model, enc, dec, a):
    from flamed_tts_b200.parallel import bucket_by_length, bucket_cost, deal_buckets
    from flamed_tts_b200.wavio import WavWriter
    target_dir = os.path.join(a.output_dir, "nfe%s-temp%s" % (a.nsteps_denoiser, a.temp_denoiser))
    os.makedirs(target_dir, exist_ok=True)
    pending = []
    with open(a.metadata_file, encoding="utf-8") as f:
        for line in (ln.strip() for ln in f):
            if not line:
                continue
            parts = line.split("|", 2)
            if len(parts) != 3:
                print("[WARN] Malformed line skipped: %s" % line)
                continue
            out_path = os.path.join(target_dir, parts[0])
            if a.skip_existing and os.path.exists(out_path):
                continue
            prompt = parts[1] if os.path.isabs(parts[1]) else os.path.join(a.prompt_dir, parts[1])
            pending.append((out_path, prompt, parts[2]))
    if not pending:
        return None
    rank, world = _dist()
    if a.bucket or a.rebucket or world > 1:
        pending.sort(key=lambda e: -len(e[2]))
    chunks = [list(range(i, min(i + a.batch_size, len(pending)))) for i in range(0, len(pending), a.batch_size)]
    if world > 1:  # one global pool, buckets dealt to the ranks by descending cost; no exchange until the end
        tl = [len(e[2]) for e in pending]
        chunks = deal_buckets(tl, chunks, world)[rank]
    cache, rtfs = {}, []
    pad_code = model.prior_generator.config["codec"]["vocab_size"]
    groups, tensors = [], []
    for idx in chunks:
        batch = [pending[i] for i in idx]
        seqs = [model._preprocess_english(t)[0].squeeze(0).cpu() for _, _, t in batch]
        feats = [prompt_features(model, enc, dec, pth, cache) for _, pth, _ in batch]
        phon = torch.nn.utils.rnn.pad_sequence(seqs, batch_first=True, padding_value=0)
        lens = torch.tensor([s.numel() for s in seqs], dtype=torch.long)
        lp = max(c.shape[-1] for c, _ in feats)
        prompts = torch.full((len(batch), feats[0][0].shape[0], lp), pad_code, dtype=feats[0][0].dtype)
        for j, (c, _) in enumerate(feats):
            prompts[j, :, : c.shape[-1]] = c
        groups.append(batch)
        tensors.append(dict(phonemes=phon, src_lens=lens, prompts=prompts, timbres=torch.stack([t for _, t in feats])))

    gathered = []  # (--gather) per-utterance (path, n_samples) in the order the PCM is concatenated
    pcm_dev = []
    with WavWriter(workers=a.writers, sr=SR) as writer:
        def on_result(i, out):
            """batch i is complete on the device and its PCM is on its way to pinned host memory"""
            if a.rebucket:  # samples come from several input batches; each waveform is trimmed to its own frames
                items = [(groups[fi][r][0], n * 200) for (fi, r), n in zip(out["index"], out["tgt_lens"])]
            else:           # the reference writes the padded batch length for every item (synthesize.py:294-296)
                items = [(g[0], out["wav"].shape[-1]) for g in groups[i]]
            per_item = out["time"] / len(items)
            if a.gather and world > 1:
                from flamed_tts_b200.engines import Context, wav_to_pcm16
                pcm = wav_to_pcm16(Context.get(model.device), out["wav"])
                for j, (path, n) in enumerate(items):
                    pcm_dev.append(pcm[j, 0, :n])
                    gathered.append((path, n))
            else:
                host = out["wav_host"].numpy()
                for j, (path, n) in enumerate(items):
                    writer.submit(path, host[j, 0, :n], ready=out["wav_ready"])
            rtfs.extend(per_item / (n / SR) for _, n in items)

        # Flamed.sample_batches = the loop of sample_batch calls of reference synthesize.py:270-299, pipelined
        model.sample_batches(tensors, codec_decoder=dec, temp_durgen=a.temp_durgen, temp_denoiser=a.temp_denoiser,
                             nsteps_durgen=a.nsteps_durgen, nsteps_denoiser=a.nsteps_denoiser, on_result=on_result,
                             rebucket=a.rebucket, row_budget=a.row_budget, max_batch=a.max_batch,
                             wav_to_host=None if (a.gather and world > 1) else "pcm16")
        if a.gather and world > 1:
            _gather_and_write(model, writer, pcm_dev, gathered, rank, world)
    return sum(rtfs) / len(rtfs) if rtfs else None


def _gather_and_write(model, writer, pcm_dev, items, rank, world):
    """final NCCL gather of the PCM to rank 0 (flm_gather_wav), which writes every file"""
    import torch.distributed as dist
    from flamed_tts_b200.parallel import WavGather
    g = WavGather(model.device, rank, world)
    flat, counts, done = g.gather(pcm_dev)
    names = [None] * world
    dist.all_gather_object(names, items)
    if rank != 0:
        done.synchronize()
        return
    host = torch.empty(flat.shape, dtype=torch.int16, pin_memory=True)
    torch.cuda.current_stream(model.device).wait_event(done)
    host.copy_(flat, non_blocking=True)
    torch.cuda.current_stream(model.device).synchronize()
    arr, off = host.numpy(), 0
    for r in range(world):
        for path, n in names[r]:
            writer.submit(path, arr[off:off + n])
            off += n


def main(args=None):
    parser = build_arg_parser()
    cli = args is None
    a = parser.parse_args() if cli else args
    if getattr(a, "prompt_dir", None) is None and hasattr(a, "input_dir"):
        a.prompt_dir = a.input_dir
    try:
        if (a.metadata_file is not None) == (a.prompt_list is not None):
            raise ValueError("Specify either --prompt-list (direct mode) or --metadata-file (batch mode), but not both.")
        if a.prompt_dir is None:
            raise ValueError("--prompt-dir/--input-dir is required.")
        if a.prompt_list is not None and not a.text:
            raise ValueError("--text is required when using --prompt-list.")
        if a.metadata_file is not None:
            if not os.path.isfile(a.metadata_file):
                raise ValueError("Metadata file not found: %s" % a.metadata_file)
            if a.batch_size < 1:
                raise ValueError("--batch-size must be >= 1.")
    except ValueError as e:
        if cli:
            parser.error(str(e))
        raise
    device = torch.device(a.device)
    if int(os.environ.get("WORLD_SIZE", 1)) > 1 and device.type == "cuda":  # torchrun: one process per GPU
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        torch.cuda.set_device(device)
    if device.type == "cuda" and not torch.cuda.is_available():
        raise SystemExit("CUDA is not available: the B200-native hot path has no CPU fallback.")
    enc, dec = get_codec(device, a.codec_encoder_ckpt, a.codec_decoder_ckpt)
    cfg = load_cfg(a.cfg_path)
    model = Flamed.from_pretrained(cfg=cfg, ckpt_path=a.ckpt_path, device=device, weights_only=a.weights_only)
    model.to(device).set_precision(a.precision)
    dec.set_precision(a.precision)
    rtf = run_metadata(model, enc, dec, a) if a.metadata_file else run_prompts(model, enc, dec, a)
    if int(os.environ.get("RANK", 0)) != 0:
        return rtf
    if rtf is not None:
        print("=" * 20, "Avg RTF", "=" * 20)
        print(">" * 5, "RTF:", round(rtf, 3))
    else:
        print("No samples were generated.")
    return rtf


if __name__ == "__main__":
    main()
