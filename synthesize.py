#!/usr/bin/env python3
"""Flamed-TTS synthesis CLI on the B200-native hot path.

Same flags, modes, output naming and RTF report as the reference's synthesize.py (328-404):
  direct mode   --text ... --prompt-list a.wav b.wav --prompt-dir DIR
  batch mode    --metadata-file meta.txt (lines `target|prompt|text`) --prompt-dir DIR [--batch-size N]
plus `--precision {bf16,fp32}` and `--bucket` (sort the pending metadata entries by text length before
batching, so that a batch pads little; the reference batches in file order).
Audio I/O uses soundfile when installed and scipy otherwise; the model config is read with OmegaConf when
installed and PyYAML otherwise.
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
    try:
        import soundfile as sf
        sf.write(path, wav, SR)
    except ImportError:
        from scipy.io import wavfile
        wavfile.write(path, SR, (np.clip(wav, -1, 1) * 32767).astype(np.int16))


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


def run_metadata(model, enc, dec, a):
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
    if a.bucket:
        pending.sort(key=lambda e: -len(e[2]))
    cache, rtfs = {}, []
    pad_code = model.prior_generator.config["codec"]["vocab_size"]
    groups, tensors = [], []
    for i in range(0, len(pending), a.batch_size):
        batch = pending[i:i + a.batch_size]
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

    def on_result(i, out):  # bucket i is enqueued (bucket i+1's front stage already overlaps it): write it out
        per_item = out["time"] / len(groups[i])
        for (out_path, _, _), w in zip(groups[i], out["wav"]):
            wav = w[0].detach().cpu().numpy()  # padded batch length, as the reference writes it
            write_wav(out_path, wav)
            rtfs.append(per_item / (len(wav) / SR))

    # Flamed.sample_batches = a loop of sample_batch calls (reference synthesize.py:270-299) with the front stage of
    # the next bucket overlapped with the denoiser / codec kernels of the current one
    model.sample_batches(tensors, codec_decoder=dec, temp_durgen=a.temp_durgen, temp_denoiser=a.temp_denoiser,
                         nsteps_durgen=a.nsteps_durgen, nsteps_denoiser=a.nsteps_denoiser, on_result=on_result)
    return sum(rtfs) / len(rtfs) if rtfs else None


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
    if device.type == "cuda" and not torch.cuda.is_available():
        raise SystemExit("CUDA is not available: the B200-native hot path has no CPU fallback.")
    enc, dec = get_codec(device, a.codec_encoder_ckpt, a.codec_decoder_ckpt)
    cfg = load_cfg(a.cfg_path)
    model = Flamed.from_pretrained(cfg=cfg, ckpt_path=a.ckpt_path, device=device, weights_only=a.weights_only)
    model.to(device).set_precision(a.precision)
    dec.set_precision(a.precision)
    rtf = run_metadata(model, enc, dec, a) if a.metadata_file else run_prompts(model, enc, dec, a)
    if rtf is not None:
        print("=" * 20, "Avg RTF", "=" * 20)
        print(">" * 5, "RTF:", round(rtf, 3))
    else:
        print("No samples were generated.")
    return rtf


if __name__ == "__main__":
    main()
