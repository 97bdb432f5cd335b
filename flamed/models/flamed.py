"""`Flamed` facade - the drop-in boundary of the B200 build.

Same constructor, `from_pretrained`, `sample`, `sample_batch`, `_preprocess_*` surface and
504-key state-dict layout as the reference's flamed/models/flamed.py (24-39, 89-217, 219-270);
the bodies are re-written around the B200 engines.  Training entry points are not provided.
"""
import os
import re
import time
from string import punctuation

import numpy as np
import torch
import torch.nn as nn

from flamed.models.synthesizer import PriorGenerator, ProbGenerator
from flamed.text import text_to_sequence

_DEFAULT_LEXICON = os.path.join(os.path.dirname(__file__), "..", "lexicon", "librispeech-lexicon.txt")


class Flamed(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.prior_generator = PriorGenerator(cfg["prior_generator"])
        self.prob_generator = ProbGenerator(cfg["prob_generator"])
        self.lexicon, self.g2p = {}, None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_pretrained(cls, cfg, ckpt_path, device, weights_only=False, training_mode=False):
        cfg["prob_generator"]["device"] = device
        cfg["prior_generator"]["device"] = device
        model = cls(cfg)
        model.lexicon = model.read_lexicon()
        model.g2p = _load_g2p()
        ckpt = torch.load(ckpt_path, map_location=device, weights_only=weights_only)
        model.load_state_dict(ckpt if weights_only else ckpt["state_dict"])
        if training_mode:
            raise NotImplementedError("the B200 build is inference-only (training_mode=True is not supported)")
        return model.eval()

    @property
    def device(self):
        return next(self.parameters()).device

    def set_precision(self, precision):
        """'bf16' (tcgen05 tensor cores, default) or 'fp32' (fp32 FMA parity mode)"""
        self.prior_generator.pva.set_precision(precision)
        self.prob_generator.set_precision(precision)
        return self

    def set_noise_device(self, where):
        """'cpu': the reference's CPU default-generator draws (bit-identical noise for a given
        torch.manual_seed); 'cuda': draw on the device (throughput mode)."""
        self.prior_generator.pva.noise_device = where
        self.prob_generator.noise_device = where
        return self

    def forward(self, *a, **k):
        raise NotImplementedError("training (Flamed.forward / compute_loss) is out of scope of the B200 hot path")

    # ------------------------------------------------------------------ inference
    @torch.inference_mode()
    def sample(self, text=None, phonemes=None, prompt_raw=None, prompt_processed=None, timbre=None, sr=16000,
               codec_cfg=None, codec_encoder=None, codec_decoder=None, temp_durgen=0.3, temp_denoiser=0.3,
               nsteps_durgen=64, nsteps_denoiser=64, lexicon_path=None, cleaners=("english_cleaners",)):
        if codec_encoder is None or codec_decoder is None:
            if codec_cfg is None:
                raise ValueError("codec_encoder / codec_decoder is None: pass them, or pass a codec_cfg to build them.")
            codec_encoder, codec_decoder = self._get_codec_models(codec_cfg)
        if (text is None) == (phonemes is None):
            raise ValueError("`text` and `phonemes` are mutually exclusive: provide exactly one of them.")
        if (prompt_raw is None) == (prompt_processed is None):
            raise ValueError("`prompt_raw` and `prompt_processed` are mutually exclusive: provide exactly one of them.")
        t0 = time.time()
        if text is not None:
            phonemes, _, _ = self._preprocess_english(text, lexicon_path, cleaners)
        else:
            phonemes = phonemes.unsqueeze(0).to(self.device)
        if prompt_raw is not None:
            enc_out = codec_encoder(self._preprocess_acoustic_prompt(prompt_raw, sr))
            _, prompts, _, _, timbre = codec_decoder(enc_out, eval_vq=False, vq=True)
            prompts = prompts.permute(1, 0, 2)
        else:
            if timbre is None:
                raise ValueError("`timbre` must be provided along with `prompt_processed`.")
            timbre = timbre.unsqueeze(0).to(self.device)
            prompts = prompt_processed.unsqueeze(0).to(self.device)
        out = self.sample_batch(
            phonemes=phonemes,
            src_lens=torch.full((phonemes.size(0),), phonemes.size(-1), dtype=torch.long, device=self.device),
            prompts=prompts, timbres=timbre, codec_decoder=codec_decoder, temp_durgen=temp_durgen,
            temp_denoiser=temp_denoiser, nsteps_durgen=nsteps_durgen, nsteps_denoiser=nsteps_denoiser)
        wav = out["wav"][0][0].detach().cpu().numpy()
        return {"wav": wav, "time": time.time() - t0}

    @torch.inference_mode()
    def sample_batch(self, phonemes, src_lens, prompts, timbres, codec_decoder=None, temp_durgen=0.3,
                     temp_denoiser=0.3, nsteps_durgen=64, nsteps_denoiser=64):
        """(B,P) phoneme ids, (B,) lengths, (B,6,Lp) prompt codes padded with vocab_size, (B,256) timbres
        -> dict(prior_embs, prior_logits, tgt_mask, latents, time[, wav]); reference flamed.py:168-217."""
        t0 = time.time()
        dev = self.device
        phonemes, src_lens = phonemes.to(dev), src_lens.to(dev)
        prompts, timbres = prompts.to(dev), timbres.to(dev)
        prior_embs, prior_logits, tgt_mask = self.prior_generator.sample(
            texts=phonemes, src_lens=src_lens, max_src_len=phonemes.size(-1), prompts=prompts,
            prompts_len=prompts.size(-1), nfe=nsteps_durgen, temperature=temp_durgen)
        latents = self.prob_generator.sample(cond=prior_embs, spk=timbres, nfe=nsteps_denoiser,
                                             temperature=temp_denoiser, mask=~tgt_mask.unsqueeze(-1))
        out = {"prior_embs": prior_embs, "prior_logits": prior_logits, "tgt_mask": tgt_mask, "latents": latents,
               "time": time.time() - t0}
        if codec_decoder is not None:
            out["wav"] = codec_decoder.inference(latents, timbres)
        return out

    @torch.inference_mode()
    def sample_batches(self, batches, codec_decoder=None, temp_durgen=0.3, temp_denoiser=0.3, nsteps_durgen=64,
                       nsteps_denoiser=64, on_result=None):
        """Pipelined form of `sample_batch` for a list of length-bucketed batches (the synthesize_via_metadata
        workload): the front stage of batch i+1 (phoneme encoder, duration ODEs, length regulator - small kernels
        and the path's only host synchronisation) runs on a side stream while the denoiser / codec kernels of
        batch i execute, so the device never drains between buckets.  `batches`: iterable of dicts with
        phonemes (B,P), src_lens (B,), prompts (B,6,Lp), timbres (B,256) (host or device tensors).  Random draws happen in
        the same order as a loop of `sample_batch` calls.  Returns the list of `sample_batch` result dicts
        (or calls on_result(i, out) and returns None)."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("Flamed.sample_batches: the hot path runs on a B200 (no CPU/PyTorch fallback)")
        batches = list(batches)
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_side_stream", None) is None or self._side_stream.device != dev:
            self._side_stream = torch.cuda.Stream(dev)
        side = self._side_stream

        side.wait_stream(main)  # inputs the caller prepared on its stream; later fronts must NOT wait for `main`

        def front(b):
            with torch.cuda.stream(side):
                t = {k: b[k].to(dev, non_blocking=True) for k in ("phonemes", "src_lens", "prompts", "timbres")}
                x, tgt_lens = self.prior_generator.front(t["phonemes"], t["src_lens"], t["phonemes"].size(-1),
                                                         nfe=nsteps_durgen, temperature=temp_durgen)
                ev = torch.cuda.Event()
                ev.record(side)
            for v in list(t.values()) + [x, tgt_lens]:
                v.record_stream(main)
            return t, x, tgt_lens, ev

        outs = []
        pending = front(batches[0]) if batches else None
        for i in range(len(batches)):
            t0 = time.time()
            t, x, tgt_lens, ev = pending
            main.wait_event(ev)
            pg = self.prior_generator
            prior_embs, prior_logits, tgt_mask = pg.decode_priors(
                x, tgt_lens, t["prompts"], t["prompts"].size(-1), bf16=pg.pva.precision == "bf16" and x.is_cuda)
            latents = self.prob_generator.sample(cond=prior_embs, spk=t["timbres"], nfe=nsteps_denoiser,
                                                 temperature=temp_denoiser, mask=~tgt_mask.unsqueeze(-1))
            out = {"prior_embs": prior_embs, "prior_logits": prior_logits, "tgt_mask": tgt_mask, "latents": latents,
                   "time": time.time() - t0}
            if codec_decoder is not None:
                out["wav"] = codec_decoder.inference(latents, t["timbres"])
            # everything of batch i is enqueued: the front stage of batch i+1 now overlaps its execution
            pending = front(batches[i + 1]) if i + 1 < len(batches) else None
            if on_result is not None:
                on_result(i, out)
            else:
                outs.append(out)
        return None if on_result is not None else outs

    # ------------------------------------------------------------------ pre-processing
    def _preprocess_acoustic_prompt(self, acoustic_prompt, sr=16000):
        if isinstance(acoustic_prompt, str):
            acoustic_prompt = _load_wav(acoustic_prompt, sr)
        if isinstance(acoustic_prompt, np.ndarray):
            return torch.from_numpy(acoustic_prompt).float().view(1, 1, -1).to(self.device)
        if isinstance(acoustic_prompt, torch.Tensor):
            return acoustic_prompt.to(self.device)
        raise ValueError("Acoustic prompt must be one of [str, np.ndarray, torch.Tensor]!")

    def _get_codec_models(self, codec_cfg):
        from flamed.models.facodec import FACodecDecoder, FACodecEncoder
        return (FACodecEncoder.from_pretrained(codec_cfg["encoder"]).eval(),
                FACodecDecoder.from_pretrained(codec_cfg["decoder"]).eval())

    def read_lexicon(self, lexicon_path=None):
        path = lexicon_path or os.environ.get("FLAMED_LEXICON") or _DEFAULT_LEXICON
        lexicon = {}
        if not os.path.exists(path):  # the blob is not shipped (reference: .MISSING_LARGE_BLOBS)
            return lexicon
        with open(path) as f:
            for line in f:
                word, *phones = re.split(r"\s+", line.strip("\n"))
                lexicon.setdefault(word.lower(), phones)
        return lexicon

    def _preprocess_english(self, text, lexicon_path=None, cleaners="english_cleaners"):
        """text -> (ids (1,P) on device, text, phone string); reference flamed.py:251-270"""
        if lexicon_path:
            self.lexicon = self.read_lexicon(lexicon_path)
        text = text.rstrip(punctuation)
        phones = []
        for w in re.split(r"([,;.\-\?\!\s+])", text):
            if w.lower() in self.lexicon:
                phones += self.lexicon[w.lower()]
            elif self.g2p is not None:
                phones += [p for p in self.g2p(w) if p != " "]
            elif w.strip():
                raise RuntimeError("word %r is not in the lexicon and g2p_en is not installed; pass `phonemes=`" % w)
        phones = "{sp " + " ".join(phones) + "}"
        phones = re.sub(r"\{[^\w\s]?\}", "{sp}", phones).replace("}{", " ")
        seq = torch.as_tensor(text_to_sequence(phones, cleaners), dtype=torch.long)
        return seq.unsqueeze(0).to(self.device), text, phones


def _load_g2p():
    try:
        from g2p_en import G2p
        return G2p()
    except Exception:  # g2p_en is optional: without it only lexicon words / phoneme inputs work
        return None


def _load_wav(path, sr):
    try:
        import soundfile as sf
        wav, file_sr = sf.read(path, dtype="float32", always_2d=False)
    except ImportError:
        from scipy.io import wavfile
        file_sr, wav = wavfile.read(path)
        wav = wav.astype(np.float32) / (32768.0 if wav.dtype != np.float32 else 1.0)
    if wav.ndim > 1:
        wav = wav.mean(axis=1)
    if file_sr != sr:
        from scipy.signal import resample_poly
        g = np.gcd(int(file_sr), int(sr))
        wav = resample_poly(wav, sr // g, file_sr // g).astype(np.float32)
    return wav
