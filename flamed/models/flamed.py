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

class LazyOutputs(dict):
    """result dict of sample_batch: plain dict semantics, plus entries that are computed on first access
    (`prior_logits`: the head projection nothing on the sampling path reads, prior_generator.py:179-181)"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._thunks = {}

    def lazy(self, key, fn):
        self._thunks[key] = fn
        super().__setitem__(key, None)  # placeholder keeps the key order / membership of the reference's dict

    def _force(self, key):
        fn = self._thunks.pop(key, None)
        if fn is not None:
            super().__setitem__(key, fn())

    def __getitem__(self, key):
        self._force(key)
        return super().__getitem__(key)

    def get(self, key, default=None):
        if key in self:
            return self[key]
        return default

    def pop(self, key, *default):
        self._force(key)
        return super().pop(key, *default)

    def values(self):
        for k in list(self._thunks):
            self._force(k)
        return super().values()

    def items(self):
        for k in list(self._thunks):
            self._force(k)
        return super().items()


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
        -> dict(prior_embs, prior_logits, tgt_mask, latents, time[, wav]); reference flamed.py:168-217.
        `prior_logits` is evaluated on first access (nothing on the sampling path reads it)."""
        t0 = time.time()
        dev = self.device
        phonemes, src_lens = phonemes.to(dev), src_lens.to(dev)
        prompts, timbres = prompts.to(dev), timbres.to(dev)
        pg = self.prior_generator
        x, tgt_lens = pg.front(phonemes, src_lens, phonemes.size(-1), nfe=nsteps_durgen, temperature=temp_durgen)
        out = self._back(x, tgt_lens, prompts, timbres, codec_decoder, temp_denoiser, nsteps_denoiser)
        # the reference's `time` covers encoder, duration ODEs, the LR host sync and the enqueue of the denoiser loop
        # (flamed.py:181,212): same span here
        out["time"] = time.time() - t0
        return out

    def _back(self, x, tgt_lens, prompts, timbres, codec_decoder, temp_denoiser, nsteps_denoiser):
        """prior decoders -> cond fold -> denoiser loop -> codec on a length-regulated batch x (B,L,192)"""
        pg = self.prior_generator
        bf16 = pg.pva.precision == "bf16" and x.is_cuda
        prior_embs, _, tgt_mask = pg.decode_priors(x, tgt_lens, prompts, prompts.size(-1), bf16=bf16, want_logits=False)
        latents = self.prob_generator.sample(cond=prior_embs, spk=timbres, nfe=nsteps_denoiser,
                                             temperature=temp_denoiser, mask=~tgt_mask.unsqueeze(-1))
        out = LazyOutputs(prior_embs=prior_embs)
        out.lazy("prior_logits", lambda: pg.logits_from(prior_embs, tgt_mask, bf16))
        out.update(tgt_mask=tgt_mask, latents=latents, time=0.0)
        if codec_decoder is not None:
            out["wav"] = codec_decoder.inference(latents, timbres)
        return out

    @torch.inference_mode()
    def sample_batches(self, batches, codec_decoder=None, temp_durgen=0.3, temp_denoiser=0.3, nsteps_durgen=64,
                       nsteps_denoiser=64, on_result=None, rebucket=False, row_budget=32768, max_batch=64, batch_overhead_rows=2500, wave_rows=4736,
                       wav_to_host=None):
        """Batched metadata entry point (the synthesize_via_metadata workload): `batches` is an iterable of dicts with
        phonemes (B,P), src_lens (B,), prompts (B,6,Lp), timbres (B,256) (host or device tensors).

        rebucket=False: a pipelined loop of `sample_batch` calls - the front stage of batch i+1 (phoneme encoder,
        duration ODEs, length regulator: small kernels and the path's only host synchronisation) runs on a side stream
        while the denoiser / codec kernels of batch i execute.  Random draws happen in the order of a loop of
        `sample_batch` calls; results are identical to that loop.

        rebucket=True: the front stage of EVERY batch runs first without any host synchronisation (durations are
        planned on the device, PVA.sample_plan), the frame counts come back in one copy, and the utterances are then
        re-grouped by their real length into batches of <= max_batch samples and <= row_budget padded rows, cut so that
        padded rows + batch_overhead_rows per batch is minimal
        (flamed_tts_b200.parallel.bucket_by_rows) for the back stage.  An utterance's front result is that of the
        reference on its front batch, its back result that of the reference on its (re-padded) back batch; padding
        waste drops from ~20 % to a few %.  Draw order: all duration draws (front batches in order), then one latent
        draw per back batch.  Each result carries `index`: [(front batch, row)] of its samples and `tgt_lens`.

        wav_to_host: None | "f32" | "pcm16": also copy the waveform to pinned host memory on a copy stream
        (`wav_host`, and the CUDA event `wav_ready` to wait on) - "pcm16" converts on the device to what
        soundfile stores for PCM_16 (lrintf(x * 32767)), halving the bytes.

        Results are delivered one batch late (batch i after batch i+1 is enqueued) so that a consumer which
        synchronises on batch i never drains the device.  `time` = device time of the batch (CUDA events; front +
        back, the front share pro rata for re-bucketed batches).  Returns the list of result dicts, or calls
        on_result(i, out) and returns None."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("Flamed.sample_batches: the hot path runs on a B200 (no CPU/PyTorch fallback)")
        batches = list(batches)
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_side_stream", None) is None or self._side_stream.device != dev:
            self._side_stream = torch.cuda.Stream(dev)
            self._copy_stream = torch.cuda.Stream(dev)
        kw = dict(codec_decoder=codec_decoder, temp_denoiser=temp_denoiser, nsteps_denoiser=nsteps_denoiser)
        outs, waiting = [], []

        def finish(i, out, ev0, ev1, extra_ms=0.0):
            """everything of batch i is enqueued on `main` between events ev0 / ev1"""
            if wav_to_host is not None and "wav" in out:
                self._copy_stream.wait_event(ev1)
                with torch.cuda.stream(self._copy_stream):
                    w = out["wav"]
                    if wav_to_host == "pcm16":
                        from flamed_tts_b200.engines import Context, wav_to_pcm16
                        w = wav_to_pcm16(Context.get(dev), w)
                    host = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
                    host.copy_(w, non_blocking=True)
                    w.record_stream(self._copy_stream)
                    out["wav"].record_stream(self._copy_stream)
                    out["wav_host"], out["wav_ready"] = host, torch.cuda.Event()
                    out["wav_ready"].record(self._copy_stream)
            waiting.append((i, out, ev0, ev1, extra_ms))
            if len(waiting) > 1:
                deliver(*waiting.pop(0))

        def deliver(i, out, ev0, ev1, extra_ms):
            ev1.synchronize()
            out["time"] = (ev0.elapsed_time(ev1) + extra_ms) / 1000.0
            if on_result is not None:
                on_result(i, out)
            else:
                outs.append(out)

        def events():
            return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        if rebucket:
            self._sample_rebucketed(batches, main, finish, events, kw, temp_durgen, nsteps_durgen, row_budget, max_batch,
                                    batch_overhead_rows, wave_rows)
        else:
            side = self._side_stream
            side.wait_stream(main)  # inputs the caller prepared on its stream; later fronts must NOT wait for `main`

            def front(b):
                with torch.cuda.stream(side):
                    ev0, ev1 = events()
                    ev0.record(side)
                    t = {k: b[k].to(dev, non_blocking=True) for k in ("phonemes", "src_lens", "prompts", "timbres")}
                    x, tgt_lens = self.prior_generator.front(t["phonemes"], t["src_lens"], t["phonemes"].size(-1),
                                                             nfe=nsteps_durgen, temperature=temp_durgen)
                    ev = torch.cuda.Event()
                    ev.record(side)
                for v in list(t.values()) + [x, tgt_lens]:
                    v.record_stream(main)
                return t, x, tgt_lens, ev, ev0, ev1

            pending = front(batches[0]) if batches else None
            for i in range(len(batches)):
                t, x, tgt_lens, ev, ev0, ev1 = pending
                main.wait_event(ev)
                out = self._back(x, tgt_lens, t["prompts"], t["timbres"], **kw)
                ev1.record(main)
                # everything of batch i is enqueued: the front stage of batch i+1 now overlaps its execution
                pending = front(batches[i + 1]) if i + 1 < len(batches) else None
                finish(i, out, ev0, ev1)
        while waiting:
            deliver(*waiting.pop(0))
        return None if on_result is not None else outs

    def _sample_rebucketed(self, batches, main, finish, events, kw, temp_durgen, nsteps_durgen, row_budget, max_batch,
                           batch_overhead_rows=2500, wave_rows=4736):
        from flamed_tts_b200.parallel import bucket_by_rows
        dev, pg = self.device, self.prior_generator
        pad_code = pg.config["codec"]["vocab_size"]
        f0, f1 = events()
        f0.record(main)
        fronts = []
        for b in batches:  # phase 1: every front batch, back to back, no host synchronisation
            t = {k: b[k].to(dev, non_blocking=True) for k in ("phonemes", "src_lens", "prompts", "timbres")}
            enc, cumsum, tgt_lens = pg.front_plan(t["phonemes"], t["src_lens"], t["phonemes"].size(-1),
                                                  nfe=nsteps_durgen, temperature=temp_durgen)
            fronts.append((t, enc, cumsum, tgt_lens))
        f1.record(main)
        if not fronts:
            return
        lens = torch.cat([f[3] for f in fronts]).cpu().tolist()  # the path's one host synchronisation
        front_ms = f0.elapsed_time(f1)
        owner = [(fi, r) for fi, f in enumerate(fronts) for r in range(f[3].numel())]
        total = max(1, sum(lens))
        eng = pg.pva.engine()
        for bi, idx in enumerate(bucket_by_rows(lens, row_budget, max_batch, batch_overhead_rows, wave_rows)):  # phase 2: pure enqueue
            ev0, ev1 = events()
            ev0.record(main)
            src = [owner[j] for j in idx]
            tl = [int(lens[j]) for j in idx]
            x = eng.expand_gather([(fronts[fi][1][r], fronts[fi][2][r]) for fi, r in src], max(tl))
            tgt_lens = torch.tensor(tl, dtype=torch.long).to(dev, non_blocking=True)
            pr = [fronts[fi][0]["prompts"][r] for fi, r in src]
            lp = max(p.shape[-1] for p in pr)
            if any(p.shape[-1] != lp for p in pr):
                pr = [torch.nn.functional.pad(p, (0, lp - p.shape[-1]), value=pad_code) for p in pr]
            prompts = torch.stack(pr)
            timbres = torch.stack([fronts[fi][0]["timbres"][r] for fi, r in src])
            out = self._back(x, tgt_lens, prompts, timbres, **kw)
            ev1.record(main)
            out["index"], out["tgt_lens"] = src, tl
            finish(bi, out, ev0, ev1, extra_ms=front_ms * sum(tl) / total)

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
