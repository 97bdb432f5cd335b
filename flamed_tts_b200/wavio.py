"""Parallel write-out of generated waveforms.

The reference writes one file per utterance with `sf.write(path, wav, 16000)` from the synthesis thread
(/root/reference/synthesize.py:293-298): float32 in, PCM_16 WAV on disk.  At thousands of audio-seconds per second
that single Python thread is the wall, so here the samples leave the device already as int16 PCM
(flm_wav_to_pcm16: lrintf(x * 32767), libsndfile's float -> PCM_16 rule), and a small thread pool writes the
canonical 44-byte-header RIFF/WAVE files (byte-identical to soundfile's / scipy's PCM_16 output); file I/O releases
the GIL, so the writers run beside the thread that keeps the GPU fed.
"""
import os
import struct
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def pcm16_from_float(wav):
    """numpy float waveform in [-1, 1] -> int16, same rule as the device kernel and libsndfile: rint(x * 32767)"""
    return np.clip(np.rint(np.asarray(wav, dtype=np.float32) * np.float32(32767.0)), -32768, 32767).astype(np.int16)


def wav_header(n_samples, sr=16000, channels=1):
    """canonical 44-byte RIFF/WAVE header for 16-bit PCM"""
    data_bytes = n_samples * channels * 2
    return (b"RIFF" + struct.pack("<I", 36 + data_bytes) + b"WAVE" + b"fmt " +
            struct.pack("<IHHIIHH", 16, 1, channels, sr, sr * channels * 2, channels * 2, 16) +
            b"data" + struct.pack("<I", data_bytes))


def write_wav_pcm16(path, pcm, sr=16000):
    pcm = np.ascontiguousarray(pcm, dtype="<i2").reshape(-1)
    tmp = path + ".part"
    with open(tmp, "wb") as f:
        f.write(wav_header(pcm.size, sr))
        f.write(memoryview(pcm).cast("B"))
    os.replace(tmp, path)  # a reader (or --skip-existing on a re-run) never sees a half-written file


class WavWriter:
    """Thread pool of file writers.  submit() returns at once; close() (or leaving the `with` block) waits for every
    file and re-raises the first I/O error."""

    def __init__(self, workers=8, sr=16000, max_pending=256):
        self.sr = sr
        self.pool = ThreadPoolExecutor(max_workers=max(1, int(workers)), thread_name_prefix="wavwriter")
        self.futures = []
        self.slots = threading.BoundedSemaphore(max_pending)  # bounds the host memory held by queued waveforms
        self.files = 0
        self.samples = 0

    def _job(self, path, pcm, ready):
        try:
            if ready is not None:
                ready.synchronize()  # CUDA event of the device -> pinned-host copy that produced `pcm`
            if pcm.dtype != np.int16:
                pcm = pcm16_from_float(pcm)
            write_wav_pcm16(path, pcm, self.sr)
        finally:
            self.slots.release()

    def submit(self, path, pcm, ready=None):
        """pcm: int16 (or float) numpy array / view of a pinned host tensor; ready: optional event to wait on first"""
        self.slots.acquire()
        self.files += 1
        self.samples += int(np.asarray(pcm).size)
        self.futures.append(self.pool.submit(self._job, path, pcm, ready))

    def close(self):
        err = None
        for f in self.futures:
            try:
                f.result()
            except Exception as e:  # noqa: BLE001 - report the first failure after draining the rest
                err = err or e
        self.futures = []
        self.pool.shutdown(wait=True)
        if err is not None:
            raise err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
