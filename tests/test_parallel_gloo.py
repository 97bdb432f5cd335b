"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucket dealing and the final waveform gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flamed_tts_b200 import parallel as P
from flamed_tts_b200 import synthetic as W


def test_bucketing_and_dealing_cover_every_utterance_once():
    wl = W.metadata_workload(300, 64, seed=3)
    lengths = [p.numel() for p in wl["phonemes"]]
    buckets = P.bucket_by_length(lengths, 64)
    assert sorted(i for b in buckets for i in b) == list(range(300))
    assert all(len(b) <= 64 for b in buckets)
    # sorted: a bucket never mixes lengths across another bucket's range
    for a, b in zip(buckets, buckets[1:]):
        assert min(lengths[i] for i in a) >= max(lengths[i] for i in b)
    for world in (1, 2, 4, 8):
        dealt = P.deal_buckets(lengths, buckets, world)
        assert sorted(i for r in dealt for b in r for i in b) == list(range(300))
        loads = [sum(P.bucket_cost(lengths, b) for b in r) for r in dealt]
        if world <= len(buckets):
            assert max(loads) - min(loads) <= max(P.bucket_cost(lengths, b) for b in buckets)
    assert P.deal_buckets(lengths, buckets, 2) == P.deal_buckets(lengths, buckets, 2)  # deterministic


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        wavs = [torch.randn(2 + rank, 1, 400 * (k + 1 + rank), generator=g) for k in range(2 + rank)]
        out = P.gather_waveforms(wavs, rank, world)
        if rank == 0:
            ok = len(out) == world
            for r in range(world):
                gr = torch.Generator().manual_seed(100 + r)
                ref = [torch.randn(2 + r, 1, 400 * (k + 1 + r), generator=gr) for k in range(2 + r)]
                ok = ok and len(out[r]) == len(ref) and all(torch.equal(a, b) for a, b in zip(out[r], ref))
            q.put(ok)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_gather_waveforms_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
