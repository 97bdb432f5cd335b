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


def test_row_budget_bucketing():
    rng = torch.Generator().manual_seed(0)
    lens = torch.randint(150, 1300, (500,), generator=rng).tolist()
    for budget, maxb in ((16384, 64), (32768, 64), (65536, 128), (10 ** 9, 64)):
        buckets = P.bucket_by_rows(lens, budget, maxb)
        assert sorted(i for b in buckets for i in b) == list(range(500))
        for b in buckets:
            longest = max(lens[i] for i in b)
            assert len(b) <= maxb and (len(b) == 1 or len(b) * longest <= budget)
            assert lens[b[0]] == longest  # sorted inside: the first member is the padded length
        for a, b in zip(buckets, buckets[1:]):
            assert min(lens[i] for i in a) >= max(lens[i] for i in b)
    # padding overhead of exact-length bucketing stays small
    buckets = P.bucket_by_rows(lens, 32768, 64)
    padded = sum(len(b) * max(lens[i] for i in b) for b in buckets)
    assert padded / sum(lens) < 1.10
    assert P.bucket_by_rows([], 100, 4) == [] and P.bucket_by_rows([5000], 100, 4) == [[0]]


def test_row_budget_bucketing_is_optimal_for_its_cost():
    """the cuts minimise sum(padded rows + overhead) over all contiguous partitions of the sorted lengths that respect the
    two limits: checked against brute force on small pools, and the overhead knob moves the batch count monotonically"""
    import itertools
    rng = torch.Generator().manual_seed(3)
    for trial in range(6):
        n = 9 + trial
        lens = torch.randint(5, 60, (n,), generator=rng).tolist()
        budget, maxb, ovh = 150, 4, 17 + 10 * trial
        srt = sorted(lens, reverse=True)

        def cost(cuts):
            c, a = 0, 0
            for b in list(cuts) + [n]:
                k = b - a
                if k > maxb or (k > 1 and k * srt[a] > budget):
                    return None
                c += k * srt[a] + ovh
                a = b
            return c
        best = min(c for r in range(n) for cuts in itertools.combinations(range(1, n), r) if (c := cost(cuts)) is not None)
        got = P.bucket_by_rows(lens, budget, maxb, ovh, wave_rows=0)
        assert sum(len(b) * max(lens[i] for i in b) + ovh for b in got) == best
    lens = torch.randint(150, 1300, (300,), generator=rng).tolist()
    counts = [len(P.bucket_by_rows(lens, 32768, 64, o)) for o in (0, 500, 2500, 10 ** 6)]
    assert counts == sorted(counts, reverse=True) and counts[0] > counts[-1]
    # the wave-aware cost: a partly filled last GEMM wave costs a full one; never below the linear cost's GEMM share
    # (18 row pairs x 4 N tiles = 72 tiles fit one wave of 74, the 19th pair opens a second wave)
    assert P.batch_cost_rows(18 * 256 + 1) > P.batch_cost_rows(18 * 256) + 0.6 * 4736 * 0.9
    assert all(P.batch_cost_rows(r) >= r - 1e-6 for r in range(1, 40000, 97)) and P.batch_cost_rows(777, 0) == 777
    wasted = lambda bs: sum(P.batch_cost_rows(len(b) * max(lens[i] for i in b)) for b in bs)
    assert wasted(P.bucket_by_rows(lens, 32768, 64, 2500)) <= wasted(P.bucket_by_rows(lens, 32768, 64, 2500, wave_rows=0))


def test_shard_of_a_global_pool_is_a_partition():
    """bench.py / synthesize.py: every rank derives its share from the same seeded pool with no communication"""
    for world in (2, 4, 8):
        wl = W.metadata_workload(256 * world, 64, seed=0)  # bench.py's default pool: 256 utterances per GPU
        lens = [p.numel() for p in wl["phonemes"]]
        buckets = P.bucket_by_length(lens, 64)
        shares = [P.deal_buckets(lens, buckets, world)[r] for r in range(world)]
        assert sorted(i for sh in shares for b in sh for i in b) == list(range(256 * world))
        cost = [sum(P.bucket_cost(lens, b) for b in sh) for sh in shares]
        assert max(cost) / (sum(cost) / world) < 1.03  # greedy LPT keeps the straggler within 3 % of the mean


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
        else:
            assert out is None
        # int16 PCM (what the GPU path ships), one rank with nothing to send
        pcm = [] if rank == 1 else [torch.arange(-5, 7, dtype=torch.int16).view(2, 1, 6)]
        out = P.gather_waveforms(pcm, rank, world)
        if rank == 0:
            ok = ok and out[1] == [] and out[0][0].dtype == torch.int16 and torch.equal(out[0][0], pcm[0])
            q.put(ok)
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
