"""Data-parallel plumbing of the batched synthesis workload (one process per GPU).

The hot path has no cross-sample reduction, so there is no collective inside either sampling loop: utterances
are length-bucketed, buckets are dealt to ranks by descending cost, every rank runs its buckets independently
and the only exchange is one final gather of the waveforms to rank 0 (NCCL over NVLink on GPUs; the same code
runs over gloo on CPU tensors in the tests).
"""
import torch
import torch.distributed as dist


def bucket_by_length(lengths, max_batch=64):
    """sort by length (descending) and cut into chunks of <= max_batch; returns lists of indices"""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return [order[i:i + max_batch] for i in range(0, len(order), max_batch)]


def bucket_cost(lengths, bucket):
    """padded work of a bucket: batch size x longest member (frames ~ phonemes)"""
    return len(bucket) * max(int(lengths[i]) for i in bucket)


def deal_buckets(lengths, buckets, world):
    """greedy longest-processing-time assignment: buckets by descending cost, each to the least loaded rank.
    Returns a list (per rank) of bucket lists; deterministic, every bucket assigned exactly once."""
    loads = [0] * world
    out = [[] for _ in range(world)]
    for b in sorted(buckets, key=lambda b: (-bucket_cost(lengths, b), b[0])):
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(b)
        loads[r] += bucket_cost(lengths, b)
    return out


def gather_waveforms(wavs, rank, world, dst=0):
    """final gather of this rank's waveforms (list of (B,1,S) tensors) to `dst`.
    Returns on dst a list over ranks of lists of tensors with the original shapes; None elsewhere."""
    if world == 1:
        return [list(wavs)]
    dev = wavs[0].device if wavs else torch.device("cpu")
    flat = torch.cat([w.reshape(-1).float() for w in wavs]) if wavs else torch.zeros(0, device=dev)
    # shapes travel as a small int64 table (n_tensors, then B and S of each), padded to the max count
    meta = torch.tensor([len(wavs)] + [d for w in wavs for d in (w.shape[0], w.shape[-1])], dtype=torch.int64, device=dev)
    sizes = torch.tensor([flat.numel(), meta.numel()], dtype=torch.int64, device=dev)
    dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
    n_flat, n_meta = int(sizes[0]), int(sizes[1])
    fbuf = torch.zeros(n_flat, dtype=torch.float32, device=dev)
    fbuf[: flat.numel()] = flat
    mbuf = torch.zeros(n_meta, dtype=torch.int64, device=dev)
    mbuf[: meta.numel()] = meta
    fl = [torch.empty_like(fbuf) for _ in range(world)] if rank == dst else None
    ml = [torch.empty_like(mbuf) for _ in range(world)] if rank == dst else None
    dist.gather(fbuf, fl, dst=dst)
    dist.gather(mbuf, ml, dst=dst)
    if rank != dst:
        return None
    out = []
    for f, m in zip(fl, ml):
        m = m.tolist()
        items, off = [], 0
        for k in range(m[0]):
            b, s = m[1 + 2 * k], m[2 + 2 * k]
            items.append(f[off: off + b * s].view(b, 1, s))
            off += b * s
        out.append(items)
    return out
