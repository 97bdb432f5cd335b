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


def batch_cost_rows(rows, wave_rows=4736, gemm_share=0.66):
    """cost of a back batch of `rows` padded rows in row units: the GEMMs (gemm_share of a step) run 256-row x 256-column
    tiles on 74 CTA pairs, i.e. in waves of 74 tiles = wave_rows rows of a 1024-column output, and a partly filled last
    wave costs a full one; everything else is linear in the rows.  wave_rows = 0: plain padded rows."""
    if not wave_rows:
        return rows
    tiles = -(-(-(-rows // 128)) // 2) * 4           # 128-row tiles paired, 4 N tiles each
    waves = -(-tiles * 64 // wave_rows)              # one tile = 64 rows of the 1024-column output
    return gemm_share * waves * wave_rows + (1.0 - gemm_share) * rows


def bucket_by_rows(lengths, row_budget=32768, max_batch=64, batch_overhead_rows=2500, wave_rows=4736):
    """Sort by length (descending) and cut into contiguous buckets of <= max_batch samples and <= row_budget padded rows
    (samples x longest member) so that  sum over buckets of (batch_cost_rows(padded rows) + batch_overhead_rows)  is
    minimal (dynamic programme over the cut positions, O(n x max_batch)).  Short utterances travel in wide batches and
    long ones in narrow batches - every launch sees about the same number of rows, what the GEMM tiles care about - while
    a bucket spans a narrow range of lengths (little padding), and batch sizes land just under a whole number of GEMM
    waves.  `batch_overhead_rows` is what one more batch costs in units of padded rows (front-loaded tables, codec /
    conditioning launches); 0 minimises the row cost alone, a large value degenerates to the greedy widest-batches cut.
    Returns lists of indices."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    n = len(order)
    lens = [max(1, int(lengths[i])) for i in order]
    inf = float("inf")
    best = [inf] * (n + 1)
    prev = [0] * (n + 1)
    best[0] = 0
    for i in range(n):
        if best[i] == inf:
            continue
        longest = lens[i]
        for k in range(1, max_batch + 1):
            if i + k > n or (k > 1 and k * longest > row_budget):
                break
            c = best[i] + batch_cost_rows(k * longest, wave_rows) + batch_overhead_rows
            if c < best[i + k]:
                best[i + k] = c
                prev[i + k] = i
    cuts, j = [], n
    while j > 0:
        cuts.append((prev[j], j))
        j = prev[j]
    return [order[a:b] for a, b in reversed(cuts)]


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


def gather_waveforms(wavs, rank, world, dst=0, device=None):
    """final gather of this rank's waveforms (list of (B,1,S) tensors of one dtype - fp32, or int16 PCM as stored on
    disk) to `dst` over torch.distributed (gloo on CPU tensors in the tests; the GPU path uses WavGather below).
    Returns on dst a list over ranks of lists of tensors with the original shapes; None elsewhere."""
    if world == 1:
        return [list(wavs)]
    if device is None:
        if wavs:
            device = wavs[0].device
        elif dist.get_backend() == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())
        else:
            device = torch.device("cpu")
    dtype = wavs[0].dtype if wavs else torch.float32
    flat = torch.cat([w.reshape(-1) for w in wavs]) if wavs else torch.zeros(0, device=device, dtype=dtype)
    # shapes travel as a small int64 table (n_tensors, dtype code, then B and S of each), padded to the max count
    code = {torch.float32: 0, torch.int16: 1}[dtype]
    meta = torch.tensor([len(wavs), code] + [d for w in wavs for d in (w.shape[0], w.shape[-1])], dtype=torch.int64,
                        device=device)
    sizes = torch.tensor([flat.numel(), meta.numel(), code if wavs else -1], dtype=torch.int64, device=device)
    dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
    n_flat, n_meta = int(sizes[0]), int(sizes[1])
    if not wavs and int(sizes[2]) == 1:
        dtype = torch.int16
        flat = flat.to(dtype)
    fbuf = torch.zeros(n_flat, dtype=dtype, device=device)
    fbuf[: flat.numel()] = flat
    mbuf = torch.zeros(n_meta, dtype=torch.int64, device=device)
    mbuf[: meta.numel()] = meta
    wire = fbuf.view(torch.uint8) if dtype == torch.int16 else fbuf  # gloo has no int16: PCM travels as bytes
    fl = [torch.empty_like(wire) for _ in range(world)] if rank == dst else None
    ml = [torch.empty_like(mbuf) for _ in range(world)] if rank == dst else None
    dist.gather(wire, fl, dst=dst)
    dist.gather(mbuf, ml, dst=dst)
    if rank != dst:
        return None
    if dtype == torch.int16:
        fl = [f.view(torch.int16) for f in fl]
    out = []
    for f, m in zip(fl, ml):
        m = m.tolist()
        items, off = [], 0
        for k in range(m[0]):
            b, s = m[2 + 2 * k], m[3 + 2 * k]
            items.append(f[off: off + b * s].view(b, 1, s))
            off += b * s
        out.append(items)
    return out


class WavGather:
    """The path's only collective on GPUs: a variable-size gather of int16 PCM waveforms to rank 0 through the C ABI
    (flm_gather_wav: ncclSend / ncclRecv group over NVLink, communicator created from an ncclUniqueId that travels
    over torch.distributed).  Receive buffers on rank 0 are allocated once per size and reused; the exchange is
    enqueued on a side stream after an event, so it overlaps whatever the caller enqueues next on its own stream."""

    def __init__(self, device, rank, world, root=0):
        import ctypes
        from .engines import Context
        self.ctx = Context.get(device)
        self.lib, self.rank, self.world, self.root = self.ctx.lib, rank, world, root
        self.device = self.ctx.device
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == root:
            buf = (ctypes.c_ubyte * 128)()
            _check(self.lib.flm_comm_unique_id(buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(self.device) if dist.get_backend() == "nccl" else uid
        dist.broadcast(uid, src=root)
        raw = (ctypes.c_ubyte * 128)(*uid.cpu().tolist())
        self.handle = ctypes.c_void_p()
        _check(self.lib.flm_comm_create(self.ctx.handle, raw, world, rank, ctypes.byref(self.handle)))
        self.stream = torch.cuda.Stream(self.device)
        self.recv = None
        # sizes travel over a host-side (gloo) group: a device collective here would make the host wait for the whole
        # step before it can enqueue the next one
        self.meta_group = dist.new_group(backend="gloo")
        self.counts = torch.zeros(world, dtype=torch.int64)

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.flm_comm_destroy(self.handle)
            self.handle = None

    def gather(self, pcm_list):
        """pcm_list: this rank's int16 device tensors.  Returns (on root) (flat int16 device buffer, counts list over
        ranks, event recorded when the data has landed); (None, counts, event) elsewhere.  Asynchronous."""
        import ctypes
        flat = torch.cat([w.reshape(-1) for w in pcm_list]) if pcm_list else torch.zeros(0, dtype=torch.int16, device=self.device)
        self.counts.zero_()
        self.counts[self.rank] = flat.numel()
        dist.all_reduce(self.counts, group=self.meta_group)  # 8 bytes per rank on the host: the only size exchange
        counts = self.counts.tolist()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(ready)
        total = int(sum(counts))
        if self.rank == self.root and (self.recv is None or self.recv.numel() < total):
            self.recv = torch.empty(total, dtype=torch.int16, device=self.device)
        carr = (ctypes.c_int64 * self.world)(*[int(c) for c in counts])
        with torch.cuda.stream(self.stream):
            _check(self.lib.flm_gather_wav(self.handle, ctypes.c_void_p(flat.data_ptr()), int(flat.numel()),
                                           ctypes.c_void_p(self.recv.data_ptr()) if self.rank == self.root else None,
                                           carr, self.root, ctypes.c_void_p(self.stream.cuda_stream)))
            flat.record_stream(self.stream)
            done = torch.cuda.Event()
            done.record(self.stream)
        return (self.recv[:total] if self.rank == self.root else None), counts, done


def _check(code):
    from ._lib import check
    check(code)
