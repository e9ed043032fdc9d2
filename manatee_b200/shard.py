"""Record-index sharding of one snapshot stream across the GPUs of a box.

Each rank verifies a contiguous run of whole records.  The only
data-path exchange is an all-gather of the 40-byte shard aggregate
``(n, A, B, C, D)`` that ``mtz_dev_aggregate`` returns; the running checksum that
precedes a shard is the fold of the aggregates of all earlier shards.  Bit 63 of
``n`` marks an aggregate whose segment contains a DRR_BEGIN: the stream checksum
restarts there, so everything before it is dropped.
"""
M64 = (1 << 64) - 1
RESET = 1 << 63


def tri2(n):
    return (n * (n + 1) // 2) & M64


def tri3(n):
    return (n * (n + 1) * (n + 2) // 6) & M64


def apply_aggregate(state, agg):
    """Running checksum ``state`` followed by a segment whose zero-state sums are ``agg``."""
    n, A, B, C, D = agg
    a, b, c, d = state
    if n & RESET:
        a = b = c = d = 0
        n &= RESET - 1
    t2, t3 = tri2(n), tri3(n)
    return ((a + A) & M64,
            (b + n * a + B) & M64,
            (c + n * b + t2 * a + C) & M64,
            (d + n * c + t2 * b + t3 * a + D) & M64)


def carry_before(rank, aggregates):
    """Running checksum at the start of shard ``rank`` given every shard's aggregate."""
    s = (0, 0, 0, 0)
    for r in range(rank):
        s = apply_aggregate(s, aggregates[r])
    return s


def _to_i64(v):
    return [x - (1 << 64) if x >= (1 << 63) else x for x in v]


def _from_i64(v):
    return tuple(int(x) & M64 for x in v)


def all_gather_aggregates(agg, device=None):
    """all-gather of one 5 x u64 aggregate per rank (NCCL on GPU boxes, gloo in CPU tests)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [tuple(agg)]
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(_to_i64(agg), dtype=torch.int64, device=device)
    outs = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [_from_i64(o.cpu().tolist()) for o in outs]


def broadcast_state(state, src, device=None):
    """broadcast a 4 x u64 checksum state from rank ``src`` (stream generation only)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tuple(state)
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(_to_i64(state), dtype=torch.int64, device=device)
    dist.broadcast(t, src=src)
    return _from_i64(t.cpu().tolist())
