"""Fan-out of one processed snapshot stream to several bootstrapping peers.

The reference starts one independent `zfs send` per restore request
(lib/backupSender.js:72-73): N peers == N traversals of the same snapshot.  Here the
stream is processed ONCE, sharded by record index over the G GPUs of the box, and every
egress GPU (one per peer connection) obtains the full processed stream by NCCL broadcast
of each shard over NVLink/NVSwitch.  The collective is only used when
more than one peer wants the same snapshot; a single peer needs no exchange beyond the
40-byte aggregates (shard.py).

The egress side consumes shard after shard, so only a rolling receive buffer is kept.
"""


def shard_sizes(nbytes, device=None):
    """all-gather of the byte count of every rank's processed shard"""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(nbytes)]
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(nbytes)], dtype=torch.int64, device=device)
    outs = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [int(o.item()) for o in outs]


def broadcast_shards(local, sizes, consume, recv=None):
    """Every rank streams the whole processed stream in shard order.

    local   : 1-D uint8 tensor, this rank's processed shard (device tensor under NCCL)
    sizes   : per-rank shard sizes (from ``shard_sizes``)
    consume : callable(rank_of_origin, tensor_view) -- the egress hook (socket writer /
              checksum); the view is only valid during the call
    recv    : optional rolling receive buffer (>= max(sizes)); allocated if None
    Returns the number of bytes this rank received from peers.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        consume(0, local[:sizes[0]])
        return 0
    if recv is None:
        recv = torch.empty(max(sizes), dtype=torch.uint8, device=local.device)
    got = 0
    for src in range(world):
        n = sizes[src]
        if src == rank:
            dist.broadcast(local[:n], src=src)
            consume(src, local[:n])
        else:
            dist.broadcast(recv[:n], src=src)
            consume(src, recv[:n])
            got += n
    return got
