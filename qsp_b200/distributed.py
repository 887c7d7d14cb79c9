"""Multi-GPU layer of the sequential path: independent MPS instances
are split into contiguous blocks, one block per rank / GPU; sites of one chain are strictly
sequential and are never split.  No data-path collective exists; the single collective is
ONE all-gather of the per-instance results (overlaps, unitaries, kind, status, ...) packed into one
byte buffer at the end."""
import torch
import torch.distributed as dist

# instance axis of every result tensor of Engine.seq_prepare (layer-major outputs keep it at dim 1)
BATCH_DIM = {"overlaps": 1, "dis_overlaps": 1, "G": 1, "kind": 1, "d2_svals": 1, "status": 0}


def shard_range(B, rank, world):
    """Contiguous block [b0, b1) of instances owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(B, world)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def _as_bytes(x, bd):
    """instance-major contiguous byte view [n_inst, bytes_per_instance] of a result tensor"""
    if torch.is_complex(x):
        x = torch.view_as_real(x)
    x = x.movedim(bd, 0).contiguous()
    return x.view(torch.uint8).reshape(x.shape[0], -1), x.shape, x.dtype


def gather_results(local, B, group=None):
    """local: dict of this rank's result tensors (instance axis per BATCH_DIM).  Returns the same dict for all B
    instances on every rank.  One collective: every tensor is laid out instance-major, the per-instance byte rows of
    all tensors are concatenated, and a single all_gather of the [max_shard, row_bytes] buffer moves everything."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dict(local)
    keys = [k for k, v in local.items() if v is not None]
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    rank = dist.get_rank(group)
    views = {k: _as_bytes(local[k], BATCH_DIM.get(k, 1)) for k in keys}
    n = sizes[rank]
    for k in keys:
        assert views[k][0].shape[0] == n, (k, views[k][0].shape, n)
    widths = [views[k][0].shape[1] for k in keys]
    row = sum(widths)
    dev = local[keys[0]].device
    buf = torch.zeros((max(sizes), row), dtype=torch.uint8, device=dev)
    o = 0
    for k, w in zip(keys, widths):
        buf[:n, o:o + w] = views[k][0]
        o += w
    out = torch.empty((world, max(sizes), row), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out.view(-1), buf.view(-1), group=group)  # the one collective of the path
    rows = torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)   # [B, row]
    res = {k: None for k in local}
    o = 0
    for k, w in zip(keys, widths):
        _, shp, dt = views[k]
        bd = BATCH_DIM.get(k, 1)
        t = rows[:, o:o + w].contiguous().view(dt).reshape((B,) + tuple(shp[1:])).movedim(0, bd)
        res[k] = torch.view_as_complex(t.contiguous()) if torch.is_complex(local[k]) else t.contiguous()
        o += w
    return res
