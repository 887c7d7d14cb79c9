"""Multi-GPU layer of the sequential path: independent MPS instances
are split into contiguous blocks, one block per rank / GPU; sites of one chain are strictly
sequential and are never split.  No data-path collective exists; the single collective is
the all-gather of the per-instance results (overlaps, unitaries, status) at the end."""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Contiguous block [b0, b1) of instances owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(B, world)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def _all_gather_batch(x, batch_dim, sizes, group=None):
    """all_gather of tensors whose `batch_dim` extent differs per rank (sizes[r])."""
    world = len(sizes)
    mx = max(sizes)
    pad_shape = list(x.shape)
    pad_shape[batch_dim] = mx
    buf = torch.zeros(pad_shape, dtype=x.dtype, device=x.device)
    buf.narrow(batch_dim, 0, x.shape[batch_dim]).copy_(x)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o.narrow(batch_dim, 0, n) for o, n in zip(out, sizes)], dim=batch_dim)


def gather_results(local, B, group=None):
    """local: dict of this rank's result tensors with the instance axis at dim 1 for the
    layer-major ones (overlaps, dis_overlaps, G, kind, d2_svals) and dim 0 for status.
    Returns the same dict for all B instances on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dict(local)
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    out = {}
    for k, v in local.items():
        if v is None:
            out[k] = None
            continue
        bd = 0 if k == "status" else 1
        if torch.is_complex(v):  # gloo has no complex all_gather; view as real pairs
            g = _all_gather_batch(torch.view_as_real(v).contiguous(), bd, sizes, group)
            out[k] = torch.view_as_complex(g.contiguous())
        else:
            out[k] = _all_gather_batch(v.contiguous(), bd, sizes, group)
    return out
