"""Ensemble sharding across GPUs: trajectories / noise realisations / architecture
fits are independent, so the batch axis is partitioned contiguously and the data path needs NO
collective.  Only single-model training over sharded batches exchanges anything: ONE flat
``all_reduce(SUM)`` of the MLP gradient (+ the loss scalar) per optimiser step."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous balanced partition of ``range(n_items)``: the first ``n_items % world_size``
    ranks get one extra item."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, world_size=None, rank=None, dim=0):
    """Slice every tensor of ``tensors`` (dict or sequence; ``None`` entries pass through) to this
    rank's share of the batch axis."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank

    def cut(x):
        if x is None:
            return None
        lo, hi = shard_bounds(x.shape[dim], world_size, rank)
        return x.narrow(dim, lo, hi - lo)

    if isinstance(tensors, dict):
        return {k: cut(v) for k, v in tensors.items()}
    return [cut(v) for v in tensors]


def lpt_assign(costs, world_size):
    """Longest-processing-time-first assignment of independent fits (e.g. the s00..s11 sweep,
    cost ~ MACs x steps) to ranks; returns a list of item indices per rank."""
    loads = [0.0] * world_size
    out = [[] for _ in range(world_size)]
    for i in sorted(range(len(costs)), key=lambda k: -costs[k]):
        r = min(range(world_size), key=lambda k: loads[k])
        out[r].append(i)
        loads[r] += costs[i]
    return out


def allreduce_flat(flat, loss=None, group=None):
    """Sum a flat gradient tensor (+ the loss scalar, carried as one extra element) over the ranks
    with ONE collective; returns (flat, loss).  No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat, loss
    buf = flat if loss is None else torch.cat([flat.reshape(-1), loss.reshape(1).to(flat)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    if loss is None:
        return buf, None
    return buf[:-1].view_as(flat), buf[-1]


def allreduce_gradients(grads, loss=None, group=None):
    """Sum the per-rank gradients (list of tensors shaped like the parameters) and the loss scalar
    with ONE collective on one flat buffer.  Returns (grads, loss) holding the global sums."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads, loss
    flat = torch.cat([g.reshape(-1) for g in grads] +
                     ([loss.reshape(1).to(grads[0])] if loss is not None else []))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, o = [], 0
    for g in grads:
        out.append(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
    return out, (flat[o] if loss is not None else None)
