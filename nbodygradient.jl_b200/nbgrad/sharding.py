"""Sharding of a batch of independent planetary systems across the GPUs of one box.

Systems never interact: rank r of `world` owns the contiguous slice shard_range(B, r, world) of the
batch, runs it on its own plan/device and writes its own slice of the outputs.  There is no exchange step on the data
path; the only cross-rank traffic is the timing reduction (max over ranks) and an optional gather of small per-rank
summaries, both through torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(nsys, rank, world):
    """Contiguous block partition: the first (nsys % world) ranks get one extra system."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank %r outside world %r" % (rank, world))
    base, extra = divmod(int(nsys), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_counts(nsys, world):
    return [shard_range(nsys, r, world)[1] - shard_range(nsys, r, world)[0] for r in range(world)]


def max_over_ranks(dist, value, device="cpu"):
    """Max of a per-rank scalar (device milliseconds) over all ranks; identity when not distributed."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, values, device="cpu"):
    """Element-wise sum of a small per-rank vector (work counters) over all ranks."""
    v = np.asarray(values, dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return v
    import torch
    t = torch.tensor(v, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def gather_slices(dist, local, nsys, device="cpu"):
    """Host-side gather of per-rank output slices (leading axis = this rank's systems) into the full batch on every rank.
    Only for small summaries (transit counts, chi^2); bulk outputs stay where they were produced."""
    local = np.ascontiguousarray(local)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world = dist.get_world_size()
    counts = shard_counts(nsys, world)
    pad = max(counts)
    buf = np.zeros((pad,) + local.shape[1:], dtype=local.dtype)
    buf[: local.shape[0]] = local
    t = torch.from_numpy(buf).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(outs, counts)], axis=0)
