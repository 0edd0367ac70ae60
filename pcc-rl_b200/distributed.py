"""Multi-GPU plumbing: the env batch shards trivially (envs are independent), one process per
GPU, no collective on the step path.  The only exchange is a tiny all-gather of per-rank
episode-return statistics (the reference prints reward_sum / reward_ewma per episode,
gym/network_sim.py:480-483) over torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
import os

import torch
import torch.distributed as dist


def shard_range(n_global, rank, world_size):
    """Contiguous env-index range [lo, hi) owned by `rank`."""
    lo = n_global * rank // world_size
    hi = n_global * (rank + 1) // world_size
    return lo, hi


def shard_flow_batch(batch, rank, world_size, n_flows_global):
    """The records of a global MI batch (numpy arrays in the structure-of-arrays + CSR layout of
    PccFlowMonitor.make_batch, `flow` = global flow id) that belong to `rank`: flows shard by contiguous id range like
    envs do, a flow's records stay on one rank in batch order, flow ids are rebased to the rank's monitor.  No
    collective is involved: every rank can cut its own slice from the global batch, or a front end can route."""
    import numpy as np
    lo, hi = shard_range(n_flows_global, rank, world_size)
    flow = np.asarray(batch["flow"])
    sel = np.nonzero((flow >= lo) & (flow < hi))[0]
    off = np.asarray(batch["rtt_off"], dtype=np.int64)
    n = (off[1:] - off[:-1])[sel]
    new_off = np.zeros(len(sel) + 1, dtype=np.int64)
    new_off[1:] = np.cumsum(n)
    rtt = np.asarray(batch["rtt"])
    parts = [rtt[off[r]:off[r + 1]] for r in sel]
    out = {k: np.asarray(v)[sel] for k, v in batch.items() if k not in ("rtt", "rtt_off", "flow")}
    out["flow"] = (flow[sel] - lo).astype(np.int32)
    out["rtt_off"] = new_off
    out["rtt"] = np.concatenate(parts) if parts else np.zeros(0)
    return out, (lo, hi)


def init_from_env(backend=None):
    """Initialises torch.distributed from torchrun's environment; returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def gather_episode_returns(finished_returns, group=None):
    """finished_returns: 1-D float64 tensor with the returns of the episodes this rank finished
    since the last call (any length, may be empty).  All-gathers [count, sum, sum of squares]
    per rank and returns dict(count, mean, std, per_rank=[(count, mean), ...]) on every rank."""
    x = finished_returns.to(torch.float64)
    stats = torch.stack([torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device),
                         x.sum(), (x * x).sum()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        out = [torch.zeros_like(stats) for _ in range(dist.get_world_size(group))]
        dist.all_gather(out, stats, group=group)
    else:
        out = [stats]
    allst = torch.stack(out).cpu()
    cnt = float(allst[:, 0].sum())
    s1 = float(allst[:, 1].sum())
    s2 = float(allst[:, 2].sum())
    mean = s1 / cnt if cnt > 0 else 0.0
    var = max(s2 / cnt - mean * mean, 0.0) if cnt > 0 else 0.0
    per_rank = [(int(r[0]), float(r[1] / r[0]) if r[0] > 0 else 0.0) for r in allst]
    return dict(count=int(cnt), mean=mean, std=var ** 0.5, per_rank=per_rank)


def max_over_ranks(value, device):
    """Max of a python float over ranks (used for device-time reporting)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
