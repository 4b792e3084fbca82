"""Process-group plumbing for multi-GPU sampling -- the job of guided_diffusion/dist_util.py (:21-53 setup_dist,
:65-85 load_state_dict broadcast, :88 sync_params) and of the sample gathering in scripts/cfg_sample.py:102-109,
without MPI: ranks come from the torchrun environment, NCCL over NVLink on GPUs (gloo on CPU, for the tests).

The sampling path shards the independent (batch x candidate) axis across ranks; nothing here runs inside a step.
"""
import os

import torch
import torch.distributed as dist


def setup_dist(device=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).
    Returns (rank, world_size).  A single process needs no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if torch.cuda.is_available() and device is not None and torch.device(device).type == "cuda":
            dist.init_process_group("nccl", device_id=torch.device(device))
        else:
            dist.init_process_group("gloo")
    return rank, world


def dev():
    """dist_util.dev(): this rank's device."""
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    return torch.device("cpu")


def broadcast_state_dict(state_dict, device, src=0):
    """Rank `src` owns the weights (it read the checkpoint); every rank returns the same tensors on `device`.
    Keys are walked in sorted order so all ranks issue the same sequence of collectives."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: v.to(device) for k, v in state_dict.items()}
    out = {}
    for k in sorted(state_dict):
        t = state_dict[k].to(device).contiguous()
        dist.broadcast(t, src=src)
        out[k] = t
    return out


def shard_range(total, rank, world):
    """Contiguous, balanced [start, stop) slice of `total` independent items (samples or candidates) for a rank."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_samples(x):
    """all_gather of finished samples in rank order (scripts/cfg_sample.py:102-109); every rank gets [world*B, ...]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, x.contiguous())
    return torch.cat(parts, dim=0)


def max_over_ranks(value, device):
    """Max of a host scalar over ranks (multi-GPU timings are the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
