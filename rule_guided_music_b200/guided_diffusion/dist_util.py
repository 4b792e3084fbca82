"""Process-group plumbing for multi-GPU sampling -- the job of guided_diffusion/dist_util.py (:21-53 setup_dist,
:65-85 load_state_dict broadcast, :88 sync_params) and of the sample gathering in scripts/cfg_sample.py:102-109,
without MPI: ranks come from the torchrun environment, NCCL over NVLink on GPUs (gloo on CPU, for the tests).

The sampling path shards the independent (batch x candidate) axis across ranks.  Sharding the BATCH needs nothing
inside a step.  Sharding the CANDIDATES of one batch (B smaller than the number of GPUs: BASELINE.json configs 4/5)
has one exchange per SCG step -- `first_max_over_ranks` below: every rank contributes its best candidate per sample
and all ranks continue from the same global winner.
"""
import os

import torch
import torch.distributed as dist


def setup_dist(device=None, port=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).
    Returns (rank, world_size).  A single process needs no group.  `port` (the reference's only argument,
    dist_util.py:21) is used as MASTER_PORT when the launcher did not set one."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if port is not None:
            os.environ.setdefault("MASTER_PORT", str(port))
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


# ---- candidate sharding (one exchange per SCG step) -----------------------------------------------------------------
_cand_shard = None  # None = off, else the process group (dist.group.WORLD when enabled with group=None)


def shard_candidates(enable=True, group=None):
    """Split the N SCG candidates of every sample across the ranks of `group` (contiguous, balanced, rank order).
    All ranks must hold the same batch and the same torch RNG state: each draws the FULL [N, B, ...] noise tensor
    and keeps its rows, so the sampled trajectory is bit-identical to a single-GPU run with the same seed."""
    global _cand_shard
    if not enable or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        _cand_shard = None
    else:
        _cand_shard = group if group is not None else dist.group.WORLD
    return _cand_shard is not None


def candidate_sharding():
    """(rank, world, group) when candidate sharding is on, else None."""
    if _cand_shard is None:
        return None
    return dist.get_rank(_cand_shard), dist.get_world_size(_cand_shard), _cand_shard


def _all_gather_rows(row, group):
    """[...] on every rank -> [world, ...] on every rank.  gloo cannot gather CUDA tensors: staged through the host
    there (only the single-GPU, two-process parity test does that); NCCL gathers in place over NVLink."""
    world = dist.get_world_size(group)
    if row.is_cuda and dist.get_backend(group) == "gloo":
        parts = [torch.empty(row.shape, dtype=row.dtype) for _ in range(world)]
        dist.all_gather(parts, row.cpu(), group=group)
        return torch.stack(parts).to(row.device)
    out = torch.empty((world * row.shape[0],) + tuple(row.shape[1:]), dtype=row.dtype, device=row.device)
    dist.all_gather_into_tensor(out, row.contiguous(), group=group)  # rank-major concatenation along dim 0
    return out.view((world,) + tuple(row.shape))


def first_max_over_ranks(best, global_idx, winner, group=None):
    """The exchange step of candidate-sharded SCG.  Per sample b this rank offers its best local candidate:
    best [B] fp32 score, global_idx [B] int64 (its index among all N candidates), winner [B, ...] fp32 latent.
    Returns (winner, global_idx) of the rank with the highest score; ties go to the LOWEST rank, which -- shards being
    contiguous in rank order and the local choice being a first-max -- is the first maximal index over all N
    candidates, i.e. exactly torch.argmax over the unsharded scores (gaussian_diffusion.py:539-540).
    One collective: score, index and latent travel as one byte row per sample (32 KB + 12 B at 4x128x16)."""
    B = best.shape[0]
    row = torch.cat([best.reshape(B, 1).float().contiguous().view(torch.uint8),
                     global_idx.reshape(B, 1).to(torch.int64).contiguous().view(torch.uint8),
                     winner.reshape(B, -1).float().contiguous().view(torch.uint8)], dim=1)
    allr = _all_gather_rows(row, group)                                  # [R, B, 12 + 4*elems] bytes
    scores = allr[:, :, 0:4].contiguous().view(torch.float32).squeeze(-1)  # [R, B]
    idxs = allr[:, :, 4:12].contiguous().view(torch.int64).squeeze(-1)     # [R, B]
    top = scores.max(dim=0).values
    is_top = (scores == top) | (scores != scores)  # NaN counts as maximal, like argmax
    first = (is_top.cumsum(0) == 0).sum(0).clamp_(max=scores.shape[0] - 1)  # rows before the first maximal one
    ar = torch.arange(B, device=best.device)
    chosen = allr[first, ar]                                             # [B, bytes]
    out = chosen[:, 12:].contiguous().view(torch.float32).reshape(winner.shape)
    return out, idxs[first, ar]


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
