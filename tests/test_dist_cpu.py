"""World-size-2 gloo run of the multi-GPU plumbing on CPU: weight broadcast, batch sharding, rank-ordered gather of
finished samples and max-over-ranks timing (the only collectives the sampling path uses)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from rule_guided_music_b200.guided_diffusion import dist_util
    from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion

    r, w = dist_util.setup_dist("cpu")
    assert (r, w) == (rank, world)
    # rank 0 "loaded the checkpoint"; the others start from garbage
    g = torch.Generator().manual_seed(0 if rank == 0 else 99)
    sd = {"b.weight": torch.randn(4, 3, generator=g), "a.bias": torch.randn(5, generator=g)}
    sd = dist_util.broadcast_state_dict(sd, "cpu", src=0)
    lo, hi = dist_util.shard_range(7, rank, world)
    # every rank samples its shard with a pure-torch model through the host sampler (no GPU involved)
    diffusion = create_diffusion(timestep_respacing="3")

    class M:
        def __call__(self, x, t, **kw):
            return x * sd["a.bias"][0]

        def parameters(self):
            yield torch.zeros(1)

    torch.manual_seed(100 + rank)
    out = diffusion.p_sample_loop(M(), (hi - lo, 4, 8, 16), device="cpu", model_kwargs={})
    pad = torch.zeros(4 - (hi - lo), 4, 8, 16)  # all_gather needs equal shapes: pad the short shard
    allx = dist_util.gather_samples(torch.cat([out, pad]))
    tmax = dist_util.max_over_ranks(10.0 + rank, "cpu")
    dist_util.barrier()
    # numpy copies are pickled by value (torch tensors would travel as shared-memory handles that die with the worker)
    q.put((rank, sd["b.weight"].numpy().copy(), (lo, hi), tuple(allx.shape), out.numpy().copy(), allx.numpy().copy(), tmax))


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    w0 = torch.randn(4, 3, generator=g)
    res = [tuple(torch.from_numpy(v) if hasattr(v, "dtype") else v for v in r) for r in res]
    assert torch.equal(res[0][1], w0) and torch.equal(res[1][1], w0)          # broadcast from rank 0
    assert res[0][2] == (0, 4) and res[1][2] == (4, 7)                          # balanced contiguous shards
    assert tuple(res[0][3]) == (8, 4, 8, 16)
    assert torch.equal(res[0][5], res[1][5])                                    # same gathered tensor everywhere
    assert torch.equal(res[0][5][:4], res[0][4]) and torch.equal(res[0][5][4:7], res[1][4])  # rank order
    assert res[0][6] == res[1][6] == 11.0                                       # slowest rank's time


# ---- candidate-sharded SCG: the per-step exchange ---------------------------------------------------------------------
def _scores(N, B):
    """Injected candidate scores with ties inside a shard, across shards, and a sample whose maximum sits in the last
    shard."""
    g = torch.Generator().manual_seed(5)
    s = torch.randn(N, B, generator=g)
    if N >= 8:
        s[1, 0] = s[4, 0] = s[6, 0] = 9.0  # three-way tie across shards: index 1 must win
        s[2, 1] = s[3, 1] = 8.0            # tie inside a shard
    s[N - 1, 2] = 7.0                      # winner in the last shard
    s[:, 3] = 0.5                          # all equal: index 0
    return s


def _exchange_worker(rank, world, port, q, N, B):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from rule_guided_music_b200.guided_diffusion import dist_util

    dist_util.setup_dist("cpu")
    assert dist_util.shard_candidates(True)
    r, w, group = dist_util.candidate_sharding()
    s = _scores(N, B)
    g = torch.Generator().manual_seed(6)
    cand = torch.randn(N, B, 4, 8, 16, generator=g)      # every rank can build all candidates; it only uses its rows
    n0, n1 = dist_util.shard_range(N, r, w)
    if n1 > n0:
        loc = s[n0:n1]
        top = loc.max(0).values
        li = ((loc == top).cumsum(0) == 0).sum(0)            # first maximal local index
        best = loc[li, torch.arange(B)]
        win = cand[n0:n1][li, torch.arange(B)]
    else:
        li = torch.zeros(B, dtype=torch.long)
        best = torch.full((B,), float("-inf"))
        win = torch.zeros(B, 4, 8, 16)
    out, gidx = dist_util.first_max_over_ranks(best, li + n0, win, group=group)
    dist_util.barrier()
    q.put((rank, out.numpy().copy(), gidx.numpy().copy()))


def _run_exchange(world, N, B=4):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q, N, B)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=60) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = _scores(N, B)
    g = torch.Generator().manual_seed(6)
    cand = torch.randn(N, B, 4, 8, 16, generator=g)
    want_idx = torch.from_numpy(s.numpy().argmax(axis=0))    # numpy argmax = first maximal index, like torch on CPU
    want = cand[want_idx, torch.arange(B)]
    for _, out, gidx in res:
        assert torch.equal(torch.from_numpy(gidx), want_idx)  # index work: exact
        assert torch.equal(torch.from_numpy(out), want)       # the winner's bits, untouched
    return want_idx


def test_candidate_exchange_two_ranks():
    idx = _run_exchange(2, N=8)
    assert idx[0] == 1 and idx[1] == 2 and idx[2] == 7 and idx[3] == 0


def test_candidate_exchange_more_ranks_than_candidates():
    """N = 2 candidates over 3 ranks: the last rank has nothing to offer and still takes part in the exchange."""
    _run_exchange(3, N=2)
