"""Candidate-sharded SCG on real kernels: two processes (gloo, both on cuda:0 -- the driver's GPU test box has one GPU;
on an 8-GPU box the same code runs over NCCL) each denoise / decode / score their share of the N candidates and
exchange winners once per step.  The trajectory must be BIT-IDENTICAL to the unsharded run with the same seed:
every rank draws the full noise tensor, per-candidate arithmetic does not depend on batch composition, and the
exchange resolves ties like a global first-max argmax."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
TARGET = [0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    try:
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
                          MASTER_PORT=str(port))
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from functools import partial
        from types import SimpleNamespace

        import torch.distributed as dist

        import golden_inputs as gi
        import gpu_util
        from rule_guided_music_b200.guided_diffusion import dist_util
        from rule_guided_music_b200.guided_diffusion.condition_functions import model_fn
        from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion

        cuda = torch.device("cuda:0")
        torch.cuda.set_device(cuda)
        dist.init_process_group("gloo")  # two ranks on one GPU: NCCL refuses that, gloo stages through the host
        model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], cuda)
        vae, _ = gpu_util.native_vae(cuda)
        B, N = 2, 5  # 5 candidates over 2 ranks: 3 + 2
        fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
        kwargs = {"y": torch.ones(B, dtype=torch.long, device=cuda),
                  "rule": {"pitch_hist": torch.tensor([TARGET], device=cuda).repeat(B, 1)}}
        guidance = SimpleNamespace(schedule=False, t_start=750, t_end=0, interval=1, method="scg", step_size=1.0, nn=False)

        def run(sharded):
            dist_util.shard_candidates(sharded)
            diffusion = create_diffusion(timestep_respacing="4")
            diffusion._trace = []
            torch.manual_seed(321)
            steps = [o["sample"].clone() for o in diffusion.ddim_sample_loop_progressive(
                fn, (B, 4, 128, 16), model_kwargs=kwargs, device=cuda, eta=1.0, embed_model=vae,
                scale_factor=gi.SCALE_FACTOR, guidance_kwargs=guidance, scg_kwargs={"num_samples": N, "pitch_hist": 1.0})]
            return steps, [i.cpu() for _, i in diffusion._trace]

        ref, ref_idx = run(False)
        got, got_idx = run(True)
        dist_util.shard_candidates(False)
        same = all(torch.equal(a, b) for a, b in zip(ref, got))
        same_idx = all(torch.equal(a, b) for a, b in zip(ref_idx, got_idx))
        worst = max((a - b).abs().max().item() for a, b in zip(ref, got))
        dist.barrier()
        q.put((rank, same, same_idx, worst, len(got), got[-1].cpu().numpy().copy()))
        dist.destroy_process_group()
    except Exception as e:  # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, False, False, float("nan"), traceback.format_exc(), None))
        raise e


def test_sharded_candidates_bit_identical(cuda):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
    for rank, same, same_idx, worst, n, _ in res:
        assert same_idx, f"rank {rank}: chosen candidate indices differ from the unsharded run ({n})"
        assert same, f"rank {rank}: trajectory differs from the unsharded run, max abs {worst} ({n})"
        assert n == 4
    import numpy as np
    assert np.array_equal(res[0][5], res[1][5])  # both ranks continue from the same x_{t-1}
