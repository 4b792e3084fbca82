"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): candidate-sharded SCG -- eager and with whole-step
CUDA graphs (the NCCL all-gather of the exchange step is captured) -- must reproduce the unsharded trajectory bit for
bit on every rank.  Prints one line per rank and a timing of sharded vs unsharded steps at B=4, N=16 (strong scaling
of one small batch, BASELINE.json config 4/5 style).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_shard_check.py
"""
import os, sys, time
from functools import partial
from types import SimpleNamespace
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi, gpu_util
from rule_guided_music_b200.guided_diffusion import dist_util
from rule_guided_music_b200.guided_diffusion.condition_functions import model_fn
from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
rank, world = dist_util.setup_dist(dev)
model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], dev)
vae, _ = gpu_util.native_vae(dev)
B, N = 4, 16
fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
kwargs = {"y": torch.ones(B, dtype=torch.long, device=dev),
          "rule": {"pitch_hist": torch.tensor([[0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]], device=dev).repeat(B, 1)}}
guidance = SimpleNamespace(schedule=False, t_start=750, t_end=0, interval=1, method="scg", step_size=1.0, nn=False)

def run(sharded, graphs, steps="8"):
    dist_util.shard_candidates(sharded)
    diffusion = create_diffusion(timestep_respacing=steps).enable_cuda_graphs(graphs)
    torch.manual_seed(99)
    torch.cuda.synchronize(); dist_util.barrier(); t0 = time.perf_counter()
    out = [o["sample"].clone() for o in diffusion.ddim_sample_loop_progressive(
        fn, (B, 4, 128, 16), model_kwargs=kwargs, device=dev, eta=1.0, embed_model=vae, scale_factor=gi.SCALE_FACTOR,
        guidance_kwargs=guidance, scg_kwargs={"num_samples": N, "pitch_hist": 1.0})]
    torch.cuda.synchronize(); dist_util.barrier()
    return out, (time.perf_counter() - t0) / len(out) * 1e3, sum(1 for g in diffusion._graphs.values() if g is not False)

ref, ms_ref, _ = run(False, False)
ref, ms_ref, _ = run(False, False)
res = {}
for name, (sh, gr) in {"sharded_eager": (True, False), "sharded_graph": (True, True), "unsharded_graph": (False, True)}.items():
    run(sh, gr)
    got, ms, ng = run(sh, gr)
    res[name] = (all(torch.equal(a, b) for a, b in zip(ref, got)), ms, ng)
dist_util.shard_candidates(False)
print(f"rank {rank}/{world}: unsharded eager {ms_ref:.1f} ms/step; " +
      "; ".join(f"{k}: bit-identical={v[0]} {v[1]:.1f} ms/step graphs={v[2]}" for k, v in res.items()), flush=True)
dist_util.barrier()
if world > 1:
    torch.distributed.destroy_process_group()
