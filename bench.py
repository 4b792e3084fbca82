#!/usr/bin/env python
"""SCG denoise-steps/sec, DiTRotary_XL_8, 4x128x16 latents, N=16 candidates (BASELINE.json config 3).

One "step" = one `ddim_sample` call (DDIM eta=1, timestep_respacing "256", guidance on every step) on a batch of
B=64 samples with the pitch-histogram rule: 1 DiT(B) + DiT(N*B) + VAE decode of N*B*8 tiles + rule scoring + argmax
(reference guided_diffusion/gaussian_diffusion.py:881-976 -> :491-554).  Synthetic seeded latents and weights.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores (oracle port)

Multi-GPU: the (batch x candidate) axis is embarrassingly parallel; every rank runs its own B=64 batch (weak
scaling), NCCL only broadcasts the weights before and gathers the finished latents after the timed region.
`value` = (n_gpus * steps) / max-over-ranks device time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from functools import partial
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "SCG denoise-steps/sec (DiTRotary_XL_8, 4x128x16, N=16)"
UNIT = "steps/s"
B_FULL, N_FULL, RESPACING, SCALE = 64, 16, "256", 1.2465
F_DIT, F_VAE_TILE = 237.4e9, 114.5e9  # algorithmic FLOPs per sample / per 16x16 tile
TARGET = [0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]
GUIDANCE = SimpleNamespace(schedule=False, t_start=750, t_end=0, interval=1, method="scg", step_size=1.0, nn=False)


def workload_config(B, N, world):
    """The `config` object both arms print (the reference arm runs a bounded sample of this workload, see its
    cpu_baseline.sample)."""
    return {"workload": "config 3: DDIM(eta=1) respacing '256', SCG N=%d, pitch_hist, batch %d per GPU, DiTRotary_XL_8 "
                        "random-init (adaLN/final re-randomised), 4x128x16" % (N, B),
            "global_batch": B * world, "candidates": N,
            "parallelism": "batch-sharded x%d, no per-step collective" % world,
            "l2": "no flush: one step streams >100 GB of activations through HBM, far beyond the 126 MB L2"}


def step_flops(B, N, tiles=8):
    return B * ((1 + N) * F_DIT + N * tiles * F_VAE_TILE)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        # under load = the upper half of the samples (idle samples before / after the region are lower)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (a restatement of the reference's algorithm, see oracle/__init__.py)
# ----------------------------------------------------------------------------------------------------------------
def cpu_steps(n_steps, warmup, B=1, N=1):
    """Times `ddim_sample` + SCG of the oracle on the host cores on a bounded sample of the workload (batch B, N
    candidates instead of 64 x 16) and scales to the full step by the algorithmic FLOP ratio."""
    from oracle import dit as odit, sampler as osampler, vae as ovae, weights as ow

    torch.manual_seed(0)
    sd = ow.make_dit_state_dict(seed=0)
    vsd = ow.make_vae_state_dict(seed=1)
    diff = osampler.OracleDiffusion(timestep_respacing=RESPACING)
    fn = lambda x, t, y=None, rule=None: odit.dit_forward(sd, x, t, y, heads=16, patch=8)  # noqa: E731
    kwargs = {"y": torch.ones(B, dtype=torch.long), "rule": {"pitch_hist": torch.tensor([TARGET]).repeat(B, 1)}}
    decode = partial(ovae.decode_latents, vsd, scale_factor=SCALE)
    x = torch.randn(B, 4, 128, 16)
    times = []
    with torch.no_grad():
        for i in range(warmup + n_steps):
            t = torch.full((B,), diff.num_timesteps - 1 - i, dtype=torch.long)
            t0 = time.perf_counter()
            out = diff.ddim_sample(fn, x, t, model_kwargs=kwargs, eta=1.0, decode_fn=decode, guidance_kwargs=GUIDANCE,
                                   scg_kwargs={"num_samples": N, "pitch_hist": 1.0})
            x = out["sample"]
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    scale = step_flops(B_FULL, N_FULL) / step_flops(B, N)
    return 1.0 / (sec * scale), sec, scale


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    value, sec, scale = cpu_steps(args.steps, min(args.warmup, 1), B=1, N=1)
    sample = (f"oracle port of the reference step at B=1, N=1 (1+1 DiT forwards, 8 VAE tiles): {sec:.2f} s per sample "
              f"step on {cores} threads, scaled x{scale:.0f} (algorithmic FLOPs) to B=64, N=16; value counts B=64 batch-steps "
              f"per second like the GPU arm's (the host's rate does not change with --gpus)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * scale * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B_FULL, N_FULL, max(args.gpus, 1)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# this repo
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    from rule_guided_music_b200 import synthetic_weights as ow  # seeded state dicts, reference keys (no checkpoints offline)
    from rule_guided_music_b200 import _lib
    from rule_guided_music_b200.guided_diffusion import dist_util
    from rule_guided_music_b200.guided_diffusion.condition_functions import model_fn
    from rule_guided_music_b200.guided_diffusion.dit import DiT_models
    from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion
    from rule_guided_music_b200.taming.models.klvae_pedal import AutoencoderKL

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    rank, world = dist_util.setup_dist(dev)
    B, N = args.batch, args.candidates

    # weights: rank 0 generates, NCCL broadcasts the packed fp32 state (dist_util.load_state_dict's job, dist_util.py:65-85)
    sd = dist_util.broadcast_state_dict(ow.make_dit_state_dict(seed=0 if rank == 0 else 7), dev, src=0)
    vsd = dist_util.broadcast_state_dict(ow.make_vae_state_dict(seed=1 if rank == 0 else 8), dev, src=0)
    model = DiT_models["DiTRotary_XL_8"](input_size=[128, 16], in_channels=4, num_classes=3, learn_sigma=False)
    model.load_state_dict(sd, strict=False)
    model.to(dev).eval()
    vae = AutoencoderKL(ddconfig=ow.VAE_DDCONFIG, embed_dim=4)
    vae.load_state_dict(vsd, strict=False)
    vae.to(dev).eval()
    del sd, vsd
    diffusion = create_diffusion(timestep_respacing=RESPACING)
    diffusion.enable_cuda_graphs(not args.no_graph)  # whole step = one graph launch (bit-identical to eager)
    fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
    kwargs = {"y": torch.ones(B, dtype=torch.long, device=dev),
              "rule": {"pitch_hist": torch.tensor([TARGET], device=dev).repeat(B, 1)}}
    scg = {"num_samples": N, "pitch_hist": 1.0}
    torch.manual_seed(1234 + rank)
    x = torch.randn(B, 4, 128, 16, device=dev)
    T = diffusion.num_timesteps

    def step(x, i):
        t = torch.full((B,), T - 1 - (i % (T - 1)), device=dev, dtype=torch.long)
        with torch.no_grad():
            return diffusion.ddim_sample(fn, x, t, model_kwargs=kwargs, eta=1.0, embed_model=vae, scale_factor=SCALE,
                                         guidance_kwargs=GUIDANCE, scg_kwargs=scg, _t_host=int(T - 1 - (i % (T - 1))))["sample"]

    def barrier():
        dist_util.barrier()
        torch.cuda.synchronize()

    k = 0
    launches_per_step = 0
    for w in range(args.warmup):
        l_w = _lib.launch_count()
        x = step(x, k)
        if w == 0:  # the first step of a kind always runs eagerly: count the kernels one step launches
            launches_per_step = _lib.launch_count() - l_w
        k += 1
    # ---- timed region: inputs resident in HBM ---------------------------------------------------------------------
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.3)
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        if args.ncu_range and i == 0:  # `ncu --profile-from-start off`: the launch list of exactly one timed step
            torch.cuda.profiler.start()
        x = step(x, k)
        if args.ncu_range and i == 0:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        k += 1
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    graphed = sum(1 for g in diffusion._graphs.values() if g is not False)
    if graphed:  # replayed kernels are not seen by the library's launch counter: kernels per step x steps
        launches = launches_per_step * args.steps
    clk = clocks.stop() if clocks else None
    # ---- end to end: host buffers in, host buffers out, every step --------------------------------------------------
    x_host = torch.empty(B, 4, 128, 16).pin_memory()
    x_host.copy_(x)
    out_host = torch.empty(B, 4, 128, 16).pin_memory()
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        xd = x_host.to(dev, non_blocking=True)
        out_host.copy_(step(xd, k), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller owns the result on the host before the next step
        x_host.copy_(out_host)
        k += 1
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    # ---- per-kernel device times for the roofline (separate pass, events around every launch) ----------------------
    vae.set_lanes(1)  # serial execution so that each launch's event pair times that launch alone
    model.set_lanes(1)
    diffusion.enable_cuda_graphs(False)  # the per-launch event pairs are host-side calls: eager
    _lib.prof_enable(True)
    for _ in range(args.prof_steps):
        x = step(x, k)
        k += 1
    prof = _lib.prof_summary()
    _lib.prof_enable(False)

    ms, ms_e2e = dist_util.max_over_ranks(ms, dev), dist_util.max_over_ranks(ms_e2e, dev)
    # gather the finished latents once (scripts/cfg_sample.py:102-109), outside the timed region
    gathered = dist_util.gather_samples(x)
    finite = bool(torch.isfinite(gathered).all().item()) and gathered.shape[0] == world * B
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md)"
        gemm = {n: v for n, v in prof.items() if n.startswith("gemm_tc")}
        # the launch shape that takes the largest share of the step, with its DRAM traffic from the committed ncu capture
        dominant, traffic = None, None
        if gemm:
            dn, dv = max(gemm.items(), key=lambda kv: kv[1]["ms"])
            d_ach = dv["flops_alg"] / (dv["ms"] * 1e-3) / 1e12 if dv["ms"] > 0 else 0.0
            dominant = {"launch": dn, "launches_per_step": dv["launches"] / max(args.prof_steps, 1),
                        "avg_launch_ms": dv["ms"] / max(dv["launches"], 1), "achieved": d_ach,
                        "frac": d_ach / tensor_peak, "algorithmic_flops_per_launch": dv["flops_alg"] / max(dv["launches"], 1)}
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(dn)
                if tr and B == B_FULL and N == N_FULL:
                    traffic = tr["traffic_bytes"]
                    dominant.update(algorithmic_bytes_per_launch=tr["algorithmic_bytes"], traffic_bytes_per_launch=traffic,
                                    traffic_source="profiles/r1_vae_convs.ncu.txt / r1_dit_linears.ncu.txt (ncu --set full)")
            except Exception:
                pass
        g_ms = sum(v["ms"] for v in gemm.values())
        g_alg = sum(v["flops_alg"] for v in gemm.values())
        g_exec = sum(v["flops_exec"] for v in gemm.values())
        g_n = sum(v["launches"] for v in gemm.values())
        all_ms = sum(v["ms"] for v in prof.values())
        achieved = g_alg / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        value = world * args.steps / (ms * 1e-3)
        e2e_value = world * args.steps / (ms_e2e * 1e-3)
        fl = step_flops(B, N)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate/residual", "data": "synthetic",
            "config": workload_config(B, N, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 4 * 128 * 16 * 4,
                    "d2h_bytes_per_step": B * 4 * 128 * 16 * 4},
            "gpu_launches": int(launches),
            "launch_mode": ("cuda graph replay, %d kernel nodes per step" % launches_per_step) if graphed else "eager",
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": achieved / tensor_peak, "traffic": traffic, "dominant_launch": dominant,
                         "kernel": "gemm_tc_kernel (tcgen05 implicit GEMM: every DiT linear and VAE convolution)",
                         "peak_source": peak_src, "launches_per_step": g_n / max(args.prof_steps, 1),
                         "executed_tflops": g_exec / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0,
                         "share_of_step_device_time": g_ms / all_ms if all_ms > 0 else None},
            "step_tflops_algorithmic": fl * world * args.steps / (ms * 1e-3) / 1e12,
            "finite": finite,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = torch.get_num_threads()
            v, sec, scale = cpu_steps(2, 1, B=1, N=1)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"oracle port at B=1, N=1: {sec:.2f} s per sample step on {cores} threads, scaled x{scale:.0f} "
                          "(algorithmic FLOPs) to B=64, N=16"}
        if args.prof_out:
            with open(args.prof_out, "w") as f:
                json.dump({"per_kernel_family": prof, "prof_steps": args.prof_steps, "ms_per_step_unprofiled":
                           ms / args.steps}, f, indent=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_FULL)
    ap.add_argument("--candidates", type=int, default=N_FULL)
    ap.add_argument("--prof-steps", type=int, default=1)
    ap.add_argument("--prof-out", default="")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-range", action="store_true",
                    help="bracket the first timed step with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
