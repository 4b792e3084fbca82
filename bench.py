#!/usr/bin/env python
"""SCG denoise-steps/sec, DiTRotary_XL_8, 4x128x16 latents (BASELINE.json).

Default workload = BASELINE config 3: one "step" = one `ddim_sample` call (DDIM eta=1, timestep_respacing "256",
guidance on every step) on a batch of B=64 samples with N=16 candidates and the pitch-histogram rule: 1 DiT(B) +
DiT(N*B) + VAE decode of N*B*8 tiles + rule scoring + argmax (reference guided_diffusion/gaussian_diffusion.py:881-976
-> :491-554).  Synthetic seeded latents and weights.

    python bench.py --gpus N --steps K --warmup W              # this repo (CUDA, one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W      # the reference algorithm on the host CPU cores
    python bench.py --impl reference-gpu                       # the same pure-PyTorch fp32 path, eager, on cuda:0
    python bench.py --config c2|c3|c5                          # BASELINE configs 2 / 3 / 5 (one JSON line each)
    python bench.py --gpus N --scaling strong                  # global batch 64 split over the N ranks
    python bench.py --gpus N --shard candidates --batch 8 --candidates 64   # all ranks share the batch, N split

Multi-GPU: the (batch x candidate) axis is embarrassingly parallel.  Default (`--scaling weak`): every rank runs its own
B=64 batch, NCCL only broadcasts the weights before and gathers the finished latents after the timed region;
`value` = (n_gpus * steps) / max-over-ranks device time.  The default run also measures, after the headline, short
extra legs that DO exercise the design's per-step behaviour and reports them under `extra`: at N > 1 the global batch
of 64 strong-scaled over the ranks, and at every N a small batch (B=8, N=64: BASELINE config 4's shape) with the
CANDIDATES sharded over the ranks -- one all-gather per step (`ncclAllGather` of [R, B, 32 780 B] rows) inside the
captured step, with the exchange's own device time.
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
# algorithmic FLOPs per sample (T = 256 / 128 tokens) and per 16x16 VAE tile
F_DIT, F_DIT_HALF, F_VAE_TILE = 237.4e9, 116.8e9, 114.5e9
TARGET = [0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]
GUIDANCE = SimpleNamespace(schedule=False, t_start=750, t_end=0, interval=1, method="scg", step_size=1.0, nn=False)
C5_NUM_IMG = 15  # W = 64 * (num_img + 1) = 1024 latent columns = 81.92 s (SURVEY.md section 8a note on config 5)

CONFIG_DEFAULTS = {"c2": (256, 0), "c3": (B_FULL, N_FULL), "c5": (1, N_FULL)}  # (batch per GPU, candidates)


def metric_name(config, N):
    if config == "c2":
        return "denoise-steps/sec (DiTRotary_XL_8, 4x128x16, unguided p_sample, class_cond)"
    if config == "c5":
        return "SCG denoise-steps/sec (DiTRotary_XL_8, diff_collage condind_long 4x1024x16, N=%d)" % N
    return METRIC if N == N_FULL else METRIC.replace("N=16", "N=%d" % N)


def workload_config(B, N, world, config="c3", scaling="weak", shard="batch"):
    """The `config` object both arms print (the reference arm runs a bounded sample of this workload, see its
    cpu_baseline.sample)."""
    if config == "c2":
        what = ("config 2: p_sample (1000-step schedule), class_cond, no guidance, batch %d per GPU, DiTRotary_XL_8 "
                "random-init (adaLN/final re-randomised), 4x128x16" % B)
    elif config == "c5":
        what = ("config 5: diff_collage CondIndSimple num_img=%d overlap 64 (latent 4x1024x16 = 81.92 s), DDIM(eta=1) "
                "respacing '256', SCG N=%d, pitch_hist, batch %d per GPU, DiTRotary_XL_8 random-init" % (C5_NUM_IMG, N, B))
    else:
        what = ("config 3: DDIM(eta=1) respacing '256', SCG N=%d, pitch_hist, batch %d per GPU, DiTRotary_XL_8 "
                "random-init (adaLN/final re-randomised), 4x128x16" % (N, B))
    if shard == "candidates":
        par = "candidate-sharded x%d: every rank holds the batch, N split, one all-gather per step" % world
        gb = B
    elif scaling == "strong":
        par = "batch-sharded x%d (strong: the global batch is split), no per-step collective" % world
        gb = B * world
    else:
        par = "batch-sharded x%d, no per-step collective" % world
        gb = B * world
    return {"workload": what, "global_batch": gb, "candidates": N, "parallelism": par,
            "l2": "no flush: one step streams >100 GB of activations through HBM, far beyond the 126 MB L2"}


def step_flops(B, N, config="c3"):
    """Algorithmic FLOPs of one step on a batch of B."""
    if config == "c2":
        return B * F_DIT
    if config == "c5":
        per_eval = C5_NUM_IMG * F_DIT + C5_NUM_IMG * F_DIT_HALF  # the reference evaluates (and zeroes) the last half tile
        return B * ((1 + N) * per_eval + N * 64 * F_VAE_TILE)
    return B * ((1 + N) * F_DIT + N * 8 * F_VAE_TILE)


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
# reference arms: the reference's own algorithm on the host cores (or, --impl reference-gpu, on cuda:0 in eager fp32)
# ----------------------------------------------------------------------------------------------------------------
def host_threads():
    """Use every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time one thread."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _unmodified_reference(device):
    """The reference itself (RGM_REFERENCE or /root/reference) behind tests/golden/ref_shims, when it is reachable --
    it is in the build container, never on the GPU box.  Returns step-building pieces or None."""
    path = os.environ.get("RGM_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(path, "guided_diffusion")):
        return None
    try:
        for p in (path, os.path.join(ROOT, "tests", "golden", "ref_shims"), os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import guided_diffusion.dit as rdit
        from guided_diffusion.condition_functions import model_fn as r_model_fn
        from guided_diffusion.script_util import create_diffusion as r_create
        from taming.modules.diffusionmodules.model import Decoder
    except Exception:
        return None
    from oracle import weights as ow

    model = rdit.DiT_models["DiTRotary_XL_8"](input_size=[128, 16], in_channels=4, num_classes=3, learn_sigma=False)
    model.load_state_dict(ow.make_dit_state_dict(seed=0), strict=True)
    model.to(device).eval()
    vsd = ow.make_vae_state_dict(seed=1)
    dec, pq = Decoder(**ow.VAE_DDCONFIG), torch.nn.Conv2d(4, 4, 1)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in vsd.items() if k.startswith("decoder.")}, strict=True)
    pq.load_state_dict({k[len("post_quant_conv."):]: v for k, v in vsd.items() if k.startswith("post_quant_conv.")})
    dec.to(device).eval()
    pq.to(device).eval()

    class Embed:  # AutoencoderKL.decode (klvae_pedal.py:80-85)
        @staticmethod
        def decode(z):
            return dec(pq(z))

    diffusion = r_create(learn_sigma=False, diffusion_steps=1000, noise_schedule="linear", timestep_respacing=RESPACING,
                         use_kl=False, predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False)
    diffusion.t_end = 0  # the loops set it (gaussian_diffusion.py:777); a single ddim_sample call reads it (:950)
    fn = partial(r_model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)

    def step(x, t, kwargs, N):
        return diffusion.ddim_sample(fn, x, t, model_kwargs=kwargs, eta=1.0, embed_model=Embed, scale_factor=SCALE,
                                     guidance_kwargs=GUIDANCE, scg_kwargs={"num_samples": N, "pitch_hist": 1.0})["sample"]

    return step, diffusion.num_timesteps


def _oracle_port(device):
    from oracle import dit as odit, sampler as osampler, vae as ovae, weights as ow

    sd = {k: v.to(device) for k, v in ow.make_dit_state_dict(seed=0).items()}
    vsd = {k: v.to(device) for k, v in ow.make_vae_state_dict(seed=1).items()}
    diff = osampler.OracleDiffusion(timestep_respacing=RESPACING, randn=lambda shape: torch.randn(*shape, device=device))
    fn = lambda x, t, y=None, rule=None: odit.dit_forward(sd, x, t, y, heads=16, patch=8)  # noqa: E731
    decode = partial(ovae.decode_latents, vsd, scale_factor=SCALE)

    def step(x, t, kwargs, N):
        return diff.ddim_sample(fn, x, t, model_kwargs=kwargs, eta=1.0, decode_fn=decode, guidance_kwargs=GUIDANCE,
                                scg_kwargs={"num_samples": N, "pitch_hist": 1.0})["sample"]

    return step, diff.num_timesteps


def reference_steps(n_steps, warmup, B, N, device="cpu"):
    """Times `ddim_sample` + SCG of the reference algorithm on a bounded sample of the workload (batch B, N candidates)
    and scales to the full step by the algorithmic FLOP ratio.  Returns (steps/s of the FULL step, seconds per sampled
    step, scale, kind)."""
    device = torch.device(device)
    torch.manual_seed(0)
    ref = _unmodified_reference(device) if os.environ.get("RGM_BENCH_PORT", "") != "1" else None
    kind = "reference" if ref is not None else "port"
    step, T = ref if ref is not None else _oracle_port(device)
    kwargs = {"y": torch.ones(B, dtype=torch.long, device=device),
              "rule": {"pitch_hist": torch.tensor([TARGET], device=device).repeat(B, 1)}}
    x = torch.randn(B, 4, 128, 16, device=device)
    times = []
    with torch.no_grad():
        for i in range(warmup + n_steps):
            t = torch.full((B,), T - 1 - i, dtype=torch.long, device=device)
            if device.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            x = step(x, t, kwargs, N)
            if device.type == "cuda":
                torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    scale = step_flops(B_FULL, N_FULL) / step_flops(B, N)
    return 1.0 / (sec * scale), sec, scale, kind


def cpu_sample_shape(cores):
    """B=1, N=16 (about 10 s per step on 32 threads); on a small host fewer candidates keep
    the run within minutes."""
    return (1, N_FULL) if cores >= 16 else (1, 4)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gpu = args.impl == "reference-gpu"
    cores = host_threads()
    if gpu:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl reference-gpu: no CUDA device")
        (B, N), dev = (4, N_FULL), "cuda:0"  # README.md:36 usage: micro-batches of 4
    else:
        (B, N), dev = cpu_sample_shape(cores), "cpu"
    if args.ref_batch > 0:
        B = args.ref_batch
    if args.ref_candidates > 0:
        N = args.ref_candidates
    value, sec, scale, kind = reference_steps(args.steps, args.warmup, B, N, device=dev)
    where = "cuda:0, eager fp32 PyTorch (TF32 only where PyTorch defaults to it: cuDNN convolutions)" if gpu else \
        f"{cores} host threads"
    what = "the unmodified reference (RGM_REFERENCE)" if kind == "reference" else "oracle port of the reference step"
    sample = (f"{what} at B={B}, N={N} ({B} + {B * N} DiT forwards, {B * N * 8} VAE tiles): {sec:.2f} s per sampled "
              f"step on {where}, scaled x{scale:.1f} (algorithmic FLOPs) to B=64, N=16; value counts B=64 batch-steps "
              f"per second like the GPU arm's (the host's rate does not change with --gpus)")
    # ms_per_step is the MEASURED time of one timed (sampled) step, so steps x ms_per_step is this run's timed region;
    # `value` is that rate scaled to the full B=64, N=16 step (ms_per_full_step_scaled = 1000 / value)
    line = {"impl": args.impl, "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "ms_per_full_step_scaled": sec * scale * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B_FULL, N_FULL, max(args.gpus, 1)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind + ("-gpu" if gpu else ""),
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# this repo
# ----------------------------------------------------------------------------------------------------------------
class Workload:
    """One configuration's step function on this rank's device."""

    def __init__(self, config, B, N, dev, model, vae, graphs=True):
        from rule_guided_music_b200 import diff_collage as dc
        from rule_guided_music_b200.guided_diffusion.condition_functions import dc_model_fn, model_fn
        from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion

        self.config, self.B, self.N, self.dev, self.vae = config, B, N, dev, vae
        self.diffusion = create_diffusion(timestep_respacing="" if config == "c2" else RESPACING)
        self.diffusion.enable_cuda_graphs(graphs)  # whole step = one graph launch (bit-identical to eager)
        self.T = self.diffusion.num_timesteps
        if config == "c5":
            def eps_fn(x, t, y=None):  # scripts/sample_rule.py:120-122
                return model(x.permute(0, 1, 3, 2), t, y=y).permute(0, 1, 3, 2)
            worker = dc.CondIndSimple((4, 16, 128), eps_fn, C5_NUM_IMG, overlap_size=64)
            self.fn = partial(dc_model_fn, model=worker.eps_scalar_t_fn, num_classes=3, class_cond=True, cfg=False, w=0.0)
            self.shape = (B, 4, worker.shape[2], 16)
        else:
            self.fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
            self.shape = (B, 4, 128, 16)
        self.kwargs = {"y": torch.ones(B, dtype=torch.long, device=dev)}
        if config != "c2":
            self.kwargs["rule"] = {"pitch_hist": torch.tensor([TARGET], device=dev).repeat(B, 1)}
        self.scg = None if config == "c2" else {"num_samples": N, "pitch_hist": 1.0}
        self.bytes_io = 4 * B * 4 * self.shape[2] * 16

    def step(self, x, i):
        ti = self.T - 1 - (i % (self.T - 1))
        t = torch.full((self.B,), ti, device=self.dev, dtype=torch.long)
        with torch.no_grad():
            if self.config == "c2":
                return self.diffusion.p_sample(self.fn, x, t, model_kwargs=self.kwargs, _t_host=ti)["sample"]
            return self.diffusion.ddim_sample(self.fn, x, t, model_kwargs=self.kwargs, eta=1.0, embed_model=self.vae,
                                              scale_factor=SCALE, guidance_kwargs=GUIDANCE, scg_kwargs=self.scg,
                                              _t_host=ti)["sample"]


def timed_region(wl, x, k, steps, barrier, ncu_range=False):
    """K steps between barriers, CUDA events on the launching stream.  Returns (ms, x, k)."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        if ncu_range and i == 0:  # `ncu --profile-from-start off`: the launch list of exactly one timed step
            torch.cuda.profiler.start()
        x = wl.step(x, k)
        if ncu_range and i == 0:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        k += 1
    e1.record()
    barrier()
    return e0.elapsed_time(e1), x, k


def extra_leg(name, B, N, shard_cands, dev, model, vae, barrier, steps=4, warmup=3):
    """A short secondary measurement (same models): returns a dict for the `extra` block of the JSON line."""
    from rule_guided_music_b200.guided_diffusion import dist_util as du

    world = du.dist.get_world_size() if du.dist.is_initialized() else 1
    rank = du.dist.get_rank() if du.dist.is_initialized() else 0
    du.shard_candidates(shard_cands)
    try:
        wl = Workload("c3", B, N, dev, model, vae, graphs=True)
        torch.manual_seed(4321 if shard_cands else 4321 + rank)
        x = torch.randn(*wl.shape, device=dev)
        k = 0
        for _ in range(warmup):
            x = wl.step(x, k)
            k += 1
        ms, x, k = timed_region(wl, x, k, steps, barrier)
        ms = du.max_over_ranks(ms, dev)
        out = {"what": name, "batch_per_rank": B, "candidates": N, "ranks": world, "steps": steps, "warmup": warmup,
               "ms_per_step": ms / steps, "value": steps / (ms * 1e-3), "unit": "steps/s of the whole job's batch",
               "finite": bool(torch.isfinite(x).all().item())}
        if shard_cands and world > 1:
            # device time of the exchange alone: the same all-gather + first-max on this step's row sizes, eager
            group = du.candidate_sharding()[2]
            best = torch.zeros(B, device=dev)
            idx = torch.zeros(B, device=dev, dtype=torch.int64)
            win = torch.zeros(*wl.shape, device=dev)
            for _ in range(3):
                du.first_max_over_ranks(best, idx, win, group=group)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                du.first_max_over_ranks(best, idx, win, group=group)
            b.record()
            torch.cuda.synchronize()
            out["exchange_ms"] = du.max_over_ranks(a.elapsed_time(b) / 20, dev)
            out["exchange"] = ("ncclAllGather of [%d ranks, %d samples, %d B] rows (score, index, winning latent) + "
                               "first-max over ranks, once per step inside the captured graph"
                               % (world, B, 12 + 4 * win[0].numel()))
        return out
    finally:
        du.shard_candidates(False)


def run_b200(args):
    from rule_guided_music_b200 import synthetic_weights as ow  # seeded state dicts, reference keys (no checkpoints offline)
    from rule_guided_music_b200 import _lib
    from rule_guided_music_b200.guided_diffusion import dist_util
    from rule_guided_music_b200.guided_diffusion.dit import DiT_models
    from rule_guided_music_b200.taming.models.klvae_pedal import AutoencoderKL

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    rank, world = dist_util.setup_dist(dev)
    config = args.config
    B = args.batch if args.batch > 0 else CONFIG_DEFAULTS[config][0]
    N = args.candidates if args.candidates >= 0 else CONFIG_DEFAULTS[config][1]
    shard_cands = args.shard == "candidates" and world > 1 and config != "c2"
    strong = args.scaling == "strong" and not shard_cands
    if strong:
        if B % world != 0:
            raise SystemExit(f"bench.py --scaling strong: the global batch {B} is not divisible by {world} ranks")
        B //= world  # `--batch` is the GLOBAL batch under strong scaling

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
    dist_util.shard_candidates(shard_cands)
    wl = Workload(config, B, N, dev, model, vae, graphs=not args.no_graph)
    torch.manual_seed(1234 + (0 if shard_cands else rank))  # candidate sharding: every rank holds the SAME batch
    x = torch.randn(*wl.shape, device=dev)

    def barrier():
        dist_util.barrier()
        torch.cuda.synchronize()

    k = 0
    launches_per_step = 0
    for w in range(args.warmup):
        l_w = _lib.launch_count()
        x = wl.step(x, k)
        if w == 0:  # the first step of a kind always runs eagerly: count the kernels one step launches
            launches_per_step = _lib.launch_count() - l_w
        k += 1
    # ---- timed region: inputs resident in HBM ---------------------------------------------------------------------
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.3)
    l0 = _lib.launch_count()
    ms, x, k = timed_region(wl, x, k, args.steps, barrier, ncu_range=args.ncu_range)
    launches = _lib.launch_count() - l0
    graphed = wl.diffusion.captured_graphs()
    if graphed:  # replayed kernels are not seen by the library's launch counter: kernels per step x steps
        launches = launches_per_step * args.steps
    clk = clocks.stop() if clocks else None
    # ---- end to end: host buffers in, host buffers out, every step --------------------------------------------------
    x_host = torch.empty(*wl.shape).pin_memory()
    x_host.copy_(x)
    out_host = torch.empty(*wl.shape).pin_memory()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        xd = x_host.to(dev, non_blocking=True)
        out_host.copy_(wl.step(xd, k), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller owns the result on the host before the next step
        x_host.copy_(out_host)
        k += 1
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    ms, ms_e2e = dist_util.max_over_ranks(ms, dev), dist_util.max_over_ranks(ms_e2e, dev)
    gathered = dist_util.gather_samples(x)  # once, outside the timed region (scripts/cfg_sample.py:102-109)
    finite = bool(torch.isfinite(gathered).all().item()) and gathered.shape[0] == world * B
    gn_timeouts = int(vae.gn_timeouts())
    dist_util.shard_candidates(False)

    # ---- extra legs (see the module docstring) ----------------------------------------------------------------------
    extra = {}
    if config == "c3" and not strong and not shard_cands and not args.no_extra and B == B_FULL and N == N_FULL:
        if world > 1:
            extra["strong_scaling"] = extra_leg("global batch 64 split over the ranks, N=16 (strong scaling of config 3)",
                                                B_FULL // world, N_FULL, False, dev, model, vae, barrier)
        extra["candidate_sharded"] = extra_leg(
            "B=8, N=64, pitch_hist (BASELINE config 4's shape): every rank holds the batch, candidates split over the ranks",
            8, 64, world > 1, dev, model, vae, barrier)

    # ---- per-kernel device times for the roofline (separate pass, events around every launch) ----------------------
    prof = {}
    if args.prof_steps > 0:
        dist_util.shard_candidates(shard_cands)
        vae.set_lanes(1)  # serial execution so that each launch's event pair times that launch alone
        model.set_lanes(1)
        wl.diffusion.enable_cuda_graphs(False)  # the per-launch event pairs are host-side calls: eager
        _lib.prof_enable(True)
        for _ in range(args.prof_steps):
            x = wl.step(x, k)
            k += 1
        prof = _lib.prof_summary()
        _lib.prof_enable(False)
        dist_util.shard_candidates(False)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md)"
        gemm = {n: v for n, v in prof.items() if n.startswith("gemm_tc") or n.startswith("conv_gn")}
        # the launch shape that takes the largest share of the step, with its DRAM traffic from the committed ncu capture
        dominant, traffic = None, None
        if gemm:
            dn, dv = max(gemm.items(), key=lambda kv: kv[1]["ms"])
            d_ach = dv["flops_alg"] / (dv["ms"] * 1e-3) / 1e12 if dv["ms"] > 0 else 0.0
            dominant = {"launch": dn, "launches_per_step": dv["launches"] / max(args.prof_steps, 1),
                        "avg_launch_ms": dv["ms"] / max(dv["launches"], 1), "achieved": d_ach,
                        "frac": d_ach / tensor_peak, "algorithmic_flops_per_launch": dv["flops_alg"] / max(dv["launches"], 1)}
            for tf in ("r2_traffic.json", "r1_traffic.json"):
                try:
                    tr = json.load(open(os.path.join(ROOT, "profiles", tf))).get(dn)
                except Exception:
                    tr = None
                if tr and config == "c3" and B == B_FULL and N == N_FULL:
                    # the captures are launches over 128 VAE tiles; a launch of this run covers RGM_VAE_CHUNK tiles
                    # (default 256) and moves proportionally more
                    per = int(os.environ.get("RGM_VAE_CHUNK", "256")) / 128.0 if " H1 " not in dn else 1.0
                    traffic = tr["traffic_bytes"] * per
                    dominant.update(algorithmic_bytes_per_launch=tr["algorithmic_bytes"] * per, traffic_bytes_per_launch=traffic,
                                    traffic_source="profiles/%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % tf)
                    break
        g_ms = sum(v["ms"] for v in gemm.values())
        g_alg = sum(v["flops_alg"] for v in gemm.values())
        g_exec = sum(v["flops_exec"] for v in gemm.values())
        g_n = sum(v["launches"] for v in gemm.values())
        all_ms = sum(v["ms"] for v in prof.values())
        one_batch = shard_cands or strong     # all ranks work on one (global) batch: a step is the whole batch
        job_steps = args.steps if one_batch else world * args.steps
        value = job_steps / (ms * 1e-3)
        e2e_value = job_steps / (ms_e2e * 1e-3)
        fl_job = step_flops(B, N, config) * (1 if shard_cands else world)  # FLOPs of one step of the whole job
        step_tflops = fl_job * args.steps / (ms * 1e-3) / 1e12
        per_gpu_tflops = step_tflops / world
        line = {
            "metric": metric_name(config, N), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if one_batch else "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate/residual", "data": "synthetic",
            "config": workload_config(B, N, world, config, args.scaling, "candidates" if shard_cands else "batch"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": wl.bytes_io, "d2h_bytes_per_step": wl.bytes_io},
            "gpu_launches": int(launches),
            "launch_mode": ("cuda graph replay, %d kernel nodes per step" % launches_per_step) if graphed else "eager",
            "clocks": clk,
            # whole step: algorithmic FLOPs of everything one step computes / the step's device time / measured peak
            "roofline": {"bound": "tensor", "achieved": per_gpu_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": per_gpu_tflops / tensor_peak, "traffic": traffic,
                         "scope": "whole step, per GPU: SURVEY.md 8(d) algorithmic FLOPs / ms_per_step / measured peak",
                         "peak_source": peak_src,
                         "gemm_family_frac": (g_alg / (g_ms * 1e-3) / 1e12 / tensor_peak) if g_ms > 0 else None,
                         "executed_frac": (g_exec / (g_ms * 1e-3) / 1e12 / tensor_peak) if g_ms > 0 else None,
                         "gemm_family": "all tcgen05 implicit-GEMM launches (DiT linears, VAE convolutions) over their own "
                                        "device time; executed_frac counts the 4 taps an upsample conv executes, not 9",
                         "gemm_share_of_step_device_time": g_ms / all_ms if all_ms > 0 else None,
                         "gemm_launches_per_step": g_n / max(args.prof_steps, 1),
                         "dominant_launch": dominant},
            "step_tflops_algorithmic": step_tflops,
            "finite": finite,
            # convolutions that normalise their own output wait for each other's tiles; a wait that gave up would have
            # produced garbage -- never in a healthy run, and checked here so that such a run cannot report a number
            "gn_timeouts": gn_timeouts,
        }
        if gn_timeouts:
            raise SystemExit("bench: a GroupNorm-in-epilogue wait gave up (rgm_vae_gn_timeouts != 0): results invalid")
        if extra:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline and config == "c3":
            cores = host_threads()
            cb, cn = cpu_sample_shape(cores)
            v, sec, scale, kind = reference_steps(2, 0, cb, cn)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": kind,
                "sample": f"{'the unmodified reference' if kind == 'reference' else 'oracle port'} at B={cb}, N={cn}: "
                          f"{sec:.2f} s per sampled step on {cores} threads (mean of 2 steps), scaled "
                          f"x{scale:.1f} (algorithmic FLOPs) to B=64, N=16"}
        if args.prof_out:
            with open(args.prof_out, "w") as f:
                json.dump({"per_kernel_family": prof, "prof_steps": args.prof_steps, "ms_per_step_unprofiled":
                           ms / args.steps}, f, indent=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        import gc

        import torch.distributed as dist

        # captured graphs that contain NCCL kernels must be gone before the communicator is torn down (destroying the
        # process group with such a graph alive blocks in ncclCommDestroy)
        wl.diffusion.enable_cuda_graphs(False)
        del wl
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c5"], help="BASELINE.json config (default c3, the headline)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch per GPU (default); strong: --batch is the global batch, split over the ranks")
    ap.add_argument("--shard", default="batch", choices=["batch", "candidates"],
                    help="candidates: every rank holds the whole batch and a share of the N candidates (one all-gather per step)")
    ap.add_argument("--batch", type=int, default=0, help="0 = the config's own (c2: 256, c3: 64, c5: 1)")
    ap.add_argument("--candidates", type=int, default=-1, help="-1 = the config's own (16)")
    ap.add_argument("--ref-batch", type=int, default=0, help="reference arms: override the bounded sample's batch")
    ap.add_argument("--ref-candidates", type=int, default=0, help="reference arms: override the bounded sample's N")
    ap.add_argument("--prof-steps", type=int, default=1)
    ap.add_argument("--prof-out", default="")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling / candidate-sharded extra legs")
    ap.add_argument("--ncu-range", action="store_true",
                    help="bracket the first timed step with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl != "b200":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
