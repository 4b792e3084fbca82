"""Short target for `ncu --set full --profile-from-start off`: one DiT forward chunk (256 samples, 2 blocks of XL
width) and one VAE decode chunk (128 tiles) at the shapes the bench step launches, bracketed by cudaProfilerStart/Stop
after a warm-up pass.  PART=dit|vae|all selects what runs inside the profiled range."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi  # noqa: E402
import gpu_util  # noqa: E402

dev = torch.device("cuda:0")
part = os.environ.get("PART", "all")
model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], dev)
vae, _ = gpu_util.native_vae(dev)
vae.set_lanes(1)
B = 256
x = torch.randn(B, 4, 128, 16, device=dev)
t = torch.full((B,), 500, device=dev)
y = torch.ones(B, dtype=torch.long, device=dev)
lat = torch.randn(16, 4, 128, 16, device=dev)


def run():
    if part in ("dit", "all"):
        model(x, t, y)
    if part in ("vae", "all"):
        vae.decode_latents(lat, 1.2465, channels=1)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
