"""Development aid: per-kernel-family device times of one VAE decode (256 tiles, serial lanes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
vae, _ = gpu_util.native_vae(dev)
lat = torch.randn(int(os.environ.get("NCAND", "32")), 4, 128, 16, device=dev)
for lanes in (1, 2):
    vae.set_lanes(lanes)
    for _ in range(2):
        vae.decode_latents(lat, 1.2465, channels=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        vae.decode_latents(lat, 1.2465, channels=1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    tiles = lat.shape[0] * 8
    print(f"lanes={lanes}: {ms:.2f} ms per decode of {tiles} tiles = {tiles*114.5e9/ms/1e9:.0f} TFLOP/s algorithmic")
vae.set_lanes(1)
_lib.prof_enable(True)
vae.decode_latents(lat, 1.2465, channels=1)
prof = _lib.prof_summary(); _lib.prof_enable(False)
tot = sum(v["ms"] for v in prof.values())
for n, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]:
    tf = v["flops_exec"] / v["ms"] / 1e9 if v["flops_exec"] else 0
    gb = v["bytes"] / v["ms"] / 1e6 if v["bytes"] else 0
    print(f"  {n:48s} n={v['launches']:4d} ms={v['ms']:7.2f} {100*v['ms']/tot:5.1f}%  TF={tf:7.1f} GB/s={gb:6.0f}")
