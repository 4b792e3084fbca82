"""Development aid: what bounds the VAE pair-kernel convolutions?  One 128-tile decode chunk with the pair kernel's
operand loads and/or epilogue switched off (RGM_GEMM_DEBUG bit 0 / bit 1; results are garbage, timings are the point)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
vae, _ = gpu_util.native_vae(dev)
vae.set_lanes(1)
lat = torch.randn(16, 4, 128, 16, device=dev)
for _ in range(6): vae.decode_latents(lat, 1.2465, channels=1)
torch.cuda.synchronize()
names = {0: "full kernel", 1: "no operand loads", 2: "no epilogue", 3: "MMA only", 4: "epilogue = TMEM reads only", 8: "epilogue without its stores"}
for rep in range(2):
    for dbg in (0, 1, 2, 3, 4, 8):
        os.environ["RGM_GEMM_DEBUG"] = str(dbg)
        vae.decode_latents(lat, 1.2465, channels=1); torch.cuda.synchronize()
        _lib.prof_enable(True)
        for _ in range(3): vae.decode_latents(lat, 1.2465, channels=1)
        prof = _lib.prof_summary(); _lib.prof_enable(False)
        row = []
        for key in ("conv1 K2304 N256", "conv1 K4608 N512", "conv2 K1024 N256", "conv2 K2048 N512"):
            v = [p for n, p in prof.items() if key in n and "gemm" in n]
            ms = sum(p["ms"] for p in v); fl = sum(p["flops_exec"] for p in v)
            row.append(f"{key}: {fl/ms/1e9:7.1f}")
        print(f"{names[dbg]:28s} executed TF/s  " + " | ".join(row), flush=True)
os.environ["RGM_GEMM_DEBUG"] = "0"
