"""Development aid: DiT linear families (real epilogues, 256-sample chunk) under different pair-kernel rasterisations
(RGM_GEMM_BAND = feature-tile pairs per band; 0 = feature pairs fastest)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi, gpu_util
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], dev)
B = 256
x = torch.randn(B, 4, 128, 16, device=dev); t = torch.full((B,), 500, device=dev); y = torch.ones(B, dtype=torch.long, device=dev)
for _ in range(2): model(x, t, y)
torch.cuda.synchronize()
for band in [0, 1, 2, 3, 5, 9]:
    os.environ["RGM_GEMM_BAND"] = str(band)
    model(x, t, y); torch.cuda.synchronize()
    _lib.prof_enable(True)
    for _ in range(4): model(x, t, y)
    prof = _lib.prof_summary(); _lib.prof_enable(False)
    row = []
    for key in ("K1152 N4608", "K1152 N3456", "K1152 N1152 epi2", "K4608 N1152"):
        v = [p for n, p in prof.items() if key in n and "gemm" in n]
        ms = sum(p["ms"] for p in v); fl = sum(p["flops_alg"] for p in v)
        row.append(f"{key}: {fl/ms/1e9:7.1f} TF/s" if ms else f"{key}: n/a")
    print(f"band {band}: " + " | ".join(row), flush=True)
