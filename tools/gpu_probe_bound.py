"""Development aid: what bounds the DiT linears?  Runs the real DiT forward (256-sample chunk) with the pair kernel's
operand loads and/or epilogue switched off (RGM_GEMM_DEBUG bit 0 / bit 1; results are garbage, timings are the point)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi, gpu_util
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], dev)
B = 256
x = torch.randn(B, 4, 128, 16, device=dev); t = torch.full((B,), 500, device=dev); y = torch.ones(B, dtype=torch.long, device=dev)
for _ in range(20): model(x, t, y)   # warm the chip into its sustained (power-capped) state
torch.cuda.synchronize()
names = {0: "full kernel", 1: "no operand loads", 2: "no epilogue", 3: "MMA only", 4: "epilogue = TMEM reads only", 5: "TMEM reads only, no loads", 8: "epilogue without its stores"}
for rep in range(2):
    for dbg in (0, 8, 2, 4):
        os.environ["RGM_GEMM_DEBUG"] = str(dbg)
        for _ in range(3): model(x, t, y)
        torch.cuda.synchronize()
        _lib.prof_enable(True)
        for _ in range(6): model(x, t, y)
        prof = _lib.prof_summary(); _lib.prof_enable(False)
        row = []
        for key in ("K1152 N4608", "K1152 N3456", "K1152 N1152 epi2", "K4608 N1152"):
            v = [p for n, p in prof.items() if key in n and "gemm" in n]
            ms = sum(p["ms"] for p in v); fl = sum(p["flops_alg"] for p in v)
            row.append(f"{key}: {fl/ms/1e9:7.1f}")
        print(f"{names[dbg]:28s} TF/s  " + " | ".join(row), flush=True)
os.environ["RGM_GEMM_DEBUG"] = "0"
