"""Development aid: cycles per tile of CTA 0 for the DiT linears (pair kernel) with parts switched off
(RGM_GEMM_DEBUG: 1 no operand loads, 2 no epilogue, 4 epilogue = TMEM reads only, 8 epilogue computes, no stores)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
def linear(M, N, K):
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half()
    bias = torch.zeros(N, device=dev); out = torch.zeros(M, N, device=dev)
    return lambda: _lib.call("rgm_gemm_f16", _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), _lib.ptr(out), M, N, K, 0, _lib.stream_ptr())
for name, fn in (("K1152 N4608 (f32 out)", linear(65536, 4608, 1152)), ("K4608 N1152 (f32 out)", linear(65536, 1152, 4608))):
    for dbg in (0, 1, 2, 4, 8, 3):
        os.environ["RGM_GEMM_DEBUG"] = str(dbg)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        tr = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
        os.environ["RGM_DEBUG_TRACE_PTR"] = str(tr.data_ptr())
        fn(); torch.cuda.synchronize()
        del os.environ["RGM_DEBUG_TRACE_PTR"]
        t = tr.view(-1, 8).cpu(); nt = int((t[:, 0] != 0).sum()); t = t[:nt].double()
        mma = (t[2:nt-1, 3] - t[2:nt-1, 1]).mean().item()
        epi = (t[2:nt-1, 5] - t[2:nt-1, 4]).mean().item()
        per = ((t[nt-1, 3] - t[1, 3]) / (nt - 2)).item()
        print(f"{name} debug {dbg}: {ms:.3f} ms, tiles of CTA0 {nt}, cycles/tile {per:.0f}, mma loop {mma:.0f}, epilogue {epi:.0f}")
os.environ["RGM_GEMM_DEBUG"] = "0"
