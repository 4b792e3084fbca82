"""Development aid: pipeline timeline (CTA 0) of one GEMM family inside a real DiT forward.  EPI=3 (QKV), 0 (fc1), 2 (proj/fc2)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util, golden_inputs as gi
dev = torch.device("cuda:0")
model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], dev)
B = 256
x = torch.randn(B, 4, 128, 16, device=dev); t = torch.full((B,), 500, device=dev); y = torch.ones(B, dtype=torch.long, device=dev)
model(x, t, y); torch.cuda.synchronize()
for epi, n in ((3, 3456), (0, 4608), (2, 1152)):
    tr = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
    os.environ["RGM_DEBUG_TRACE_PTR"] = str(tr.data_ptr()); os.environ["RGM_DEBUG_TRACE_EPI"] = str(epi); os.environ["RGM_DEBUG_TRACE_N"] = str(n)
    model(x, t, y); torch.cuda.synchronize()
    for k in ("RGM_DEBUG_TRACE_PTR", "RGM_DEBUG_TRACE_EPI", "RGM_DEBUG_TRACE_N"): del os.environ[k]
    tt = tr.view(-1, 8).cpu(); nt = int((tt[:, 0] != 0).sum()); tt = tt[:nt].double(); t0 = tt[0, 0]
    print("epi", epi, "N", n, "tiles of CTA0:", nt)
    for i in (3, 4, 5):
        r = tt[i]
        print(f"  tile {i:3d} prod {r[0]-t0:9.0f} acc_free {r[1]-t0:9.0f} first_land {r[2]-t0:9.0f} last_issue {r[3]-t0:9.0f} epi_start {r[4]-t0:9.0f} epi_done {r[5]-t0:9.0f} | mma {r[3]-r[1]:7.0f} epi {r[5]-r[4]:7.0f}")
    print(f"  avg cycles/tile {(tt[nt-1,5]-tt[1,5])/(nt-2):.0f}")
