"""Development aid: per-tile pipeline timeline of CTA 0 for the 256->256 conv at 64x64 (pair kernel), with the
epilogue's stores on and off (RGM_GEMM_DEBUG=8)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
n, H, cin, cout = 128, 64, 256, 256
x = torch.randn(n, H, H, cin, device=dev).half()
w = torch.randn(cout * 9 * cin, device=dev).half() * 0.02
bias = torch.zeros(cout, device=dev)
out = torch.empty(n, H, H, cout, device=dev, dtype=torch.float16)
part = torch.zeros(n * H * H // 128 * cout // 4 * 2 + 16, device=dev)
fn = lambda: _lib.call("rgm_conv_f16", _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), None, _lib.ptr(out), n, H, H, cin, cout, 1, 256, _lib.ptr(part), _lib.stream_ptr())
for dbg in (0, 8, 4):
    os.environ["RGM_GEMM_DEBUG"] = str(dbg)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tr = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
    os.environ["RGM_DEBUG_TRACE_PTR"] = str(tr.data_ptr())
    fn(); torch.cuda.synchronize()
    del os.environ["RGM_DEBUG_TRACE_PTR"]
    t = tr.view(-1, 8).cpu(); nt = int((t[:, 0] != 0).sum()); t = t[:nt].double(); t0 = t[0, 0]
    print("debug", dbg, "tiles of CTA0:", nt)
    for i in (2, 3, 4):
        r = t[i]
        print(f"  tile {i:3d} acc_free {r[1]-t0:9.0f} first_land {r[2]-t0:9.0f} last_issue {r[3]-t0:9.0f} epi_start {r[4]-t0:9.0f} epi_done {r[5]-t0:9.0f} | mma {r[3]-r[1]:7.0f} epi {r[5]-r[4]:7.0f}")
    print(f"  avg cycles/tile {(t[nt-1,5]-t[1,5])/(nt-2):.0f}")
os.environ["RGM_GEMM_DEBUG"] = "0"
