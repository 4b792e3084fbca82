"""Development aid: per-tile pipeline timeline of CTA 0 of the implicit-GEMM kernel (GemmParams::trace)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
def linear(M, N, K, bn):
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half()
    bias = torch.zeros(N, device=dev); out = torch.zeros(M, N, device=dev)
    return lambda: _lib.call("rgm_gemm_f16", _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), _lib.ptr(out), M, N, K, bn, _lib.stream_ptr())
def conv(n, H, cin, cout, kind, bn):
    x = torch.randn(n, H, H, cin, device=dev).half()
    taps = {0: 1, 1: 9, 2: 4}[kind]; npar = 4 if kind == 2 else 1
    w = torch.randn(npar * cout * taps * cin, device=dev).half() * 0.02
    bias = torch.zeros(cout, device=dev); s = 2 if kind == 2 else 1
    out = torch.empty(n, H * s, H * s, cout, device=dev, dtype=torch.float16)
    part = torch.zeros(n * H * H * s * s // 32 * cout // 4 * 2 + 16, device=dev)
    return lambda: _lib.call("rgm_conv_f16", _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), None, _lib.ptr(out), n, H, H, cin, cout, kind, bn, _lib.ptr(part), _lib.stream_ptr())
cases = {"lin_k1152_n1152": linear(262144, 1152, 1152, 128), "lin_k1152_n4608": linear(65536, 4608, 1152, 256),
         "conv128_k1152_n128": conv(64, 128, 128, 128, 1, 128), "conv64_k2304_n256": conv(64, 64, 256, 256, 1, 256),
         "up2_k1024_n256": conv(64, 64, 256, 256, 2, 256)}
for name, fn in cases.items():
    tr = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
    fn(); torch.cuda.synchronize()
    os.environ["RGM_DEBUG_TRACE_PTR"] = str(tr.data_ptr())
    fn(); torch.cuda.synchronize()
    del os.environ["RGM_DEBUG_TRACE_PTR"]
    t = tr.view(-1, 8).cpu()
    n = int((t[:, 0] != 0).sum())
    t = t[:n].double(); t0 = t[0, 0]
    print(name, "tiles of CTA0:", n)
    for i in list(range(min(n, 6))) + ([n - 2, n - 1] if n > 8 else []):
        r = t[i]
        print(f"  tile {i:3d} prod_start {r[0]-t0:9.0f} acc_free {r[1]-t0:9.0f} first_land {r[2]-t0:9.0f} last_issue {r[3]-t0:9.0f} epi_start {r[4]-t0:9.0f} epi_done {r[5]-t0:9.0f} | mma {r[3]-r[1]:7.0f} epi {r[5]-r[4]:7.0f} tmem_ld {r[6]:7.0f}")
    if n > 3:
        print(f"  avg cycles/tile {(t[n-1,5]-t[1,5])/(n-2):.0f}")
