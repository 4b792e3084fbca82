"""One conv / linear shape, few launches: target for ncu.  CASE=conv128|lin4608"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
case = os.environ.get("CASE", "conv128")
if case in ("conv128", "conv64"):
    n, H, cin, cout = (64, 128, 128, 128) if case == "conv128" else (128, 64, 256, 256)
    x = torch.randn(n, H, H, cin, device=dev).half()
    w = torch.randn(cout * 9 * cin, device=dev).half() * 0.02
    bias = torch.zeros(cout, device=dev)
    out = torch.empty(n, H, H, cout, device=dev, dtype=torch.float16)
    part = torch.zeros(n * H * H // 32 * cout // 4 * 2 + 16, device=dev)
    fn = lambda: _lib.call("rgm_conv_f16", _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), None, _lib.ptr(out), n, H, H, cin, cout, 1, 0, _lib.ptr(part), _lib.stream_ptr())
else:
    M, N, K = 65536, 4608, 1152
    a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half()
    bias = torch.zeros(N, device=dev); out = torch.zeros(M, N, device=dev)
    fn = lambda: _lib.call("rgm_gemm_f16", _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), _lib.ptr(out), M, N, K, 0, _lib.stream_ptr())
for _ in range(3):
    fn()
torch.cuda.synchronize()
