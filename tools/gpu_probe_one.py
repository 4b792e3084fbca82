"""One GEMM shape, few launches: target for ncu. Env: SHAPE=M,N,K,bn  SKIP_STORE=1"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib
M, N, K, bn = [int(v) for v in os.environ.get("SHAPE", "262144,1152,1152,128").split(",")]
iters = int(os.environ.get("ITERS", "3"))
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).half(); b = torch.randn(N, K, device=dev).half()
bias = torch.zeros(N, device=dev); out = torch.zeros(M, N, device=dev)
def run():
    _lib.call("rgm_gemm_f16", _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), _lib.ptr(out), M, N, K, bn, _lib.stream_ptr())
run(); torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(iters): run()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / iters
print(f"SHAPE {M},{N},{K},{bn} ms {ms:.3f} TF {2.0*M*N*K/ms/1e9:.1f}", flush=True)
