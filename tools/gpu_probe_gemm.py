"""Quick throughput probe of the implicit-GEMM kernel on a B200 (not a bench line; development aid)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda:0")
    res = []
    for (M, N, K, bn) in [(262144, 3456, 1152, 128), (262144, 1152, 1152, 128), (262144, 4608, 1152, 256),
                          (262144, 4608, 1152, 128), (262144, 1152, 4608, 128), (65536, 512, 4608, 256),
                          (65536, 512, 4608, 128)]:
        a = torch.randn(M, K, device=dev).half()
        b = torch.randn(N, K, device=dev).half()
        bias = torch.zeros(N, device=dev)
        out = torch.empty(M, N, device=dev)
        ms = timeit(lambda: _lib.call("rgm_gemm_f16", _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), _lib.ptr(out), M, N, K,
                                      bn, _lib.stream_ptr()))
        tf = 2.0 * M * N * K / ms / 1e9
        ms_t = timeit(lambda: torch.matmul(a, b.t()))
        res.append(dict(op="linear", M=M, N=N, K=K, bn=bn, ms=ms, tflops=tf, torch_ms=ms_t,
                        torch_tflops=2.0 * M * N * K / ms_t / 1e9))
        print(res[-1], flush=True)
        del a, b, out
    for (n, H, cin, cout, kind, bn) in [(64, 128, 128, 128, 1, 128), (64, 128, 256, 128, 1, 128),
                                        (64, 64, 256, 256, 1, 256), (64, 64, 256, 256, 1, 128),
                                        (256, 16, 512, 512, 1, 256), (64, 64, 256, 256, 2, 256)]:
        x = torch.randn(n, H, H, cin, device=dev).half()
        taps = 9 if kind == 1 else 4
        npar = 4 if kind == 2 else 1
        wp = torch.randn(npar * cout * taps * cin, device=dev).half() * 0.01
        bias = torch.zeros(cout, device=dev)
        s = 2 if kind == 2 else 1
        out = torch.empty(n, H * s, H * s, cout, device=dev, dtype=torch.float16)
        ms = timeit(lambda: _lib.call("rgm_conv_f16", _lib.ptr(x), _lib.ptr(wp), _lib.ptr(bias), None, _lib.ptr(out), n,
                                      H, H, cin, cout, kind, bn, None, _lib.stream_ptr()))
        flops = 2.0 * n * (H * s) * (H * s) * cout * cin * 9  # reference-equivalent flops (3x3 on the output grid)
        res.append(dict(op="conv", n=n, H=H, cin=cin, cout=cout, kind=kind, bn=bn, ms=ms, ref_tflops=flops / ms / 1e9))
        print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/probe_gemm.json", "w"), indent=1)


if __name__ == "__main__":
    main()
