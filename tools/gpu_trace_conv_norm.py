"""Development aid: per-tile pipeline timeline of CTA 0 for the convolutions that normalise their own output in the
epilogue (gn_epilogue_loop), next to the plain convolution of the same shape.  RGM_GEMM_DEBUG=16 skips the wait."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
for (n, H, cin, cout) in ((128, 128, 128, 128), (128, 64, 256, 256)):
    x = torch.randn(n, H, H, cin, device=dev).half()
    w = torch.randn(cout * 9 * cin, device=dev).half() * 0.02
    bias = torch.zeros(cout, device=dev)
    gamma = torch.ones(cout, device=dev); beta = torch.zeros(cout, device=dev)
    out = torch.empty(n, H, H, cout, device=dev, dtype=torch.float16)
    part = torch.zeros(n * H * H // 128 * cout // 4 * 2 + 16, device=dev)
    scratch = torch.zeros(n * 128, device=dev, dtype=torch.int32)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    plain = lambda: _lib.call("rgm_conv_f16", _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), None, _lib.ptr(out), n, H, H, cin, cout, 1, 0, _lib.ptr(part), _lib.stream_ptr())
    fused = lambda: _lib.call("rgm_conv_norm_f16", _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta), None, None, _lib.ptr(out), n, H, H, cin, cout, 1, 1, _lib.ptr(scratch), _lib.ptr(err), _lib.stream_ptr())
    raw = torch.empty_like(out)
    dual = lambda: _lib.call("rgm_conv_norm_f16", _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(x) if cin == cout else None, _lib.ptr(raw), _lib.ptr(out), n, H, H, cin, cout, 1, 1, _lib.ptr(scratch), _lib.ptr(err), _lib.stream_ptr())
    plain_res = lambda: _lib.call("rgm_conv_f16", _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(x), _lib.ptr(raw), n, H, H, cin, cout, 1, 0, _lib.ptr(part), _lib.stream_ptr())
    for name, fn, dbg in (("plain", plain, 0), ("norm in epilogue", fused, 0), ("norm in epilogue, no wait", fused, 16), ("plain + residual", plain_res, 0), ("dual: raw + residual and normalised copy", dual, 0)):
        os.environ["RGM_GEMM_DEBUG"] = str(dbg)
        for _ in range(3): fn()
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
        t = tr.view(-1, 8).cpu(); nt = int((t[:, 0] != 0).sum()); t = t[:nt].double(); t0 = t[0, 0]
        print(f"conv3x3 {cin}->{cout} @{H}x{H} x{n}: {name}: {ms:.3f} ms per launch, tiles of CTA 0: {nt}, err flag {err.item()}")
        for i in (4, 5, 6, 7):
            r = t[i]
            extra = f" published {r[6]-t0:9.0f} complete {r[7]-t0:9.0f} | pass1 {r[6]-r[4]:6.0f} wait {r[7]-r[6]:6.0f} pass2 {r[5]-r[7]:6.0f}" if r[6] > 0 else ""
            print(f"  tile {i:3d} acc_free {r[1]-t0:9.0f} first_land {r[2]-t0:9.0f} last_issue {r[3]-t0:9.0f} epi_start {r[4]-t0:9.0f} epi_done {r[5]-t0:9.0f} | mma {r[3]-r[1]:7.0f} epi {r[5]-r[4]:7.0f}{extra}")
        print(f"  avg cycles/tile {(t[nt-1,5]-t[1,5])/(nt-2):.0f}")
os.environ["RGM_GEMM_DEBUG"] = "0"
