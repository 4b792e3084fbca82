"""Development aid: per-tile pipeline timeline of CTA 0 of the fused GroupNorm + conv kernel (csrc/conv_gn.cuh) at the
128 -> 128 @ 128x128 shape of one bench chunk, full and with parts switched off (RGM_GEMM_DEBUG bits: 16 = transform
warps skip the arithmetic, 2 = no epilogue), next to the two-pass form (gn_apply + gemm_sw_kernel) timed with events."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
n, H, W, cin, cout = 128, 128, 128, int(os.environ.get("CIN", "128")), 128
x = torch.randn(n, H, W, cin, device=dev).half()
ab = torch.stack((torch.rand(n, cin, device=dev) + 0.5, torch.randn(n, cin, device=dev) * 0.3), dim=-1).contiguous()
w = torch.randn(cout * 9 * cin, device=dev).half() * 0.02
bias = torch.zeros(cout, device=dev)
out = torch.empty(n, H, W, cout, device=dev, dtype=torch.float16)
y = torch.empty_like(x)
part = torch.zeros(n * H * W // 128 * cout // 4 * 2 + 16, device=dev)
fused = lambda: _lib.call("rgm_conv_gn_f16", _lib.ptr(x), _lib.ptr(ab), _lib.ptr(w), _lib.ptr(bias), None, _lib.ptr(out), n, H, W, cin, cout, _lib.ptr(part), _lib.stream_ptr())
def two_pass():
    _lib.call("rgm_gn_apply_f16", _lib.ptr(x), _lib.ptr(ab), _lib.ptr(y), n, H * W, cin, 1, _lib.stream_ptr())
    _lib.call("rgm_conv_f16", _lib.ptr(y), _lib.ptr(w), _lib.ptr(bias), None, _lib.ptr(out), n, H, W, cin, cout, 1, 0, _lib.ptr(part), _lib.stream_ptr())
def timeit(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
fl = 2.0 * n * H * W * cout * 9 * cin
print(f"Cin {cin}: two-pass (gn_apply + conv) {timeit(two_pass):.3f} ms")
for dbg in (0, 16, 18, 2):
    os.environ["RGM_GEMM_DEBUG"] = str(dbg)
    ms = timeit(fused)
    tr = torch.zeros(32 * 256, dtype=torch.int64, device=dev)
    os.environ["RGM_DEBUG_TRACE_PTR"] = str(tr.data_ptr())
    fused(); torch.cuda.synchronize()
    del os.environ["RGM_DEBUG_TRACE_PTR"]
    t = tr.view(-1, 32).cpu(); nt = int((t[:, 8] != 0).sum()); t = t[:nt].double(); t0 = t[0, 0]
    print(f"debug {dbg:2d}: fused {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s; tiles of CTA0: {nt}, avg cycles/tile {(t[nt-1,17]-t[1,17])/(nt-2):.0f}")
    names = ["tma0", "tma1", "tr0_start", "tr0_half", "tr0_done", "tr1_start", "tr1_half", "tr1_done", "acc_free", "mma0_h0", "mma0_h1",
             "mma0_end", "mma1_h0", "mma1_h1", "mma1_end", "-", "epi_start", "epi_done"]
    for i in (5, 6):
        r = t[i]
        print("   tile %d: " % i + " ".join(f"{nm}={r[j]-t0:.0f}" for j, nm in enumerate(names) if nm != "-"))
        print(f"      transform kb0 {r[4]-r[2]:.0f} kb1 {r[7]-r[5]:.0f} | tma->landed kb0 {r[2]-r[0]:.0f} kb1 {r[5]-r[1]:.0f} | "
              f"mma kb0 issue {r[11]-r[9]:.0f} (wait h1 {r[10]-r[9]:.0f}) kb1 {r[14]-r[12]:.0f} | epi {r[17]-r[16]:.0f}")
os.environ["RGM_GEMM_DEBUG"] = "0"
