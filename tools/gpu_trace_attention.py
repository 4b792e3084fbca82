"""Development aid: per-pair pipeline timeline of CTA 0 of the tcgen05 attention kernel (csrc/attention.cu) at the DiT's
shape (16 heads x 72, T = 256, 256 samples = one bench chunk)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rule_guided_music_b200 import _lib
dev = torch.device("cuda:0")
B, heads, T, dh = 256, 16, int(os.environ.get("T", "256")), 72
q = torch.randn(B, heads, T, dh, device=dev).half()
k = torch.randn(B, heads, T, dh, device=dev).half()
vt = torch.randn(B, heads, dh, T, device=dev).half()
out = torch.empty(B * T, heads * dh, device=dev, dtype=torch.float16)
fn = lambda: _lib.call("rgm_attention_f16", _lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(out), B, heads, T, dh, dh ** -0.5, _lib.stream_ptr())
for _ in range(3): fn()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): fn()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"T {T}: {ms*1e3:.1f} us per launch ({B*heads} pairs), {4.0*B*heads*T*T*dh/ms/1e9:.0f} TFLOP/s")
tr = torch.zeros(16 * 64, dtype=torch.int64, device=dev)
os.environ["RGM_DEBUG_TRACE_PTR"] = str(tr.data_ptr())
fn(); torch.cuda.synchronize()
del os.environ["RGM_DEBUG_TRACE_PTR"]
t = tr.view(-1, 16).cpu(); n = int((t[:, 0] != 0).sum()); t = t[:n].double(); t0 = t[0, 0]
names = ["start", "A_landed", "S_issued", "P0_seen", "PV0_issued", "P1_seen", "PV1_issued", "-", "S0_ready", "max0_done", "P0_written",
         "S1_ready", "P1_written", "O0_ready", "epi_done"]
print("pairs of CTA 0:", n, " avg cycles/pair", (t[n - 1, 14] - t[1, 14]) / (n - 2))
for i in (4, 5):
    r = t[i]
    print(f" pair {i}: " + " ".join(f"{nm}={r[j]-r[0]:.0f}" for j, nm in enumerate(names) if nm != "-") + f" | next start {t[i+1,0]-r[0]:.0f}")
