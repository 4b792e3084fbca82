"""Development aid: DiT forward time vs the library's sample chunk (RGM_DIT_CHUNK): smaller chunks keep the residual
stream and the q/k/v tensors in the 126 MB L2 between the kernels of a block."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi, gpu_util
dev = torch.device("cuda:0")
B = 1024
x = torch.randn(B, 4, 128, 16, device=dev); t = torch.full((B,), 500, device=dev); y = torch.ones(B, dtype=torch.long, device=dev)
for chunk in (256, 128, 64, 48, 32, 256):
    os.environ["RGM_DIT_CHUNK"] = str(chunk)
    model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], dev)
    for _ in range(2): model(x, t, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): model(x, t, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"chunk {chunk:4d}: {ms:7.2f} ms per forward of {B} samples (2 blocks)", flush=True)
    del model
