"""Short target for `ncu --set full --profile-from-start off`: one DiT forward chunk (256 samples, 2 blocks of XL
width) and one VAE decode chunk (128 tiles) at the shapes the bench step launches, bracketed by cudaProfilerStart/Stop
after a warm-up pass.  PART=dit|vae|all selects what runs inside the profiled range."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi  # noqa: E402
import gpu_util  # noqa: E402

dev = torch.device("cuda:0")
part = os.environ.get("PART", "all")
model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], dev)
vae, _ = gpu_util.native_vae(dev)
vae.set_lanes(1)
B = 256
x = torch.randn(B, 4, 128, 16, device=dev)
t = torch.full((B,), 500, device=dev)
y = torch.ones(B, dtype=torch.long, device=dev)
lat = torch.randn(16, 4, 128, 16, device=dev)


from rule_guided_music_b200 import _lib  # noqa: E402
from rule_guided_music_b200.music_rule_guidance import music_rules  # noqa: E402

# the fused GroupNorm + conv kernel at one bench chunk's shape (opt-in path, profiled for the bound analysis)
n_cg, cin = 128, 128
xg = torch.randn(n_cg, 128, 128, cin, device=dev).half()
abg = torch.stack((torch.rand(n_cg, cin, device=dev) + 0.5, torch.randn(n_cg, cin, device=dev) * 0.3), dim=-1).contiguous()
wg = torch.randn(128 * 9 * cin, device=dev).half() * 0.02
bg = torch.zeros(128, device=dev)
og = torch.empty(n_cg, 128, 128, 128, device=dev, dtype=torch.float16)


# convolutions that normalise their own output in the epilogue, at one bench chunk's shape: conv1 of a 128-wide block
# (accumulators wait in tensor memory) and conv2 (dual form: raw + residual and the normalised copy)
gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
scr = torch.zeros(n_cg * 128, device=dev, dtype=torch.int32)
gerr = torch.zeros(1, device=dev, dtype=torch.int32)
rawg = torch.empty_like(og)


def run():
    if part in ("convnorm",):
        _lib.call("rgm_conv_norm_f16", _lib.ptr(xg), _lib.ptr(wg), _lib.ptr(bg), _lib.ptr(gam), _lib.ptr(bet), None, None,
                  _lib.ptr(og), n_cg, 128, 128, cin, 128, 1, 1, _lib.ptr(scr), _lib.ptr(gerr), _lib.stream_ptr())
        _lib.call("rgm_conv_norm_f16", _lib.ptr(xg), _lib.ptr(wg), _lib.ptr(bg), _lib.ptr(gam), _lib.ptr(bet), _lib.ptr(xg),
                  _lib.ptr(rawg), _lib.ptr(og), n_cg, 128, 128, cin, 128, 1, 1, _lib.ptr(scr), _lib.ptr(gerr),
                  _lib.stream_ptr())
        _lib.call("rgm_conv_f16", _lib.ptr(xg), _lib.ptr(wg), _lib.ptr(bg), None, _lib.ptr(og), n_cg, 128, 128, cin, 128, 1, 0,
                  None, _lib.stream_ptr())
    if part in ("dit", "all"):
        model(x, t, y)
    if part in ("vae", "all"):
        roll = vae.decode_latents(lat, 1.2465, channels=1)
        music_rules.total_pitch_class_histogram(roll)
        music_rules.note_density(roll)
    if part in ("convgn",):
        _lib.call("rgm_conv_gn_f16", _lib.ptr(xg), _lib.ptr(abg), _lib.ptr(wg), _lib.ptr(bg), None, _lib.ptr(og), n_cg, 128,
                  128, cin, 128, None, _lib.stream_ptr())


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
