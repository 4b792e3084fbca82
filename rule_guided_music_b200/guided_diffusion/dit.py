"""DiTRotary denoiser backed by librgm_b200.so -- host-side mirror of guided_diffusion/dit.py of the reference.

Same constructor arguments, registry names (``DiT_models``) and state-dict keys as the reference
(dit.py:538-576, 893-983; keys listed in SURVEY.md section 8b), so ``scripts/sample_rule.py:55-73`` works unchanged:
``DiT_models[name](input_size=[H, W], in_channels=..., num_classes=..., learn_sigma=...)``, ``load_state_dict(sd,
strict=False)``, ``.to(device)``, ``.eval()``, ``model(x, t, y)``.  The forward pass is entirely in CUDA
(csrc/dit.cu); there is no PyTorch fallback.  Only the rotary DiT family is on this path; the non-rotary DiT and the
classifier presets stay with the reference.
"""
import ctypes
import math

import torch

from .. import _lib


class DiTRotary:
    """Inference-only DiTRotary (reference dit.py:538-634)."""

    def __init__(self, input_size=32, patch_size=8, in_channels=3, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=9, learn_sigma=True):
        if isinstance(input_size, int):
            input_size = [input_size, input_size]
        self.input_size = list(input_size)
        self.patch_size = patch_size
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.learn_sigma = learn_sigma
        self.hidden_size = hidden_size
        self.depth = depth
        self.num_heads = num_heads
        self.num_classes = num_classes
        self.mlp_hidden = int(hidden_size * mlp_ratio)
        # LabelEmbedder keeps one extra row for the null class when class_dropout_prob > 0 (dit.py:79-80)
        self.label_rows = (num_classes + (1 if class_dropout_prob > 0 else 0)) if num_classes else 0
        self.training = False
        self._h = None
        self._device = None
        self._host_sd = {}
        self._loaded = set()

    # ---- nn.Module-like surface --------------------------------------------------------------------------------
    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.RgmError("rule_guided_music_b200.DiTRotary runs on a B200 only (no CPU path); got " + str(device))
        if self._h is None or device != self._device:
            self._create(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise _lib.RgmError("DiTRotary (B200 sampling path) is inference-only")
        return self

    def parameters(self):
        # the reference uses next(model.parameters()).device to find the device (gaussian_diffusion.py:836-837)
        if self._device is None:
            raise _lib.RgmError("DiTRotary: call .to(device) first")
        yield torch.empty(0, device=self._device)

    def convert_to_fp16(self):
        """scripts/sample_rule.py:75-76 calls this when use_fp16 is set; operands are already fp16 here."""
        return self

    def load_state_dict(self, state_dict, strict=True):
        """Accepts the reference's keys.  Returns (missing_keys, unexpected_keys) like torch."""
        unexpected = []
        for k, v in state_dict.items():
            self._host_sd[k] = v.detach().to(torch.float32)
        if self._h is not None:
            unexpected = self._push(self._host_sd)
        missing = sorted(set(self._required_keys()) - set(self._host_sd))
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]}..., unexpected {unexpected[:5]}...")
        return missing, unexpected

    def state_dict(self):
        return dict(self._host_sd)

    # ---- internals ------------------------------------------------------------------------------------------
    def _required_keys(self):
        keys = ["x_embedder.MLP.0", "x_embedder.MLP.2", "t_embedder.mlp.0", "t_embedder.mlp.2", "final_layer.linear",
                "final_layer.adaLN_modulation.1"]
        for i in range(self.depth):
            p = f"blocks.{i}."
            keys += [p + "attn.qkv", p + "attn.proj", p + "mlp.fc1", p + "mlp.fc2", p + "adaLN_modulation.1"]
        out = [k + s for k in keys for s in (".weight", ".bias")]
        if self.label_rows:
            out.append("y_embedder.embedding_table.weight")
        return out

    def _create(self, device):
        self._destroy()
        with torch.cuda.device(device):
            h = ctypes.c_void_p()
            _lib.call("rgm_dit_create", ctypes.byref(h), self.depth, self.hidden_size, self.num_heads, self.patch_size,
                      self.in_channels, self.out_channels, self.label_rows, self.input_size[1], self.mlp_hidden)
            self._h = h
            self._device = device
            # defaults the reference builds in its constructors rather than loads from a checkpoint
            half = 128  # TimestepEmbedder frequency_embedding_size = 256 (dit.py:37)
            freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
            rot = int(self.hidden_size // self.num_heads * 0.5)
            rope = 1.0 / (10000 ** (torch.arange(0, rot, 2).float() / rot))
            self._push({"__timestep_freqs": freqs, "rotary_emb.freqs": rope})
            if self._host_sd:
                self._push(self._host_sd)

    def _push(self, sd):
        unexpected = []
        with torch.cuda.device(self._device):
            stream = _lib.stream_ptr()
            for k, v in sd.items():
                t = v.to(self._device, torch.float32).contiguous()
                rc = _lib.lib().rgm_dit_load(self._h, k.encode(), _lib.ptr(t), t.numel(), stream)
                if rc < 0:
                    _lib.check(rc)
                if rc == 1:
                    unexpected.append(k)
            torch.cuda.current_stream().synchronize()  # the staging tensors above die with this scope
        return unexpected

    def set_lanes(self, lanes):
        """2 = overlap one sample chunk's attention / LayerNorm passes with the next chunk's GEMMs (default), 1 = serial."""
        _lib.call("rgm_dit_set_lanes", self._h, int(lanes))

    def _destroy(self):
        if self._h is not None:
            _lib.lib().rgm_dit_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # ---- forward ------------------------------------------------------------------------------------------------
    def forward(self, x, t, y=None):
        """x [N, C, H, W] fp32 cuda, t [N] (int or float), y [N] int or None -> [N, C_out, H, W]  (dit.py:618-634)."""
        if self._h is None:
            self.to(x.device)
        if x.device != self._device:
            raise _lib.RgmError(f"DiTRotary lives on {self._device}, input is on {x.device}")
        N, C, H, W = x.shape
        if C != self.in_channels or W != self.input_size[1]:
            raise _lib.RgmError(f"DiTRotary: expected [N,{self.in_channels},H,{self.input_size[1]}], got {tuple(x.shape)}")
        x = x.contiguous().float()
        tf = t.to(torch.float32).contiguous()
        yy = None
        if self.num_classes and y is not None:
            yy = y.to(torch.int64).contiguous()
        out = torch.empty(N, self.out_channels, H, W, device=x.device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            _lib.call("rgm_dit_forward", self._h, _lib.ptr(x), _lib.ptr(tf), _lib.ptr(yy), _lib.ptr(out), N, H,
                      _lib.stream_ptr())
        return out

    __call__ = forward


# ---- presets (reference dit.py:893-966) and registry (dit.py:969-983) -----------------------------------------------
def DiTRotary_XL_16(**kwargs):
    return DiTRotary(depth=28, hidden_size=1152, patch_size=16, num_heads=16, **kwargs)


def DiTRotary_XL_8(**kwargs):
    return DiTRotary(depth=28, hidden_size=1152, patch_size=8, num_heads=16, **kwargs)


def DiTRotary_B_16(**kwargs):
    return DiTRotary(depth=12, hidden_size=768, patch_size=16, num_heads=12, **kwargs)


def DiTRotary_B_8(**kwargs):
    return DiTRotary(depth=12, hidden_size=768, patch_size=8, num_heads=12, **kwargs)


DiT_models = {
    "DiTRotary_B_16": DiTRotary_B_16, "DiTRotary_B_8": DiTRotary_B_8,
    "DiTRotary_XL_16": DiTRotary_XL_16, "DiTRotary_XL_8": DiTRotary_XL_8,
}
