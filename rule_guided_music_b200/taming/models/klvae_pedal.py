"""AutoencoderKL (decoder side) backed by librgm_b200.so -- mirror of taming/models/klvae_pedal.py:13-85.

Only what the sampling path calls is here: construction from the reference's ``ddconfig`` (the decoder fields of
taming-transformers/configs/pr/kl/f8-all-onset.yaml:5-16), ``init_from_ckpt`` / ``load_state_dict`` with the
reference's checkpoint keys (``decoder.*``, ``post_quant_conv.*``; encoder/loss keys are ignored like strict=False),
``.to(device)``, ``.eval()``, ``decode(z)`` and -- for scripts/edit.py, which encodes the ground-truth roll once per
run through gaussian_diffusion._encode -- ``encode_save(x)`` / ``encode(x)`` (klvae_pedal.py:60-78, checkpoint keys
``encoder.*``, ``quant_conv.*``).  ``decode_latents`` is the fused form of gaussian_diffusion._decode (re-tiling +
decode + roll assembly) that the native sampler uses.
"""
import ctypes

import torch

from ... import _lib

DEFAULT_DDCONFIG = dict(double_z=True, z_channels=4, resolution=128, in_channels=3, out_ch=3, ch=128,
                        ch_mult=[1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)


class AutoencoderKL:
    def __init__(self, ddconfig=None, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), **_unused):
        cfg = dict(DEFAULT_DDCONFIG)
        cfg.update(ddconfig or {})
        if list(cfg.get("attn_resolutions", [])):
            raise _lib.RgmError("AutoencoderKL (B200 path): attn_resolutions must be [] (mid attention only)")
        if cfg["z_channels"] != embed_dim:
            raise _lib.RgmError("AutoencoderKL (B200 path): embed_dim must equal z_channels")
        self.ddconfig = cfg
        self.embed_dim = embed_dim
        self._h = None
        self._device = None
        self._host_sd = {}
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    # ---- loading (klvae_pedal.py:50-59) -------------------------------------------------------------------------
    def init_from_ckpt(self, path, ignore_keys=()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        sd = {k: v for k, v in sd.items() if not any(k.startswith(ik) for ik in ignore_keys)}
        self.load_state_dict(sd, strict=False)

    def load_state_dict(self, state_dict, strict=True):
        for k, v in state_dict.items():
            if k.startswith(("decoder.", "post_quant_conv.", "encoder.", "quant_conv.")):
                self._host_sd[k] = v.detach().to(torch.float32)
        unexpected = self._push(self._host_sd) if self._h is not None else []
        return [], unexpected

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.RgmError("rule_guided_music_b200.AutoencoderKL runs on a B200 only (no CPU path)")
        if self._h is None or device != self._device:
            self._create(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def eval(self):
        return self

    def parameters(self):
        if self._device is None:
            raise _lib.RgmError("AutoencoderKL: call .to(device) first")
        yield torch.empty(0, device=self._device)

    def _create(self, device):
        self._destroy()
        c = self.ddconfig
        mult = (ctypes.c_int * len(c["ch_mult"]))(*c["ch_mult"])
        with torch.cuda.device(device):
            h = ctypes.c_void_p()
            _lib.call("rgm_vae_create", ctypes.byref(h), c["ch"], mult, len(c["ch_mult"]), c["num_res_blocks"],
                      c["z_channels"], c["out_ch"])
        self._h, self._device = h, device
        if self._host_sd:
            self._push(self._host_sd)

    def _push(self, sd):
        unexpected = []
        with torch.cuda.device(self._device):
            stream = _lib.stream_ptr()
            for k, v in sd.items():
                t = v.to(self._device, torch.float32).contiguous()
                rc = _lib.lib().rgm_vae_load(self._h, k.encode(), _lib.ptr(t), t.numel(), stream)
                if rc < 0:
                    _lib.check(rc)
                if rc == 1:
                    unexpected.append(k)
            torch.cuda.current_stream().synchronize()
        return unexpected

    def set_lanes(self, lanes):
        """2 = overlap GroupNorm passes of one tile chunk with the convolutions of the next (default), 1 = serial."""
        _lib.call("rgm_vae_set_lanes", self._h, int(lanes))

    def gn_timeouts(self):
        """Diagnostic (rgm_vae_gn_timeouts): 1 if a convolution that normalises its own output ever gave up waiting for
        its image's other tiles -- never in a healthy run.  Synchronises the device."""
        if self._h is None:
            return 0
        with torch.cuda.device(self._device):
            rc = _lib.lib().rgm_vae_gn_timeouts(self._h)
        if rc < 0:
            _lib.check(rc)
        return rc

    def _destroy(self):
        if self._h is not None:
            _lib.lib().rgm_vae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # ---- compute ------------------------------------------------------------------------------------------------
    def decode_latents(self, latents, scale_factor=1.0, channels=None):
        """gaussian_diffusion._decode fused: latents [n, 4, H, 16] (H multiple of 16) -> roll [n, ch, 128, 8H]."""
        if self._h is None:
            self.to(latents.device)
        n, c, H, W = latents.shape
        if c != self.ddconfig["z_channels"] or W != 16:
            raise _lib.RgmError(f"decode_latents: expected [n,{self.ddconfig['z_channels']},H,16], got {tuple(latents.shape)}")
        ch = self.ddconfig["out_ch"] if channels is None else channels
        lat = latents.contiguous().float()
        roll = torch.empty(n, ch, 128, 8 * H, device=lat.device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            _lib.call("rgm_vae_decode_latents", self._h, _lib.ptr(lat), float(scale_factor), _lib.ptr(roll), n, H, ch,
                      _lib.stream_ptr())
        return roll

    def decode(self, z):
        """z [n, 4, 16, 16] -> [n, out_ch, 128, 128]   (klvae_pedal.py:80-85)."""
        if z.shape[-1] != 16 or z.shape[-2] != 16:
            raise _lib.RgmError("AutoencoderKL.decode (B200 path) takes 16x16 latent tiles")
        # a tile is [pitch, time]; decode_latents takes [time, pitch] like the sampler's latents
        return self.decode_latents(z.permute(0, 1, 3, 2), 1.0)

    def encode_save(self, x, range_fix=False):
        """x [n, 3, 128, 128] in [-1, 1] -> moments [n, 2*z_channels, 16, 16] (mean | logvar)  (klvae_pedal.py:60-68)."""
        if self._h is None:
            self.to(x.device)
        n, c, H, W = x.shape
        if c != self.ddconfig["in_channels"] or H != 128 or W != 128:
            raise _lib.RgmError(f"encode_save: expected [n,{self.ddconfig['in_channels']},128,128], got {tuple(x.shape)}")
        if not any(k.startswith("encoder.") for k in self._host_sd):
            raise _lib.RgmError("encode_save: no encoder.* weights were loaded into this AutoencoderKL")
        xin = x.contiguous().float()
        zz = 2 * self.ddconfig["z_channels"]
        moments = torch.empty(n, zz, 16, 16, device=xin.device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            _lib.call("rgm_vae_encode", self._h, _lib.ptr(xin), _lib.ptr(moments), n, _lib.stream_ptr())
        if range_fix:
            mean, logvar = torch.chunk(moments, 2, dim=1)
            moments = torch.concat((torch.sigmoid(mean) * 2 - 1, logvar), dim=1)
        return moments

    def encode(self, x, range_fix=False):
        """-> DiagonalGaussianDistribution over the latent (klvae_pedal.py:70-78)."""
        return DiagonalGaussianDistribution(self.encode_save(x, range_fix=range_fix))


class DiagonalGaussianDistribution:
    """taming/modules/distributions/distributions.py:24-62 (the parts sampling uses: mean, logvar clamp, sample, mode)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.parameters.device)

    def mode(self):
        return self.mean
