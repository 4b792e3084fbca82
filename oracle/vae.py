"""taming KL-VAE decoder (post_quant_conv + Decoder) and encoder (Encoder + quant_conv), torch fp32 on CPU, functional
over a reference-keyed state dict.

Follows taming/models/klvae_pedal.py:80-85 and taming/modules/diffusionmodules/model.py: nonlinearity :29-31,
Normalize :34-35 (GroupNorm 32 groups, eps 1e-6), Upsample :49-53, ResnetBlock :117-137, AttnBlock :168-192,
Decoder.forward :506-537; and the latent re-tiling of gaussian_diffusion.py:1347-1358 (_decode).
Encoder side: Downsample :55-75 (pad right/bottom by one, 3x3 stride 2), Encoder.forward :403-433,
AutoencoderKL.encode_save klvae_pedal.py:60-68, and the roll re-tiling of gaussian_diffusion.py:1382-1395 (_encode).
"""
import torch
import torch.nn.functional as F


def _gn(sd, name, x):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(sd, name, x, pad):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=pad)


def _res(sd, p, x):
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), 1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), 1)
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, 0)
    return x + h


def _attn(sd, p, x):
    h = _gn(sd, p + ".norm", x)
    q, k, v = _conv(sd, p + ".q", h, 0), _conv(sd, p + ".k", h, 0), _conv(sd, p + ".v", h, 0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w = torch.bmm(q, k) * (int(c) ** (-0.5))
    w = F.softmax(w, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", h, 0)


def vae_decode(sd, z, num_levels=4, num_res_blocks=2, collect=None):
    """z [n,4,16,16] -> [n,3,128,128]   (AutoencoderKL.decode)."""
    h = _conv(sd, "post_quant_conv", z, 0)
    h = _conv(sd, "decoder.conv_in", h, 1)
    h = _res(sd, "decoder.mid.block_1", h)
    h = _attn(sd, "decoder.mid.attn_1", h)
    h = _res(sd, "decoder.mid.block_2", h)
    if collect is not None:
        collect.append(("mid", h.clone()))
    for lvl in reversed(range(num_levels)):
        for b in range(num_res_blocks + 1):
            h = _res(sd, f"decoder.up.{lvl}.block.{b}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", h, 1)
        if collect is not None:
            collect.append((f"up{lvl}", h.clone()))
    h = _swish(_gn(sd, "decoder.norm_out", h))
    return _conv(sd, "decoder.conv_out", h, 1)


def decode_latents(sd, pred_zstart, scale_factor=1.0, threshold=False):
    """gaussian_diffusion.py:1347-1358: latent [B,4,H,16] -> piano roll [B,3,128,8H], tiles batched tile-major."""
    H, W = pred_zstart.shape[-2], pred_zstart.shape[-1]
    s = (pred_zstart / scale_factor).permute(0, 1, 3, 2)
    s = torch.cat(torch.chunk(s, H // W, dim=-1), dim=0)
    s = vae_decode(sd, s)
    roll = torch.cat(torch.chunk(s, H // W, dim=0), dim=-1)
    if threshold:
        roll[roll <= -0.95] = -1.0
    return roll


def vae_encode(sd, x, num_levels=4, num_res_blocks=2):
    """x [n,3,128,128] -> moments [n,8,16,16]   (AutoencoderKL.encode_save, range_fix=False)."""
    h = _conv(sd, "encoder.conv_in", x, 1)
    for lvl in range(num_levels):
        for b in range(num_res_blocks):
            h = _res(sd, f"encoder.down.{lvl}.block.{b}", h)
        if lvl != num_levels - 1:
            p = f"encoder.down.{lvl}.downsample.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1), mode="constant", value=0), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    h = _res(sd, "encoder.mid.block_1", h)
    h = _attn(sd, "encoder.mid.attn_1", h)
    h = _res(sd, "encoder.mid.block_2", h)
    h = _swish(_gn(sd, "encoder.norm_out", h))
    h = _conv(sd, "encoder.conv_out", h, 1)
    return _conv(sd, "quant_conv", h, 0)


def encode_rolls(sd, roll, scale_factor=1.0):
    """gaussian_diffusion.py:1382-1395: piano roll [B,3,128,L] -> latent [B,4,L/8,16] (posterior mean * scale)."""
    H, W = roll.shape[-2], roll.shape[-1]
    seq = W // H
    micro = torch.cat(torch.chunk(roll, seq, dim=-1), dim=0)
    micro = vae_encode(sd, micro)
    z = torch.chunk(micro, 2, dim=1)[0] if micro.shape[1] == 8 else micro
    z = torch.cat(torch.chunk(z, seq, dim=0), dim=-1)
    return z.permute(0, 1, 3, 2) * scale_factor
