"""Deterministic synthetic weights with the reference's state-dict key names and shapes (benchmark / test data; not
part of the sampling path itself).

Checkpoints are not available offline, so every parity test and the benchmark use these.  A freshly constructed
reference DiT outputs exactly zero (adaLN-zero and a zero final layer, dit.py:597-606), which would hide every
kernel bug, so all adaLN / final-layer tensors are drawn non-zero here.  Generation is pure torch-CPU from a seed,
in a fixed key order, so the build container and the GPU box produce identical tensors.
Key names: SURVEY.md section 8b (probed from the reference modules).
"""
import math

import torch

DIT_PRESETS = {
    # name: (depth, hidden, patch, heads)      reference dit.py:893-966
    "DiTRotary_XL_8": (28, 1152, 8, 16),
    "DiTRotary_XL_16": (28, 1152, 16, 16),
    "DiTRotary_B_8": (12, 768, 8, 12),
    "DiTRotary_B_16": (12, 768, 16, 12),
}


def _xavier(g, out_f, in_f):
    a = math.sqrt(6.0 / (in_f + out_f))
    return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a


def make_dit_state_dict(seed=0, depth=28, hidden=1152, patch=8, heads=16, in_channels=4, num_classes=3,
                        learn_sigma=False, mlp_ratio=4.0, class_dropout=True, std_zero_init=0.02, bias_std=0.02):
    """State dict of DiTRotary (dit.py:538-606).  Linear weights Xavier-uniform like the reference init; the
    tensors the reference zero-inits are N(0, std_zero_init); biases N(0, bias_std) so bias paths are exercised."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    out_ch = in_channels * 2 if learn_sigma else in_channels
    hd = hidden // heads
    mlp = int(hidden * mlp_ratio)

    def lin(name, out_f, in_f, w_std=None):
        sd[name + ".weight"] = _xavier(g, out_f, in_f) if w_std is None else torch.randn(out_f, in_f, generator=g) * w_std
        sd[name + ".bias"] = torch.randn(out_f, generator=g) * bias_std

    lin("x_embedder.MLP.0", 256, in_channels * patch)
    lin("x_embedder.MLP.2", hidden, 256)
    lin("t_embedder.mlp.0", hidden, 256, w_std=0.02)
    lin("t_embedder.mlp.2", hidden, hidden, w_std=0.02)
    if num_classes:
        sd["y_embedder.embedding_table.weight"] = torch.randn(num_classes + (1 if class_dropout else 0), hidden,
                                                              generator=g) * 0.02
    rot = int(hd * 0.5)
    freqs = 1.0 / (10000 ** (torch.arange(0, rot, 2).float() / rot))
    sd["rotary_emb.freqs"] = freqs.clone()
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "attn.rotary_emb.freqs"] = freqs.clone()
        lin(p + "attn.qkv", 3 * hidden, hidden)
        lin(p + "attn.proj", hidden, hidden)
        lin(p + "mlp.fc1", mlp, hidden)
        lin(p + "mlp.fc2", hidden, mlp)
        lin(p + "adaLN_modulation.1", 6 * hidden, hidden, w_std=std_zero_init)
    lin("final_layer.linear", patch * out_ch, hidden, w_std=std_zero_init)
    lin("final_layer.adaLN_modulation.1", 2 * hidden, hidden, w_std=std_zero_init)
    return sd


VAE_DDCONFIG = dict(  # taming-transformers/configs/pr/kl/f8-all-onset.yaml:5-16
    double_z=True, z_channels=4, resolution=128, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 2, 4],
    num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def vae_decoder_layout(ch=128, ch_mult=(1, 2, 2, 4), num_res_blocks=2, z_channels=4, out_ch=3):
    """(name, kind, cin, cout) for every parameterised layer of Decoder + post_quant_conv, in forward order
    (model.py:436-537, klvae_pedal.py:80-85).  kind in {conv3, conv1, norm}."""
    L = []
    L.append(("post_quant_conv", "conv1", z_channels, z_channels))
    block_in = ch * ch_mult[-1]
    L.append(("decoder.conv_in", "conv3", z_channels, block_in))

    def res(prefix, cin, cout):
        L.append((prefix + ".norm1", "norm", cin, cin))
        L.append((prefix + ".conv1", "conv3", cin, cout))
        L.append((prefix + ".norm2", "norm", cout, cout))
        L.append((prefix + ".conv2", "conv3", cout, cout))
        if cin != cout:
            L.append((prefix + ".nin_shortcut", "conv1", cin, cout))

    res("decoder.mid.block_1", block_in, block_in)
    L.append(("decoder.mid.attn_1.norm", "norm", block_in, block_in))
    for n in ("q", "k", "v", "proj_out"):
        L.append((f"decoder.mid.attn_1.{n}", "conv1", block_in, block_in))
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(len(ch_mult))):
        block_out = ch * ch_mult[lvl]
        for b in range(num_res_blocks + 1):
            res(f"decoder.up.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            L.append((f"decoder.up.{lvl}.upsample.conv", "conv3", block_in, block_in))
    L.append(("decoder.norm_out", "norm", block_in, block_in))
    L.append(("decoder.conv_out", "conv3", block_in, out_ch))
    return L


def vae_encoder_layout(ch=128, ch_mult=(1, 2, 2, 4), num_res_blocks=2, z_channels=4, in_channels=3):
    """(name, kind, cin, cout) for every parameterised layer of Encoder + quant_conv, in forward order
    (model.py:342-433, klvae_pedal.py:60-63)."""
    L = [("encoder.conv_in", "conv3", in_channels, ch)]

    def res(prefix, cin, cout):
        L.append((prefix + ".norm1", "norm", cin, cin))
        L.append((prefix + ".conv1", "conv3", cin, cout))
        L.append((prefix + ".norm2", "norm", cout, cout))
        L.append((prefix + ".conv2", "conv3", cout, cout))
        if cin != cout:
            L.append((prefix + ".nin_shortcut", "conv1", cin, cout))

    block_in = ch
    for lvl in range(len(ch_mult)):
        block_out = ch * ch_mult[lvl]
        for b in range(num_res_blocks):
            res(f"encoder.down.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
        if lvl != len(ch_mult) - 1:
            L.append((f"encoder.down.{lvl}.downsample.conv", "conv3", block_in, block_in))
    res("encoder.mid.block_1", block_in, block_in)
    L.append(("encoder.mid.attn_1.norm", "norm", block_in, block_in))
    for n in ("q", "k", "v", "proj_out"):
        L.append((f"encoder.mid.attn_1.{n}", "conv1", block_in, block_in))
    res("encoder.mid.block_2", block_in, block_in)
    L.append(("encoder.norm_out", "norm", block_in, block_in))
    L.append(("encoder.conv_out", "conv3", block_in, 2 * z_channels))
    L.append(("quant_conv", "conv1", 2 * z_channels, 2 * z_channels))
    return L


def _fill(layout, g):
    sd = {}
    for name, kind, cin, cout in layout:
        if kind == "norm":
            sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(cin, generator=g)
            sd[name + ".bias"] = 0.1 * torch.randn(cin, generator=g)
        else:
            k = 3 if kind == "conv3" else 1
            bound = 1.0 / math.sqrt(cin * k * k)
            sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
            sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
    return sd


def make_vae_encoder_state_dict(seed=2, **cfg):
    """Encoder + quant_conv weights (separate generator from the decoder's, so the decoder tensors and every golden
    vector made from them are unchanged)."""
    return _fill(vae_encoder_layout(**cfg), torch.Generator(device="cpu").manual_seed(seed))


def make_vae_state_dict(seed=1, **cfg):
    """Decoder + post_quant_conv weights, PyTorch-default-like scale (uniform +-1/sqrt(fan_in)), GroupNorm affine
    perturbed away from (1, 0) so the affine path is exercised."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for name, kind, cin, cout in vae_decoder_layout(**cfg):
        if kind == "norm":
            sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(cin, generator=g)
            sd[name + ".bias"] = 0.1 * torch.randn(cin, generator=g)
        else:
            k = 3 if kind == "conv3" else 1
            bound = 1.0 / math.sqrt(cin * k * k)
            sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
            sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
    return sd


def make_classifier_state_dict(seed=0, depth=4, hidden=384, patch=8, heads=6, in_channels=4, num_classes=9,
                               mlp_ratio=4.0, std_zero_init=0.02, bias_std=0.02):
    """State dict of DiTRotaryClassifier, chord=False (dit.py:735-800; `DiTRotary-XS/8-cls` by default).  Same init
    policy as make_dit_state_dict: Xavier-uniform linears, the tensors the reference zero-inits drawn N(0, 0.02) so
    that the blocks are not the identity, small non-zero biases, a visible class token."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    hd = hidden // heads
    mlp = int(hidden * mlp_ratio)

    def lin(name, out_f, in_f, w_std=None):
        sd[name + ".weight"] = _xavier(g, out_f, in_f) if w_std is None else torch.randn(out_f, in_f, generator=g) * w_std
        sd[name + ".bias"] = torch.randn(out_f, generator=g) * bias_std

    lin("x_embedder.MLP.0", 256, in_channels * patch)
    lin("x_embedder.MLP.2", hidden, 256)
    lin("t_embedder.mlp.0", hidden, 256, w_std=0.02)
    lin("t_embedder.mlp.2", hidden, hidden, w_std=0.02)
    rot = int(hd * 0.5)
    freqs = 1.0 / (10000 ** (torch.arange(0, rot, 2).float() / rot))
    sd["rotary_emb.freqs"] = freqs.clone()
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "attn.rotary_emb.freqs"] = freqs.clone()
        lin(p + "attn.qkv", 3 * hidden, hidden)
        lin(p + "attn.proj", hidden, hidden)
        lin(p + "mlp.fc1", mlp, hidden)
        lin(p + "mlp.fc2", hidden, mlp)
        lin(p + "adaLN_modulation.1", 6 * hidden, hidden, w_std=std_zero_init)
    sd["cls_token"] = torch.randn(1, 1, hidden, generator=g) * 0.5
    sd["norm.weight"] = 1.0 + 0.1 * torch.randn(hidden, generator=g)
    sd["norm.bias"] = 0.05 * torch.randn(hidden, generator=g)
    lin("classifier_head.0", hidden // 4, hidden)
    lin("classifier_head.2", num_classes, hidden // 4)
    return sd
