"""DiTRotary forward, torch fp32 on CPU, functional over a reference-keyed state dict.

Follows guided_diffusion/dit.py: FlattenPatchify1D :219-227, TimestepEmbedder :47-70, LabelEmbedder :95-100,
DiTBlockRotary :332-336, RotaryAttention :263-288, modulate :25-26, FinalLayerPatch1D :372-376, unpatchify :613-616,
DiTRotary.forward :618-634.  timm.Mlp (fc1 -> GELU(tanh) -> fc2) and rotary_embedding_torch.RotaryEmbedding
(interleaved-pair rotation of the first rot_dim features, positions 0..T-1) are restated from their published
behaviour.
"""
import math

import torch
import torch.nn.functional as F


def timestep_embedding(t, dim=256, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def rotate_queries_or_keys(x, freqs):
    """x [B, heads, T, hd]; freqs [rot_dim/2].  Rotates features [0, rot_dim) in interleaved pairs."""
    T = x.shape[-2]
    ang = torch.arange(T, dtype=freqs.dtype, device=freqs.device)[:, None] * freqs[None, :]  # [T, rot/2]
    ang = ang.repeat_interleave(2, dim=-1)                              # [T, rot]  (f0,f0,f1,f1,...)
    rot = ang.shape[-1]
    xr, xp = x[..., :rot], x[..., rot:]
    x2 = xr.reshape(*xr.shape[:-1], rot // 2, 2)
    half = torch.stack((-x2[..., 1], x2[..., 0]), dim=-1).reshape(xr.shape)
    return torch.cat((xr * ang.cos() + half * ang.sin(), xp), dim=-1)


def _ln(x):
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def _modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def _blocks(sd, h, sc, heads, depth, collect=None):
    """The DiTBlockRotary stack (dit.py:315-336) on tokens h [B,T,D] with the SiLU'd conditioning vector sc [B,D]."""
    B, T, hidden = h.shape
    hd = hidden // heads
    freqs = sd["rotary_emb.freqs"]
    for i in range(depth):
        p = f"blocks.{i}."
        mod = F.linear(sc, sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
        s1, c1, g1, s2, c2, g2 = mod.chunk(6, dim=1)
        a = _modulate(_ln(h), s1, c1)
        qkv = F.linear(a, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        qkv = qkv.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q = rotate_queries_or_keys(q, freqs)
        k = rotate_queries_or_keys(k, freqs)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(B, T, hidden)
        o = F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        h = h + g1.unsqueeze(1) * o
        m = _modulate(_ln(h), s2, c2)
        m = F.linear(m, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        m = F.linear(F.gelu(m, approximate="tanh"), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        h = h + g2.unsqueeze(1) * m
        if collect is not None:
            collect.append(h.clone())
    return h


def dit_forward(sd, x, t, y=None, *, heads, patch, collect=None):
    """x [B,C,H,W] fp32, t [B] (int or float), y [B] int or None -> [B,C_out,H,W]."""
    B, C, H, W = x.shape
    hidden = sd["x_embedder.MLP.2.weight"].shape[0]
    hd = hidden // heads
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    tok = x.permute(0, 2, 3, 1).reshape(B, H * W // patch, C * patch)
    h = F.linear(tok, sd["x_embedder.MLP.0.weight"], sd["x_embedder.MLP.0.bias"])
    h = F.linear(F.silu(h), sd["x_embedder.MLP.2.weight"], sd["x_embedder.MLP.2.bias"])
    c = F.linear(timestep_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    c = F.linear(F.silu(c), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    if y is not None and "y_embedder.embedding_table.weight" in sd:
        c = c + sd["y_embedder.embedding_table.weight"][y]
    sc = F.silu(c)
    h = _blocks(sd, h, sc, heads, depth, collect)
    mod = F.linear(sc, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=1)
    h = _modulate(_ln(h), shift, scale)
    h = F.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    c_out = h.shape[-1] // patch
    return h.reshape(B, -1, W, c_out).permute(0, 3, 1, 2).contiguous()


def classifier_forward(sd, x, t, *, heads, patch):
    """DiTRotaryClassifier.forward, chord=False (dit.py:801-831): tokens of FlattenPatchify1D with the class token
    prepended (T = H*W/patch + 1; the rotary positions run over all T tokens, class token = position 0), conditioning
    = t_embedder(t) alone, the same DiTBlockRotary stack, LayerNorm WITH affine (nn.LayerNorm default eps 1e-5) of the
    class token, classifier_head = Linear -> SiLU -> Linear.  x [B,C,H,W] fp32, t [B] -> logits [B, num_classes].
    Groundwork for SURVEY.md 8(f) rank 1 (classifier guidance natively): pinned against the unmodified reference
    (tests/golden/classifier.npz), forward and the input-gradient condition_functions.py:45-55 takes."""
    B, C, H, W = x.shape
    hidden = sd["x_embedder.MLP.2.weight"].shape[0]
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    tok = x.permute(0, 2, 3, 1).reshape(B, H * W // patch, C * patch)
    h = F.linear(tok, sd["x_embedder.MLP.0.weight"], sd["x_embedder.MLP.0.bias"])
    h = F.linear(F.silu(h), sd["x_embedder.MLP.2.weight"], sd["x_embedder.MLP.2.bias"])
    h = torch.cat((sd["cls_token"].expand(B, -1, -1), h), dim=1)
    c = F.linear(timestep_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    c = F.linear(F.silu(c), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    h = _blocks(sd, h, F.silu(c), heads, depth)
    z = F.layer_norm(h[:, 0, :], (hidden,), sd["norm.weight"], sd["norm.bias"], eps=1e-5)
    z = F.linear(z, sd["classifier_head.0.weight"], sd["classifier_head.0.bias"])
    return F.linear(F.silu(z), sd["classifier_head.2.weight"], sd["classifier_head.2.bias"])


def classifier_xentropy_grad(sd, x, labels, *, heads, patch):
    """condition_functions.grad_nn_zt_xentropy (:45-55): d/dx of sum_b log_softmax(classifier(x, t=0))[b, labels[b]]."""
    t = torch.zeros(x.shape[0])
    with torch.enable_grad():
        x_in = x.detach().requires_grad_(True)
        logits = classifier_forward(sd, x_in, t, heads=heads, patch=patch)
        sel = F.log_softmax(logits, dim=-1)[range(len(logits)), labels.view(-1)]
        return torch.autograd.grad(sel.sum(), x_in)[0]
