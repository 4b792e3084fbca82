#!/usr/bin/env python
"""CPU study: which fp16 operand roundings of the native DiT forward dominate its error against the fp32 reference.

Emulates csrc/dit.cu's precision choices inside the oracle's forward (oracle/dit.py): every GEMM operand (weights and
activations) rounded to fp16, accumulation / residual stream / LayerNorm / softmax statistics fp32, P stored
un-normalised in fp16.  Each rounding site can be switched off to measure its contribution on the 28-layer XL/8 golden
case (tests/golden/dit.npz, produced by the unmodified reference).  No GPU needed.

    python tools/cpu_precision_study.py            # table of rel-L2 errors, one site switched to fp32 at a time
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi  # noqa: E402
from oracle import dit as odit  # noqa: E402
from oracle import weights as ow  # noqa: E402

SITES = ["w_cond", "a_cond", "w_embed", "a_embed", "w_qkv", "a_ln1", "a_qk", "a_v", "a_p", "a_o", "w_proj", "w_fc1",
         "a_ln2", "a_h", "w_fc2", "w_fin", "a_fin"]


def h16(x, on=True):
    return x.half().float() if on else x


def split3(x):
    """fp16 hi + lo split: x ~ hi + lo with both representable in fp16 (what a 3-pass split GEMM would consume)."""
    hi = x.half().float()
    lo = (x - hi).half().float()
    return hi, lo


def lin(a, w, b, ra, rw, split=False):
    if split:  # hi*hi + hi*lo + lo*hi
        ah, al = split3(a)
        wh, wl = split3(w)
        return F.linear(ah, wh) + F.linear(ah, wl) + F.linear(al, wh) + b
    return F.linear(h16(a, ra), h16(w, rw), b)


def forward(sd, x, t, y, heads, patch, r, split_cond=False):
    B, C, H, W = x.shape
    hidden = sd["x_embedder.MLP.2.weight"].shape[0]
    hd = hidden // heads
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    tok = x.permute(0, 2, 3, 1).reshape(B, H * W // patch, C * patch)
    h = lin(tok, sd["x_embedder.MLP.0.weight"], sd["x_embedder.MLP.0.bias"], r["a_embed"], r["w_embed"])
    h = lin(F.silu(h), sd["x_embedder.MLP.2.weight"], sd["x_embedder.MLP.2.bias"], r["a_embed"], r["w_embed"])
    c = lin(odit.timestep_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"], r["a_cond"],
            r["w_cond"], split_cond)
    c = lin(F.silu(c), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"], r["a_cond"], r["w_cond"], split_cond)
    if y is not None:
        c = c + sd["y_embedder.embedding_table.weight"][y]
    sc = F.silu(c)
    freqs = sd["rotary_emb.freqs"]
    T = h.shape[1]
    scale = 1.0 / np.sqrt(hd)
    for i in range(depth):
        p = f"blocks.{i}."
        mod = lin(sc, sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"], r["a_cond"], r["w_cond"],
                  split_cond)
        s1, c1, g1, s2, c2, g2 = mod.chunk(6, dim=1)
        a = odit._modulate(odit._ln(h), s1, c1)
        qkv = lin(a, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"], r["a_ln1"], r["w_qkv"])
        qkv = qkv.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q = h16(odit.rotate_queries_or_keys(q, freqs), r["a_qk"])
        k = h16(odit.rotate_queries_or_keys(k, freqs), r["a_qk"])
        v = h16(v, r["a_v"])
        s = q @ k.transpose(-1, -2)
        e = torch.exp((s - s.amax(-1, keepdim=True)) * scale)
        o = (h16(e, r["a_p"]) @ v) / e.sum(-1, keepdim=True)
        o = o.transpose(1, 2).reshape(B, T, hidden)
        o = lin(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"], r["a_o"], r["w_proj"])
        h = h + g1.unsqueeze(1) * o
        m = odit._modulate(odit._ln(h), s2, c2)
        m = lin(m, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"], r["a_ln2"], r["w_fc1"])
        m = lin(F.gelu(m, approximate="tanh"), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"], r["a_h"], r["w_fc2"])
        h = h + g2.unsqueeze(1) * m
    mod = lin(sc, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"], r["a_cond"],
              r["w_cond"], split_cond)
    shift, scl = mod.chunk(2, dim=1)
    h = odit._modulate(odit._ln(h), shift, scl)
    h = lin(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"], r["a_fin"], r["w_fin"])
    c_out = h.shape[-1] // patch
    return h.reshape(B, -1, W, c_out).permute(0, 3, 1, 2).contiguous()


def main():
    torch.set_grad_enabled(False)
    tag = sys.argv[1] if len(sys.argv) > 1 else "xl8"
    cfg = gi.DIT_CASES[tag]
    w = cfg["weights"]
    sd = ow.make_dit_state_dict(**w)
    x, t, y = gi.dit_inputs(cfg)
    gold = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "dit.npz"))[tag])

    def err(r, **kw):
        out = forward(sd, x, t, y, w["heads"], w["patch"], r, **kw)
        return ((out.double() - gold.double()).norm() / gold.double().norm()).item()

    none = {s: False for s in SITES}
    allr = {s: True for s in SITES}
    print(f"{tag}: fp32 emulation vs reference golden      {err(none):.3e}")
    base = err(allr)
    print(f"{tag}: all sites fp16 (what dit.cu does)       {base:.3e}")
    print(f"{tag}: all fp16, conditioning path split-fp16  {err(allr, split_cond=True):.3e}")
    for s in SITES:
        r = dict(allr)
        r[s] = False
        e = err(r)
        print(f"  {s:8s} in fp32: {e:.3e}  (removes {max(base * base - e * e, 0) ** 0.5:.3e} in quadrature)")
    for s in SITES:
        r = dict(none)
        r[s] = True
        print(f"  only {s:8s} fp16: {err(r):.3e}")


if __name__ == "__main__":
    main()
