"""Restatement of rotary-embedding-torch 0.3.2's RotaryEmbedding as used by the reference (dit.py:15, 269-271,
571-572): `freqs` parameter = 1/10000^(arange(0,dim,2)/dim); rotate_queries_or_keys rotates the first `dim` features in
interleaved pairs by position * freqs, positions along dim -2.  Used ONLY to run the unmodified reference when
generating golden vectors; the package itself is not installed in this image."""
import torch
from torch import nn


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        self.freqs = nn.Parameter(1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim)),
                                  requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        seq_len = t.shape[seq_dim]
        pos = torch.arange(seq_len, device=t.device).type(self.freqs.dtype)
        freqs = torch.einsum("..., f -> ... f", pos, self.freqs)
        freqs = freqs.repeat_interleave(2, dim=-1)
        rot_dim = freqs.shape[-1]
        t_left, t_right = t[..., :rot_dim], t[..., rot_dim:]
        x = t_left.reshape(*t_left.shape[:-1], rot_dim // 2, 2)
        x1, x2 = x.unbind(dim=-1)
        rot_half = torch.stack((-x2, x1), dim=-1).reshape(t_left.shape)
        t_left = t_left * freqs.cos() + rot_half * freqs.sin()
        return torch.cat((t_left, t_right), dim=-1)
