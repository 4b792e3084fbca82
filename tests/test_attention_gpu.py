"""tcgen05 attention kernel vs torch SDPA in fp32 on the same fp16-rounded q, k, v.  P is rounded to fp16 before the
P.V product (like every flash-attention kernel), so the tolerance is 2e-3 of max|ref|."""
import pytest
import torch
import torch.nn.functional as F

from rule_guided_music_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,heads,T,dh", [(1, 1, 128, 64), (2, 3, 256, 72), (3, 16, 256, 72), (40, 16, 128, 72),
                                           (2, 2, 256, 128), (5, 6, 128, 64), (3, 4, 192, 72), (7, 2, 64, 64),
                                           (1, 16, 192, 72)])
def test_attention(cuda, B, heads, T, dh):
    g = torch.Generator(device="cpu").manual_seed(B * 100 + heads + T + dh)
    q = torch.randn(B, heads, T, dh, generator=g).to(cuda).half()
    k = torch.randn(B, heads, T, dh, generator=g).to(cuda).half()
    v = torch.randn(B, heads, T, dh, generator=g).to(cuda).half()
    vt = v.transpose(2, 3).contiguous()
    out = torch.empty(B * T, heads * dh, device=cuda, dtype=torch.float16)
    scale = dh ** -0.5
    _lib.call("rgm_attention_f16", _lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(out), B, heads, T, dh, scale,
              _lib.stream_ptr())
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float())
    ref = ref.transpose(1, 2).reshape(B * T, heads * dh)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, err
