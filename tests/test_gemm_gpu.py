"""Parity of the tcgen05 implicit-GEMM kernel (linear / conv3x3 / conv1x1 / upsample-conv) against torch fp32 on the
same fp16-rounded operands.  Tolerances: the kernel accumulates in fp32 like the reference math, so the only
differences are summation order (and, for the upsample conv, pre-summed fp16 weights): 2e-3 * max|ref| absolute."""
import pytest
import torch
import torch.nn.functional as F

from rule_guided_music_b200 import _lib

pytestmark = pytest.mark.gpu


def _gemm(a16, b16, bias, block_n=0):
    M, K = a16.shape
    N = b16.shape[0]
    out = torch.empty(M, N, device=a16.device, dtype=torch.float32)
    _lib.call("rgm_gemm_f16", _lib.ptr(a16), _lib.ptr(b16), _lib.ptr(bias), _lib.ptr(out), M, N, K, block_n,
              _lib.stream_ptr())
    return out


@pytest.mark.parametrize("M,N,K,bn", [
    (256, 128, 64, 128),
    (128, 128, 256, 128),
    (1000, 384, 1152, 128),
    (4096, 512, 512, 256),
    (512, 32, 1152, 32),
    (300, 256, 4608, 0),
    (33000, 1152, 1152, 0),
])
def test_linear(cuda, M, N, K, bn):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda).half()
    b = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda).half()
    bias = torch.randn(N, generator=g).to(cuda)
    out = _gemm(a, b, bias, bn)
    ref = a.float() @ b.float().t() + bias
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item(), err


def _pack(w, kind, cin_pad=None, cout_pad=None):
    cout, cin = w.shape[:2]
    cin_pad = cin_pad or cin
    cout_pad = cout_pad or cout
    taps = {0: 1, 1: 9, 2: 4, 3: 9}[kind]
    npar = 4 if kind == 2 else 1
    out = torch.empty(npar * cout_pad * taps * cin_pad, device=w.device, dtype=torch.float16)
    _lib.call("rgm_pack_conv_weight", _lib.ptr(w.contiguous()), _lib.ptr(out), cout, cin, cout_pad, cin_pad, kind,
              _lib.stream_ptr())
    return out


def _conv(x_nhwc16, wp, bias, cout, kind, resid=None, bn=0, gn_part=None):
    n, H, W, cin = x_nhwc16.shape
    Ho, Wo = (H * 2, W * 2) if kind == 2 else ((H // 2, W // 2) if kind == 3 else (H, W))
    out = torch.empty(n, Ho, Wo, cout, device=x_nhwc16.device, dtype=torch.float16)
    _lib.call("rgm_conv_f16", _lib.ptr(x_nhwc16), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(resid), _lib.ptr(out), n, H,
              W, cin, cout, kind, bn, _lib.ptr(gn_part), _lib.stream_ptr())
    return out


@pytest.mark.parametrize("n,H,cin,cout,kind", [
    (3, 16, 64, 128, 1),
    (2, 32, 128, 256, 1),
    (1, 128, 64, 128, 1),
    (2, 64, 256, 256, 1),
    (5, 16, 512, 512, 0),
    (2, 16, 64, 128, 2),
    (2, 64, 128, 256, 2),
    (3, 32, 256, 128, 0),
    (2, 128, 128, 128, 3),   # Downsample (stride 2, pad right/bottom): the three encoder shapes
    (3, 64, 256, 256, 3),
    (5, 32, 256, 256, 3),
])
def test_conv(cuda, n, H, cin, cout, kind):
    g = torch.Generator(device="cpu").manual_seed(n * 1000 + H + cin + cout + kind)
    x = torch.randn(n, cin, H, H, generator=g).to(cuda).half()
    k = 1 if kind == 0 else 3
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    wp = _pack(w, kind)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    out = _conv(x_nhwc, wp, bias, cout, kind).float().permute(0, 3, 1, 2)
    xin = x.float()
    if kind == 2:
        xin = F.interpolate(xin, scale_factor=2.0, mode="nearest")
    if kind == 3:
        ref = F.conv2d(F.pad(xin, (0, 1, 0, 1)), w.half().float(), bias, stride=2)
    else:
        ref = F.conv2d(xin, w.half().float(), bias, padding=k // 2)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    tol = (4e-3 if kind == 2 else 2e-3) * ref.abs().max().item()
    assert err <= tol, (err, tol)


def test_conv_resid_and_gn_partials(cuda):
    n, H, cin, cout = 2, 32, 128, 128
    g = torch.Generator(device="cpu").manual_seed(7)
    x = torch.randn(n, H, H, cin, generator=g).to(cuda).half()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    resid = torch.randn(n, H, H, cout, generator=g).to(cuda).half()
    part = torch.zeros(n * H * H // 128, cout // 4, 2, device=cuda)
    out = _conv(x, _pack(w, 1), bias, cout, 1, resid=resid, gn_part=part)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), bias, padding=1).permute(0, 2, 3, 1) + resid.float()
    torch.cuda.synchronize()
    assert (out.float() - ref).abs().max().item() <= 4e-3 * ref.abs().max().item()
    o = out.float().reshape(n * H * H // 128, 128, cout // 4, 4)  # partial sums per (128 rows x 4 channels)
    assert torch.allclose(part[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=3e-2)
    assert torch.allclose(part[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=3e-2)


@pytest.mark.parametrize("n,H,cin,resid", [(1, 128, 128, False), (3, 128, 128, True), (2, 128, 256, False),
                                            (150, 128, 128, True), (2, 6, 64, False)])
def test_conv_with_fused_groupnorm_input(cuda, n, H, cin, resid):
    """conv3x3(swish(a*x + b)) in one kernel (csrc/conv_gn.cuh: halo tile normalised in shared memory, nine taps as
    shifted descriptor views) against (i) the two-pass form through the same library -- rgm_gn_apply_f16 then
    rgm_conv_f16: identical operand values, only the accumulation order differs -- and (ii) torch in fp32."""
    W, cout = 128, 128
    g = torch.Generator(device="cpu").manual_seed(n * 31 + H + cin)
    x = (torch.randn(n, H, W, cin, generator=g) * 1.5 + 0.3).to(cuda).half()
    ab = torch.stack((torch.rand(n, cin, generator=g) + 0.5, torch.randn(n, cin, generator=g) * 0.5), dim=-1).to(cuda)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    res = torch.randn(n, H, W, cout, generator=g).to(cuda).half() if resid else None
    wp = _pack(w, 1)
    part = torch.zeros(n * H * W // 128, cout // 4, 2, device=cuda)
    out = torch.empty(n, H, W, cout, device=cuda, dtype=torch.float16)
    _lib.call("rgm_conv_gn_f16", _lib.ptr(x), _lib.ptr(ab), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out),
              n, H, W, cin, cout, _lib.ptr(part), _lib.stream_ptr())
    y = torch.empty_like(x)
    _lib.call("rgm_gn_apply_f16", _lib.ptr(x), _lib.ptr(ab), _lib.ptr(y), n, H * W, cin, 1, _lib.stream_ptr())
    two_pass = torch.empty_like(out)
    _lib.call("rgm_conv_f16", _lib.ptr(y), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(two_pass), n, H, W,
              cin, cout, 1, 0, None, _lib.stream_ptr())
    torch.cuda.synchronize()
    # (i) same operands, different fp32 accumulation order: at most one fp16 ulp on a few outputs
    d = (out.float() - two_pass.float()).abs()
    scale = two_pass.float().abs().max().item()
    assert d.max().item() <= 2e-3 * scale, (d.max().item(), scale)
    assert (d > 0).float().mean().item() < 0.05
    # (ii) the math itself
    act = torch.nn.functional.silu(x.float() * ab[:, None, None, :, 0] + ab[:, None, None, :, 1]).half().float()
    ref = F.conv2d(act.permute(0, 3, 1, 2), w.half().float(), bias, padding=1).permute(0, 2, 3, 1)
    if resid:
        ref = ref + res.float()
    assert (out.float() - ref).abs().max().item() <= 4e-3 * ref.abs().max().item()
    # partial sums are taken of the fp32 values before the fp16 rounding of the store: 512 roundings of 2^-11 |v| apart
    o = out.float().reshape(n * H * W // 128, 128, cout // 4, 4)
    assert torch.allclose(part[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=0.15)


@pytest.mark.parametrize("n,H,cin,cout,kind,swish", [
    (3, 32, 128, 128, 1, 1),     # single-CTA kernel, 4 tiles per image
    (3, 128, 128, 128, 1, 1),    # 64 tiles per image, 192 tiles: the third image straddles two waves of the grid
    (40, 16, 256, 512, 1, 1),    # CTA pairs, 2 feature pairs per image, groups of 16 channels (4 quads)
    (10, 64, 256, 256, 1, 1),    # CTA pairs, 16 row tiles per image, groups of 8 channels; images straddle waves
    (130, 128, 128, 128, 1, 1),  # a full bench chunk of the 128x128 level: 8320 tiles, 57 waves
    (80, 16, 128, 256, 0, 0),    # 1x1 conv on CTA pairs, one tile per image, norm without swish
])
def test_conv_with_groupnorm_of_its_own_output(cuda, n, H, cin, cout, kind, swish):
    """swish(GroupNorm(conv(x))) with the normalisation applied inside the convolution's epilogue (gemm_tc.cuh
    gn_epilogue_loop: statistics exchanged between the CTAs of an image while the accumulators wait in tensor
    memory) against torch fp32 on the same fp16-rounded operands, and against the two-pass form of this library
    (rgm_conv_f16 + torch statistics + rgm_gn_apply_f16), which differs only by the fp16 rounding of the stored raw
    tensor.  The wait must never give up."""
    g = torch.Generator(device="cpu").manual_seed(n * 977 + H + cin + cout + kind)
    k = 1 if kind == 0 else 3
    x = (torch.randn(n, H, H, cin, generator=g) * 1.3 + 0.2).to(cuda).half()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    gamma = (torch.rand(cout, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(cout, generator=g) * 0.3).to(cuda)
    wp = _pack(w, kind)
    scratch = torch.full((n * 128,), 12345, device=cuda, dtype=torch.int32)  # n * 512 bytes; the call must zero it
    err = torch.zeros(1, device=cuda, dtype=torch.int32)
    out = torch.empty(n, H, H, cout, device=cuda, dtype=torch.float16)
    for _ in range(2):  # twice: accumulators and counters are re-armed by every call
        _lib.call("rgm_conv_norm_f16", _lib.ptr(x), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta),
                  None, None, _lib.ptr(out), n, H, H, cin, cout, kind, swish, _lib.ptr(scratch), _lib.ptr(err),
                  _lib.stream_ptr())
    torch.cuda.synchronize()
    assert err.item() == 0, "a GroupNorm-in-epilogue wait gave up"
    count = scratch.view(n, 64, 2)[:, :, 0] & 255   # low byte of every accumulator word = contributions received
    assert int(count.min()) == int(count.max()) == 2 * (H * H // 256)
    conv = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), bias, padding=k // 2)
    ref = F.group_norm(conv, 32, gamma, beta, eps=1e-6)
    if swish:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    e = (out.float() - ref).abs().max().item()
    assert e <= 2e-3 * ref.abs().max().item() + 1e-3, e
    assert torch.isfinite(out.float()).all()
    # the two-pass form: raw fp16 tensor, statistics of the fp32 result, normalise pass
    raw = _conv(x, wp, bias, cout, kind)
    cpg = cout // 32
    c32 = conv.permute(0, 2, 3, 1).reshape(n, H * H, 32, cpg)
    mean = c32.mean(dim=(1, 3))
    rstd = (c32.var(dim=(1, 3), unbiased=False) + 1e-6).rsqrt()
    a = rstd.repeat_interleave(cpg, dim=1) * gamma
    b = beta - mean.repeat_interleave(cpg, dim=1) * a
    ab = torch.stack((a, b), dim=-1).contiguous()
    two = torch.empty_like(out)
    _lib.call("rgm_gn_apply_f16", _lib.ptr(raw), _lib.ptr(ab), _lib.ptr(two), n, H * H, cout, swish, _lib.stream_ptr())
    torch.cuda.synchronize()
    d = (out.float() - two.float()).abs().max().item()
    assert d <= 4e-3 * two.float().abs().max().item() + 1e-3, d


@pytest.mark.parametrize("n,H,cin,cout,resid", [
    (3, 32, 128, 128, "separate"),
    (130, 128, 128, 128, "separate"),  # a full bench chunk of the 128x128 level
    (10, 64, 256, 256, "in_place"),    # CTA pairs; the shortcut is found in the output buffer (nin_shortcut case)
    (40, 16, 512, 512, None),
])
def test_conv_dual_raw_and_normalised_output(cuda, n, H, cin, cout, resid):
    """conv2 of a ResnetBlock: raw = conv(x) + shortcut AND swish(GroupNorm(raw)) from one launch (gemm_tc.cuh
    gn_dual_loop).  The raw tensor must equal the plain convolution's bit for bit; the normalised copy is compared with
    the library's own normalise pass on that raw tensor (same fp16 inputs, statistics of the fp32 values) and torch."""
    g = torch.Generator(device="cpu").manual_seed(n * 131 + H + cin + cout)
    x = (torch.randn(n, H, H, cin, generator=g) * 1.2).to(cuda).half()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    gamma = (torch.rand(cout, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(cout, generator=g) * 0.3).to(cuda)
    res = torch.randn(n, H, H, cout, generator=g).to(cuda).half() if resid else None
    wp = _pack(w, 1)
    plain = torch.empty(n, H, H, cout, device=cuda, dtype=torch.float16)
    if resid == "in_place":
        plain.copy_(res)
        _lib.call("rgm_conv_f16", _lib.ptr(x), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(plain), _lib.ptr(plain), n, H, H,
                  cin, cout, 1, 0, None, _lib.stream_ptr())
    else:
        plain = _conv(x, wp, bias, cout, 1, resid=res)
    scratch = torch.full((n * 128,), -1, device=cuda, dtype=torch.int32)
    err = torch.zeros(1, device=cuda, dtype=torch.int32)
    raw = torch.empty_like(plain)
    out = torch.empty_like(plain)
    rp = res
    if resid == "in_place":
        raw.copy_(res)
        rp = raw
    _lib.call("rgm_conv_norm_f16", _lib.ptr(x), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta),
              _lib.ptr(rp), _lib.ptr(raw), _lib.ptr(out), n, H, H, cin, cout, 1, 1, _lib.ptr(scratch), _lib.ptr(err),
              _lib.stream_ptr())
    torch.cuda.synchronize()
    assert err.item() == 0, "a GroupNorm-in-epilogue wait gave up"
    assert torch.equal(raw, plain)
    conv = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), bias, padding=1)
    if res is not None:
        conv = conv + res.float().permute(0, 3, 1, 2)
    cpg = cout // 32
    c32 = conv.permute(0, 2, 3, 1).reshape(n, H * H, 32, cpg)
    mean = c32.mean(dim=(1, 3))
    rstd = (c32.var(dim=(1, 3), unbiased=False) + 1e-6).rsqrt()
    a = rstd.repeat_interleave(cpg, dim=1) * gamma
    b = beta - mean.repeat_interleave(cpg, dim=1) * a
    two = torch.empty_like(out)
    ab = torch.stack((a, b), dim=-1).contiguous()
    _lib.call("rgm_gn_apply_f16", _lib.ptr(raw), _lib.ptr(ab), _lib.ptr(two), n, H * H, cout, 1, _lib.stream_ptr())
    torch.cuda.synchronize()
    d = (out.float() - two.float()).abs()
    assert d.max().item() <= 2e-3 * two.float().abs().max().item() + 1e-3, d.max().item()
    assert (d > 0).float().mean().item() < 0.02   # same inputs, statistics equal to ~1e-6: rare last-bit differences
    ref = F.silu(F.group_norm(conv, 32, gamma, beta, eps=1e-6)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 3e-3 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("n,H,c", [(6, 32, 256), (20, 64, 256)])
def test_upsample_conv_dual_raw_and_normalised_output(cuda, n, H, c):
    """The nearest-2x-upsample + 3x3 conv (four parity sub-convolutions) in the dual form: raw output bit-identical to
    the plain launch, normalised copy = this library's normalise pass on it up to rare last-bit differences."""
    g = torch.Generator(device="cpu").manual_seed(n + H + c)
    x = (torch.randn(n, H, H, c, generator=g) * 1.1).to(cuda).half()
    w = (torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5).to(cuda)
    bias = torch.randn(c, generator=g).to(cuda)
    gamma = (torch.rand(c, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(c, generator=g) * 0.3).to(cuda)
    wp = _pack(w, 2)
    plain = _conv(x, wp, bias, c, 2)
    scratch = torch.full((n * 128,), -1, device=cuda, dtype=torch.int32)
    err = torch.zeros(1, device=cuda, dtype=torch.int32)
    raw = torch.empty_like(plain)
    out = torch.empty_like(plain)
    _lib.call("rgm_conv_norm_f16", _lib.ptr(x), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta), None,
              _lib.ptr(raw), _lib.ptr(out), n, H, H, c, c, 2, 1, _lib.ptr(scratch), _lib.ptr(err), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert err.item() == 0, "a GroupNorm-in-epilogue wait gave up"
    assert torch.equal(raw, plain)
    count = scratch.view(n, 64, 2)[:, :, 0] & 255
    assert int(count.min()) == int(count.max()) == 2 * 4 * (H * H // 256)
    conv = F.conv2d(F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest"), w.half().float(), bias,
                    padding=1)
    ref = F.silu(F.group_norm(conv, 32, gamma, beta, eps=1e-6)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 6e-3 * ref.abs().max().item() + 1e-3
    cpg = c // 32
    r32 = raw.float().reshape(n, 4 * H * H, 32, cpg)
    mean = r32.mean(dim=(1, 3))
    rstd = (r32.var(dim=(1, 3), unbiased=False) + 1e-6).rsqrt()
    a = rstd.repeat_interleave(cpg, dim=1) * gamma
    b = beta - mean.repeat_interleave(cpg, dim=1) * a
    two = torch.empty_like(out)
    ab = torch.stack((a, b), dim=-1).contiguous()
    _lib.call("rgm_gn_apply_f16", _lib.ptr(raw), _lib.ptr(ab), _lib.ptr(two), n, 4 * H * H, c, 1, _lib.stream_ptr())
    torch.cuda.synchronize()
    d = (out.float() - two.float()).abs()
    assert d.max().item() <= 3e-3 * two.float().abs().max().item() + 1e-3, d.max().item()


def test_conv_norm_flags_statistics_outside_the_fixed_point_range(cuda):
    """The in-epilogue GroupNorm accumulates its statistics as 36.20 fixed-point integers: activations with an rms of
    several hundred over a whole image would wrap them, so the launch must raise the error flag (code 2) instead."""
    n, H, c = 2, 32, 128
    g = torch.Generator(device="cpu").manual_seed(3)
    x = (torch.randn(n, H, H, c, generator=g) * 40.0).to(cuda).half()
    w = (torch.randn(c, c, 3, 3, generator=g) * 0.5).to(cuda)          # conv output rms ~ 40 * 0.5 * sqrt(1152) ~ 680
    bias = torch.zeros(c, device=cuda)
    gamma, beta = torch.ones(c, device=cuda), torch.zeros(c, device=cuda)
    scratch = torch.zeros(n * 128, device=cuda, dtype=torch.int32)
    err = torch.zeros(1, device=cuda, dtype=torch.int32)
    out = torch.empty(n, H, H, c, device=cuda, dtype=torch.float16)
    _lib.call("rgm_conv_norm_f16", _lib.ptr(x), _lib.ptr(_pack(w, 1)), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta),
              None, None, _lib.ptr(out), n, H, H, c, c, 1, 1, _lib.ptr(scratch), _lib.ptr(err), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert err.item() == 2
