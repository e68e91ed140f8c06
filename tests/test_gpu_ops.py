"""Per-kernel parity against the oracle's operators, through the C ABI.

Inputs and weights are rounded to bf16 first, so the oracle (fp32 on those
rounded values) and the kernels (bf16 operands, fp32 accumulation, one bf16
rounding of the output) differ only by accumulation order, the activation's
approximation and the final rounding.  Stated tolerance for bf16 outputs:
rel-L2 <= 4e-3 and max-abs <= 2^-7 * max|ref| (two bf16 ulps at the top of the
range); fp32 outputs of the heads: rel-L2 <= 2e-3.
"""
import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import bf16_round, nchw_f32, nhwc_bf16, report

pytestmark = pytest.mark.gpu

REL_TOL = 4e-3
MAX_TOL = 2.0 ** -7


@pytest.fixture(scope="module")
def lib():
    from hgr_b200 import _lib
    return _lib.load()


def _chk(rc, what):
    from hgr_b200 import _lib
    _lib.check(rc, what)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _act(v, act):
    return F.silu(v) if act == 1 else (F.gelu(v) if act == 2 else v)


def run_conv(lib, x, w, scale, shift, k, s, act, res=None, in_pad=(0, 0), out_pad=(0, 0), res_pad=(0, 0)):
    """x NCHW fp32 (already bf16-rounded) on CPU.  *_pad = (channels before, channels after) of garbage
    around the slice inside its NHWC buffer, to exercise the concat-slice addressing."""
    dev = torch.device("cuda")
    b, cin, h, wd = x.shape
    cout = w.shape[0]

    def embed(t, pad):
        full = torch.randn(t.shape[0], pad[0] + t.shape[1] + pad[1], t.shape[2], t.shape[3])
        full[:, pad[0]: pad[0] + t.shape[1]] = t
        return nhwc_bf16(full, dev)

    xin = embed(x, in_pad)
    ho, wo = h // s, wd // s
    out_ctot = out_pad[0] + cout + out_pad[1]
    out = torch.full((b, ho, wo, out_ctot), 7.0, dtype=torch.bfloat16, device=dev)
    wp = w.permute(0, 2, 3, 1).contiguous().to(dev, torch.bfloat16)
    sc = None if scale is None else scale.to(dev).float().contiguous()
    sh = None if shift is None else shift.to(dev).float().contiguous()
    rbuf = None if res is None else embed(res, res_pad)
    _chk(lib.hgr_conv_bn_act(xin.data_ptr(), b, h, wd, xin.shape[-1], in_pad[0], cin, wp.data_ptr(), _ptr(sc), _ptr(sh),
                             k, s, act, _ptr(rbuf), 0 if res is None else rbuf.shape[-1], res_pad[0], out.data_ptr(),
                             out_ctot, out_pad[0], cout, _stream()), "hgr_conv_bn_act")
    torch.cuda.synchronize()
    got = nchw_f32(out)
    # untouched channels must still hold the fill value
    if out_pad[0]:
        assert torch.all(got[:, : out_pad[0]] == 7.0)
    if out_pad[1]:
        assert torch.all(got[:, out_pad[0] + cout:] == 7.0)
    y = F.conv2d(x, bf16_round(w), None, stride=s, padding=k // 2)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    return got[:, out_pad[0]: out_pad[0] + cout], _act(y, act)


CONV_CASES = [
    # name, B, H, cin, cout, k, s, act, res, in_pad, out_pad
    ("1x1_128_128_48", 2, 48, 128, 128, 1, 1, 1, False, (0, 0), (0, 128)),
    ("3x3_64_64_48_slice_res", 2, 48, 64, 64, 3, 1, 1, True, (64, 128), (128, 64)),
    ("3x3_128_128_24_ragged_batch", 3, 24, 128, 128, 3, 1, 1, True, (128, 256), (256, 128)),
    ("3x3_256_256_12_ragged_batch", 5, 12, 256, 256, 3, 1, 1, False, (0, 0), (0, 0)),
    ("1x1_1024_512_12", 9, 12, 1024, 512, 1, 1, 1, False, (0, 0), (0, 0)),
    ("3x3s2_64_128_96", 2, 96, 64, 128, 3, 2, 1, False, (0, 0), (0, 0)),
    ("3x3s2_128_256_48", 3, 48, 128, 256, 3, 2, 1, False, (0, 0), (0, 0)),
    ("3x3s2_256_512_24", 9, 24, 256, 512, 3, 2, 1, False, (0, 0), (0, 0)),
    ("3x3_64_64_64_noact", 1, 64, 64, 64, 3, 1, 0, False, (0, 0), (0, 0)),
    ("3x3_128_128_32", 2, 32, 128, 128, 3, 1, 1, True, (0, 0), (0, 0)),
    ("3x3_256_256_16", 3, 16, 256, 256, 3, 1, 1, True, (0, 0), (0, 0)),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_bn_act(lib, case):
    name, b, h, cin, cout, k, s, act, use_res, in_pad, out_pad = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    x = bf16_round(torch.randn(b, cin, h, h, generator=g))
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.3
    res = bf16_round(torch.randn(b, cout, h // s, h // s, generator=g)) if use_res else None
    got, ref = run_conv(lib, x, w, scale, shift, k, s, act, res, in_pad, out_pad, res_pad=(64, 0))
    r, m = report("conv " + name, got, ref)
    assert r <= REL_TOL and m <= MAX_TOL


LINEAR_CASES = [
    ("qkv_300", 300, 256, 768, 0, False, False),
    ("out_res_1160", 1160, 256, 256, 0, False, True),
    ("ff1_gelu_bias_77", 77, 256, 256, 2, True, False),
    ("ff2_bias_res_4096", 4096, 256, 256, 0, True, True),
    ("wide_k_512_129", 129, 512, 256, 0, True, False),
]


@pytest.mark.parametrize("case", LINEAR_CASES, ids=[c[0] for c in LINEAR_CASES])
def test_linear(lib, case):
    name, rows, cin, cout, act, use_bias, use_res = case
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    x = bf16_round(torch.randn(rows, cin, generator=g))
    w = bf16_round(torch.randn(cout, cin, generator=g) * cin ** -0.5)
    bias = torch.randn(cout, generator=g) * 0.2 if use_bias else None
    res = bf16_round(torch.randn(rows, cout, generator=g)) if use_res else None
    xd, wd = x.to(dev, torch.bfloat16), w.to(dev, torch.bfloat16)
    bd = None if bias is None else bias.to(dev)
    rd = None if res is None else res.to(dev, torch.bfloat16)
    y = torch.full((rows, cout), 7.0, dtype=torch.bfloat16, device=dev)
    _chk(lib.hgr_linear(xd.data_ptr(), rows, cin, wd.data_ptr(), None, _ptr(bd), act, _ptr(rd), y.data_ptr(), cout,
                        None, None, _stream()), "hgr_linear")
    torch.cuda.synchronize()
    ref = F.linear(x, w, bias)
    if res is not None:
        ref = ref + res
    ref = _act(ref, act)
    r, m = report("linear " + name, y, ref)
    assert r <= REL_TOL and m <= MAX_TOL


@pytest.mark.parametrize("rows,cout,act", [(300, 768, 0), (1160, 256, 2), (77, 256, 0)])
def test_linear_with_folded_layernorm(lib, rows, cout, act):
    """LayerNorm -> Linear (reference model/transformer.py:33-34, 63-65) as ONE GEMM on the un-normalised rows:
    the producer GEMM's epilogue leaves (mean, rstd) per row, the consumer folds gamma into W and applies
    rstd * acc - rstd * mean * c + d.  Checked as a producer/consumer pair against the oracle's operators."""
    from hgr_b200 import packing
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(rows + cout)
    # producer: x1 = a Wo^T + x0 (to_out + residual), emits row statistics of x1
    a = bf16_round(torch.randn(rows, 256, generator=g))
    x0 = bf16_round(torch.randn(rows, 256, generator=g) * 3 + 0.5)
    wo = bf16_round(torch.randn(256, 256, generator=g) / 16)
    ad, x0d, wod = (t.to(dev, torch.bfloat16) for t in (a, x0, wo))
    x1 = torch.empty(rows, 256, dtype=torch.bfloat16, device=dev)
    stats = torch.full((rows, 2), 7.0, dtype=torch.float32, device=dev)
    _chk(lib.hgr_linear(ad.data_ptr(), rows, 256, wod.data_ptr(), None, None, 0, x0d.data_ptr(), x1.data_ptr(), 256,
                        None, stats.data_ptr(), _stream()), "hgr_linear stats_out")
    torch.cuda.synchronize()
    x1_ref = F.linear(a, wo) + x0
    r, m = report("folded-LN producer x1", x1, x1_ref)
    assert r <= REL_TOL and m <= MAX_TOL
    mean_ref, var_ref = x1_ref.mean(1), x1_ref.var(1, unbiased=False)
    torch.testing.assert_close(stats[:, 0].cpu(), mean_ref, rtol=0, atol=2e-3)
    torch.testing.assert_close(stats[:, 1].cpu(), torch.rsqrt(var_ref + 1e-5), rtol=2e-3, atol=0)
    # consumer: y = act(LN(x1) W^T + b) from the bf16 x1 and the statistics above
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.1
    w = torch.randn(cout, 256, generator=g) / 16
    bias = torch.randn(cout, generator=g) * 0.2
    wp = (w * gamma[None, :]).to(torch.bfloat16)
    c, d = wp.float().sum(1), w @ beta + bias
    wpd, cd, dd = wp.to(dev), c.to(dev), d.to(dev)
    y = torch.full((rows, cout), 7.0, dtype=torch.bfloat16, device=dev)
    _chk(lib.hgr_linear(x1.data_ptr(), rows, 256, wpd.data_ptr(), cd.data_ptr(), dd.data_ptr(), act, None,
                        y.data_ptr(), cout, stats.data_ptr(), None, _stream()), "hgr_linear stats_in")
    torch.cuda.synchronize()
    ref = _act(F.linear(F.layer_norm(x1.float().cpu(), (256,), gamma, beta, 1e-5), w, bias), act)
    r, m = report(f"folded-LN consumer rows={rows} cout={cout} act={act}", y, ref)
    assert r <= 6e-3 and m <= 2 * MAX_TOL  # W' = bf16(gamma*W) and the cancellation add about one more bf16 ulp
    del packing


@pytest.mark.parametrize("b,h", [(2, 96), (3, 96), (1, 32), (5, 64), (41, 96), (3, 128)])
def test_conv_chain(lib, b, h):
    """conv2 -> cspelan1.cv1 (reference model/gelan.py:156, :127) as one CTA-pair kernel against the fp32 operators
    and against the two separate hgr_conv_bn_act launches it replaces (same rounding point for the tensor between)."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(b * 1000 + h)
    x = bf16_round(torch.randn(b, 64, h, h, generator=g))
    w1 = torch.randn(128, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5
    w2 = torch.randn(128, 128, 1, 1, generator=g) * (2.0 / 128) ** 0.5
    s1, s2 = torch.rand(128, generator=g) + 0.5, torch.rand(128, generator=g) + 0.5
    t1, t2 = torch.randn(128, generator=g) * 0.3, torch.randn(128, generator=g) * 0.3
    xin = nhwc_bf16(x, dev)
    w1d = w1.permute(0, 2, 3, 1).contiguous().to(dev, torch.bfloat16)
    w2d = w2.reshape(128, 128).contiguous().to(dev, torch.bfloat16)
    s1d, s2d, t1d, t2d = (t.to(dev).float().contiguous() for t in (s1, s2, t1, t2))
    out = torch.full((b, h // 2, h // 2, 256), 7.0, dtype=torch.bfloat16, device=dev)
    _chk(lib.hgr_conv_chain(xin.data_ptr(), b, h, h, w1d.data_ptr(), s1d.data_ptr(), t1d.data_ptr(), w2d.data_ptr(),
                            s2d.data_ptr(), t2d.data_ptr(), out.data_ptr(), 256, 64, _stream()), "hgr_conv_chain")
    torch.cuda.synchronize()
    got = nchw_f32(out)
    assert torch.all(got[:, :64] == 7.0) and torch.all(got[:, 192:] == 7.0)
    # the two separate launches
    mid = torch.empty(b, h // 2, h // 2, 128, dtype=torch.bfloat16, device=dev)
    two = torch.empty_like(mid)
    _chk(lib.hgr_conv_bn_act(xin.data_ptr(), b, h, h, 64, 0, 64, w1d.data_ptr(), s1d.data_ptr(), t1d.data_ptr(), 3, 2, 1,
                             None, 0, 0, mid.data_ptr(), 128, 0, 128, _stream()), "conv2")
    _chk(lib.hgr_conv_bn_act(mid.data_ptr(), b, h // 2, h // 2, 128, 0, 128, w2d.data_ptr(), s2d.data_ptr(),
                             t2d.data_ptr(), 1, 1, 1, None, 0, 0, two.data_ptr(), 128, 0, 128, _stream()), "cv1")
    torch.cuda.synchronize()
    a2 = F.silu(F.conv2d(x, bf16_round(w1), None, stride=2, padding=1) * s1.view(1, -1, 1, 1) + t1.view(1, -1, 1, 1))
    ref = F.silu(F.conv2d(a2, bf16_round(w2)) * s2.view(1, -1, 1, 1) + t2.view(1, -1, 1, 1))
    r, m = report(f"conv_chain b={b} h={h} vs fp32 operators", got[:, 64:192], ref)
    assert r <= 6e-3 and m <= 2 * MAX_TOL

    r2, _ = report(f"conv_chain b={b} h={h} vs two launches", got[:, 64:192], nchw_f32(two))
    assert r2 <= 1e-3


@pytest.mark.parametrize("b,h", [(2, 48), (3, 48), (1, 16), (41, 48), (5, 64)])
def test_gelan_tail(lib, b, h):
    """cspelan1.cv3.0.cv2 (+ residual, SiLU) -> cspelan1.cv4 (reference model/gelan.py:73-87, :137-142) as one CTA-pair
    kernel against the fp32 operators and against the two hgr_conv_bn_act launches it replaces (same rounding point
    for y3, same K order in the 1x1 layer)."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(b * 100 + h)
    t = bf16_round(torch.randn(b, 64, h, h, generator=g))
    y012 = bf16_round(torch.randn(b, 192, h, h, generator=g))
    wh = torch.randn(64, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5
    w4 = torch.randn(128, 256, 1, 1, generator=g) * (2.0 / 256) ** 0.5
    sh_, s4 = torch.rand(64, generator=g) + 0.5, torch.rand(128, generator=g) + 0.5
    th_, t4 = torch.randn(64, generator=g) * 0.3, torch.randn(128, generator=g) * 0.3
    td = nhwc_bf16(t, dev)
    gbuf = torch.full((b, h, h, 256), 7.0, dtype=torch.bfloat16, device=dev)
    gbuf[..., :192] = nhwc_bf16(y012, dev)
    whd = wh.permute(0, 2, 3, 1).contiguous().to(dev, torch.bfloat16)
    w4d = w4.reshape(128, 256).contiguous().to(dev, torch.bfloat16)
    shd, s4d, thd, t4d = (v.to(dev).float().contiguous() for v in (sh_, s4, th_, t4))
    out = torch.full((b, h, h, 128), 7.0, dtype=torch.bfloat16, device=dev)
    _chk(lib.hgr_gelan_tail(td.data_ptr(), gbuf.data_ptr(), b, h, h, whd.data_ptr(), shd.data_ptr(), thd.data_ptr(),
                            w4d.data_ptr(), s4d.data_ptr(), t4d.data_ptr(), out.data_ptr(), _stream()), "hgr_gelan_tail")
    torch.cuda.synchronize()
    assert torch.all(gbuf[..., 192:] == 7.0)  # y3 is not written
    # the two launches
    g2 = gbuf.clone()
    two = torch.empty_like(out)
    _chk(lib.hgr_conv_bn_act(td.data_ptr(), b, h, h, 64, 0, 64, whd.data_ptr(), shd.data_ptr(), thd.data_ptr(), 3, 1, 1,
                             g2.data_ptr(), 256, 128, g2.data_ptr(), 256, 192, 64, _stream()), "cv3.0.cv2")
    _chk(lib.hgr_conv_bn_act(g2.data_ptr(), b, h, h, 256, 0, 256, w4d.data_ptr(), s4d.data_ptr(), t4d.data_ptr(), 1, 1, 1,
                             None, 0, 0, two.data_ptr(), 128, 0, 128, _stream()), "cv4")
    torch.cuda.synchronize()
    y2 = y012[:, 128:192]
    y3 = F.silu(F.conv2d(t, bf16_round(wh), None, padding=1) * sh_.view(1, -1, 1, 1) + th_.view(1, -1, 1, 1) + y2)
    ref = F.silu(F.conv2d(torch.cat([y012, y3], 1), bf16_round(w4)) * s4.view(1, -1, 1, 1) + t4.view(1, -1, 1, 1))
    r, m = report(f"gelan_tail b={b} h={h} vs fp32 operators", nchw_f32(out), ref)
    assert r <= 6e-3 and m <= 2 * MAX_TOL
    r2, _ = report(f"gelan_tail b={b} h={h} vs two launches", nchw_f32(out), nchw_f32(two))
    assert r2 <= 1e-3


@pytest.mark.parametrize("b,size", [(2, 192), (3, 192), (1, 64), (5, 128), (41, 192), (3, 256)])
def test_stem_fused(lib, b, size):
    """conv1 -> conv2 -> cspelan1.cv1 (reference model/gelan.py:155, :156, :127) as one kernel against the fp32
    operators and against the two launches it replaces (hgr_conv1 + hgr_conv_chain: the same conv1 arithmetic and
    rounding points, conv2's taps summed in another order).  b = 41 gives every CTA pair several tiles (both input
    patch buffers, both accumulator stages, the per-plane barriers through several phases); b = 3 an odd image count."""
    from hgr_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(b * 1000 + size)
    x = bf16_round(torch.randn(b, 3, size, size, generator=g))
    w0 = torch.randn(64, 3, 3, 3, generator=g) * (2.0 / 27) ** 0.5
    s0, t0 = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.3
    w1 = torch.randn(128, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5
    w2 = torch.randn(128, 128, 1, 1, generator=g) * (2.0 / 128) ** 0.5
    s1, s2 = torch.rand(128, generator=g) + 0.5, torch.rand(128, generator=g) + 0.5
    t1, t2 = torch.randn(128, generator=g) * 0.3, torch.randn(128, generator=g) * 0.3
    wk = torch.zeros(64, 32)
    wk[:, :27] = (w0 * s0.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(64, 27)
    wk = wk.to(dev, torch.bfloat16)
    t0d = t0.to(dev)
    w1d = w1.permute(0, 2, 3, 1).contiguous().to(dev, torch.bfloat16)
    w2d = w2.reshape(128, 128).contiguous().to(dev, torch.bfloat16)
    s1d, s2d, t1d, t2d = (t.to(dev).float().contiguous() for t in (s1, s2, t1, t2))
    xd = x.to(dev, torch.bfloat16).contiguous()
    q = size // 4
    out = torch.full((b, q, q, 256), 7.0, dtype=torch.bfloat16, device=dev)
    _chk(lib.hgr_stem_fused(xd.data_ptr(), b, size, wk.data_ptr(), t0d.data_ptr(), w1d.data_ptr(), s1d.data_ptr(),
                            t1d.data_ptr(), w2d.data_ptr(), s2d.data_ptr(), t2d.data_ptr(), out.data_ptr(), 256, 64,
                            _stream()), "hgr_stem_fused")
    torch.cuda.synchronize()
    got = nchw_f32(out)
    assert torch.all(got[:, :64] == 7.0) and torch.all(got[:, 192:] == 7.0)
    # the launches it replaces
    a1 = torch.empty(b, size // 2, size // 2, 64, dtype=torch.bfloat16, device=dev)
    two = torch.full((b, q, q, 256), 7.0, dtype=torch.bfloat16, device=dev)
    _chk(lib.hgr_conv1(xd.data_ptr(), _lib.BF16, b, size, wk.data_ptr(), t0d.data_ptr(), a1.data_ptr(), _stream()),
         "hgr_conv1")
    _chk(lib.hgr_conv_chain(a1.data_ptr(), b, size // 2, size // 2, w1d.data_ptr(), s1d.data_ptr(), t1d.data_ptr(),
                            w2d.data_ptr(), s2d.data_ptr(), t2d.data_ptr(), two.data_ptr(), 256, 64, _stream()),
         "hgr_conv_chain")
    torch.cuda.synchronize()
    w0ref = wk.float().cpu()[:, :27].reshape(64, 3, 3, 3).permute(0, 3, 1, 2)  # the bf16-rounded folded weights
    a1r = F.silu(F.conv2d(x, w0ref, None, stride=2, padding=1) + t0.view(1, -1, 1, 1))
    a2r = F.silu(F.conv2d(a1r, bf16_round(w1), None, stride=2, padding=1) * s1.view(1, -1, 1, 1) + t1.view(1, -1, 1, 1))
    ref = F.silu(F.conv2d(a2r, bf16_round(w2)) * s2.view(1, -1, 1, 1) + t2.view(1, -1, 1, 1))
    r, m = report(f"stem_fused b={b} size={size} vs fp32 operators", got[:, 64:192], ref)
    assert r <= 8e-3 and m <= 2 * MAX_TOL
    r2, _ = report(f"stem_fused b={b} size={size} vs conv1 + conv_chain", got[:, 64:192], nchw_f32(two)[:, 64:192])
    assert r2 <= 2e-3


@pytest.mark.parametrize("rows", [1160, 128 * 148 * 2 + 77, 77])
@pytest.mark.parametrize("inplace", [False, True], ids=["out", "inplace"])
def test_vit_block(lib, rows, inplace):
    """to_out + residual -> LayerNorm -> Linear -> GELU -> Linear -> residual (reference model/transformer.py:75, 93,
    29-42, 94) as one chained kernel, against the fp32 operators and against the three separate hgr_linear launches
    it replaces (same rounding points, so the two CUDA paths agree to bf16 rounding of the last layer)."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(rows)
    a = bf16_round(torch.randn(rows, 256, generator=g))
    x0 = bf16_round(torch.randn(rows, 256, generator=g) * 2 + 0.3)
    wo = bf16_round(torch.randn(256, 256, generator=g) / 16)
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.1
    w1 = torch.randn(256, 256, generator=g) / 16
    b1 = torch.randn(256, generator=g) * 0.2
    w2 = bf16_round(torch.randn(256, 256, generator=g) / 16)
    b2 = torch.randn(256, generator=g) * 0.2
    w1p = (w1 * gamma[None, :]).to(torch.bfloat16)
    c1, d1 = w1p.float().sum(1), w1 @ beta + b1
    ad, x0d, wod, w2d = (t.to(dev, torch.bfloat16) for t in (a, x0, wo, w2))
    w1d, c1d, d1d, b2d = w1p.to(dev), c1.to(dev), d1.to(dev), b2.to(dev)
    # fused
    xin = x0d.clone()
    x2 = xin if inplace else torch.full((rows, 256), 7.0, dtype=torch.bfloat16, device=dev)
    stats = torch.full((rows, 2), 7.0, dtype=torch.float32, device=dev)
    _chk(lib.hgr_vit_block(ad.data_ptr(), xin.data_ptr(), rows, wod.data_ptr(), w1d.data_ptr(), c1d.data_ptr(),
                           d1d.data_ptr(), w2d.data_ptr(), b2d.data_ptr(), x2.data_ptr(), stats.data_ptr(), _stream()),
         "hgr_vit_block")
    torch.cuda.synchronize()
    # the three separate launches
    x1 = torch.empty(rows, 256, dtype=torch.bfloat16, device=dev)
    h = torch.empty_like(x1)
    y = torch.empty_like(x1)
    st1 = torch.empty(rows, 2, dtype=torch.float32, device=dev)
    st2 = torch.empty_like(st1)
    _chk(lib.hgr_linear(ad.data_ptr(), rows, 256, wod.data_ptr(), None, None, 0, x0d.data_ptr(), x1.data_ptr(), 256,
                        None, st1.data_ptr(), _stream()), "to_out")
    _chk(lib.hgr_linear(x1.data_ptr(), rows, 256, w1d.data_ptr(), c1d.data_ptr(), d1d.data_ptr(), 2, None,
                        h.data_ptr(), 256, st1.data_ptr(), None, _stream()), "net.1")
    _chk(lib.hgr_linear(h.data_ptr(), rows, 256, w2d.data_ptr(), None, b2d.data_ptr(), 0, x1.data_ptr(), y.data_ptr(),
                        256, None, st2.data_ptr(), _stream()), "net.4")
    torch.cuda.synchronize()
    # fp32 operators
    x1_ref = F.linear(a, wo) + x0
    h_ref = F.gelu(F.linear(F.layer_norm(x1_ref, (256,), gamma, beta, 1e-5), w1, b1))
    x2_ref = F.linear(h_ref, w2, b2) + x1_ref
    r, m = report(f"vit_block rows={rows} vs fp32 operators", x2, x2_ref)
    assert r <= 6e-3 and m <= 3 * MAX_TOL
    r3, m3 = report(f"vit_block rows={rows} vs three launches", x2, y.float().cpu())
    assert r3 <= 2e-3
    mean_ref, var_ref = x2_ref.mean(1), x2_ref.var(1, unbiased=False)
    torch.testing.assert_close(stats[:, 0].cpu(), mean_ref, rtol=0, atol=4e-3)
    torch.testing.assert_close(stats[:, 1].cpu(), torch.rsqrt(var_ref + 1e-5), rtol=4e-3, atol=0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("size,b", [(64, 3), (192, 3), (192, 41), (256, 5), (128, 1), (96, 2)])
def test_conv1(lib, dtype, size, b):
    """Sides that are multiples of 64 run the tcgen05 kernel (conv1_tc.cu; b = 41 gives every CTA several bands, both
    patch buffers and both accumulator stages), 96 the mma.sync kernel (conv1.cu)."""
    from hgr_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(size)
    x = torch.randn(b, 3, size, size, generator=g)
    if dtype == torch.bfloat16:
        x = bf16_round(x)
    w = torch.randn(64, 3, 3, 3, generator=g) * (2.0 / 27) ** 0.5
    scale = torch.rand(64, generator=g) + 0.5
    shift = torch.randn(64, generator=g) * 0.3
    wk = torch.zeros(64, 32)
    wk[:, :27] = (w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(64, 27)
    wk = wk.to(dev, torch.bfloat16)
    out = torch.full((b, size // 2, size // 2, 64), 7.0, dtype=torch.bfloat16, device=dev)
    xd = x.to(dev, dtype).contiguous()
    shd = shift.to(dev)
    _chk(lib.hgr_conv1(xd.data_ptr(), _lib.F32 if dtype == torch.float32 else _lib.BF16, b, size, wk.data_ptr(),
                       shd.data_ptr(), out.data_ptr(), _stream()), "hgr_conv1")
    torch.cuda.synchronize()
    wref = wk.float().cpu()[:, :27].reshape(64, 3, 3, 3).permute(0, 3, 1, 2)  # the bf16-rounded folded weights
    ref = F.silu(F.conv2d(bf16_round(x), wref, None, stride=2, padding=1) + shift.view(1, -1, 1, 1))
    r, m = report(f"conv1 {size} b={b} {dtype}", nchw_f32(out), ref)
    assert r <= REL_TOL and m <= MAX_TOL


def test_layernorm(lib):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    rows = 1237
    x = bf16_round(torch.randn(rows, 256, generator=g) * 3 + 0.5)
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.1
    xd = x.to(dev, torch.bfloat16)
    y = torch.empty_like(xd)
    gd, bd = gamma.to(dev), beta.to(dev)
    _chk(lib.hgr_layernorm(xd.data_ptr(), y.data_ptr(), gd.data_ptr(), bd.data_ptr(), rows, _stream()), "hgr_layernorm")
    torch.cuda.synchronize()
    r, m = report("layernorm", y, F.layer_norm(x, (256,), gamma, beta, 1e-5))
    assert r <= REL_TOL and m <= MAX_TOL


@pytest.mark.parametrize("tokens", [145, 257, 17])
@pytest.mark.parametrize("probs", [None, torch.float32, torch.bfloat16], ids=["noprobs", "p32", "pbf16"])
def test_attention(lib, tokens, probs):
    from hgr_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(tokens)
    b = 3
    qkv = bf16_round(torch.randn(b, tokens, 768, generator=g) * 1.5)
    qd = qkv.to(dev, torch.bfloat16)
    out = torch.full((b, tokens, 256), 7.0, dtype=torch.bfloat16, device=dev)
    pd = None if probs is None else torch.full((b, 8, tokens, tokens), 7.0, dtype=probs, device=dev)
    _chk(lib.hgr_attention(qd.data_ptr(), out.data_ptr(), _ptr(pd), _lib.BF16 if probs == torch.bfloat16 else _lib.F32,
                           b, tokens, _stream()), "hgr_attention")
    torch.cuda.synchronize()
    q, k, v = qkv.chunk(3, dim=-1)
    sp = lambda t: t.reshape(b, tokens, 8, 32).permute(0, 2, 1, 3)
    attn = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) * 32 ** -0.5, dim=-1)
    ref = (attn @ sp(v)).permute(0, 2, 1, 3).reshape(b, tokens, 256)
    r, m = report(f"attention T={tokens}", out, ref)
    assert r <= 6e-3 and m <= 2 * MAX_TOL  # P is rounded to bf16 before P.V, like the reference's bf16 path
    if pd is not None:
        r2, m2 = report(f"attention probs T={tokens} {probs}", pd, attn)
        assert r2 <= (REL_TOL if probs == torch.bfloat16 else 1e-4) and m2 <= MAX_TOL


@pytest.mark.parametrize("tokens,b", [(145, 3), (145, 40), (160, 2), (129, 5), (145, 300), (145, 1)])
def test_attention_tc(lib, tokens, b):
    """The tcgen05 attention kernel (csrc/attention_tc.cu, opt-in through HGR_ATTN_TC=1) against the fp32 operators."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(tokens * 7 + b)
    qkv = bf16_round(torch.randn(b, tokens, 768, generator=g) * 1.5)
    qd = qkv.to(dev, torch.bfloat16)
    out = torch.full((b, tokens, 256), 7.0, dtype=torch.bfloat16, device=dev)
    _chk(lib.hgr_attention_tc(qd.data_ptr(), out.data_ptr(), b, tokens, _stream()), "hgr_attention_tc")
    torch.cuda.synchronize()
    q, k, v = qkv.chunk(3, dim=-1)
    sp = lambda t: t.reshape(b, tokens, 8, 32).permute(0, 2, 1, 3)
    attn = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) * 32 ** -0.5, dim=-1)
    ref = (attn @ sp(v)).permute(0, 2, 1, 3).reshape(b, tokens, 256)
    r, m = report(f"attention_tc T={tokens} B={b}", out, ref)
    assert r <= 6e-3 and m <= 2 * MAX_TOL


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_cls_head(lib, out_dtype):
    from hgr_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(11)
    b, t, c = 13, 145, 19
    tok = bf16_round(torch.randn(b, t, 256, generator=g) * 2)
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.1
    w, bias = torch.randn(c, 256, generator=g) * 0.1, torch.randn(c, generator=g)
    td = tok.to(dev, torch.bfloat16)
    out = torch.empty(b, c, dtype=out_dtype, device=dev)
    args = [x.to(dev) for x in (gamma, beta, w, bias)]
    _chk(lib.hgr_cls_head(td.data_ptr(), *[a.data_ptr() for a in args], out.data_ptr(),
                          _lib.F32 if out_dtype == torch.float32 else _lib.BF16, b, t, c, _stream()), "hgr_cls_head")
    torch.cuda.synchronize()
    ref = F.linear(F.layer_norm(tok[:, 0], (256,), gamma, beta, 1e-5), w, bias)
    r, m = report(f"cls_head {out_dtype}", out, ref)
    assert r <= (1e-5 if out_dtype == torch.float32 else REL_TOL)


@pytest.mark.parametrize("feat", [12, 16])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_pose_head(lib, feat, out_dtype):
    from hgr_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(feat)
    b, j = 3, 21
    tok = bf16_round(torch.randn(b, feat * feat + 1, 256, generator=g) * 2)
    w = bf16_round(torch.randn(j, 256, generator=g) * 0.1)
    bias = torch.randn(j, generator=g)
    td, wd, bd = tok.to(dev, torch.bfloat16), w.to(dev, torch.bfloat16), bias.to(dev)
    out = torch.full((b, j, 4 * feat, 4 * feat), 7.0, dtype=out_dtype, device=dev)
    _chk(lib.hgr_pose_head(td.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(),
                           _lib.F32 if out_dtype == torch.float32 else _lib.BF16, b, feat, j, _stream()),
         "hgr_pose_head")
    torch.cuda.synchronize()
    fmap = tok[:, 1:].reshape(b, feat, feat, 256).permute(0, 3, 1, 2)
    up = F.relu(F.interpolate(fmap, scale_factor=(4, 4), mode="bilinear", align_corners=True))
    ref = F.conv2d(bf16_round(up), w.view(j, 256, 1, 1), bias)
    r, m = report(f"pose_head F={feat} {out_dtype}", out, ref)
    # the interpolated operand is formed in bf16 arithmetic (value + weight * slope): about two bf16 roundings more
    # than the oracle's single rounding of the fp32 up-sampled tensor
    assert r <= 6e-3 and m <= 2 * MAX_TOL


@pytest.mark.parametrize("feat,b", [(12, 3), (12, 310), (16, 5), (8, 4), (4, 2), (20, 2)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_pose_head_decode(lib, feat, b, out_dtype):
    """The keypoint decode fused into the pose head's epilogue (libs/utils.py:4-32 after transformer.py:146-150):
    bit-identical to get_max_preds on the heatmaps the same call writes, and identical again in the keypoints-only
    mode where the heatmaps never reach memory.  b = 310 gives every CTA several images (arg-max state reset)."""
    from hgr_b200 import _lib
    from oracle import multitasknet_oracle as O
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(feat * 31 + b)
    j = 21
    tok = bf16_round(torch.randn(b, feat * feat + 1, 256, generator=g) * 2)
    w = bf16_round(torch.randn(j, 256, generator=g) * 0.1)
    bias = torch.randn(j, generator=g) * 0.5
    bias[3] = -100.0  # an all-negative map: the prediction is masked to (0, 0)
    td, wd, bd = tok.to(dev, torch.bfloat16), w.to(dev, torch.bfloat16), bias.to(dev)
    so = 4 * feat
    dt = _lib.F32 if out_dtype == torch.float32 else _lib.BF16
    heat = torch.full((b, j, so, so), 7.0, dtype=out_dtype, device=dev)
    preds = torch.full((b, j, 2), -1.0, device=dev)
    maxv = torch.full((b, j, 1), -1.0, device=dev)
    _chk(lib.hgr_pose_head_decode(td.data_ptr(), wd.data_ptr(), bd.data_ptr(), heat.data_ptr(), dt, preds.data_ptr(),
                                  maxv.data_ptr(), b, feat, j, _stream()), "hgr_pose_head_decode")
    torch.cuda.synchronize()
    rp, rv = O.get_max_preds(heat.float().cpu().numpy())
    assert np.array_equal(preds.cpu().numpy().view(np.uint32), rp.view(np.uint32))
    assert np.array_equal(maxv.cpu().numpy().view(np.uint32), rv.view(np.uint32))
    assert float(preds[:, 3].abs().max()) == 0.0
    # the heatmaps are the ones the plain call writes
    heat2 = torch.empty_like(heat)
    _chk(lib.hgr_pose_head(td.data_ptr(), wd.data_ptr(), bd.data_ptr(), heat2.data_ptr(), dt, b, feat, j, _stream()),
         "hgr_pose_head")
    assert torch.equal(heat, heat2)
    # keypoints only
    p2, v2 = torch.full_like(preds, -1.0), torch.full_like(maxv, -1.0)
    _chk(lib.hgr_pose_head_decode(td.data_ptr(), wd.data_ptr(), bd.data_ptr(), None, dt, p2.data_ptr(), v2.data_ptr(),
                                  b, feat, j, _stream()), "hgr_pose_head_decode (keypoints only)")
    torch.cuda.synchronize()
    assert torch.equal(p2, preds) and torch.equal(v2, maxv)
    # and the values are the operator's (same tolerance as test_pose_head)
    fmap = tok[:, 1:].reshape(b, feat, feat, 256).permute(0, 3, 1, 2)
    up = F.relu(F.interpolate(fmap, scale_factor=(4, 4), mode="bilinear", align_corners=True))
    ref = F.conv2d(bf16_round(up), w.view(j, 256, 1, 1), bias)
    r, m = report(f"pose_head_decode F={feat} B={b} {out_dtype}", heat, ref)
    assert r <= 6e-3 and m <= 2 * MAX_TOL
