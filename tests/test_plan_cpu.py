"""CPU tests of the conv planners: descriptors produced by libfpg_b200.so (pure host code) are executed by the
numpy interpreter in igemm_interp.py and compared with torch.nn.functional on the same inputs."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import igemm_interp as interp
from fpgan import lib as L

SMS = 148


def make_act(n, h, w, c, c_stride=None, halo=0, fp32=0, base=interp.FAKE_BASE):
    a = L.Act()
    a.data = base
    a.n, a.h, a.w, a.c = n, h, w, c
    a.c_stride = c_stride or c
    a.halo = halo
    a.fp32 = fp32
    return a


def geom(r, s, stride, pad, c_in, c_out):
    g = L.ConvGeom()
    g.r, g.s, g.stride, g.pad, g.c_in, g.c_out = r, s, stride, pad, c_in, c_out
    return g


def nhwc_buffer(x_nchw, c_pad, halo=0, mode="reflect"):
    """NCHW tensor -> flat NHWC buffer with channel padding and (reflect or zero) halo."""
    if halo:
        x_nchw = F.pad(x_nchw, (halo,) * 4, mode) if mode == "reflect" else F.pad(x_nchw, (halo,) * 4)
    n, c, h, w = x_nchw.shape
    buf = torch.zeros(n, h, w, c_pad, dtype=torch.float64)
    buf[..., :c] = x_nchw.permute(0, 2, 3, 1)
    return buf.reshape(-1).numpy().copy()


def pack_fprop(w, c_in_pad, c_out_pad, taps_padded):
    k, c, r, s = w.shape
    out = torch.zeros(c_out_pad, taps_padded, c_in_pad, dtype=torch.float64)
    out[:k, :r * s, :c] = w.reshape(k, c, r * s).permute(0, 2, 1)
    return out.reshape(-1).numpy().copy()


def pack_dgrad(fpglib, w, g):
    """numpy twin of fpg_pack_weights_dgrad driven by fpg_dgrad_class_info."""
    k, c, r, s = w.shape
    src_tap = (C.c_int32 * L.FPG_MAX_TAPS)()
    ntp, off, ncls = C.c_int32(), C.c_int64(), C.c_int32()
    total = fpglib.fpg_packed_weight_dgrad_bytes(C.byref(g)) // 2
    buf = np.zeros(total)
    L.call("fpg_dgrad_class_info", C.byref(g), 0, src_tap, C.byref(ntp), C.byref(off), C.byref(ncls))
    wf = w.reshape(k, c, r * s).double().numpy()
    for cls in range(ncls.value):
        L.call("fpg_dgrad_class_info", C.byref(g), cls, src_tap, C.byref(ntp), C.byref(off), C.byref(ncls))
        m = np.zeros((g.c_in, ntp.value, g.c_out))
        for t in range(ntp.value):
            if src_tap[t] >= 0:
                m[:c, t, :k] = wf[:, :, src_tap[t]].T
        buf[off.value:off.value + m.size] = m.reshape(-1)
    return buf


FPROP_CASES = [
    # (n, h, w, c_real, c_pad, k_real, k_pad, r, stride, zero_pad, reflect_halo)
    (1, 16, 16, 9, 16, 64, 64, 7, 1, 0, 3),     # generator stem (model_architectures.py:312)
    (2, 16, 16, 64, 64, 128, 128, 3, 2, 1, 0),  # downsample (:314)
    (1, 8, 8, 64, 64, 64, 64, 3, 1, 0, 1),      # residual conv, reflect halo (:407)
    (1, 12, 12, 64, 64, 27, 32, 7, 1, 0, 3),    # content head (:328)
    (1, 8, 8, 64, 64, 10, 16, 1, 1, 0, 0),      # attention head (:334)
    (1, 16, 16, 12, 16, 64, 64, 4, 2, 1, 0),    # PatchGAN model.0 (:424)
    (1, 9, 9, 64, 64, 128, 128, 4, 1, 1, 0),    # PatchGAN model.8-style 4x4 s1 p1, odd extent (:434)
    (1, 7, 7, 128, 128, 1, 16, 4, 1, 1, 0),     # PatchGAN model.11 (:437)
    (1, 6, 6, 27, 32, 64, 64, 3, 1, 1, 0),      # 32-channel input (SW64 path)
]


@pytest.mark.parametrize("case", FPROP_CASES)
def test_fprop_plan_matches_conv2d(fpglib, case):
    n, h, w, c, cp, k, kp, r, stride, pad, halo = case
    torch.manual_seed(0)
    x = torch.randn(n, c, h, w, dtype=torch.float64)
    wt = torch.randn(k, c, r, r, dtype=torch.float64)
    bias = torch.randn(kp, dtype=torch.float64)
    bias[k:] = 0
    xin = F.pad(x, (halo,) * 4, "reflect") if halo else x
    ref = F.conv2d(xin, wt, bias[:k], stride=stride, padding=pad)
    ho, wo = ref.shape[2:]
    xa = make_act(n, h, w, cp, halo=halo)
    ya = make_act(n, ho, wo, kp, halo=1, base=interp.FAKE_BASE)  # output with its own halo: interior is written
    g = geom(r, r, stride, pad, cp, kp)
    d = L.FpropDesc()
    L.call("fpg_conv2d_fprop_plan", C.byref(xa), interp.FAKE_BASE, None, L.ACT_LEAKY, C.byref(g), C.byref(ya), SMS,
           C.byref(d))
    assert d.num_sub % (64 // d.cblk) == 0 and d.tile_h * d.tile_w in (128, 256)
    abuf = nhwc_buffer(x, cp, halo)
    bbuf = pack_fprop(wt, cp, kp, d.num_taps)
    assert bbuf.size * 2 == fpglib.fpg_packed_weight_bytes(C.byref(g))
    out = np.full(n * (ho + 2) * (wo + 2) * kp, np.nan)
    interp.run_fprop(d, abuf, bbuf, out, bias.numpy())
    out = torch.from_numpy(out).reshape(n, ho + 2, wo + 2, kp)
    got = out[:, 1:-1, 1:-1, :k].permute(0, 3, 1, 2)
    torch.testing.assert_close(got, F.leaky_relu(ref, 0.2), rtol=1e-9, atol=1e-9)
    assert torch.isnan(out[:, 0]).all() and torch.isnan(out[:, :, 0]).all()  # halo untouched
    assert (out[:, 1:-1, 1:-1, k:] == 0).all()  # padded channels stay exactly zero


DGRAD_CASES = [
    # (n, h, w, c_real, c_pad, k_real, k_pad, r, stride, zero_pad, reflect_halo) of the FORWARD conv
    (1, 8, 8, 64, 64, 64, 64, 3, 1, 0, 1),      # residual conv: gradient w.r.t. the reflect-padded input
    (1, 16, 16, 64, 64, 128, 128, 3, 2, 1, 0),  # stride-2 3x3 (== ConvTranspose2d 3x3 s2 p1 op1 forward, :324)
    (2, 16, 16, 64, 64, 128, 128, 4, 2, 1, 0),  # PatchGAN 4x4 s2
    (1, 9, 9, 64, 64, 128, 128, 4, 1, 1, 0),    # PatchGAN 4x4 s1 p1
    (1, 7, 7, 128, 128, 1, 16, 4, 1, 1, 0),     # model.11: 16-channel dy (SW32 path)
    (1, 12, 12, 64, 64, 27, 32, 7, 1, 0, 3),    # content head: 32-channel dy (SW64 path), halo 3
    (1, 16, 16, 12, 16, 64, 64, 4, 2, 1, 0),    # model.0: gradient w.r.t. the 16-channel D input
    (16, 64, 64, 64, 64, 64, 64, 3, 1, 0, 1),   # 66x66 haloed gradient: two tile regions (64x2 tiles + edge columns)
]


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_dgrad_plan_matches_autograd(fpglib, case):
    n, h, w, c, cp, k, kp, r, stride, pad, halo = case
    torch.manual_seed(1)
    xin = torch.randn(n, c, h + 2 * halo, w + 2 * halo, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(k, c, r, r, dtype=torch.float64)
    y = F.conv2d(xin, wt, None, stride=stride, padding=pad)
    dy = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, xin, dy)
    ho, wo = y.shape[2:]
    g = geom(r, r, stride, pad, cp, kp)
    dya = make_act(n, ho, wo, kp)
    dxa = make_act(n, h, w, cp, halo=halo)
    descs = (L.FpropDesc * 4)()
    nd = C.c_int()
    L.call("fpg_conv2d_dgrad_plan", C.byref(dya), interp.FAKE_BASE, None, L.ACT_NONE, C.byref(g), C.byref(dxa), SMS,
           descs, C.byref(nd))
    assert nd.value == (4 if stride == 2 else 1)
    if h == 64:
        assert descs[0].tiles_x1 == 1 and descs[0].x_org1 == 64 and descs[0].tile_w == 64, "expected two tile regions"
    abuf = nhwc_buffer(dy, kp)
    bbuf = pack_dgrad(fpglib, wt, g)
    hp, wp = h + 2 * halo, w + 2 * halo
    out = np.full(n * hp * wp * cp, np.nan)
    for q in range(nd.value):
        interp.run_fprop(descs[q], abuf, bbuf, out)
    out = torch.from_numpy(out).reshape(n, hp, wp, cp)
    assert not torch.isnan(out).any()
    torch.testing.assert_close(out[..., :c].permute(0, 3, 1, 2), ref, rtol=1e-9, atol=1e-9)
    assert (out[..., c:] == 0).all()


def test_conv_transpose_forward_is_dgrad(fpglib):
    """nn.ConvTranspose2d(k3, s2, p1, op1) forward == fpg_conv2d_dgrad with the weight read as [c_out][c_in][r][s]."""
    torch.manual_seed(2)
    n, ci, co, h = 1, 64, 128, 8
    x = torch.randn(n, ci, h, h, dtype=torch.float64)
    wt = torch.randn(ci, co, 3, 3, dtype=torch.float64)  # ConvTranspose2d layout [Cin_T][Cout_T][R][S]
    ref = F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1)
    g = geom(3, 3, 2, 1, co, ci)  # equivalent forward conv: c_in = Cout_T, c_out = Cin_T
    dya = make_act(n, h, h, ci)
    dxa = make_act(n, 2 * h, 2 * h, co)
    descs = (L.FpropDesc * 4)()
    nd = C.c_int()
    L.call("fpg_conv2d_dgrad_plan", C.byref(dya), interp.FAKE_BASE, None, L.ACT_NONE, C.byref(g), C.byref(dxa), SMS,
           descs, C.byref(nd))
    out = np.full(n * 4 * h * h * co, np.nan)
    for q in range(nd.value):
        interp.run_fprop(descs[q], nhwc_buffer(x, ci), pack_dgrad(fpglib, wt, g), out)
    out = torch.from_numpy(out).reshape(n, 2 * h, 2 * h, co).permute(0, 3, 1, 2)
    torch.testing.assert_close(out, ref, rtol=1e-9, atol=1e-9)


WGRAD_CASES = [
    (1, 8, 8, 64, 64, 128, 128, 3, 1, 0, 1),    # residual conv (reflect halo)
    (1, 16, 16, 64, 64, 128, 128, 3, 2, 1, 0),  # stride-2 3x3
    (1, 16, 16, 9, 16, 64, 64, 7, 1, 0, 3),     # stem: 16-channel input, taps-as-atoms
    (1, 16, 16, 12, 16, 64, 64, 4, 2, 1, 0),    # PatchGAN model.0
    (1, 12, 12, 64, 64, 27, 32, 7, 1, 0, 3),    # content head: small c_out
    (1, 8, 8, 64, 64, 10, 16, 1, 1, 0, 0),      # attention head 1x1
    (1, 7, 7, 128, 128, 1, 16, 4, 1, 1, 0),     # model.11
    (2, 9, 9, 64, 64, 128, 128, 4, 1, 1, 0),    # 4x4 s1 p1, odd extent
    (1, 8, 8, 64, 64, 256, 256, 3, 1, 0, 1),    # c_out 256: M = 256 tile (two MMAs sharing the input tile)
    # shifted operands (64-pixel rows): column taps read from ONE box at shifted pixel rows
    (1, 3, 64, 128, 128, 256, 256, 3, 1, 0, 1),  # residual conv: S shift groups, M = N = 128, 2x2 channel groups
    (2, 4, 64, 9, 16, 64, 64, 7, 1, 0, 3),       # stem: dy row pairs x 7 shift atoms of the 16-channel input
    (1, 5, 58, 64, 64, 27, 32, 7, 1, 0, 3),      # content head: input row pairs x 7 shift atoms of dy (wp = 64)
    (1, 4, 130, 64, 64, 3, 16, 7, 1, 0, 3),      # tanh head (3 -> 16 channels), ragged row of 136 pixels
    (1, 6, 64, 27, 32, 64, 64, 3, 1, 1, 0),      # zero-padded 3x3 on 32 input channels: shift atoms with pad
]


@pytest.mark.parametrize("case", [(2, 16, 16, 256, 256, 256, 256, 3, 1, 0, 1),  # residual conv of the 256x256 step
                                  (1, 8, 8, 256, 256, 256, 256, 2, 1, 0, 0),     # even tap count
                                  (1, 11, 11, 256, 256, 256, 256, 4, 1, 1, 0)])  # 16 taps, zero pad, ragged k tiles
def test_wgrad_pair_plan_matches_autograd(fpglib, case, monkeypatch):
    """256 x 256-channel layers plan the CTA-pair kernel (an item = two taps, D[256, 2 x 256])."""
    monkeypatch.delenv("FPG_WGRAD_SHIFT_WIDE", raising=False)
    d = _check_wgrad_plan(fpglib, case)
    assert d.cta_pair == 1 and d.y_sets == 2 and d.stages <= 4
    monkeypatch.setenv("FPG_DISABLE_WGRAD_PAIR", "1")
    assert _check_wgrad_plan(fpglib, case, run=False).cta_pair == 0


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad_plan_matches_autograd(fpglib, case, monkeypatch):
    monkeypatch.setenv("FPG_WGRAD_SHIFT_WIDE", "1")  # also cover the opt-in shift-group plan of the wide layers
    n, h, w, c, cp, k, kp, r, stride, pad, halo = case
    d = _check_wgrad_plan(fpglib, case)
    if w >= 58 and stride == 1 and r > 1:
        assert d.y_shift_atoms or d.y_shifts > 1, "wide stride-1 layers must plan shifted operands"


def _check_wgrad_plan(fpglib, case, run=True):
    n, h, w, c, cp, k, kp, r, stride, pad, halo = case
    torch.manual_seed(3)
    x = torch.randn(n, c, h, w, dtype=torch.float64)
    xin = F.pad(x, (halo,) * 4, "reflect") if halo else x
    wt = torch.randn(k, c, r, r, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xin, wt, None, stride=stride, padding=pad)
    dy = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, wt, dy)
    ho, wo = y.shape[2:]
    g = geom(r, r, stride, pad, cp, kp)
    xa = make_act(n, h, w, cp, halo=halo)
    dya = make_act(n, ho, wo, kp)
    d = L.WgradDesc()
    L.call("fpg_conv2d_wgrad_plan", C.byref(xa), C.byref(dya), C.byref(g), SMS, C.byref(d))
    assert fpglib.fpg_conv2d_wgrad_ws_bytes(C.byref(xa), C.byref(dya), C.byref(g), SMS) > 0
    if not run:
        return d
    dw = np.full(k * c * r * r, np.nan)
    xb, yb = nhwc_buffer(x, cp, halo), nhwc_buffer(dy, kp)
    if d.x_is_dy:
        interp.run_wgrad(d, yb, xb, dw, c * r * r, r * r, k, c)
    else:
        interp.run_wgrad(d, xb, yb, dw, c * r * r, r * r, k, c)
    assert not np.isnan(dw).any()
    torch.testing.assert_close(torch.from_numpy(dw).reshape(k, c, r, r), ref, rtol=1e-9, atol=1e-9)
    return d


def test_library_exports_every_declared_symbol(fpglib):
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "fpg.h")).read()
    declared = set(re.findall(r"\b(fpg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(fpglib, name), f"{name} declared in include/fpg.h but not exported"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)


@pytest.mark.parametrize("kind", ["fprop", "dgrad_s1_halo", "dgrad_s2"])
def test_cta_pair_plans(fpglib, kind):
    """2-CTA (cta_group::2) descriptors: pair tiles, half weight boxes, odd number of row tiles (masked dummy half).
    A small sm_count makes the planner choose pairing on shapes the interpreter can afford."""
    torch.manual_seed(4)
    sms = 2
    if kind == "fprop":
        n, h, w, c, k, r, stride, pad, halo = 1, 24, 16, 64, 128, 3, 1, 0, 1   # 24 rows -> 3 row tiles of 8: odd
        x = torch.randn(n, c, h, w, dtype=torch.float64)
        wt = torch.randn(k, c, r, r, dtype=torch.float64)
        xin = F.pad(x, (halo,) * 4, "reflect")
        ref = F.conv2d(xin, wt, None, stride=stride, padding=pad)
        xa, ya, g = make_act(n, h, w, c, halo=halo), make_act(n, h, w, k), geom(r, r, stride, pad, c, k)
        d = L.FpropDesc()
        L.call("fpg_conv2d_fprop_plan", C.byref(xa), interp.FAKE_BASE, None, L.ACT_NONE, C.byref(g), C.byref(ya), sms,
               C.byref(d))
        assert d.cta_pair == 1 and int(d.b.box[1]) * 2 == d.block_n
        out = np.full(n * h * w * k, np.nan)
        interp.run_fprop(d, nhwc_buffer(x, c, halo), pack_fprop(wt, c, k, d.num_taps), out)
        got = torch.from_numpy(out).reshape(n, h, w, k).permute(0, 3, 1, 2)
        torch.testing.assert_close(got, ref, rtol=1e-9, atol=1e-9)
        return
    if kind == "dgrad_s1_halo":
        n, h, w, c, k, r, stride, pad, halo = 1, 22, 14, 128, 64, 3, 1, 0, 1
    else:
        n, h, w, c, k, r, stride, pad, halo = 1, 32, 32, 128, 64, 4, 2, 1, 0
    xin = torch.randn(n, c, h + 2 * halo, w + 2 * halo, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(k, c, r, r, dtype=torch.float64)
    y = F.conv2d(xin, wt, None, stride=stride, padding=pad)
    dy = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, xin, dy)
    g = geom(r, r, stride, pad, c, k)
    dya, dxa = make_act(n, y.shape[2], y.shape[3], k), make_act(n, h, w, c, halo=halo)
    descs = (L.FpropDesc * 4)()
    nd = C.c_int()
    L.call("fpg_conv2d_dgrad_plan", C.byref(dya), interp.FAKE_BASE, None, L.ACT_NONE, C.byref(g), C.byref(dxa), sms,
           descs, C.byref(nd))
    assert all(descs[q].cta_pair == 1 for q in range(nd.value))
    hp, wp = h + 2 * halo, w + 2 * halo
    out = np.full(n * hp * wp * c, np.nan)
    for q in range(nd.value):
        interp.run_fprop(descs[q], nhwc_buffer(dy, k), pack_dgrad(fpglib, wt, g), out)
    out = torch.from_numpy(out).reshape(n, hp, wp, c)
    assert not torch.isnan(out).any()
    torch.testing.assert_close(out.permute(0, 3, 1, 2), ref, rtol=1e-9, atol=1e-9)


ROWS_CASES = [
    # (kind, n, h, w, c_real, c_pad, k_real, k_pad, r, zero_pad, reflect_halo) of the FORWARD conv (stride 1)
    ("fprop", 1, 6, 100, 9, 16, 64, 64, 7, 0, 3),     # generator stem on a wide image: resident filter, SW32
    ("fprop", 2, 5, 130, 64, 64, 27, 32, 7, 0, 3),    # content head: two column tiles, filter-row ring
    ("fprop", 1, 9, 128, 27, 32, 64, 64, 3, 1, 0),    # zero padding: negative offsets / TMA zero fill, SW64
    ("dgrad", 1, 6, 96, 64, 64, 27, 32, 7, 0, 3),     # content head data gradient incl. halo (dy has 32 channels)
    ("dgrad", 1, 7, 100, 9, 16, 64, 64, 7, 0, 3),     # stem data gradient (cycle models): N = 16
    ("dgrad", 1, 10, 120, 64, 64, 32, 32, 3, 1, 0),   # zero-padded 3x3
]


@pytest.mark.parametrize("case", ROWS_CASES)
def test_rows_plan_matches_torch(fpglib, case):
    """Row-stationary descriptors (igemm_rows.cu) interpreted on the CPU against conv2d / its input gradient."""
    kind, n, h, w, c, cp, k, kp, r, pad, halo = case
    torch.manual_seed(5)
    xin = torch.randn(n, c, h + 2 * halo, w + 2 * halo, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(k, c, r, r, dtype=torch.float64)
    bias = torch.randn(kp, dtype=torch.float64)
    bias[k:] = 0
    y = F.conv2d(xin, wt, None, stride=1, padding=pad)
    ho, wo = y.shape[2:]
    g = geom(r, r, 1, pad, cp, kp)
    d = L.RowsDesc()
    if kind == "fprop":
        xa, ya = make_act(n, h, w, cp, halo=halo), make_act(n, ho, wo, kp, halo=1)
        rc = fpglib.fpg_conv2d_rows_plan(C.byref(xa), interp.FAKE_BASE, None, L.ACT_LEAKY, C.byref(g), C.byref(ya), 0,
                                         SMS, C.byref(d))
        assert rc == 0
        abuf = torch.zeros(n, h + 2 * halo, w + 2 * halo, cp, dtype=torch.float64)
        abuf[..., :c] = xin.detach().permute(0, 2, 3, 1)
        bbuf = pack_fprop(wt, cp, kp, int(d.b.dims[0]) // cp)
        assert bbuf.size * 2 == fpglib.fpg_packed_weight_bytes(C.byref(g))
        out = np.full(n * (ho + 2) * (wo + 2) * kp, np.nan)
        interp.run_rows(d, abuf.reshape(-1).numpy(), bbuf, out, bias.numpy())
        out = torch.from_numpy(out).reshape(n, ho + 2, wo + 2, kp)
        ref = F.leaky_relu(y.detach() + bias[:k].view(1, -1, 1, 1), 0.2)
        torch.testing.assert_close(out[:, 1:-1, 1:-1, :k].permute(0, 3, 1, 2), ref, rtol=1e-9, atol=1e-9)
        assert torch.isnan(out[:, 0]).all() and torch.isnan(out[:, :, 0]).all()
        assert (out[:, 1:-1, 1:-1, k:] == 0).all()
    else:
        dy = torch.randn_like(y)
        (ref,) = torch.autograd.grad(y, xin, dy)
        dya, dxa = make_act(n, ho, wo, kp), make_act(n, h, w, cp, halo=halo)
        rc = fpglib.fpg_conv2d_rows_plan(C.byref(dya), interp.FAKE_BASE, None, L.ACT_NONE, C.byref(g), C.byref(dxa), 1,
                                         SMS, C.byref(d))
        assert rc == 0
        hp, wp = h + 2 * halo, w + 2 * halo
        out = np.full(n * hp * wp * cp, np.nan)
        interp.run_rows(d, nhwc_buffer(dy, kp), pack_dgrad(fpglib, wt, g), out)
        out = torch.from_numpy(out).reshape(n, hp, wp, cp)
        assert not torch.isnan(out).any()
        torch.testing.assert_close(out[..., :c].permute(0, 3, 1, 2), ref, rtol=1e-9, atol=1e-9)
        assert (out[..., c:] == 0).all()


def test_rows_plan_declines_other_shapes(fpglib):
    """Narrow images, strided, 1x1 and wide-channel convolutions stay on the tiled kernel (return code 1)."""
    d = L.RowsDesc()
    for (h, w, cp, kp, r, stride, pad, halo) in [(64, 64, 64, 64, 3, 1, 0, 1), (256, 256, 16, 64, 4, 2, 1, 0),
                                                 (256, 256, 64, 16, 1, 1, 0, 0), (128, 128, 128, 64, 3, 1, 1, 0)]:
        g = geom(r, r, stride, pad, cp, kp)
        hp, wp = h + 2 * halo, w + 2 * halo
        ho, wo = (hp + 2 * pad - r) // stride + 1, (wp + 2 * pad - r) // stride + 1
        xa, ya = make_act(1, h, w, cp, halo=halo), make_act(1, ho, wo, kp)
        assert fpglib.fpg_conv2d_rows_plan(C.byref(xa), interp.FAKE_BASE, None, 0, C.byref(g), C.byref(ya), 0, SMS,
                                           C.byref(d)) == 1
