"""GPU parity tests of the individual sm_100a kernels, called through the C ABI (fpgan.ops -> libfpg_b200.so),
against plain fp32 PyTorch on the same (bf16-rounded) inputs."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def bf16r(t):
    return t.to(torch.bfloat16).float()


def close_rms(got, ref, max_tol, mean_tol, what=""):
    """Abs error relative to the RMS of the reference. bf16 storage rounds every element to 2^-9 relative, so
    the element-wise bound also allows 2^-7 of the element's own magnitude (heavy-tailed values)."""
    rms = ref.float().pow(2).mean().sqrt().item() + 1e-12
    err = (got.float() - ref.float()).abs()
    mean = err.mean().item() / rms
    mx = ((err - ref.float().abs() * 2.0 ** -7).clamp_min(0)).max().item() / rms
    assert mx < max_tol and mean < mean_tol, f"{what}: max {mx:.4g} (tol {max_tol}) mean {mean:.4g} (tol {mean_tol})"


CONV_CASES = [
    # (n, h, w, c_real, k_real, r, stride, zero_pad, reflect_halo)
    (6, 64, 64, 256, 256, 3, 1, 0, 1),    # residual conv: 192 tiles > 148 SMs (persistent loop, both TMEM stages)
    (1, 64, 64, 256, 256, 3, 1, 0, 1),    # B=1: block_n shrinks to fill the SMs
    (2, 64, 64, 9, 64, 7, 1, 0, 3),       # stem, 16-channel SW32 path
    (2, 64, 64, 64, 128, 3, 2, 1, 0),     # downsample s2
    (2, 32, 32, 128, 256, 3, 2, 1, 0),
    (2, 64, 64, 64, 27, 7, 1, 0, 3),      # content head
    (2, 64, 64, 64, 10, 1, 1, 0, 0),      # attention head
    (2, 64, 64, 12, 64, 4, 2, 1, 0),      # PatchGAN model.0
    (2, 32, 32, 128, 256, 4, 2, 1, 0),    # model.5
    (2, 32, 32, 256, 512, 4, 1, 1, 0),    # model.8 (31x31 output, masked tiles, 2 n-blocks)
    (2, 31, 31, 512, 1, 4, 1, 1, 0),      # model.11 (30x30 output)
    (1, 8, 8, 27, 64, 3, 1, 1, 0),        # 32-channel input (SW64 path)
    # wide images: the 7x7 layers run on the row-stationary kernel (igemm_rows.cu), fprop and dgrad
    (3, 42, 256, 9, 64, 7, 1, 0, 3),      # stem at full width: resident filter; 3*2*11 = 66 tiles
    (2, 130, 256, 64, 27, 7, 1, 0, 3),    # content head: filter-row ring, 132 tiles, ragged last row tile
    (1, 20, 200, 27, 64, 3, 1, 1, 0),     # zero-padded 3x3 on a ragged width (two column tiles, SW64)
    (19, 64, 128, 64, 27, 7, 1, 0, 3),    # 304 tiles > 148 SMs: persistent loop over both accumulator stages
    (16, 64, 64, 256, 256, 3, 1, 0, 1),   # B=16 residual conv: dgrad over the 66x66 haloed grid uses two tile regions
]


def _conv_setup(case, seed):
    from fpgan import ops
    n, h, w, c, k, r, stride, pad, halo = case
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = bf16r(torch.randn(n, c, h, w, device="cuda", generator=g))
    wt = bf16r(torch.randn(k, c, r, r, device="cuda", generator=g) * (1.0 / (c * r * r) ** 0.5))
    spec = ops.ConvSpec(r, r, stride, pad, ops.pad16(c), ops.pad16(k), c_in_valid=c, c_out_valid=k)
    spec.pack(wt.contiguous())
    xin = F.pad(x, (halo,) * 4, "reflect") if halo else x
    return ops, x, wt, spec, xin


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fprop(case):
    ops, x, wt, spec, xin = _conv_setup(case, 0)
    n, h, w, c, k, r, stride, pad, halo = case
    bias = torch.randn(ops.pad16(k), device="cuda")
    bias[k:] = 0
    ref = F.conv2d(xin, wt, bias[:k], stride=stride, padding=pad)
    ho, wo = ref.shape[2:]
    xb = ops.ActBuf.from_nchw(x, halo=halo)
    for fp32 in (False, True):
        yb = ops.ActBuf(n, ho, wo, ops.pad16(k), halo=1, fp32=fp32)
        yb.t.fill_(7.0)
        ops.conv_fprop(xb, spec, yb, bias=bias, act=ops.ACT_LEAKY)
        torch.cuda.synchronize()
        got = yb.to_nchw(k)
        close_rms(got, F.leaky_relu(ref, 0.2), 0.03 if not fp32 else 2e-3, 0.004 if not fp32 else 2e-4,
                  f"fprop {case} fp32={fp32}")
        assert (yb.t[:, 0] == 7.0).all() and (yb.t[:, :, -1] == 7.0).all(), "output halo was overwritten"
        assert (yb.interior()[..., k:] == 0).all(), "padded output channels must stay zero"


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_dgrad(case):
    ops, x, wt, spec, xin = _conv_setup(case, 1)
    n, h, w, c, k, r, stride, pad, halo = case
    xin = xin.clone().requires_grad_(True)
    y = F.conv2d(xin, wt, None, stride=stride, padding=pad)
    dy = bf16r(torch.randn_like(y))
    (ref,) = torch.autograd.grad(y, xin, dy)
    dyb = ops.ActBuf.from_nchw(dy)
    dxb = ops.ActBuf(n, h, w, ops.pad16(c), halo=halo)
    dxb.t.fill_(float("nan"))
    ops.conv_dgrad(dyb, spec, dxb)
    torch.cuda.synchronize()
    got = dxb.t[..., :c].permute(0, 3, 1, 2).float()
    assert not torch.isnan(dxb.t).any()
    close_rms(got, ref, 0.03, 0.004, f"dgrad {case}")
    assert (dxb.t[..., c:] == 0).all()


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_wgrad(case):
    ops, x, wt, spec, xin = _conv_setup(case, 2)
    n, h, w, c, k, r, stride, pad, halo = case
    wt = wt.clone().requires_grad_(True)
    y = F.conv2d(xin, wt, None, stride=stride, padding=pad)
    dy = bf16r(torch.randn_like(y))
    (ref,) = torch.autograd.grad(y, wt, dy)
    xb = ops.ActBuf.from_nchw(x, halo=halo)
    dyb = ops.ActBuf.from_nchw(dy)
    dw = torch.full_like(ref, float("nan"))
    ops.conv_wgrad(xb, dyb, spec, dw)
    torch.cuda.synchronize()
    assert not torch.isnan(dw).any()
    close_rms(dw, ref, 5e-3, 5e-4, f"wgrad {case}")


def test_conv_transpose_roundtrip():
    """ConvTranspose2d(3, s2, p1, op1) forward / input-grad / weight-grad through dgrad / fprop / wgrad."""
    from fpgan import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    n, ci, co, h = 2, 256, 128, 64
    x = bf16r(torch.randn(n, ci, h, h, device="cuda", generator=g)).requires_grad_(True)
    wt = bf16r(torch.randn(ci, co, 3, 3, device="cuda", generator=g) / 48).requires_grad_(True)
    y = F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1)
    dy = bf16r(torch.randn_like(y))
    dx_ref, dw_ref = torch.autograd.grad(y, (x, wt), dy)
    spec = ops.ConvSpec(3, 3, 2, 1, co, ci)  # equivalent forward conv: c_in = Cout_T, c_out = Cin_T
    spec.pack(wt.detach().contiguous())
    xb = ops.ActBuf.from_nchw(x.detach())
    yb = ops.ActBuf(n, 2 * h, 2 * h, co)
    ops.conv_dgrad(xb, spec, yb)
    close_rms(yb.to_nchw(), y.detach(), 0.03, 0.004, "convT forward")
    dyb = ops.ActBuf.from_nchw(dy)
    dxb = ops.ActBuf(n, h, h, ci)
    ops.conv_fprop(dyb, spec, dxb)
    close_rms(dxb.to_nchw(), dx_ref, 0.03, 0.004, "convT input grad")
    dw = torch.empty_like(dw_ref)
    ops.conv_wgrad(dyb, xb, spec, dw)
    close_rms(dw, dw_ref, 5e-3, 5e-4, "convT weight grad")


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("shape,halo,act", [((3, 64, 64, 256), 1, 1), ((2, 128, 128, 64), 3, 1),
                                            ((2, 31, 31, 512), 0, 2), ((2, 32, 32, 128), 0, 0),
                                            ((2, 16, 16, 16), 0, 1)])
def test_instnorm_fwd_bwd(shape, halo, act, f16):
    """f16: the pre-norm tensor y and the residual are fp16 buffers (FPG_DT_FP16), as in the generator trunk"""
    from fpgan import ops
    n, h, w, c = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    rnd = (lambda t: t.half().float()) if f16 else bf16r
    y = rnd(torch.randn(n, c, h, w, device="cuda", generator=g) * 3 + 0.5).requires_grad_(True)
    res = rnd(torch.randn(n, c, h, w, device="cuda", generator=g))
    fn = {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.2)}[act]
    zi = fn(F.instance_norm(y, eps=1e-5)) + res
    z = F.pad(zi, (halo,) * 4, "reflect") if halo else zi
    dz = bf16r(torch.randn_like(z))
    dz2 = bf16r(torch.randn_like(zi))
    (dy_ref,) = torch.autograd.grad([z, zi], y, [dz, dz2])

    yb = ops.ActBuf.from_nchw(y.detach(), f16=f16)
    rb = ops.ActBuf.from_nchw(res, f16=f16)
    zb = ops.ActBuf(n, h, w, c, halo=halo)
    skip = ops.ActBuf(n, h, w, c, f16=True, zero=False)
    stats = torch.empty(n * c * 2, device="cuda")
    ops.instnorm_stats(yb, stats)
    ops.instnorm_apply(yb, stats, act, zb, residual=rb, skip_out=skip)
    # the skip-stream copy: same values, rounded to fp16 instead of bf16 (8x finer)
    close_rms(skip.to_nchw(), zi.detach(), 0.003, 0.0005, "instnorm apply: fp16 skip stream")
    st = stats.view(n, c, 2)
    torch.testing.assert_close(st[..., 0], y.detach().mean((2, 3)), rtol=1e-3, atol=2e-3)
    torch.testing.assert_close(st[..., 1], (y.detach().var((2, 3), unbiased=False) + 1e-5).rsqrt(), rtol=1e-3,
                               atol=1e-4)
    got = zb.t.permute(0, 3, 1, 2).float()
    close_rms(got, z.detach(), 0.02, 0.003, "instnorm apply (incl. halo)")

    def fresh_dz():  # fpg_instnorm_bwd consumes dz (folds its mirror band in place)
        b = ops.ActBuf(n, h, w, c, halo=halo)
        b.t.copy_(dz.permute(0, 2, 3, 1))
        return b

    dzb = fresh_dz()
    dz2b = ops.ActBuf.from_nchw(dz2)
    dyb = ops.ActBuf(n, h, w, c)
    dres = ops.ActBuf(n, h, w, c)
    ops.instnorm_bwd(dzb, yb, stats, act, dyb, dz2=dz2b, dres=dres)
    close_rms(dyb.to_nchw(), dy_ref, 0.03, 0.004, "instnorm bwd")
    # dres = fold(dz) + dz2 == gradient w.r.t. the residual input
    res_g = res.clone().requires_grad_(True)
    zi2 = fn(F.instance_norm(y.detach(), eps=1e-5)) + res_g
    z2 = F.pad(zi2, (halo,) * 4, "reflect") if halo else zi2
    (dres_ref,) = torch.autograd.grad([z2, zi2], res_g, [dz, dz2])
    close_rms(dres.to_nchw(), dres_ref, 0.02, 0.003, "halo fold")
    # without dres the apply pass recomputes the fold
    dyb2 = ops.ActBuf(n, h, w, c)
    ops.instnorm_bwd(fresh_dz(), yb, stats, act, dyb2, dz2=dz2b)
    close_rms(dyb2.to_nchw(), dy_ref, 0.03, 0.004, "instnorm bwd (no dres)")
    out = ops.ActBuf(n, h, w, c)
    ops.halo_fold(fresh_dz(), dz2b, out)
    close_rms(out.to_nchw(), dres_ref, 0.02, 0.003, "halo_fold op")


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("act,with_add,batch", [(1, False, 2), (0, True, 2), (1, True, 2), (0, True, 16), (1, False, 16)])
def test_dgrad_with_instnorm_backward_statistics(act, with_add, batch, f16):
    """conv_dgrad_inbwd + instnorm_bwd_apply == autograd of conv(reflect_pad(z)) w.r.t. y (and the block input), where
    z = relu(IN(y)) (act 1) or z = zprev + IN(y) (act 0: residual block output): the reduction pass of the InstanceNorm
    backward runs in the data-gradient kernel's epilogue, fed by the saved conv input z (and zprev) staged through a
    shared-memory ring; the skip gradient is merged on the interior. batch 16 = the benchmarked shape (two tile regions,
    block_n 256); f16: the pre-norm tensor read by the apply pass is an fp16 buffer."""
    from fpgan import ops
    n, c, k, h = batch, 256, 256, 64
    g = torch.Generator(device="cuda").manual_seed(11)
    y = (torch.randn(n, c, h, h, device="cuda", generator=g) * 2 + 0.3)
    y = (y.half().float() if f16 else bf16r(y)).requires_grad_(True)
    zprev = bf16r(torch.randn(n, c, h, h, device="cuda", generator=g)).requires_grad_(True)
    wt = bf16r(torch.randn(k, c, 3, 3, device="cuda", generator=g) / 48)
    zhat = F.instance_norm(y, eps=1e-5)
    z = F.relu(zhat) if act == 1 else zprev + zhat
    out = F.conv2d(F.pad(z, (1,) * 4, "reflect"), wt)
    dout = bf16r(torch.randn_like(out) * 0.1)
    dskip_up = bf16r(torch.randn(n, c, h, h, device="cuda", generator=g) * 0.1)  # gradient arriving on z directly
    ups = [dout, dskip_up] if with_add else [dout]
    outs = [out, z] if with_add else [out]
    dy_ref, dprev_ref = torch.autograd.grad(outs, (y, zprev), ups, allow_unused=True)

    spec = ops.ConvSpec(3, 3, 1, 0, c, k)
    spec.pack(wt.contiguous())
    yb = ops.ActBuf.from_nchw(y.detach(), f16=f16)
    stats = torch.empty(n * c * 2, device="cuda")
    ops.instnorm_stats(yb, stats)
    zb = ops.ActBuf.from_nchw(z.detach(), halo=1)          # the convolution's saved input (bf16, reflect halo)
    pb = ops.ActBuf.from_nchw(zprev.detach(), halo=1) if act == 0 else None
    dyb = ops.ActBuf.from_nchw(dout)
    dx = ops.ActBuf(n, h, h, c, halo=1, zero=False)
    add = None
    if with_add:  # the skip gradient lives in the interior of a haloed buffer, as in the trunk backward
        add = ops.ActBuf(n, h, h, c, halo=1)
        add.t.normal_()  # the halo of that buffer holds stale values: they must not be read as gradient
        add.t[:, 1:-1, 1:-1, :] = dskip_up.permute(0, 2, 3, 1)
    red = ops.conv_dgrad_inbwd(dyb, spec, dx, zb, pb, add, force=True)
    if red is None and os.environ.get("FPG_DISABLE_TILE_REGIONS"):
        pytest.skip("the single-tile-shape plan of the haloed gradient has no statistics epilogue (caller falls back)")
    assert red is not None, "the residual conv must plan the statistics epilogue"
    dy = ops.ActBuf(n, h, h, c, zero=False)
    ops.instnorm_bwd_apply(dx, yb, stats, red, act, dy)
    close_rms(dy.to_nchw(), dy_ref, 0.03, 0.004, "fused instnorm bwd")
    if act == 0:  # dx's interior now holds the total gradient w.r.t. z = gradient w.r.t. the block input zprev
        close_rms(dx.t[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).float(), dprev_ref, 0.02, 0.003, "merged skip gradient")
    # and the unfused path agrees
    dx2 = ops.ActBuf(n, h, h, c, halo=1, zero=False)
    ops.conv_dgrad(dyb, spec, dx2)
    dy2 = ops.ActBuf(n, h, h, c, zero=False)
    ops.instnorm_bwd(dx2, yb, stats, act, dy2, dz2=ops.ActBuf.from_nchw(dskip_up) if with_add else None)
    # (one bf16 rounding of dz + skip here, two there)
    close_rms(dy.to_nchw(), dy2.to_nchw(), 0.03, 0.004, "fused vs two-pass")


def _blend_ref(content_pre, logits, image):
    """model_architectures.py:353-399 restated with torch ops"""
    content = torch.tanh(content_pre)
    att = torch.softmax(logits, dim=1)
    out = image * att[:, 9:10]
    for k in range(9):
        out = out + content[:, 3 * k:3 * k + 3] * att[:, k:k + 1]
    return out, att[:, 9]


def test_blend_fwd_bwd():
    from fpgan import ops
    n, h, w = 2, 32, 48
    g = torch.Generator(device="cuda").manual_seed(11)
    cpre = torch.randn(n, 27, h, w, device="cuda", generator=g).requires_grad_(True)
    logits = (torch.randn(n, 10, h, w, device="cuda", generator=g) * 2).requires_grad_(True)
    image = bf16r(torch.rand(n, 9, h, w, device="cuda", generator=g) * 2 - 1)
    img_rgb = image[:, :3].clone().requires_grad_(True)
    out_ref, mask_ref = _blend_ref(cpre, logits, img_rgb)
    dout = torch.randn_like(out_ref)
    dc_ref, dl_ref, dimg_ref = torch.autograd.grad(out_ref, (cpre, logits, img_rgb), dout)

    cb = ops.ActBuf.from_nchw(torch.tanh(cpre.detach()), c_pad=32, fp32=True)
    lb = ops.ActBuf.from_nchw(logits.detach(), c_pad=16, fp32=True)
    ib = ops.ActBuf.from_nchw(image, halo=3)
    ob = ops.ActBuf(n, h, w, 16)
    out_nchw = torch.empty(n, 3, h, w, device="cuda")
    mask = torch.empty(n, h, w, device="cuda")
    ops.blend_fwd(cb, lb, ib, out=ob, out_c0=9, out_nchw=out_nchw, mask=mask)
    torch.testing.assert_close(out_nchw, out_ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(mask, mask_ref.detach(), rtol=1e-4, atol=1e-6)
    close_rms(ob.interior()[..., 9:12].permute(0, 3, 1, 2), out_ref.detach(), 0.01, 0.002, "blend bf16 out")
    assert (ob.interior()[..., :9] == 0).all() and (ob.interior()[..., 12:] == 0).all()

    dcb = ops.ActBuf(n, h, w, 32)
    dlb = ops.ActBuf(n, h, w, 16)
    dimg = torch.empty(n, 3, h, w, device="cuda")
    # split the upstream gradient over the two supported sources
    half = bf16r(dout * 0.5)
    gb = ops.ActBuf(n, h, w, 16)
    gb.t[..., 9:12] = half.permute(0, 2, 3, 1).to(torch.bfloat16)
    ops.blend_bwd(cb, lb, ib, dcb, dlb, dout_nchw=(dout - half).contiguous(), dout_nhwc=gb, dout_c0=9,
                  dimage_nchw=dimg)
    close_rms(dcb.to_nchw(27), dc_ref, 0.02, 0.003, "blend dcontent")
    close_rms(dlb.to_nchw(10), dl_ref, 0.02, 0.003, "blend dlogits")
    torch.testing.assert_close(dimg, dimg_ref, rtol=1e-4, atol=1e-5)
    assert (dcb.t[..., 27:] == 0).all() and (dlb.t[..., 10:] == 0).all()


def test_losses():
    from fpgan import ops
    g = torch.Generator(device="cuda").manual_seed(13)
    n, h, w = 4, 30, 30
    logit = torch.randn(n, 1, h, w, device="cuda", generator=g).requires_grad_(True)
    for target in (0.0, 1.0):
        ref = F.mse_loss(logit, torch.full_like(logit, target))
        (dref,) = torch.autograd.grad(ref * 0.5, logit)
        lb = ops.ActBuf.from_nchw(logit.detach(), c_pad=16, fp32=True)
        db = ops.ActBuf(n, h, w, 16)
        db.t.fill_(3.0)
        loss = torch.zeros(1, device="cuda")
        ops.mse_const_loss(lb, target, 1.0, 0.5, loss, dlogits=db)
        torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-5, atol=1e-7)
        close_rms(db.to_nchw(1), dref, 0.01, 0.003, "mse grad")
        assert (db.t[..., 1:] == 0).all()
    pred = torch.randn(3, 3, 64, 64, device="cuda", generator=g).requires_grad_(True)
    tgt = torch.randn(3, 3, 64, 64, device="cuda", generator=g)
    ref = F.l1_loss(pred, tgt) * 100
    (dref,) = torch.autograd.grad(ref, pred)
    loss = torch.zeros(1, device="cuda")
    dp = torch.ones_like(tgt)
    ops.l1_loss(pred.detach(), tgt, 100.0, 1.0, loss, dpred=dp, accumulate=True)
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dp, dref + 1.0, rtol=1e-6, atol=1e-9)


def test_adam_matches_torch():
    from fpgan import ops
    g = torch.Generator(device="cuda").manual_seed(17)
    p = torch.randn(100003, device="cuda", generator=g)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=2e-4, betas=(0.5, 0.999))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn(p.shape, device="cuda", generator=g) * 10 ** (step - 3)
        ref_p.grad = grad.clone()
        opt.step()
        ops.adam_step(p, grad, m, v, 2e-4, 0.5, 0.999, 1e-8, step)
        torch.testing.assert_close(p, ref_p.detach(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("family", ["pairedattention", "pix2pix"])
def test_fused_adam_and_repack_equals_the_two_launches(family):
    """fpg_adam_pack_step (Adam + every bf16 operand layout + padded bias vectors in one launch) against
    fpg_adam_step_dev followed by fpg_pack_weights_batched: parameters, both moments and every packed operand
    bit-identical over three steps, also with the gradient given as three sources summed in order."""
    from fpgan import networks, ops
    from fpgan.trainer import FlatParams
    from models import model_architectures as A
    import ctypes as C

    def build():
        torch.manual_seed(5)
        if family == "pix2pix":
            mods = [A.Pix2PixGenerator(9).cuda(), A.Pix2PixDiscriminator(9).cuda()]
        else:
            mods = [A.PairedAttentionGenerator(9).cuda(), A.PairedAttentionDiscriminator(9).cuda()]
        fp = FlatParams(*mods)
        return fp, [m._executor() for m in mods]

    fa, ea = build()
    fb, eb = build()
    assert torch.equal(fa.flat, fb.flat)
    fused = networks.AdamPack(fb, eb)
    for ex in ea:
        ex.repack(force=True)
    g = torch.Generator(device="cuda").manual_seed(3)
    for step in range(3):
        parts = [torch.randn(fa.flat.shape, device="cuda", generator=g) * 10.0 ** (step - 2) for _ in range(3)]
        total = (parts[0] + parts[1]) + parts[2]
        fa.set_lr(2e-4)
        fb.set_lr(2e-4)
        fa.grads.flat.copy_(total)
        fa.adam(grad_scale=0.5)
        for ex in ea:
            ex.repack(force=True)
        if step == 1:
            srcs = (C.c_void_p * 3)(*[t.data_ptr() for t in parts])
            gsum = torch.empty_like(total)
            fused.run(0.5, sources=srcs, n_src=3, gsum=gsum)
            assert torch.equal(gsum, total)
        else:
            fb.grads.flat.copy_(total)
            fused.run(0.5)
        torch.cuda.synchronize()
        assert torch.equal(fa.flat, fb.flat) and torch.equal(fa.m, fb.m) and torch.equal(fa.v, fb.v), step
        assert torch.equal(fa.state, fb.state)
        for xa, xb in zip(ea, eb):
            for name, la in xa.layers.items():
                lb = xb.layers[name]
                assert torch.equal(la.spec.w_fprop, lb.spec.w_fprop), (step, name, "fprop operand")
                assert torch.equal(la.spec.w_dgrad, lb.spec.w_dgrad), (step, name, "dgrad operand")
                if la.use_bias:
                    assert torch.equal(la.bias_pad, lb.bias_pad), (step, name, "bias")


@pytest.mark.parametrize("shape", [(3, 64, 96), (2, 256, 256)])
def test_space_to_depth_stem_matches_the_strided_convolution(shape, monkeypatch):
    """(opt-in experiment, FPG_S2D_STEM=1) The PatchGAN stem (4x4 stride 2 pad 1 over 16 channels) as a 2x2 stride-1 convolution over the space-to-depth
    copy of its input: the copy is exact, the operand is a tap permutation of the ordinary one, and the outputs of both
    formulations agree with each other (fp32 accumulation order differs: a few bf16 roundings) and with torch."""
    from fpgan import ops
    from models import model_architectures as A
    n, h, w = shape
    monkeypatch.setenv("FPG_S2D_STEM", "1")
    torch.manual_seed(3)
    D = A.PairedAttentionDiscriminator(9).cuda()
    ex = D._executor()
    ex.repack()
    layer = ex.layers["model.0"]
    assert layer.s2d_spec is not None
    g = torch.Generator(device="cuda").manual_seed(4)
    x = bf16r(torch.randn(n, 12, h, w, device="cuda", generator=g))
    din = ops.ActBuf(n, h, w, 16, zero=False)
    ops.pack_nchw(x, din, 0, zero_rest=True)
    y_s2d, xs = ex._stem_conv(din, "model.0", ops.ACT_LEAKY)
    assert xs is not None and tuple(xs.t.shape) == (n, h // 2 + 1, w // 2 + 1, 64)
    padded = F.pad(din.t, (0, 0, 1, 1, 1, 1))  # zero ring: pixel (y, x) sits at (y + 1, x + 1)
    blocks = padded.reshape(n, h // 2 + 1, 2, w // 2 + 1, 2, 16).permute(0, 1, 3, 2, 4, 5).reshape(xs.t.shape)
    assert torch.equal(xs.t, blocks)
    wf = layer.spec.w_fprop.view(64, 4, 4, 16)            # [k][r][s][c]
    w2 = layer.s2d_spec.w_fprop.view(64, 2, 2, 2, 2, 16)  # [k][ty][tx][i][j][c]
    assert torch.equal(w2, wf.view(64, 2, 2, 2, 2, 16).permute(0, 1, 3, 2, 4, 5))  # r = 2 ty + i, s = 2 tx + j
    y_ref = ex._conv(din, "model.0", act=ops.ACT_LEAKY)
    close_rms(y_s2d.t, y_ref.t, 2e-2, 1e-3, "space-to-depth vs strided kernel")
    ref = F.leaky_relu(F.conv2d(x, bf16r(D.model[0].weight.detach()), D.model[0].bias.detach(), stride=2, padding=1), 0.2)
    close_rms(y_s2d.to_nchw(64), ref, 2e-2, 2e-3, "space-to-depth stem vs torch")


def test_pack_unpack_bias_grad():
    from fpgan import ops
    g = torch.Generator(device="cuda").manual_seed(19)
    x = torch.rand(2, 9, 32, 40, device="cuda", generator=g) * 2 - 1
    y = torch.rand(2, 3, 32, 40, device="cuda", generator=g) * 2 - 1
    b = ops.ActBuf(2, 32, 40, 16, halo=3)
    b.t.fill_(5.0)
    ops.pack_nchw(x, b, 0, zero_rest=True)
    ops.pack_nchw(y, b, 9)
    ref = F.pad(torch.cat([x, y], 1), (3,) * 4, "reflect")
    torch.testing.assert_close(b.t[..., :12].permute(0, 3, 1, 2).float(), bf16r(ref))
    assert (b.t[..., 12:] == 0).all()
    out = torch.full((2, 3, 32, 40), 1.0, device="cuda")
    ops.unpack_nchw(b, out, c0=9, accumulate=True)
    torch.testing.assert_close(out, bf16r(y) + 1.0)
    dy = ops.ActBuf.from_nchw(bf16r(torch.randn(3, 27, 20, 20, device="cuda", generator=g)), c_pad=32)
    db = torch.zeros(27, device="cuda")
    ops.bias_grad(dy, db, 27)
    torch.testing.assert_close(db, dy.to_nchw(27).sum((0, 2, 3)), rtol=1e-4, atol=1e-3)


def test_pack_paired_inputs_matches_the_separate_packs():
    """the one-pass packer of a paired batch (generator input with reflect halo + both discriminator inputs) against
    torch and against the three pack_nchw calls it replaces (bit-exact)"""
    from fpgan import ops
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.rand(3, 9, 24, 40, device="cuda", generator=g) * 2 - 1
    y = torch.rand(3, 3, 24, 40, device="cuda", generator=g) * 2 - 1
    gin = ops.ActBuf(3, 24, 40, 16, halo=3, zero=False)
    din = ops.ActBuf(6, 24, 40, 16, zero=False)
    for b in (gin, din):
        b.t.fill_(7.0)
    fake, real = din.batch_slice(0, 3), din.batch_slice(3, 3)
    ops.pack_paired_inputs(x, y, gin, fake, real)
    ref = F.pad(x, (3,) * 4, "reflect")
    assert torch.equal(gin.t[..., :9].permute(0, 3, 1, 2).float(), bf16r(ref)) and (gin.t[..., 9:] == 0).all()
    assert torch.equal(fake.t[..., :9].permute(0, 3, 1, 2).float(), bf16r(x)) and (fake.t[..., 9:] == 0).all()
    assert torch.equal(real.t[..., :12].permute(0, 3, 1, 2).float(), bf16r(torch.cat([x, y], 1)))
    assert (real.t[..., 12:] == 0).all()
    old = ops.ActBuf(3, 24, 40, 16, zero=False)
    ops.pack_nchw(x, old, 0, zero_rest=True)
    ops.pack_nchw(y, old, 9)
    assert torch.equal(old.t, real.t)


def test_flood_mask_bit_exact_and_confusion():
    """(sigmoid(x) > 0.5).float() -- model.py:399-400 -- bit-exact against the fp32 CPU expression, including the
    interval 0 < x < ~9e-8 where the fp32 sigmoid rounds to exactly 0.5."""
    from fpgan import ops
    g = torch.Generator().manual_seed(23)
    x = torch.randn(1 << 20, generator=g) * 3
    # every float in a neighbourhood of the rounding threshold, both signs, plus zeros / denormals / infinities
    near = torch.arange(0x33000000, 0x34400000, 37, dtype=torch.int32).view(torch.float32)
    special = torch.tensor([0.0, -0.0, 1e-45, -1e-45, 5.9e-8, 6e-8, 8.9e-8, 9e-8, 1.2e-7, float("inf"),
                            -float("inf"), 88.0, -88.0, 104.0, -104.0])
    x = torch.cat([x, near, -near, special])
    ref = (torch.sigmoid(x) > 0.5).float()
    xd = x.cuda()
    mask = torch.empty_like(xd)
    ops.flood_mask(xd, mask)
    assert torch.equal(mask.cpu(), ref)
    truth = (torch.rand(x.shape, generator=g) > 0.5).float()
    counts = torch.zeros(4, dtype=torch.int64, device="cuda")
    ops.confusion_counts(mask, truth.cuda(), counts)
    tp = int(((ref == 1) & (truth == 1)).sum())
    fp = int(((ref == 1) & (truth == 0)).sum())
    tn = int(((ref == 0) & (truth == 0)).sum())
    fn = int(((ref == 0) & (truth == 1)).sum())
    assert counts.tolist() == [tp, fp, tn, fn]


@pytest.mark.parametrize("shape,use_mask,two", [((2, 16, 16, 64), False, True), ((1, 2, 2, 512), True, False),
                                               ((3, 8, 8, 128), True, True), ((2, 31, 31, 256), False, False)])
def test_batchnorm_fwd_bwd(shape, use_mask, two):
    """BatchNorm2d (batch statistics, affine, running statistics) + dropout mask + up to two activated outputs
    (lrelu for the next down-convolution, relu for the skip connection) against torch autograd."""
    from fpgan import ops
    n, h, w, c = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    y = bf16r(torch.randn(n, c, h, w, device="cuda", generator=g) * 2 + 0.3).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(c, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(c, device="cuda", generator=g)).requires_grad_(True)
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    rm_ref, rv_ref = rm.clone(), rv.clone()
    v = F.batch_norm(y, rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    keep = (torch.rand(n, c, h, w, device="cuda", generator=g) < 0.5) if use_mask else None
    if use_mask:
        v = v * keep * 2.0
    z1, z2 = F.leaky_relu(v, 0.2), F.relu(v)
    dz1, dz2 = bf16r(torch.randn_like(z1)), bf16r(torch.randn_like(z2))
    outs, gouts = ([z1, z2], [dz1, dz2]) if two else ([z1], [dz1])
    dy_ref, dg_ref, db_ref = torch.autograd.grad(outs, (y, gamma, beta), gouts)

    yb = ops.ActBuf.from_nchw(y.detach())
    stats = torch.empty(c * 2, device="cuda")
    ops.batch_stats(yb, stats)
    ops.batchnorm_running_update(stats, n * h * w, rm, rv)
    torch.testing.assert_close(rm, rm_ref, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(rv, rv_ref, rtol=1e-3, atol=1e-4)
    mask = keep.permute(0, 2, 3, 1).contiguous().to(torch.uint8).reshape(-1) if use_mask else None
    wide = ops.ActBuf(n, h, w, 2 * c)  # the second output lands in a channel slice of a wider buffer
    z1b = ops.ActBuf(n, h, w, c)
    ops.batchnorm_apply(yb, stats, gamma.detach(), beta.detach(), ops.ACT_LEAKY, z1b, ops.ACT_RELU,
                        wide.channels(c, c) if two else None, mask=mask)
    close_rms(z1b.to_nchw(), z1.detach(), 0.03, 0.004, "bn apply lrelu")
    if two:
        close_rms(wide.t[..., c:].permute(0, 3, 1, 2).float(), z2.detach(), 0.03, 0.004, "bn apply relu slice")
        assert (wide.t[..., :c] == 0).all()
    dyb = ops.ActBuf(n, h, w, c)
    dg, db = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    gw = ops.ActBuf(n, h, w, 2 * c)
    gw.t[..., :c] = dz2.permute(0, 2, 3, 1)
    ops.batchnorm_bwd(ops.ActBuf.from_nchw(dz1), ops.ACT_LEAKY, yb, stats, gamma.detach(), beta.detach(), dyb, dg, db,
                      dz2=gw.channels(0, c) if two else None, act2=ops.ACT_RELU, mask=mask)
    close_rms(dyb.to_nchw(), dy_ref, 0.05, 0.006, "bn bwd dy")
    close_rms(dg, dg_ref, 0.02, 0.005, "bn bwd dgamma")
    close_rms(db, db_ref, 0.02, 0.005, "bn bwd dbeta")


def test_dropout_mask_is_reproducible_and_fair():
    from fpgan import ops
    m1 = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    m2 = torch.empty_like(m1)
    ops.dropout_mask(m1, 1234)
    ops.dropout_mask(m2, 1234)
    assert (m1 == m2).all() and set(m1.unique().tolist()) == {0, 1}
    assert abs(m1.float().mean().item() - 0.5) < 5e-3
    ops.dropout_mask(m2, 1235)
    assert (m1 != m2).float().mean().item() > 0.4


@pytest.mark.parametrize("case", [
    # (n, h, w, c, k, r, stride, pad, halo, transposed)
    (16, 64, 64, 256, 256, 3, 1, 0, 1, False),   # residual conv on the 2-CTA kernel
    (2, 64, 64, 64, 128, 3, 2, 1, 0, False),     # stride-2 down conv, 1-CTA kernel
    (3, 31, 31, 256, 512, 4, 1, 1, 0, False),    # PatchGAN model.8: ragged tiles, two n-blocks
    (2, 32, 32, 256, 128, 3, 2, 1, 0, True),     # transposed conv: four parity-class launches share the partials
])
@pytest.mark.parametrize("f16", [False, True])
def test_conv_with_epilogue_statistics(case, f16):
    """{mean, rstd} per (image, channel) from the conv epilogue (no separate pass) vs torch on the stored output;
    also in batch mode (BatchNorm statistics). f16: the output is an fp16 buffer (pre-norm tensors of the InstanceNorm
    networks) -- the stored values must be the fp16 rounding of what the bf16 launch rounds to bf16."""
    from fpgan import ops
    n, h, w, c, k, r, stride, pad, halo, transposed = case
    g = torch.Generator(device="cuda").manual_seed(21)
    if transposed:
        x = bf16r(torch.randn(n, c, h, w, device="cuda", generator=g))
        wt = bf16r(torch.randn(c, k, r, r, device="cuda", generator=g) / (c * r * r) ** 0.5 * 2)
        spec = ops.ConvSpec(r, r, stride, pad, ops.pad16(k), ops.pad16(c), c_in_valid=k, c_out_valid=c)
        spec.pack(wt.contiguous())
        xb = ops.ActBuf.from_nchw(x)
        yb = ops.ActBuf(n, 2 * h, 2 * w, ops.pad16(k), f16=f16)
    else:
        x = bf16r(torch.randn(n, c, h, w, device="cuda", generator=g) + 0.2)
        wt = bf16r(torch.randn(k, c, r, r, device="cuda", generator=g) / (c * r * r) ** 0.5)
        spec = ops.ConvSpec(r, r, stride, pad, ops.pad16(c), ops.pad16(k), c_in_valid=c, c_out_valid=k)
        spec.pack(wt.contiguous())
        xb = ops.ActBuf.from_nchw(x, halo=halo)
        hp, wp = h + 2 * halo, w + 2 * halo
        yb = ops.ActBuf(n, (hp + 2 * pad - r) // stride + 1, (wp + 2 * pad - r) // stride + 1, ops.pad16(k), f16=f16)
    for batch in (False, True):
        stats = torch.zeros((1 if batch else n) * yb.c * 2, device="cuda")
        assert ops.conv_with_stats(xb, spec, yb, stats, transposed=transposed, batch=batch)
        y = yb.t.float()  # the stored (bf16-rounded) output [n, h, w, c]
        dims = (0, 1, 2) if batch else (1, 2)
        mean, var = y.mean(dims), y.var(dims, unbiased=False)
        st = stats.view(-1, yb.c, 2)
        torch.testing.assert_close(st[..., 0].reshape(mean.shape), mean, rtol=1e-3, atol=2e-4)
        torch.testing.assert_close(st[..., 1].reshape(var.shape), (var + 1e-5).rsqrt(), rtol=1e-3, atol=1e-3)
    ref = ops.ActBuf(yb.n, yb.h, yb.w, yb.c, fp32=f16)
    if transposed:
        ops.conv_dgrad(xb, spec, ref)
    else:
        ops.conv_fprop(xb, spec, ref)
    if f16:  # the fp32 launch shows the accumulators: the fp16 store is their round-to-nearest
        assert torch.equal(ref.t.half(), yb.t), "fp16 output != fp16 rounding of the fp32 accumulators"
    else:
        assert torch.equal(ref.t, yb.t), "the statistics epilogue must not change the convolution output"
