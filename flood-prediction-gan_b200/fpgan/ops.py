"""Python-side operator wrappers over the C ABI: device buffers are torch tensors (plumbing only), every
operator is one call into libfpg_b200.so on the current CUDA stream. No operator here has a torch fallback."""
import contextlib
import ctypes as C
import os

import torch

from . import lib as L

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH = L.ACT_NONE, L.ACT_RELU, L.ACT_LEAKY, L.ACT_TANH


LAUNCHES = 0     # kernels launched through this module (bench.py reports it as gpu_launches)
PROFILE = None   # when a dict: key -> [(start_event, end_event), ...] around every operator call


def _run(key, n_kernels, name, *args):
    global LAUNCHES
    LAUNCHES += n_kernels
    if PROFILE is None:
        L.call(name, *args)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.call(name, *args)
    b.record()
    PROFILE.setdefault(key, []).append((a, b))


def _conv_key(kind, n, h, w, spec):
    g = spec.g
    return f"{kind} n{n} {h}x{w} c{g.c_in} k{g.c_out} r{g.r} s{g.stride}"


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def pad16(c):
    return (c + 15) // 16 * 16


class ActBuf:
    """An NHWC activation buffer [n, h+2*halo, w+2*halo, c_stride] (bf16 or fp32) plus its fpg_act descriptor.

    `c` channels starting at channel offset `c0` of the underlying storage are visible to the kernels."""

    def __init__(self, n, h, w, c, halo=0, fp32=False, device="cuda", tensor=None, c0=0, c_stride=None, zero=True,
                 f16=False):
        """fp32: network heads; f16: tensors that only elementwise kernels read (pre-normalisation conv outputs, the
        residual skip stream) -- fp16 keeps 3 more mantissa bits than bf16 at the same size; default bf16."""
        assert not (fp32 and f16)
        self.n, self.h, self.w, self.c, self.halo, self.fp32, self.f16 = n, h, w, c, halo, fp32, f16
        self.c_stride = c_stride or c
        self.c0 = c0
        dtype = torch.float32 if fp32 else (torch.float16 if f16 else torch.bfloat16)
        if tensor is None:
            shape = (n, h + 2 * halo, w + 2 * halo, self.c_stride)
            tensor = torch.zeros(shape, dtype=dtype, device=device) if zero else torch.empty(shape, dtype=dtype,
                                                                                             device=device)
        self.t = tensor
        self.desc = L.Act()
        self.desc.data = tensor.data_ptr() + c0 * tensor.element_size()
        self.desc.n, self.desc.h, self.desc.w, self.desc.c = n, h, w, c
        self.desc.c_stride = self.c_stride
        self.desc.halo = halo
        self.desc.fp32 = L.DT_FP32 if fp32 else (L.DT_FP16 if f16 else L.DT_BF16)

    def ref(self):
        return C.byref(self.desc)

    def channels(self, c0, c):
        """A view of `c` channels starting at c0 (shares storage)."""
        return ActBuf(self.n, self.h, self.w, c, self.halo, self.fp32, tensor=self.t, c0=self.c0 + c0,
                      c_stride=self.c_stride, f16=self.f16)

    def batch_slice(self, start, n):
        """A view of images [start, start+n) (shares storage)."""
        return ActBuf(n, self.h, self.w, self.c, self.halo, self.fp32, tensor=self.t[start:start + n], c0=self.c0,
                      c_stride=self.c_stride, f16=self.f16)

    def interior(self):
        """torch view [n, h, w, c] of the interior."""
        hl = self.halo
        t = self.t[:, hl:hl + self.h, hl:hl + self.w] if hl else self.t
        return t[..., self.c0:self.c0 + self.c]

    def to_nchw(self, c=None):
        return self.interior()[..., :c].permute(0, 3, 1, 2).float().contiguous()

    @staticmethod
    def from_nchw(x, c_pad=None, halo=0, mode="reflect", fp32=False, f16=False):
        """Test helper: NCHW float tensor -> ActBuf (torch ops; the product path uses pack_nchw)."""
        n, c, h, w = x.shape
        cp = c_pad or pad16(c)
        buf = ActBuf(n, h, w, cp, halo=halo, fp32=fp32, device=x.device, f16=f16)
        xp = torch.nn.functional.pad(x, (halo,) * 4, mode) if halo else x
        buf.t[..., :c] = xp.permute(0, 2, 3, 1).to(buf.t.dtype)
        return buf


class ConvSpec:
    """Geometry + packed bf16 operands of one convolution layer (forward conv view)."""

    def __init__(self, r, s, stride, pad, c_in, c_out, c_in_valid=None, c_out_valid=None):
        self.g = L.ConvGeom()
        self.g.r, self.g.s, self.g.stride, self.g.pad = r, s, stride, pad
        self.g.c_in, self.g.c_out = c_in, c_out
        self.c_in_valid = c_in_valid if c_in_valid is not None else c_in
        self.c_out_valid = c_out_valid if c_out_valid is not None else c_out
        self.w_fprop = None
        self.w_dgrad = None

    def gref(self):
        return C.byref(self.g)

    def stride_k(self, transposed=False):
        """element strides of the fp32 parameter for (output channel k, input channel c) of the forward-conv view"""
        rs = self.g.r * self.g.s
        return self.c_in_valid * rs  # [K][C][R][S]; for ConvTranspose2d the parameter already is [K=Cin_T][C=Cout_T]

    def alloc(self, device):
        """Allocate the packed bf16 operands (fprop and dgrad layouts)."""
        lib = L.load()
        if self.w_fprop is None:
            nbytes = lib.fpg_packed_weight_bytes(self.gref())
            assert nbytes > 0
            self.w_fprop = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=device)
        if self.w_dgrad is None:
            nbytes = lib.fpg_packed_weight_dgrad_bytes(self.gref())
            assert nbytes > 0
            self.w_dgrad = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=device)

    def pack(self, weight, fprop=True, dgrad=True):
        """(Re)pack the fp32 parameter `weight` ([K][C][R][S] in the forward-conv view) into the bf16 operands."""
        lib = L.load()
        rs = self.g.r * self.g.s
        sk, sc = self.c_in_valid * rs, rs
        assert weight.is_contiguous() and weight.dtype == torch.float32
        assert weight.numel() == self.c_out_valid * self.c_in_valid * rs, (weight.shape, self.c_out_valid,
                                                                           self.c_in_valid, rs)
        if fprop:
            if self.w_fprop is None:
                nbytes = lib.fpg_packed_weight_bytes(self.gref())
                assert nbytes > 0
                self.w_fprop = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=weight.device)
            _run("pack_weights", 1, "fpg_pack_weights", _ptr(weight), sk, sc, self.c_out_valid, self.c_in_valid, self.gref(),
                   _ptr(self.w_fprop), _stream())
        if dgrad:
            if self.w_dgrad is None:
                nbytes = lib.fpg_packed_weight_dgrad_bytes(self.gref())
                assert nbytes > 0
                self.w_dgrad = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=weight.device)
            _run("pack_weights", 4 if self.g.stride == 2 else 1, "fpg_pack_weights_dgrad", _ptr(weight), sk, sc, self.c_out_valid, self.c_in_valid, self.gref(),
                   _ptr(self.w_dgrad), _stream())


def pack_weights_batched(jobs, block_job, block_first, n_blocks):
    _run("pack_weights", 1, "fpg_pack_weights_batched", _ptr(jobs), _ptr(block_job), _ptr(block_first), n_blocks,
         _stream())


def conv_fprop(x, spec, y, bias=None, act=ACT_NONE):
    _run(_conv_key("fprop", y.n, y.h, y.w, spec), 1, "fpg_conv2d_fprop", x.ref(), _ptr(spec.w_fprop), _ptr(bias), act,
         spec.gref(), y.ref(), _stream())


_dgrad_launches = {}


def dgrad_launches(dy, spec, dx):
    """kernels one conv_dgrad launches for this geometry (the parity classes of a stride-2 layer run as one launch
    where their plans qualify)"""
    g = spec.g
    key = (dy.n, dy.h, dy.w, dx.h, dx.w, dx.halo, g.r, g.s, g.stride, g.pad, g.c_in, g.c_out)
    if key not in _dgrad_launches:
        n = L.load().fpg_conv2d_dgrad_launches(dy.ref(), spec.gref(), dx.ref()) if g.stride == 2 else 1
        _dgrad_launches[key] = max(1, n)
    return _dgrad_launches[key]


def conv_dgrad(dy, spec, dx, bias=None, act=ACT_NONE):
    _run(_conv_key("dgrad", dx.n, dx.h, dx.w, spec), dgrad_launches(dy, spec, dx), "fpg_conv2d_dgrad", dy.ref(),
         _ptr(spec.w_dgrad), _ptr(bias), act, spec.gref(), dx.ref(), _stream())


_stat_ws = {}


def _stat_workspace(nfloats, device):
    """partials of the epilogue statistics (own buffer: the generic workspace is in use by other kernels)"""
    key = str(device)
    cur = _stat_ws.get(key)
    if cur is None or cur.numel() < nfloats:
        cur = torch.empty(max(nfloats, 1 << 22), dtype=torch.float32, device=device)
        _stat_ws[key] = cur
    return cur


def conv_with_stats(x, spec, y, stats, transposed=False, eps=1e-5, batch=False):
    """y = conv(x) (or the transposed-conv forward) and stats = per-(image, channel) {mean, rstd} of y computed from
    the conv epilogue; `batch`: statistics over the whole batch (BatchNorm). Returns False if this layer's kernel has no
    statistics epilogue (the caller then runs instnorm_stats)."""
    lib = L.load()
    rows = lib.fpg_conv_stats_rows(x.ref(), spec.gref(), y.ref(), 1 if transposed else 0)
    if rows <= 0:
        return False
    ws = _stat_workspace(int(1.02 * y.n * rows * y.c * 2) + 4096, y.t.device)
    if transposed:
        _run(_conv_key("dgrad", y.n, y.h, y.w, spec), dgrad_launches(x, spec, y), "fpg_conv2d_dgrad_stats", x.ref(),
             _ptr(spec.w_dgrad), None, ACT_NONE, spec.gref(), y.ref(), _ptr(ws), _stream())
    else:
        _run(_conv_key("fprop", y.n, y.h, y.w, spec), 1, "fpg_conv2d_fprop_stats", x.ref(), _ptr(spec.w_fprop), None,
             ACT_NONE, spec.gref(), y.ref(), _ptr(ws), _stream())
    n, per = (1, y.n * rows) if batch else (y.n, rows)
    folds, left = 0, per  # row lists longer than 512 are first folded 256:1 (one launch per round)
    while left > 512:
        left, folds = (left + 255) // 256, folds + 1
    _run("instnorm_stats", 1 + folds, "fpg_instnorm_stats_finalize", _ptr(ws), per, n, y.c,
         (y.n if batch else 1) * y.h * y.w, eps, _ptr(stats), _stream())
    return True


_ws_cache = {}


def workspace(nbytes, device):
    """Grow-only fp32 scratch shared by the wgrad / reduction kernels of one device and stream (stream-ordered reuse)."""
    key = (str(device), torch.cuda.current_stream().cuda_stream)
    cur = _ws_cache.get(key)
    if cur is None or cur.numel() * 4 < nbytes:
        cur = torch.empty(max(nbytes // 4 + 1, 1 << 20), dtype=torch.float32, device=device)
        _ws_cache[key] = cur
    return cur


class _WgradSide:
    """The weight-gradient kernels of a backward pass run on a second stream: they depend only on the layer's saved
    input and its output gradient, not on the dgrad -> InstanceNorm-backward chain, so their CTAs fill the SMs that the
    chain's kernels leave idle (the 64x64 convs run 3.46 waves of tiles: 80 SMs idle through the fourth).
    `with wgrad_side(x, dy):` switches to the side stream after making it wait for the producing stream and keeps the
    operands alive until `wgrad_join()`, which every executor's backward() calls before it returns.
    Opt-in (FPG_WGRAD_STREAM=1): measured on the B=16 PairedAttention step it is no gain (10.04 ms against 9.96 ms on
    one stream) -- every kernel here is a one-CTA-per-SM persistent grid, so two streams only trade places."""

    def __init__(self):
        self.streams, self.keep, self.dirty = {}, [], False
        self.enabled = os.environ.get("FPG_WGRAD_STREAM", "0") == "1"

    def stream(self, device):
        st = self.streams.get(str(device))
        if st is None:
            st = self.streams[str(device)] = torch.cuda.Stream(device=device)
        return st


_SIDE = _WgradSide()


@contextlib.contextmanager
def wgrad_side(*operands):
    if not _SIDE.enabled or PROFILE is not None:
        yield
        return
    main = torch.cuda.current_stream()
    side = _SIDE.stream(operands[0].t.device)
    side.wait_stream(main)
    _SIDE.keep.append(operands)
    _SIDE.dirty = True
    with torch.cuda.stream(side):
        yield


def wgrad_join():
    """the current stream waits for the weight-gradient stream; the operands it kept alive are released"""
    if not _SIDE.dirty:
        return
    main = torch.cuda.current_stream()
    main.wait_stream(_SIDE.stream(main.device))
    _SIDE.keep.clear()
    _SIDE.dirty = False


def conv_wgrad(x, dy, spec, dw):
    """dw (fp32 parameter-layout gradient, contiguous [K][C][R][S]) = conv_backward_weight(x, dy); overwritten."""
    lib = L.load()
    sms = lib.fpg_sm_count()
    nbytes = lib.fpg_conv2d_wgrad_ws_bytes(x.ref(), dy.ref(), spec.gref(), sms)
    if nbytes <= 0:
        L.check(-22, "fpg_conv2d_wgrad_ws_bytes")
    ws = workspace(nbytes, dw.device)
    rs = spec.g.r * spec.g.s
    _run(_conv_key("wgrad", dy.n, dy.h, dy.w, spec), 2, "fpg_conv2d_wgrad", x.ref(), dy.ref(), spec.gref(), _ptr(dw),
         spec.c_in_valid * rs, rs, spec.c_out_valid, spec.c_in_valid, _ptr(ws), _stream())


def bias_grad(dy, db, k_valid):
    ws = workspace(592 * dy.c * 4, db.device)
    _run("bias_grad", 2, "fpg_bias_grad", dy.ref(), _ptr(db), k_valid, _ptr(ws), _stream())


def _scratch_for(y):
    n = L.load().fpg_instnorm_scratch_floats(y.ref())
    return workspace(n * 4, y.t.device)


_counter_cache = {}


def _counters(device):
    """zero-initialised int32 ticket counters (kernels leave them zero)"""
    key = str(device)
    if key not in _counter_cache:
        _counter_cache[key] = torch.zeros(4096 + 64, dtype=torch.int32, device=device)
    return _counter_cache[key]


def instnorm_stats(y, stats, eps=1e-5):
    _run("instnorm_stats", 1, "fpg_instnorm_stats", y.ref(), eps, _ptr(stats), _ptr(_scratch_for(y)),
         _ptr(_counters(y.t.device)), _stream())


def instnorm_apply(y, stats, act, z, residual=None, skip_out=None):
    """z = act(IN(y)) (+ residual) with its reflect halo; skip_out (halo-free fp16 / bf16) also receives the values
    before they are rounded to z's bf16 (the trunk's skip stream)"""
    _run("instnorm_apply", 1, "fpg_instnorm_apply", y.ref(), _ptr(stats), act, residual.ref() if residual is not None else None,
           z.ref(), skip_out.ref() if skip_out is not None else None, _stream())


def instnorm_bwd(dz, y, stats, act, dy, dz2=None, dres=None):
    _run("instnorm_bwd", 3 if dz.halo else 2, "fpg_instnorm_bwd", dz.ref(), dz2.ref() if dz2 is not None else None, y.ref(), _ptr(stats),
         act, dy.ref(), dres.ref() if dres is not None else None, _ptr(_scratch_for(y)), _ptr(_counters(y.t.device)),
         _stream())


# The reduction pass of the InstanceNorm backward runs in the epilogue of the data-gradient kernel that produces its
# upstream gradient (FPG_INBWD_EPILOGUE=0 restores the separate streamed pass for A/B measurements).
# Values: "0" off, "relu" only where the staged operand is the convolution's own input (relu-type norms: one tensor),
# "1" everywhere in the residual trunk (residual-type norms stage three tensors).
INBWD_EPILOGUE = os.environ.get("FPG_INBWD_EPILOGUE", "relu")


def conv_dgrad_inbwd(dy, spec, dx, z, zprev=None, add=None, force=False):
    """dx (incl. halo) = conv_backward_data(dy, w) [+ add on the interior], the gradient w.r.t. the convolution's
    saved input z = relu(IN(y)) (zprev None) or z = zprev + IN(y) (residual block output); the reduction pass of that
    InstanceNorm's backward runs in the conv epilogue, fed by z (and zprev) staged through shared memory. Returns
    red [n, c, 2] = {mean g', mean g' * zhat} for instnorm_bwd_apply, or None when this layer has no such epilogue
    (nothing was launched: the caller runs conv_dgrad + instnorm_bwd)."""
    lib = L.load()
    enabled = force or INBWD_EPILOGUE == "1" or (INBWD_EPILOGUE == "relu" and zprev is None and add is None)
    if not enabled or spec.g.stride != 1 or z.halo != dx.halo or dx.c % 64:
        return None
    rows = lib.fpg_conv_stats_rows(dy.ref(), spec.gref(), dx.ref(), 1)
    if rows <= 0:
        return None
    ws = _stat_workspace(int(1.02 * dx.n * rows * dx.c * 2) + 4096, dx.t.device)
    rows_out = C.c_int32(0)
    ev = None
    if PROFILE is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    rc = lib.fpg_conv2d_dgrad_inbwd(dy.ref(), _ptr(spec.w_dgrad), spec.gref(), dx.ref(), z.ref(),
                                    zprev.ref() if zprev is not None else None,
                                    add.ref() if add is not None else None, _ptr(ws), C.byref(rows_out), _stream())
    if rc == 1:
        return None
    L.check(rc, "fpg_conv2d_dgrad_inbwd")
    if ev is not None:
        ev[1].record()
        PROFILE.setdefault(_conv_key("dgrad", dx.n, dx.h, dx.w, spec), []).append(ev)
    global LAUNCHES
    LAUNCHES += 1
    red = torch.empty(dx.n, dx.c, 2, dtype=torch.float32, device=dx.t.device)
    _run("instnorm_bwd", 1, "fpg_instnorm_bwd_sums_finalize", _ptr(ws), rows_out.value, dx.n, dx.c, dx.h * dx.w,
         _ptr(red), _stream())
    return red


def instnorm_bwd_apply(dz, y, stats, red, act, dy):
    """the apply half of instnorm_bwd with the plane means `red` given (conv_dgrad_inbwd); dz is consumed (folded)"""
    _run("instnorm_bwd", 2 if dz.halo else 1, "fpg_instnorm_bwd_apply", dz.ref(), y.ref(), _ptr(stats), _ptr(red), act, dy.ref(),
         _stream())


def batch_stats(y, stats, eps=1e-5):
    """Per-channel {mean, rstd} over the whole batch: InstanceNorm statistics of the batch viewed as one image."""
    flat = ActBuf(1, y.n * y.h, y.w, y.c, tensor=y.t, c0=y.c0, c_stride=y.c_stride, f16=y.f16)
    instnorm_stats(flat, stats, eps)


def batchnorm_apply(y, stats, gamma, beta, act1, z1, act2=ACT_NONE, z2=None, mask=None, mask_scale=2.0):
    _run("batchnorm_apply", 1, "fpg_batchnorm_apply", y.ref(), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(mask),
         float(mask_scale), act1, z1.ref(), act2, z2.ref() if z2 is not None else None, _stream())


def batchnorm_bwd(dz1, act1, y, stats, gamma, beta, dy, dgamma=None, dbeta=None, dz2=None, act2=ACT_NONE, mask=None,
                  mask_scale=2.0, accumulate=False):
    n = L.load().fpg_batchnorm_scratch_floats(y.ref())
    ws = workspace(n * 4, y.t.device)
    _run("batchnorm_bwd", 3, "fpg_batchnorm_bwd", dz1.ref(), act1, dz2.ref() if dz2 is not None else None, act2,
         _ptr(mask), float(mask_scale), y.ref(), _ptr(stats), _ptr(gamma), _ptr(beta), dy.ref(), _ptr(dgamma),
         _ptr(dbeta), 1 if accumulate else 0, _ptr(ws), _stream())


def batchnorm_running_update(stats, count, running_mean, running_var, eps=1e-5, momentum=0.1):
    _run("batchnorm_running", 1, "fpg_batchnorm_running_update", _ptr(stats), running_mean.numel(), int(count),
         float(eps), float(momentum), _ptr(running_mean), _ptr(running_var), _stream())


def maxpool2(x, y):
    _run("maxpool2", 1, "fpg_maxpool2", x.ref(), y.ref(), _stream())


def dropout_mask(mask, seed, keep=0.5):
    _run("dropout_mask", 1, "fpg_dropout_mask", _ptr(mask), mask.numel(), C.c_uint64(int(seed) & (2 ** 64 - 1)),
         float(keep), _stream())


def dropout_mask_dev(mask, seed_dev, seed_add=0, keep=0.5):
    """seed_dev: one int64 element on the device (the step's seed); seed_add is added to it"""
    assert seed_dev.dtype == torch.int64 and seed_dev.numel() == 1
    _run("dropout_mask", 1, "fpg_dropout_mask_dev", _ptr(mask), mask.numel(), _ptr(seed_dev),
         C.c_uint64(int(seed_add) & (2 ** 64 - 1)), float(keep), _stream())


def act_bwd(dz, z, act, dx):
    _run("act_bwd", 1, "fpg_act_bwd", dz.ref(), z.ref(), act, dx.ref(), _stream())


def halo_fold(a, b, c):
    _run("halo_fold", 1, "fpg_halo_fold", a.ref(), b.ref() if b is not None else None, c.ref(), _stream())


def blend_fwd(content, logits, inp, out=None, out_c0=0, out_nchw=None, mask=None, input_lo_offset=0):
    _run("blend_fwd", 1, "fpg_blend_fwd", content.ref(), logits.ref(), inp.ref(), input_lo_offset,
         out.ref() if out is not None else None, out_c0, _ptr(out_nchw), _ptr(mask), _stream())


def norm_split_f32(y, out, norm=True, act=ACT_NONE, residual=None, skip_out=None, eps=1e-5):
    """fp32 parity mode: out [hi | lo | hi] = act(IN(y) or y) (+ residual), fp32 throughout (see fpg_norm_split_f32)"""
    _run("norm_split_f32", 1, "fpg_norm_split_f32", y.ref(), 1 if norm else 0, float(eps), act, _ptr(residual),
         _ptr(skip_out), out.ref(), _stream())


def pack_nchw_split(src, dst):
    assert src.dtype == torch.float32 and src.is_contiguous()
    _run("pack_nchw_split", 1, "fpg_pack_nchw_split", _ptr(src), src.shape[1], dst.ref(), _stream())


def blend_bwd(content, logits, inp, dcontent, dlogits, dout_nchw=None, dout_nhwc=None, dout_c0=0, dimage_nchw=None):
    _run("blend_bwd", 1, "fpg_blend_bwd", _ptr(dout_nchw), dout_nhwc.ref() if dout_nhwc is not None else None, dout_c0,
           content.ref(), logits.ref(), inp.ref(), dcontent.ref(), dlogits.ref(), _ptr(dimage_nchw), _stream())


def mse_const_loss(logits, target, weight, grad_scale, loss, dlogits=None):
    ws = workspace(4096 * 4, logits.t.device)
    # int32 #4096 of the shared counter block is this kernel's ticket (the InstanceNorm kernels use the first 4096)
    _run("mse_const_loss", 1, "fpg_mse_const_loss", logits.ref(), float(target), float(weight), float(grad_scale), _ptr(loss),
           dlogits.ref() if dlogits is not None else None, _ptr(ws), _ptr(_counters(logits.t.device)[4096:]), _stream())


def l1_loss(pred, target, weight, grad_scale, loss, dpred=None, accumulate=False):
    """loss = weight * mean|pred - target| (+ its gradient). pred: contiguous fp32 [B, c, H, W]; target: contiguous, or
    the leading c channels of a wider contiguous NCHW tensor (real_image[:, :3])."""
    ws = workspace(4096, pred.device)
    per_image, t_stride = 0, 0
    assert pred.is_contiguous() and pred.dtype == torch.float32 and target.dtype == torch.float32
    if not target.is_contiguous():
        per_image, t_stride = pred[0].numel(), target.stride(0)
        assert target.shape == pred.shape and target[0].is_contiguous(), "target must be a leading-channel slice"
    _run("l1_loss", 2, "fpg_l1_loss", _ptr(pred), _ptr(target), pred.numel(), per_image, t_stride, float(weight),
         float(grad_scale), _ptr(loss), _ptr(dpred), 1 if accumulate else 0, _ptr(ws), _stream())


def pack_nchw(src, dst, c0=0, zero_rest=False):
    """src: fp32 [B, c, H, W], contiguous or a channel slice x[:, a:b] of a contiguous NCHW tensor"""
    assert src.dtype == torch.float32
    c_img = 0
    if not src.is_contiguous():
        assert src[0].is_contiguous() and src.stride(0) % (src.shape[2] * src.shape[3]) == 0, "not a channel slice"
        c_img = src.stride(0) // (src.shape[2] * src.shape[3])
    _run("pack_nchw", 1, "fpg_pack_nchw", _ptr(src), src.shape[1], c_img, dst.ref(), c0, 1 if zero_rest else 0,
         _stream())


def pack_paired_inputs(x, y, gin, fake, real):
    """x [B, cx, H, W], y [B, cy, H, W] contiguous fp32 -> the generator input (reflect halo) and both discriminator
    inputs, one pass (fpg_pack_paired_inputs)"""
    assert x.dtype == y.dtype == torch.float32 and x.is_contiguous() and y.is_contiguous()
    _run("pack_nchw", 1, "fpg_pack_paired_inputs", _ptr(x), x.shape[1], _ptr(y), y.shape[1], gin.ref(), fake.ref(),
         real.ref(), _stream())


def space_to_depth16(src, dst):
    """16-channel bf16 [n, h, w, 16] -> [n, h/2+1, w/2+1, 64] (fpg_space_to_depth16): the PatchGAN stem's input"""
    _run("space_to_depth", 1, "fpg_space_to_depth16", src.ref(), dst.ref(), _stream())


def add_f32(dst, src):
    """dst += src (flat fp32 gradient buffers)"""
    assert dst.dtype == src.dtype == torch.float32 and dst.numel() == src.numel()
    _run("add_f32", 1, "fpg_add_f32", _ptr(dst), _ptr(src), dst.numel(), _stream())


def history_exchange(cur, pool, ctrl, out):
    """device-resident get_buffer_image: see fpg_history_exchange. cur/out: same-size tensors, pool: [50, ...]"""
    nbytes = cur.numel() * cur.element_size()
    assert out.numel() * out.element_size() == nbytes and pool[0].numel() * pool.element_size() == nbytes
    _run("history_exchange", 1, "fpg_history_exchange", _ptr(cur), _ptr(pool), _ptr(ctrl), _ptr(out), nbytes, _stream())


def unpack_nchw(src, dst, c0=0, accumulate=False):
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    _run("unpack_nchw", 1, "fpg_unpack_nchw", src.ref(), c0, _ptr(dst), dst.shape[1], 1 if accumulate else 0, _stream())


def tanh_bwd_pack(dout_nchw, out, dpre, c_valid=3):
    assert dout_nchw.dtype == torch.float32 and dout_nchw.is_contiguous()
    _run("tanh_bwd_pack", 1, "fpg_tanh_bwd_pack", _ptr(dout_nchw), out.ref(), c_valid, dpre.ref(), _stream())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    _run("adam_step", 1, "fpg_adam_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), float(beta1), float(beta2),
           float(eps), int(step), float(grad_scale), _stream())


def adam_step_dev(p, g, m, v, state, beta1, beta2, eps, grad_scale=1.0):
    """state: int32[4] device tensor {step, lr bits, -, -}"""
    _run("adam_step", 2, "fpg_adam_step_dev", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(beta1), float(beta2),
         float(eps), _ptr(state), float(grad_scale), _stream())


def flood_mask(logits, mask):
    _run("flood_mask", 1, "fpg_flood_mask", _ptr(logits), _ptr(mask), logits.numel(), _stream())


def confusion_counts(pred, truth, counts):
    _run("confusion_counts", 1, "fpg_confusion_counts", _ptr(pred), _ptr(truth), pred.numel(), _ptr(counts), _stream())


# ---------------------------------------------------------------------------------------------- input pipeline
def resize_bicubic_aa(src_hwc, channel_map, out_h, out_w, flip_w=False, out=None):
    """src_hwc: decoded stack [H, W, C] fp32 on the device. Returns [len(channel_map), out_h, out_w] fp32: the selected
    channels, optionally mirrored left-right, resized like torchvision Resize(BICUBIC, antialias=True)."""
    assert src_hwc.is_cuda and src_hwc.dtype == torch.float32 and src_hwc.is_contiguous() and src_hwc.dim() == 3
    h, w, c = src_hwc.shape
    n = len(channel_map)
    if out is None:
        out = torch.empty(n, out_h, out_w, dtype=torch.float32, device=src_hwc.device)
    nbytes = L.load().fpg_resize_aa_scratch_bytes(h, w, out_h, out_w, n)
    if nbytes <= 0:
        L.check(-22, "fpg_resize_aa_scratch_bytes")
    ws = workspace(nbytes + 256, src_hwc.device)
    base = (ws.data_ptr() + 255) & ~255
    cmap = (C.c_int32 * n)(*channel_map)
    _run("resize_bicubic_aa", 3, "fpg_resize_bicubic_aa", _ptr(src_hwc), h, w, c, cmap, n, 1 if flip_w else 0, out_h,
         out_w, _ptr(out), C.c_void_p(base), _stream())
    return out


def tile_gather(image_ptrs, crop_index, channels, height, width, divisions, out, mean=0.5, std=0.5):
    """image_ptrs: int64 device tensor of B pointers to resident [channels, height, width] fp32 images; crop_index:
    int32 device tensor [B]; out: [B, channels, height // divisions, width // divisions] fp32."""
    _run("tile_gather", 1, "fpg_tile_gather", _ptr(image_ptrs), channels, height, width, _ptr(crop_index),
         image_ptrs.numel(), divisions, mean, std, _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------------------------- image-quality metrics
def ssim_stats(pred, target, data_range=1.0, k1=0.01, k2=0.03, sigma=1.5):
    """pred, target: [B, C, H, W] fp32 CUDA. Returns [B, 2] = per-image (mean SSIM, mean contrast sensitivity)."""
    assert pred.shape == target.shape and pred.is_cuda and pred.dtype == torch.float32
    pred, target = pred.contiguous(), target.contiguous()
    b, c, h, w = pred.shape
    nbytes = L.load().fpg_ssim_scratch_bytes(b, c, h, w)
    if nbytes <= 0:
        L.check(-22, "fpg_ssim_scratch_bytes")
    ws = workspace(nbytes, pred.device)
    out = torch.empty(b, 2, dtype=torch.float32, device=pred.device)
    _run("ssim_stats", 2, "fpg_ssim_stats", _ptr(pred), _ptr(target), b, c, h, w, data_range, k1, k2, sigma, _ptr(out),
         _ptr(ws), _stream())
    return out


def avgpool2_f32(x):
    b, c, h, w = x.shape
    y = torch.empty(b, c, h // 2, w // 2, dtype=torch.float32, device=x.device)
    _run("avgpool2_f32", 1, "fpg_avgpool2_f32", _ptr(x.contiguous()), _ptr(y), b * c, h, w, _stream())
    return y


def sq_err_sum(a, b, clamp=(0.0, 1.0)):
    """sum (clamp(a) - clamp(b))^2 as a float64 device scalar (fixed summation order)"""
    a, b = a.contiguous(), b.contiguous()
    ws = workspace(L.load().fpg_sq_err_scratch_bytes(), a.device)
    out = torch.empty(1, dtype=torch.float64, device=a.device)
    _run("sq_err_sum", 2, "fpg_sq_err_sum", _ptr(a), _ptr(b), a.numel(), clamp[0], clamp[1], _ptr(out), _ptr(ws),
         _stream())
    return out
